"""In-tree build of libasb200.so (hand-written sm_100a CUDA + C ABI).  nvcc cross-compiles without a GPU."""
from __future__ import annotations

import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(_HERE, "csrc", "asb200.cu")
DEPS = [SRC, os.path.join(_HERE, "csrc", "myers_band.cuh"), os.path.join(_HERE, "csrc", "lines.cuh"), os.path.join(_HERE, "csrc", "text.cuh"), os.path.join(os.path.dirname(_HERE), "include", "asb200.h")]
LIB = os.path.join(_HERE, "_lib", "libasb200.so")
PYHOST_SRC = os.path.join(_HERE, "csrc", "pyhost.c")
PYHOST_LIB = os.path.join(_HERE, "_lib", "libasb_pyhost.so")  # Python-host glue (records -> pointers), plain C, no CUDA

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-shared",
              "-Xcompiler", "-fPIC", "-cudart", "static"]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libasb200.so cannot be built")


def stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in DEPS)


def build_pyhost(force: bool = False) -> str:
    """The Python host's record walker (csrc/pyhost.c): gcc against this interpreter's headers."""
    import sysconfig

    if not force and os.path.exists(PYHOST_LIB) and os.path.getmtime(PYHOST_LIB) >= os.path.getmtime(PYHOST_SRC):
        return PYHOST_LIB
    os.makedirs(os.path.dirname(PYHOST_LIB), exist_ok=True)
    cc = shutil.which("gcc") or shutil.which("cc")
    if not cc:
        raise RuntimeError("gcc not found: libasb_pyhost.so cannot be built")
    subprocess.check_call([cc, "-O2", "-shared", "-fPIC", "-I" + sysconfig.get_paths()["include"], "-o", PYHOST_LIB, PYHOST_SRC])
    return PYHOST_LIB


def build(force: bool = False, verbose: bool = False) -> str:
    build_pyhost(force)
    if not force and not stale():
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    cmd = [nvcc_path(), *NVCC_FLAGS, "-o", LIB, SRC]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd))
    env = dict(os.environ)
    env.pop("CC", None)
    env.pop("CXX", None)
    subprocess.check_call(cmd, env=env)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
