"""CPU: the C-ABI library loads and exports every symbol include/asb200.h declares (no compute)."""
import ctypes as C
import os
import re

from amplicon_sorter_b200 import _ffi, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_builds_and_exports_header_symbols():
    build.build()
    lib = _ffi.load()
    header = open(os.path.join(ROOT, "include", "asb200.h")).read()
    declared = set(re.findall(r"\b(asb_[a-z_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in asb200.h but not exported"
    assert set(_ffi.SYMBOLS) <= declared
    assert lib.asb_version() >= 100


def test_struct_layouts_match_header():
    assert C.sizeof(_ffi.StepInfo) == 6 * 8 + 6 * 4 + 3 * 8 + 8 + 4 * 4
    assert _ffi.RECORD.itemsize == 16


def test_no_cpu_fallback_engine_creation_fails_loudly_without_gpu():
    """On a box without a CUDA device the product must raise, never fall back (there is no CPU path)."""
    import pytest
    import torch

    from amplicon_sorter_b200.engine import Engine
    from amplicon_sorter_b200._ffi import EngineError

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(EngineError):
        Engine(0)


def test_product_package_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under amplicon_sorter_b200/ may import or load it."""
    import glob

    for path in glob.glob(os.path.join(ROOT, "amplicon_sorter_b200", "**", "*.py"), recursive=True):
        src = open(path).read()
        assert "oracle" not in src.replace("# oracle", ""), path
    for path in glob.glob(os.path.join(ROOT, "amplicon_sorter_b200", "csrc", "*")):
        assert "asref" not in open(path).read(), path


def test_format_records_host_helper_matches_python():
    import numpy as np

    from amplicon_sorter_b200 import host

    rng = np.random.default_rng(3)
    lens = np.sort(rng.integers(300, 1100, 500))
    rec = np.empty(2000, dtype=_ffi.RECORD)
    rec["i_pos"] = rng.integers(0, 499, 2000)
    rec["j_pos"] = np.minimum(rec["i_pos"] + rng.integers(1, 50, 2000), 499)
    rec["d"] = rng.integers(0, 200, 2000)
    rec["reverse"] = rng.integers(0, 2, 2000)
    idx = rng.permutation(500)
    text = host.format_records(rec, idx, lens)
    want = "".join(f"{idx[r['i_pos']]}:{idx[r['j_pos']]}:{round(1 - int(r['d']) / int(lens[r['j_pos']]), 3)}" + (":reverse\n" if r["reverse"] else "\n") for r in rec)
    assert text == want
