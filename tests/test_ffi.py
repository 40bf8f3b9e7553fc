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
    assert C.sizeof(_ffi.StepInfo) == 6 * 8 + 6 * 4 + 8
    assert _ffi.RECORD.itemsize == 16
