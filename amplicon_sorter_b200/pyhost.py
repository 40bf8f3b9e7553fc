"""Python-host glue: one C pass over comparelist2's records (csrc/pyhost.c) instead of three Python passes.

Not part of the C ABI (include/asb200.h stays free of Python types) and not required: when the helper library is
missing, or a record is not what read_file builds (amplicon_sorter.py:551-561), `collect` returns None and
host.process_list takes its generic Python path -- same result, a few dozen milliseconds later."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build

_LIB = None


def _load():
    global _LIB
    if _LIB is None:
        if not os.path.exists(_build.PYHOST_LIB):
            _LIB = False
        else:
            L = C.PyDLL(_build.PYHOST_LIB)  # PyDLL: the GIL stays held, the helper calls the CPython API
            L.asbpy_collect_records.argtypes = [C.py_object, C.c_void_p, C.c_void_p, C.c_void_p]
            L.asbpy_collect_records.restype = C.c_int
            _LIB = L
    return _LIB


def collect(records: list):
    """(idx keys int64[n], pointers uint64[n] to the ASCII bytes of every SEQ, lengths uint32[n]) or None.
    The pointers are borrowed from the str objects: `records` must stay alive and unchanged while they are used."""
    L = _load()
    if not L or not isinstance(records, list):
        return None
    n = len(records)
    keys = np.empty(n, dtype=np.int64)
    ptrs = np.empty(n, dtype=np.uint64)
    lens = np.empty(n, dtype=np.uint32)
    if L.asbpy_collect_records(records, keys.ctypes.data, ptrs.ctypes.data, lens.ctypes.data) != 0:
        return None
    return keys, ptrs, lens
