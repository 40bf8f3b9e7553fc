"""ctypes binding of include/asb200.h.  There is no fallback: a missing library is a hard error."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build

RECORD = np.dtype([("i_pos", "<u4"), ("j_pos", "<u4"), ("d", "<u4"), ("reverse", "<u4")])

ASB_OK, ASB_DONE = 0, 1
TEXT_MAX_LINE = 40  # >= kTextMaxLine of csrc/text.cuh: bytes of the longest possible line
TEXT_CHUNK = 1 << 19  # records per asb_text_step call: bounds the pinned buffers (6 x 32 MB; pinning costs ~0.3 ms per MB,
# on the critical path of a one-shot run) and feeds the writer threads early


class StepInfo(C.Structure):
    _fields_ = [("pairs", C.c_uint64), ("n_records", C.c_uint64), ("fwd_survivors", C.c_uint64),
                ("rc_survivors", C.c_uint64), ("zone_checks", C.c_uint64), ("word_updates", C.c_uint64),
                ("row_begin", C.c_uint32), ("row_end", C.c_uint32), ("screen_ms", C.c_float), ("total_ms", C.c_float),
                ("launches", C.c_uint32), ("reserved", C.c_uint32), ("screen_word_updates", C.c_uint64),
                ("useful_word_updates", C.c_uint64), ("screen_useful_word_updates", C.c_uint64),
                ("pruned_pairs", C.c_uint64), ("cluster_ms", C.c_float), ("n_pivots", C.c_uint32),
                ("lists_ms", C.c_float), ("reserved2", C.c_uint32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class EngineError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"asb200 error {code}: {msg}")
        self.code = code


_LIB = None

SYMBOLS = ["asb_version", "asb_create", "asb_destroy", "asb_last_error", "asb_set_param", "asb_upload_reads", "asb_upload_reads_dev", "asb_upload_reads_scattered", "asb_prepare_pruning", "asb_uploaded_ascii_dev",
           "asb_batch_begin", "asb_batch_step", "asb_batch_records", "asb_distance_pairs", "asb_debug_read", "asb_batch_records_dev", "asb_int_peak", "asb_format_records",
           "asb_kmer_build", "asb_kmer_shared_pairs", "asb_kmer_shared_tile", "asb_threeway_pairs",
           "asb_lines_upload", "asb_lines_hist", "asb_lines_besthit", "asb_lines_besthit_fetch", "asb_components",
           "asb_text_begin", "asb_text_load", "asb_text_step", "asb_text_measure", "asb_lines_append_dev", "asb_host_alloc", "asb_host_free", "asb_lines_count", "asb_lines_fetch"]


def lib_path() -> str:
    return os.environ.get("ASB200_LIB") or _build.LIB  # ASB200_LIB: a differently tuned build of the same library (experiments)


def load():
    """Load libasb200.so (built in-tree by build.py / __graft_entry__.build()).  Raises if absent."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.exists(path):
        raise EngineError(-100, f"{path} is missing: run `python -m amplicon_sorter_b200.build` "
                                "(the CUDA library is the product; there is no CPU fallback)")
    L = C.CDLL(path)
    vp, u8p, u32p, u64p, i32p = C.c_void_p, C.POINTER(C.c_uint8), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64), C.POINTER(C.c_int32)
    L.asb_version.restype = C.c_int
    L.asb_create.argtypes = [C.c_int, vp, C.POINTER(vp)]
    L.asb_destroy.argtypes = [vp]
    L.asb_destroy.restype = None
    L.asb_last_error.argtypes = [vp]
    L.asb_last_error.restype = C.c_char_p
    L.asb_set_param.argtypes = [vp, C.c_char_p, C.c_double]
    L.asb_upload_reads.argtypes = [vp, u8p, u64p, C.c_uint32]
    L.asb_upload_reads_dev.argtypes = [vp, vp, u64p, C.c_uint32]
    L.asb_upload_reads_scattered.argtypes = [vp, vp, u32p, C.c_uint32]
    L.asb_prepare_pruning.argtypes = [vp, C.c_uint32]
    L.asb_uploaded_ascii_dev.argtypes = [vp, C.POINTER(vp), u64p]
    L.asb_batch_begin.argtypes = [vp, u32p, C.c_uint32, u32p, u32p, u32p, C.c_uint32, C.c_uint32, C.c_uint32]
    L.asb_batch_step.argtypes = [vp, C.POINTER(StepInfo)]
    L.asb_batch_records.argtypes = [vp, vp]
    L.asb_distance_pairs.argtypes = [vp, u32p, u32p, u8p, C.c_uint64, C.c_int, i32p]
    L.asb_batch_records_dev.argtypes = [vp, vp]
    L.asb_int_peak.argtypes = [vp, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.asb_format_records.argtypes = [vp, C.c_uint64, u32p, u32p, u64p, C.c_uint32, u32p, C.c_uint64, C.c_char_p, vp, C.c_uint64]
    L.asb_format_records.restype = C.c_int64
    L.asb_kmer_build.argtypes = [vp, C.c_int]
    L.asb_kmer_shared_pairs.argtypes = [vp, u32p, u32p, C.c_uint64, u32p]
    L.asb_kmer_shared_tile.argtypes = [vp, u32p, C.c_uint32, u32p, C.c_uint32, u32p]
    L.asb_threeway_pairs.argtypes = [vp, u32p, u32p, C.c_uint64, u32p, u32p, C.c_uint32, C.POINTER(StepInfo)]
    L.asb_debug_read.argtypes = [vp, C.c_uint32, C.c_int, u8p, C.c_uint32]
    fp = C.POINTER(C.c_float)
    L.asb_lines_upload.argtypes = [vp, u32p, u32p, u32p, C.c_uint64]
    L.asb_lines_hist.argtypes = [vp, u64p, fp]
    L.asb_lines_besthit.argtypes = [vp, C.c_uint32, u32p, C.c_uint32, u64p, fp]
    L.asb_lines_besthit_fetch.argtypes = [vp, u32p, u32p]
    L.asb_components.argtypes = [vp, u32p, u32p, C.c_uint64, C.c_uint32, u32p, fp]
    u16p = C.POINTER(C.c_uint16)
    L.asb_text_begin.argtypes = [vp, u32p, C.c_uint32, u32p, C.c_uint32, u32p, u16p, C.c_uint32, C.c_char_p, C.c_uint32]
    L.asb_text_load.argtypes = [vp, vp, C.c_uint64, C.c_int]
    L.asb_text_step.argtypes = [vp, C.c_uint64, C.c_uint64, C.c_int, vp, C.c_uint64, u64p]
    L.asb_text_measure.argtypes = [vp, u64p]
    L.asb_lines_append_dev.argtypes = [vp, vp, C.c_uint64]
    L.asb_host_alloc.argtypes = [C.c_uint64, C.POINTER(vp)]
    L.asb_host_free.argtypes = [vp]
    L.asb_host_free.restype = None
    L.asb_lines_count.argtypes = [vp]
    L.asb_lines_count.restype = C.c_uint64
    L.asb_lines_fetch.argtypes = [vp, u32p, u32p, u32p, u8p]
    _LIB = L
    return L


def ptr(a: np.ndarray, t):
    return a.ctypes.data_as(C.POINTER(t))
