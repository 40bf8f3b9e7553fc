"""ctypes front end of oracle/libasref.so plus a pure-Python restatement for tiny cases.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
legs; never by amplicon_sorter_b200/ (the product).  Citations "AS:n" = /root/reference/amplicon_sorter.py:n.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

RECORD = np.dtype([("i_pos", "<u4"), ("j_pos", "<u4"), ("d", "<u4"), ("reverse", "<u4")])


class Stats(C.Structure):
    _fields_ = [("pairs", C.c_uint64), ("rc_retries", C.c_uint64), ("records", C.c_uint64), ("alignments", C.c_uint64)]


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libasref.so")
    src = os.path.join(_HERE, "asref.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libasref.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        u8p, u32p, u64p, i32p = (C.POINTER(t) for t in (C.c_uint8, C.c_uint32, C.c_uint64, C.c_int32))
        for name in ("asref_dp_nw", "asref_dp_hw", "asref_myers_nw", "asref_edlib_like_nw"):
            f = getattr(L, name)
            f.argtypes = [C.c_char_p, C.c_int32, C.c_char_p, C.c_int32]
            f.restype = C.c_int32
        L.asref_compl_reverse.argtypes = [C.c_char_p, C.c_int32, C.c_char_p]
        L.asref_iden.argtypes = [C.c_int32, C.c_int32]
        L.asref_iden.restype = C.c_double
        L.asref_host_threads.restype = C.c_int
        L.asref_process_batch.argtypes = [u8p, u64p, u32p, C.c_uint32, C.c_double, C.c_int, C.c_uint32, C.c_uint32,
                                          C.c_uint32, C.c_void_p, C.c_uint64, C.POINTER(Stats), C.c_int]
        L.asref_process_batch.restype = C.c_int64
        L.asref_format_lines.argtypes = [C.c_void_p, C.c_uint64, u32p, u32p, u64p, C.c_char_p, C.c_uint64]
        L.asref_format_lines.restype = C.c_int64
        L.asref_distance_pairs.argtypes = [u8p, u64p, u32p, u32p, C.c_uint64, C.c_int, C.c_int, i32p, C.c_int]
        L.asref_distance_pairs.restype = C.c_int64
        L.asref_kmer_shared.argtypes = [C.c_char_p, C.c_int32, C.c_char_p, C.c_int32, C.c_int]
        L.asref_kmer_shared.restype = C.c_uint32
        L.asref_nw_path.argtypes = [C.c_char_p, C.c_int32, C.c_char_p, C.c_int32, C.c_char_p, C.c_char_p, i32p]
        L.asref_nw_path.restype = C.c_int64
        _LIB = L
    return _LIB


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


ALGO = {"myers": 0, "edlib_like": 1, "dp": 2}


def nw(a: bytes, b: bytes, algo="dp") -> int:
    f = {"dp": lib().asref_dp_nw, "myers": lib().asref_myers_nw, "edlib_like": lib().asref_edlib_like_nw}[algo]
    if len(a) > len(b):
        a, b = b, a
    return f(a, len(a), b, len(b))


def hw(q: bytes, t: bytes) -> int:
    if len(q) > len(t):
        q, t = t, q
    return lib().asref_dp_hw(q, len(q), t, len(t))


def compl_reverse(s: bytes) -> bytes:
    out = C.create_string_buffer(len(s))
    lib().asref_compl_reverse(s, len(s), out)
    return out.raw[: len(s)]


def iden(d: int, L: int) -> float:
    return lib().asref_iden(d, L)


def host_threads() -> int:
    return lib().asref_host_threads()


def stable_length_order(lengths: np.ndarray) -> np.ndarray:
    """d.sort(key=lambda x: len(x[1]))  AS:669 -- Python's sort is stable."""
    return np.argsort(np.asarray(lengths), kind="stable").astype(np.uint32)


def process_batch(seqs, offs, order, similar_genes=80.0, algo="myers", rows=None, nthreads=0, cap=None):
    """One batch of process_list (AS:647-807).  Returns (records, stats dict)."""
    order = np.ascontiguousarray(order, dtype=np.uint32)
    n = order.shape[0]
    rb, re_, rs = rows if rows is not None else (0, n, 1)
    st = Stats()
    cap = cap or 1 << 16
    while True:
        out = np.empty(cap, dtype=RECORD)
        r = lib().asref_process_batch(_p(seqs, C.c_uint8), _p(offs, C.c_uint64), _p(order, C.c_uint32), n,
                                      float(similar_genes), ALGO[algo], rb, re_, rs, out.ctypes.data, cap,
                                      C.byref(st), nthreads)
        if r >= 0:
            break
        cap = int(st.records) + 16
    stats = {k: int(getattr(st, k)) for k, _ in Stats._fields_}
    return out[:r].copy(), stats


def format_lines(records, order, idx, offs) -> bytes:
    order = np.ascontiguousarray(order, dtype=np.uint32)
    idx = np.ascontiguousarray(idx, dtype=np.uint32)
    records = np.ascontiguousarray(records)
    cap = 64 * (len(records) + 1)
    buf = C.create_string_buffer(cap)
    r = lib().asref_format_lines(records.ctypes.data, len(records), _p(order, C.c_uint32), _p(idx, C.c_uint32),
                                 _p(offs, C.c_uint64), buf, cap)
    assert r >= 0
    return buf.raw[:r]


def distance_pairs(seqs, offs, a, b, algo="myers", mode="NW", nthreads=0) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint32)
    b = np.ascontiguousarray(b, dtype=np.uint32)
    out = np.empty(a.shape[0], dtype=np.int32)
    lib().asref_distance_pairs(_p(seqs, C.c_uint8), _p(offs, C.c_uint64), _p(a, C.c_uint32), _p(b, C.c_uint32),
                               a.shape[0], ALGO[algo], 1 if mode == "HW" else 0, _p(out, C.c_int32), nthreads)
    return out


def kmer_shared(a: bytes, b: bytes, k: int) -> int:
    return lib().asref_kmer_shared(a, len(a), b, len(b), k)


def nw_path(q: bytes, t: bytes, eq_table: bytes):
    ops = C.create_string_buffer(len(q) + len(t) + 1)
    d = C.c_int32()
    n = lib().asref_nw_path(q, len(q), t, len(t), eq_table, ops, C.byref(d))
    return ops.raw[:n].decode(), d.value


# ----------------------------------------------------------------------------------------------
# Pure-Python restatement (tiny inputs only): mirrors the reference statement by statement.
# ----------------------------------------------------------------------------------------------
def py_lev(a: str, b: str) -> int:
    prev = list(range(len(a) + 1))
    for j, cb in enumerate(b, 1):
        cur = [j]
        for i, ca in enumerate(a, 1):
            cur.append(min(prev[i] + 1, cur[i - 1] + 1, prev[i - 1] + (ca != cb)))
        prev = cur
    return prev[-1]


def py_distance(X1: str, X2: str) -> float:
    """distance() AS:224-234 with edlib replaced by the DP definition."""
    if len(X1) > len(X2):
        A2, A1 = X1, X2
    else:
        A1, A2 = X1, X2
    d = py_lev(A1, A2)
    return round(1 - d / len(A2), 3)


def py_compl_reverse(s: str) -> str:
    """compl_reverse() AS:236-241."""
    return s[::-1].translate("".maketrans("ATCGRYKMSW", "TAGCYRMKSW"))


def py_process_list(batches, similar_genes=80.0):
    """queuer AS:662-715 + similarity AS:776-807 -> list of text lines (order of -np 1)."""
    similarg = similar_genes / 100
    lines = []
    for d in batches:
        d.sort(key=lambda x: len(x[1]))
        for position in range(0, len(d) - 1):
            for position2 in range(position + 1, len(d)):
                A1, A2 = d[position], d[position2]
                if len(A1[1]) * 1.05 < len(A2[1]):
                    continue
                iden_ = py_distance(A1[1], A2[1])
                if iden_ >= similarg:
                    lines.append(str(A1[3]) + ":" + str(A2[3]) + ":" + str(iden_))
                elif iden_ < 0.5:
                    iden_ = py_distance(A1[1], py_compl_reverse(A2[1]))
                    if iden_ >= similarg:
                        lines.append(str(A1[3]) + ":" + str(A2[3]) + ":" + str(iden_) + ":reverse")
    return lines


def py_process_consensuslist(indexes, grouplist, comparelist2, similar):
    """process_consensuslist AS:1627-1690 + similarity_species AS:1692-1715 at -np 1 -> list of lines
    (None when the reference would not run the comparison at all: empty last spool chunk, AS:1687)."""
    indexes2 = indexes.copy()
    for x in grouplist:
        for y in x:
            if y.isdigit():
                indexes2.discard(y)
    consensuslist = [[x, y[-1]] for x, y in enumerate(grouplist)]
    comparelist4 = [i for i in comparelist2 if str(i[3]) in indexes2]
    todo, k, chunk = [], 0, 0
    for A1 in comparelist4:
        for A2 in consensuslist:
            if len(A1[1]) * 1.05 < len(A2[1]) or len(A2[1]) * 1.05 < len(A1[1]):
                continue
            todo.append([A1, A2])
            chunk += 1
            if chunk == 2000000:
                chunk = 0
                k += 1
        if k == 100:
            break
    if chunk == 0:
        return None
    lines = []
    for A1, A2 in todo:
        iden = py_distance(A1[1], A2[1])
        if iden >= similar - 0.01:
            lines.append(str(A1[3]) + ":" + str(A2[0]) + ":" + str(iden))
        elif iden < 0.5:
            iden = py_distance(A1[1], py_compl_reverse(A2[1]))
            if iden >= similar - 0.01:
                lines.append(str(A1[3]) + ":" + str(A2[0]) + ":" + str(iden))
    return lines


def py_hw(q: str, t: str) -> int:
    """edlib HW by definition: min over substrings of t (free leading/trailing target gaps)."""
    prev = list(range(len(q) + 1))
    best = prev[-1]
    for cb in t:
        cur = [0]
        for i, ca in enumerate(q, 1):
            cur.append(min(prev[i] + 1, cur[i - 1] + 1, prev[i - 1] + (ca != cb)))
        prev = cur
        best = min(best, prev[-1])
    return best


def py_iden_consensus(todolist):
    """iden_consensus AS:1139-1158 -> lines 'y,z,iden' (distance() with mode='HW', AS:224-234)."""
    def dist_hw(X1, X2):
        A1, A2 = (X2, X1) if len(X1) > len(X2) else (X1, X2)
        return round(1 - py_hw(A1, A2) / len(A2), 3)

    lines = []
    for A1, A2, y, z in todolist:
        idenlist = [dist_hw(A1, A2), dist_hw(A1, py_compl_reverse(A2))]
        idenlist.sort(reverse=True)
        if idenlist[0] >= 0.60:
            lines.append(str(y) + "," + str(z) + "," + str(idenlist[0]))
    return lines
