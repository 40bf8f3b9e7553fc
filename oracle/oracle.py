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
                                          C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint64, C.POINTER(Stats), C.c_int]
        L.asref_process_batch.restype = C.c_int64
        L.asref_format_lines.argtypes = [C.c_void_p, C.c_uint64, u32p, u32p, u64p, C.c_char_p, C.c_uint64]
        L.asref_format_lines.restype = C.c_int64
        L.asref_distance_pairs.argtypes = [u8p, u64p, u32p, u32p, C.c_uint64, C.c_int, C.c_int, i32p, C.c_int]
        L.asref_distance_pairs.restype = C.c_int64
        L.asref_kmer_shared.argtypes = [C.c_char_p, C.c_int32, C.c_char_p, C.c_int32, C.c_int]
        L.asref_kmer_shared.restype = C.c_uint32
        L.asref_nw_path.argtypes = [C.c_char_p, C.c_int32, C.c_char_p, C.c_int32, C.c_char_p, C.c_char_p, i32p]
        L.asref_nw_path.restype = C.c_int64
        L.asref_besthit.argtypes = [u32p, u32p, u32p, C.c_uint64, C.c_uint32, u32p, C.c_uint32, u32p, u32p]
        L.asref_besthit.restype = C.c_int64
        L.asref_components.argtypes = [u32p, u32p, C.c_uint64, C.c_uint32, u32p]
        L.asref_components.restype = None
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


def process_batch(seqs, offs, order, similar_genes=80.0, algo="myers", rows=None, nthreads=0, cap=None, col_step=1):
    """One batch of process_list (AS:647-807).  Returns (records, stats dict).
    rows = (begin, end, step) and col_step restrict the run to a strided sample of the pair set (bench.py)."""
    order = np.ascontiguousarray(order, dtype=np.uint32)
    n = order.shape[0]
    rb, re_, rs = rows if rows is not None else (0, n, 1)
    st = Stats()
    cap = cap or 1 << 16
    while True:
        out = np.empty(cap, dtype=RECORD)
        r = lib().asref_process_batch(_p(seqs, C.c_uint8), _p(offs, C.c_uint64), _p(order, C.c_uint32), n,
                                      float(similar_genes), ALGO[algo], rb, re_, rs, int(col_step), out.ctypes.data, cap,
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


def c_lev(a: str, b: str) -> int:
    """The same textbook DP, compiled (asref_dp_nw): for restatement runs on fixture-sized inputs."""
    return nw(a.encode("latin-1"), b.encode("latin-1"), "dp")


def c_hw(q: str, t: str) -> int:
    return hw(q.encode("latin-1"), t.encode("latin-1"))


def py_distance(X1: str, X2: str, lev=py_lev) -> float:
    """distance() AS:224-234 with edlib replaced by the DP definition."""
    if len(X1) > len(X2):
        A2, A1 = X1, X2
    else:
        A1, A2 = X1, X2
    d = lev(A1, A2)
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


def py_process_consensuslist(indexes, grouplist, comparelist2, similar, lev=py_lev):
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
        iden = py_distance(A1[1], A2[1], lev)
        if iden >= similar - 0.01:
            lines.append(str(A1[3]) + ":" + str(A2[0]) + ":" + str(iden))
        elif iden < 0.5:
            iden = py_distance(A1[1], py_compl_reverse(A2[1]), lev)
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


def py_iden_consensus(todolist, hw_fn=None):
    """iden_consensus AS:1139-1158 -> lines 'y,z,iden' (distance() with mode='HW', AS:224-234)."""
    hw_fn = hw_fn or py_hw

    def dist_hw(X1, X2):
        A1, A2 = (X2, X1) if len(X1) > len(X2) else (X1, X2)
        return round(1 - hw_fn(A1, A2) / len(A2), 3)

    lines = []
    for A1, A2, y, z in todolist:
        idenlist = [dist_hw(A1, A2), dist_hw(A1, py_compl_reverse(A2))]
        idenlist.sort(reverse=True)
        if idenlist[0] >= 0.60:
            lines.append(str(y) + "," + str(z) + "," + str(idenlist[0]))
    return lines


# ----------------------------------------------------------------------------------------------
# Consumers of <stem>_compare.tmp.
# ----------------------------------------------------------------------------------------------
def besthit(a, b, milli, min_milli=0, member_bits=None):
    """asref_besthit -> (line numbers, first admitted line of each survivor's key); key-ascending, list order."""
    a = np.ascontiguousarray(a, dtype=np.uint32)
    b = np.ascontiguousarray(b, dtype=np.uint32)
    milli = np.ascontiguousarray(milli, dtype=np.uint32)
    n = a.shape[0]
    n_keys = int(max(a.max(initial=0), b.max(initial=0))) + 1 if n else 1
    out_line = np.empty(max(n, 1), dtype=np.uint32)
    out_first = np.empty(max(n, 1), dtype=np.uint32)
    mp = None
    if member_bits is not None:
        member_bits = np.ascontiguousarray(member_bits, dtype=np.uint32)
        assert member_bits.shape[0] * 32 >= n_keys
        mp = _p(member_bits, C.c_uint32)
    r = lib().asref_besthit(_p(a, C.c_uint32), _p(b, C.c_uint32), _p(milli, C.c_uint32), n, int(min_milli), mp, n_keys,
                            _p(out_line, C.c_uint32), _p(out_first, C.c_uint32))
    assert r >= 0
    return out_line[:r].copy(), out_first[:r].copy()


def components(a, b, n_nodes):
    a = np.ascontiguousarray(a, dtype=np.uint32)
    b = np.ascontiguousarray(b, dtype=np.uint32)
    label = np.empty(max(n_nodes, 1), dtype=np.uint32)
    lib().asref_components(_p(a, C.c_uint32), _p(b, C.c_uint32), a.shape[0], n_nodes, _p(label, C.c_uint32))
    return label[:n_nodes]


def py_ssg(text: str):
    """SSG AS:809-835 on the text of the file."""
    totalsimil = 0
    tempdict = {}
    for line in text.splitlines():
        simil = float(line.strip().split(":")[2])
        tempdict[simil] = tempdict.get(simil, 0) + 1
        totalsimil += simil
    templist = sorted(tempdict, reverse=True)
    b = int(totalsimil * 0.06)
    N6 = 0
    for x in templist:
        N6 += tempdict[x] * x
        if N6 >= b:
            return int(x * 100)
    return None


def py_besthit_templist(text: str, ssg=None, indexes=None):
    """The filter of update_list AS:986-1012 (ssg is None) or read_indexes AS:1364-1398 (ssg = similar_species_groups
    as a fraction, indexes = set of idx strings) -> the sorted templist of [idxA, idxB, iden] string triples."""
    tempdict = {}
    for line in text.splitlines():
        e = line.strip().split(":")
        if ssg is None:
            e = e[:3]
        elif not (float(e[2]) >= ssg and len({e[0], e[1]}.intersection(indexes)) > 0):
            continue
        lst = tempdict.setdefault(e[1], [])
        lst.append(e)
        lst.sort(key=lambda x: (int(x[1]), float(x[2])))
        keep = [x for i, x in enumerate(lst) if i + 1 == len(lst) or not (x[2] < lst[i + 1][2])]
        lst[:] = keep
    templist = [x for sub in tempdict.values() for x in sub]
    if ssg is None:
        templist.sort(key=lambda x: (float(x[2]), int(x[1])), reverse=True)
    else:
        templist.sort(key=lambda x: float(x[2]), reverse=True)
    return templist


def py_groups(templist, update_with_list=False):
    """Greedy grouping AS:1022-1031 (read_indexes' variant AS:1403-1409 grows a group from a list) + merge_groups
    AS:1057-1086 -> (number of greedy groups, merged groups in order)."""
    grouplist = []
    for x in templist:
        for s in grouplist:
            if len({x[0], x[1]}.intersection(s)) > 0:
                s.update([x[0], x[1]] if update_with_list else {x[0], x[1]})
                break
        else:
            grouplist.append({x[0], x[1]})
    n_greedy = len(grouplist)
    if n_greedy > 1:
        grouplist = [set(g) for g in grouplist if len(g) > 1]
        a1, a2 = len(grouplist), 0
        while a1 > a2:
            a1 = len(grouplist)
            for p1 in range(len(grouplist) - 1):
                for p2 in range(p1 + 1, len(grouplist)):
                    if len(grouplist[p1].intersection(grouplist[p2])) > 0:
                        grouplist[p1] = grouplist[p1].union(grouplist[p2])
                        grouplist[p2].clear()
            grouplist = [g for g in grouplist if len(g) > 0]
            a2 = len(grouplist)
    return n_greedy, grouplist
