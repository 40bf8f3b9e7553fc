"""Host-side mirror of the reference's all-pairs stage: same names, arguments and side effects.

``process_list(comparelist2, tempfile)`` replaces amplicon_sorter.py:647-774 (queuer / feeder /
consumer threads, forked ``similarity`` workers :776-807) with one GPU engine call per batch.  The
Python host keeps what the reference keeps in Python: the stable length sort (:669), the window test
``len_i*1.05 < len_j`` (:679) in the reference's float arithmetic, the text formatting of
``iden = round(1 - d/len, 3)`` (:233) and the tempfile protocol.  Everything O(N^2) happens on the GPU.
"""
from __future__ import annotations

import glob
import os

import numpy as np

from . import pyhost, thresholds
from .engine import Engine


class NoReadsToCompare(Exception):
    """amplicon_sorter.py:702-706 + :768-772: zero comparable pairs -> skip this input file."""


def batch_geometry(lengths_sorted: np.ndarray):
    """hi[p] = last position j kept by the window test of amplicon_sorter.py:679 for row p.

    The batch is sorted by length, so the kept partners of row p are the contiguous positions
    p+1 .. hi[p].  The test is `len_i*1.05 < len_j` -> skip, evaluated in float64 exactly as Python
    does (int*float -> float; int < 2**53 compares exactly).
    """
    L = np.asarray(lengths_sorted, dtype=np.int64)
    lim = L.astype(np.float64) * 1.05
    # last j with not (lim_i < L_j)  <=>  L_j <= lim_i
    hi = np.searchsorted(L.astype(np.float64), lim, side="right").astype(np.int64) - 1
    hi = np.maximum(hi, np.arange(L.shape[0]))
    return hi.astype(np.uint32)


def _iden_table(lens_sorted: np.ndarray, records: np.ndarray, dcap: np.ndarray | None = None):
    """String table of str(round(1 - d/L, 3)) -- Python's own float/round/str, evaluated once per
    (L, d) that can occur: every longer-read length L present in the records, d = 0 .. max d seen for it."""
    L = lens_sorted[records["j_pos"]].astype(np.int64)
    d = records["d"].astype(np.int64)
    lmax = int(L.max())
    if dcap is not None and dcap.shape[0] > lmax:
        # emitted distances never exceed the pass cut-off of their length: no scan of d needed
        present = np.bincount(L, minlength=lmax + 1) > 0
        cap = dcap[: lmax + 1].astype(np.int64)
        dmax = np.where(present & (cap < 0x7FFFFFFF), cap, -1)  # 0xFFFFFFFF = no distance passes for that length
    else:
        dmax = np.zeros(lmax + 1, dtype=np.int64) - 1
        np.maximum.at(dmax, L, d)
    lbase = np.full(lmax + 1, np.uint64(0xFFFFFFFFFFFFFFFF), dtype=np.uint64)
    strs, soff, n, milli = [], [0], 0, []
    for length in np.nonzero(dmax >= 0)[0].tolist():
        lbase[length] = n
        for dd in range(int(dmax[length]) + 1):
            v = round(1 - dd / length, 3)
            t = str(v)
            strs.append(t)
            soff.append(soff[-1] + len(t))
            milli.append(int(round(v * 1000)))  # exact: v has <= 3 decimals
        n += int(dmax[length]) + 1
    return lbase, np.asarray(soff, dtype=np.uint32), "".join(strs).encode("ascii"), n, np.asarray(milli, dtype=np.uint32)


def format_records(records: np.ndarray, idx_sorted: np.ndarray, lens_sorted: np.ndarray, dcap: np.ndarray | None = None,
                   with_lines: bool = False):
    """Records -> the lines of amplicon_sorter.py:792-798: 'idxA:idxB:iden' or '...:reverse'.
    Text assembly runs in the library's host-side C (asb_format_records); the iden strings are Python's.
    with_lines: also return the same lines in integer form (idxA, idxB, iden*1000, reverse flag) for groups.Lines."""
    import ctypes as C

    from . import _ffi

    n = int(records.shape[0])
    if n == 0:
        z = np.zeros(0, dtype=np.uint32)
        return ("", (z, z, z, z.astype(bool))) if with_lines else ""
    records = np.ascontiguousarray(records)
    idx32 = np.ascontiguousarray(idx_sorted, dtype=np.uint32)
    len32 = np.ascontiguousarray(lens_sorted, dtype=np.uint32)
    lbase, soff, sbuf, nstr, milli = _iden_table(np.asarray(lens_sorted), records, dcap)
    cap = n * (32 + 8) + 64
    out = C.create_string_buffer(cap)
    lib = _ffi.load()
    k = lib.asb_format_records(records.ctypes.data, n, _ffi.ptr(idx32, C.c_uint32), _ffi.ptr(len32, C.c_uint32),
                               _ffi.ptr(lbase, C.c_uint64), lbase.shape[0], _ffi.ptr(soff, C.c_uint32), nstr, sbuf,
                               C.cast(out, C.c_void_p), cap)
    if k < 0:
        raise RuntimeError(f"asb_format_records failed ({k})")
    text = out.raw[:k].decode("ascii")
    if not with_lines:
        return text
    e = lbase[np.asarray(lens_sorted)[records["j_pos"]].astype(np.int64)].astype(np.int64) + records["d"].astype(np.int64)
    return text, (idx32[records["i_pos"]], idx32[records["j_pos"]], milli[e], records["reverse"] != 0)


def text_tables(idx_sorted: np.ndarray, lens_sorted: np.ndarray, dpass: np.ndarray):
    """What asb_text_begin needs to print the lines of amplicon_sorter.py:792-798 on the device: the idx printed for
    every sorted position, and Python's own ``str(round(1 - d/L, 3))`` (:233) for every (L, d) a record can carry --
    every length L present in the batch, d = 0 .. dpass[L] (an emitted distance never exceeds the pass cut-off of its
    longer read).  String number lbase[L] + d; milli = iden * 1000 is the integer form the consumers of the file use."""
    lens = np.unique(np.asarray(lens_sorted, dtype=np.int64))
    lmax = int(lens[-1]) if lens.size else 0
    lbase = np.full(lmax + 1, 0xFFFFFFFF, dtype=np.uint32)
    strs, n = [], 0
    for length in lens.tolist():
        cap = int(dpass[length]) if length < dpass.shape[0] else 0xFFFFFFFF
        if length <= 0 or cap == 0xFFFFFFFF:  # no distance passes for that length
            continue
        lbase[length] = n
        strs.extend(round(1 - dd / length, 3) for dd in range(cap + 1))
        n += cap + 1
    milli = np.fromiter((int(round(v * 1000)) for v in strs), dtype=np.uint16, count=n)  # exact: v has <= 3 decimals
    strs = [str(v) for v in strs]
    soff = np.zeros(n + 1, dtype=np.uint32)
    np.cumsum(np.fromiter((len(t) for t in strs), dtype=np.uint32, count=n), out=soff[1:])
    return np.ascontiguousarray(idx_sorted, dtype=np.uint32), lbase, soff, milli, "".join(strs).encode("ascii")


def ascii_view(s: str) -> np.ndarray:
    """The bytes of `s` as a uint8 array.  Reads are ASCII, and CPython keeps an ASCII str as one byte per character:
    PyUnicode_AsUTF8AndSize then returns the str's own buffer, so 100 MB of reads are not copied a second time.  Any
    other text (the reference would accept it: every distinct character is a symbol) takes the latin-1 copy."""
    import ctypes as C

    if s.isascii():
        size = C.c_ssize_t()
        fn = C.pythonapi.PyUnicode_AsUTF8AndSize
        fn.restype, fn.argtypes = C.c_void_p, [C.py_object, C.POINTER(C.c_ssize_t)]
        p = fn(s, C.byref(size))
        if p and size.value == len(s):
            return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(len(s),)) if len(s) else np.zeros(0, np.uint8)
    return np.frombuffer(s.encode("latin-1"), dtype=np.uint8)


class TextSink:
    """Appends the chunks of text to the tempfile from a small pool of writer threads, so that write(2) overlaps the
    GPU work on the next slab (the reference appends per 1 M-pair chunk, amplicon_sorter.py:802-807).  Every chunk's
    file offset is fixed when it is handed in (chunks arrive in file order), so the threads write disjoint ranges
    with pwrite -- one thread copies ~1.5 GB/s into the page cache, a 540 MB tempfile needs several.
    Callable: sink(chunk)."""

    THREADS = int(os.environ.get("ASB200_WRITER_THREADS", "4"))

    def __init__(self, path: str, existing: bool = False):
        import queue
        import threading

        self.path = path
        if existing:  # a worker rank writing its pieces into the file rank 0 created
            self.base = 0
        else:
            self.fd = os.open(path, os.O_WRONLY | os.O_CREAT | os.O_APPEND, 0o666)  # :803 append mode ...
            self.base = os.lseek(self.fd, 0, os.SEEK_END)
            os.close(self.fd)
        self.fd = os.open(path, os.O_WRONLY)  # ... realised with explicit offsets (pwrite ignores them under O_APPEND)
        self.pos = self.base  # where the next chunk goes
        self.q = queue.Queue()
        self.err = None
        self.bytes = 0
        self.threads = [threading.Thread(target=self._run, daemon=True) for _ in range(self.THREADS)]
        for t in self.threads:
            t.start()

    def seek(self, offset: int):
        """The next chunks go to `offset` onwards (multi-GPU: this rank's piece of a slab)."""
        self.pos = int(offset)

    def __call__(self, chunk):
        n = len(chunk.data)
        self.q.put((chunk, self.pos))
        self.pos += n
        self.bytes += n

    def _run(self):
        while True:
            item = self.q.get()
            if item is None:
                return
            c, off = item
            try:
                if self.err is None:
                    mv = memoryview(c.data)
                    done = 0
                    while done < len(mv):
                        done += os.pwrite(self.fd, mv[done:], off + done)
            except BaseException as exc:  # reported by close()
                self.err = exc
            finally:
                c.release()

    def close(self):
        for _ in self.threads:
            self.q.put(None)
        for t in self.threads:
            t.join()
        os.close(self.fd)
        if self.err is not None:
            raise self.err


class NullSink:
    """Drops the text (measurement runs: the lines are assembled and copied out, not written)."""

    path = None
    base = 0

    def __init__(self):
        self.bytes = 0

    def seek(self, offset: int):
        pass

    def __call__(self, chunk):
        self.bytes += len(chunk.data)
        chunk.release()

    def close(self):
        pass


class AllPairs:
    """State shared by the batches of one input file: one engine, reads uploaded once."""

    def __init__(self, engine: Engine | None = None, device: int = 0):
        self.engine = engine or Engine(device)
        self.stats = {"pairs": 0, "records": 0, "fwd_survivors": 0, "rc_survivors": 0, "zone_checks": 0,
                      "word_updates": 0, "gpu_ms": 0.0, "screen_ms": 0.0}

    def set_lengths(self, seqs: list[str]):
        lens = np.fromiter(map(len, seqs), dtype=np.uint64, count=len(seqs))
        self.lens = lens.astype(np.int64)
        return lens

    def upload(self, seqs: list[str], lens=None):
        if lens is None:
            lens = self.set_lengths(seqs)
        offs = np.zeros(len(seqs) + 1, dtype=np.uint64)
        np.cumsum(lens, out=offs[1:])
        joined = "".join(seqs)  # kept alive: `buf` may be a view of it
        buf = ascii_view(joined)
        self.engine.upload_reads(buf, offs)
        del buf, joined

    def _account(self, tot):
        self.stats["pairs"] += tot["pairs"]
        self.stats["records"] += tot["n_records"]
        for k in ("fwd_survivors", "rc_survivors", "zone_checks", "word_updates", "screen_ms"):
            self.stats[k] += tot.get(k, 0)
        self.stats["gpu_ms"] += tot.get("total_ms", 0.0)

    def compare(self, read_ids: np.ndarray, similar_genes: float, rank=0, world=1):
        """One batch (read ids in the batch's CURRENT list order) -> (perm, order, lens_sorted, records, tl)."""
        perms, order, lens_sorted, hi, tl = self.plan([read_ids])
        if tl == 0:
            return perms[0], order, lens_sorted, np.empty(0, dtype=self._rec_dtype()), 0
        dpass, drev = thresholds.tables(similar_genes / 100, int(lens_sorted[-1]) + 1)  # similarg :783
        recs, tot = self.engine.compare_batch(order, hi, dpass, drev, rank, world)
        self._account(tot)
        self.last_dpass = dpass
        return perms[0], order, lens_sorted, recs, tl

    def plan(self, batches_read_ids):
        """All batches of one input file as ONE engine batch: the length-sorted batches are laid end to end
        (position = batch offset + position in the batch), a row's window never leaves its own batch, and the
        records come back in (batch, i, j) order -- the order of the reference's file.  A default-mode run is
        dozens of 1,000-read batches of 0.5 M pairs each; one launch per batch cannot fill 148 SMs.
        Returns (perms per batch, order, lens_sorted, hi with GLOBAL positions, tl)."""
        perms, orders, lens_all, his = [], [], [], []
        base = 0
        for read_ids in batches_read_ids:
            read_ids = np.asarray(read_ids, dtype=np.int64)
            lens = self.lens[read_ids]
            perm = np.argsort(lens, kind="stable")  # d.sort(key=lambda x: len(x[1]))  :669
            perms.append(perm)
            orders.append(read_ids[perm].astype(np.uint32))
            lens_all.append(lens[perm])
            his.append(batch_geometry(lens[perm]).astype(np.int64) + base)
            base += read_ids.shape[0]
        order = np.concatenate(orders) if orders else np.zeros(0, np.uint32)
        lens_sorted = np.concatenate(lens_all) if lens_all else np.zeros(0, np.int64)
        hi = (np.concatenate(his) if his else np.zeros(0, np.int64)).astype(np.uint32)
        tl = int((hi.astype(np.int64) - np.arange(hi.shape[0])).sum())  # the reference's tl (:684)
        return perms, order, lens_sorted, hi, tl

    @staticmethod
    def prepare_tables(lens_sorted, idx_sorted, similar_genes: float):
        """The integer cut-offs and the iden string tables of one planned batch (pure host work)."""
        import time

        t0 = time.perf_counter()
        dpass, drev = thresholds.tables(similar_genes / 100, int(lens_sorted.max()) + 1)  # similarg :783
        tables = text_tables(idx_sorted, lens_sorted, dpass)
        return dpass, drev, tables, (time.perf_counter() - t0) * 1e3

    def compare_to_file(self, order, lens_sorted, hi, prep, out_path: str):
        """Decide every pair of the planned batch and append the lines to `out_path` as they are produced."""
        import time

        dpass, drev, tables, prep_ms = prep
        t1 = time.perf_counter()
        sink = TextSink(out_path)
        try:
            tot = self.engine.compare_text(order, hi, dpass, drev, tables, sink)
            t2 = time.perf_counter()
        finally:
            sink.close()
        t3 = time.perf_counter()
        self.stats["phases_ms"] = {"cut-off + iden string tables": prep_ms, "all slabs (GPU)": (t2 - t1) * 1e3,
                                   "of which pivots + read assignment": tot.get("cluster_ms", 0.0), "writer thread drain": (t3 - t2) * 1e3,
                                   **{"loop: " + k: v for k, v in tot.get("host_ms", {}).items()}}
        self._account(tot)
        self.stats["text_bytes"] = self.stats.get("text_bytes", 0) + sink.bytes
        return tot

    @staticmethod
    def _rec_dtype():
        from ._ffi import RECORD
        return RECORD


def process_list(self, tempfile, args, engine: Engine | None = None, stats_out: dict | None = None, device: int = 0):
    """Drop-in for ``process_list(self, tempfile)`` (amplicon_sorter.py:647).

    self     : comparelist2 -- list of batches of [id, SEQ, tag, idx] records (:568-622)
    tempfile : os.path.join(outputfolder, '<stem>_compare.tmp') as passed at :2041
    args     : the reference's argparse namespace (uses .outputfolder, .similar_genes)

    Side effects kept: (1) every batch list is left sorted by length in place (:669); (2) the
    tempfile holds one line per passing pair in the -np 1 order (batch, i, j) and exists (possibly
    empty) whenever at least one pair was compared (:802-807 opens it in append mode per chunk);
    (3) zero comparable pairs -> 'No reads to compare, exiting...' is appended to results.txt and an
    Exception skips the file (:702-706, :768-772); (4) stale *.todo spool files are removed (:655-660).

    The lines are assembled on the GPU slab by slab (csrc/text.cuh) and appended by a writer thread while the next
    slab is compared; the same lines stay on the device in integer form for SSG / update_list / read_indexes.
    """
    outputfolder = args.outputfolder
    for x in glob.glob(os.path.join(outputfolder, "*.todo")):
        try:
            os.remove(x)
        except FileNotFoundError:
            pass
    out_path = os.path.join(outputfolder, tempfile)  # :779-780 joins again; absolute paths win
    try:
        os.remove(out_path)
    except FileNotFoundError:
        pass

    import time

    t0 = time.perf_counter()
    # distinct records of all batches, keyed by their idx field (:560-561); -ra batches overlap
    from operator import itemgetter

    flat = self[0] if len(self) == 1 else [rec for d in self for rec in d]
    ap = AllPairs(engine, device)
    own = isinstance(ap.engine, Engine)
    can_scatter = own or bool(getattr(ap.engine, "scattered_ok", False))  # (the sharded facade on GPUs; not the CPU test engines)
    fast = pyhost.collect(flat) if can_scatter else None  # one C pass: idx keys, and where every SEQ's bytes live
    seqs = ptrs = None
    if fast is not None:
        keys, ptrs, lens = fast
    else:
        keys = np.fromiter(map(itemgetter(3), flat), dtype=np.int64, count=len(flat))
    if keys.shape[0] < 2 or bool((keys[1:] > keys[:-1]).all()):  # read_file numbers the records in file order: nothing to merge
        rid_to_idx, inv = keys, np.arange(keys.shape[0], dtype=np.int64)
        if fast is None:
            seqs = list(map(itemgetter(1), flat))
    else:
        rid_to_idx, first, inv = np.unique(keys, return_index=True, return_inverse=True)  # read id = rank of the idx value
        if fast is None:
            seqs = [flat[i][1] for i in first.tolist()]
        else:
            ptrs, lens = ptrs[first], lens[first]
    batch_rids = np.split(inv.astype(np.int64), np.cumsum([len(d) for d in self])[:-1]) if len(self) else []

    from . import groups

    t1 = time.perf_counter()
    # The upload runs beside the host's own preparation (length sort, windows, side effect (1), string tables) when the
    # engine is this process's own: the C call releases the GIL.  With the record walker the reads go to the GPU
    # straight from the str objects (asb_upload_reads_scattered); without it they are joined first.  A sharded facade
    # broadcasts the job with collectives and stays on the calling thread.
    if fast is not None:
        ap.lens = lens.astype(np.int64)
    else:
        lens = ap.set_lengths(seqs)
    up = None
    if own:
        import threading

        box = []

        def _up():
            try:
                if fast is not None:
                    ap.engine.upload_reads_scattered(ptrs, lens)  # `flat` keeps the strs alive
                else:
                    ap.upload(seqs, lens)
                if lens.shape[0]:  # the pivots of the pivot bound, while this thread's owner sorts and builds tables
                    dp, _ = thresholds.tables(args.similar_genes / 100, int(lens.max()) + 1)
                    ok = dp[dp != 0xFFFFFFFF]
                    ap.engine.prepare_pruning(int(ok.max()) if ok.size else 0)
            except BaseException as exc:  # re-raised on the calling thread
                box.append(exc)

        up = threading.Thread(target=_up, name="asb200-upload")
        up.start()
    elif fast is not None:
        ap.engine.upload_reads_scattered(ptrs, lens)
    else:
        ap.upload(seqs, lens)
    t2 = time.perf_counter()
    live = [(d, rids) for d, rids in zip(self, batch_rids) if len(d)]
    perms, order, lens_sorted, hi, tl_total = ap.plan([rids for _, rids in live])
    for (d, _), perm in zip(live, perms):
        d[:] = [d[i] for i in perm.tolist()]  # side effect (1): batch left length-sorted in place
    for old, _ in groups.CACHE.values():  # lines of an earlier input file are never asked for again (:2179 deletes that file)
        old.discard()
    groups.CACHE.clear()
    prep = ap.prepare_tables(lens_sorted, rid_to_idx[order.astype(np.int64)], args.similar_genes) if tl_total else None
    if up is not None:
        up.join()
        if box:
            raise box[0]
    t3 = time.perf_counter()
    if tl_total:
        # records come out sorted by (global i, global j) = (batch, i, j): the reference's -np 1 file order
        ap.compare_to_file(order, lens_sorted, hi, prep, out_path)
        # the consumers of the file (SSG, update_list, read_indexes) get its lines without parsing the text
        groups.CACHE[os.path.abspath(out_path)] = (groups.DeviceLines(ap.engine), os.path.getsize(out_path))
    t4 = time.perf_counter()
    ap.stats["phases_ms"] = {"records -> read set": (t1 - t0) * 1e3, "join + upload (threaded on one engine)": (t2 - t1) * 1e3,
                             "sort + windows + tables, upload joined": (t3 - t2) * 1e3,
                             "compare + write": (t4 - t3) * 1e3, **ap.stats.get("phases_ms", {})}
    if stats_out is not None:
        stats_out.update(ap.stats)
        stats_out["tl"] = tl_total
    if engine is None:
        ap.engine.close()  # an engine handed in by the caller (shared across files / ranks) stays open
    if tl_total == 0:
        print("No reads to compare, exiting...")
        with open(os.path.join(outputfolder, "results.txt"), "a") as rf:
            rf.write("No reads to compare, exiting...")
        raise NoReadsToCompare()
    return None


def process_consensuslist(indexes, grouplist, group_filename, *, args, comparelist2, similar, engine):
    """Drop-in for ``process_consensuslist(indexes, grouplist, group_filename)`` (amplicon_sorter.py:1627-1690)
    together with its worker ``similarity_species`` (:1692-1715): reads not yet in a sub-group x the
    consensus of every sub-group, same three-way rule as ``similarity`` with the cut ``similar - 0.01``
    evaluated in the reference's float arithmetic (:1700), lines ``read_idx:group_no:iden`` in the
    -np 1 order (read, then group) appended to ``<group>.tmp``.

    Quirks kept: the two-sided 5 % window (:1663); the cap of 100 spool files x 2,000,000 comparisons,
    checked after each read (:1675-1676); nothing at all is compared when the LAST spool chunk is empty
    (:1687 -- a comparison count that is an exact multiple of 2,000,000, including zero).
    `args`, `comparelist2` and `similar` are the reference's module globals at call time."""
    outputfolder = args.outputfolder
    group_tempfile = os.path.join(outputfolder, group_filename).replace(".group", ".tmp")
    for x in [group_tempfile] + glob.glob(os.path.join(outputfolder, "*.todo")):
        try:
            os.remove(x)
        except FileNotFoundError:
            pass
    indexes2 = indexes.copy()
    for x in grouplist:
        for y in x:
            if y.isdigit():
                indexes2.discard(y)
    consensuslist = [[x, y[-1]] for x, y in enumerate(grouplist)]
    comparelist4 = [i for i in comparelist2 if str(i[3]) in indexes2]
    rl = np.fromiter((len(r[1]) for r in comparelist4), dtype=np.int64, count=len(comparelist4))
    cl = np.fromiter((len(c[1]) for c in consensuslist), dtype=np.int64, count=len(consensuslist))
    # window :1663  `len(A1)*1.05 < len(A2) or len(A2)*1.05 < len(A1)` -> skip   (float64, as in Python)
    keep = ~((rl[:, None].astype(np.float64) * 1.05 < cl[None, :]) | (cl[None, :].astype(np.float64) * 1.05 < rl[:, None])) \
        if rl.size and cl.size else np.zeros((rl.size, cl.size), dtype=bool)
    per_read = keep.sum(axis=1)
    cum = np.cumsum(per_read)
    stop = np.nonzero(cum // 2000000 >= 100)[0]  # :1675 `if k == 100: break`, tested after each read
    n_reads_used = int(stop[0]) + 1 if stop.size else rl.size
    keep[n_reads_used:, :] = False
    l_total = int(cum[n_reads_used - 1]) if n_reads_used else 0
    print(group_filename + "----> " + str(l_total) + " comparisons to calculate")
    if l_total % 2000000 == 0:  # :1687 `if len(todolist) > 0:` -- the last chunk is empty, nothing runs
        return None
    xs, ys = np.nonzero(keep)  # enumeration order: read, then group
    seqs = [r[1] for r in comparelist4[:n_reads_used]] + [c[1] for c in consensuslist]
    lens = np.concatenate([rl[:n_reads_used], cl])
    offs = np.zeros(len(seqs) + 1, dtype=np.uint64)
    np.cumsum(lens, out=offs[1:])
    buf = np.frombuffer("".join(seqs).encode("latin-1"), dtype=np.uint8)
    engine.upload_reads(buf, offs)
    cut = similar - 0.01  # :1700, raw float (0.94 - 0.01 = 0.9299999999999999)
    dpass, drev = thresholds.tables(cut, int(lens.max()) + 1)
    q = (ys + n_reads_used).astype(np.uint32)  # the consensus is the DP query: many reads share its match masks
    t = xs.astype(np.uint32)
    recs, _ = engine.threeway_pairs(q, t, dpass, drev)
    order = np.lexsort((recs["i_pos"], recs["j_pos"]))  # back to (read, group) order
    lines = []
    for r in recs[order].tolist():
        qi, ti, d, _rev = r
        y, x = qi - n_reads_used, ti
        iden = round(1 - d / max(int(rl[x]), int(cl[y])), 3)  # :233 len(longer)
        lines.append(str(comparelist4[x][3]) + ":" + str(consensuslist[y][0]) + ":" + str(iden) + "\n")
    with open(os.path.join(outputfolder, group_tempfile), "a") as f:  # :1709
        f.writelines(lines)
    return None


def iden_consensus_files(outputfolder, consensus_tempfile, stringx, *, engine):
    """Drop-in for ``do_parallel(..., iden_consensus, ...)`` (amplicon_sorter.py:1160-1203 driving the worker
    ``iden_consensus`` :1139-1158): every spooled ``[A1, A2, y, z]`` gets the edlib-HW (infix) identity on
    both strands, the larger of the two is kept, and ``y,z,iden`` is appended to the consensus tempfile
    if it is >= 0.60 (:1151) -- in the -np 1 order (spool files by mtime, entries in list order)."""
    import pickle

    names = [n for n in os.listdir(outputfolder) if n.endswith(".todo")]
    names.sort(key=lambda x: os.path.getmtime(os.path.join(outputfolder, x)))
    for name in names:
        print(stringx + name)
        with open(os.path.join(outputfolder, name), "rb") as rf:
            todolist = pickle.load(rf)
        os.remove(os.path.join(outputfolder, name))
        lines = []
        if todolist:
            ids: dict = {}
            seqs: list = []
            a = np.empty(len(todolist), dtype=np.uint32)
            b = np.empty(len(todolist), dtype=np.uint32)
            for p, (A1, A2, _y, _z) in enumerate(todolist):
                for arr, s in ((a, A1), (b, A2)):
                    k = ids.get(s)
                    if k is None:
                        k = ids[s] = len(seqs)
                        seqs.append(s)
                    arr[p] = k
            lens = np.fromiter((len(x) for x in seqs), dtype=np.uint64, count=len(seqs))
            offs = np.zeros(len(seqs) + 1, dtype=np.uint64)
            np.cumsum(lens, out=offs[1:])
            engine.upload_reads(np.frombuffer("".join(seqs).encode("latin-1"), dtype=np.uint8), offs)
            n = len(todolist)
            strand = np.concatenate([np.zeros(n, np.uint8), np.ones(n, np.uint8)])
            d = engine.distance_pairs(np.concatenate([a, a]), np.concatenate([b, b]), strand, mode="HW")
            for p, (A1, A2, y, z) in enumerate(todolist):
                L = max(len(A1), len(A2))
                iden = max(round(1 - int(d[p]) / L, 3), round(1 - int(d[n + p]) / L, 3))  # :1145-1150
                if iden >= 0.60:  # :1151
                    lines.append(str(y) + "," + str(z) + "," + str(iden) + "\n")
        with open(os.path.join(outputfolder, consensus_tempfile), "a") as f:  # :1154 (created even when empty)
            f.writelines(lines)
