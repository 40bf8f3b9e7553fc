"""Thin object wrapper over the C ABI (include/asb200.h): one Engine = one GPU context."""
from __future__ import annotations

import ctypes as C
import threading

import numpy as np

from . import _ffi
from ._ffi import RECORD, EngineError, StepInfo, ptr


class TextChunk:
    """One slab of tempfile text: `data` is a bytes-like view that stays valid until release()."""

    __slots__ = ("data", "_slot")

    def __init__(self, data, slot=None):
        self.data, self._slot = data, slot

    def release(self):
        if self._slot is not None:
            self._slot.free.set()
            self._slot = None


class _Slot:
    __slots__ = ("ptr", "cap", "free")

    def __init__(self):
        self.ptr, self.cap, self.free = None, 0, threading.Event()
        self.free.set()


TOTALS = ("pairs", "n_records", "fwd_survivors", "rc_survivors", "zone_checks", "word_updates", "screen_word_updates",
          "useful_word_updates", "screen_useful_word_updates", "screen_ms", "total_ms", "launches", "pruned_pairs", "cluster_ms", "lists_ms")


class EngineBase:
    """What every engine offers on top of batch_begin / batch_step / text_begin / text_load / text_chunks."""

    _pool = None

    def _text_async(self, sink, append_lines: bool = True):
        """Lines of the staged record set -> sink, on a helper thread (the text stage has its own CUDA stream): the
        comparison of the next slab runs meanwhile.  Returns a future; wait for it before the next text_load."""
        if self._pool is None:
            from concurrent.futures import ThreadPoolExecutor

            self._pool = ThreadPoolExecutor(1, thread_name_prefix="asb200-text")

        def emit(n):
            for chunk in self.text_chunks(n, append_lines):
                sink(chunk)

        return lambda n: self._pool.submit(emit, n)

    def compare_text(self, order, hi, dpass, drev, text_tables, sink, rank=0, world=1):
        """All steps of one batch, the lines of every step handed to `sink(chunk)` in file order as they are produced:
        while slab k + 1 is compared, the text of slab k is assembled on the GPU's second stream, copied out and
        appended by the writer threads.  Returns the totals dict."""
        self.batch_begin(order, hi, dpass, drev, rank, world)
        self.text_begin(*text_tables)
        tot = dict.fromkeys(TOTALS, 0)
        tot["steps"] = 0
        import time

        submit, pending = self._text_async(sink), None
        t_step = t_wait = t_load = 0.0
        try:
            while True:
                t0 = time.perf_counter()
                info = self.batch_step()
                t1 = time.perf_counter()
                if pending is not None:
                    pending.result()  # the previous slab's text is out: its staging buffers are free again
                    pending = None
                t2 = time.perf_counter()
                t_step += t1 - t0
                t_wait += t2 - t1
                if info is None:
                    break
                for k in TOTALS:
                    tot[k] += info.get(k, 0)
                tot["steps"] += 1
                if info["n_records"]:
                    self.text_load(None, info["n_records"])
                    pending = submit(info["n_records"])
                t_load += time.perf_counter() - t2
        finally:
            if pending is not None:
                pending.result()
        # host view of the loop: time inside asb_batch_step, time waiting for text that did not overlap, staging
        tot["host_ms"] = {"batch_step calls": t_step * 1e3, "waiting for the text stage": t_wait * 1e3, "staging records": t_load * 1e3}
        return tot


class Engine(EngineBase):
    def __init__(self, device: int = 0, stream: int | None = None):
        self._lib = _ffi.load()
        h = C.c_void_p()
        rc = self._lib.asb_create(int(device), C.c_void_p(stream) if stream else None, C.byref(h))
        if rc != 0:
            raise EngineError(rc, "asb_create failed (no usable CUDA device?)")
        self._h = h
        self.device = device
        self.n_reads = 0
        self._lines_token = None  # groups.upload: the Lines object whose arrays are resident on the device
        self._slots = [_Slot() for _ in range(6)]  # pinned host buffers of the text ring (writer threads + one being filled)
        self._slot_next = 0

    # -- plumbing ------------------------------------------------------------------------------
    def _check(self, rc):
        if rc < 0:
            raise EngineError(rc, self._lib.asb_last_error(self._h).decode())
        return rc

    def close(self):
        if getattr(self, "_h", None):
            self._evict_lines()
            if self._pool is not None:
                self._pool.shutdown(wait=True)
                self._pool = None
            for sl in self._slots:
                sl.free.wait()
                if sl.ptr:
                    self._lib.asb_host_free(sl.ptr)
                    sl.ptr, sl.cap = None, 0
            self._lib.asb_destroy(self._h)
            self._h = None

    def _evict_lines(self):
        """The resident line set is about to be replaced: let its owner (groups.DeviceLines) take a host copy first."""
        tok, self._lines_token = self._lines_token, None
        if tok is not None and hasattr(tok, "materialize"):
            tok.materialize()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_param(self, name: str, value: float):
        self._check(self._lib.asb_set_param(self._h, name.encode(), float(value)))

    # -- reads ---------------------------------------------------------------------------------
    def upload_reads(self, ascii_buf: np.ndarray, offs: np.ndarray):
        ascii_buf = np.ascontiguousarray(ascii_buf, dtype=np.uint8)
        offs = np.ascontiguousarray(offs, dtype=np.uint64)
        n = offs.shape[0] - 1
        self._check(self._lib.asb_upload_reads(self._h, ptr(ascii_buf, C.c_uint8), ptr(offs, C.c_uint64), n))
        self.n_reads = n
        self._lens = (offs[1:] - offs[:-1]).astype(np.uint32)

    def upload_reads_scattered(self, ptrs: np.ndarray, lens: np.ndarray):
        """upload_reads with every read in a host buffer of its own: ptrs[r] = address of read r's bytes (uint64),
        lens[r] = its length.  The buffers must stay valid during the call (pyhost.collect borrows them from strs)."""
        ptrs = np.ascontiguousarray(ptrs, dtype=np.uint64)
        lens = np.ascontiguousarray(lens, dtype=np.uint32)
        n = lens.shape[0]
        self._check(self._lib.asb_upload_reads_scattered(self._h, C.c_void_p(ptrs.ctypes.data), ptr(lens, C.c_uint32), n))
        self.n_reads = n
        self._lens = lens.copy()

    def uploaded_ascii_tensor(self, dev):
        """The read bytes of the last host-side upload as a uint8 torch tensor VIEW of the engine's device staging
        buffer (valid until the next upload): what rank 0 broadcasts to the other ranks."""
        import torch

        p, n = C.c_void_p(), C.c_uint64()
        self._check(self._lib.asb_uploaded_ascii_dev(self._h, C.byref(p), C.byref(n)))
        if not n.value:
            return torch.empty(0, dtype=torch.uint8, device=dev)

        class _View:  # zero-copy: torch reads the CUDA array interface
            __cuda_array_interface__ = {"shape": (int(n.value),), "typestr": "|u1", "data": (int(p.value), False), "version": 2}

        return torch.as_tensor(_View(), device=dev)

    def prepare_pruning(self, kmax: int):
        """Build the read clusters of the pivot bound now (cut-off kmax = the largest dpass of the coming batch)."""
        self._check(self._lib.asb_prepare_pruning(self._h, int(kmax)))

    def upload_reads_tensor(self, t, offs: np.ndarray):
        """upload_reads with the read bytes already on this engine's GPU (a uint8 torch tensor)."""
        offs = np.ascontiguousarray(offs, dtype=np.uint64)
        n = offs.shape[0] - 1
        self._check(self._lib.asb_upload_reads_dev(self._h, C.c_void_p(t.data_ptr()), ptr(offs, C.c_uint64), n))
        self.n_reads = n
        self._lens = (offs[1:] - offs[:-1]).astype(np.uint32)

    def debug_read(self, r: int, strand: int = 0) -> bytes:
        out = np.empty(int(self._lens[r]), dtype=np.uint8)
        self._check(self._lib.asb_debug_read(self._h, r, strand, ptr(out, C.c_uint8), out.shape[0]))
        return out.tobytes()

    # -- all-pairs batch -------------------------------------------------------------------------
    def batch_begin(self, order, hi, dpass, drev, rank=0, world=1):
        order = np.ascontiguousarray(order, dtype=np.uint32)
        hi = np.ascontiguousarray(hi, dtype=np.uint32)
        dpass = np.ascontiguousarray(dpass, dtype=np.uint32)
        drev = np.ascontiguousarray(drev, dtype=np.uint32)
        assert dpass.shape == drev.shape
        self._keep = (order, hi, dpass, drev)
        self._check(self._lib.asb_batch_begin(self._h, ptr(order, C.c_uint32), order.shape[0], ptr(hi, C.c_uint32),
                                              ptr(dpass, C.c_uint32), ptr(drev, C.c_uint32), dpass.shape[0], rank, world))

    def batch_step(self):
        """Run the next slab.  Returns a StepInfo dict, or None when the batch is exhausted."""
        info = StepInfo()
        rc = self._check(self._lib.asb_batch_step(self._h, C.byref(info)))
        if rc == _ffi.ASB_DONE:
            return None
        return info.as_dict()

    def batch_records(self, n_records: int) -> np.ndarray:
        out = np.empty(n_records, dtype=RECORD)
        if n_records:
            self._check(self._lib.asb_batch_records(self._h, out.ctypes.data))
        return out

    def batch_records_dev(self, dev_ptr: int):
        """Copy the last step's records into device memory at `dev_ptr` (same GPU)."""
        self._check(self._lib.asb_batch_records_dev(self._h, C.c_void_p(dev_ptr)))

    def int_peak(self, iters: int = 2000):
        """(pure LOP3, Myers-mix) INT32 ALU throughput of this GPU in 1e12 lane-ops/s, measured live."""
        a, b = C.c_double(), C.c_double()
        self._check(self._lib.asb_int_peak(self._h, iters, C.byref(a), C.byref(b)))
        return a.value, b.value

    def compare_batch(self, order, hi, dpass, drev, rank=0, world=1, fetch=True):
        """All steps of one batch.  Returns (records sorted by (i_pos, j_pos), totals dict)."""
        self.batch_begin(order, hi, dpass, drev, rank, world)
        recs = []
        tot = dict.fromkeys(TOTALS, 0)
        tot["steps"] = 0
        while True:
            info = self.batch_step()
            if info is None:
                break
            for k in tot:
                if k in info:
                    tot[k] += info[k]
            tot["steps"] += 1
            if fetch and info["n_records"]:
                recs.append(self.batch_records(info["n_records"]))
        out = np.concatenate(recs) if recs else np.empty(0, dtype=RECORD)
        return out, tot

    # -- three-way rule on an explicit pair list (similarity_species) ---------------------------------
    def threeway_pairs(self, q, t, dpass, drev):
        """Records (i_pos = q id, j_pos = t id, d, reverse) sorted by (q, t), plus the step info."""
        q = np.ascontiguousarray(q, dtype=np.uint32)
        t = np.ascontiguousarray(t, dtype=np.uint32)
        dpass = np.ascontiguousarray(dpass, dtype=np.uint32)
        drev = np.ascontiguousarray(drev, dtype=np.uint32)
        info = StepInfo()
        self._check(self._lib.asb_threeway_pairs(self._h, ptr(q, C.c_uint32), ptr(t, C.c_uint32), q.shape[0], ptr(dpass, C.c_uint32),
                                                 ptr(drev, C.c_uint32), dpass.shape[0], C.byref(info)))
        return self.batch_records(info.n_records), info.as_dict()

    # -- consumers of <stem>_compare.tmp on integer lines (SSG, best-hit filter, grouping) -----------------
    def lines_upload(self, a, b, milli):
        a = np.ascontiguousarray(a, dtype=np.uint32)
        b = np.ascontiguousarray(b, dtype=np.uint32)
        milli = np.ascontiguousarray(milli, dtype=np.uint32)
        self._evict_lines()
        self._check(self._lib.asb_lines_upload(self._h, ptr(a, C.c_uint32), ptr(b, C.c_uint32), ptr(milli, C.c_uint32), a.shape[0]))

    def lines_count(self) -> int:
        return int(self._lib.asb_lines_count(self._h))

    def lines_fetch(self):
        """(a, b, milli, rev) of the resident line set written by text_step."""
        n = self.lines_count()
        a, b, m = (np.empty(n, dtype=np.uint32) for _ in range(3))
        r = np.empty(n, dtype=np.uint8)
        self._check(self._lib.asb_lines_fetch(self._h, ptr(a, C.c_uint32), ptr(b, C.c_uint32), ptr(m, C.c_uint32), ptr(r, C.c_uint8)))
        return a, b, m, r.astype(bool)

    # -- the tempfile as text, assembled on the device --------------------------------------------------
    def text_begin(self, idx_sorted, lbase, soff, milli, sbuf: bytes):
        idx_sorted = np.ascontiguousarray(idx_sorted, dtype=np.uint32)
        lbase = np.ascontiguousarray(lbase, dtype=np.uint32)
        soff = np.ascontiguousarray(soff, dtype=np.uint32)
        milli = np.ascontiguousarray(milli, dtype=np.uint16)
        self._evict_lines()
        self._check(self._lib.asb_text_begin(self._h, ptr(idx_sorted, C.c_uint32), idx_sorted.shape[0], ptr(lbase, C.c_uint32), lbase.shape[0],
                                             ptr(soff, C.c_uint32), ptr(milli, C.c_uint16), milli.shape[0], sbuf, len(sbuf)))

    def _acquire_slot(self, cap: int) -> _Slot:
        sl = self._slots[self._slot_next]
        self._slot_next = (self._slot_next + 1) % len(self._slots)
        sl.free.wait()  # the writer thread is done with this buffer
        if sl.cap < cap:
            if sl.ptr:
                self._lib.asb_host_free(sl.ptr)
                sl.ptr, sl.cap = None, 0
            want = 1 << max(20, (int(cap) - 1).bit_length())  # powers of two: a slot is re-pinned a handful of times at most
            p = C.c_void_p()
            if self._lib.asb_host_alloc(want, C.byref(p)) != 0:
                raise EngineError(-3, f"cannot pin {want} bytes of host memory for the tempfile text")
            sl.ptr, sl.cap = p.value, want
        return sl

    def text_step(self, first: int, count: int, append_lines: bool = True) -> TextChunk:
        """Lines of records [first, first + count) of the staged record set (text_load).  append_lines: also append
        them in integer form to the resident line set (off when they get there through lines_append_tensor)."""
        sl = self._acquire_slot(_ffi.TEXT_MAX_LINE * int(count) + 64)
        nb = C.c_uint64()
        self._check(self._lib.asb_text_step(self._h, int(first), int(count), int(bool(append_lines)), C.c_void_p(sl.ptr), sl.cap, C.byref(nb)))
        sl.free.clear()
        return TextChunk(memoryview((C.c_char * nb.value).from_address(sl.ptr)).cast("B"), sl)

    def text_chunks(self, n_records: int, append_lines: bool = True):
        """The staged record set as text, in chunks of at most TEXT_CHUNK records, in file order."""
        for first in range(0, int(n_records), _ffi.TEXT_CHUNK):
            yield self.text_step(first, min(_ffi.TEXT_CHUNK, int(n_records) - first), append_lines)

    def text_measure(self) -> int:
        """Bytes the staged record set prints to."""
        nb = C.c_uint64()
        self._check(self._lib.asb_text_measure(self._h, C.byref(nb)))
        return int(nb.value)

    def lines_append_tensor(self, recs):
        """Append an (n, 4) int32 tensor of records on this GPU, already in file order, to the resident line set."""
        if int(recs.shape[0]):
            self._check(self._lib.asb_lines_append_dev(self._h, C.c_void_p(recs.data_ptr()), int(recs.shape[0])))

    def text_load(self, dev_ptr: int | None, n_records: int, sort: bool = True):
        """Stage the record set of the text stage: the last step's records (dev_ptr None), or n_records asb_records in
        device memory at dev_ptr (sorted if asked)."""
        self._check(self._lib.asb_text_load(self._h, C.c_void_p(dev_ptr) if dev_ptr else None, int(n_records), int(bool(sort))))

    def step_records_tensor(self, n_records: int, dev):
        """The last step's records as an (n, 4) int32 tensor on this engine's GPU (for the NCCL gather)."""
        import torch

        t = torch.empty((int(n_records), 4), dtype=torch.int32, device=dev)
        if n_records:
            self.batch_records_dev(t.data_ptr())
        return t

    def text_load_tensor(self, recs, sort: bool = True):
        """Stage an (n, 4) int32 tensor of records on this engine's GPU (the NCCL gather of several ranks' lists)."""
        n = int(recs.shape[0])
        if n:
            self.text_load(recs.data_ptr(), n, sort)
        return n

    def lines_hist(self):
        """(hist[1001] of iden*1000 over the resident lines, device ms)."""
        hist = np.zeros(1001, dtype=np.uint64)
        ms = C.c_float()
        self._check(self._lib.asb_lines_hist(self._h, ptr(hist, C.c_uint64), C.byref(ms)))
        return hist, ms.value

    def lines_besthit(self, min_milli: int = 0, member_bits=None):
        """(surviving line numbers, first admitted line of each survivor's key, device ms); key-ascending, list order."""
        n, ms = C.c_uint64(), C.c_float()
        mp, mw = None, 0
        if member_bits is not None:
            member_bits = np.ascontiguousarray(member_bits, dtype=np.uint32)
            mp, mw = ptr(member_bits, C.c_uint32), member_bits.shape[0]
        self._check(self._lib.asb_lines_besthit(self._h, int(min_milli), mp, mw, C.byref(n), C.byref(ms)))
        line = np.empty(n.value, dtype=np.uint32)
        first = np.empty(n.value, dtype=np.uint32)
        if n.value:
            self._check(self._lib.asb_lines_besthit_fetch(self._h, ptr(line, C.c_uint32), ptr(first, C.c_uint32)))
        return line, first, ms.value

    def components(self, a, b, n_nodes: int):
        """(label[n_nodes] = smallest node id of the node's component, device ms)."""
        a = np.ascontiguousarray(a, dtype=np.uint32)
        b = np.ascontiguousarray(b, dtype=np.uint32)
        label = np.empty(max(int(n_nodes), 1), dtype=np.uint32)
        ms = C.c_float()
        self._check(self._lib.asb_components(self._h, ptr(a, C.c_uint32), ptr(b, C.c_uint32), a.shape[0], int(n_nodes),
                                             ptr(label, C.c_uint32), C.byref(ms)))
        return label[: int(n_nodes)], ms.value

    # -- k-mer side output (new; no reference counterpart) --------------------------------------------
    def kmer_build(self, k: int = 6):
        self._check(self._lib.asb_kmer_build(self._h, int(k)))

    def kmer_shared_pairs(self, a, b) -> np.ndarray:
        a = np.ascontiguousarray(a, dtype=np.uint32)
        b = np.ascontiguousarray(b, dtype=np.uint32)
        out = np.empty(a.shape[0], dtype=np.uint32)
        self._check(self._lib.asb_kmer_shared_pairs(self._h, ptr(a, C.c_uint32), ptr(b, C.c_uint32), a.shape[0], ptr(out, C.c_uint32)))
        return out

    def kmer_shared_tile(self, rows, cols) -> np.ndarray:
        rows = np.ascontiguousarray(rows, dtype=np.uint32)
        cols = np.ascontiguousarray(cols, dtype=np.uint32)
        out = np.empty((rows.shape[0], cols.shape[0]), dtype=np.uint32)
        self._check(self._lib.asb_kmer_shared_tile(self._h, ptr(rows, C.c_uint32), rows.shape[0], ptr(cols, C.c_uint32),
                                                   cols.shape[0], ptr(out, C.c_uint32)))
        return out

    # -- distance() on explicit pairs --------------------------------------------------------------
    def distance_pairs(self, a, b, strand=None, mode: str = "NW") -> np.ndarray:
        a = np.ascontiguousarray(a, dtype=np.uint32)
        b = np.ascontiguousarray(b, dtype=np.uint32)
        out = np.empty(a.shape[0], dtype=np.int32)
        sp = None
        if strand is not None:
            strand = np.ascontiguousarray(strand, dtype=np.uint8)
            sp = ptr(strand, C.c_uint8)
        self._check(self._lib.asb_distance_pairs(self._h, ptr(a, C.c_uint32), ptr(b, C.c_uint32), sp, a.shape[0], 1 if mode == "HW" else 0,
                                                 ptr(out, C.c_int32)))
        return out
