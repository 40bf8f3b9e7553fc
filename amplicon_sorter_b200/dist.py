"""One process per GPU: shard the all-pairs stage over ranks with torch.distributed.

The pair set shards with no exchange step: every rank holds all reads and
decides its own contiguous range of the rows of every slab (ranges of equal pair counts, asb_batch_step),
prints the lines of that range and writes them at their offset of the tempfile.  The only communication is
(1) rank 0 broadcasting the job (reads + batch composition) to the workers, (2) one small all_gather per slab
(record counts, status, bytes printed) and (3) the gather of the compacted per-rank record lists to rank 0,
where they become the resident integer lines -- NCCL over NVLink on GPUs, gloo in CPU tests.

    torchrun --nproc-per-node 8 -m amplicon_sorter_b200 --script amplicon_sorter.py -i reads.fastq ...

Rank 0 runs the reference script; ranks > 0 sit in `worker_loop` and serve its process_list calls.
"""
from __future__ import annotations

import os

import numpy as np

from ._ffi import RECORD
from .engine import EngineBase


def world():
    return int(os.environ.get("WORLD_SIZE", "1"))


def rank():
    return int(os.environ.get("RANK", "0"))


def local_rank():
    return int(os.environ.get("LOCAL_RANK", "0"))


def init_from_env(backend: str | None = None):
    """Initialise torch.distributed when launched under torchrun; returns (rank, world, device)."""
    import torch
    import torch.distributed as dist

    w = world()
    use_cuda = torch.cuda.is_available() and backend != "gloo"
    dev = torch.device("cuda", local_rank()) if use_cuda else torch.device("cpu")
    if use_cuda:
        torch.cuda.set_device(dev)
    if w > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        if use_cuda:
            dist.init_process_group("nccl", device_id=dev)
        else:
            dist.init_process_group("gloo")
    return rank(), w, dev


def _bcast_array(arr, dev, src=0, keep_on_device=False):
    """Broadcast a numpy array (shape/dtype known on src only).  keep_on_device: return the uint8 tensor on `dev`."""
    import torch
    import torch.distributed as dist

    meta = [None]
    if dist.get_rank() == src:
        meta = [(arr.shape, arr.dtype.str)]
    dist.broadcast_object_list(meta, src=src)
    shape, dt = meta[0]
    if dist.get_rank() == src:
        import warnings

        with warnings.catch_warnings():  # a read-only view of a bytes object: it is only read (copied to `dev`)
            warnings.simplefilter("ignore", UserWarning)
            t = torch.from_numpy(np.ascontiguousarray(arr).view(np.uint8).reshape(-1)).to(dev)
    else:
        t = torch.empty(int(np.prod(shape)) * np.dtype(dt).itemsize, dtype=torch.uint8, device=dev)
    if t.numel():
        dist.broadcast(t, src=src)
    if keep_on_device:
        return t
    return t.cpu().numpy().view(np.dtype(dt)).reshape(shape)


INLINE_BYTES = 1 << 20  # arrays up to this size travel inside the job's pickle; larger ones as device tensors


def broadcast_job(job: dict | None, dev):
    """Rank 0 passes {'op': ..., arrays...}; every rank returns the same dict.  ONE object broadcast carries the op,
    the scalars and every small array (a batch is a handful of 400 KB arrays: two dozen separate collectives cost
    more than the comparison of a slab at 8 GPUs); only big arrays (the read bytes) go as tensors of their own."""
    import torch.distributed as dist

    box = [None]
    if dist.get_rank() == 0:
        box = [{k: (("__tensor__",) if isinstance(v, np.ndarray) and v.nbytes > INLINE_BYTES else v) for k, v in job.items()}]
    dist.broadcast_object_list(box, src=0)
    out = {}
    for k, v in box[0].items():
        if isinstance(v, tuple) and v == ("__tensor__",):
            # the read bytes go straight from the broadcast buffer into the engine (no host round trip on any rank)
            out[k] = _bcast_array(job[k] if job is not None else None, dev, keep_on_device=(k == "buf" and dev.type == "cuda"))
        else:
            out[k] = v
    return out


class DistributedAbort(SystemExit):
    """Some rank failed inside a sharded call.  A SystemExit on purpose: the reference's per-file
    `except Exception: continue` (amplicon_sorter.py:2184) must not swallow it and walk into the next
    collective while the other ranks are gone."""


def agree(ok: bool, dev, what: str = ""):
    """Collective: every rank reports whether its last engine call worked; if any did not, ALL ranks leave."""
    import torch
    import torch.distributed as dist

    flag = torch.tensor([0 if ok else 1], dtype=torch.int32, device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MAX)
    if int(flag.item()):
        raise DistributedAbort(f"amplicon_sorter_b200: rank {dist.get_rank()}: a rank failed in {what or 'a sharded call'}; aborting all ranks")


def gather_step(mine, status: int, done: bool, dev, extra: int = 0, counts=(0, 0)):
    """Per-slab K6: the ranks' record tensors ((k, 4) int32 on `dev`, each sorted) -> their concatenation in rank order
    on rank 0 (None elsewhere), all on the device: counts all_gather, padded gather over NCCL / NVLink, no host round
    trip.  The counts message also carries every rank's status and done flag, so that a failed rank (or ranks that
    disagree on the number of slabs) stops ALL ranks at the same step instead of leaving them parked in a collective --
    and three more integers per rank (`extra`: the bytes its piece of the slab prints to; `counts`: its pairs and its
    list entries, for the slab-size hints every rank derives from the same totals).
    Returns (records on rank 0 | None, done, list of every rank's extra); the per-rank heads stay in gather_step.heads."""
    import torch
    import torch.distributed as dist

    w, r = dist.get_world_size(), dist.get_rank()
    n = int(mine.shape[0]) if mine is not None else 0
    head = torch.tensor([n, int(status), int(bool(done)), int(extra), int(counts[0]), int(counts[1])], dtype=torch.int64, device=dev)
    heads = torch.empty((w, 6), dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(heads, head) if dev.type == "cuda" else dist.all_gather(list(heads.unbind(0)), head)
    heads = heads.tolist()
    gather_step.heads = heads
    if any(h[1] for h in heads):
        bad = [i for i, h in enumerate(heads) if h[1]]
        raise DistributedAbort(f"amplicon_sorter_b200: rank {r}: engine failure on rank(s) {bad}; aborting all ranks")
    dones = {h[2] for h in heads}
    if len(dones) != 1:
        raise DistributedAbort(f"amplicon_sorter_b200: rank {r}: ranks disagree on the number of slabs; aborting all ranks")
    extras = [h[3] for h in heads]
    if dones.pop():
        return None, True, extras
    sizes = [h[0] for h in heads]
    mx = max(sizes)
    if mx == 0:
        return (mine[:0] if r == 0 else None), False, extras
    pad = torch.empty((mx, 4), dtype=torch.int32, device=dev)
    pad[:n] = mine
    buf = torch.empty((w, mx, 4), dtype=torch.int32, device=dev) if r == 0 else None
    dist.gather(pad, list(buf.unbind(0)) if r == 0 else None, dst=0)
    if r != 0:
        return None, False, extras
    return torch.cat([buf[i, :s] for i, s in enumerate(sizes)]), False, extras


def gather_records(recs: np.ndarray, dev, status: int = 0):
    """Per-rank record arrays -> merged array in (i_pos, j_pos) order on rank 0 (None elsewhere)."""
    import torch

    mine = torch.from_numpy(np.ascontiguousarray(recs).view(np.uint32).reshape(-1, 4).view(np.int32).copy()).to(dev)
    allr, _, _ = gather_step(mine, status, False, dev)
    if allr is None:
        return None
    key = (allr[:, 0].to(torch.int64) << 32) | (allr[:, 1].to(torch.int64) & 0xFFFFFFFF)
    allr = allr[torch.argsort(key, stable=True)]
    return allr.cpu().numpy().view(np.uint32).reshape(-1).view(RECORD)


class ShardedEngine(EngineBase):
    """Engine facade for rank 0: every all-pairs call is executed by all ranks on their shard of the rows."""

    def __init__(self, engine, dev):
        self.engine, self.dev = engine, dev
        self.aborted = False

    def _guard(self, fn, what):
        """Run this rank's part of a sharded call, then agree with the other ranks that it worked."""
        err = None
        try:
            out = fn()
        except Exception as exc:  # noqa: BLE001 -- reported through the collective below
            err, out = exc, None
        try:
            agree(err is None, self.dev, what)
        except DistributedAbort:
            self.aborted = True
            if err is not None:
                import traceback

                traceback.print_exception(err)
            raise
        return out

    def upload_reads(self, buf, offs):
        job = broadcast_job({"op": "upload", "buf": np.ascontiguousarray(buf), "offs": np.ascontiguousarray(offs)}, self.dev)
        self._guard(lambda: _upload(self.engine, job), "upload_reads")

    @property
    def scattered_ok(self):
        return self.dev.type == "cuda"

    def upload_reads_scattered(self, ptrs, lens):
        """Engine.upload_reads_scattered on rank 0 (reads gathered from the records' own buffers straight into device
        memory), then the device copy of the bytes is broadcast over NCCL: the host never joins the reads."""
        import torch.distributed as dist

        err, t = None, None
        try:
            self.engine.upload_reads_scattered(ptrs, lens)
            t = self.engine.uploaded_ascii_tensor(self.dev)
        except Exception as exc:  # noqa: BLE001 -- reported through the collective below
            err = exc
        offs = np.zeros(len(lens) + 1, dtype=np.uint64)
        np.cumsum(np.asarray(lens, dtype=np.uint64), out=offs[1:])
        broadcast_job({"op": "upload_dev", "ok": err is None, "nbytes": int(t.numel()) if t is not None else 0, "offs": offs}, self.dev)
        if err is None and t.numel():
            dist.broadcast(t, src=0)
        try:
            agree(err is None, self.dev, "upload_reads")
        except DistributedAbort:
            self.aborted = True
            if err is not None:
                import traceback

                traceback.print_exception(err)
            raise

    def set_param(self, name, value):
        broadcast_job({"op": "param", "name": name, "value": float(value)}, self.dev)
        self.engine.set_param(name, value)

    def compare_batch(self, order, hi, dpass, drev, rank=0, world=1, fetch=True):
        import torch.distributed as dist

        job = broadcast_job({"op": "batch", "order": np.asarray(order, np.uint32), "hi": np.asarray(hi, np.uint32),
                             "dpass": np.asarray(dpass, np.uint32), "drev": np.asarray(drev, np.uint32)}, self.dev)
        try:
            return _run_shard(self.engine, job, self.dev, dist.get_rank(), dist.get_world_size())
        except DistributedAbort:
            self.aborted = True
            raise

    def compare_text(self, order, hi, dpass, drev, text_tables, sink, rank=0, world=1):
        """Engine.compare_text over all ranks: per slab, the ranks' records are gathered on rank 0's GPU (gather_step),
        merged, printed there (csrc/text.cuh) and handed to `sink`; the resident line set ends up on rank 0's engine."""
        import torch.distributed as dist

        idx_sorted, lbase, soff, milli, sbuf = text_tables
        job = broadcast_job({"op": "batch_text", "order": np.asarray(order, np.uint32), "hi": np.asarray(hi, np.uint32),
                             "dpass": np.asarray(dpass, np.uint32), "drev": np.asarray(drev, np.uint32),
                             "t_idx": np.asarray(idx_sorted, np.uint32), "t_lbase": np.asarray(lbase, np.uint32), "t_soff": np.asarray(soff, np.uint32),
                             "t_milli": np.asarray(milli, np.uint16), "t_sbuf": np.frombuffer(sbuf, dtype=np.uint8),
                             "path": getattr(sink, "path", None), "base": int(getattr(sink, "pos", 0))}, self.dev)
        try:
            return _run_shard_text(self.engine, job, self.dev, dist.get_rank(), dist.get_world_size(), sink)
        except DistributedAbort:
            self.aborted = True
            raise

    @property
    def _lines_token(self):
        return self.engine._lines_token

    @_lines_token.setter
    def _lines_token(self, v):
        self.engine._lines_token = v

    def lines_count(self):
        return self.engine.lines_count()

    def lines_fetch(self):
        return self.engine.lines_fetch()

    # Only the all-pairs stage shards.  The other stages (reads x consensuses, consensus x consensus, the consumers of the
    # tempfile) are milliseconds of work: the launcher runs them on `self.engine`, rank 0's own engine ("replicas only").
    def close(self):
        if not self.aborted:
            broadcast_job({"op": "stop"}, self.dev)
        self.engine.close()


def _upload(engine, job):
    if isinstance(job["buf"], np.ndarray):
        engine.upload_reads(job["buf"], job["offs"])
    else:
        engine.upload_reads_tensor(job["buf"], job["offs"])


def _upload_dev(facade, engine, job, dev):
    """Worker side of ShardedEngine.upload_reads_scattered: receive rank 0's device copy of the read bytes."""
    import torch
    import torch.distributed as dist

    if not job["ok"]:  # rank 0 failed before the broadcast: only the agreement is left
        agree(True, dev, "upload_reads")
        return
    t = torch.empty(int(job["nbytes"]), dtype=torch.uint8, device=dev)
    if t.numel():
        dist.broadcast(t, src=0)
    facade._guard(lambda: engine.upload_reads_tensor(t, job["offs"]), "upload_reads")


def _run_shard(engine, job, dev, r, w):
    """One rank's part of ShardedEngine.compare_batch (records back on the host of rank 0)."""
    recs, tot, status = np.empty(0, dtype=RECORD), None, 0
    try:
        recs, tot = engine.compare_batch(job["order"], job["hi"], job["dpass"], job["drev"], r, w)
    except Exception:  # noqa: BLE001 -- gather_step tells every rank
        import traceback

        traceback.print_exc()
        status = 1
    merged = gather_records(recs, dev, status)
    tot = _sum_totals(tot, dev)
    tot["n_records"] = int(merged.shape[0]) if merged is not None else 0
    return merged, tot


def _sum_totals(tot, dev):
    import torch
    import torch.distributed as dist

    keys = sorted(k for k, v in tot.items() if isinstance(v, (int, float)) and not isinstance(v, bool))
    t = torch.tensor([float(tot[k]) for k in keys], dtype=torch.float64, device=dev)
    dist.all_reduce(t)  # sums; times become rank-sums (reported as such)
    return {k: (int(v) if isinstance(tot[k], int) else float(v)) for k, v in zip(keys, t.tolist())}


def _run_shard_text(engine, job, dev, r, w, sink):
    """One rank's part of ShardedEngine.compare_text.  Every rank compares its contiguous piece of every slab, prints
    ITS OWN lines (second CUDA stream + helper thread, while the next slab is compared) and writes them at its own
    offset of the tempfile: one small all_gather per slab (record counts, status, done flag, bytes of the piece) fixes
    the offsets, so formatting, the device-to-host copy and write(2) all scale with the ranks.  The records are also
    gathered on rank 0's GPU (NCCL), where they become the resident integer lines of the later stages."""
    from . import host
    from .engine import TOTALS

    tot = dict.fromkeys(TOTALS, 0)
    tot["steps"] = 0
    status = 0
    my_sink = sink
    try:
        engine.batch_begin(job["order"], job["hi"], job["dpass"], job["drev"], r, w)
        engine.text_begin(job["t_idx"], job["t_lbase"], job["t_soff"], job["t_milli"], job["t_sbuf"].tobytes())
        if r != 0:
            my_sink = host.TextSink(job["path"], existing=True) if job["path"] else host.NullSink()
    except Exception:  # noqa: BLE001
        import traceback

        traceback.print_exc()
        status = 1
    n_rec, file_pos = 0, int(job["base"])
    g_pairs = g_left = g_recs = 0
    submit, pending = engine._text_async(my_sink, append_lines=False), None
    try:
        while True:
            info, mine, nbytes = None, None, 0
            if not status:
                try:
                    info = engine.batch_step()
                    if pending is not None:  # the previous slab's text went out while this slab was compared
                        pending.result()
                        pending = None
                    if info is not None:
                        mine = engine.step_records_tensor(info["n_records"], dev)
                        if info["n_records"]:
                            engine.text_load(None, info["n_records"])
                            nbytes = engine.text_measure()
                except Exception:  # noqa: BLE001 -- gather_step tells every rank
                    import traceback

                    traceback.print_exc()
                    status = 1
            counts = (info.get("pairs", 0), info.get("pairs", 0) - info.get("pruned_pairs", 0)) if info is not None else (0, 0)
            merged, done, sizes = gather_step(mine, status, info is None, dev, nbytes, counts)  # raises DistributedAbort on every rank if any failed
            if done:
                break
            # the next slab's size: lines and list entries per pair over ALL ranks so far -- the same numbers on every rank
            # (each rank's own counts would cut different slabs)
            g_pairs += sum(h[4] for h in gather_step.heads)
            g_left += sum(h[5] for h in gather_step.heads)
            g_recs += sum(h[0] for h in gather_step.heads)
            if not status and g_pairs:
                engine.set_param("slab_keep_ratio", g_left / g_pairs)
                engine.set_param("slab_rec_ratio", g_recs / g_pairs)
            for k in TOTALS:
                tot[k] += info.get(k, 0)
            tot["steps"] += 1
            try:
                if nbytes:
                    my_sink.seek(file_pos + sum(sizes[:r]))
                    pending = submit(info["n_records"])
                file_pos += sum(sizes)
                if r == 0 and merged is not None and merged.shape[0]:
                    n_rec += int(merged.shape[0])
                    engine.lines_append_tensor(merged)
            except Exception:  # noqa: BLE001 -- the next gather_step tells the other ranks
                import traceback

                traceback.print_exc()
                status = 1
    finally:
        err = None
        try:
            if pending is not None:
                pending.result()
            if r != 0:
                my_sink.close()
        except Exception as exc:  # noqa: BLE001
            import traceback

            traceback.print_exc()
            err = exc
    agree(err is None, dev, "writing the tempfile")  # also the barrier: rank 0 returns when every piece is in the file
    tot = _sum_totals(tot, dev)
    tot["n_records"] = n_rec
    return tot


_REGION: dict = {}


def _region_begin(dev):
    import time

    import torch
    import torch.distributed as dist

    if dist.is_initialized():
        dist.barrier()
    if dev.type == "cuda":
        torch.cuda.synchronize(dev)
        _REGION["ev0"] = torch.cuda.Event(enable_timing=True)
        _REGION["ev0"].record(torch.cuda.current_stream(dev))
    _REGION["t0"] = time.perf_counter()


def _region_end(dev) -> float:
    import time

    import torch
    import torch.distributed as dist

    dev_s = 0.0
    if dev.type == "cuda":
        ev1 = torch.cuda.Event(enable_timing=True)
        ev1.record(torch.cuda.current_stream(dev))
    if dist.is_initialized():
        dist.barrier()
    if dev.type == "cuda":
        torch.cuda.synchronize(dev)
        dev_s = _REGION["ev0"].elapsed_time(ev1) / 1e3
    t = torch.tensor([max(time.perf_counter() - _REGION["t0"], dev_s)], dtype=torch.float64, device=dev)
    if dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


class Region:
    """A timed region over all ranks (measurement hook of the sharded path): barrier + synchronize on both sides,
    device time from CUDA events on the shared stream, wall time beside it, MAX over ranks.  Rank 0 calls begin() /
    end(); the workers take part from worker_loop."""

    def __init__(self, facade, dev, world: int):
        self.facade, self.dev, self.world = facade, dev, world

    def begin(self):
        if self.world > 1:
            broadcast_job({"op": "region_begin"}, self.dev)
        _region_begin(self.dev)

    def end(self) -> float:
        if self.world > 1:
            broadcast_job({"op": "region_end"}, self.dev)
        return _region_end(self.dev)


def worker_loop(engine, dev):
    """Ranks > 0: serve rank 0's jobs until it sends 'stop'."""
    import torch.distributed as dist

    facade = ShardedEngine(engine, dev)
    while True:
        job = broadcast_job(None, dev)
        op = job["op"]
        if op == "stop":
            engine.close()
            return
        if op == "upload":
            facade._guard(lambda: _upload(engine, job), "upload_reads")
        elif op == "upload_dev":
            _upload_dev(facade, engine, job, dev)
        elif op == "param":
            engine.set_param(job["name"], job["value"])
        elif op == "batch":
            _run_shard(engine, job, dev, dist.get_rank(), dist.get_world_size())
        elif op == "region_begin":
            _region_begin(dev)
        elif op == "region_end":
            _region_end(dev)
        elif op == "batch_text":
            _run_shard_text(engine, job, dev, dist.get_rank(), dist.get_world_size(), None)
