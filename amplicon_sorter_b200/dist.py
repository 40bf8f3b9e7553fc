"""One process per GPU: shard the all-pairs stage over ranks with torch.distributed.

The pair set shards with no exchange step: every rank holds all reads and
decides the rows `p % world == rank` of the sorted batch.  The only communication is
(1) rank 0 broadcasting the job (reads + batch composition) to the workers and (2) the gather of
the compacted per-rank record lists to rank 0 -- NCCL over NVLink on GPUs, gloo in CPU tests.

    torchrun --nproc-per-node 8 -m amplicon_sorter_b200 --script amplicon_sorter.py -i reads.fastq ...

Rank 0 runs the reference script; ranks > 0 sit in `worker_loop` and serve its process_list calls.
"""
from __future__ import annotations

import os

import numpy as np

from ._ffi import RECORD


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


def _bcast_array(arr, dev, src=0):
    """Broadcast a numpy array (shape/dtype known on src only)."""
    import torch
    import torch.distributed as dist

    meta = [None]
    if dist.get_rank() == src:
        meta = [(arr.shape, arr.dtype.str)]
    dist.broadcast_object_list(meta, src=src)
    shape, dt = meta[0]
    if dist.get_rank() == src:
        t = torch.from_numpy(np.ascontiguousarray(arr).view(np.uint8).reshape(-1)).to(dev)
    else:
        t = torch.empty(int(np.prod(shape)) * np.dtype(dt).itemsize, dtype=torch.uint8, device=dev)
    if t.numel():
        dist.broadcast(t, src=src)
    return t.cpu().numpy().view(np.dtype(dt)).reshape(shape)


def broadcast_job(job: dict | None, dev):
    """Rank 0 passes {'op': ..., arrays...}; every rank returns the same dict."""
    import torch.distributed as dist

    keys = [None]
    if dist.get_rank() == 0:
        keys = [[(k, isinstance(v, np.ndarray)) for k, v in job.items()]]
    dist.broadcast_object_list(keys, src=0)
    out = {}
    for k, is_arr in keys[0]:
        if is_arr:
            out[k] = _bcast_array(job[k] if job is not None else None, dev)
        else:
            box = [job[k] if job is not None else None]
            dist.broadcast_object_list(box, src=0)
            out[k] = box[0]
    return out


def gather_records(recs: np.ndarray, dev):
    """Per-rank record arrays -> merged array in (i_pos, j_pos) order on rank 0 (None elsewhere)."""
    import torch
    import torch.distributed as dist

    w, r = dist.get_world_size(), dist.get_rank()
    cnt = torch.tensor([recs.shape[0]], dtype=torch.int64, device=dev)
    cnts = [torch.zeros_like(cnt) for _ in range(w)]
    dist.all_gather(cnts, cnt)
    sizes = [int(c.item()) for c in cnts]
    mx = max(max(sizes), 1)
    pad = torch.zeros((mx, 4), dtype=torch.int32, device=dev)
    if recs.shape[0]:
        pad[: recs.shape[0]] = torch.from_numpy(recs.view(np.uint32).reshape(-1, 4).view(np.int32)).to(dev)
    parts = [torch.empty_like(pad) for _ in range(w)] if r == 0 else None
    dist.gather(pad, parts, dst=0)
    if r != 0:
        return None
    allr = torch.cat([p[:s] for p, s in zip(parts, sizes)])
    key = (allr[:, 0].to(torch.int64) << 32) | (allr[:, 1].to(torch.int64) & 0xFFFFFFFF)
    allr = allr[torch.argsort(key, stable=True)]
    return allr.cpu().numpy().view(np.uint32).reshape(-1).view(RECORD)


class ShardedEngine:
    """Engine facade for rank 0: every compare_batch is executed by all ranks on their shard."""

    def __init__(self, engine, dev):
        self.engine, self.dev = engine, dev

    def upload_reads(self, buf, offs):
        job = broadcast_job({"op": "upload", "buf": np.ascontiguousarray(buf), "offs": np.ascontiguousarray(offs)}, self.dev)
        self.engine.upload_reads(job["buf"], job["offs"])

    def set_param(self, name, value):
        broadcast_job({"op": "param", "name": name, "value": float(value)}, self.dev)
        self.engine.set_param(name, value)

    def compare_batch(self, order, hi, dpass, drev, rank=0, world=1, fetch=True):
        import torch.distributed as dist

        job = broadcast_job({"op": "batch", "order": np.asarray(order, np.uint32), "hi": np.asarray(hi, np.uint32),
                             "dpass": np.asarray(dpass, np.uint32), "drev": np.asarray(drev, np.uint32)}, self.dev)
        return _run_shard(self.engine, job, self.dev, dist.get_rank(), dist.get_world_size())

    # Only the all-pairs stage shards.  The other stages (reads x consensuses, consensus x consensus, the consumers of the
    # tempfile) are milliseconds of work: the launcher runs them on `self.engine`, rank 0's own engine ("replicas only").
    def close(self):
        broadcast_job({"op": "stop"}, self.dev)
        self.engine.close()


def _run_shard(engine, job, dev, r, w):
    import torch
    import torch.distributed as dist

    recs, tot = engine.compare_batch(job["order"], job["hi"], job["dpass"], job["drev"], r, w)
    merged = gather_records(recs, dev)
    keys = sorted(k for k, v in tot.items() if isinstance(v, (int, float)))
    t = torch.tensor([float(tot[k]) for k in keys], dtype=torch.float64, device=dev)
    dist.all_reduce(t)  # sums; times become rank-sums (reported as such)
    tot = {k: (int(v) if isinstance(tot[k], int) else float(v)) for k, v in zip(keys, t.tolist())}
    tot["n_records"] = int(merged.shape[0]) if merged is not None else 0
    return merged, tot


def worker_loop(engine, dev):
    """Ranks > 0: serve rank 0's jobs until it sends 'stop'."""
    import torch.distributed as dist

    while True:
        job = broadcast_job(None, dev)
        op = job["op"]
        if op == "stop":
            engine.close()
            return
        if op == "upload":
            engine.upload_reads(job["buf"], job["offs"])
        elif op == "param":
            engine.set_param(job["name"], job["value"])
        elif op == "batch":
            _run_shard(engine, job, dev, dist.get_rank(), dist.get_world_size())
