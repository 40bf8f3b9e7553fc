#!/usr/bin/env python3
"""bench.py -- read-pair comparisons/sec of the all-pairs read-similarity path (BASELINE.json metric).

Default workload = configs[4] of BASELINE.json, the configuration the metric is quoted on: `--all` on 100,000
synthetic ~1 kb ONT-like amplicon reads (200 templates, 6 % error, either strand, 1 % of reads carrying N), i.e.
4,999,949,987 length-compatible pairs.  `--config 1..4` runs the other BASELINE configs (parity-test shapes, not the
headline).  One STEP = the whole job: every pair decided exactly as amplicon_sorter.py:776-807 decides it and the
lines of <stem>_compare.tmp produced in the reference's order.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config C]     one rank per GPU under torchrun for N > 1
  python bench.py --impl reference ...                                   the CPU arm (oracle port, all host cores)

Both arms go through the PRODUCT's code: a single GPU through `Engine`, N > 1 through `dist.ShardedEngine` on rank 0
with the other ranks in `dist.worker_loop` (job broadcast, every slab's rows split into contiguous ranges of equal pair
counts, every rank printing and writing its own piece of the tempfile, per-slab device-side NCCL gather of the records).

value  : pairs/s with the reads already resident in HBM: `compare_text` (all slabs: screen + list kernels, sort, text
         assembled on the device and copied to pinned host memory) with the text dropped instead of written.
e2e    : pairs/s of `host.process_list(comparelist2, tempfile)` -- the drop-in for amplicon_sorter.py:647 -- from
         Python lists of records to the finished <stem>_compare.tmp on disk: string join, H2D of the reads, all slabs,
         D2H of the text, write(2).  This is the call the reference script makes.
roofline: the dominant kernel (asb_lists on clustered data, asb_screen otherwise) is integer-ALU bound;
         see DESIGN.md section 4.
"""
from __future__ import annotations

import argparse
import json
import os
import random
import shutil
import statistics
import sys
import tempfile as tempfile_mod
import threading
import time
import types
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from amplicon_sorter_b200 import host, synth, thresholds  # noqa: E402

_OUT = sys.stdout
METRIC = "read-pair comparisons/sec (all-vs-all, ~1 kb reads)"
UNIT = "pairs/s"
ALU_OPS_PER_WORD_UPDATE = 10  # ALU-pipe instructions per Myers word-update in asb_screen's SASS (profiles/)

# CRC-32 of the <stem>_compare.tmp text of the full-size job (and its line count), recorded from a 1-GPU run whose
# sampled rows matched the oracle (tests/test_gpu_fullsize.py); every later run -- any kernel version, any number of
# GPUs -- must reproduce it.  None = no constant recorded for that config yet (the run prints its CRC).
KNOWN_TEXT_CRC = {
    1: 1131388478,  # 99,424 lines (the count of the unmodified script's own run), 1,756,869 bytes
    2: 3681597654,  # 4,527,411 lines, 88,959,979 bytes
    3: 3356075816,  # 916,299 lines, 17,988,322 bytes
    4: 3352788876,  # 62,448,413 lines, 1,339,098,740 bytes
    5: 1976580653,  # 24,887,125 lines, 539,029,280 bytes (same line count as every r1 kernel version)
    6: 2667264339,  # 24,895,379 lines (recorded with --prune 0)
}  # all recorded with every pair going through the banded passes (1-5: before the pivot bound existed; 6: --prune 0)
DESCR = {
    1: "cfg1: default batch mode on 1,000 synthetic ~700 bp reads (5 templates)",
    2: "cfg2: --all on 10,000 synthetic reads, 3 genes (1.8 / 0.7 / 1.0 kb) x 4 species",
    3: "cfg3: -ra, 20 random batches of 1,000 from a 10,000-read 50-species mix (~700 bp)",
    4: "cfg4: --all on 50,000 synthetic reads, 10 loci as full (~1 kb) and nested (~870 bp) amplicons",
    5: "cfg5: --all on 100,000 synthetic ~1 kb reads (200 templates, 6% ONT-like error, both strands, 1% with N)",
    6: "cfg6 (not a BASELINE config): cfg5's shape with RELATED templates, 20 ancestors x 10 siblings at 8-15% from the ancestor",
}


def batches_of(cfg: int, n: int):
    """Read ids per batch, the way read_file builds comparelist2 for the config's command line
    (amplicon_sorter.py:568-622): default = consecutive slices of 1,000 up to maxreads (10 batches, most of them empty,
    when exactly 1,000 reads exist), -ra = random samples of 1,000 (overlapping), -a = one batch."""
    if cfg == 1:
        return [list(range(k, min(k + 1000, n))) for k in range(0, 10000, 1000)]
    if cfg == 3:
        rnd = random.Random(103)
        return [rnd.sample(range(n), min(1000, n)) for _ in range(max(1, (2 * n) // 1000))]
    return [list(range(n))]


def make_workload(cfg: int, scale: float):
    """The config's reads, its comparelist2 batches, and the engine-level geometry of the whole job."""
    cache = f"/tmp/asb200_cfg{cfg}_{scale:g}.npz"
    if os.path.exists(cache):
        z = np.load(cache)
        buf, offs = z["buf"], z["offs"]
    else:
        reads, _, _ = synth.make_config(cfg, scale=scale)
        buf, offs = synth.pack_reads(reads)
        try:
            np.savez(cache + f".{os.getpid()}.npz", buf=buf, offs=offs)
            os.replace(cache + f".{os.getpid()}.npz", cache)
        except OSError:
            pass
    n = int(offs.shape[0] - 1)
    text = buf.tobytes().decode("ascii")
    o = offs.astype(np.int64)
    seqs = [text[o[i]:o[i + 1]] for i in range(n)]
    batches = batches_of(cfg, n)
    ap = host.AllPairs.__new__(host.AllPairs)  # geometry only: no engine
    ap.lens = (o[1:] - o[:-1])
    perms, order, lens_sorted, hi, tl = ap.plan([np.asarray(b, dtype=np.int64) for b in batches if len(b)])
    dpass, drev = thresholds.tables(0.80, int(ap.lens.max()) + 1)
    tables = host.text_tables(order, lens_sorted, dpass)  # idx == read id in the bench
    return dict(cfg=cfg, buf=buf, offs=offs, seqs=seqs, batches=batches, order=order, lens_sorted=lens_sorted, hi=hi, dpass=dpass,
                drev=drev, tables=tables, tl=tl, n_reads=n, mean_len=float(ap.lens.mean()))


def comparelist2_of(w):
    """[[id, SEQ, 'u', idx], ...] batches as read_file leaves them (amplicon_sorter.py:551-561); -ra batches share records."""
    recs = [[f"r{i}", s, "u", i] for i, s in enumerate(w["seqs"])]
    return [[recs[i] for i in b] for b in w["batches"]]


def nominal_ops(w):
    """SURVEY 8(d) nominal figure: 20 * ceil(m/32) * n integer ops per alignment, 2 alignments per
    unrelated pair -- what a full-matrix Myers would execute.  Reported next to the executed count."""
    L = w["lens_sorted"].astype(np.float64)
    n = L.shape[0]
    hi = w["hi"].astype(np.int64)
    csum = np.concatenate([[0.0], np.cumsum(L)])
    tsum = csum[hi + 1] - csum[np.arange(n) + 1]  # sum of partner lengths per row
    return float((20.0 * np.ceil(L / 32.0) * tsum).sum() * 2.0)


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons of one GPU during the timed region (NVML, else nvidia-smi)."""

    BAD = {"hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40, "hw_power_brake": 0x80}
    NOTE = {"sw_power_cap": 0x4}

    def __init__(self, index: int, period=0.2):
        super().__init__(daemon=True)
        self.index, self.period, self.samples, self.reasons, self.max_mhz = index, period, [], set(), None
        self._stop_evt = threading.Event()
        self.h = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and vis.split(",")[index].isdigit() else index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.h = None

    def run(self):
        while not self._stop_evt.is_set():
            try:
                if self.h is not None:
                    self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                    r = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(self.nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                        else self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                    for name, bit in {**self.BAD, **self.NOTE}.items():
                        if r & bit:
                            self.reasons.add(name)
            except Exception:
                pass
            self._stop_evt.wait(self.period)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        med = statistics.median(self.samples) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def cpu_baseline(w, seconds=20.0, algo="edlib_like", nthreads=0):
    """The reference's CPU arithmetic (oracle/asref.c: edlib-like band-doubling Myers, three-way rule of
    similarity()) on all host cores, on a 2-D strided sample of the job's pair set sized for ~`seconds`:
    every `rs`-th row, every `cs`-th partner of it -- a few hundred rows of equal size per thread, so the
    sample scales with the thread count (a rows-only sample of this job is a dozen rows of unequal length)."""
    from oracle import oracle

    n = w["n_reads"]
    cores = oracle.host_threads() if nthreads <= 0 else nthreads

    def run(rs, cs):
        t0 = time.perf_counter()
        _, st = oracle.process_batch(w["buf"], w["offs"], w["order"], 80.0, algo=algo, rows=(rs // 2, n, rs), nthreads=nthreads, col_step=cs)
        return st, max(time.perf_counter() - t0, 1e-3)

    rs = max(1, n // (cores * 16))
    cs = max(1, int(round(w["tl"] / (rs * 20000.0 * cores))))  # calibration: ~20 k pairs per core
    st, dt = run(rs, cs)
    rate = st["pairs"] / dt
    cs = max(1, int(round(w["tl"] / rs / max(rate * seconds, 1.0))))
    st, dt = run(rs, cs)
    return {"value": st["pairs"] / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"rows {rs // 2}::{rs}, every {cs}-th partner of each, of the same job ({st['pairs']} pairs, {st['alignments']} alignments, {dt:.1f} s)",
            "algorithm": "oracle/asref.c edlib-like band-doubling Myers (64-bit words) + similarity() three-way rule, pthreads",
            "us_per_alignment_per_core": dt * cores / max(st["alignments"], 1) * 1e6,
            "pairs": st["pairs"], "seconds": dt}


def run_reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = make_workload(a.config, a.scale)
    vals, secs = [], []
    per_step = max(3.0, min(30.0, a.cpu_seconds))
    for s in range(a.warmup + a.steps):
        r = cpu_baseline(w, seconds=per_step)
        if s >= a.warmup:
            vals.append(r["value"])
            secs.append(r["seconds"])
    v = float(np.mean(vals))
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": float(np.mean(secs)) * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u64", "data": "synthetic",
            "config": {"workload": f"{DESCR[a.config]}; each step a 2-D strided sample (~{per_step:.0f} s) of the {w['tl']}-pair job",
                       "reads": w["n_reads"], "pairs_full_job": w["tl"]},
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample", "us_per_alignment_per_core")} | {"value": v},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    ref = _reference_script_record()
    if ref:
        line["cpu_baseline"]["reference_script"] = ref
    print(json.dumps(line), file=_OUT, flush=True)


def _reference_script_record():
    """Timing of the UNMODIFIED reference script (+ in-repo edlib/Bio shims) on a GPU box's host, recorded by
    tests/perf_reference_script.py (too slow to repeat in every bench run: the script sleeps >= 8 s per file)."""
    try:
        with open(os.path.join(ROOT, "profiles", "r2_reference_script.json")) as f:
            return json.load(f)
    except Exception:
        return None


def parity_check(facade, world):
    """Before anything is timed: a small job of the same kind through host.process_list on THIS engine stack (all
    ranks, NCCL gather, device-side text) must give the file the CPU oracle writes."""
    from oracle import oracle

    reads, _, _ = synth.make_config(5, scale=0.012)  # 1,200 reads, 200 templates, some with N: 0.72 M pairs
    buf, offs = synth.pack_reads(reads)
    lens = (offs[1:] - offs[:-1]).astype(np.int64)
    order = oracle.stable_length_order(lens)
    want_recs, st = oracle.process_batch(buf, offs, order, 80.0)
    want = oracle.format_lines(want_recs, order, np.arange(len(reads), dtype=np.uint32), offs)
    tmp = tempfile_mod.mkdtemp(prefix="asb200_parity_")
    try:
        args = types.SimpleNamespace(outputfolder=tmp, similar_genes=80.0)
        c2 = [[[f"r{i}", s.decode(), "u", i] for i, s in enumerate(reads)]]
        stats = {}
        facade.set_param("pair_cap", 1 << 17)  # several slabs, so that the per-slab gather and merge are exercised
        host.process_list(c2, "p_compare.tmp", args, engine=facade, stats_out=stats)
        facade.set_param("pair_cap", float(1 << 26))
        with open(os.path.join(tmp, "p_compare.tmp"), "rb") as f:
            got = f.read()
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    ok = got == want and stats["pairs"] == st["pairs"]
    return {"parity_check": "ok" if ok else "FAILED", "parity_job": f"cfg5 x 0.012 through host.process_list on {world} rank(s): {st['pairs']} pairs, "
            f"{len(want_recs)} lines, file {'identical to' if ok else 'DIFFERS from'} the CPU oracle's", "parity_crc": zlib.crc32(got)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=5, choices=[1, 2, 3, 4, 5, 6],
                    help="BASELINE.json config (default 5, the headline); 6 = config 5 with related templates (worst case for the pivot bound)")
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the config's read count (debugging)")
    ap.add_argument("--reads", type=int, default=0, help="shorthand for --scale on config 5")
    ap.add_argument("--cpu-seconds", type=float, default=20.0)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--pair-cap", type=float, default=0)
    ap.add_argument("--prune", type=int, default=1, choices=[0, 1], help="0 = switch the pivot bound off (every pair goes through the screen kernel)")
    ap.add_argument("--param", action="append", default=[], metavar="NAME=VALUE",
                    help="engine parameter (asb_set_param) for experiments, e.g. two_rows=0, class_sort=0, list_path=0, slab_pairs=2147483648")
    ap.add_argument("--tmpdir", default=None, help="where the e2e arm writes <stem>_compare.tmp (default: a fresh temp dir)")
    a = ap.parse_args()
    if a.reads:
        a.scale = a.reads / 100000.0
    # stdout carries exactly ONE line (the JSON): everything else any library prints to fd 1 -- NCCL's
    # "NCCL version ..." banner under NCCL_DEBUG=VERSION/WARN for one -- is sent to stderr
    global _OUT
    sys.stdout.flush()
    _OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if a.impl == "reference":
        return run_reference_arm(a)

    import torch

    from amplicon_sorter_b200 import dist as adist
    from amplicon_sorter_b200.engine import Engine

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    rank, world, dev = adist.init_from_env()
    stream = torch.cuda.Stream(dev)  # engine and torch.distributed share one non-default stream
    torch.cuda.set_stream(stream)
    eng = Engine(dev.index, stream=stream.cuda_stream)
    if rank != 0:
        adist.worker_loop(eng, dev)  # serves rank 0's uploads, batches and timing marks until 'stop'
        torch.distributed.destroy_process_group()
        return
    facade = adist.ShardedEngine(eng, dev) if world > 1 else eng
    region = adist.Region(facade, dev, world)

    w = make_workload(a.config, a.scale)
    parity = {} if a.no_parity else parity_check(facade, world)
    if a.pair_cap:
        facade.set_param("pair_cap", a.pair_cap)
    facade.set_param("prune", a.prune)
    for kv in a.param:
        name, _, val = kv.partition("=")
        facade.set_param(name, float(val))

    # ---- resident-input arm -------------------------------------------------------------------
    codes_bytes = 2 * int(w["buf"].nbytes)  # forward + compl_reverse symbol codes
    flush_buf = None if codes_bytes > 126 * (1 << 20) else torch.empty(192 * (1 << 20), dtype=torch.uint8, device=dev)
    facade.upload_reads(w["buf"], w["offs"])
    sink = host.NullSink()

    def job():
        return facade.compare_text(w["order"], w["hi"], w["dpass"], w["drev"], w["tables"], sink)

    for _ in range(a.warmup):
        if flush_buf is not None:
            flush_buf.zero_()  # (also loads torch's fill kernel outside the timed region)
        job()
    sampler = ClockSampler(dev.index)
    sampler.start()
    agg = None
    steps_ms = []
    region.begin()
    for _ in range(a.steps):
        if flush_buf is not None:
            flush_buf.zero_()  # inputs smaller than L2 (small configs): evict them between timed steps
        t_job = time.perf_counter()
        tot = job()
        steps_ms.append(round((time.perf_counter() - t_job) * 1e3, 2))
        host_ms = tot.pop("host_ms", None)
        agg = tot if agg is None else {k: agg[k] + tot[k] for k in agg}
    t_res = region.end()
    clocks = sampler.stop()
    value = w["tl"] * a.steps / t_res
    n_records = agg["n_records"] // a.steps
    # own kernels: pruning / screen + list kernels (info.launches, summed over ranks) + the text kernels (length + write per
    # chunk of TEXT_CHUNK records, at least one chunk per slab and rank) + at N > 1 one record pack per rank and slab and
    # the line append on rank 0; cub sorts / scans are not counted
    from amplicon_sorter_b200 import _ffi as _f
    text_chunks = agg["n_records"] // _f.TEXT_CHUNK + agg["steps"]
    gpu_launches = int(agg["launches"] + 2 * text_chunks + (0 if world == 1 else agg["steps"] + agg["steps"] // world))

    # ---- end-to-end arm: the reference-facing call, Python lists in, tempfile on disk out ----------------
    e2e, crc = None, None
    if a.e2e_steps > 0:
        tmp = a.tmpdir or tempfile_mod.mkdtemp(prefix="asb200_e2e_")
        os.makedirs(tmp, exist_ok=True)
        args = types.SimpleNamespace(outputfolder=tmp, similar_genes=80.0)
        lists = [comparelist2_of(w) for _ in range(a.e2e_steps + 1)]  # process_list sorts its batches in place: one fresh copy per step
        stats = {}
        # one untimed call first: the reference-facing path has one-time costs of its own that the resident arm's warm-up
        # does not touch (pinned staging of the scattered upload, the writer pool's first file, NCCL's first broadcast
        # of the job on the worker ranks)
        host.process_list(lists.pop(), "bench_compare_warmup.tmp", args, engine=facade, stats_out={})
        os.remove(os.path.join(tmp, "bench_compare_warmup.tmp"))
        # every step writes a file of its own, as a run of the script does (its output folder is fresh); the previous
        # step's 0.5 GB file is unlinked beside the next step, not inside it
        removers = []
        region.begin()
        for s in range(a.e2e_steps):
            host.process_list(lists[s], f"bench_compare_{s}.tmp", args, engine=facade, stats_out=stats)
            if s:
                removers.append(threading.Thread(target=os.remove, args=(os.path.join(tmp, f"bench_compare_{s - 1}.tmp"),)))
                removers[-1].start()
        t_e2e = region.end()
        for t in removers:
            t.join()
        path = os.path.join(tmp, f"bench_compare_{a.e2e_steps - 1}.tmp")
        fbytes = os.path.getsize(path)
        with open(path, "rb") as f:
            crc = 0
            while True:
                blk = f.read(1 << 26)
                if not blk:
                    break
                crc = zlib.crc32(blk, crc)
        if not a.tmpdir:
            shutil.rmtree(tmp, ignore_errors=True)
        small = w["order"].nbytes + w["hi"].nbytes + w["dpass"].nbytes + w["drev"].nbytes
        e2e = {"value": w["tl"] * a.e2e_steps / t_e2e, "unit": UNIT,
               # the read bytes go host -> device once (rank 0; the other ranks receive rank 0's device copy over NCCL), the
               # batch geometry, cut-off tables and iden string tables on every rank
               "h2d_bytes_per_step": int(w["buf"].nbytes + (w["offs"].nbytes + small + sum(np.asarray(t).nbytes if not isinstance(t, bytes) else len(t) for t in w["tables"])) * world),
               "d2h_bytes_per_step": int(fbytes), "steps": a.e2e_steps, "ms_per_step": t_e2e / a.e2e_steps * 1e3,
               "tempfile_bytes": int(fbytes), "tempfile_crc32": crc,
               "rank0_phases_ms_last_step": {k: round(v, 1) for k, v in stats.get("phases_ms", {}).items()},
               "api": "host.process_list(comparelist2, tempfile) -- the drop-in for amplicon_sorter.py:647 -- "
                      + ("on one Engine" if world == 1 else f"on dist.ShardedEngine over {world} ranks (NCCL broadcast of the job, per-slab device-side gather)")
                      + ": str join + H2D of the reads, all slabs, text assembled on the GPU, D2H, write(2) on a writer thread"}
    known = KNOWN_TEXT_CRC.get(a.config) if a.scale == 1.0 else None
    if crc is not None and known is not None:
        parity["records_crc"] = crc
        parity["records_crc_check"] = "ok" if crc == known else f"FAILED (expected {known})"
    elif crc is not None:
        parity["records_crc"] = crc

    lop3, mix = eng.int_peak(4000)
    peak = max(lop3, mix)
    screen_s = agg["screen_ms"] / 1e3 / world  # rank-sums of device times -> mean per rank
    lists_s = agg["lists_ms"] / 1e3 / world
    total_s = agg["total_ms"] / 1e3 / world
    pruned_frac = agg["pruned_pairs"] / max(agg["pairs"], 1)
    traffic = _ncu_traffic(a.config)
    # the dominant kernel: asb_screen when every pair goes through the banded pass; asb_lists (the exact passes on
    # the pairs the pivot bound leaves undecided) when most pairs are decided by asb_prune
    if pruned_frac > 0.5 or lists_s > screen_s:
        kern, k_s = "asb_lists", lists_s
        k_wu, k_useful = agg["word_updates"] - agg["screen_word_updates"], agg["useful_word_updates"] - agg["screen_useful_word_updates"]
        k_launches = 3 * agg["steps"]
    else:
        kern, k_s = "asb_screen", screen_s
        k_wu, k_useful = agg["screen_word_updates"], agg["screen_useful_word_updates"]
        k_launches = agg["steps"]
    wu_rate = k_wu / max(k_s, 1e-9) / world  # per GPU
    achieved = wu_rate * ALU_OPS_PER_WORD_UPDATE / 1e12
    useful = k_useful / max(k_wu, 1)
    roofline = {"bound": "int_alu", "achieved": achieved, "peak": peak, "unit": "Tops/s (INT32 ALU-pipe lane-ops, per GPU)",
                "frac": achieved / peak if peak else None,
                "achieved_useful": achieved * useful, "frac_useful": achieved * useful / peak if peak else None,
                "live_lane_fraction": useful,
                "traffic": traffic["bytes_per_launch"] if traffic and traffic.get("kernel") == kern else None,
                "traffic_source": traffic["source"] if traffic and traffic.get("kernel") == kern else None,
                "kernel": kern, "launches": k_launches,
                "avg_launch_ms": k_s * 1e3 * world / max(k_launches, 1) / world,
                "kernel_share_of_step": k_s / max(t_res, 1e-9),
                "kernel_share_of_device_time": k_s / max(total_s, 1e-9),
                "device_time_ms_per_step": {"asb_screen or asb_prune": screen_s * 1e3 / a.steps, "asb_lists": lists_s * 1e3 / a.steps,
                                            "all kernels of the steps (events)": total_s * 1e3 / a.steps,
                                            "pivot selection + read assignment (first step of a job)": agg["cluster_ms"] / world / a.steps},
                "host_loop_ms_last_step": {k: round(v, 1) for k, v in (host_ms or {}).items()},
                "word_updates_per_s": wu_rate, "alu_ops_per_word_update": ALU_OPS_PER_WORD_UPDATE,
                "word_updates_per_pair": agg["word_updates"] / max(agg["pairs"], 1),
                "useful_word_updates_per_pair": agg["useful_word_updates"] / max(agg["pairs"], 1),
                "pairs_decided_by_pivot_bound": pruned_frac,
                "note": "achieved counts EXECUTED lane-slots (32 lanes x active words x columns of every warp) of the dominant kernel; "
                        "achieved_useful only the words each lane itself still needed (live_lane_fraction = useful / executed, counted by the kernel)",
                "peak_source": "asb_int_peak measured live on this GPU: best of LOP3-chain probe (%.2f) and LOP3/SHF/IADD3/LEA mix probe (%.2f); nominal 148 SM x 64 lanes x 1.965 GHz = 18.61" % (lop3, mix),
                "nominal": {"ops_per_job": nominal_ops(w), "note": "SURVEY 8(d): 20*ceil(m/32)*n*2 per pair (full-matrix Myers, both strands)",
                            "equivalent_tops": nominal_ops(w) * a.steps / t_res / 1e12 / max(world, 1)},
                "hbm": {"peak_gbs": _measured_peak("hbm_gbs"), "note": "path is not HBM-bound"}}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": t_res / a.steps * 1e3, "rank0_wall_ms_of_each_step": steps_ms, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": f"{DESCR[a.config]}, -sg 80; one step = the whole {w['tl']}-pair job",
                       "reads": w["n_reads"], "pairs_per_step": w["tl"], "records_per_step": n_records, "mean_read_len": w["mean_len"],
                       "l2_policy": ("inputs (2 x %.0f MB symbol codes) exceed the 126 MB L2; no flush needed" % (w["buf"].nbytes / 1e6)) if flush_buf is None
                       else "inputs fit in L2: a 192 MB buffer is overwritten between timed steps",
                       "sharding": "every slab's rows split into `world` contiguous ranges of equal pair counts; each rank prints and writes its own piece "
                                   "of the tempfile; per-slab NCCL gather of the records to rank 0 for the resident lines (dist.gather_step)",
                       **({"params": a.param} if a.param else {})},
            "clocks": clocks, "e2e": e2e, "gpu_launches": gpu_launches, "roofline": roofline, **parity}
    if not a.no_cpu and world == 1:
        line["cpu_baseline"] = cpu_baseline(w, seconds=a.cpu_seconds)
        ref = _reference_script_record()
        if ref:
            line["cpu_baseline"]["reference_script"] = ref
    print(json.dumps(line), file=_OUT, flush=True)
    failed = parity.get("parity_check") == "FAILED" or str(parity.get("records_crc_check", "ok")).startswith("FAILED")
    facade.close()
    if world > 1:
        torch.distributed.destroy_process_group()
    if failed:
        raise SystemExit("bench.py: PARITY FAILED -- the numbers above are void")


def _measured_peak(key):
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f).get(key)
    except Exception:
        return None


def _ncu_traffic(cfg):
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE asb_screen launch of this config's full-size job, from an
    `ncu --set full` capture (profiles/r2_traffic.json records the command and the report it came from)."""
    try:
        with open(os.path.join(ROOT, "profiles", "r2_traffic.json")) as f:
            return json.load(f).get(f"cfg{cfg}")
    except Exception:
        return None


if __name__ == "__main__":
    main()
