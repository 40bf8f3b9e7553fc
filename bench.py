#!/usr/bin/env python3
"""bench.py -- read-pair comparisons/sec of the all-pairs read-similarity path (BASELINE.json metric).

Workload (configs[4] of BASELINE.json, the configuration the metric is quoted on): `--all` on
100,000 synthetic ~1 kb ONT-like amplicon reads (200 templates, 6 % error, either strand, 1 % of
reads carrying N), i.e. 4,999,950,000 length-compatible pairs.  One STEP = the whole job: every
pair decided exactly as amplicon_sorter.py:776-807 decides it, records gathered on rank 0.

  python bench.py [--gpus N] [--steps K] [--warmup W]          one rank per GPU under torchrun for N > 1
  python bench.py --impl reference ...                          the CPU arm (oracle port, all host cores)

value  : pairs/s with the reads already resident in HBM (timed region = K whole jobs, barrier +
         synchronize on both sides, max over ranks).  Strong scaling: total work is fixed as N grows.
e2e    : pairs/s through the C-ABI with HOST buffers: asb_upload_reads (H2D of the ASCII reads and
         offsets) + all steps + the D2H of the merged records, per step.
roofline: the dominant kernel asb_screen is integer-ALU bound; see DESIGN.md.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from amplicon_sorter_b200 import host, synth, thresholds  # noqa: E402

_OUT = sys.stdout
METRIC = "read-pair comparisons/sec (all-vs-all, ~1 kb reads)"
UNIT = "pairs/s"
ALU_OPS_PER_WORD_UPDATE = 10  # ALU-pipe instructions per Myers word-update in asb_screen's SASS (profiles/)
NCU_TRAFFIC_BYTES_PER_LAUNCH = 48.7e6  # dram__bytes_read+write.sum of one asb_screen launch (ncu --set full, profiles/r1_asb_screen_ncu_full_v7.txt)


def make_workload(n_reads: int):
    """cfg5 reads + the host-side geometry of the one `--all` batch (amplicon_sorter.py:610, :669, :679)."""
    cache = f"/tmp/asb200_cfg5_{n_reads}.npz"
    if os.path.exists(cache):
        z = np.load(cache)
        buf, offs = z["buf"], z["offs"]
    else:
        reads, _, _ = synth.make_config(5, scale=n_reads / 100000.0)
        buf, offs = synth.pack_reads(reads)
        try:
            np.savez(cache + f".{os.getpid()}.npz", buf=buf, offs=offs)
            os.replace(cache + f".{os.getpid()}.npz", cache)
        except OSError:
            pass
    lens = (offs[1:] - offs[:-1]).astype(np.int64)
    order = np.argsort(lens, kind="stable").astype(np.uint32)
    lens_sorted = lens[order]
    hi = host.batch_geometry(lens_sorted)
    dpass, drev = thresholds.tables(0.80, int(lens.max()) + 1)
    tl = int((hi.astype(np.int64) - np.arange(hi.shape[0])).sum())
    return dict(buf=buf, offs=offs, order=order, lens_sorted=lens_sorted, hi=hi, dpass=dpass, drev=drev, tl=tl,
                n_reads=int(lens.shape[0]), mean_len=float(lens.mean()))


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
    similarity()) on all host cores, on a strided sample of the job's rows sized for ~`seconds`."""
    from oracle import oracle

    n = w["n_reads"]
    cores = oracle.host_threads() if nthreads <= 0 else nthreads
    # calibrate on a tiny strided sample, then size the real one
    stride = max(1, n // 8)
    t0 = time.perf_counter()
    _, st = oracle.process_batch(w["buf"], w["offs"], w["order"], 80.0, algo=algo, rows=(0, n, stride), nthreads=nthreads)
    dt = max(time.perf_counter() - t0, 1e-3)
    rate = st["pairs"] / dt
    want_pairs = rate * seconds
    stride = int(max(1, min(n // 2, round(w["tl"] / max(want_pairs, 1.0)))))
    t0 = time.perf_counter()
    _, st = oracle.process_batch(w["buf"], w["offs"], w["order"], 80.0, algo=algo, rows=(stride // 2, n, stride), nthreads=nthreads)
    dt = time.perf_counter() - t0
    return {"value": st["pairs"] / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"rows {stride // 2}::{stride} of the same job ({st['pairs']} pairs, {st['alignments']} alignments, {dt:.1f} s)",
            "algorithm": "oracle/asref.c edlib-like band-doubling Myers (64-bit words) + similarity() three-way rule, pthreads",
            "pairs": st["pairs"], "seconds": dt}


def run_reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = make_workload(a.reads)
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
            "config": {"workload": f"cfg5: --all on {w['n_reads']} synthetic ~1 kb reads, 200 templates; each step a strided row sample (~{per_step:.0f} s) of the {w['tl']}-pair job",
                       "reads": w["n_reads"], "pairs_full_job": w["tl"]},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"]},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), file=_OUT, flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--reads", type=int, default=100000, help="reads in the --all batch (BASELINE config 5 = 100000)")
    ap.add_argument("--cpu-seconds", type=float, default=20.0)
    ap.add_argument("--e2e-steps", type=int, default=1)
    ap.add_argument("--no-cpu", action="store_true")
    a = ap.parse_args()
    # stdout carries exactly ONE line (the JSON): everything else any library prints to fd 1 -- NCCL's
    # "NCCL version ..." banner under NCCL_DEBUG=VERSION/WARN for one -- is sent to stderr
    global _OUT
    sys.stdout.flush()
    _OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if a.impl == "reference":
        return run_reference_arm(a)

    import torch
    import torch.distributed as dist

    from amplicon_sorter_b200.engine import Engine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    w = make_workload(a.reads)
    stream = torch.cuda.current_stream()
    eng = Engine(local, stream=stream.cuda_stream)
    pinned_buf = torch.from_numpy(w["buf"]).pin_memory()
    pinned_offs = torch.from_numpy(w["offs"].view(np.int64)).pin_memory()
    h_buf, h_offs = pinned_buf.numpy(), pinned_offs.numpy().view(np.uint64)

    rec_cap = 1 << 22
    rec_dev = torch.empty((rec_cap, 4), dtype=torch.int32, device=dev)
    launches = [0]
    agg = {"pairs": 0, "word_updates": 0, "screen_word_updates": 0, "screen_ms": 0.0, "total_ms": 0.0, "screen_launches": 0, "n_records": 0}

    def job(collect=None):
        """One whole job on this rank's shard; returns the merged record tensor on rank 0 (device)."""
        nonlocal rec_dev, rec_cap
        eng.batch_begin(w["order"], w["hi"], w["dpass"], w["drev"], rank, world)
        n_rec = 0
        while True:
            info = eng.batch_step()
            if info is None:
                break
            launches[0] += info["launches"] + (1 if info["n_records"] else 0)  # screen/list kernels (+ pack); cub sorts not counted
            if collect is not None:
                for k in ("pairs", "word_updates", "screen_word_updates", "screen_ms", "total_ms", "n_records"):
                    collect[k] += info[k]
                collect["screen_launches"] += 1
            k = info["n_records"]
            if n_rec + k > rec_cap:
                rec_cap = max(2 * rec_cap, n_rec + k)
                bigger = torch.empty((rec_cap, 4), dtype=torch.int32, device=dev)
                bigger[:n_rec] = rec_dev[:n_rec]
                rec_dev = bigger
            if k:
                eng.batch_records_dev(rec_dev[n_rec:].data_ptr())
                n_rec += k
        mine = rec_dev[:n_rec]
        if world == 1:
            return mine
        # K6: per-rank lists -> rank 0 over NCCL (counts, then padded all_gather; lists are small)
        cnt = torch.tensor([n_rec], dtype=torch.int64, device=dev)
        cnts = [torch.zeros_like(cnt) for _ in range(world)]
        dist.all_gather(cnts, cnt)
        mx = int(max(int(c.item()) for c in cnts))
        pad = torch.zeros((mx, 4), dtype=torch.int32, device=dev)
        pad[:n_rec] = mine
        parts = [torch.empty_like(pad) for _ in range(world)] if rank == 0 else None
        dist.gather(pad, parts, dst=0)
        if rank != 0:
            return mine
        allr = torch.cat([p[: int(c.item())] for p, c in zip(parts, cnts)])
        key = (allr[:, 0].to(torch.int64) << 32) | allr[:, 1].to(torch.int64)
        return allr[torch.argsort(key)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- resident-input arm -------------------------------------------------------------------
    codes_bytes = 2 * int(w["buf"].nbytes)  # forward + compl_reverse symbol codes
    flush_buf = None if codes_bytes > 126 * (1 << 20) else torch.empty(192 * (1 << 20), dtype=torch.uint8, device=dev)
    eng.upload_reads(h_buf, h_offs)
    for _ in range(a.warmup):
        job()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches[0] = 0
    t0 = time.perf_counter()
    ev0.record(stream)
    for _ in range(a.steps):
        if flush_buf is not None:
            flush_buf.zero_()  # inputs smaller than L2 (reduced --reads only): evict them between timed steps
        out = job(agg)
    ev1.record(stream)
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    dev_s = ev0.elapsed_time(ev1) / 1e3
    tmax = torch.tensor([max(wall, dev_s)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    t_res = float(tmax.item())
    n_records = int(out.shape[0]) if rank == 0 else 0
    value = w["tl"] * a.steps / t_res
    gpu_launches = launches[0]

    # ---- end-to-end arm: host buffers in, merged records out, every step ------------------------
    e2e = None
    if a.e2e_steps > 0:
        # pinned destination of the result, sized from the resident arm's record count (buffer set-up, like the pinned inputs)
        host_pin = torch.empty((max(n_records, 1), 4), dtype=torch.int32).pin_memory() if rank == 0 else None
        barrier()
        t0 = time.perf_counter()
        d2h = 0
        parts_ms = {"upload": 0.0, "job_and_gather": 0.0, "d2h": 0.0}
        for _ in range(a.e2e_steps):
            ta = time.perf_counter()
            eng.upload_reads(h_buf, h_offs)
            tb = time.perf_counter()
            out = job()
            torch.cuda.synchronize()
            tc = time.perf_counter()
            if rank == 0:
                if host_pin is None or host_pin.shape[0] < out.shape[0]:
                    host_pin = torch.empty((max(out.shape[0], 1), 4), dtype=torch.int32).pin_memory()
                host_pin[: out.shape[0]].copy_(out, non_blocking=True)  # D2H of the merged records into pinned memory
                torch.cuda.synchronize()
                d2h = out.numel() * 4
            td = time.perf_counter()
            for k, v in zip(parts_ms, (tb - ta, tc - tb, td - tc)):
                parts_ms[k] += v * 1e3 / a.e2e_steps
        barrier()
        te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        h2d = (h_buf.nbytes + h_offs.nbytes + w["order"].nbytes + w["hi"].nbytes + w["dpass"].nbytes + w["drev"].nbytes) * world
        e2e = {"value": w["tl"] * a.e2e_steps / float(te.item()), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "steps": a.e2e_steps, "rank0_ms_per_step": {k: round(v, 1) for k, v in parts_ms.items()},
               "api": "asb_upload_reads + asb_batch_begin/step + asb_batch_records_dev + NCCL gather + D2H on rank 0"}

    if rank == 0:
        lop3, mix = eng.int_peak(4000)
        screen_s = agg["screen_ms"] / 1e3
        # asb_screen alone: its own word-update counter over its own launch durations (CUDA events in the library)
        wu_rate = agg["screen_word_updates"] / max(screen_s, 1e-9)
        achieved = wu_rate * ALU_OPS_PER_WORD_UPDATE / 1e12
        peak = max(lop3, mix)
        roofline = {"bound": "int_alu", "achieved": achieved, "peak": peak, "unit": "Tops/s (INT32 ALU-pipe lane-ops)",
                    "frac": achieved / peak if peak else None, "traffic": NCU_TRAFFIC_BYTES_PER_LAUNCH,
                    "kernel": "asb_screen", "launches": agg["screen_launches"],
                    "avg_launch_ms": agg["screen_ms"] / max(agg["screen_launches"], 1),
                    "kernel_share_of_step": screen_s / max(agg["total_ms"] / 1e3, 1e-9),
                    "word_updates_per_s": wu_rate, "alu_ops_per_word_update": ALU_OPS_PER_WORD_UPDATE,
                    "word_updates_per_pair": agg["word_updates"] / max(agg["pairs"], 1),
                    "peak_source": "asb_int_peak measured live on this GPU: best of LOP3-chain probe (%.2f) and LOP3/SHF/IADD3/LEA mix probe (%.2f); nominal 148 SM x 64 lanes x 1.965 GHz = 18.61" % (lop3, mix),
                    "ncu": "profiles/r1_asb_screen_ncu_full_v7.txt: sm__inst_executed_pipe_alu 91.4 % of peak, dram 48.7 MB per launch",
                    "nominal": {"ops_per_job": nominal_ops(w), "note": "SURVEY 8(d): 20*ceil(m/32)*n*2 per pair (full-matrix Myers, both strands)",
                                "equivalent_tops": nominal_ops(w) * a.steps / t_res / 1e12 / max(world, 1)},
                    "hbm": {"peak_gbs": _measured_peak("hbm_gbs"), "note": "path is not HBM-bound: ~200 MB of symbol codes stay L2-resident"}}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
                "ms_per_step": t_res / a.steps * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "u32", "data": "synthetic",
                "config": {"workload": f"cfg5: --all on {w['n_reads']} synthetic ~1 kb reads (200 templates, 6% ONT-like error, both strands, 1% with N), -sg 80; one step = the whole {w['tl']}-pair job",
                           "reads": w["n_reads"], "pairs_per_step": w["tl"], "records_per_step": n_records, "mean_read_len": w["mean_len"],
                           "l2_policy": ("inputs (2 x %.0f MB symbol codes) exceed the 126 MB L2; no flush needed" % (w["buf"].nbytes / 1e6)) if flush_buf is None
                           else "inputs fit in L2: a 192 MB buffer is overwritten between timed steps",
                           "sharding": "rows of the length-sorted batch dealt cyclically over ranks; NCCL gather of records to rank 0"},
                "clocks": clocks, "e2e": e2e, "gpu_launches": gpu_launches, "roofline": roofline}
        if not a.no_cpu:
            line["cpu_baseline"] = cpu_baseline(w, seconds=a.cpu_seconds) if world == 1 else None
        print(json.dumps(line), file=_OUT, flush=True)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


def _measured_peak(key):
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f).get(key)
    except Exception:
        return None


if __name__ == "__main__":
    main()
