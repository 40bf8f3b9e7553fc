"""CPU baseline the north star asks for: the UNMODIFIED reference script (a copy under git-ignored baseline/_ref/, or
/root/reference) on the in-repo edlib / Bio shims, timed on THIS host -- run it on the GPU box:

    python tests/perf_reference_script.py [--np N] > gpurun_out/r2_reference_script.json

B1: `-np 1` on config 1 (1,000 reads, default batches);  B2: `-np <all cores>` on config 2 (`-a`, 10,000 reads; pass
--cfg2-scale to shrink it).  Only sort_genes() runs (the all-pairs stage; --stop-after-genes).  Reported per run: the
wall time of process_list, the reference's own pair count tl (amplicon_sorter.py:684), pairs/s, and pairs/s with the
stage's FIXED sleeps subtracted (1 s :658, 5 s :767, 2 s per spool chunk :732/:744 -- SURVEY F7).  Not a product path."""
import argparse
import json
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from amplicon_sorter_b200 import synth  # noqa: E402

SCRIPTS = [os.path.join(ROOT, "baseline", "_ref", "amplicon_sorter.py"), "/root/reference/amplicon_sorter.py"]


def run(script, cfg, scale, nproc, work):
    reads, _, cli = synth.make_config(cfg, scale=scale)
    fq = os.path.join(work, f"c{cfg}.fastq")
    synth.write_fastq(fq, reads)
    cli = [a for a in cli if a not in ("-np", "1")]
    cmd = [sys.executable, os.path.join(ROOT, "oracle", "run_reference.py"), "--script", script, "--stage", "reference", "--stop-after-genes",
           "--", "-i", fq, "-o", os.path.join(work, f"out{cfg}"), "-np", str(nproc)] + cli
    out = subprocess.run(cmd, env=dict(os.environ, PYTHONHASHSEED="0"), capture_output=True, text=True)
    m = re.search(r"ASB_TIMING process_list seconds=([0-9.]+) tl=(-?\d+)", out.stdout)
    if not m:
        raise SystemExit(out.stdout[-2000:] + out.stderr[-2000:])
    secs, tl = float(m.group(1)), int(m.group(2))
    chunks = out.stdout.count("processing: file_")
    sleeps = 1 + 5 + 2 * chunks
    return {"config": cfg, "reads": len(reads), "cli": " ".join(cli + ["-np", str(nproc)]), "np": nproc, "pairs_tl": tl, "seconds": secs,
            "pairs_per_s": tl / secs, "spool_chunks": chunks, "fixed_sleep_seconds": sleeps,
            "pairs_per_s_without_fixed_sleeps": tl / max(secs - sleeps, 1e-9)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--np", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--cfg2-scale", type=float, default=1.0)
    ap.add_argument("--skip-cfg2", action="store_true")
    a = ap.parse_args()
    script = next((p for p in SCRIPTS if os.path.isfile(p)), None)
    if script is None:
        raise SystemExit("no copy of the reference script (run __graft_entry__.build() where /root/reference exists)")
    with tempfile.TemporaryDirectory() as work:
        res = {"kind": "reference+shim", "what": "unmodified amplicon_sorter.py on oracle/shims (in-repo edlib-compatible C Myers behind edlib.align, "
                                                 "minimal Bio.SeqIO), all-pairs stage only", "host_cores": os.cpu_count(),
               "B1": run(script, 1, 1.0, 1, work)}
        if not a.skip_cfg2:
            res["B2"] = run(script, 2, a.cfg2_scale, a.np, work)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
