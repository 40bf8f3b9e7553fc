"""Dev tool (not a test): throughput of the consumers of <stem>_compare.tmp on the lines of a
scaled BASELINE config 5 job, next to the CPU restatements (oracle/asref.c, and the statement-by-statement Python
one on a bounded prefix -- that is what the reference executes)."""
import argparse
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from amplicon_sorter_b200 import groups, synth  # noqa: E402
from amplicon_sorter_b200.engine import Engine  # noqa: E402
from oracle import oracle  # noqa: E402
from tests import util  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--scale", type=float, default=0.2)
ap.add_argument("--py-lines", type=int, default=300000)
a = ap.parse_args()
reads, _, _ = synth.make_config(5, scale=a.scale)
eng = Engine(0)
buf, offs, order, lens_sorted, hi, dpass, drev = util.batch_inputs(reads)
eng.upload_reads(buf, offs)
recs, tot = eng.compare_batch(order, hi, dpass, drev)
L = lens_sorted[recs["j_pos"]].astype(np.int64)
d = recs["d"].astype(np.int64)
tab = {}
milli = np.empty(len(recs), dtype=np.uint32)
for length in np.unique(L).tolist():
    sel = L == length
    lut = np.array([int(round(round(1 - dd / length, 3) * 1000)) for dd in range(int(d[sel].max()) + 1)], dtype=np.uint32)
    milli[sel] = lut[d[sel]]
lines = groups.Lines(order[recs["i_pos"]], order[recs["j_pos"]], milli, recs["reverse"] != 0)
out = {"reads": len(reads), "pairs": tot["pairs"], "lines": len(lines)}
t0 = time.perf_counter(); groups.upload(eng, lines); out["upload_s"] = time.perf_counter() - t0
st = {}
for rep in range(2):
    t0 = time.perf_counter(); ssg = groups.ssg_estimate(eng, lines, st); out["ssg_wall_s"] = time.perf_counter() - t0
    t0 = time.perf_counter(); templist, ta, tb, tm = groups.best_hits(eng, lines, stats=st); out["besthit_wall_s"] = time.perf_counter() - t0
    t0 = time.perf_counter(); n_greedy, grp = groups.make_groups(eng, templist, False, st); out["groups_wall_s"] = time.perf_counter() - t0
    members = grp[0]
    t0 = time.perf_counter(); tl2, *_ = groups.best_hits(eng, lines, 0.93, members, stats=None); out["read_indexes_wall_s"] = time.perf_counter() - t0
out.update(st, ssg=ssg, templist=len(templist), n_greedy=n_greedy, groups=len(grp), read_indexes_templist=len(tl2))
out["lines_per_s_besthit_device"] = len(lines) / (st["besthit_ms"] / 1e3)
out["lines_per_s_hist_device"] = len(lines) / (st["hist_ms"] / 1e3)
# CPU: C restatement on everything, Python restatement (= what the reference runs) on a prefix
t0 = time.perf_counter(); wl, wf = oracle.besthit(lines.a, lines.b, lines.milli); out["c_oracle_besthit_s"] = time.perf_counter() - t0
line, first, _ = eng.lines_besthit()
assert np.array_equal(line, wl) and np.array_equal(first, wf)
k = min(a.py_lines, len(lines))
text = "".join(f"{x}:{y}:{groups.IDEN_STR[z]}\n" for x, y, z in zip(lines.a[:k].tolist(), lines.b[:k].tolist(), lines.milli[:k].tolist()))
t0 = time.perf_counter(); oracle.py_besthit_templist(text); dt = time.perf_counter() - t0
out["python_restatement_lines_per_s"] = k / dt
t0 = time.perf_counter(); oracle.py_ssg(text); out["python_ssg_lines_per_s"] = k / (time.perf_counter() - t0)
print(json.dumps(out))
