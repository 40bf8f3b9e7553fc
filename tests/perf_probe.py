"""Dev tool (not a test): quick throughput probe of the engine on a scaled BASELINE config."""
import argparse
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from amplicon_sorter_b200 import synth  # noqa: E402
from amplicon_sorter_b200.engine import Engine  # noqa: E402
from tests import util  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--cfg", type=int, default=5)
ap.add_argument("--scale", type=float, default=0.2)
ap.add_argument("--frac", type=float, nargs="*", default=[0.62])
ap.add_argument("--push", type=int, nargs="*", default=[1])
ap.add_argument("--cont", type=int, nargs="*", default=[4])
ap.add_argument("--seed", type=int, nargs="*", default=[1])
ap.add_argument("--sg", type=float, default=80.0)
ap.add_argument("--check", action="store_true")
a = ap.parse_args()
t0 = time.time()
reads, _, _ = synth.make_config(a.cfg, scale=a.scale)
print(f"synth {len(reads)} reads in {time.time()-t0:.1f}s", flush=True)
eng = Engine(0)
buf, offs, order, lens_sorted, hi, dpass, drev = util.batch_inputs(reads, a.sg)
t0 = time.time()
eng.upload_reads(buf, offs)
print(f"upload {time.time()-t0:.3f}s")
for seed in a.seed:
 for frac in a.frac:
  for cont in a.cont:
    for push in a.push:
        eng.set_param("seed_lb", seed)
        eng.set_param("screen_frac", frac)
        eng.set_param("push_thresh", push)
        eng.set_param("cont_thresh", cont)
        for rep in range(2):
            t0 = time.time()
            recs, tot = eng.compare_batch(order, hi, dpass, drev)
            dt = time.time() - t0
        tot.update(seed=seed, frac=frac, push=push, cont=cont, wall_s=round(dt, 4), pairs_per_s=round(tot["pairs"] / dt / 1e6, 2),
                   wu_per_s_T=round(tot["word_updates"] / (tot["total_ms"] / 1e3) / 1e12, 3))
        print(json.dumps(tot), flush=True)
if a.check:
    from oracle import oracle
    want, st = oracle.process_batch(buf, offs, order, a.sg)
    util.assert_same_records(recs, want)
    print("parity ok", st)
