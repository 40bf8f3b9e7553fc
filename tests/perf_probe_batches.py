"""Dev tool (not a test): default-mode shape (BASELINE config 3: 20 random batches of 1,000 reads) through the host
stage -- all batches as one engine batch (host.process_list) vs one engine batch per 1,000 reads."""
import json
import os
import random
import sys
import tempfile
import time
import types

import numpy as np

sys.path.insert(0, ".")
from amplicon_sorter_b200 import host, synth  # noqa: E402
from amplicon_sorter_b200.engine import Engine  # noqa: E402

reads, _, _ = synth.make_config(3, scale=1.0)
recs = [[f"r{i}", s.decode(), "u", i] for i, s in enumerate(reads)]
rnd = random.Random(0)
batches = [rnd.sample(recs, 1000) for _ in range(20)]
out = tempfile.mkdtemp()
open(os.path.join(out, "results.txt"), "w").close()
args = types.SimpleNamespace(outputfolder=out, similar_genes=80.0, nprocesses=1)
eng = Engine(0)
res = {}
for rep in range(3):
    b2 = [list(b) for b in batches]
    st = {}
    t0 = time.perf_counter()
    host.process_list(b2, os.path.join(out, "x_compare.tmp"), args, engine=eng, stats_out=st)
    res["one_engine_batch_s"] = time.perf_counter() - t0
res.update(pairs=st["pairs"], records=st["records"], gpu_ms=st["gpu_ms"])
text_all = open(os.path.join(out, "x_compare.tmp")).read()
# the same work, one engine batch per 1,000 reads (the previous host path)
ap = host.AllPairs(eng)
idx_to_rid = {r[3]: t for t, r in enumerate(recs)}
ap.upload([r[1] for r in recs])
for rep in range(3):
    t0 = time.perf_counter()
    parts = []
    gpu_ms0 = ap.stats["gpu_ms"]
    for b in batches:
        rids = np.array([idx_to_rid[r[3]] for r in b], dtype=np.int64)
        perm, order, lens_sorted, rr, tl = ap.compare(rids, 80.0)
        parts.append(host.format_records(rr, order.astype(np.int64), lens_sorted, ap.last_dpass))
    res["per_batch_s"] = time.perf_counter() - t0
    res["per_batch_gpu_ms"] = ap.stats["gpu_ms"] - gpu_ms0
assert "".join(parts) == text_all
res["pairs_per_s_one"] = res["pairs"] / res["one_engine_batch_s"]
res["pairs_per_s_per_batch"] = res["pairs"] / res["per_batch_s"]
print(json.dumps(res))
