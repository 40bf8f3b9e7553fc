"""Dev tool (not a test): throughput of the "next" rows -- reads x consensuses (asb_threeway_pairs) and
consensus x consensus (asb_distance_pairs, HW mode) -- next to the CPU oracle on a bounded sample."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from amplicon_sorter_b200 import synth, thresholds  # noqa: E402
from amplicon_sorter_b200.engine import Engine  # noqa: E402
from oracle import oracle  # noqa: E402

rng = np.random.default_rng(5)
ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
G, M, L = 200, 50000, 1000
T = [synth.random_template(rng, L) for _ in range(G)]
reads = []
for i in range(M):
    r = synth.mutate(rng, T[i % G])
    if rng.random() < 0.5:
        r = synth.revcomp_codes(r)
    reads.append(ACGT[r].tobytes())
cons = [ACGT[t].tobytes() for t in T]
buf, offs = synth.pack_reads(reads + cons)
eng = Engine(0)
eng.upload_reads(buf, offs)
# every read against every consensus (lengths all within the 5 % window): M x G pairs, consensus = DP query
q = np.repeat(np.arange(M, M + G, dtype=np.uint32), M)
t = np.tile(np.arange(M, dtype=np.uint32), G)
dpass, drev = thresholds.tables(0.95 - 0.01, 1200)
for rep in range(2):
    t0 = time.time()
    recs, info = eng.threeway_pairs(q, t, dpass, drev)
    dt = time.time() - t0
out = {"stage": "reads x consensuses (similarity_species rule, cut 0.94)", "pairs": int(q.shape[0]), "records": int(len(recs)),
       "wall_s": round(dt, 3), "gpu_ms": round(info["total_ms"], 1), "pairs_per_s": round(q.shape[0] / dt)}
sub = rng.choice(q.shape[0], 20000, replace=False)
t0 = time.time()
d = oracle.distance_pairs(buf, offs, q[sub], t[sub], algo="edlib_like")
out["cpu_pairs_per_s_fwd_only"] = round(sub.shape[0] / (time.time() - t0))
out["cpu_threads"] = oracle.host_threads()
print(json.dumps(out), flush=True)

# consensus x consensus, HW mode, both strands (iden_consensus): G^2/2 pairs
a, b = np.triu_indices(G, 1)
a = (a + M).astype(np.uint32)
b = (b + M).astype(np.uint32)
n = a.shape[0]
strand = np.concatenate([np.zeros(n, np.uint8), np.ones(n, np.uint8)])
for rep in range(2):
    t0 = time.time()
    dd = eng.distance_pairs(np.concatenate([a, a]), np.concatenate([b, b]), strand, mode="HW")
    dt = time.time() - t0
out = {"stage": "consensus x consensus (HW, both strands)", "alignments": int(2 * n), "wall_s": round(dt, 3), "alignments_per_s": round(2 * n / dt)}
sub = rng.choice(n, 300, replace=False)
t0 = time.time()
want = oracle.distance_pairs(buf, offs, a[sub], b[sub], mode="HW")
out["cpu_alignments_per_s_plain_dp"] = round(sub.shape[0] / (time.time() - t0))
assert np.array_equal(dd[:n][sub], want)
print(json.dumps(out), flush=True)
