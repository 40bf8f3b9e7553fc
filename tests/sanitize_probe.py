"""Dev tool: small end-to-end run for compute-sanitizer (memcheck / racecheck)."""
import sys
sys.path.insert(0, ".")
import numpy as np
from amplicon_sorter_b200 import synth
from amplicon_sorter_b200.engine import Engine
from tests import util

reads, _, _ = synth.make_config(5, scale=0.003)
reads += [b"ACGTN" * 10, b"A", b"ACGT" * 600]
with Engine(0) as eng:
    got, tot = util.gpu_batch(eng, reads, pair_cap=20000)
    want, st = util.oracle_batch(reads)
    util.assert_same_records(got, want)
    a = np.arange(0, 40, dtype=np.uint32); b = a[::-1].copy()
    eng.distance_pairs(a, b); eng.distance_pairs(a, b, mode="HW")
    eng.kmer_build(6); eng.kmer_shared_tile(a, b)
print("ok", tot["pairs"], len(got))
