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
    # consumers of the tempfile: histogram, best-hit filter (with and without the read_indexes admission test), components
    from amplicon_sorter_b200 import groups
    from oracle import oracle
    rng = np.random.default_rng(3)
    la = rng.integers(0, 300, 20000).astype(np.uint32)
    lb = ((la + 1 + rng.integers(0, 299, 20000)) % 300).astype(np.uint32)
    lm = (900 + 10 * rng.integers(0, 11, 20000)).astype(np.uint32)
    eng.lines_upload(la, lb, lm)
    eng.lines_hist()
    member = groups.member_bitmap({str(v) for v in range(0, 300, 3)}, 300)
    for mm, mb in ((0, None), (930, member)):
        line, first, _ = eng.lines_besthit(mm, mb)
        wl, wf = oracle.besthit(la, lb, lm, mm, mb)
        assert np.array_equal(line, wl) and np.array_equal(first, wf)
    label, _ = eng.components(la, lb, 300)
    assert np.array_equal(label, oracle.components(la, lb, 300))
print("ok", tot["pairs"], len(got))
