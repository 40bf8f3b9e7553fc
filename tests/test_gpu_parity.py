"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on identical inputs.

Bar: bit-exact.  Edit distances, pass/fail decisions, the reverse flag and the record order are
integers and must match the oracle exactly."""
import numpy as np
import pytest

from amplicon_sorter_b200 import synth, thresholds
from oracle import oracle
from tests import util

pytestmark = pytest.mark.gpu


def test_encode_roundtrip_and_compl_reverse(engine):
    reads = [b"ACGT", b"A", b"AACGTNRYKMSWBDHV-X", b"N" * 33, b"ACGT" * 300 + b"ACG"]
    buf, offs = synth.pack_reads(reads)
    engine.upload_reads(buf, offs)
    for r, s in enumerate(reads):
        assert engine.debug_read(r, 0) == s
        assert engine.debug_read(r, 1) == oracle.compl_reverse(s)


def test_exact_distance_known_and_boundaries(engine):
    rng = np.random.default_rng(21)
    al = np.frombuffer(b"ACGT", dtype=np.uint8)
    reads = [b"A", b"C", b"ACGT", b"AGGT", b"ACGGT", b"AGT", b"AAAA", b"TTTT", b"KITTEN", b"SITTING", b"ACGTN", b"ACGTA"]
    for m in (31, 32, 33, 63, 64, 65, 95, 96, 97, 1023, 1024, 1025, 1500):
        c = rng.integers(0, 4, m).astype(np.uint8)
        a = al[c]
        reads += [a.tobytes(), a[1:].tobytes(), al[synth.mutate(rng, c)].tobytes(), al[rng.integers(0, 4, m)].tobytes()]
    buf, offs = synth.pack_reads(reads)
    engine.upload_reads(buf, offs)
    n = len(reads)
    a = rng.integers(0, n, 600).astype(np.uint32)
    b = rng.integers(0, n, 600).astype(np.uint32)
    a[:12], b[:12] = np.arange(12), np.roll(np.arange(12), 1)
    got = engine.distance_pairs(a, b)
    want = oracle.distance_pairs(buf, offs, a, b, algo="dp")
    assert np.array_equal(got, want)
    # reverse strand: distance to compl_reverse of the longer read
    strand = np.ones(a.shape[0], dtype=np.uint8)
    got_r = engine.distance_pairs(a, b, strand)
    for p in range(0, 600, 7):
        x, y = reads[a[p]], reads[b[p]]
        if len(x) > len(y):
            x, y = y, x
        assert got_r[p] == oracle.nw(x, oracle.compl_reverse(y), "myers")


@pytest.mark.parametrize("cfg,scale", [(1, 0.3), (2, 0.06), (3, 0.05), (4, 0.016), (5, 0.012)])
def test_batch_parity_on_baseline_configs(engine, cfg, scale):
    reads, _, _ = synth.make_config(cfg, scale=scale)
    got, tot = util.gpu_batch(engine, reads)
    want, st = util.oracle_batch(reads)
    assert tot["pairs"] == st["pairs"]
    util.assert_same_records(got, want)
    assert st["records"] > 0


@pytest.mark.parametrize("sg", [50.0, 65.0, 80.0, 90.0, 97.0, 100.0])
def test_batch_parity_thresholds(engine, sg):
    rng = np.random.default_rng(int(sg))
    reads = util.random_reads(rng, 260, 300, 420, families=6, err=0.10)
    reads += util.random_reads(rng, 40, 300, 420)                        # unrelated
    reads += [reads[0], reads[0], reads[1][:-1], oracle.compl_reverse(reads[2])]  # identical / near-identical / exact RC
    got, tot = util.gpu_batch(engine, reads, sg)
    want, st = util.oracle_batch(reads, sg)
    assert tot["pairs"] == st["pairs"]
    util.assert_same_records(got, want)


def test_batch_parity_non_acgt_and_ragged_lengths(engine):
    rng = np.random.default_rng(33)
    reads = util.random_reads(rng, 120, 1, 70, alphabet=b"ACGTN", families=0)
    reads += util.random_reads(rng, 80, 30, 40, families=2, err=0.05)
    reads += [b"A", b"A", b"N", b"ACGTRYKMSWN" * 3, oracle.compl_reverse(b"ACGTRYKMSWN" * 3), b"ACGTRYKMSWN" * 3]
    got, tot = util.gpu_batch(engine, reads, 70.0)
    want, st = util.oracle_batch(reads, 70.0)
    assert tot["pairs"] == st["pairs"]
    util.assert_same_records(got, want)


def test_long_reads_use_wide_and_dynamic_windows(engine):
    rng = np.random.default_rng(34)
    reads = util.random_reads(rng, 24, 2400, 2500, families=2, err=0.06)     # 18S-like
    reads += util.random_reads(rng, 10, 7000, 7200, families=1, err=0.06)    # needs the dynamic window at sg=50
    for sg in (80.0, 50.0):
        got, tot = util.gpu_batch(engine, reads, sg)
        want, st = util.oracle_batch(reads, sg)
        assert tot["pairs"] == st["pairs"]
        util.assert_same_records(got, want)


def test_screen_parameters_do_not_change_results(engine):
    """Early-termination / survivor routing is an optimisation: any setting gives identical records."""
    reads, _, _ = synth.make_config(1, scale=0.25)
    want, _ = util.oracle_batch(reads)
    for frac, push, cont in ((0.2, 0, 0), (0.4, 31, 32), (1.0, 0, 14), (0.62, 3, 32), (0.05, 8, 1), (0.5, 1, 14)):
        got, _ = util.gpu_batch(engine, reads, screen_frac=frac, push_thresh=push, cont_thresh=cont)
        util.assert_same_records(got, want)


def test_small_pair_cap_multi_step_and_sharding(engine):
    reads, _, _ = synth.make_config(2, scale=0.04)
    want, st = util.oracle_batch(reads)
    got, tot = util.gpu_batch(engine, reads, pair_cap=2048)
    assert tot["steps"] > 3 and tot["pairs"] == st["pairs"]
    util.assert_same_records(got, want)
    for world in (2, 3, 8):
        got, tot = util.gpu_batch(engine, reads, world=world, pair_cap=5000)
        assert tot["pairs"] == st["pairs"]  # the shards partition the pair set
        util.assert_same_records(got, want)
    engine.set_param("pair_cap", float(1 << 26))


def _edge_pairs(rng, L, sg):
    """Engineer reads whose forward/reverse distances sit exactly on the integer cut-offs."""
    dp, dr = thresholds.tables(sg / 100, L + 64)
    al = np.frombuffer(b"ACGT", dtype=np.uint8)
    q = al[rng.integers(0, 4, L)]
    reads = [q.tobytes()]
    want_fwd = [int(dp[L]) - 1, int(dp[L]), int(dp[L]) + 1, int(dp[L]) + 2]
    # forward edge: substitutions at distinct spaced positions, adjusted by the oracle
    for target in want_fwd:
        t = q.copy()
        pos = rng.permutation(L)
        k = 0
        while oracle.nw(q.tobytes(), t.tobytes(), "myers") < target:
            t[pos[k]] = al[(np.where(al == t[pos[k]])[0][0] + 1 + rng.integers(0, 3)) % 4]
            k += 1
        reads.append(t.tobytes())
    # zone edge: T with d(q, rc(T)) small but d(q, T) exactly drev-1 / drev (needs a near-palindromic q)
    return reads, dp, dr


def test_threshold_edges_forward(engine):
    rng = np.random.default_rng(35)
    for L, sg in ((700, 80.0), (1024, 80.0), (333, 93.0)):
        reads, dp, dr = _edge_pairs(rng, L, sg)
        got, _ = util.gpu_batch(engine, reads, sg)
        want, _ = util.oracle_batch(reads, sg)
        util.assert_same_records(got, want)
        row0 = want[want["i_pos"] == 0]
        assert dp[L] in row0["d"].tolist() and dp[L] + 1 not in row0["d"].tolist()


def test_dead_zone_edges(engine):
    """Pairs whose compl_reverse passes, with the forward distance walked across drev-1 / drev:
    the record must appear iff d_fwd >= drev (the reference's `elif iden < 0.5`, AS:794)."""
    rng = np.random.default_rng(36)
    al = np.frombuffer(b"ACGT", dtype=np.uint8)
    comp = {65: 84, 84: 65, 67: 71, 71: 67}
    L, sg = 400, 80.0
    dp, dr = thresholds.tables(sg / 100, L + 8)
    found = {int(dr[L]) - 2: 0, int(dr[L]) - 1: 0, int(dr[L]): 0, int(dr[L]) + 1: 0}
    reads = []
    for attempt in range(60):
        # q = Y + M + rc(Y): rc(q) = Y + rc(M) + rc(Y), so d(q, rc(q)) = d(M, rc(M)), tunable via |M|
        mid = int(rng.integers(int(0.90 * L), L - 2)) & ~1
        y = al[rng.integers(0, 4, (L - mid) // 2)]
        M = al[rng.integers(0, 4, mid)]
        rcy = np.array([comp[c] for c in y[::-1]], dtype=np.uint8)
        q = np.concatenate([y, M, rcy])
        t = np.frombuffer(oracle.compl_reverse(q.tobytes()), dtype=np.uint8).copy()  # rc(t) == q: reverse passes with d=0
        d = oracle.nw(q.tobytes(), t.tobytes(), "myers")
        # walk single substitutions in the middle until d_fwd hits a wanted value
        for _ in range(400):
            if d in found and found[d] < 3:
                found[d] += 1
                reads += [q.tobytes(), t.tobytes()]
                break
            p = int(rng.integers(len(y), len(y) + mid))
            t2 = t.copy()
            t2[p] = al[rng.integers(0, 4)]
            d2 = oracle.nw(q.tobytes(), t2.tobytes(), "myers")
            goal = min(found, key=lambda v: (found[v] >= 3, abs(v - d)))
            if abs(d2 - goal) <= abs(d - goal):
                t, d = t2, d2
        if all(v >= 2 for v in found.values()):
            break
    assert all(v >= 1 for v in found.values()), found
    got, tot = util.gpu_batch(engine, reads, sg)
    want, st = util.oracle_batch(reads, sg)
    util.assert_same_records(got, want)
    assert tot["zone_checks"] >= sum(found.values())
    assert (want["reverse"] == 1).sum() >= 2  # some dead-zone pairs emit, some are suppressed


@pytest.mark.parametrize("k", [2, 5, 6, 8])
def test_kmer_side_output_matches_its_oracle(engine, k):
    """K2/K3 have no reference counterpart: checked against oracle/asref.c's restatement."""
    rng = np.random.default_rng(50 + k)
    reads = util.random_reads(rng, 70, 40, 400, families=4, err=0.08) + util.random_reads(rng, 10, 1, 30, alphabet=b"ACGTN")
    reads += [b"A", b"ACGT" * 50, oracle.compl_reverse(reads[0])]
    buf, offs = synth.pack_reads(reads)
    engine.upload_reads(buf, offs)
    engine.kmer_build(k)
    n = len(reads)
    a = rng.integers(0, n, 300).astype(np.uint32)
    b = rng.integers(0, n, 300).astype(np.uint32)
    got = engine.kmer_shared_pairs(a, b)
    want = np.array([oracle.kmer_shared(reads[x], reads[y], k) for x, y in zip(a, b)], dtype=np.uint32)
    assert np.array_equal(got, want)
    rows = rng.permutation(n)[:45].astype(np.uint32)
    cols = rng.permutation(n)[:70].astype(np.uint32)
    tile = engine.kmer_shared_tile(rows, cols)
    for r in range(0, 45, 4):
        for c in range(0, 70, 5):
            assert tile[r, c] == oracle.kmer_shared(reads[rows[r]], reads[cols[c]], k)
    # same template, opposite strands share almost everything; unrelated reads share little (k=8)
    if k == 8:
        assert engine.kmer_shared_pairs([0], [n - 1])[0] == oracle.kmer_shared(reads[0], reads[0], k)


def test_degenerate_batches(engine):
    """Empty, singleton and no-partner batches: no step runs, no records (AS:673 loops over range(0,-1))."""
    from amplicon_sorter_b200 import thresholds as th
    dpass, drev = th.tables(0.8, 500)
    buf, offs = synth.pack_reads([b"ACGT" * 80, b"ACGT" * 100])  # 320 vs 400: outside the 5 % window
    engine.upload_reads(buf, offs)
    for order, hi in (([], []), ([0], [0]), ([0, 1], [0, 1])):
        recs, tot = engine.compare_batch(np.array(order, np.uint32), np.array(hi, np.uint32), dpass, drev)
        assert len(recs) == 0 and tot["pairs"] == 0 and tot["steps"] == 0


def test_large_alphabet_uses_fewer_warps_per_block(engine):
    """edlib semantics: every distinct byte is its own symbol.  60 symbols x long reads still fit (fewer warps/block)."""
    rng = np.random.default_rng(91)
    al = np.frombuffer(bytes(range(48, 108)), dtype=np.uint8)
    base = [al[rng.integers(0, al.size, 900)] for _ in range(3)]
    reads = []
    for i in range(60):
        r = base[i % 3].copy()
        pos = rng.integers(0, 900, 40)
        r[pos] = al[rng.integers(0, al.size, 40)]
        reads.append(r.tobytes())
    got, tot = util.gpu_batch(engine, reads)
    want, st = util.oracle_batch(reads)
    assert tot["pairs"] == st["pairs"]
    util.assert_same_records(got, want)
    assert len(want) > 300


def test_seed_bound_is_only_an_optimisation(engine):
    """The q-mer seed lower bound (myers_band.cuh::SeedLB) may only end hopeless alignments earlier:
    identical records with and without it, and fewer word-updates with it."""
    try:
        for cfg, scale, sg in ((5, 0.012, 80.0), (1, 0.25, 80.0), (4, 0.012, 96.0), (3, 0.03, 90.0)):
            reads, _, _ = synth.make_config(cfg, scale=scale)
            want, _ = util.oracle_batch(reads, sg)
            off, t0 = util.gpu_batch(engine, reads, sg, seed_lb=0)
            on, t1 = util.gpu_batch(engine, reads, sg, seed_lb=1)
            util.assert_same_records(off, want)
            util.assert_same_records(on, want)
            assert t1["word_updates"] <= t0["word_updates"]
            if cfg == 5:
                assert t1["word_updates"] < 0.8 * t0["word_updates"], (t0["word_updates"], t1["word_updates"])
    finally:
        engine.set_param("seed_lb", 1)


def test_seed_bound_adversarial_edges(engine):
    """Pairs built so that the bound is TIGHT: every edit destroys a different seed of the target and the
    distance sits exactly on / one above the cut-off.  d == dpass must still be emitted (admissibility)."""
    rng = np.random.default_rng(77)
    al = np.frombuffer(b"ACGT", dtype=np.uint8)
    comp = {65: 84, 84: 65, 67: 71, 71: 67}
    q_len = 7
    for L, sg in ((994, 96.0), (700, 90.0), (1001, 80.0), (350, 97.0)):
        dp, _ = thresholds.tables(sg / 100, L + 64)
        k = int(dp[L])
        q = al[rng.integers(0, 4, L)]
        reads = [q.tobytes()]
        n_seeds = L // q_len
        for extra in (-1, 0, 1, 2):
            for kind in ("sub", "mixed"):
                t = q.copy()
                want_d = k + extra - (2 if kind == "mixed" else 0)  # the grid shift below costs 2 edits
                seeds = rng.permutation(n_seeds)
                # one substitution inside each of `want_d` different seeds (wraps around if k > #seeds)
                for e in range(want_d):
                    s = int(seeds[e % n_seeds])
                    p = s * q_len + (e // n_seeds * 3 + int(rng.integers(0, 2))) % q_len
                    t[p] = al[(int(np.where(al == t[p])[0][0]) + 1 + int(rng.integers(0, 3))) % 4]
                tb = t.tobytes()
                if kind == "mixed":  # shift the seed grid: delete the first base, append one
                    tb = tb[1:] + b"A"
                reads.append(tb)
                reads.append(oracle.compl_reverse(tb))             # same pair through the ':reverse' branch
                reads.append(tb[: L // 2] + b"N" + tb[L // 2 + 1:])  # a seed with a non-ACGT symbol
        reads += util.random_reads(rng, 20, L - 10, L + 10)
        want, _ = util.oracle_batch(reads, sg)
        got, _ = util.gpu_batch(engine, reads, sg, seed_lb=1)
        util.assert_same_records(got, want)
        assert len(want) > 8 and k - 1 in want["d"].tolist() and k in want["d"].tolist()


def _force_prune(engine, on=True):
    engine.set_param("prune", 1 if on else 0)
    engine.set_param("prune_min_reads", 0 if on else 1024)
    engine.set_param("prune_min_pairs", 0 if on else float(1 << 22))


@pytest.mark.parametrize("cfg,scale", [(5, 0.02), (4, 0.03), (2, 0.08), (3, 0.1), (1, 1.0)])
def test_cluster_pruning_is_only_an_optimisation(engine, cfg, scale):
    """The pivot bound (asb_prune) decides pairs without aligning them; records must not depend on it.  Forced on
    for these small inputs; cfg5 / cfg4 have unrelated amplicons (most pairs pruned), cfg2 / cfg3 / cfg1 are one
    family per length window (nothing to prune: every pair goes through the class-sorted list passes, or -- with
    list_path 0, see test_list_pass_layouts_do_not_change_results -- back to the screen kernel)."""
    reads, _, _ = synth.make_config(cfg, scale=scale)
    want, st = util.oracle_batch(reads)
    try:
        _force_prune(engine, True)
        got, tot = util.gpu_batch(engine, reads, pair_cap=1 << 16)
        assert tot["pairs"] == st["pairs"]
        util.assert_same_records(got, want)
        if cfg in (5, 4):
            assert tot["pruned_pairs"] > 0.5 * tot["pairs"], tot
        parts = util.gpu_batch(engine, reads, world=3, pair_cap=1 << 16)[0]
        util.assert_same_records(parts, want)
        _force_prune(engine, False)
        engine.set_param("prune", 0)
        off, tot0 = util.gpu_batch(engine, reads, pair_cap=1 << 16)
        util.assert_same_records(off, want)
        assert tot0["pruned_pairs"] == 0
    finally:
        _force_prune(engine, False)
        engine.set_param("prune", 1)
        engine.set_param("pair_cap", float(1 << 26))


def test_cluster_pruning_adversarial_boundaries(engine):
    """Pairs that sit right on the cut-off across cluster boundaries: families whose centres are 1.0-1.6 x dpass apart
    (the bound D - a - b is close to k), reads on both strands, reverse-complement palindromes (a read that is its own
    compl_reverse has BOTH orientations close to its pivot) and a low-complexity family."""
    rng = np.random.default_rng(99)
    al = np.frombuffer(b"ACGT", dtype=np.uint8)
    L = 400
    base = rng.integers(0, 4, L, dtype=np.uint8)
    centres = [base]
    for frac in (0.05, 0.10, 0.16, 0.20, 0.24, 0.30):
        centres.append(synth.diverge(rng, base, frac))
    half = rng.integers(0, 4, L // 2, dtype=np.uint8)
    centres.append(np.concatenate([half, synth.revcomp_codes(half)]))          # palindrome: equal to its compl_reverse
    centres.append(np.tile(np.array([0, 3], dtype=np.uint8), L // 2))           # ATAT...: low complexity, self-complementary
    centres += [rng.integers(0, 4, L, dtype=np.uint8) for _ in range(6)]       # unrelated
    reads = []
    for c in centres:
        for _ in range(40):
            r = synth.mutate(rng, c, sub=0.02, ins=0.01, dele=0.01)
            if rng.random() < 0.5:
                r = synth.revcomp_codes(r)
            reads.append(al[r].tobytes())
    want, st = util.oracle_batch(reads)
    try:
        _force_prune(engine, True)
        for sg in (80.0, 90.0):
            w, _ = util.oracle_batch(reads, sg)
            got, tot = util.gpu_batch(engine, reads, sg, pair_cap=1 << 14)
            util.assert_same_records(got, w)
            assert tot["pruned_pairs"] > 0
    finally:
        _force_prune(engine, False)
        engine.set_param("pair_cap", float(1 << 26))
    assert (want["reverse"] == 1).any() and st["records"] > 1000


@pytest.mark.parametrize("cfg,scale", [(5, 0.02), (2, 0.08), (3, 0.1), (6, 0.02)])
def test_list_pass_layouts_do_not_change_results(engine, cfg, scale):
    """two_rows (a list warp takes the pairs of two queries at a time), class_sort (a row's entries ordered by the
    cluster class of their target), prune_rows (the pivot bound skipping whole clusters per row instead of looking at
    every pair) and list_path (clustered but unprunable data through the list passes instead of
    the screen kernel) only change WHICH lanes share a warp: every combination gives the oracle's records, also
    sharded and with many small slabs."""
    reads, _, _ = synth.make_config(cfg, scale=scale)
    reads = reads + [b"", reads[0][:40], reads[1] + reads[2]]  # an empty read, a very short and a very long one ride along
    want, st = util.oracle_batch(reads)
    try:
        _force_prune(engine, True)
        pruned = set()
        for two, cs, lp, pr in [(1, 1, 1, 1), (0, 1, 1, 1), (1, 0, 1, 1), (1, 1, 0, 1), (0, 0, 0, 1), (1, 1, 1, 0), (0, 0, 0, 0)]:
            got, tot = util.gpu_batch(engine, reads, pair_cap=1 << 15, two_rows=two, class_sort=cs, list_path=lp, prune_rows=pr)
            assert tot["pairs"] == st["pairs"]
            util.assert_same_records(got, want)
            if lp == 1 and cs == 1:  # (without them, unprunable data goes back to the screen kernel and reports 0)
                pruned.add(tot["pruned_pairs"])
        assert len(pruned) == 1, pruned  # asb_prune_rows (clusters skipped per row) leaves exactly the pairs asb_prune leaves
        parts = util.gpu_batch(engine, reads, world=3, pair_cap=1 << 15, two_rows=1, class_sort=1, list_path=1)[0]
        util.assert_same_records(parts, want)
    finally:
        _force_prune(engine, False)
        for k in ("two_rows", "class_sort", "list_path", "prune_rows"):
            engine.set_param(k, 1)
        engine.set_param("pair_cap", float(1 << 26))


def test_scattered_upload_equals_contiguous_upload(engine):
    """asb_upload_reads_scattered (reads gathered from separate host buffers through pinned staging, several threads,
    4 MB pieces) must leave the same codes as asb_upload_reads: reads that straddle piece and thread boundaries,
    empty reads, and the Python host's record walker as the source of the pointers."""
    from amplicon_sorter_b200 import pyhost

    rng = np.random.default_rng(5)
    al = np.frombuffer(b"ACGTN", dtype=np.uint8)
    lens = rng.integers(2000, 4000, 6000)
    lens[[0, 17, 5999]] = 0
    seqs = [al[rng.integers(0, 5, int(n))].tobytes().decode() for n in lens]
    recs = [[f"r{i}", s, "u", 3 * i] for i, s in enumerate(seqs)]
    keys, ptrs, ln = pyhost.collect(recs)
    assert keys.tolist() == [3 * i for i in range(len(recs))] and ln.tolist() == lens.tolist()
    engine.upload_reads_scattered(ptrs, ln)
    total = np.concatenate([[0], np.cumsum(lens)])
    edges = set()
    for b in range(0, int(total[-1]), 1 << 20):  # every read holding a multiple of 1 MB: all piece / thread boundaries
        edges.add(int(np.searchsorted(total, b, side="right") - 1))
    pick = sorted(edges | set(rng.integers(0, len(seqs), 150).tolist()) | {0, 1, 16, 17, 18, 5998, 5999})
    for r in pick:
        r = min(r, len(seqs) - 1)
        if not seqs[r]:
            continue  # (empty reads are stepped over by the gather; their neighbours 1, 16, 18, 5998 are checked)
        assert engine.debug_read(r, 0) == seqs[r].encode(), r
        assert engine.debug_read(r, 1) == oracle.compl_reverse(seqs[r].encode()), r
