"""CPU: pin the oracle (oracle/asref.c) to the mathematical definitions and to the reference's own
Python expressions.  No GPU, no /root/reference needed."""
import numpy as np
import pytest

from amplicon_sorter_b200 import host, synth, thresholds
from oracle import oracle
from tests import util

KNOWN = [  # (a, b, NW distance) -- hand-checkable
    ("", "", 0), ("A", "", 1), ("", "ACGT", 4), ("A", "A", 0), ("A", "C", 1), ("ACGT", "ACGT", 0),
    ("ACGT", "AGGT", 1), ("ACGT", "ACGGT", 1), ("ACGT", "AGT", 1), ("AAAA", "TTTT", 4),
    ("kitten".upper(), "sitting".upper(), 3), ("ACGTN", "ACGTN", 0), ("ACGTN", "ACGTA", 1), ("NNNN", "ACGT", 4),
    ("GATTACA", "GCATGCT", 4),
]


@pytest.mark.parametrize("a,b,d", KNOWN)
def test_known_answers(a, b, d):
    for algo in ("dp", "myers", "edlib_like"):
        assert oracle.nw(a.encode(), b.encode(), algo) == d
    assert oracle.py_lev(a, b) == d


def test_word_boundaries_and_random_agree():
    rng = np.random.default_rng(7)
    al = np.frombuffer(b"ACGT", dtype=np.uint8)
    for m in (1, 31, 32, 33, 63, 64, 65, 127, 128, 129, 1023, 1024, 1025):
        a = al[rng.integers(0, 4, m)].tobytes()
        for b in (a, a[1:], a + b"A", al[rng.integers(0, 4, m)].tobytes(), al[rng.integers(0, 4, max(1, m - 7))].tobytes()):
            d = oracle.nw(a, b, "dp")
            assert oracle.nw(a, b, "myers") == d
            assert oracle.nw(a, b, "edlib_like") == d
    for _ in range(200):
        a = al[rng.integers(0, 4, int(rng.integers(0, 150)))].tobytes()
        b = al[rng.integers(0, 4, int(rng.integers(0, 150)))].tobytes()
        d = oracle.py_lev(a.decode(), b.decode())
        assert oracle.nw(a, b, "dp") == oracle.nw(a, b, "myers") == oracle.nw(a, b, "edlib_like") == d


def test_hw_is_min_over_substrings():
    rng = np.random.default_rng(8)
    al = np.frombuffer(b"ACGT", dtype=np.uint8)
    for _ in range(30):
        q = al[rng.integers(0, 4, int(rng.integers(1, 12)))].tobytes()
        t = al[rng.integers(0, 4, int(rng.integers(12, 30)))].tobytes()
        brute = min(oracle.py_lev(q.decode(), t[i:j].decode()) for i in range(len(t) + 1) for j in range(i, len(t) + 1))
        assert oracle.hw(q, t) == brute


def test_compl_reverse_matches_python_expression():
    s = "AACGTNRYKMSWBDHV-acgtX"
    assert oracle.compl_reverse(s.encode()).decode() == oracle.py_compl_reverse(s)
    assert oracle.py_compl_reverse("ATCGRYKMSW") == "WSKMRYCGAT"[::1]  # involution on the table's support
    assert oracle.py_compl_reverse(oracle.py_compl_reverse(s)) == s


def test_iden_matches_python_round():
    rng = np.random.default_rng(9)
    for L in list(range(1, 400)) + [700, 1000, 1024, 1800, 2999]:
        for d in set([0, 1, L // 5, L // 2, L - 1, L] + rng.integers(0, L + 1, 8).tolist()):
            assert oracle.iden(d, L) == round(1 - d / L, 3), (d, L)


def test_threshold_tables_are_the_float_rule():
    """dpass/drev (product host code) == exhaustive scan of the reference's float expressions."""
    for sg in (50.0, 80.0, 85.0, 93.0, 96.0, 99.9, 100.0):
        cut = sg / 100
        dp, dr = thresholds.tables(cut, 1300)
        for L in list(range(1, 130)) + [699, 700, 1000, 1023, 1024, 1025, 1299]:
            idens = [round(1 - d / L, 3) for d in range(L + 1)]
            passing = [d for d in range(L + 1) if idens[d] >= cut]
            assert dp[L] == max(passing)
            assert dr[L] == min(d for d in range(L + 1) if idens[d] < 0.5)
            assert passing == list(range(len(passing)))  # monotone: a prefix of d values
    dp, dr = thresholds.tables(0.8, 1100)
    assert (dp[700], dr[700], dp[1000], dr[1000], dp[1024], dr[1024]) == (140, 351, 200, 501, 205, 513)


def test_window_geometry_is_the_float_test():
    rng = np.random.default_rng(10)
    L = np.sort(rng.integers(300, 900, 300))
    hi = host.batch_geometry(L)
    for i in range(L.size):
        kept = [j for j in range(i + 1, L.size) if not (int(L[i]) * 1.05 < int(L[j]))]
        assert hi[i] == (kept[-1] if kept else i)
        assert kept == list(range(i + 1, int(hi[i]) + 1))


def test_c_batch_equals_python_restatement():
    """oracle.process_batch (C) == py_process_list (statement-by-statement Python) on a small batch."""
    rng = np.random.default_rng(11)
    reads = util.random_reads(rng, 40, 60, 80, families=3, err=0.08) + util.random_reads(rng, 6, 60, 80, alphabet=b"ACGTN")
    for sg in (80.0, 92.0):
        recs, st = util.oracle_batch(reads, sg)
        buf, offs, order, lens_sorted, *_ = util.batch_inputs(reads, sg)
        text = oracle.format_lines(recs, order, np.arange(len(reads), dtype=np.uint32), offs).decode()
        batch = [[f"r{i}", r.decode(), "u", i] for i, r in enumerate(reads)]
        lines = oracle.py_process_list([batch], sg)
        assert text == "".join(l + "\n" for l in lines)
        assert st["records"] == len(lines) and st["pairs"] > 0
        recs_dp, _ = util.oracle_batch(reads, sg, algo="dp")
        recs_el, _ = util.oracle_batch(reads, sg, algo="edlib_like")
        util.assert_same_records(recs, recs_dp)
        util.assert_same_records(recs, recs_el)
        # product-side formatter (pure host code) produces the same bytes
        assert host.format_records(recs.astype(recs.dtype), np.arange(len(reads))[order], lens_sorted) == text


def test_synth_configs_are_deterministic():
    a, la, _ = synth.make_config(1, scale=0.05)
    b, lb, _ = synth.make_config(1, scale=0.05)
    assert a == b and la == lb and len(a) == 50
    r5, _, args5 = synth.make_config(5, scale=0.004)
    assert len(r5) == 400 and args5[0] == "-a"


def test_kmer_oracle_small_cases():
    # AAAA: one 2-mer AA, canonical min(AA=0, TT=15) = 0; TTTT hits the same canonical k-mer
    assert oracle.kmer_shared(b"AAAA", b"TTTT", 2) == 1
    assert oracle.kmer_shared(b"AAAA", b"CCCC", 2) == 0
    assert oracle.kmer_shared(b"ACGT", b"ACGT", 2) == 2  # AC/GT -> same canonical, CG (palindrome): 2 distinct
    assert oracle.kmer_shared(b"ACNGT", b"ACGT", 2) == 1  # the windows over N are skipped
    s = b"ACGTTGCAAGGCTTAACCGGTT"
    assert oracle.kmer_shared(s, oracle.compl_reverse(s), 5) == oracle.kmer_shared(s, s, 5)  # strand-independent


def test_pyhost_record_walker_matches_the_python_passes():
    """csrc/pyhost.c hands host.process_list the idx keys and the bytes of every SEQ in one C pass; anything that is
    not what read_file builds (amplicon_sorter.py:551-561) makes it step aside (None -> generic Python path)."""
    import ctypes

    from amplicon_sorter_b200 import pyhost

    recs = [["r0", "ACGT", "u", 7], ("r1", "", "u", 9), ["r2", "ACGTNNRY" * 40, "u", 2 ** 40]]
    got = pyhost.collect(recs)
    assert got is not None, "libasb_pyhost.so missing: run python -m amplicon_sorter_b200.build"
    keys, ptrs, lens = got
    assert keys.tolist() == [7, 9, 2 ** 40] and lens.tolist() == [4, 0, 320]
    for (_, s, _, _), p, n in zip(recs, ptrs.tolist(), lens.tolist()):
        assert ctypes.string_at(p, n) == s.encode("ascii")
    assert pyhost.collect([["r0", "ACGÄ", "u", 1]]) is None      # not ASCII: one byte per character does not hold
    assert pyhost.collect([["r0", b"ACGT", "u", 1]]) is None          # not a str
    assert pyhost.collect([["r0", "ACGT", "u", "1"]]) is None         # idx not an int
    assert pyhost.collect([["r0", "ACGT", "u", 2 ** 70]]) is None     # idx beyond int64
    assert pyhost.collect([["r0", "ACGT"]]) is None                   # short record
    assert pyhost.collect(tuple(recs)) is None                        # not a list
    k, p, n = pyhost.collect([])
    assert k.shape == p.shape == n.shape == (0,)
