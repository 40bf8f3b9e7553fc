"""Consumers of <stem>_compare.tmp: SSG, the best-hit filter, grouping.

Golden vectors tests/golden/h_*.json.gz hold what the UNMODIFIED reference functions (SSG, update_list,
read_indexes, merge_groups) computed from a given file (tests/golden/make_golden_groups.py).  They must be reproduced

  * by the CPU oracle (oracle.py restatements; pins the oracle to the reference)                      -- CPU
  * by the product's host code (amplicon_sorter_b200/groups.py) driven by the oracle-backed engine    -- CPU
  * by the product (host code + CUDA kernels of lines.cuh through the C ABI)                          -- GPU
and on large seeded inputs the CUDA kernels must equal the C oracle exactly.                          -- GPU
"""
import glob
import gzip
import json
import os
import random

import numpy as np
import pytest

from amplicon_sorter_b200 import groups
from oracle import oracle
from tests.fake_engine import OracleEngine

HERE = os.path.dirname(os.path.abspath(__file__))
FIXTURES = sorted(glob.glob(os.path.join(HERE, "golden", "h_*.json.gz")))
IDS = [os.path.basename(p)[:-8] for p in FIXTURES]


def load(path):
    with gzip.open(path, "rt") as f:
        return json.load(f)


def as_sorted(groups_):
    return [sorted(g, key=int) for g in groups_]


def test_fixtures_present():
    assert len(FIXTURES) >= 7


def test_milli_is_an_exact_stand_in_for_iden():
    """float(text) == milli/1000 and str(milli/1000) == text for everything round(x, 3) can print, and string
    order == numeric order (the reference compares the TEXT at amplicon_sorter.py:998)."""
    seen = set()
    for L in list(range(300, 1300, 7)) + [1000, 1024, 2500]:
        for d in range(0, L // 2 + 2, 3):
            x = round(1 - d / L, 3)
            m = int(round(x * 1000))
            assert m / 1000 == x and groups.IDEN_STR[m] == str(x)
            seen.add(m)
    assert len(seen) > 500
    vals = [m for m in range(450, 1001)]
    assert sorted(vals, key=lambda m: groups.IDEN_STR[m]) == vals


@pytest.mark.parametrize("path", FIXTURES, ids=IDS)
def test_oracle_reproduces_reference(path):
    fx = load(path)
    text = fx["compare_tmp"]
    assert oracle.py_ssg(text) == fx["ssg"]
    templist = oracle.py_besthit_templist(text)
    assert templist == fx["update_list"]["templist"]
    n_greedy, merged = oracle.py_groups(templist)
    assert n_greedy == fx["update_list"]["n_greedy"]
    assert as_sorted(merged) == fx["update_list"]["groups"]
    for case in fx["read_indexes"]:
        tl = oracle.py_besthit_templist(text, case["ssg"] / 100, set(case["members"]))
        assert tl == case["templist"]
        n_greedy, merged = oracle.py_groups(tl)
        assert n_greedy == case["n_greedy"] and as_sorted(merged) == case["groups"]


def check_product(fx, engine):
    lines = groups.Lines.from_text(fx["compare_tmp"])
    assert len(lines) == fx["compare_tmp"].count("\n")
    assert groups.ssg_estimate(engine, lines) == fx["ssg"]
    templist, a, b, m = groups.best_hits(engine, lines)
    assert templist == fx["update_list"]["templist"]
    n_greedy, merged = groups.make_groups(engine, templist)
    assert n_greedy == fx["update_list"]["n_greedy"]
    assert as_sorted(merged) == fx["update_list"]["groups"]  # same partition, same numbering
    # ... and the same set internals as the reference's own sequence of set operations (same process = same hash seed)
    assert [list(g) for g in merged] == [list(g) for g in oracle.py_groups(templist)[1]]
    for case in fx["read_indexes"]:
        tl, a, b, m = groups.best_hits(engine, lines, case["ssg"] / 100, set(case["members"]))
        assert tl == case["templist"]
        n_greedy, merged = groups.make_groups(engine, tl, update_with_list=True)
        assert n_greedy == case["n_greedy"] and as_sorted(merged) == case["groups"]
        assert [list(g) for g in merged] == [list(g) for g in oracle.py_groups(tl, update_with_list=True)[1]]


@pytest.mark.parametrize("path", FIXTURES, ids=IDS)
def test_host_code_with_oracle_engine_reproduces_reference(path):
    check_product(load(path), OracleEngine())


@pytest.mark.gpu
@pytest.mark.parametrize("path", FIXTURES, ids=IDS)
def test_product_reproduces_reference(path, engine):
    check_product(load(path), engine)


def random_lines(seed, n_idx, n, lo, hi, step):
    rng = np.random.default_rng(seed)
    a = rng.integers(0, n_idx, n).astype(np.uint32)
    b = ((a + 1 + rng.integers(0, n_idx - 1, n)) % n_idx).astype(np.uint32)
    milli = (lo + step * rng.integers(0, (hi - lo) // step + 1, n)).astype(np.uint32)
    return a, b, milli


def test_c_oracle_equals_python_restatement():
    """Pins oracle/asref.c::asref_besthit (used at sizes Python cannot reach) to the statement-by-statement one."""
    for seed, (n_idx, n, lo, hi, step) in enumerate([(40, 4000, 900, 1000, 20), (500, 20000, 600, 1000, 1), (7, 300, 990, 1000, 5)]):
        a, b, milli = random_lines(seed, n_idx, n, lo, hi, step)
        text = "".join(f"{x}:{y}:{groups.IDEN_STR[z]}\n" for x, y, z in zip(a.tolist(), b.tolist(), milli.tolist()))
        rnd = random.Random(seed)
        members = set(str(v) for v in rnd.sample(range(n_idx), max(1, n_idx // 3)))
        for ssg, idx in ((None, None), (0.93, members)):
            want = oracle.py_besthit_templist(text, ssg, idx)
            lines = groups.Lines(a, b, milli)
            got, *_ = groups.best_hits(OracleEngine(), lines, ssg, idx)
            assert got == want


@pytest.mark.gpu
@pytest.mark.parametrize("n_idx,n,lo,hi,step", [(1000, 200000, 800, 1000, 1), (50, 100000, 900, 1000, 10), (100000, 3000000, 800, 1000, 1),
                                                (3, 50000, 1000, 1000, 1), (20000, 1000000, 990, 1000, 2)])
def test_kernels_equal_c_oracle_on_large_inputs(engine, n_idx, n, lo, hi, step):
    a, b, milli = random_lines(n_idx + n, n_idx, n, lo, hi, step)
    engine.lines_upload(a, b, milli)
    hist, _ = engine.lines_hist()
    assert np.array_equal(hist, np.bincount(milli, minlength=1001).astype(np.uint64))
    rng = np.random.default_rng(5)
    member = groups.member_bitmap(set(str(v) for v in rng.integers(0, n_idx, max(1, n_idx // 4)).tolist()), n_idx)
    for min_milli, mb in ((0, None), (930, member), (1001, None)):
        line, first, _ = engine.lines_besthit(min_milli, mb)
        wl, wf = oracle.besthit(a, b, milli, min_milli, mb)
        assert np.array_equal(line, wl) and np.array_equal(first, wf)
        if min_milli == 0:
            label, _ = engine.components(a[line], b[line], n_idx)
            assert np.array_equal(label, oracle.components(a[line], b[line], n_idx))
    # components on the raw (dense) edge list too: long chains and big stars
    label, _ = engine.components(a, b, n_idx)
    assert np.array_equal(label, oracle.components(a, b, n_idx))


@pytest.mark.gpu
def test_degenerate_line_sets(engine):
    z = np.zeros(0, dtype=np.uint32)
    engine.lines_upload(z, z, z)
    assert engine.lines_hist()[0].sum() == 0
    line, first, _ = engine.lines_besthit()
    assert line.shape[0] == 0
    label, _ = engine.components(z, z, 5)
    assert label.tolist() == [0, 1, 2, 3, 4]
    with pytest.raises(Exception):
        engine.lines_upload(np.array([1], np.uint32), np.array([2], np.uint32), np.array([1001], np.uint32))


def test_process_list_leaves_the_lines_of_the_file_it_wrote(tmp_path):
    """host.process_list hands the integer lines to the consumers (groups.CACHE): they must be the file's lines."""
    import types

    from amplicon_sorter_b200 import host
    from tests.test_golden import load as load_g, rebuild_comparelist2

    fx = load_g(os.path.join(HERE, "golden", "g2_random.json.gz"))  # -ra mode: several overlapping batches, one engine batch
    args = types.SimpleNamespace(outputfolder=str(tmp_path), similar_genes=fx["similar_genes"], nprocesses=1)
    open(os.path.join(str(tmp_path), "results.txt"), "w").close()
    path = os.path.join(str(tmp_path), "x_compare.tmp")
    host.process_list(rebuild_comparelist2(fx), path, args, engine=OracleEngine())
    cached = groups.lines_for(path)
    parsed = groups.Lines.from_text(open(path).read())
    assert open(path).read() == fx["compare_tmp"] and len(cached) == len(parsed) > 0
    for f in ("a", "b", "milli", "rev"):
        assert np.array_equal(getattr(cached, f), getattr(parsed, f)), f
    # a file that changed on disk is re-parsed, a missing one raises like the reference's open()
    with open(path, "a") as f:
        f.write("1:2:0.9\n")
    assert len(groups.lines_for(path)) == len(parsed) + 1
    os.remove(path)
    with pytest.raises(FileNotFoundError):
        groups.lines_for(path)


@pytest.mark.parametrize("seed", range(6))
def test_randomised_host_logic_against_the_statement_by_statement_restatement(seed):
    """Small random files with heavy ties, repeated pairs and both admission modes."""
    rnd = random.Random(100 + seed)
    n_idx = rnd.choice([5, 12, 40])
    n = rnd.choice([1, 30, 400, 2000])
    rows = []
    for _ in range(n):
        a, b = rnd.sample(range(n_idx), 2)
        rows.append(f"{a}:{b}:{groups.IDEN_STR[rnd.choice([1000, 990, 985, 950, 900, rnd.randrange(500, 1001)])]}" + (":reverse" if rnd.random() < 0.3 else ""))
    text = "\n".join(rows) + "\n"
    lines = groups.Lines.from_text(text)
    eng = OracleEngine()
    assert groups.ssg_estimate(eng, lines) == oracle.py_ssg(text)
    members = set(str(v) for v in rnd.sample(range(n_idx), max(1, n_idx // 2)))
    for ssg, idx, as_list in ((None, None, False), (0.95, members, True), (0.5, members, True), (1.01, members, True)):
        want = oracle.py_besthit_templist(text, ssg, idx)
        got, *_ = groups.best_hits(eng, lines, ssg, idx)
        assert got == want
        n_greedy, merged = groups.make_groups(eng, got, update_with_list=as_list)
        w_greedy, w_merged = oracle.py_groups(want, update_with_list=as_list)
        assert n_greedy == w_greedy and [list(g) for g in merged] == [list(g) for g in w_merged]
