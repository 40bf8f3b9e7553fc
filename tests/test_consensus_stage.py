"""The "next" row: reads x group consensuses (process_consensuslist AS:1627-1690 + similarity_species
AS:1692-1715).  CPU: host code on the oracle-backed engine == statement-by-statement restatement.
GPU: the same through the CUDA engine; plus asb_threeway_pairs against the oracle on raw pair lists."""
import os
import types

import numpy as np
import pytest

from amplicon_sorter_b200 import host, synth, thresholds
from oracle import oracle
from tests import util
from tests.fake_engine import OracleEngine

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def make_case(seed, n_reads=60, n_groups=4, L=120):
    rng = np.random.default_rng(seed)
    T = [rng.integers(0, 4, int(L * f), dtype=np.uint8) for f in (1.0, 1.0, 1.03, 1.2)[:n_groups]]
    comparelist2 = []
    for i in range(n_reads):
        t = T[i % n_groups]
        r = synth.mutate(rng, t, sub=0.02, ins=0.01, dele=0.01)
        if rng.random() < 0.5:
            r = synth.revcomp_codes(r)
        comparelist2.append([f"r{i}", ACGT[r].tobytes().decode(), "u", i])
    # groups: a few member indexes (strings) followed by the consensus (last item), AS:1653
    grouplist = []
    for g in range(n_groups):
        members = [str(i) for i in range(g, 12, n_groups)]
        cons = ACGT[T[g]].tobytes().decode()
        if g == 1:
            cons = cons[:40] + "R" + cons[41:80] + "N" + cons[81:]  # consensus with ambiguity codes
        grouplist.append(members + [cons])
    indexes = {str(i) for i in range(n_reads)}
    return indexes, grouplist, comparelist2


def run_host(engine, case, similar, tmp_path):
    indexes, grouplist, comparelist2 = case
    args = types.SimpleNamespace(outputfolder=str(tmp_path), nprocesses=1)
    host.process_consensuslist(indexes, grouplist, "x_0.group", args=args, comparelist2=comparelist2, similar=similar, engine=engine)
    p = os.path.join(str(tmp_path), "x_0.tmp")
    return open(p).read().splitlines() if os.path.exists(p) else None


@pytest.mark.parametrize("similar", [0.95, 0.94, 0.9, 0.85])
def test_host_consensus_stage_matches_restatement(similar, tmp_path):
    case = make_case(int(similar * 100))
    want = oracle.py_process_consensuslist(*case, similar)
    got = run_host(OracleEngine(), case, similar, tmp_path)
    assert got == want and len(want) > 10


@pytest.mark.gpu
@pytest.mark.parametrize("similar", [0.95, 0.94, 0.88])
def test_product_consensus_stage_matches_restatement(engine, similar, tmp_path):
    case = make_case(7 + int(similar * 100), n_reads=90)
    want = oracle.py_process_consensuslist(*case, similar)
    got = run_host(engine, case, similar, tmp_path)
    assert got == want and len(want) > 10


@pytest.mark.gpu
def test_threeway_pairs_any_length_order(engine):
    rng = np.random.default_rng(77)
    reads = util.random_reads(rng, 150, 280, 340, families=5, err=0.06) + util.random_reads(rng, 30, 280, 340)
    buf, offs = synth.pack_reads(reads)
    engine.upload_reads(buf, offs)
    n = len(reads)
    q = rng.integers(0, n, 4000).astype(np.uint32)
    t = rng.integers(0, n, 4000).astype(np.uint32)
    key = q.astype(np.uint64) << np.uint64(32) | t
    _, first = np.unique(key, return_index=True)
    q, t = q[np.sort(first)], t[np.sort(first)]  # distinct pairs
    for sg in (80.0, 93.0):
        dpass, drev = thresholds.tables(sg / 100, 400)
        got, info = engine.threeway_pairs(q, t, dpass, drev)
        fe = OracleEngine()
        fe.upload_reads(buf, offs)
        want, _ = fe.threeway_pairs(q, t, dpass, drev)
        util.assert_same_records(got, want)
        assert info["pairs"] == q.shape[0] and len(want) > 10


@pytest.mark.gpu
def test_threeway_pairs_queries_of_very_different_lengths(engine):
    """Adjacent rows of an explicit pair list need not be alike: a list warp may hold the pairs of a 90-base and of a
    650-base query at the same time (two query slots), targets longer and shorter than either."""
    rng = np.random.default_rng(78)
    reads = []
    for lo, hi in ((80, 100), (300, 330), (620, 680)):
        reads += util.random_reads(rng, 60, lo, hi, families=3, err=0.06)
    perm = rng.permutation(len(reads))
    reads = [reads[i] for i in perm]  # read ids (= rows) in random length order
    buf, offs = synth.pack_reads(reads)
    engine.upload_reads(buf, offs)
    n = len(reads)
    q = np.repeat(np.arange(n, dtype=np.uint32), 12)
    t = rng.integers(0, n, q.shape[0]).astype(np.uint32)
    key = q.astype(np.uint64) << np.uint64(32) | t
    _, first = np.unique(key, return_index=True)
    q, t = q[np.sort(first)], t[np.sort(first)]
    dpass, drev = thresholds.tables(0.80, 800)
    fe = OracleEngine()
    fe.upload_reads(buf, offs)
    want, _ = fe.threeway_pairs(q, t, dpass, drev)
    try:
        for two in (1, 0):
            engine.set_param("two_rows", two)
            got, info = engine.threeway_pairs(q, t, dpass, drev)
            util.assert_same_records(got, want)
    finally:
        engine.set_param("two_rows", 1)
    assert len(want) > 50 and (want["reverse"] == 1).any()


def make_todolist(seed, n=40, L=150):
    """[A1, A2, y, z] entries as comp_consensus_groups / compare_consensus spool them (AS:1275-1299)."""
    rng = np.random.default_rng(seed)
    base = [rng.integers(0, 4, int(L * f), dtype=np.uint8) for f in (1.0, 1.1, 0.8, 1.0, 1.25)]
    cons = []
    for i in range(n):
        r = synth.mutate(rng, base[i % len(base)], sub=0.03, ins=0.01, dele=0.01)
        if i % 3 == 0:
            r = r[10:-5]  # nested consensus: HW must find it inside the longer one
        if rng.random() < 0.5:
            r = synth.revcomp_codes(r)
        cons.append(ACGT[r].tobytes().decode())
    cons[3] = cons[3][:20] + "N" + cons[3][21:]
    return [[cons[y], cons[z], y, z] for y in range(n) for z in range(y + 1, n) if (y + z) % 3 != 1]


def run_iden(engine, todolist, tmp_path):
    import pickle
    with open(os.path.join(str(tmp_path), "file_0.todo"), "wb") as f:
        pickle.dump(todolist, f)
    host.iden_consensus_files(str(tmp_path), "consensus.tmp", "...comparing consensuses ", engine=engine)
    assert not os.path.exists(os.path.join(str(tmp_path), "file_0.todo"))
    return open(os.path.join(str(tmp_path), "consensus.tmp")).read().splitlines()


def test_hw_oracle_definition():
    assert oracle.py_hw("ACGT", "TTACGTTT") == 0 and oracle.hw(b"ACGT", b"TTACGTTT") == 0
    assert oracle.py_hw("ACGT", "TTAGGTTT") == 1 and oracle.hw(b"ACGT", b"TTAGGTTT") == 1
    rng = np.random.default_rng(5)
    for _ in range(40):
        q = ACGT[rng.integers(0, 4, int(rng.integers(1, 40)))].tobytes()
        t = ACGT[rng.integers(0, 4, int(rng.integers(40, 90)))].tobytes()
        assert oracle.hw(q, t) == oracle.py_hw(q.decode(), t.decode())


def test_host_iden_consensus_matches_restatement(tmp_path):
    todo = make_todolist(1, n=16, L=60)
    assert run_iden(OracleEngine(), todo, tmp_path) == oracle.py_iden_consensus(todo)


@pytest.mark.gpu
def test_product_iden_consensus_matches_restatement(engine, tmp_path):
    todo = make_todolist(2, n=30, L=150)
    want = oracle.py_iden_consensus(todo)
    assert run_iden(engine, todo, tmp_path) == want and len(want) > 20


@pytest.mark.gpu
def test_hw_distance_pairs(engine):
    rng = np.random.default_rng(9)
    reads = util.random_reads(rng, 40, 30, 700, families=3, err=0.05) + [b"ACGT", b"TTACGTTT", b"A" * 33, b"C"]
    reads += [reads[0][40:300], reads[1][5:-7]]
    buf, offs = synth.pack_reads(reads)
    engine.upload_reads(buf, offs)
    n = len(reads)
    a = rng.integers(0, n, 500).astype(np.uint32)
    b = rng.integers(0, n, 500).astype(np.uint32)
    strand = rng.integers(0, 2, 500).astype(np.uint8)
    got = engine.distance_pairs(a, b, strand, mode="HW")
    fe = OracleEngine()
    fe.upload_reads(buf, offs)
    want = fe.distance_pairs(a, b, strand, mode="HW")
    assert np.array_equal(got, want)


# ---- pinned to the reference: fixtures written by the UNMODIFIED process_consensuslist / similarity_species / do_parallel /
# ---- iden_consensus (tests/golden/make_golden_stages.py, run where /root/reference exists)
import glob  # noqa: E402
import gzip  # noqa: E402
import json  # noqa: E402

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
F1 = sorted(glob.glob(os.path.join(GOLD, "s_consensuslist_*.json.gz")))
F2 = sorted(glob.glob(os.path.join(GOLD, "s_iden_consensus_*.json.gz")))


def _load(path):
    with gzip.open(path, "rt") as f:
        return json.load(f)


def _f1_case(fx):
    return set(fx["indexes"]), [list(g) for g in fx["grouplist"]], [list(r) for r in fx["comparelist2"]]


def test_stage_fixtures_present():
    assert len(F1) == 3 and len(F2) == 2
    assert {_load(p)["similar"] for p in F1} == {0.95, 0.94, 0.88}
    for p in F1 + F2:
        assert "unmodified" in _load(p)["reference"]


@pytest.mark.parametrize("path", F1, ids=[os.path.basename(p) for p in F1])
def test_reference_group_tmp_oracle_and_host(path, tmp_path):
    """CPU: the restatement (pins the oracle) and the product's host code on the oracle-backed engine reproduce the
    <group>.tmp the unmodified reference wrote."""
    fx = _load(path)
    want = fx["group_tmp"].splitlines()
    assert len(want) > 50 and any(len(line.split(":")) == 3 for line in want)
    assert oracle.py_process_consensuslist(*_f1_case(fx), fx["similar"], lev=oracle.c_lev) == want
    assert run_host(OracleEngine(), _f1_case(fx), fx["similar"], tmp_path) == want


@pytest.mark.gpu
@pytest.mark.parametrize("path", F1, ids=[os.path.basename(p) for p in F1])
def test_reference_group_tmp_product(engine, path, tmp_path):
    fx = _load(path)
    assert run_host(engine, _f1_case(fx), fx["similar"], tmp_path) == fx["group_tmp"].splitlines()


@pytest.mark.parametrize("path", F2, ids=[os.path.basename(p) for p in F2])
def test_reference_consensus_tmp_oracle_and_host(path, tmp_path):
    fx = _load(path)
    want = fx["consensus_tmp"].splitlines()
    todo = [list(e) for e in fx["todolist"]]
    assert 20 < len(want) < len(todo)  # some pairs fall below 0.60 (:1151)
    assert oracle.py_iden_consensus(todo, hw_fn=oracle.c_hw) == want
    assert run_iden(OracleEngine(), todo, tmp_path) == want


@pytest.mark.gpu
@pytest.mark.parametrize("path", F2, ids=[os.path.basename(p) for p in F2])
def test_reference_consensus_tmp_product(engine, path, tmp_path):
    fx = _load(path)
    assert run_iden(engine, [list(e) for e in fx["todolist"]], tmp_path) == fx["consensus_tmp"].splitlines()
