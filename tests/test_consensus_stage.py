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
