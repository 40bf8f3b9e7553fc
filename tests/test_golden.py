"""Golden vectors: <stem>_compare.tmp written by the UNMODIFIED reference script (run in the build
container by tests/golden/make_golden.py) must be reproduced byte for byte

  * by the CPU oracle (pins the oracle's control flow to the reference)         -- CPU
  * by the product's host code driven by the oracle behind the engine interface  -- CPU
  * by the product (host code + CUDA engine through the C ABI)                   -- GPU
"""
import glob
import gzip
import json
import os
import types

import numpy as np
import pytest

from amplicon_sorter_b200 import host, synth
from oracle import oracle
from tests.fake_engine import OracleEngine

HERE = os.path.dirname(os.path.abspath(__file__))
FIXTURES = sorted(glob.glob(os.path.join(HERE, "golden", "g*.json.gz")))


def load(path):
    with gzip.open(path, "rt") as f:
        return json.load(f)


def rebuild_comparelist2(fx):
    """[[id, SEQ, 'u', idx], ...] batches exactly as read_file built them (amplicon_sorter.py:551-622).
    Overlapping -ra batches share record objects, as in the reference."""
    records = {}
    batches = []
    for b in fx["batches_before"]:
        batch = []
        for idx in b:
            if idx not in records:
                records[idx] = [f"r{idx}", fx["records"][str(idx)], "u", idx]
            batch.append(records[idx])
        batches.append(batch)
    return batches


def run_host(fx, tmp_path, engine):
    args = types.SimpleNamespace(outputfolder=str(tmp_path), similar_genes=fx["similar_genes"], nprocesses=1)
    open(os.path.join(str(tmp_path), "results.txt"), "w").close()
    batches = rebuild_comparelist2(fx)
    tempfile = os.path.join(str(tmp_path), fx["name"] + "_compare.tmp")
    stats = {}
    host.process_list(batches, tempfile, args, engine=engine, stats_out=stats)
    with open(tempfile) as f:
        text = f.read()
    return text, batches, stats


def test_fixtures_present():
    assert len(FIXTURES) >= 3


@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p) for p in FIXTURES])
def test_oracle_reproduces_reference_file(path):
    fx = load(path)
    idxs = sorted(int(k) for k in fx["records"])
    rid = {idx: t for t, idx in enumerate(idxs)}
    reads = [fx["records"][str(idx)].encode() for idx in idxs]
    buf, offs = synth.pack_reads(reads)
    lens = (offs[1:] - offs[:-1]).astype(np.int64)
    out = []
    for b in fx["batches_before"]:
        ids = np.asarray([rid[x] for x in b], dtype=np.int64)
        order = ids[np.argsort(lens[ids], kind="stable")].astype(np.uint32)
        recs, _ = oracle.process_batch(buf, offs, order, fx["similar_genes"])
        out.append(oracle.format_lines(recs, order, np.asarray(idxs, dtype=np.uint32), offs).decode())
    assert "".join(out) == fx["compare_tmp"]


@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p) for p in FIXTURES])
def test_host_stage_with_oracle_engine_reproduces_reference_file(path, tmp_path):
    fx = load(path)
    eng = OracleEngine()
    text, batches, stats = run_host(fx, tmp_path, eng)
    assert text == fx["compare_tmp"]
    # side effect (1): batches left length-sorted in place, in the reference's stable order
    assert [[rec[3] for rec in b] for b in batches] == fx["batches_after"]
    # the reads went up through the record walker (csrc/pyhost.c): pointers into the records' own str objects, merged
    # by idx for the overlapping -ra batches
    assert getattr(eng, "scattered_uploads", 0) == 1


@pytest.mark.parametrize("path", FIXTURES[:1], ids=[os.path.basename(p) for p in FIXTURES[:1]])
def test_host_stage_without_the_record_walker_is_the_same(path, tmp_path, monkeypatch):
    """Records the walker does not handle (or a missing helper library) take the generic Python path: same file."""
    from amplicon_sorter_b200 import pyhost

    monkeypatch.setattr(pyhost, "collect", lambda records: None)
    fx = load(path)
    eng = OracleEngine()
    text, batches, stats = run_host(fx, tmp_path, eng)
    assert text == fx["compare_tmp"]
    assert [[rec[3] for rec in b] for b in batches] == fx["batches_after"]
    assert getattr(eng, "scattered_uploads", 0) == 0


@pytest.mark.gpu
@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p) for p in FIXTURES])
def test_product_reproduces_reference_file(path, tmp_path, engine):
    fx = load(path)
    text, batches, stats = run_host(fx, tmp_path, engine)
    assert text == fx["compare_tmp"]
    assert [[rec[3] for rec in b] for b in batches] == fx["batches_after"]
    assert stats["pairs"] == stats["tl"] > 0


def test_no_comparable_pairs_is_the_reference_error(tmp_path):
    """amplicon_sorter.py:702-706/:768-772: nothing to compare -> note in results.txt + Exception."""
    args = types.SimpleNamespace(outputfolder=str(tmp_path), similar_genes=80.0, nprocesses=1)
    open(os.path.join(str(tmp_path), "results.txt"), "w").close()
    batches = [[["a", "A" * 300, "u", 0], ["b", "C" * 400, "u", 1]], []]
    with pytest.raises(Exception):
        host.process_list(batches, os.path.join(str(tmp_path), "x_compare.tmp"), args, engine=OracleEngine())
    assert "No reads to compare" in open(os.path.join(str(tmp_path), "results.txt")).read()
    assert not os.path.exists(os.path.join(str(tmp_path), "x_compare.tmp"))
