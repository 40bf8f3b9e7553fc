"""GPU: the tempfile text assembled on the device (csrc/text.cuh) against the host-side formatter and the oracle.

Bar: byte-exact text, and the resident integer lines equal to what parsing that text gives."""
import os
import subprocess
import sys
import types

import numpy as np
import pytest

from amplicon_sorter_b200 import groups, host, synth, thresholds
from oracle import oracle
from tests import util

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class Collect:
    def __init__(self):
        self.parts = []

    def __call__(self, chunk):
        self.parts.append(bytes(chunk.data))
        chunk.release()

    def text(self):
        return b"".join(self.parts)


def oracle_text(reads, idx, sg=80.0):
    buf, offs, order, *_ = util.batch_inputs(reads, sg)
    recs, st = oracle.process_batch(buf, offs, order, sg)
    return oracle.format_lines(recs, order, np.asarray(idx, dtype=np.uint32), offs), recs, st


@pytest.mark.parametrize("cfg,scale,pair_cap", [(5, 0.012, 1 << 26), (5, 0.012, 1 << 16), (2, 0.05, 1 << 15), (1, 0.3, 1 << 26)])
def test_device_text_equals_oracle_text(engine, cfg, scale, pair_cap):
    reads, _, _ = synth.make_config(cfg, scale=scale)
    rng = np.random.default_rng(5)
    idx = rng.permutation(4 * len(reads))[: len(reads)].astype(np.uint32)  # idx values with 1..4 digits, not the read ids
    idx[:3] = [0, 9, 4000000000]                                           # 1-digit and 10-digit extremes
    buf, offs, order, lens_sorted, hi, dpass, drev = util.batch_inputs(reads)
    engine.upload_reads(buf, offs)
    engine.set_param("pair_cap", pair_cap)
    sink = Collect()
    try:
        tot = engine.compare_text(order, hi, dpass, drev, host.text_tables(idx[order], lens_sorted, dpass), sink)
    finally:
        engine.set_param("pair_cap", float(1 << 26))
    want, recs, st = oracle_text(reads, idx)
    assert sink.text() == want
    assert tot["pairs"] == st["pairs"] and tot["n_records"] == len(recs)
    if pair_cap < (1 << 20):
        assert tot["steps"] > 1 and len(sink.parts) > 1
    # the resident line set = the printed lines in integer form
    parsed = groups.Lines.from_text(want.decode())
    a, b, m, r = engine.lines_fetch()
    assert engine.lines_count() == len(parsed)
    assert np.array_equal(a, parsed.a) and np.array_equal(b, parsed.b) and np.array_equal(m, parsed.milli) and np.array_equal(r, parsed.rev)
    # ... and the consumers run on it without an upload
    hist, _ = engine.lines_hist()
    assert np.array_equal(hist, np.bincount(parsed.milli, minlength=1001).astype(np.uint64))


def test_text_step_on_gathered_records_sorts_them(engine):
    """The N > 1 path hands asb_text_step the concatenation of several ranks' lists in device memory."""
    import torch

    reads, _, _ = synth.make_config(5, scale=0.008)
    buf, offs, order, lens_sorted, hi, dpass, drev = util.batch_inputs(reads)
    engine.upload_reads(buf, offs)
    parts = [engine.compare_batch(order, hi, dpass, drev, r, 3)[0] for r in range(3)]
    engine.batch_begin(order, hi, dpass, drev, 0, 1)
    engine.text_begin(*host.text_tables(order, lens_sorted, dpass))
    allr = np.concatenate(parts[::-1])  # rank order reversed: nothing may depend on it
    t = torch.from_numpy(allr.view(np.uint32).reshape(-1, 4).view(np.int32).copy()).cuda()
    got = b""
    for chunk in engine.text_chunks(engine.text_load_tensor(t, sort=True)):
        got += bytes(chunk.data)
        chunk.release()
    want, _, _ = oracle_text(reads, np.arange(len(reads)))
    assert got == want


def test_text_errors_are_reported(engine):
    from amplicon_sorter_b200._ffi import EngineError

    reads, _, _ = synth.make_config(1, scale=0.05)
    buf, offs, order, lens_sorted, hi, dpass, drev = util.batch_inputs(reads)
    engine.upload_reads(buf, offs)
    with pytest.raises(EngineError):  # no batch yet
        engine.text_begin(*host.text_tables(order, lens_sorted, dpass))
    engine.batch_begin(order, hi, dpass, drev)
    short = thresholds.tables(0.99, int(lens_sorted.max()) + 1)[0]  # strings only up to the 0.99 cut-off: records at 0.80 have none
    engine.text_begin(*host.text_tables(order, lens_sorted, short))
    info = engine.batch_step()
    assert info["n_records"] > 0
    engine.text_load(None, info["n_records"])
    with pytest.raises(EngineError):
        engine.text_step(0, info["n_records"])
    with pytest.raises(EngineError):  # a range outside the current record set
        engine.text_step(info["n_records"], 1)


def test_process_list_writes_the_file_and_leaves_resident_lines(engine, tmp_path):
    reads, _, _ = synth.make_config(3, scale=0.04)
    c2 = [[[f"r{i}", s.decode(), "u", i] for i, s in enumerate(reads)]]
    args = types.SimpleNamespace(outputfolder=str(tmp_path), similar_genes=80.0)
    stats = {}
    host.process_list(c2, "f_compare.tmp", args, engine=engine, stats_out=stats)
    path = os.path.join(str(tmp_path), "f_compare.tmp")
    want, recs, st = oracle_text(reads, np.arange(len(reads)))
    assert open(path, "rb").read() == want
    lines = groups.lines_for(path)
    assert isinstance(lines, groups.DeviceLines) and lines._host is None and len(lines) == len(recs)
    # a consumer on the same engine uses the resident set; a later upload first hands the owner its host copy
    tl, a, b, m = groups.best_hits(engine, lines)
    engine.lines_upload(np.zeros(1, np.uint32), np.zeros(1, np.uint32), np.zeros(1, np.uint32))
    assert lines._host is not None and np.array_equal(lines.a, groups.Lines.from_text(want.decode()).a)


@pytest.mark.parametrize("world", [2])
def test_torchrun_sharded_engine_matches_oracle(world, tmp_path):
    """Real NCCL: dist.ShardedEngine on rank 0, worker_loop elsewhere, per-slab device gather, text on rank 0."""
    import torch

    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    port = 29700 + os.getpid() % 200
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
                          "--master-port", str(port), os.path.join(ROOT, "tests", "dist_gpu_check.py")], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert out.stdout.count("identical to the oracle") >= 3, out.stdout


def test_torchrun_engine_failure_aborts_every_rank(tmp_path):
    """Real NCCL: an engine error on one rank stops all ranks at the same step -- no hang, no swallowed exception."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    port = 29750 + os.getpid() % 200
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                          "--master-port", str(port), os.path.join(ROOT, "tests", "dist_gpu_check.py"), "--fail"], capture_output=True, text=True, timeout=300)
    assert out.returncode != 0
    assert out.stdout.count("aborted:") == 2 and "SWALLOWED" not in out.stdout and "returned normally" not in out.stdout, out.stdout + out.stderr[-2000:]
