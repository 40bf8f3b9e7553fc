"""CPU, world_size 2, gloo: the multi-rank host path (broadcast of the job, row sharding, gather of
the per-rank record lists, merge on rank 0) reproduces the single-rank result and the golden file."""
import gzip
import json
import os
import types

import numpy as np
import pytest
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


def _worker(rank, world, port, fixture, outdir):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from amplicon_sorter_b200 import dist, host
    from tests.fake_engine import OracleEngine
    from tests.test_golden import load, rebuild_comparelist2

    r, w, dev = dist.init_from_env("gloo")
    eng = OracleEngine()
    if r != 0:
        dist.worker_loop(eng, dev)
        return
    sharded = dist.ShardedEngine(eng, dev)
    fx = load(fixture)
    args = types.SimpleNamespace(outputfolder=outdir, similar_genes=fx["similar_genes"], nprocesses=1)
    open(os.path.join(outdir, "results.txt"), "w").close()
    batches = rebuild_comparelist2(fx)
    stats = {}
    host.process_list(batches, os.path.join(outdir, "x_compare.tmp"), args, engine=sharded, stats_out=stats)
    sharded.close()
    with open(os.path.join(outdir, "stats.json"), "w") as f:
        json.dump({k: (int(v) if isinstance(v, (int, np.integer)) else float(v)) for k, v in stats.items() if not isinstance(v, dict)}, f)


@pytest.mark.parametrize("name", ["g1_default", "g3_all_mixed"])
def test_two_ranks_reproduce_the_reference_file(name, tmp_path):
    fixture = os.path.join(HERE, "golden", name + ".json.gz")
    port = 29600 + (os.getpid() % 300)
    mp.spawn(_worker, args=(2, port, fixture, str(tmp_path)), nprocs=2, join=True)
    with gzip.open(fixture, "rt") as f:
        fx = json.load(f)
    assert open(os.path.join(str(tmp_path), "x_compare.tmp")).read() == fx["compare_tmp"]
    stats = json.load(open(os.path.join(str(tmp_path), "stats.json")))
    assert stats["pairs"] == stats["tl"] > 0  # the two shards partition the pair set


def _failing_worker(rank, world, port, fixture, outdir, where):
    """One rank's engine fails inside a sharded call; every rank must leave through DistributedAbort (a SystemExit),
    nobody may stay parked in a collective (the torchrun path had no failure protocol)."""
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from amplicon_sorter_b200 import dist, host
    from amplicon_sorter_b200._ffi import EngineError
    from tests.fake_engine import OracleEngine
    from tests.test_golden import load, rebuild_comparelist2

    class Broken(OracleEngine):
        def upload_reads(self, buf, offs):
            if where == "upload":
                raise EngineError(-3, "injected: out of memory in upload_reads")
            return super().upload_reads(buf, offs)

        def batch_step(self):
            if where == "step":
                raise EngineError(-1, "injected: CUDA error in batch_step")
            return super().batch_step()

    bad_rank = 1 if where != "rank0-step" else 0
    if where == "rank0-step":
        where = "step"
    r, w, dev = dist.init_from_env("gloo")
    eng = Broken() if r == bad_rank else OracleEngine()
    outcome = "returned"
    try:
        if r != 0:
            dist.worker_loop(eng, dev)
        else:
            sharded = dist.ShardedEngine(eng, dev)
            fx = load(fixture)
            args = types.SimpleNamespace(outputfolder=outdir, similar_genes=fx["similar_genes"], nprocesses=1)
            open(os.path.join(outdir, "results.txt"), "w").close()
            try:
                try:
                    host.process_list(rebuild_comparelist2(fx), os.path.join(outdir, "x_compare.tmp"), args, engine=sharded)
                except Exception:  # the reference's per-file handler (amplicon_sorter.py:2184) -- must NOT see the abort
                    outcome = "swallowed"
            finally:
                sharded.close()  # skips the 'stop' broadcast after an abort
    except dist.DistributedAbort:
        outcome = "abort"
    with open(os.path.join(outdir, f"outcome_{rank}.txt"), "w") as f:
        f.write(outcome)


@pytest.mark.timeout(120)
@pytest.mark.parametrize("where", ["upload", "step", "rank0-step"])
def test_engine_failure_on_one_rank_aborts_all_ranks_without_hanging(where, tmp_path):
    fixture = os.path.join(HERE, "golden", "g1_default.json.gz")
    port = 29900 + (os.getpid() % 90) + {"upload": 0, "step": 1, "rank0-step": 2}[where] * 100
    mp.spawn(_failing_worker, args=(2, port, fixture, str(tmp_path), where), nprocs=2, join=True)
    assert [open(os.path.join(str(tmp_path), f"outcome_{r}.txt")).read() for r in range(2)] == ["abort", "abort"]
