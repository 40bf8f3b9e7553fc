"""CPU, world_size 2, gloo: the multi-rank host path (broadcast of the job, cyclic sharding, gather of
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
        json.dump({k: (int(v) if isinstance(v, (int, np.integer)) else float(v)) for k, v in stats.items()}, f)


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
