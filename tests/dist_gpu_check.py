"""Multi-GPU check, run under torchrun on a GPU box (tests/test_gpu_text.py spawns it when >= 2 GPUs are visible):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/dist_gpu_check.py
Every rank decides its shard of the rows on its own GPU; what rank 0 gathers over NCCL must equal the CPU oracle --
as records (compare_batch) and as the text of the tempfile (host.process_list on the ShardedEngine)."""
import os
import sys
import tempfile
import types

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402
import torch  # noqa: E402

from amplicon_sorter_b200 import dist, host, synth  # noqa: E402
from amplicon_sorter_b200.engine import Engine  # noqa: E402
from oracle import oracle  # noqa: E402
from tests import util  # noqa: E402


def failing(r, w, dev, eng):
    """--fail: the engine of the LAST rank fails inside its first batch_step; every rank must leave through
    DistributedAbort (no rank parked in an NCCL collective), and say so."""
    from amplicon_sorter_b200._ffi import EngineError

    if r == w - 1:
        def boom():
            raise EngineError(-1, "injected: CUDA error in batch_step")
        eng.batch_step = boom
    reads, _, _ = synth.make_config(5, scale=0.01)
    try:
        if r != 0:
            dist.worker_loop(eng, dev)
        else:
            sh = dist.ShardedEngine(eng, dev)
            with tempfile.TemporaryDirectory() as tmp:
                args = types.SimpleNamespace(outputfolder=tmp, similar_genes=80.0)
                try:
                    host.process_list([[[f"r{i}", s.decode(), "u", i] for i, s in enumerate(reads)]], "x_compare.tmp", args, engine=sh)
                except Exception:  # the reference's per-file handler must not see the abort
                    print("SWALLOWED", flush=True)
        print(f"rank {r}: returned normally", flush=True)
    except dist.DistributedAbort as exc:
        print(f"rank {r}: aborted: {exc}", flush=True)
        os._exit(3)


def main():
    r, w, dev = dist.init_from_env()
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)
    eng = Engine(dev.index, stream=stream.cuda_stream)
    if "--fail" in sys.argv:
        return failing(r, w, dev, eng)
    if r != 0:
        dist.worker_loop(eng, dev)
        return
    sh = dist.ShardedEngine(eng, dev)
    for cfg, scale in ((5, 0.03), (2, 0.05)):
        reads, _, _ = synth.make_config(cfg, scale=scale)
        buf, offs, order, lens_sorted, hi, dpass, drev = util.batch_inputs(reads)
        want, st = util.oracle_batch(reads)
        sh.upload_reads(buf, offs)
        sh.set_param("pair_cap", 1 << 20)
        got, tot = sh.compare_batch(order, hi, dpass, drev)
        util.assert_same_records(got, want)
        assert tot["pairs"] == st["pairs"], (tot["pairs"], st["pairs"])
        print(f"world={w} cfg{cfg}: {st['pairs']} pairs, {len(want)} records identical to the oracle", flush=True)
        if cfg == 5:
            with tempfile.TemporaryDirectory() as tmp:
                args = types.SimpleNamespace(outputfolder=tmp, similar_genes=80.0)
                stats = {}
                host.process_list([[[f"r{i}", s.decode(), "u", i] for i, s in enumerate(reads)]], "x_compare.tmp", args, engine=sh, stats_out=stats)
                text = open(os.path.join(tmp, "x_compare.tmp"), "rb").read()
            assert text == oracle.format_lines(want, order, np.arange(len(reads), dtype=np.uint32), offs)
            assert stats["pairs"] == st["pairs"] and sh.lines_count() == len(want)
            print(f"world={w} cfg{cfg}: tempfile of host.process_list ({len(text)} bytes) identical to the oracle", flush=True)
    sh.close()


if __name__ == "__main__":
    main()
