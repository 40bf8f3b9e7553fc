"""Multi-GPU check (run under torchrun on a GPU box; not collected by pytest):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/dist_gpu_check.py
Every rank decides its cyclic shard on its own GPU; the NCCL gather on rank 0 must equal the CPU oracle."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402

from amplicon_sorter_b200 import dist, synth  # noqa: E402
from amplicon_sorter_b200.engine import Engine  # noqa: E402
from tests import util  # noqa: E402


def main():
    r, w, dev = dist.init_from_env()
    eng = Engine(dev.index)
    if r != 0:
        dist.worker_loop(eng, dev)
        return
    sh = dist.ShardedEngine(eng, dev)
    for cfg, scale in ((5, 0.03), (2, 0.05)):
        reads, _, _ = synth.make_config(cfg, scale=scale)
        buf, offs, order, lens_sorted, hi, dpass, drev = util.batch_inputs(reads)
        sh.upload_reads(buf, offs)
        sh.set_param("pair_cap", 1 << 20)
        got, tot = sh.compare_batch(order, hi, dpass, drev)
        want, st = util.oracle_batch(reads)
        util.assert_same_records(got, want)
        assert tot["pairs"] == st["pairs"], (tot["pairs"], st["pairs"])
        print(f"world={w} cfg{cfg}: {st['pairs']} pairs, {len(want)} records identical to the oracle", flush=True)
    sh.close()


if __name__ == "__main__":
    main()
