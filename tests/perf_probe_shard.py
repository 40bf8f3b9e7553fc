"""One rank's share of the config-5 job on ONE GPU (rank r of `world`), for several slab sizes: what a GPU of an
N-GPU run spends in its kernels, without the box.  python tests/perf_probe_shard.py [world] [rank]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import bench  # noqa: E402
from amplicon_sorter_b200 import host  # noqa: E402
from amplicon_sorter_b200.engine import Engine  # noqa: E402

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
rank = int(sys.argv[2]) if len(sys.argv) > 2 else 0
w = bench.make_workload(5, 1.0)
with Engine(0) as eng:
    eng.upload_reads(w["buf"], w["offs"])
    for slab in (1 << 30, 1 << 31, 1 << 32):
        eng.set_param("slab_pairs", float(slab))
        for rep in range(2):
            tot = eng.compare_text(w["order"], w["hi"], w["dpass"], w["drev"], w["tables"], host.NullSink(), rank, world)
        print(json.dumps({"world": world, "rank": rank, "slab_pairs": slab, "steps": tot["steps"], "pairs": tot["pairs"], "records": tot["n_records"],
                          "total_ms": round(tot["total_ms"], 2), "lists_ms": round(tot["lists_ms"], 2), "screen_ms": round(tot["screen_ms"], 2),
                          "host_ms": {k: round(v, 1) for k, v in tot["host_ms"].items()}}), flush=True)
