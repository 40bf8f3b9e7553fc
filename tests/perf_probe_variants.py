"""A/B of the engine parameters on full-size configs, one process per box: every variant must reproduce the pinned
CRC-32 of the config's tempfile text (bench.KNOWN_TEXT_CRC).  python tests/perf_probe_variants.py 5 2 3 4 6"""
import json
import os
import sys
import time
import zlib

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import bench  # noqa: E402
from amplicon_sorter_b200.engine import Engine  # noqa: E402


class CrcSink:
    path, base = None, 0

    def __init__(self):
        self.crc, self.bytes = 0, 0

    def seek(self, offset):
        pass

    def __call__(self, chunk):
        self.crc = zlib.crc32(chunk.data, self.crc)
        self.bytes += len(chunk.data)
        chunk.release()

    def close(self):
        pass


VARIANTS = [
    ("default", {}),
    ("prune_rows=0", {"prune_rows": 0}),
    ("two_rows=0", {"two_rows": 0}),
    ("class_sort=0", {"class_sort": 0}),
    ("list_path=0", {"list_path": 0}),
    ("all off", {"two_rows": 0, "class_sort": 0, "list_path": 0, "prune_rows": 0}),
]
DEFAULTS = {"two_rows": 1, "class_sort": 1, "list_path": 1, "prune_rows": 1}

cfgs = [int(x) for x in sys.argv[1:]] or [5, 2]
with Engine(0) as eng:
    for cfg in cfgs:
        w = bench.make_workload(cfg, 1.0)
        eng.upload_reads(w["buf"], w["offs"])
        for name, params in VARIANTS:
            if cfg in (4, 5, 6) and name == "list_path=0":
                continue  # the pivot bound decides most pairs there: the switch has no effect
            for k, v in {**DEFAULTS, **params}.items():
                eng.set_param(k, float(v))
            best, tot, sink = None, None, None
            for rep in range(3):
                sink = CrcSink()
                t0 = time.perf_counter()
                tot = eng.compare_text(w["order"], w["hi"], w["dpass"], w["drev"], w["tables"], sink)
                dt = (time.perf_counter() - t0) * 1e3
                best = dt if best is None or (rep > 0 and dt < best) else best
            ok = sink.crc == bench.KNOWN_TEXT_CRC.get(cfg)
            print(json.dumps({"cfg": cfg, "variant": name, "ms_per_job": round(best, 2), "Mpairs_per_s": round(w["tl"] / best / 1e3, 1),
                              "crc_ok": ok, "crc": sink.crc, "records": tot["n_records"], "steps": tot["steps"],
                              "device_ms": round(tot["total_ms"], 2), "lists_ms": round(tot["lists_ms"], 2), "screen_ms": round(tot["screen_ms"], 2),
                              "cluster_ms": round(tot["cluster_ms"], 2), "pruned": tot["pruned_pairs"],
                              "wu_per_pair": round(tot["word_updates"] / max(tot["pairs"], 1), 1),
                              "live": round(tot["useful_word_updates"] / max(tot["word_updates"], 1), 3)}), flush=True)
