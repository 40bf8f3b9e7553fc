"""Seeded harness around the UNMODIFIED reference script (test infrastructure only).

    PYTHONHASHSEED=0 python oracle/run_reference.py --script /root/reference/amplicon_sorter.py \
        [--stage oracle|reference] [--stop-after-genes] [--dump DIR] -- <reference CLI args>

* puts oracle/shims (edlib, Bio.SeqIO stand-ins) on sys.path -- neither package is installed here;
* random.seed(0) before the reference's __main__ body (the script never seeds);
* --stage reference : the script's own process_list (forked workers, sleeps and all) = the oracle run;
  --stage oracle    : amplicon_sorter_b200.host.process_list driven by tests/fake_engine.OracleEngine
                      (the CPU oracle behind the engine interface) -- exercises the drop-in host logic
                      without a GPU;
  --stage gpu       : the product path (needs a GPU);
* --dump DIR : saves <stem>_compare.tmp (the script deletes it at :2179) and the batch composition.
"""
import argparse
import json
import os
import random
import shutil
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "shims"))

from amplicon_sorter_b200 import launcher  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--script", required=True)
    ap.add_argument("--stage", default="reference", choices=["reference", "oracle", "gpu"])
    ap.add_argument("--stop-after-genes", action="store_true")
    ap.add_argument("--dump")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("rest", nargs=argparse.REMAINDER)
    a = ap.parse_args()
    rest = a.rest[1:] if a.rest and a.rest[0] == "--" else a.rest
    if os.environ.get("PYTHONHASHSEED") != "0":
        print("warning: PYTHONHASHSEED is not 0; set iteration order downstream is not reproducible", file=sys.stderr)

    ns, main_code = launcher.load_reference(a.script)
    ns["check_version"] = lambda version: None
    if a.stage == "gpu":
        launcher.install_gpu_stage(ns)
    elif a.stage == "oracle":
        from tests.fake_engine import OracleEngine

        shared = OracleEngine()  # the CPU oracle behind the engine interface: the product's host code, no GPU
        launcher.install_gpu_stage(ns, engine_factory=lambda: shared)

    inner_pl = ns["process_list"]
    inner_sg = ns["sort_groups"]

    def process_list_probe(self, tempfile):
        before = [[rec[3] for rec in d] for d in self]
        records = {str(rec[3]): rec[1] for d in self for rec in d}  # idx (:560-561) -> upper-cased SEQ (:551)
        t0 = time.perf_counter()
        try:
            return inner_pl(self, tempfile)
        finally:
            # wall time of the all-pairs stage alone (the reference's own stage includes >= 8 s of fixed sleeps)
            print("ASB_TIMING process_list seconds=%.3f tl=%s" % (time.perf_counter() - t0, ns.get("tl")), flush=True)
            if a.dump:
                os.makedirs(a.dump, exist_ok=True)
                stem = os.path.basename(tempfile).replace("_compare.tmp", "")
                with open(os.path.join(a.dump, stem + "_batches.json"), "w") as f:
                    json.dump({"batches_before": before, "batches_after": [[rec[3] for rec in d] for d in self],
                               "similar_genes": ns["args"].similar_genes, "records": records}, f)

    def sort_groups_probe():
        if a.dump:
            src = os.path.join(ns["args"].outputfolder, ns["tempfile"])
            if os.path.exists(src):
                shutil.copyfile(src, os.path.join(a.dump, os.path.basename(src)))
        if a.stop_after_genes:
            raise Exception("stop after sort_genes (harness)")  # caught by the per-file handler :2184
        return inner_sg()

    ns["process_list"] = process_list_probe
    ns["sort_groups"] = sort_groups_probe
    random.seed(a.seed)
    launcher.execute(ns, main_code, rest)


if __name__ == "__main__":
    main()
