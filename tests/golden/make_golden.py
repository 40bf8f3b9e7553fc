"""Generates tests/golden/*.json.gz by running the UNMODIFIED reference script
(/root/reference/amplicon_sorter.py, -np 1, PYTHONHASHSEED=0, random.seed(0)) on top of
oracle/shims, and capturing <stem>_compare.tmp before the script deletes it (:2179).

Run here (the container that has /root/reference); the fixtures travel, the reference does not.
    python tests/golden/make_golden.py
Each fixture holds: the length-filtered records (idx -> SEQ) the script compared, the CLI, the batch composition the script built (record idx
per batch, before process_list sorted them), and the exact bytes of _compare.tmp.
"""
import gzip
import json
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from amplicon_sorter_b200 import synth  # noqa: E402

REF = "/root/reference/amplicon_sorter.py"
ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def case_default():
    reads, _, _ = synth.make_config(1, scale=0.12)  # 120 reads x ~700 bp, 5 templates
    return "g1_default", reads, ["-np", "1"]


def case_random_batches():
    rng = np.random.default_rng(202)
    T = [synth.random_template(rng, int(rng.integers(110, 131))) for _ in range(40)]
    reads, _ = synth._emit(rng, T, synth._split(1100, 40))
    return "g2_random", reads, ["-np", "1", "-ra", "-maxr", "2200", "-min", "100"]


def case_all_mixed():
    rng = np.random.default_rng(203)
    T = [synth.random_template(rng, L) for L in (300, 300, 312, 330, 331, 360)]
    reads, _ = synth._emit(rng, T, synth._split(180, 6), n_frac=0.1)
    reads += [ACGT[rng.integers(0, 4, int(rng.integers(300, 380)))].tobytes() for _ in range(20)]
    return "g3_all_mixed", reads, ["-np", "1", "-a", "-maxr", "200", "-sg", "85"]


def run_case(name, reads, cli):
    work = tempfile.mkdtemp(prefix="golden_")
    try:
        fq = os.path.join(work, name + ".fastq")
        synth.write_fastq(fq, reads)
        env = dict(os.environ, PYTHONHASHSEED="0")
        cmd = [sys.executable, os.path.join(ROOT, "oracle", "run_reference.py"), "--script", REF, "--stage", "reference",
               "--stop-after-genes", "--dump", os.path.join(work, "dump"), "--", "-i", fq, "-o", os.path.join(work, "out"), *cli]
        subprocess.check_call(cmd, env=env, cwd=work, stdout=subprocess.DEVNULL)
        with open(os.path.join(work, "dump", name + "_batches.json")) as f:
            meta = json.load(f)
        with open(os.path.join(work, "dump", name + "_compare.tmp")) as f:
            compare = f.read()
        fixture = {"name": name, "cli": cli, "reference": "avierstr/amplicon_sorter amplicon_sorter.py version 2025-05-28",
                   "harness": "oracle/run_reference.py --stage reference (edlib/Bio shims), PYTHONHASHSEED=0, random.seed(0)",
                   "n_input_reads": len(reads), "records": meta["records"], "similar_genes": meta["similar_genes"],
                   "batches_before": meta["batches_before"], "batches_after": meta["batches_after"], "compare_tmp": compare}
        with gzip.open(os.path.join(HERE, name + ".json.gz"), "wt", compresslevel=9) as f:
            json.dump(fixture, f)
        print(name, len(reads), "reads", len(meta["batches_before"]), "batches", compare.count("\n"), "lines")
    finally:
        shutil.rmtree(work, ignore_errors=True)


if __name__ == "__main__":
    for case in (case_default, case_random_batches, case_all_mixed):
        run_case(*case())
