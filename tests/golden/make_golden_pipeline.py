"""Generates tests/golden/p_pipeline.json: the user-visible output files of a COMPLETE run of the UNMODIFIED reference
script (amplicon_sorter.py on oracle/shims, -np 1, PYTHONHASHSEED=0, random.seed(0)) on reduced BASELINE configs --
per-species fasta, <stem>_consensussequences.fasta, consensusfile.fasta, results.csv (the contract
test).  results.txt is left out (dates and paths).  tests/test_pipeline_contract.py then runs the SAME command with the
stages of this repo swapped in (launcher + oracle-backed engine on CPU, launcher + CUDA engine on the GPU box) and
compares file by file.

Run in the container that has /root/reference (each run takes minutes: the script sleeps between its stages):
    python tests/golden/make_golden_pipeline.py
"""
import hashlib
import json
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from amplicon_sorter_b200 import synth  # noqa: E402

REF = "/root/reference/amplicon_sorter.py"
# name -> (BASELINE config, scale, extra CLI); the CLI of the config comes from synth.make_config
CASES = {"p_cfg1": (1, 0.3), "p_cfg4": (4, 0.012)}


def case_input(name, folder):
    cfg, scale = CASES[name]
    reads, _, cli = synth.make_config(cfg, scale=scale)
    path = os.path.join(folder, name + ".fastq")
    synth.write_fastq(path, reads)
    return path, [a if a != "-np" else a for a in cli], len(reads)


def digest_folder(folder):
    out = {}
    for fn in sorted(os.listdir(folder)):
        p = os.path.join(folder, fn)
        if fn == "results.txt" or not os.path.isfile(p):
            continue
        data = open(p, "rb").read()
        out[fn] = {"sha256": hashlib.sha256(data).hexdigest(), "bytes": len(data)}
        if fn.endswith(".csv"):
            out[fn]["text"] = data.decode()
    return out


def launch(script, stage, fastq, cli, outdir, np_="1", extra_env=None):
    env = dict(os.environ, PYTHONHASHSEED="0", **(extra_env or {}))
    cmd = [sys.executable, os.path.join(ROOT, "oracle", "run_reference.py"), "--script", script, "--stage", stage, "--",
           "-i", fastq, "-o", outdir] + [a for a in cli if a not in ("-np", "1")] + ["-np", np_]
    return subprocess.Popen(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)


def main():
    work = tempfile.mkdtemp(prefix="golden_pipeline_")
    fixture = {"reference": "avierstr/amplicon_sorter amplicon_sorter.py (unmodified, /root/reference) on oracle/shims",
               "harness": "oracle/run_reference.py --stage reference, PYTHONHASHSEED=0, random.seed(0), -np 1", "cases": {}}
    try:
        procs = {}
        for name in CASES:
            fastq, cli, n = case_input(name, work)
            procs[name] = (launch(REF, "reference", fastq, cli, os.path.join(work, name + "_out")), cli, n)
        for name, (p, cli, n) in procs.items():
            log, _ = p.communicate()
            if p.returncode != 0:
                raise SystemExit(f"{name}: reference run failed\n{log[-3000:]}")
            files = digest_folder(os.path.join(work, name + "_out"))
            fixture["cases"][name] = {"config": CASES[name][0], "scale": CASES[name][1], "reads": n, "cli": cli, "files": files}
            print(name, n, "reads ->", len(files), "files:", ", ".join(files))
    finally:
        shutil.rmtree(work, ignore_errors=True)
    with open(os.path.join(HERE, "p_pipeline.json"), "w") as f:
        json.dump(fixture, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
