"""Generates tests/golden/h*.json.gz: what the UNMODIFIED reference's consumers of <stem>_compare.tmp
compute from a given file -- SSG (amplicon_sorter.py:809-835), update_list's best-hit filter, greedy grouping
and merge_groups (:986-1033, :1057-1086) and the same trio inside read_indexes (:1364-1412).

The reference's functions are executed as they are (loaded with amplicon_sorter_b200.launcher.load_reference
on top of oracle/shims); their local variables are captured from the frame of a patched callee
(comp_consensus_groups / merge_groups), which then aborts the function before it touches anything else.

Run here (the container that has /root/reference); the fixtures travel, the reference does not.
    PYTHONHASHSEED=0 python tests/golden/make_golden_groups.py
"""
import glob
import gzip
import json
import os
import random
import shutil
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "shims"))
from amplicon_sorter_b200 import launcher  # noqa: E402

REF = "/root/reference/amplicon_sorter.py"


class Abort(Exception):
    pass


def synthetic_lines(seed, n_reads, n_lines, lo, hi, step, reverse_frac=0.3):
    """Lines with many tied scores and keys met in scattered order: the reference's leftovers of lower scores
    (k copies of v, then a higher score, leave k-1 copies) depend on exactly this order."""
    rng = random.Random(seed)
    out = []
    for _ in range(n_lines):
        a, b = rng.sample(range(n_reads), 2)
        iden = round(rng.randrange(lo, hi + 1, step) / 1000, 3)
        out.append(f"{a}:{b}:{iden}" + (":reverse" if rng.random() < reverse_frac else ""))
    return "\n".join(out) + "\n"


def cases():
    for path in sorted(glob.glob(os.path.join(HERE, "g*.json.gz"))):
        with gzip.open(path, "rt") as f:
            fx = json.load(f)
        yield "h_" + fx["name"], fx["compare_tmp"], max(int(k) for k in fx["records"]) + 1
    yield "h_ties_dense", synthetic_lines(1, 60, 3000, 800, 1000, 25), 60          # 9 distinct scores: ties everywhere
    yield "h_ties_sparse", synthetic_lines(2, 400, 6000, 700, 1000, 1), 400          # few ties, long tail of keys
    yield "h_two_scores", synthetic_lines(3, 30, 1500, 900, 950, 50), 30             # worst case of the leftover rule
    yield "h_single_line", "3:7:0.912\n", 8


def run_case(ns, name, text, n_reads, work):
    out = os.path.join(work, name)
    os.makedirs(out)
    tmpname = name + "_compare.tmp"
    with open(os.path.join(out, tmpname), "w") as f:
        f.write(text)
    ns["args"] = types.SimpleNamespace(outputfolder=out, similar_species_groups="Estimate", nprocesses=1)
    ns["infile"] = name + ".fastq"
    ns["num_seq"] = n_reads
    ns["tempfile"] = tmpname
    with open(os.path.join(out, "results.txt"), "w") as f:
        f.write("- similar_species_groups = Estimate\n")
    fixture = {"name": name, "reference": "avierstr/amplicon_sorter amplicon_sorter.py version 2025-05-28",
               "harness": "tests/golden/make_golden_groups.py, PYTHONHASHSEED=0", "compare_tmp": text, "n_reads": n_reads}
    # ---- SSG
    fixture["ssg"] = ns["SSG"](tmpname)
    # ---- update_list: capture templist (filtered + sorted) and the grouping
    cap = {}
    real_merge = ns["merge_groups"]

    def merge_probe(grouplist):
        cap["greedy"] = [sorted(g, key=int) for g in grouplist]
        res = real_merge(grouplist)
        cap["merged"] = [sorted(g, key=int) for g in res]
        return res

    def comp_probe(grouplist):
        cap["templist"] = [list(e) for e in sys._getframe(1).f_locals["templist"]]
        raise Abort()

    ns["merge_groups"], ns["comp_consensus_groups"] = merge_probe, comp_probe
    try:
        ns["update_list"](os.path.join(out, tmpname))
    except Abort:
        pass
    fixture["update_list"] = {"ssg_arg": ns["args"].similar_species_groups, "templist": cap["templist"],
                              "n_greedy": len(cap["greedy"]), "groups": cap["merged"]}
    # ---- read_indexes: for the two largest groups and one threshold each side of the estimate
    fixture["read_indexes"] = []
    groups = sorted(cap["merged"], key=len, reverse=True)[:2]
    for gi, members in enumerate(groups):
        for ssg in sorted({int(fixture["ssg"] or 90), 93}):
            gname = f"{name}_{gi}.group"
            with open(os.path.join(out, gname), "w") as f:
                f.write("".join(m + "\n" for m in members))
            ns["args"].similar_species_groups = ssg
            cap2 = {}

            def merge_probe2(grouplist, cap2=cap2):
                cap2["templist"] = [list(e) for e in sys._getframe(1).f_locals["templist"]]
                cap2["greedy"] = len(grouplist)
                res = real_merge(grouplist)
                cap2["merged"] = [sorted(g, key=int) for g in res]
                raise Abort()

            ns["merge_groups"] = merge_probe2
            try:
                ns["read_indexes"](gname)
            except Abort:
                pass
            fixture["read_indexes"].append({"members": members, "ssg": ssg, "templist": cap2["templist"],
                                            "n_greedy": cap2["greedy"], "groups": cap2["merged"]})
    ns["merge_groups"] = real_merge
    with gzip.open(os.path.join(HERE, name + ".json.gz"), "wt", compresslevel=9) as f:
        json.dump(fixture, f)
    print(name, text.count("\n"), "lines: ssg", fixture["ssg"], "templist", len(cap["templist"]), "greedy", len(cap["greedy"]),
          "groups", len(cap["merged"]), "read_indexes cases", len(fixture["read_indexes"]))


def main():
    if os.environ.get("PYTHONHASHSEED") != "0":
        raise SystemExit("run with PYTHONHASHSEED=0")
    ns, _ = launcher.load_reference(REF)
    work = tempfile.mkdtemp(prefix="golden_groups_")
    try:
        for name, text, n_reads in cases():
            run_case(ns, name, text, n_reads, work)
    finally:
        shutil.rmtree(work, ignore_errors=True)


if __name__ == "__main__":
    main()
