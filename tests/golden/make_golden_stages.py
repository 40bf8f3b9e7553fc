"""Generates tests/golden/s_*.json.gz: what the UNMODIFIED reference writes in the two "next" stages
(SURVEY 8(f) rows 1 and 2):

  s_consensuslist_*  <group>.tmp as written by process_consensuslist (amplicon_sorter.py:1627-1690) driving its
                     worker similarity_species (:1692-1715) through do_parallel (:1160-1203) at nprocesses = 1,
                     for the similarity levels 0.95 / 0.94 / 0.88 (0.94 - 0.01 = 0.9299999999999999 in float, :1700);
  s_iden_consensus_* consensus.tmp as written by do_parallel (:1160-1203) driving iden_consensus (:1139-1158) on a
                     spool file of [A1, A2, y, z] entries laid out the way comp_consensus_groups builds it (:1268-1296),
                     with nested amplicons (-ldc 20, :1281), reverse-complemented consensuses and N / IUPAC symbols.

The reference's functions run as they are (amplicon_sorter_b200.launcher.load_reference on top of oracle/shims: the
in-repo edlib / Bio stand-ins, because neither package is installed here and edlib's distances are defined by the
DP the shim is tested against).  Workers are the reference's own forked processes.

Run in the container that has /root/reference; the fixtures travel, the reference does not.
    PYTHONHASHSEED=0 python tests/golden/make_golden_stages.py
"""
import gzip
import json
import os
import pickle
import shutil
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "shims"))
from amplicon_sorter_b200 import launcher, synth  # noqa: E402

REF = "/root/reference/amplicon_sorter.py"
ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
META = {"reference": "avierstr/amplicon_sorter amplicon_sorter.py (unmodified, /root/reference) on oracle/shims",
        "harness": "tests/golden/make_golden_stages.py, PYTHONHASHSEED=0, nprocesses=1"}


def consensuslist_case(seed, n_reads, L, similar):
    """Reads of one gene group: 5 species at 1-9 % from an ancestor plus a nested (shorter) amplicon and an unrelated
    contaminant, 3 % read error, either strand, a few N.  Four sub-groups already exist (members + consensus last)."""
    rng = np.random.default_rng(seed)
    anc = rng.integers(0, 4, L, dtype=np.uint8)
    species = [synth.diverge(rng, anc, f) for f in (0.01, 0.03, 0.05, 0.07, 0.09)]
    species.append(species[0][int(0.06 * L):int(0.97 * L)].copy())   # nested amplicon: inside the 5 % window of some reads only
    species.append(rng.integers(0, 4, L, dtype=np.uint8))            # unrelated
    comparelist2 = []
    for i in range(n_reads):
        t = species[i % len(species)]
        r = synth.mutate(rng, t, sub=0.015, ins=0.0075, dele=0.0075)
        if rng.random() < 0.5:
            r = synth.revcomp_codes(r)
        s = ACGT[r].copy()
        if i % 17 == 5:
            s[int(rng.integers(0, s.size))] = ord("N")
        comparelist2.append([f"r{i}", s.tobytes().decode(), "u", i])
    grouplist = []
    for g in range(4):
        members = [str(i) for i in range(g, 4 * len(species), len(species))][:4]
        cons = ACGT[species[g]].tobytes().decode()
        if g == 2:
            cons = cons[:30] + "R" + cons[31:90] + "N" + cons[91:]   # ambiguity codes in a consensus (:277 writes them)
        if g == 3:
            cons = ACGT[synth.revcomp_codes(species[g])].tobytes().decode()  # a consensus on the other strand
        grouplist.append(members + [cons])
    indexes = {str(i) for i in range(n_reads) if i % 11 != 3}       # the gene group does not hold every read
    return {"similar": similar, "indexes": sorted(indexes, key=int), "grouplist": grouplist,
            "comparelist2": [[r[0], r[1], r[2], r[3]] for r in comparelist2]}


def run_consensuslist(ns, case, work, name):
    out = os.path.join(work, name)
    os.makedirs(out)
    ns["args"] = types.SimpleNamespace(outputfolder=out, nprocesses=1)
    ns["comparelist2"] = [list(r) for r in case["comparelist2"]]
    ns["similar"] = case["similar"]
    ns["process_consensuslist"](set(case["indexes"]), [list(g) for g in case["grouplist"]], name + "_0.group")
    p = os.path.join(out, name + "_0.tmp")
    return open(p).read() if os.path.exists(p) else None


def iden_case(seed, n, L, length_diff_c):
    """Consensus sequences of n groups: full and nested amplicons of a few loci (nested = inside the -ldc window),
    each on a random strand, light consensus error, one with N and one with an IUPAC code; the spool entries are the
    pairs comp_consensus_groups would write (:1275-1286: all position < position2 inside the length window)."""
    rng = np.random.default_rng(seed)
    loci = [rng.integers(0, 4, L, dtype=np.uint8) for _ in range(4)]
    cons = []
    for i in range(n):
        t = loci[i % len(loci)]
        if i % 3 == 1:
            t = t[int(0.065 * L):int(0.935 * L)]                      # nested amplicon (870 of 1000)
        r = synth.mutate(rng, t, sub=0.01 * (i % 5), ins=0.002, dele=0.002, homopolymer_boost=1.0)
        if rng.random() < 0.5:
            r = synth.revcomp_codes(r)
        cons.append(ACGT[r].tobytes().decode())
    cons[2] = cons[2][:25] + "N" + cons[2][26:]
    cons[5] = cons[5][:40] + "Y" + cons[5][41:]
    todolist = []
    for y in range(n - 1):
        for z in range(y + 1, n):
            A1, A2 = cons[y], cons[z]
            if len(A1) * length_diff_c < len(A2) or len(A2) * length_diff_c < len(A1):
                continue
            todolist.append([A1, A2, y, z])
    return {"length_diff_c": length_diff_c, "todolist": todolist}


def run_iden(ns, case, work, name):
    out = os.path.join(work, name)
    os.makedirs(out)
    with open(os.path.join(out, "file_0.todo"), "wb") as wf:
        pickle.dump(case["todolist"], wf)
    consensus_tempfile = os.path.join(out, "consensus.tmp")          # :1212
    ns["do_parallel"](out, 1, consensus_tempfile, ns["iden_consensus"], "...comparing consensuses ", "_")
    return open(consensus_tempfile).read()


def save(name, fixture):
    with gzip.open(os.path.join(HERE, name + ".json.gz"), "wt", compresslevel=9) as f:
        json.dump(dict(META, name=name, **fixture), f)


def main():
    if os.environ.get("PYTHONHASHSEED") != "0":
        raise SystemExit("run with PYTHONHASHSEED=0")
    ns, _ = launcher.load_reference(REF)
    work = tempfile.mkdtemp(prefix="golden_stages_")
    try:
        for similar, seed in ((0.95, 11), (0.94, 12), (0.88, 13)):
            name = "s_consensuslist_%d" % round(similar * 100)
            case = consensuslist_case(seed, n_reads=154, L=400, similar=similar)
            text = run_consensuslist(ns, case, work, name)
            save(name, dict(case, group_tmp=text))
            print(name, "lines", 0 if text is None else text.count("\n"))
        for ldc, seed in ((1.08, 21), (1.2, 22)):
            name = "s_iden_consensus_ldc%d" % round((ldc - 1) * 100)
            case = iden_case(seed, n=26, L=400, length_diff_c=ldc)
            text = run_iden(ns, case, work, name)
            save(name, dict(case, consensus_tmp=text))
            print(name, "entries", len(case["todolist"]), "lines", text.count("\n"))
    finally:
        shutil.rmtree(work, ignore_errors=True)


if __name__ == "__main__":
    main()
