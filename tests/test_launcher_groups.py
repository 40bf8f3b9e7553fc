"""The group stage as the launcher installs it: the user's own update_list / read_indexes re-compiled with the scan of
the tempfile and the grouping statements swapped for this package's calls (launcher.install_group_stage).

  * on a stand-in script written for this test (no reference needed: runs everywhere)
  * on the real reference, when /root/reference is present (build container only), against the golden vectors
"""
import glob
import gzip
import json
import os
import sys
import types

import pytest

from amplicon_sorter_b200 import launcher
from tests.fake_engine import OracleEngine

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/amplicon_sorter.py"

STAND_IN = '''
import os, sys

def SSG(tempfile):
    return -1

def merge_groups(grouplist):
    raise AssertionError("the rewritten functions must not call merge_groups")

def update_list(tempfile):
    outputfolder = args.outputfolder
    templist = []
    try:
        with open(os.path.join(outputfolder, tempfile), 'r') as tf:
            for line in tf:
                templist.append(line.split(':'))
    except FileNotFoundError:
        sys.exit()
    grouplist = []
    for x in templist:
        for s in grouplist:
            if x[0] in s or x[1] in s:
                s.update({x[0], x[1]})
                break
        else:
            grouplist.append({x[0], x[1]})
    grouplist = merge_groups(grouplist)
    return templist, grouplist

def read_indexes(group_filename):
    outputfolder = args.outputfolder
    similar_species_groups = args.similar_species_groups / 100
    indexes = set(open(os.path.join(outputfolder, group_filename)).read().split())
    if group_filename.endswith('nogroup.group'):
        grouplist = []
    else:
        templist = []
        try:
            with open(os.path.join(outputfolder, tempfile), 'r') as tf:
                for line in tf:
                    templist.append(line.split(':'))
        except FileNotFoundError:
            pass
        grouplist = []
        templist.sort(key=lambda x: float(x[2]), reverse=True)
        for x in templist:
            for s in grouplist:
                if x[0] in s or x[1] in s:
                    s.update([x[0], x[1]])
                    break
            else:
                grouplist.append({x[0], x[1]})
        grouplist = merge_groups(grouplist)
    return indexes, templist, grouplist

if __name__ == '__main__':
    pass
'''


def load(path):
    with gzip.open(path, "rt") as f:
        return json.load(f)


def test_rewrite_of_a_stand_in_script(tmp_path):
    script = tmp_path / "stand_in.py"
    script.write_text(STAND_IN)
    ns, _ = launcher.load_reference(str(script))
    fx = load(os.path.join(HERE, "golden", "h_ties_dense.json.gz"))
    (tmp_path / "x_compare.tmp").write_text(fx["compare_tmp"])
    ns["args"] = types.SimpleNamespace(outputfolder=str(tmp_path), similar_species_groups=93)
    ns["tempfile"] = "x_compare.tmp"
    eng = OracleEngine()
    assert sorted(launcher.install_group_stage(ns, lambda: eng)) == ["SSG", "read_indexes", "update_list"]
    assert ns["SSG"]("x_compare.tmp") == fx["ssg"]
    templist, grouplist = ns["update_list"]("x_compare.tmp")
    assert templist == fx["update_list"]["templist"]
    assert [sorted(g, key=int) for g in grouplist] == fx["update_list"]["groups"]
    case = fx["read_indexes"][-1]
    (tmp_path / "x_0.group").write_text("".join(m + "\n" for m in case["members"]))
    ns["args"].similar_species_groups = case["ssg"]
    indexes, templist, grouplist = ns["read_indexes"]("x_0.group")
    assert indexes == set(case["members"]) and templist == case["templist"]
    assert [sorted(g, key=int) for g in grouplist] == case["groups"]
    # a missing file: update_list exits (:1014-1015), read_indexes carries on with nothing (:1392-1393)
    os.remove(tmp_path / "x_compare.tmp")
    with pytest.raises(SystemExit):
        ns["update_list"]("x_compare.tmp")
    assert ns["read_indexes"]("x_0.group")[1:] == ([], [])


def test_unexpected_shape_is_left_alone(tmp_path):
    script = tmp_path / "other.py"
    script.write_text("def update_list(tempfile):\n    return 1\n\ndef SSG(tempfile):\n    return 2\n\nif __name__ == '__main__':\n    pass\n")
    ns, _ = launcher.load_reference(str(script))
    assert launcher.install_group_stage(ns, lambda: OracleEngine()) == ["SSG"]
    assert ns["update_list"]("x") == 1


class Abort(Exception):
    pass


@pytest.mark.skipif(not os.path.exists(REF), reason="the reference is only mounted in the build container")
@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(HERE, "golden", "h_*.json.gz"))), ids=lambda p: os.path.basename(p)[:-8])
def test_rewritten_reference_functions_reproduce_the_golden_vectors(path, tmp_path):
    sys.path.insert(0, os.path.join(os.path.dirname(HERE), "oracle", "shims"))
    try:
        ns, _ = launcher.load_reference(REF)
    finally:
        sys.path.pop(0)
    fx = load(path)
    out = str(tmp_path)
    tmpname = fx["name"] + "_compare.tmp"
    with open(os.path.join(out, tmpname), "w") as f:
        f.write(fx["compare_tmp"])
    with open(os.path.join(out, "results.txt"), "w") as f:
        f.write("- similar_species_groups = Estimate\n")
    ns["args"] = types.SimpleNamespace(outputfolder=out, similar_species_groups="Estimate", nprocesses=1)
    ns["infile"], ns["num_seq"], ns["tempfile"] = fx["name"] + ".fastq", fx["n_reads"], tmpname
    eng = OracleEngine()
    assert sorted(launcher.install_group_stage(ns, lambda: eng)) == ["SSG", "read_indexes", "update_list"]
    cap = {}

    def comp_probe(grouplist):
        cap["templist"] = [list(e) for e in sys._getframe(1).f_locals["templist"]]
        cap["groups"] = [sorted(g, key=int) for g in grouplist]
        raise Abort()

    ns["comp_consensus_groups"] = comp_probe
    ns["merge_groups"] = lambda g: (_ for _ in ()).throw(AssertionError("merge_groups must not run"))
    with pytest.raises(Abort):
        ns["update_list"](os.path.join(out, tmpname))
    assert ns["args"].similar_species_groups == fx["ssg"] == fx["update_list"]["ssg_arg"]
    assert "(Estimated)" in open(os.path.join(out, "results.txt")).read()  # the reference's own bookkeeping still ran (:970-979)
    assert cap["templist"] == fx["update_list"]["templist"] and cap["groups"] == fx["update_list"]["groups"]
    for k, case in enumerate(fx["read_indexes"]):
        gname = f"{fx['name']}_{k}.group"
        with open(os.path.join(out, gname), "w") as f:
            f.write("".join(m + "\n" for m in case["members"]))
        ns["args"].similar_species_groups = case["ssg"]
        seen = {}

        def sort_probe(*a, **kw):
            raise Abort()

        # read_indexes goes on to build consensuses; stop it right after the grouping: comparelist.sort is the next call (:1418)
        ns["comparelist"] = types.SimpleNamespace(sort=sort_probe)
        import builtins
        real_print = builtins.print

        def print_probe(*a, **kw):
            if a and str(a[0]).startswith("--> Number of groups after removing"):
                fr = sys._getframe(1)
                seen["templist"] = [list(e) for e in fr.f_locals["templist"]]
            return real_print(*a, **kw)

        ns["print"] = print_probe
        with pytest.raises(Abort):
            ns["read_indexes"](gname)
        del ns["print"]
        assert seen["templist"] == case["templist"]


def test_small_stages_use_the_local_engine_under_torchrun(tmp_path):
    """dist.ShardedEngine only shards compare_batch; every other stage must get rank 0's own engine."""
    script = tmp_path / "stub.py"
    script.write_text("def process_list(self, tempfile):\n    pass\n\ndef process_consensuslist(indexes, grouplist, group_filename):\n    pass\n\n"
                      "def do_parallel(*a):\n    pass\n\ndef SSG(tempfile):\n    pass\n\nif __name__ == '__main__':\n    pass\n")
    ns, _ = launcher.load_reference(str(script))
    local = OracleEngine()
    facade = types.SimpleNamespace(engine=local, compare_batch=None)  # what dist.ShardedEngine looks like
    seen = []
    import amplicon_sorter_b200.host as host_mod
    real = host_mod.iden_consensus_files
    host_mod.iden_consensus_files = lambda outputfolder, consensus_tempfile, stringx, engine: seen.append(engine)
    try:
        launcher.install_gpu_stage(ns, engine_factory=lambda: facade)

        def iden_consensus():
            pass

        ns["do_parallel"]("out", 1, "c.tmp", iden_consensus, "x", "g")
    finally:
        host_mod.iden_consensus_files = real
    assert seen == [local]


def test_launcher_fails_loudly_without_an_engine(tmp_path, monkeypatch):
    """No CPU fallback: the launcher must stop before the reference's `except Exception: continue` can hide the failure."""
    import amplicon_sorter_b200.engine as engine_mod
    from amplicon_sorter_b200._ffi import EngineError

    def boom(*a, **k):
        raise EngineError(-1, "asb_create failed (no usable CUDA device?)")

    monkeypatch.setattr(engine_mod, "Engine", boom)
    monkeypatch.delenv("WORLD_SIZE", raising=False)
    script = tmp_path / "stub.py"
    script.write_text("def process_list(self, tempfile):\n    pass\n\nif __name__ == '__main__':\n    raise AssertionError('the script must not start')\n")
    with pytest.raises(SystemExit) as e:
        launcher.main(["--script", str(script), "-i", "x.fastq"])
    assert "cannot start the CUDA engine" in str(e.value)


def test_stage_errors_are_not_swallowed_by_the_reference(tmp_path, capsys):
    """an engine failure inside a replaced stage must not become a silently skipped input file.  The
    reference's main loop catches Exception (amplicon_sorter.py:2184); `loud` turns the failure into SystemExit."""
    from amplicon_sorter_b200 import host
    from amplicon_sorter_b200._ffi import EngineError

    class Broken(OracleEngine):
        def upload_reads(self, buf, offs):
            raise EngineError(-3, "out of memory (injected)")

    script = tmp_path / "stub.py"
    script.write_text("import types\nargs = types.SimpleNamespace(outputfolder=%r, similar_genes=80.0)\n"
                      "def process_list(self, tempfile):\n    pass\n\ndef do_parallel(*a):\n    pass\n\n"
                      "if __name__ == '__main__':\n    for f in (1, 2):\n        try:\n            process_list(BATCHES, 'x_compare.tmp')\n"
                      "        except Exception:\n            continue\n" % str(tmp_path))
    ns, main_code = launcher.load_reference(str(script))
    ns["BATCHES"] = [[["a", "ACGT" * 100, "u", 0], ["b", "ACGT" * 100, "u", 1]]]
    launcher.install_gpu_stage(ns, engine_factory=lambda: Broken(), stages={"process_list"})
    with pytest.raises(SystemExit) as e:
        launcher.execute(ns, main_code, [])
    assert "process_list failed" in str(e.value)
    assert "out of memory (injected)" in capsys.readouterr().err
    # ... while the reference's own "nothing to compare" protocol still skips the file quietly (:702-706, :768-772)
    ns["BATCHES"] = [[["a", "A" * 300, "u", 0], ["b", "C" * 400, "u", 1]]]
    open(tmp_path / "results.txt", "w").close()
    launcher.install_gpu_stage(ns, engine_factory=lambda: OracleEngine(), stages={"process_list"})
    launcher.execute(ns, main_code, [])
    assert "No reads to compare" in open(tmp_path / "results.txt").read()
    assert host.NoReadsToCompare.__mro__[1] is Exception


def test_stage_selection_is_parsed_once(monkeypatch):
    monkeypatch.delenv("ASB200_STAGES", raising=False)
    assert launcher.enabled_stages() == set(launcher.ALL_STAGES)
    monkeypatch.setenv("ASB200_STAGES", "process_list, iden_consensus")
    assert launcher.enabled_stages() == {"process_list", "iden_consensus"}  # independent of process_consensuslist
    monkeypatch.setenv("ASB200_STAGES", "process_list,process_consensuslist")
    assert launcher.enabled_stages() == {"process_list", "process_consensuslist"}
    monkeypatch.setenv("ASB200_STAGES", "proces_list")
    with pytest.raises(SystemExit):
        launcher.enabled_stages()
