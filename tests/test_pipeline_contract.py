"""Contract test: a COMPLETE run of the user's script with this repo's stages swapped in writes the
same user-visible files as the unmodified script -- per-species fasta, <stem>_consensussequences.fasta,
consensusfile.fasta, results.csv.  The expected files were written by the UNMODIFIED reference
(tests/golden/make_golden_pipeline.py -> tests/golden/p_pipeline.json: sha256 per file).

  CPU : launcher + the oracle behind the engine interface (the product's host code end to end, no GPU); needs the
        reference script, i.e. runs where /root/reference exists.
  GPU : launcher + CUDA engine, one case at -np 4 -- the reference forks its worker pools (amplicon_sorter.py:1189-1198)
        while the engine's CUDA context is alive; needs a copy of the script on the box (baseline/_ref/, git-ignored,
        placed there by __graft_entry__.build()).

The script sleeps between its stages, so each run takes about two minutes; the cases run side by side."""
import json
import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden_pipeline as gp  # noqa: E402

SCRIPTS = ["/root/reference/amplicon_sorter.py", os.path.join(ROOT, "baseline", "_ref", "amplicon_sorter.py")]
SCRIPT = next((p for p in SCRIPTS if os.path.isfile(p)), None)
GOLDEN = os.path.join(HERE, "golden", "p_pipeline.json")


def run_cases(stage, tmp_path, np_by_case, extra_env=None):
    with open(GOLDEN) as f:
        fixture = json.load(f)
    procs = {}
    for name, case in fixture["cases"].items():
        fastq, cli, n = gp.case_input(name, str(tmp_path))
        assert n == case["reads"] and cli == case["cli"]
        procs[name] = gp.launch(SCRIPT, stage, fastq, cli, os.path.join(str(tmp_path), name + "_out"), np_by_case.get(name, "1"), extra_env)
    for name, p in procs.items():
        log, _ = p.communicate(timeout=1200)
        assert p.returncode == 0, f"{name}: run failed\n{log[-3000:]}"
        got = gp.digest_folder(os.path.join(str(tmp_path), name + "_out"))
        want = fixture["cases"][name]["files"]
        assert sorted(got) == sorted(want), (name, sorted(got), sorted(want))
        for fn in want:
            assert got[fn]["sha256"] == want[fn]["sha256"], f"{name}: {fn} differs from the file the unmodified reference wrote"
        assert any(fn.endswith("_consensussequences.fasta") for fn in want) and "results.csv" in want
    return fixture


def test_golden_pipeline_fixture_is_committed():
    with open(GOLDEN) as f:
        fixture = json.load(f)
    assert set(fixture["cases"]) == {"p_cfg1", "p_cfg4"} and "unmodified" in fixture["reference"]
    for case in fixture["cases"].values():
        assert len(case["files"]) >= 5


@pytest.mark.skipif(SCRIPT is None, reason="the reference script is not on this box")
@pytest.mark.timeout(1500)
def test_launcher_on_the_oracle_engine_writes_the_reference_files(tmp_path):
    run_cases("oracle", tmp_path, {})


@pytest.mark.gpu
@pytest.mark.skipif(SCRIPT is None, reason="no copy of the reference script on this box (baseline/_ref/)")
@pytest.mark.timeout(1500)
def test_launcher_on_the_cuda_engine_writes_the_reference_files(tmp_path):
    run_cases("gpu", tmp_path, {"p_cfg1": "4"})  # -np 4: worker pools forked after CUDA is up
