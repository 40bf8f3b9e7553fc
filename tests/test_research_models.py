"""Keeps the next-round research code honest (CPU only, nothing here is on a product path): the draft of the chained
seed bound (tests/research/chain_profile_draft.cuh, host/device-shared source) must stay ADMISSIBLE against the full DP
matrix, and strong enough to matter."""
import os
import re
import shutil
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_chained_seed_bound_draft_is_admissible(tmp_path):
    exe = str(tmp_path / "chain_check")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(HERE, "research", "chain_profile_check.cpp")])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "NOT ADMISSIBLE" not in out.stdout
    m = re.search(r"admissible on (\d+) related pairs", out.stdout)
    assert m and int(m.group(1)) > 300
    m = re.search(r"dead at column ([0-9.]+)", out.stdout)
    assert m and float(m.group(1)) < 100  # today's 7-mer presence bound: 192
