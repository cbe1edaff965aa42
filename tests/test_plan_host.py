"""CPU: invariants of the launch plan and of the stream-K piece iterator (focal_b200/csrc/plan.h), compiled with g++
and run on the host -- the pieces of all CTAs must tile every launch exactly once, and workspace regions must not overlap."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_plan_and_piece_iterator_invariants(tmp_path):
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("g++ not available")
    exe = str(tmp_path / "test_plan")
    src = os.path.join(ROOT, "tests", "csrc", "test_plan.cpp")
    res = subprocess.run([gxx, "-O1", "-std=c++17", "-o", exe, src], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    run = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert run.returncode == 0 and "plan ok" in run.stdout, run.stdout[-2000:] + run.stderr[-2000:]
