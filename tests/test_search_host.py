"""The branch-pick search of the SR kernel (mole_b200/csrc/mole_search.h, compiled unchanged into
sr_pick_gather_tiled_kernel) against std::upper_bound on the host: all-alive, mostly-dead, empty-tile, ragged-tile
and N^2-total weight patterns.  CPU only; the GPU parity of the whole branch step is tests/test_gpu_parity.py."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_tiled_pick_is_upper_bound(tmp_path):
    cxx = shutil.which("g++")
    if not cxx:
        pytest.skip("no g++")
    exe = str(tmp_path / "search_check")
    subprocess.check_call([cxx, "-O2", "-std=c++17", "-Wall", "-I", os.path.join(ROOT, "mole_b200", "csrc"),
                           os.path.join(ROOT, "tests", "native", "search_check.cpp"), "-o", exe])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    tag, cases, draws = r.stdout.split()
    assert tag == "ok" and int(cases) >= 200 and int(draws) > 10 ** 6
