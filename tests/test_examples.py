"""The reference's examples re-expressed on the GPU path (examples/*.py) run end to end at reference sizes."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(script, *args):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "examples", script), *args], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return r.stdout


def test_custom_operator_example():
    out = _run("custom_operator.py")
    assert out.count("Energy: 1.5") >= 19 and "Energy:     1.5" in out.replace("1.4999999999999", "1.5")


def test_hydrogen_molecule_example_faithful():
    out = _run("hydrogen_molecule.py", "--faithful", "--dmc-iters", "2000")
    assert "STOCHASTIC RECONFIGURATION" in out and "STEEPEST DESCENT" in out and out.count("accept:") == 20
    e = float(out.split("DMC Energy:")[1].split()[0])
    assert -1.35 < e < -1.0                                        # exact -1.17447; 100 walkers x 2000 steps


def test_helium_example_faithful():
    out = _run("helium_atom_singlet.py", "--faithful")
    assert out.count("accept:") == 20
    last_sr = [l for l in out.split("STEEPEST")[0].splitlines() if l.startswith("Energy:")][-1]
    assert -3.1 < float(last_sr.split()[1]) < -2.6                 # alpha -> 27/16, E -> -2.85


def test_dmc_example():
    out = _run("dmc.py", "--dmc-iters", "8000")
    assert "VMC Energy:" in out
    e = float(out.split("DMC Energy:")[1].split()[0])
    assert abs(e + 0.5) < 0.05


def test_lcao_example():
    out = _run("lcao.py", "--walkers", "2048")
    e1 = float(out.split("H2+ LCAO   Energy:")[1].split()[0])
    e2 = float(out.split("He LCAO    Energy:")[1].split()[0])
    assert abs(e1 + 0.5648) < 2e-3 and abs(e2 + 2.8477) < 1e-2      # closed forms: LCAO integrals at R = 2.5; alpha^2 - 27 alpha / 8


def test_lcao_chain_example_large_p():
    out = _run("lcao_chain.py", "--walkers", "2048", "--iterations", "6")
    assert "P = 36 variational parameters" in out and out.count("Energy:") == 6
    e = [float(l.split()[1]) for l in out.splitlines() if l.startswith("Energy:")]
    assert e[-1] < e[0] and "non-finite samples skipped: 0" in out
