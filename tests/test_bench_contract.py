"""bench.py's reference arm (the oracle port timed on the host cores) keeps the driver's JSON contract.
Runs on CPU: the arm never touches the GPU library.  The mole_b200 arm needs a B200 and is exercised by the
driver; here we only check that it refuses to run without a device instead of falling back."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BENCH = os.path.join(ROOT, "bench.py")


def _run(extra, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, BENCH] + extra, cwd=ROOT, env=e, capture_output=True, text=True, timeout=600)


def test_reference_arm_prints_one_contract_line():
    r = _run(["--impl", "reference", "--gpus", "1", "--steps", "2", "--warmup", "1", "--ref-walkers", "64", "--ref-sweeps", "20"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference"
    # BASELINE.json: "VMC walker-steps/sec incl. local energy ..."
    assert d["metric"] == "vmc_walker_steps_per_sec_incl_local_energy"
    assert d["unit"] == "walker-steps/s" and d["higher_is_better"] is True and d["scaling"] == "weak"
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 1
    assert d["dtype"] == "f64" and d["data"] == "synthetic" and d["vs_baseline"] is None
    assert d["value"] > 0 and d["ms_per_step"] > 0
    # value is the walker-steps of the bounded sample over the timed steps
    assert abs(d["value"] - 64 * 20 / (d["ms_per_step"] * 1e-3)) <= 1e-6 * d["value"]
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "64 walkers x 20 sweeps" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


def test_reference_arm_other_ranks_exit_without_work():
    r = _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"], env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_product_arm_refuses_to_run_without_a_device():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is present: the product arm is the driver's bench run")
    r = _run(["--steps", "1", "--warmup", "3", "--no-cpu-baseline"])
    assert r.returncode != 0
    assert not any(l.startswith("{") for l in r.stdout.splitlines())
