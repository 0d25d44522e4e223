"""Series statistics, checkpoint/restart (8(f)4) and the block-level Log callback.

CPU: the oracle's restatement of scripts/statfor.rs against golden vectors produced by the reference's
own scripts/statfor.py (tests/golden/make_statfor_golden.py).  GPU: mole_series_analyze per walker
against the oracle on the same device-resident series; bit-identical continuation after save/load."""
import json
import os

import numpy as np
import pytest

from common import SEED0

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "statfor_golden.json")))


@pytest.fixture(scope="module")
def orc():
    import oracle
    oracle.build()
    return oracle


@pytest.mark.parametrize("case", GOLD["cases"], ids=[c["name"] for c in GOLD["cases"]])
def test_oracle_statfor_matches_reference_script(orc, case):
    x = np.array(case["series"])
    r = orc.statfor(x)
    assert abs(r["average"] - case["mean"]) <= 1e-14 * max(1, abs(case["mean"]))
    assert abs(r["variance"] - case["variance"]) <= 1e-13 * case["variance"]
    assert len(r["corr"]) == len(case["corr"]) == min(200, x.size - 1)
    assert np.allclose(r["corr"], case["corr"], rtol=2e-10, atol=2e-10)      # corr.out carries 10 digits
    for k in ("tcorr", "n_eff", "sigma"):
        assert abs(r[k] - case[k]) <= 1e-11 * abs(case[k])


def test_oracle_blocking_is_statfor_rs(orc):
    """statfor.rs:58-81 restated independently with numpy: chunks() keeps the ragged last chunk and the
    divisor uses ndata // size (this differs from statfor.py's array_split, so it is pinned here)."""
    x = np.random.default_rng(3).normal(size=1013).cumsum() * 0.05
    r = orc.statfor(x)
    large = x.size // 20
    step = max(large // 100, 1)
    sizes = list(range(1, large + 1, step))
    assert list(r["block_sizes"]) == sizes
    for size, err in zip(sizes, r["block_errors"]):
        m = np.array([x[b:b + size].mean() for b in range(0, x.size, size)])
        ref = np.sqrt(((m ** 2).mean() - m.mean() ** 2) / (x.size // size - 1))
        assert abs(err - ref) <= 1e-12 * ref
    assert orc.statfor_block_sizes(5000)[:3].tolist() == [1, 3, 5] and orc.statfor_block_sizes(19).size == 0


# ------------------------------------------------------------------------------------------------ GPU
@pytest.fixture(scope="module")
def mole():
    import mole_b200
    return mole_b200


def _he(mole, W, seed=SEED0):
    wf = mole.HeliumAtomWaveFunction(1.69)
    op = mole.ElectronicHamiltonian.from_ions([[0, 0, 0]], [2])
    met = mole.MetropolisDiffuse.from_rng(0.1, seed)
    ens = mole.Ensemble(W, 2, seed)
    ens.init_uniform()
    return wf, op, met, ens


@pytest.mark.gpu
@pytest.mark.parametrize("W,n,drop", [(37, 260, False), (200, 57, True), (1030, 401, False)])
def test_series_analyze_matches_oracle_per_walker(mole, orc, W, n, drop):
    wf, op, met, ens = _he(mole, W)
    ens.sweep(wf, met, op, n_sweeps=20, observables=0)
    tr = ens.sweep(wf, met, op, n_sweeps=n, observables=mole.ffi.OBS_ENERGY, traces=("energy",), keep_series=True)["energy"]
    assert ens.series_length() == n
    assert np.array_equal(ens.series_get(W - 1), tr[W - 1])           # the kept series IS the trace
    m = n - 1 if drop else n
    sizes = orc.statfor_block_sizes(m)
    assert np.array_equal(mole.series_block_sizes(m), sizes)
    if sizes.size == 0:
        sizes = np.array([1, 2, 5], dtype=np.int32)
    got = ens.series_analyze(block_sizes=sizes, drop_last=drop, per_walker=True)
    keys = ("average", "variance", "tcorr", "n_eff", "sigma")
    ref = [orc.statfor(tr[w][:m], sizes) for w in range(W)]
    rs = np.array([[r[k] for k in keys] for r in ref])
    rc = np.array([r["corr"] for r in ref])
    rb = np.array([r["block_errors"] for r in ref])
    assert np.allclose(got["per_walker"][:, :2], rs[:, :2], rtol=1e-12, atol=0)
    assert np.allclose(got["per_walker_corr"], rc, rtol=0, atol=1e-11)
    # tcorr folds up to 200 correlations; a corr_i within rounding of zero may flip the cut-off f
    close = np.isclose(got["per_walker"][:, 2], rs[:, 2], rtol=1e-9)
    assert close.mean() > 0.99
    assert np.allclose(got["per_walker"][close][:, 3:], rs[close][:, 3:], rtol=1e-9)
    assert np.allclose(got["per_walker_block_errors"], rb, rtol=1e-9, atol=1e-15)
    # walker means, reduced on the device
    for i, k in enumerate(keys):
        assert abs(got[k] - got["per_walker"][:, i].mean()) <= 1e-12 * abs(got[k])
    assert np.allclose(got["corr"], got["per_walker_corr"].mean(axis=0), rtol=0, atol=1e-13)
    assert np.allclose(got["block_errors"], got["per_walker_block_errors"].mean(axis=0), rtol=1e-12)


@pytest.mark.gpu
def test_series_append_and_text_format(mole, orc, tmp_path):
    wf, op, met, ens = _he(mole, 16)
    a = ens.sweep(wf, met, op, n_sweeps=30, observables=mole.ffi.OBS_ENERGY, traces=("energy",), keep_series=True)["energy"]
    b = ens.sweep(wf, met, op, n_sweeps=45, observables=mole.ffi.OBS_ENERGY, traces=("energy",), append_series=True)["energy"]
    assert ens.series_length() == 75
    full = np.concatenate([a, b], axis=1)
    assert np.array_equal(ens.series_get(3), full[3])
    p = tmp_path / "energy.out"
    ens.series_write_text(3, p)
    back = np.array([float(line.split()[0]) for line in open(p)])      # scripts/statfor.py:17-19 read_data
    assert np.array_equal(back, full[3])
    ens.sweep(wf, met, op, n_sweeps=10, observables=mole.ffi.OBS_ENERGY, keep_series=True)   # keep replaces
    assert ens.series_length() == 10
    ens.series_clear()
    assert ens.series_length() == 0
    with pytest.raises(mole.MoleError) as ei:
        ens.series_analyze()
    assert ei.value.kind == "EmptyCacheError"


@pytest.mark.gpu
def test_checkpoint_restart_continues_bit_identically(mole, tmp_path):
    wf, op, met, ens = _he(mole, 300, seed=bytes(range(32)))
    obs = mole.ffi.OBS_ENERGY | mole.ffi.OBS_PGRAD | mole.ffi.OBS_WFVALUE
    ens.sweep(wf, met, op, n_sweeps=37, block_size=10, observables=obs)      # leaves a partial block of 7
    ck = tmp_path / "ens.ckpt"
    ens.save(ck)
    t1 = ens.sweep(wf, met, op, n_sweeps=43, block_size=10, observables=obs, traces=("energy", "accept"))
    acc1, x1 = ens.acc_get(), ens.get_configs()
    # a fresh ensemble with a different seed and state, restored from the file
    ens2 = mole.Ensemble(300, 2, SEED0)
    ens2.init_normal(2.0)
    ens2.load(ck)
    assert ens2.step == 37
    t2 = ens2.sweep(wf, met, op, n_sweeps=43, block_size=10, observables=obs, traces=("energy", "accept"))
    acc2 = ens2.acc_get()
    assert np.array_equal(t1["accept"], t2["accept"]) and np.array_equal(t1["energy"], t2["energy"])
    assert np.array_equal(x1, ens2.get_configs())
    for f, _ in mole.ffi.AccHost._fields_[:10]:
        assert getattr(acc1, f) == getattr(acc2, f), f
    assert list(acc1.sum_o) == list(acc2.sum_o) and list(acc1.sum_oo) == list(acc2.sum_oo)
    # shape mismatch and garbage files are refused
    with pytest.raises(mole.MoleError) as ei:
        mole.Ensemble(299, 2, SEED0).load(ck)
    assert ei.value.kind == "ShapeError"
    bad = tmp_path / "bad.ckpt"
    bad.write_bytes(b"not a checkpoint")
    with pytest.raises(mole.MoleError):
        ens2.load(bad)


@pytest.mark.gpu
def test_dmc_checkpoint_restart(mole, tmp_path):
    seed = bytes([1] * 32)
    wf = mole.STO(0.9)
    op = mole.ElectronicHamiltonian.from_ions([[0, 0, 0]], [1])
    met = mole.MetropolisDiffuse.from_rng(0.025, seed)
    dmc = mole.DmcRunner.new(wf, 512, -0.45, op, met, mole.SRBrancher.new(), identical_start=False)
    ens = dmc.ensemble
    for _ in range(5):
        ens.dmc_step(wf, met, op, 0.025, -0.45); ens.branch(mole.ffi.BRANCH_SR)
    ck = tmp_path / "dmc.ckpt"
    ens.save(ck)
    a = [ens.dmc_step(wf, met, op, 0.025, -0.45) for _ in range(1)]
    ens.branch(mole.ffi.BRANCH_SR)
    a.append(ens.dmc_step(wf, met, op, 0.025, -0.45))
    dmc2 = mole.DmcRunner.new(wf, 512, -0.45, op, met, mole.SRBrancher.new(), identical_start=True)
    e2 = dmc2.ensemble
    e2.load(ck)
    b = [e2.dmc_step(wf, met, op, 0.025, -0.45)]
    e2.branch(mole.ffi.BRANCH_SR)
    b.append(e2.dmc_step(wf, met, op, 0.025, -0.45))
    assert a == b
    assert np.array_equal(ens.get_weights(), e2.get_weights()) and np.array_equal(ens.get_configs(), e2.get_configs())


@pytest.mark.gpu
def test_runner_log_callback_sees_block_reductions(mole, capsys):
    wf, op, met, ens = _he(mole, 64)
    obs = mole.ffi.OBS_ENERGY | mole.ffi.OBS_KINETIC | mole.ffi.OBS_WFVALUE
    ens.snapshot()
    seen = []

    def log(d):
        seen.append(d)
        return "Energy: %.6f" % d["block_energy"] if d["block_nr"] == 2 else ""

    ens.run_logged(wf, met, op, steps=53, block_size=10, log=log, observables=obs, keep_series=True)
    assert [d["block_nr"] for d in seen] == [1, 2, 3, 4] and ens.step == 50
    assert capsys.readouterr().out.strip() == "Energy: %.6f" % seen[1]["block_energy"]   # printed once per block if non-empty
    # the same run in one launch with traces: blocks of the trace reproduce the logged reductions
    ens.restore(); ens.step = 0; ens.acc_reset()
    tr = ens.sweep(wf, met, op, n_sweeps=50, n_discard=10, block_size=10, observables=obs, traces=("energy", "kinetic", "wfvalue", "accept"))
    for b, d in enumerate(seen):
        sl = slice(10 * b, 10 * (b + 1))
        assert d["n_samples"] == 640 and d["block_size"] == 10
        assert abs(d["block_energy"] - tr["energy"][:, sl].mean()) < 1e-11
        assert abs(d["running_energy"] - tr["energy"][:, :10 * (b + 1)].mean()) < 1e-11
        assert abs(d["block_kinetic"] - tr["kinetic"][:, sl].mean()) < 1e-11
        assert abs(d["block_wfvalue"] - tr["wfvalue"][:, sl].mean()) < 1e-12
        assert abs(d["acceptance"] - tr["accept"][:, 10 * (b + 1):10 * (b + 2)].mean()) < 1e-12
    # Runner + logger object (reference API shape: Runner::new(sampler, logger))
    class Logger:
        def __init__(self): self.n = 0
        def log(self, d): self.n += 1; return ""
    lg = Logger()
    s = mole.Sampler(wf, met, {"Energy": op}, n_walkers=32)
    mole.Runner(s, lg).run(40, 10, traces=False)
    assert lg.n == 3
    with pytest.raises(mole.MoleError):
        ens.run_logged(wf, met, op, steps=15, block_size=10, log=log)


@pytest.mark.gpu
@pytest.mark.parametrize("W", [700, 5000])
def test_dmc_block_equals_step_by_step(mole, W):
    """mole_dmc_block (no host reads, fused scan, device-side normalisation) == n x (mole_dmc_step, mole_branch)."""
    seed = bytes(range(1, 33))
    wf = mole.STO(0.9)
    op = mole.ElectronicHamiltonian.from_ions([[0, 0, 0]], [1])
    met = mole.MetropolisDiffuse.from_rng(0.025, seed)
    a = mole.DmcRunner.new(wf, W, -0.45, op, met, mole.SRBrancher.new(), identical_start=False).ensemble
    b = mole.DmcRunner.new(wf, W, -0.45, op, met, mole.SRBrancher.new(), identical_start=False).ensemble
    ea = []
    for _ in range(23):
        swe, sw = a.dmc_step(wf, met, op, 0.025, -0.47)
        ea.append(swe / sw)
        a.branch(mole.ffi.BRANCH_SR)
    eb = b.dmc_block(wf, met, op, mole.ffi.BRANCH_SR, 0.025, -0.47, 23)
    assert np.array_equal(np.array(ea), eb)
    assert a.step == b.step == 23
    assert np.array_equal(a.get_configs(), b.get_configs()) and np.array_equal(a.get_weights(), b.get_weights())
    assert np.array_equal(a.branch_sources(), b.branch_sources())
    # SimpleBranching goes through the step-by-step path inside mole_dmc_block
    es = b.dmc_block(wf, met, op, mole.ffi.BRANCH_SIMPLE, 0.025, -0.47, 3)
    assert np.isfinite(es).all() and b.step == 26


def _dmc_case(mole, name):
    if name == "sto":
        return mole.STO(0.9), mole.ElectronicHamiltonian.from_ions([[0, 0, 0]], [1]), 1, -0.47
    if name == "gaussian":
        return mole.GaussianWaveFunction(1.3), mole.ElectronicHamiltonian.from_ions([[0, 0, 0]], [1]), 1, -0.42
    if name == "h2":
        return (mole.HydrogenMoleculeWaveFunction(1.4, [0.5]),
                mole.ElectronicHamiltonian.from_ions([[-0.7, 0, 0], [0.7, 0, 0]], [1, 1]), 2, -1.1)
    if name == "he":
        return mole.HeliumAtomWaveFunction(1.6875), mole.ElectronicHamiltonian.from_ions([[0, 0, 0]], [2]), 2, -2.85
    ions = [[-1.25, 0, 0], [1.25, 0, 0]]
    if name == "h2p":
        return mole.H2WF(2.5, 1.0), mole.ElectronicHamiltonian.from_ions(ions, [1, 1]), 1, -0.55
    if name == "lcao_h2p":
        return (mole.SingleDeterminant([mole.Orbital([[1.0], [1.0]], mole.Hydrogen1sBasis(ions, [1.0]))]),
                mole.ElectronicHamiltonian.from_ions(ions, [1, 1]), 1, -0.55)
    if name == "lcao_he":
        b1 = mole.Hydrogen1sBasis([[0, 0, 0]], [1.0 / 1.6875])
        return (mole.SpinDeterminantProduct([mole.Orbital([[1.0]], b1), mole.Orbital([[1.0]], b1)], 1),
                mole.ElectronicHamiltonian.from_ions([[0, 0, 0]], [2]), 2, -2.85)
    bs = mole.Hydrogen1sBasis([[-0.7, 0, 0], [0.7, 0, 0]], [0.85])
    wf = mole.SingleDeterminant([mole.Orbital([[1.0], [1.0]], bs), mole.Orbital([[1.0], [-1.0]], bs)])
    return wf, mole.ElectronicHamiltonian.from_ions([[-0.7, 0, 0], [0.7, 0, 0]], [1, 1]), 2, -0.8


@pytest.mark.gpu
@pytest.mark.parametrize("name,W,n_steps", [("sto", 1, 4), ("sto", 129, 7), ("sto", 32768, 40), ("sto", 100003, 9), ("sto", 300000, 6),
                                            ("gaussian", 4097, 10), ("h2", 3000, 11), ("lcao_triplet", 2500, 8),
                                            ("he", 2000, 6), ("h2p", 1500, 6), ("lcao_h2p", 1500, 6), ("lcao_he", 1200, 6)])
def test_dmc_block_persistent_launch_is_bit_identical_to_per_step_launches(mole, name, W, n_steps):
    """the whole-block cooperative launch (mole_dmc_block.cuh: three phases, two grid barriers per step, virtual blocks of
    128 walkers) (mole_dmc_block_select(2)) against the per-step launches (select(1)): configurations, weights, cached
    local energies (through the next block), branching sources, step energies and launch counts.  100 003 walkers need
    more virtual blocks than one co-resident grid holds; 300 000 is the largest population class served by the
    persistent launch (2368 partial rows)."""
    seed = bytes(range(7, 39))
    wf, op, ne, eref = _dmc_case(mole, name)
    met = mole.MetropolisDiffuse.from_rng(0.02, seed)
    ens = []
    for impl in (2, 1):                                      # 2: the persistent launch whatever the population
        e = mole.Ensemble(W, ne, seed)
        e.init_normal(0.8)
        e.dmc_block_select(impl)
        ens.append(e)
    a, b = ens
    ctx = a.ctx
    n0 = ctx.launch_count()
    ea = a.dmc_block(wf, met, op, mole.ffi.BRANCH_SR, 0.02, eref, n_steps)
    n1 = ctx.launch_count()
    eb = b.dmc_block(wf, met, op, mole.ffi.BRANCH_SR, 0.02, eref, n_steps)
    n2 = ctx.launch_count()
    assert n1 - n0 == 1 and n2 - n1 == 3 * n_steps
    assert np.array_equal(ea, eb) and np.isfinite(ea).all()
    assert a.step == b.step == n_steps
    assert np.array_equal(a.get_configs(), b.get_configs()) and np.array_equal(a.get_weights(), b.get_weights())
    assert np.array_equal(a.branch_sources(), b.branch_sources())
    # a second block continues from the carried state (cached E_L, swapped buffers after an odd step count)
    ea2 = a.dmc_block(wf, met, op, mole.ffi.BRANCH_SR, 0.02, eref, 3)
    eb2 = b.dmc_block(wf, met, op, mole.ffi.BRANCH_SR, 0.02, eref, 3)
    assert np.array_equal(ea2, eb2) and np.array_equal(a.get_configs(), b.get_configs())
    assert a.health() == b.health()
    # the default picks the persistent launch only while one virtual block per CTA fits the co-resident grid
    c = mole.Ensemble(W, ne, seed)
    c.init_normal(0.8)
    n3 = ctx.launch_count()
    ec = c.dmc_block(wf, met, op, mole.ffi.BRANCH_SR, 0.02, eref, n_steps)
    assert ctx.launch_count() - n3 == (1 if W <= 40000 else 3 * n_steps) and np.array_equal(ec, eb)
