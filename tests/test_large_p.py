"""The general LCAO Slater-Jastrow kind (MOLE_WF_LCAO_SJ, SURVEY.md 8(f)3: N_e > 2, more than two centres, P up to 44)
and the large-P optimisation moments: the per-sample rows (1, E_L, O_k) contracted into the Gram matrix on the tensor
cores (mma.sync m8n8k4 f64) and on the FP64 vector pipe, the SR step on them, all against the oracle."""
import json
import os

import numpy as np
import pytest

from common import SEED0, cases, random_cfgs, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-10
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "wf_golden.json")))
LSJ = ["lsj_h4", "lsj_h3", "lsj_h8"]


def close(a, b, tol=TOL):
    a, b = np.asarray(a), np.asarray(b)
    scale = np.maximum(np.abs(b), 1e-3 * np.max(np.abs(b)) if b.size else 1.0)
    return np.max(np.abs(a - b) / scale) < tol if a.size else True


@pytest.mark.parametrize("name", LSJ)
def test_lsj_pointwise_and_batched_match_golden_and_oracle(mole, orc, name):
    c = cases()[name]
    wf, op = c["make"](mole)
    assert wf.num_parameters() == c["np"] and wf.num_electrons() == c["ne"]
    for e in GOLD[name]["entries"]:                                         # 40-digit mpmath fixtures
        cfg = np.array(e["cfg"]).reshape(-1, 3)
        assert rel_err(wf.value(cfg), float(e["psi"])) < 2e-10
        gref = np.array([float(t) for t in e["grad"]]).reshape(-1, 3)
        assert np.max(np.abs(wf.gradient(cfg) - gref)) < 2e-10 * max(1.0, np.max(np.abs(gref)))
        assert abs(wf.laplacian(cfg) - float(e["lap"])) < 2e-10 * max(1.0, abs(float(e["lap"])))
        pref = np.array([float(t) for t in e["pgrad"]])
        assert np.max(np.abs(wf.parameter_gradient(cfg) - pref)) < 2e-10 * max(1.0, np.max(np.abs(pref)))
        assert abs(op.act_on(wf, cfg) / wf.value(cfg) - float(e["eloc"])) < 2e-10 * max(1.0, abs(float(e["eloc"])))
    W = 777
    cfgs = random_cfgs(W, c["ne"], seed=5, scale=1.5)
    ens = mole.Ensemble(W, c["ne"], SEED0)
    ens.set_configs(cfgs)
    got = ens.eval_vgl(wf, op)
    ref = orc.eval_batch(c["owf"], c["oham"], cfgs)
    assert rel_err(got["psi"], ref["psi"]) < TOL
    assert close(got["grad"], ref["grad"]) and close(got["lap"], ref["lap"])
    assert close(got["hpsi"] / got["psi"], ref["hpsi"] / ref["psi"]) and close(got["pgrad"], ref["pgrad"])


@pytest.mark.parametrize("name", LSJ)
@pytest.mark.parametrize("metrop", ["box", "diffuse"])
def test_lsj_sweep_parity_and_gram_moments(mole, orc, name, metrop):
    """shared Philox stream: accept/reject bit-exact, traces within 1e-9; the Gram matrix of the rows (both
    implementations) against the same contraction of the oracle's traces; the SR step on it against the oracle's."""
    c = cases()[name]
    wf, op = c["make"](mole)
    P, ne = c["np"], c["ne"]
    W, steps, bs = 96, 30, 10
    seed = bytes([7] * 32)
    param = 0.6 if metrop == "box" else 0.05
    kind = orc.METROP_BOX if metrop == "box" else orc.METROP_DIFFUSE
    m = mole.MetropolisBox(param, seed) if metrop == "box" else mole.MetropolisDiffuse(param, seed)
    cfgs = np.array([orc.init_uniform(seed, w, ne, -1.5, 1.5) for w in range(W)])
    obs = orc.OBS_ENERGY | orc.OBS_WFVALUE | orc.OBS_PGRAD
    ref = orc.ensemble_run(c["owf"], c["oham"], orc.run_options(kind, param, obs, nan_reject=1), cfgs, seed, steps, bs)
    o = ref["pgrad"][:, :, :P] / ref["wfvalue"][:, :, None]
    rows = np.concatenate([np.ones_like(ref["energy"])[:, :, None], ref["energy"][:, :, None], o], axis=2).reshape(-1, P + 2)
    gram_ref = rows.T @ rows
    for impl in (0, 1):
        ens = mole.Ensemble(W, ne, seed)
        ens.init_uniform(-1.5, 1.5)
        ens.gram_select(impl)
        got = ens.sweep(wf, m, op, n_sweeps=steps, n_discard=bs, block_size=bs, observables=obs,
                        traces=("energy", "wfvalue", "pgrad", "accept"))
        assert np.array_equal(got["accept"], ref["accept"]), "accept/reject decisions differ"
        assert close(ens.get_configs(), ref["cfgs"])
        assert close(got["energy"], ref["energy"], 1e-9) and close(got["wfvalue"], ref["wfvalue"])
        assert close(got["pgrad"], ref["pgrad"][:, :, :P], 1e-9)
        acc = ens.acc_get()
        assert acc.n_samples == ref["energy"].size and acc.n_accept == ref["accept"].sum()
        assert abs(acc.sum_e - ref["energy"].sum()) < 1e-9 * np.abs(ref["energy"]).sum()
        g = ens.gram_get()
        assert g.shape == (P + 2, P + 2) and g[0, 0] == ref["energy"].size
        assert np.max(np.abs(g - gram_ref)) < 1e-9 * np.max(np.abs(gram_ref))
        assert np.array_equal(g, g.T)
        e, grad = mole.gram_finalize(g)
        flat = ref["energy"].reshape(-1)
        gref = orc.energy_gradient(ref["wfvalue"].reshape(-1), ref["pgrad"].reshape(-1, ref["pgrad"].shape[2])[:, :P], flat, orc.mean_fold(flat))
        assert abs(e - flat.mean()) < 1e-10 * abs(e) and np.max(np.abs(grad - gref)) < 1e-8 * max(1.0, np.max(np.abs(gref)))
        # a second sweep accumulates; acc_reset clears
        ens.sweep(wf, m, op, n_sweeps=bs, n_discard=0, block_size=bs, observables=obs)
        assert ens.gram_get()[0, 0] == ref["energy"].size + W * bs
        ens.acc_reset()
        assert np.all(ens.gram_get() == 0.0)
    # SR step on the Gram matrix against the oracle's optimizer on the raw samples
    opt = mole.StochasticReconfiguration(0.1, P).set_regularization(1.01, 1e-3)
    dp = opt.compute_parameter_update(wf.parameters(), g)
    S = opt.sr_matrix(g)
    cov = np.cov(o.reshape(-1, P).T, bias=True)
    Sref = cov.copy()
    Sref[np.diag_indices(P)] = np.diag(cov) * 1.01 + 1e-3
    assert np.max(np.abs(S - Sref)) < 1e-8 * np.max(np.abs(Sref))
    dref = 0.1 * np.linalg.solve(Sref, -0.5 * gref)
    assert np.max(np.abs(dp - dref)) < 1e-6 * max(1.0, np.max(np.abs(dref)))


def test_lsj_h8_vmc_sr_lowers_the_energy_and_dmc_runs(mole):
    """P = 36: a short SR optimisation of the H8 chain through mole_vmc_run_optimization (Gram path) lowers the energy
    without incident; one DMC block on the optimised function runs and gives a finite energy below the VMC one."""
    c = cases()["lsj_h8"]
    wf, op = c["make"](mole)
    seed = bytes([3] * 32)
    W = 2048
    ens = mole.Ensemble(W, 8, seed)
    ens.init_uniform(-4.0, 4.0)
    met = mole.MetropolisDiffuse.from_rng(0.05, seed)
    ens.sweep(wf, mole.MetropolisBox.from_rng(1.0, seed), op, n_sweeps=100, observables=0)
    sampler = mole.Sampler.with_initial_configuration(wf, met, mole.operators(**{"Energy": op, "Parameter gradient": mole.ParameterGradient,
                                                                                   "Wavefunction value": mole.WavefunctionValue}),
                                                      ens.get_configs(), n_walkers=W, independent=True)
    opt = mole.StochasticReconfiguration(0.05, 36).set_regularization(1.01, 1e-2)
    runner = mole.VmcRunner(sampler, opt)
    _, en, er = runner.run_optimization(8, 60 * W, 10, W, restart_each_iter=False)
    assert np.isfinite(en).all() and en[-1] < en[0] - 3 * er[0]
    assert runner.ensemble.health() == (0, 0)
    d = mole.Ensemble(W, 8, seed)
    d.set_configs(ens.get_configs())
    se = d.dmc_block(wf, met, op, mole.ffi.BRANCH_SR, 0.01, float(en[-1]), 50)
    assert np.isfinite(se).all()


def test_gram_dmma_matches_the_vector_pipe_at_size(mole):
    """both contractions on the same rows at a size where the SYRK is a real GEMM (P = 36, 2^14 walkers x 40 samples)"""
    c = cases()["lsj_h8"]
    wf, op = c["make"](mole)
    seed = bytes([4] * 32)
    W = 1 << 14
    met = mole.MetropolisDiffuse.from_rng(0.05, seed)
    obs = mole.ffi.OBS_ENERGY | mole.ffi.OBS_PGRAD | mole.ffi.OBS_WFVALUE
    out = []
    for impl in (0, 1):
        ens = mole.Ensemble(W, 8, seed)
        ens.init_uniform(-4.0, 4.0)
        ens.gram_select(impl)
        ens.sweep(wf, met, op, n_sweeps=50, n_discard=10, block_size=10, observables=obs)
        out.append(ens.gram_get())
    assert out[0][0, 0] == W * 40
    assert np.max(np.abs(out[0] - out[1])) < 1e-11 * np.max(np.abs(out[0]))


@pytest.mark.parametrize("W,ns,cols", [(333, 7, 38), (334, 5, 46), (16, 3, 3), (4099, 2, 14), (1 << 15, 20, 38)])
def test_gram_kernels_agree_on_synthetic_rows(mole, W, ns, cols):
    """odd and even walker counts (8-byte / 16-byte load paths of the DMMA kernel), ragged tails, 3 .. 46 columns: the
    tensor-core contraction and the vector-pipe one give the same Gram matrix (checksum of the upper triangle)."""
    ctx = mole.default_context()
    a = ctx.bench_gram(W, ns, cols, 0, 1)
    b = ctx.bench_gram(W, ns, cols, 1, 1)
    assert abs(a[2] - b[2]) <= 1e-11 * max(abs(b[2]), 1.0), (a, b)
