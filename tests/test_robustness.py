"""Robustness of the hot path (missing#4; ADVICE high/medium):
  * a non-finite E_L / O_k sample is counted and kept out of every accumulator instead of poisoning them
    (the reference silently rejects a non-finite psi in the move, metrop.rs:81,197, and has no guard in the sums);
  * a DMC walker with a non-finite local energy dies (weight 0) and is counted;
  * the bench workload's SR loop (Ne Slater-Jastrow, P = 7) stays bounded for 25+ iterations at the bench's step
    and regularisation, and its first-iteration energy agrees with the oracle's run of a sub-sample."""
import importlib.util
import os

import numpy as np
import pytest

from common import SEED0, cases

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("name", ["he", "sj_be"])
def test_nonfinite_samples_are_counted_and_skipped(mole, name):
    c = cases()[name]
    wf, op = c["make"](mole)
    ne, W, nbad, steps, bs = c["ne"], 64, 4, 20, 10
    rng = np.random.default_rng(2)
    cfgs = rng.normal(0.0, 0.8, size=(W, ne, 3))
    cfgs[:nbad, 0, :] = 0.0                                   # electron 0 ON the nucleus: 1/r = inf
    ens = mole.Ensemble(W, ne, SEED0)
    ens.set_configs(cfgs)
    met = mole.MetropolisBox(1e-200, SEED0)                   # nobody moves: |x'|^2 underflows, r stays 0
    obs = mole.ffi.OBS_ENERGY | mole.ffi.OBS_PGRAD | mole.ffi.OBS_WFVALUE
    got = ens.sweep(wf, met, op, n_sweeps=steps, n_discard=0, block_size=bs, observables=obs, traces=("energy",))
    en = got["energy"]
    assert not np.isfinite(en[:nbad]).any() and np.isfinite(en[nbad:]).all()   # the traces show what happened
    assert ens.health() == (nbad * steps, 0)
    acc = ens.acc_get()
    assert acc.n_samples == (W - nbad) * steps and acc.n_blocks == (W - nbad) * steps // bs
    good = en[nbad:]
    assert np.isfinite(acc.sum_e) and abs(acc.sum_e - good.sum()) < 1e-9 * np.abs(good).sum()
    for k in range(acc.n_params):
        assert np.isfinite(acc.sum_o[k]) and np.isfinite(acc.sum_oe[k])
    e, err, _, g = mole.acc_finalize(acc)
    assert np.isfinite(e) and np.isfinite(g).all()
    opt = mole.StochasticReconfiguration(0.01, acc.n_params).set_regularization(1.01, 1e-3)
    assert np.isfinite(opt.compute_parameter_update(wf.parameters(), acc)).all()
    ens.acc_reset()
    assert ens.health() == (0, 0)


def test_dmc_walker_with_nonfinite_local_energy_dies(mole):
    c = cases()["gauss_h"]                                     # E_L = 3a - 2a^2 r^2 - 1/r: -inf on the nucleus
    wf, op = c["make"](mole)
    W, nbad = 256, 3
    rng = np.random.default_rng(4)
    cfgs = rng.normal(0.0, 1.0, size=(W, 1, 3))
    cfgs[:nbad] = 0.0
    seed = bytes([1] * 32)
    ens = mole.Ensemble(W, 1, seed)
    ens.set_configs(cfgs)
    met = mole.MetropolisDiffuse.from_rng(0.025, seed)
    swe, sw = ens.dmc_step(wf, met, op, 0.025, -0.45)
    w = ens.get_weights()
    assert ens.health() == (0, nbad)
    assert np.all(w[:nbad] == 0.0) and np.all(w[nbad:] > 0.0) and np.isfinite(w).all()
    assert sw == W - nbad and np.isfinite(swe)
    ens.branch(mole.ffi.BRANCH_SR)                              # dead walkers are never picked
    assert np.all(ens.branch_sources() >= nbad)
    se = ens.dmc_block(wf, met, op, mole.ffi.BRANCH_SR, 0.025, -0.45, 8)
    assert np.isfinite(se).all()


def test_config5_sr_loop_stays_bounded_at_the_bench_settings(mole, orc, capsys):
    """25 + 3 + 2 iterations (the driver runs --warmup 3 --steps 20..25) at the bench's SR step / regularisation on
    2^12 walkers: parameters bounded, cond(S) bounded and reported, energies inside the physical window and never
    rising by more than noise; the first iteration (starting parameters) against the oracle on a sub-sample."""
    B = _bench()
    W, iters, sweeps = 1 << 12, 30, 200
    wf = mole.SlaterJastrow(5, 5, B.ZETA, B.JB, B.KAPPA)
    op = mole.ElectronicHamiltonian.from_ions([[0, 0, 0]], [10])
    met = mole.MetropolisDiffuse.from_rng(B.TAU, B.SEED)
    opt = mole.StochasticReconfiguration(B.SR_STEP, 7).set_regularization(*B.SR_DIAG)
    ens = mole.Ensemble(W, 10, B.SEED)
    ens.init_normal(0.5)
    ens.sweep(wf, mole.MetropolisBox.from_rng(0.5, B.SEED), op, n_sweeps=200, observables=0)
    ens.sweep(wf, met, op, n_sweeps=50, observables=0)
    obs = mole.ffi.OBS_ENERGY | mole.ffi.OBS_PGRAD | mole.ffi.OBS_WFVALUE
    p0 = wf.parameters().copy()
    hist, conds = [], []
    x_start = ens.get_configs().copy()
    first_trace = None
    for it in range(iters):
        ens.reseed(mole.derive_seed(B.SEED, it))
        ens.acc_reset()
        got = ens.sweep(wf, met, op, n_sweeps=sweeps, n_discard=B.BLOCK, block_size=B.BLOCK, observables=obs,
                        traces=("energy",) if it == 0 else ())
        if it == 0:
            first_trace = got["energy"]
        acc = ens.acc_get()
        e, err, accp, g = mole.acc_finalize(acc)
        conds.append(np.linalg.cond(opt.sr_matrix(acc)))
        wf.update_parameters(opt.compute_parameter_update(wf.parameters(), acc))
        hist.append((e, err, wf.parameters().copy()))
        assert ens.health() == (0, 0)
    es = np.array([h[0] for h in hist])
    errs = np.array([h[1] for h in hist])
    ps = np.array([h[2] for h in hist])
    with capsys.disabled():
        print("\n  config 5 SR loop, %d walkers: E %.4f -> %.4f (+/- %.4f), cond(S) %.2e .. %.2e, p_end %s" % (
            W, es[0], es[-1], errs[-1], min(conds), max(conds), np.array2string(ps[-1], precision=4)))
    assert np.isfinite(es).all() and B.E_WINDOW[0] < es.min() and es.max() < B.E_WINDOW[1]
    assert max(conds) < 1e6                                        # diag x 1.01 alone: 1e4 -> 2e10
    assert np.all(np.abs(ps[:, :3] - p0[:3]) < 0.5 * p0[:3]) and np.all(np.abs(ps[:, 3:] - p0[3:]) < 3.0)
    step_sizes = np.linalg.norm(np.diff(ps, axis=0), axis=1)
    assert step_sizes.max() < 10 * max(step_sizes[0], 1e-3)         # no run-away: |dp| does not blow up
    assert es[-5:].mean() < es[:3].mean() + 5 * errs.max()          # the optimisation does not climb
    # first iteration against the oracle on a sub-sample of the same walkers (same seeds, same start)
    sub = 48
    owf = orc.wf_desc(orc.WF_SLATER_JASTROW, list(B.ZETA) + list(B.JB), [B.KAPPA, 5, 5])
    oham = orc.ham_desc(orc.HAM_ELECTRONIC, [[0, 0, 0]], [10])
    ref = orc.ensemble_run(owf, oham, orc.run_options(orc.METROP_DIFFUSE, B.TAU, orc.OBS_ENERGY, nan_reject=1), x_start[:sub],
                           mole.derive_seed(B.SEED, 0), sweeps, B.BLOCK)
    assert np.max(np.abs(first_trace[:sub] - ref["energy"]) / np.maximum(np.abs(ref["energy"]), 1.0)) < 1e-6
    sub_mean = ref["energy"].mean()
    sub_err = ref["energy"].mean(axis=1).std(ddof=1) / np.sqrt(sub)
    assert abs(sub_mean - es[0]) < 5 * np.hypot(sub_err, errs[0])


def test_two_contexts_driven_from_two_threads(mole, orc):
    """include/mole_b200.h: "different ctxs may be driven from different threads".  Two contexts on the same GPU, each
    with its own stream, ensemble and Slater-Jastrow / SimpleBranching work (the kernels whose shared-memory opt-ins used
    to sit behind unsynchronised process-wide flags), run concurrently and reproduce the serial results bit for bit."""
    import threading
    c = cases()["sj_be"]
    obs = mole.ffi.OBS_ENERGY | mole.ffi.OBS_PGRAD | mole.ffi.OBS_WFVALUE

    def work(ctx, out, key):
        wf = mole.SlaterJastrow(2, 2, (3.68, 0.96, 0.96), (0.5, 1.0, 0.2, 0.1), 1.0, ctx=ctx)
        op = mole.ElectronicHamiltonian.from_ions([[0, 0, 0]], [4], ctx=ctx)
        met = mole.MetropolisDiffuse.from_rng(0.05, SEED0)
        ens = mole.Ensemble(3000, 4, SEED0, ctx=ctx)
        ens.init_uniform(-1.0, 1.0)
        for _ in range(5):
            ens.sweep(wf, met, op, n_sweeps=20, n_discard=0, block_size=10, observables=obs)
        acc = ens.acc_get()
        g = mole.STO(0.9, ctx=ctx)
        gop = mole.ElectronicHamiltonian.from_ions([[0, 0, 0]], [1], ctx=ctx)
        d = mole.Ensemble(70000, 1, SEED0, ctx=ctx)          # > 65k walkers: SimpleBranching needs the > 48 KB opt-in
        d.init_normal(1.0)
        d.dmc_step(g, mole.MetropolisDiffuse.from_rng(0.025, SEED0), gop, 0.025, -0.5)
        d.branch(mole.ffi.BRANCH_SIMPLE)
        out[key] = (ens.get_configs(), acc.sum_e, acc.oo(2, 5), d.get_configs())

    serial, par = {}, {}
    work(mole.Context(0), serial, "a")
    ctxs = [mole.Context(0), mole.Context(0)]
    ths = [threading.Thread(target=work, args=(ctxs[i], par, i)) for i in range(2)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    assert set(par) == {0, 1}
    for i in range(2):
        for a, b in zip(par[i], serial["a"]):
            assert np.array_equal(np.asarray(a), np.asarray(b))


def test_rebalance_on_one_rank_is_the_identity_after_equalising_weights(mole):
    """mole_rebalance with a single rank: unequal weights are first resampled (SRBrancher, stratified), then every
    walker keeps its slot and carries the mean weight; total weight conserved."""
    c = cases()["sto_h"]
    wf, op = c["make"](mole)
    W = 512
    ens = mole.Ensemble(W, 1, SEED0)
    ens.init_normal(1.0)
    x0 = ens.get_configs().copy()
    ens.rebalance()                                              # uniform weights: nothing changes
    assert np.array_equal(ens.get_configs(), x0) and np.all(ens.get_weights() == 1.0)
    w = np.linspace(0.5, 1.5, W)
    ens.set_weights(w)
    ens.rebalance()
    got = ens.get_weights()
    assert np.all(got == got[0]) and abs(got[0] - w.mean()) < 1e-12
    src = ens.branch_sources()
    assert np.array_equal(ens.get_configs()[:, 0, :], x0[src, 0, :])
    # one rank = one island: the imbalance the DMC drivers test between blocks is 1, before and after a block
    assert ens.island_imbalance() == 1.0
    met = mole.MetropolisDiffuse.from_rng(0.025, SEED0)
    ens.dmc_block(wf, met, op, mole.ffi.BRANCH_SR, 0.025, -0.5, 4)
    assert ens.island_imbalance() == 1.0
    with pytest.raises(mole.MoleError):
        ens.dmc_block_select(3)
