"""Parity tests proper: the CUDA path (through the C ABI) against the CPU oracle on identical
walker configurations and a shared Philox stream.  Bars (BASELINE.json north_star):
psi, grad psi, lap psi, E_L, parameter gradients within 1e-10 relative; accept/reject bit-exact."""
import json
import os

import numpy as np
import pytest

from common import CFG2, SEED0, cases, random_cfgs, rel_err

pytestmark = pytest.mark.gpu

TOL = 1e-10
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "wf_golden.json")))
LCAO = ["lcao_h2p", "lcao_he", "lcao_h2_singlet", "lcao_h2_triplet"]
SMALL = ["h2", "he", "h2p", "gauss_sho", "gauss_h", "sto_h"] + LCAO
SJ = ["sj_ne", "sj_be", "sj_li"]
ALL = SMALL + SJ


def close(a, b, tol=TOL):
    a, b = np.asarray(a), np.asarray(b)
    scale = np.maximum(np.abs(b), 1e-3 * np.max(np.abs(b)) if b.size else 1.0)
    return np.max(np.abs(a - b) / scale) < tol if a.size else True


@pytest.mark.parametrize("name", ALL)
def test_eval_vgl_matches_oracle(mole, orc, name):
    c = cases()[name]
    wf, op = c["make"](mole)
    W = 4096 if name in SMALL else 1531        # not a multiple of the 6 walkers a warp holds
    cfgs = random_cfgs(W, c["ne"], seed=5, scale=1.0 if name in SMALL else 0.7)
    ens = mole.Ensemble(W, c["ne"], SEED0)
    ens.set_configs(cfgs)
    assert np.array_equal(ens.get_configs(), cfgs)
    got = ens.eval_vgl(wf, op)
    ref = orc.eval_batch(c["owf"], c["oham"], cfgs)
    assert rel_err(got["psi"], ref["psi"]) < TOL
    assert close(got["grad"], ref["grad"])
    assert close(got["lap"], ref["lap"])
    assert close(got["hpsi"], ref["hpsi"])
    assert close(got["hpsi"] / got["psi"], ref["hpsi"] / ref["psi"])   # E_L
    if c["np"]:
        assert close(got["pgrad"], ref["pgrad"])


@pytest.mark.parametrize("name", ALL)
def test_pointwise_traits_match_mpmath_golden(mole, name):
    g = GOLD[name]
    wf, op = cases()[name]["make"](mole)
    for e in g["entries"]:
        cfg = np.array(e["cfg"]).reshape(-1, 3)
        assert rel_err(wf.value(cfg), float(e["psi"])) < TOL
        gref = np.array([float(t) for t in e["grad"]]).reshape(-1, 3)
        assert np.max(np.abs(wf.gradient(cfg) - gref)) < TOL * max(1.0, np.max(np.abs(gref)))
        assert abs(wf.laplacian(cfg) - float(e["lap"])) < TOL * max(1.0, abs(float(e["lap"])))
        if g["n_params"]:
            assert rel_err(wf.parameter_gradient(cfg), [float(t) for t in e["pgrad"]]) < TOL
        assert abs(op.act_on(wf, cfg) / wf.value(cfg) - float(e["eloc"])) < TOL * max(1.0, abs(float(e["eloc"])))


def test_lcao_helium_is_the_reference_helium(mole):
    """tests/helium_lcao.rs:94-102: the commented-out SpinDeterminantProduct over a Hydrogen1sBasis of width 1/1.69 and
    the HeliumAtomWaveFunction(1.69) the test runs in its place are the same function -> with one Philox stream the
    two kinds take the same decisions and give the same energies (and the reference's criterion, :134)."""
    c = cases()
    he, op = c["he"]["make"](mole)
    lc, _ = c["lcao_he"]["make"](mole)
    out = []
    for wf in (he, lc):
        ens = mole.Ensemble(4096, 2, SEED0)
        ens.init_uniform(-1.0, 1.0)
        m = mole.MetropolisDiffuse(0.1, SEED0)                              # helium_lcao.rs:110
        got = ens.sweep(wf, m, op, n_sweeps=300, n_discard=50, block_size=50, observables=mole.ffi.OBS_ENERGY,
                        traces=("energy", "accept"))
        e, err, _, _ = mole.acc_finalize(ens.acc_get())
        out.append((got["accept"], got["energy"], ens.get_configs(), e, err))
    assert np.array_equal(out[0][0], out[1][0])
    assert close(out[1][1], out[0][1], 1e-9) and close(out[1][2], out[0][2])
    assert abs(out[1][3] - (-2.84765625)) < 5 * out[1][4] + 2e-3          # <E> of exp(-alpha(r1+r2)) at alpha = 1.69 is alpha^2 - 27 alpha / 8


def test_lcao_descriptor_errors(mole):
    b3 = [[0, 0, 0], [1, 0, 0], [2, 0, 0]]
    with pytest.raises(mole.MoleError):
        mole.Hydrogen1sBasis(b3, [1.0])                                     # closed set: one or two centres
    with pytest.raises(mole.MoleError):
        mole.Hydrogen1sBasis([[0, 0, 0]], [1.0, 2.0])                       # one width
    b = mole.Hydrogen1sBasis([[0, 0, 0]], [1.0])
    with pytest.raises(mole.MoleError):
        mole.SingleDeterminant([mole.Orbital([[1.0]], b)])                  # (1 electron, 1 centre) is the STO kind
    with pytest.raises(mole.MoleError):
        mole.SpinDeterminantProduct([mole.Orbital([[1.0]], b)] * 2, 2)
    bneg = mole.Hydrogen1sBasis([[0, 0, 0]], [-1.0])
    with pytest.raises(mole.MoleError):
        mole.SpinDeterminantProduct([mole.Orbital([[1.0]], bneg)] * 2, 1)   # width <= 0 is rejected by the library


def test_init_draws_match_oracle(mole, orc):
    W, ne = 300, 2
    seed = bytes(range(32))
    ens = mole.Ensemble(W, ne, seed, walker_offset=1000)
    ens.init_uniform(-1.0, 1.0)
    got = ens.get_configs()
    ref = np.array([orc.init_uniform(seed, 1000 + w, ne) for w in range(W)])
    assert np.array_equal(got, ref)          # integer -> double conversion and one mul/add: bit-exact
    ens.init_normal(1.0)
    got = ens.get_configs()
    ref = np.array([orc.init_normal(seed, 1000 + w, ne) for w in range(W)])
    assert np.max(np.abs(got - ref)) < 1e-14  # log/sincos differ in the last ulp between libm and CUDA
    ens.init_uniform(-1.0, 1.0, broadcast_walker0=True)   # vmc.rs:56: clones of ONE sampler
    got = ens.get_configs()
    assert np.array_equal(got, np.broadcast_to(orc.init_uniform(seed, 0, ne), got.shape))


@pytest.mark.parametrize("name", ALL)
@pytest.mark.parametrize("metrop", ["box", "diffuse", "diffuse_nan_accept"])
def test_sweep_parity_shared_philox(mole, orc, name, metrop):
    """Identical starting walkers + shared Philox stream: accept/reject decisions bit-exact,
    E_L traces, stored observables and final configurations within 1e-10.  "diffuse_nan_accept" runs
    the reference-faithful NaN policy (MOLE_COMPAT_NAN_ACCEPT); it matters for the nodal Slater-Jastrow
    kinds, whose far-from-equilibrium start puts walkers next to nodes (t_high = t_low = 0)."""
    c = cases()[name]
    wf, op = c["make"](mole)
    W, steps, bs = (256, 60, 10) if name in SMALL else (100, 30, 10)
    seed = bytes([7] * 32)
    nan_accept = metrop.endswith("nan_accept")
    metrop = metrop.split("_")[0]
    param = (1.0 if metrop == "box" else 0.25) if name in SMALL else (0.4 if metrop == "box" else 0.02)
    m = mole.MetropolisBox(param, seed) if metrop == "box" else mole.MetropolisDiffuse(param, seed)
    if nan_accept:
        m.set_compat(mole.ffi.COMPAT_NAN_ACCEPT)
    cfgs = np.array([orc.init_uniform(seed, w, c["ne"]) for w in range(W)])
    obs = orc.OBS_ENERGY | orc.OBS_WFVALUE | orc.OBS_KINETIC | (orc.OBS_PGRAD if c["np"] else 0)
    ref = orc.ensemble_run(c["owf"], c["oham"], orc.run_options(orc.METROP_BOX if metrop == "box" else orc.METROP_DIFFUSE,
                                                                 param, obs, nan_reject=0 if nan_accept else 1),
                           cfgs, seed, steps, bs)
    ens = mole.Ensemble(W, c["ne"], seed)
    ens.init_uniform(-1.0, 1.0)
    got = ens.sweep(wf, m, op, n_sweeps=steps, n_discard=bs, block_size=bs, observables=obs,
                    traces=("energy", "wfvalue", "kinetic", "pgrad", "accept"))
    assert np.array_equal(got["accept"], ref["accept"]), "accept/reject decisions differ"
    assert 0.05 < got["accept"].mean() <= 1.0
    assert close(ens.get_configs(), ref["cfgs"])
    assert close(got["energy"], ref["energy"], 1e-9)
    assert close(got["wfvalue"], ref["wfvalue"])
    assert close(got["kinetic"], ref["kinetic"], 1e-9)
    if c["np"]:
        assert close(got["pgrad"], ref["pgrad"][:, :, :c["np"]])
    # accumulators against the raw series
    acc = ens.acc_get()
    en = got["energy"]
    assert acc.n_samples == en.size and acc.n_moves == W * steps * c["ne"]
    assert acc.n_accept == got["accept"].sum()
    assert abs(acc.sum_e - en.sum()) < 1e-9 * abs(en.sum())
    assert abs(acc.sum_e2 - (en ** 2).sum()) < 1e-9 * (en ** 2).sum()
    bm = en.reshape(W, -1, bs).mean(axis=2)
    assert acc.n_blocks == bm.size and abs(acc.sum_b2 - (bm ** 2).sum()) < 1e-9 * (bm ** 2).sum()
    e, err, accp, g = mole.acc_finalize(acc)
    flat = ref["energy"].reshape(-1)
    assert abs(e - orc.mean_fold(flat)) < 1e-9 * abs(e)
    assert abs(err - orc.blocking_error(flat, bs, orc.mean_fold(flat))) < 1e-7 * err
    if c["np"]:
        P = c["np"]
        o = ref["pgrad"][:, :, :P] / ref["wfvalue"][:, :, None]
        assert acc.n_params == P
        for k in range(P):
            sk = np.abs(o[:, :, k]).sum() + 1e-300
            assert abs(acc.sum_o[k] - o[:, :, k].sum()) < 1e-9 * sk
            assert abs(acc.sum_oe[k] - (o[:, :, k] * ref["energy"]).sum()) < 1e-9 * np.abs(o[:, :, k] * ref["energy"]).sum() + 1e-300
            for l in range(k, P):
                assert abs(acc.oo(k, l) - (o[:, :, k] * o[:, :, l]).sum()) < 1e-9 * np.abs(o[:, :, k] * o[:, :, l]).sum() + 1e-300
        gref =orc.energy_gradient(ref["wfvalue"].reshape(-1), ref["pgrad"].reshape(-1, max(c["np"], 1))[:, :c["np"]],
                                   flat, orc.mean_fold(flat))
        assert np.max(np.abs(g - gref)) < 1e-8 * max(1.0, np.max(np.abs(gref)))


OPS = ["kinetic", "ionic_pot", "elec_pot", "ionic"]


def _op_pair(mole, orc, c, opname):
    """(oracle ham desc, mole operator) of one operator kind over the case's ions (operator.rs:16-149)."""
    pos = np.array(c["oham"].ion_pos)[:3 * c["oham"].n_ions].reshape(-1, 3)
    z = list(c["oham"].ion_charge)[:c["oham"].n_ions]
    if opname == "kinetic":
        return orc.ham_desc(orc.HAM_KINETIC), mole.KineticEnergy()
    if opname == "ionic_pot":
        return orc.ham_desc(orc.HAM_IONIC_POT, pos, z), mole.IonicPotential(pos, z)
    if opname == "elec_pot":
        return orc.ham_desc(orc.HAM_ELEC_POT), mole.ElectronicPotential()
    return orc.ham_desc(orc.HAM_IONIC, pos, z), mole.IonicHamiltonian(mole.KineticEnergy(), mole.IonicPotential(pos, z))


@pytest.mark.parametrize("name", ["h2", "he", "lcao_h2_singlet", "sj_ne", "sj_li"])
@pytest.mark.parametrize("opname", OPS)
def test_operator_kinds_act_on_and_as_sweep_operator(mole, orc, name, opname):
    """KineticEnergy, IonicPotential, ElectronicPotential and IonicHamiltonian
    (operator.rs:59-61,94-96,122-124,146-148) as the operator of mole_op_act_on, of the batched evaluation and
    of a fused sweep ("Energy" => op), for a thread-per-walker kind and the cooperative Slater-Jastrow kind."""
    c = cases()[name]
    wf, _ = c["make"](mole)
    oham, op = _op_pair(mole, orc, c, opname)
    ne = c["ne"]
    # LocalOperator::act_on, one configuration at a time (H psi, not divided)
    for cfg in random_cfgs(4, ne, seed=11, scale=0.8):
        ref = orc.ham_act_on(oham, c["owf"], cfg)
        assert abs(op.act_on(wf, cfg) - ref) < TOL * max(abs(ref), 1e-3 * abs(orc.wf_value(c["owf"], cfg)))
    # batched
    W = 515
    cfgs = random_cfgs(W, ne, seed=12, scale=0.8)
    ens = mole.Ensemble(W, ne, SEED0)
    ens.set_configs(cfgs)
    got = ens.eval_vgl(wf, op, want=("psi", "hpsi"))
    ref = orc.eval_batch(c["owf"], oham, cfgs, want_pgrad=False)
    assert close(got["hpsi"] / got["psi"], ref["hpsi"] / ref["psi"])
    # as the "Energy" operator of a sweep: same decisions (the operator does not enter the moves), its own E_L trace
    steps, bs = 30, 10
    seed = bytes([9] * 32)
    tau = 0.25 if name in SMALL else 0.02
    m = mole.MetropolisDiffuse(tau, seed)
    start = np.array([orc.init_uniform(seed, w, ne) for w in range(64)])
    refr = orc.ensemble_run(c["owf"], oham, orc.run_options(orc.METROP_DIFFUSE, tau, orc.OBS_ENERGY, nan_reject=1), start, seed, steps, bs)
    e2 = mole.Ensemble(64, ne, seed)
    e2.init_uniform(-1.0, 1.0)
    out = e2.sweep(wf, m, op, n_sweeps=steps, n_discard=bs, block_size=bs, observables=mole.ffi.OBS_ENERGY, traces=("energy", "accept"))
    assert np.array_equal(out["accept"], refr["accept"])
    assert close(out["energy"], refr["energy"], 1e-9)
    acc = e2.acc_get()
    assert abs(acc.sum_e - refr["energy"].sum()) < 1e-9 * np.abs(refr["energy"]).sum()


def test_sweep_is_independent_of_launch_split_and_sharding(mole, orc):
    """Philox keys are (walker, step): splitting the sweeps over launches, or the walkers over
    ensembles (ranks), must not change a single bit of the trajectory."""
    c = cases()["h2"]
    wf, op = c["make"](mole)
    m = mole.MetropolisDiffuse(0.25, SEED0)
    W, steps = 512, 40
    a = mole.Ensemble(W, 2, SEED0); a.init_uniform()
    a.sweep(wf, m, op, n_sweeps=steps, block_size=10)
    b = mole.Ensemble(W, 2, SEED0); b.init_uniform()
    for _ in range(4):
        b.sweep(wf, m, op, n_sweeps=steps // 4, block_size=10)
    assert np.array_equal(a.get_configs(), b.get_configs())
    accA, accB = a.acc_get(), b.acc_get()
    assert accA.n_samples == accB.n_samples and accA.n_blocks == accB.n_blocks == W * steps // 10
    assert abs(accA.sum_e - accB.sum_e) < 1e-9 * abs(accA.sum_e) and abs(accA.sum_b2 - accB.sum_b2) < 1e-9 * accA.sum_b2
    h0 = mole.Ensemble(W // 2, 2, SEED0, walker_offset=0); h0.init_uniform()
    h1 = mole.Ensemble(W // 2, 2, SEED0, walker_offset=W // 2); h1.init_uniform()
    for h in (h0, h1):
        h.sweep(wf, m, op, n_sweeps=steps, block_size=10)
    assert np.array_equal(np.concatenate([h0.get_configs(), h1.get_configs()]), a.get_configs())
    s0, s1 = h0.acc_get(), h1.acc_get()
    assert abs((s0.sum_e + s1.sum_e) - accA.sum_e) < 1e-9 * abs(accA.sum_e)


def test_compat_vector_div_quirk(mole, orc):
    """MOLE_COMPAT_VECTOR_DIV reproduces operator/src/traits.rs:149-150 (Vector/Scalar = scalar/array)."""
    c = cases()["he"]
    wf, op = c["make"](mole)
    W, steps, bs = 64, 30, 10
    m = mole.MetropolisDiffuse(0.25, SEED0)
    obs = orc.OBS_ENERGY | orc.OBS_WFVALUE | orc.OBS_PGRAD
    cfgs = np.array([orc.init_uniform(SEED0, w, 2) for w in range(W)])
    ref = orc.ensemble_run(c["owf"], c["oham"], orc.run_options(orc.METROP_DIFFUSE, 0.25, obs, quirk_vector_div=1), cfgs,
                           SEED0, steps, bs)
    ens = mole.Ensemble(W, 2, SEED0); ens.init_uniform()
    got = ens.sweep(wf, m, op, n_sweeps=steps, n_discard=bs, block_size=bs, observables=obs,
                    compat=mole.ffi.COMPAT_VECTOR_DIV, traces=("pgrad",))
    assert close(got["pgrad"], ref["pgrad"])
    acc = ens.acc_get()
    o = ref["pgrad"][:, :, 0] / ref["wfvalue"]
    assert abs(acc.sum_o[0] - o.sum()) < 1e-9 * abs(o.sum())


def test_runner_known_answers(mole):
    """The reference's integration tests re-expressed with proper error bars."""
    # tests/helium_lcao.rs:89-138 (He, alpha=1.69, Diffuse(0.1), run(1000,100)), many chains
    wf = mole.HeliumAtomWaveFunction(1.69)
    h = mole.ElectronicHamiltonian(mole.KineticEnergy(), mole.IonicPotential([[0, 0, 0]], [2]), mole.ElectronicPotential())
    s = mole.Sampler(wf, mole.MetropolisDiffuse.from_rng(0.1, SEED0), mole.operators(Energy=h), n_walkers=4096, independent=True)
    r = mole.Runner(s).run(1000, 100)
    en = r.data["Energy"]
    assert en.shape == (4096, 900)
    exact = 0.5 * 1.5 ** 6 * (-0.5)
    chain_means = en.mean(axis=1)
    err = chain_means.std() / np.sqrt(len(chain_means))
    assert abs(chain_means.mean() - exact) < 5 * err and err < 5e-3
    assert abs(en.mean() - exact) < 2.0 * en.std()      # the reference's own (loose) assertion, :137
    # tests/hydrogen_molecular_ion_lcao.rs:100-142 (H2+, Box(1.0), run(10000,100)) on 256 chains of 2000 steps
    wf = mole.H2WF(2.5, 1.0)
    h = mole.ElectronicHamiltonian(mole.KineticEnergy(), mole.IonicPotential([[-1.25, 0, 0], [1.25, 0, 0]], [1, 1]),
                                   mole.ElectronicPotential())
    s = mole.Sampler(wf, mole.MetropolisBox.from_rng(1.0, SEED0), mole.operators(Energy=h), n_walkers=256, independent=True)
    en = mole.Runner(s).run(2000, 100).data["Energy"]
    assert abs(en.mean() - (-0.565)) < en.std()          # :141
    # examples/custom_operator.rs:100-139: exact eigenstate, zero variance
    wf = mole.GaussianWaveFunction(np.sqrt(2.0))
    s = mole.Sampler(wf, mole.MetropolisBox.from_rng(1.0, SEED0), mole.operators(Energy=mole.HarmonicHamiltonian(1.0)), n_walkers=64,
                     independent=True)
    r = mole.Runner(s).run(1000, 1)
    en = r.data["Energy"]
    assert abs(en.mean() - 1.5) < 1e-14 and en.std() < 1e-14
    assert 0.0 < r.acceptance <= 64 * 1000


def test_runner_assert_and_errors(mole):
    wf = mole.GaussianWaveFunction(1.0)
    s = mole.Sampler(wf, mole.MetropolisBox(1.0, SEED0), mole.operators(Energy=mole.HarmonicHamiltonian(1.0)))
    with pytest.raises(mole.MoleError) as ei:            # montecarlo.rs:29
        mole.Runner(s).run(10, 10)
    assert ei.value.code == mole.ffi.ERR_ASSERT
    mock = mole.WaveFunctionMock(1.0)
    with pytest.raises(mole.MoleError) as ei:            # metrop.rs:248-250 unimplemented!()
        mock.gradient([[0.0, 0.0, 0.0]])
    assert ei.value.code == mole.ffi.ERR_FUNC
    ens = mole.Ensemble(4, 1, SEED0)
    with pytest.raises(mole.MoleError):
        ens.sweep(mock, mole.MetropolisDiffuse(0.1, SEED0), None, n_sweeps=2, observables=0)
    with pytest.raises(mole.MoleError) as ei:            # shape mismatch
        mole.Ensemble(4, 2, SEED0).sweep(wf, mole.MetropolisBox(1.0, SEED0), None, n_sweeps=2, observables=0)
    assert ei.value.code == mole.ffi.ERR_SHAPE


def test_uniform_wf_always_accepted(mole):
    """proptest of metrop.rs:259-269: for psi == 1 every MetropolisBox proposal is accepted."""
    mock = mole.WaveFunctionMock(1.0)
    W = 1024
    ens = mole.Ensemble(W, 1, SEED0)
    rng = np.random.default_rng(0)
    ens.set_configs(rng.normal(size=(W, 1, 3)) * 10.0 ** rng.integers(-30, 30, size=(W, 1, 1)))
    got = ens.sweep(mock, mole.MetropolisBox(1.0, SEED0), None, n_sweeps=8, observables=0, traces=("accept",))
    assert got["accept"].all()


@pytest.mark.parametrize("compat", [0, 3])
def test_vmc_run_optimization_matches_oracle(mole, orc, compat):
    """VmcRunner::run_optimization on the H2 example's shape (hydrogen_molecule.rs:171-199,235-273),
    8 workers, restart-from-master semantics, SR; energies/errors/parameters per iteration."""
    c = cases()["h2"]
    iters, total, bs, nw = 4, 4000, 10, 8
    step = 0.05
    cfg0 = orc.init_uniform(SEED0, 0, 2)
    opts = orc.run_options(orc.METROP_DIFFUSE, 0.25, quirk_vector_div=1 if compat else 0)
    ropt = orc.Optimizer(orc.OPT_SR, 1, step, quirk_sr_subtract=1 if compat else 0)
    ref = orc.vmc_run_optimization(c["owf"], c["oham"], opts, ropt, SEED0, cfg0, iters, total, bs, nw)
    wf, op = c["make"](mole)
    obs = mole.operators(**{"Energy": op, "Parameter gradient": mole.ParameterGradient, "Wavefunction value": mole.WavefunctionValue})
    sampler = mole.Sampler.new(wf, mole.MetropolisDiffuse.from_rng(0.25, SEED0), obs, compat=compat)
    vmc = mole.VmcRunner.__new__(mole.VmcRunner)
    vmc.__init__(sampler, mole.StochasticReconfiguration(step, 1, compat=compat))
    wf_out, en, er = vmc.run_optimization(iters, total, bs, nw)
    assert np.max(np.abs(en - ref["energies"])) < 1e-8
    assert np.max(np.abs(er - ref["errors"]) / ref["errors"]) < 1e-6
    assert np.max(np.abs(vmc.acceptance - ref["acceptance"])) < 1e-12
    assert np.max(np.abs(vmc.param_history - ref["param_history"])) < 1e-7 * max(1.0, np.max(np.abs(ref["param_history"])))
    assert abs(wf_out.parameters()[0] - ref["params"][0]) < 1e-7 * max(1.0, abs(ref["params"][0]))


def test_sho_optimize(mole):
    """tests/sho_optimize.rs:117-141 with enough walkers for a meaningful error bar."""
    wf = mole.GaussianWaveFunction(1.0)
    obs = mole.operators(**{"Energy": mole.HarmonicHamiltonian(1.0), "Parameter gradient": mole.ParameterGradient,
                            "Wavefunction value": mole.WavefunctionValue})
    sampler = mole.Sampler(wf, mole.MetropolisDiffuse.from_rng(0.1, SEED0), obs)
    vmc = mole.VmcRunner(sampler, mole.SteepestDescent(2e-1))
    _, en, er = vmc.run_optimization(40, 4096 * 200, 10, 4096, restart_each_iter=False)
    assert abs(wf.parameters()[0] - np.sqrt(2.0)) < 0.02
    assert abs(en[-1] - 1.5) < max(5 * er[-1], 1e-3)
    assert en[-1] < en[0]


def _dmc_setup(mole, orc, W, identical):
    c = cases()["gauss_h"]
    wf, op = c["make"](mole)
    seed = bytes([1] * 32)        # examples/dmc.rs:198
    m = mole.MetropolisDiffuse.from_rng(0.025, seed).fix_nodes()
    cfgs = np.array([orc.init_normal(seed, 0 if identical else w, 1) for w in range(W)])
    return c, wf, op, m, seed, cfgs


def test_dmc_step_and_sr_branch_match_oracle(mole, orc):
    """One DMC time step (dmc.rs:87-130) and one SRBrancher::branch per iteration, re-synchronised on
    the oracle's state each step: k_i = trunc(w_i N / w_max) sits exactly on an integer for the
    heaviest walker, so a last-ulp difference in exp() flips it; with bit-identical weights the
    integer weights, the draws and therefore the picked walkers must be bit-exact."""
    W = 2048
    c, wf, op, m, seed, cfgs = _dmc_setup(mole, orc, W, identical=False)
    dmc = mole.DmcRunner(wf, W, 0.6, op, m, mole.SRBrancher.new(), identical_start=False)
    assert np.max(np.abs(dmc.ensemble.get_configs() - cfgs)) < 1e-14
    ens = dmc.ensemble
    w, x, e_ref = np.ones(W), cfgs.copy(), 0.6
    for t in range(6):
        ens.set_configs(x); ens.set_weights(w); ens.step = t
        e_o, tw_o, w, x = orc.dmc_step(c["owf"], c["oham"], w, x, 0.025, 0.025, e_ref, seed, t)
        swe, sw = ens.dmc_step(wf, m, op, 0.025, e_ref)
        assert abs(swe / sw - e_o) < 1e-10 * abs(e_o) and abs(sw - tw_o) < 1e-10 * tw_o
        assert close(ens.get_weights(), w) and close(ens.get_configs(), x)
        ens.set_configs(x); ens.set_weights(w)
        w, x = orc.branch(orc.BRANCH_SR, 1, w, x, seed, t)
        ens.branch(mole.ffi.BRANCH_SR)
        assert np.array_equal(ens.get_configs(), x)      # same walkers picked
        assert np.array_equal(ens.get_weights(), w)
        assert ens.step == t + 1
    # without re-synchronisation the cached E_L and the device-side sum/max feed the next step
    ens.set_configs(x); ens.set_weights(w); ens.step = 6
    for t in range(6, 9):
        e_o, tw_o, w, x = orc.dmc_step(c["owf"], c["oham"], w, x, 0.025, 0.025, e_ref, seed, t)
        swe, sw = ens.dmc_step(wf, m, op, 0.025, e_ref)
        assert abs(swe / sw - e_o) < 1e-10 * abs(e_o)
        ens.branch(mole.ffi.BRANCH_SR)
        src = ens.branch_sources()
        w, x = np.full(W, w.mean()), x[src]
        assert close(ens.get_weights(), w) and close(ens.get_configs(), x)


def test_simple_branching_matches_oracle(mole, orc):
    W = 3000
    rng = np.random.default_rng(4)
    cfgs = rng.normal(size=(W, 1, 3))
    for trial, scale in enumerate((1.0, 0.7, 1.4)):     # balanced, shrinking (clone), growing (random removal)
        w = rng.gamma(4.0, 0.25, size=W) * scale
        ens = mole.Ensemble(W, 1, SEED0)
        ens.set_configs(cfgs); ens.set_weights(w); ens.step = 11 + trial
        wo, xo = orc.branch(orc.BRANCH_SIMPLE, 1, w, cfgs, SEED0, 11 + trial)
        ens.branch(mole.ffi.BRANCH_SIMPLE)
        assert np.array_equal(ens.get_weights(), wo)
        assert np.array_equal(ens.get_configs(), xo)
    # SR on hand-set weights
    w = rng.gamma(2.0, 0.5, size=W)
    ens = mole.Ensemble(W, 1, SEED0)
    ens.set_configs(cfgs); ens.set_weights(w); ens.step = 3
    wo, xo = orc.branch(orc.BRANCH_SR, 1, w, cfgs, SEED0, 3)
    ens.branch(mole.ffi.BRANCH_SR)
    assert np.array_equal(ens.get_configs(), xo) and close(ens.get_weights(), wo)
    src = ens.branch_sources()
    assert np.array_equal(cfgs[src], xo)


def _update_energies(step_energies, e_ref, bs, neq):
    """DmcRunner::update_energies (dmc.rs:155-202) restated on the host for the recursion check."""
    en, var = [], []
    for b, blk in enumerate(np.asarray(step_energies).reshape(-1, bs)):
        e = blk.sum() / bs
        if b == neq:
            e_ref = (e_ref + e) / 2; en.append(e); var.append(0.0)
        if b > neq:
            prev, k = en[-1], b - neq
            en.append(prev + (e - prev) / k)
            e_ref = (e_ref + en[-1]) / 2
            var.append(var[-1] + ((e - prev) * (e - en[-1]) - var[-1]) / k)
    return np.array(en), np.sqrt(np.array(var) / np.arange(1, len(var) + 1)), e_ref


@pytest.mark.parametrize("identical", [True, False])
def test_dmc_diffuse_matches_oracle(mole, orc, identical):
    """DmcRunner::diffuse on the examples/dmc.rs shape (tau=0.025, SRBrancher).  The first steps are
    compared one to one; after the first last-ulp flip of an integer weight the two populations are
    different samples of the same process, so the run is compared within statistical error bars, and
    the block / E_ref / variance recursion is checked exactly on the GPU's own step energies."""
    W, iters, bs, neq = 400, 2000, 50, 4
    c, wf, op, m, seed, cfgs = _dmc_setup(mole, orc, W, identical)
    ref = orc.dmc_diffuse(c["owf"], c["oham"], np.ones(W), cfgs, 0.025, 0.62, orc.BRANCH_SR, seed, 0.025, iters, bs, neq)
    dmc = mole.DmcRunner.new(wf, W, 0.62, op, m, mole.SRBrancher.new(), identical_start=identical)
    en, er = dmc.diffuse(0.025, iters, bs, neq, want_steps=True)
    assert len(en) == len(ref["energies"]) == iters // bs - neq
    assert abs(dmc.step_energies[0] - ref["step_energies"][0]) < 1e-10 * abs(ref["step_energies"][0])
    en2, er2, eref2 = _update_energies(dmc.step_energies, 0.62, bs, neq)
    assert np.allclose(en, en2, rtol=0, atol=1e-13) and np.allclose(er, er2, rtol=0, atol=1e-13)
    assert abs(dmc.reference_energy - eref2) < 1e-13
    sigma = np.hypot(er[-1], ref["errors"][-1])
    assert abs(en[-1] - ref["energies"][-1]) < 5 * sigma + 2e-3


def test_dmc_hydrogen_energy(mole):
    """examples/dmc.rs:224: DMC with a nodeless guiding function converges to the exact -0.5 Ha."""
    wf = mole.GaussianWaveFunction(1.3)
    op = mole.ElectronicHamiltonian.from_ions([[0, 0, 0]], [1])
    m = mole.MetropolisDiffuse.from_rng(0.01, bytes([1] * 32)).fix_nodes()
    dmc = mole.DmcRunner.new(wf, 8192, -0.45, op, m, mole.SRBrancher.new(), identical_start=False)
    en, er = dmc.diffuse(0.01, 3000, 100, 10)
    assert abs(en[-1] + 0.5) < max(6 * er[-1], 3e-3)


def test_vmc_sr_slater_jastrow_p7_matches_oracle(mole, orc):
    """SR with P = 7 (untested upstream, SURVEY §4 gaps): Be Slater-Jastrow, 6 workers, 3 iterations;
    energies, blocking errors and all seven parameters per iteration against the oracle's raw-sample path."""
    # Be uses only the 1s/2s orbitals: d psi/d zeta3 == 0, S is singular -> Error::LinalgError on both sides
    cb = cases()["sj_be"]
    with pytest.raises(RuntimeError):
        orc.vmc_run_optimization(cb["owf"], cb["oham"], orc.run_options(orc.METROP_DIFFUSE, 0.05, nan_reject=1),
                                 orc.Optimizer(orc.OPT_SR, 7, 0.02), SEED0, orc.init_uniform(SEED0, 0, 4), 1, 360, 10, 6)
    wfb, opb = cb["make"](mole)
    obsb = mole.operators(**{"Energy": opb, "Parameter gradient": mole.ParameterGradient, "Wavefunction value": mole.WavefunctionValue})
    with pytest.raises(mole.MoleError) as ei:
        mole.VmcRunner(mole.Sampler.new(wfb, mole.MetropolisDiffuse.from_rng(0.05, SEED0), obsb),
                       mole.StochasticReconfiguration(0.02, 7)).run_optimization(1, 360, 10, 6)
    assert ei.value.code == mole.ffi.ERR_LINALG
    c = cases()["sj_ne"]
    iters, total, bs, nw = 3, 6 * 60, 10, 6
    cfg0 = orc.init_uniform(SEED0, 0, 10)
    opts = orc.run_options(orc.METROP_DIFFUSE, 0.05, nan_reject=1)
    ropt = orc.Optimizer(orc.OPT_SR, 7, 0.02)
    ref = orc.vmc_run_optimization(c["owf"], c["oham"], opts, ropt, SEED0, cfg0, iters, total, bs, nw)
    wf, op = c["make"](mole)
    obs = mole.operators(**{"Energy": op, "Parameter gradient": mole.ParameterGradient, "Wavefunction value": mole.WavefunctionValue})
    sampler = mole.Sampler.new(wf, mole.MetropolisDiffuse.from_rng(0.05, SEED0), obs)
    vmc = mole.VmcRunner(sampler, mole.StochasticReconfiguration(0.02, 7))
    _, en, er = vmc.run_optimization(iters, total, bs, nw)
    assert np.max(np.abs(en - ref["energies"]) / np.abs(ref["energies"])) < 1e-8
    assert np.max(np.abs(er - ref["errors"]) / ref["errors"]) < 1e-5
    assert np.max(np.abs(vmc.param_history - ref["param_history"]) / np.maximum(np.abs(ref["param_history"]), 1e-2)) < 1e-5


def test_dmc_step_slater_jastrow_matches_oracle(mole, orc):
    """One DMC time step (dmc.rs:87-130) of the cooperative Slater-Jastrow kernel, state re-synchronised
    on the oracle each step."""
    c = cases()["sj_li"]
    wf, op = c["make"](mole)
    W, tau, seed = 200, 0.01, bytes([3] * 32)
    m = mole.MetropolisDiffuse.from_rng(tau, seed)
    x = np.array([orc.init_normal(seed, w, 3, 0.8) for w in range(W)])
    w = np.ones(W)
    ens = mole.Ensemble(W, 3, seed)
    e_ref = -7.0
    for t in range(3):
        ens.set_configs(x); ens.set_weights(w); ens.step = t
        # the oracle's dmc_step uses the reference-faithful NaN policy; Li with tau = 0.01 from N(0, 0.8) stays finite
        e_o, tw_o, w2, x2 = orc.dmc_step(c["owf"], c["oham"], w, x, tau, tau, e_ref, seed, t)
        swe, sw = ens.dmc_step(wf, m.set_compat(mole.ffi.COMPAT_NAN_ACCEPT), op, tau, e_ref)
        assert abs(swe / sw - e_o) < 1e-9 * abs(e_o)
        assert close(ens.get_configs(), x2) and close(ens.get_weights(), w2, 1e-8)
        ens.branch(mole.ffi.BRANCH_SR)
        src = ens.branch_sources()
        w, x = np.full(W, w2.mean()), x2[src]
        assert close(ens.get_configs(), x)


@pytest.mark.parametrize("name", ["lcao_h2p", "lcao_h2_singlet", "lcao_h2_triplet"])
def test_dmc_step_lcao_matches_oracle(mole, orc, name):
    """One DMC time step (dmc.rs:87-130) with an LCAO guiding function, state re-synchronised on the oracle each step;
    the triplet has a node, so the fixed-node rejection (metrop.rs:178-180) is exercised."""
    c = cases()[name]
    wf, op = c["make"](mole)
    W, tau, seed, ne = 512, 0.02, bytes([5] * 32), c["ne"]
    m = mole.MetropolisDiffuse.from_rng(tau, seed).fix_nodes().set_compat(mole.ffi.COMPAT_NAN_ACCEPT)
    x = np.array([orc.init_normal(seed, w, ne, 0.9) for w in range(W)])
    w = np.ones(W)
    ens = mole.Ensemble(W, ne, seed)
    e_ref = -0.6 if ne == 1 else -1.0
    for t in range(3):
        ens.set_configs(x); ens.set_weights(w); ens.step = t
        e_o, tw_o, w2, x2 = orc.dmc_step(c["owf"], c["oham"], w, x, tau, tau, e_ref, seed, t)
        swe, sw = ens.dmc_step(wf, m, op, tau, e_ref)
        assert abs(swe / sw - e_o) < 1e-9 * abs(e_o) and abs(sw - tw_o) < 1e-10 * tw_o
        assert close(ens.get_configs(), x2) and close(ens.get_weights(), w2, 1e-8)
        ens.branch(mole.ffi.BRANCH_SR)
        src = ens.branch_sources()
        w, x = np.full(W, w2.mean()), x2[src]
        assert close(ens.get_configs(), x)


def test_dmc_lcao_h2plus_energy(mole):
    """DMC of H2+ at R = 2.5 guided by the LCAO sigma_g of tests/hydrogen_molecular_ion_lcao.rs:101-107: the guide is
    nodeless, so the run converges to the exact total energy -0.5938235 Ha (electronic -0.9938235 + 1/R), well below
    the guide's own VMC energy -0.56483."""
    c = cases()["lcao_h2p"]
    wf, op = c["make"](mole)
    m = mole.MetropolisDiffuse.from_rng(0.01, bytes([1] * 32)).fix_nodes()
    dmc = mole.DmcRunner.new(wf, 8192, -0.565, op, m, mole.SRBrancher.new(), identical_start=False)
    en, er = dmc.diffuse(0.01, 3000, 100, 10)
    assert abs(en[-1] + 0.5938235) < max(6 * er[-1], 8e-3)     # the oracle's run of 4000 walkers: -0.5908 +/- 0.0023
    assert en[-1] < -0.58


def test_slater_jastrow_long_run_no_drift(mole, orc):
    """The kernel carries the inverse Slater matrices and grad ln D through Sherman-Morrison updates and
    rebuilds them from scratch only every SJ_REFRESH_EVERY (8) sweeps; the oracle rebuilds everything for
    every evaluation.  Over 120 sweeps (1200 single-electron moves per walker) decisions stay bit-exact and
    E_L stays within 1e-9 of the oracle's."""
    c = cases()["sj_ne"]
    wf, op = c["make"](mole)
    W, steps, bs = 30, 120, 10
    seed = bytes([11] * 32)
    m = mole.MetropolisDiffuse(0.02, seed)
    # start from walkers equilibrated by box moves: a raw N(0, sigma) start puts some walkers next to a node,
    # where the drift is huge, t_high / t_low underflow and the very first decision is ill-conditioned
    ens = mole.Ensemble(W, c["ne"], seed)
    ens.init_normal(0.6)
    ens.sweep(wf, mole.MetropolisBox(0.4, seed), op, n_sweeps=60, observables=0)
    cfgs = ens.get_configs().copy()
    ens.step = 0
    obs = orc.OBS_ENERGY | orc.OBS_WFVALUE | orc.OBS_PGRAD
    ref = orc.ensemble_run(c["owf"], c["oham"], orc.run_options(orc.METROP_DIFFUSE, 0.02, obs, nan_reject=1), cfgs, seed, steps, bs)
    got = ens.sweep(wf, m, op, n_sweeps=steps, n_discard=bs, block_size=bs, observables=obs, traces=("energy", "pgrad", "accept"))
    assert np.array_equal(got["accept"], ref["accept"])
    assert close(got["energy"], ref["energy"], 1e-9) and close(ens.get_configs(), ref["cfgs"])
    assert close(got["pgrad"][:, -10:], ref["pgrad"][:, -10:, :c["np"]], 1e-9)
