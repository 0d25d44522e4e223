"""Edge cases of the Metropolis acceptance: NaN ratios, ragged ensemble sizes, empty schedules.

Upstream `acceptance.min(1.0)` uses Rust's NaN-dropping f64::min (metrop.rs:80,195): a NaN ratio is
ACCEPTED.  The product rejects NaN ratios by default and reproduces the upstream behaviour under
MOLE_COMPAT_NAN_ACCEPT; both policies are checked against the oracle."""
import numpy as np
import pytest

from common import SEED0, cases

pytestmark = pytest.mark.gpu


def _far_configs(W):
    # psi underflows to 0 at these radii -> psi'^2/psi^2 = 0/0 = NaN
    rng = np.random.default_rng(9)
    x = rng.normal(size=(W, 1, 3))
    x[::2] *= 1e3
    return x


@pytest.mark.parametrize("nan_accept", [False, True])
def test_nan_ratio_policy_box(mole, orc, nan_accept):
    W, steps, bs = 64, 20, 10
    cfgs = _far_configs(W)
    wf = mole.GaussianWaveFunction(1.0)
    op = mole.HarmonicHamiltonian(1.0)
    m = mole.MetropolisBox(1.0, SEED0)
    if nan_accept:
        m.set_compat(mole.ffi.COMPAT_NAN_ACCEPT)
    ens = mole.Ensemble(W, 1, SEED0)
    ens.set_configs(cfgs)
    got = ens.sweep(wf, m, op, n_sweeps=steps, n_discard=bs, block_size=bs, observables=0, traces=("accept",))
    c = cases()["gauss_sho"]
    ref = orc.ensemble_run(c["owf"], c["oham"], orc.run_options(orc.METROP_BOX, 1.0, 0, nan_reject=0 if nan_accept else 1),
                           cfgs, SEED0, steps, bs)
    assert np.array_equal(got["accept"], ref["accept"])
    far = got["accept"][::2]
    assert far.all() if nan_accept else not far.any()
    assert 0 < got["accept"][1::2].mean() < 1
    assert np.allclose(ens.get_configs(), ref["cfgs"], rtol=1e-12, atol=0)


def test_nan_psi_is_rejected_by_node_test(mole, orc):
    """Diffusion sampler: psi = 0 -> drift NaN -> psi' NaN -> signum(NaN) != signum(psi): rejected upstream too."""
    W, steps, bs = 32, 20, 10
    cfgs = _far_configs(W)
    wf = mole.GaussianWaveFunction(1.0)
    op = mole.HarmonicHamiltonian(1.0)
    for compat in (0, mole.ffi.COMPAT_NAN_ACCEPT):
        m = mole.MetropolisDiffuse(0.1, SEED0).set_compat(compat)
        ens = mole.Ensemble(W, 1, SEED0)
        ens.set_configs(cfgs)
        got = ens.sweep(wf, m, op, n_sweeps=steps, block_size=bs, observables=0, traces=("accept",))
        c = cases()["gauss_sho"]
        ref = orc.ensemble_run(c["owf"], c["oham"], orc.run_options(orc.METROP_DIFFUSE, 0.1, 0, nan_reject=0 if compat else 1),
                               cfgs, SEED0, steps, bs)
        assert np.array_equal(got["accept"], ref["accept"])
        assert not got["accept"][::2].any()


@pytest.mark.parametrize("W", [1, 5, 7, 129, 1000])
def test_ragged_ensemble_sizes(mole, orc, W):
    """Ensemble sizes that do not fill a warp / CTA / the 6-walker Slater-Jastrow warp layout."""
    for name in ("he", "sj_be"):
        c = cases()[name]
        wf, op = c["make"](mole)
        tau = 0.25 if name == "he" else 0.02
        m = mole.MetropolisDiffuse(tau, SEED0)
        ens = mole.Ensemble(W, c["ne"], SEED0)
        ens.init_uniform()
        cfgs = ens.get_configs()
        got = ens.sweep(wf, m, op, n_sweeps=20, n_discard=10, block_size=10, traces=("energy", "accept"))
        ref = orc.ensemble_run(c["owf"], c["oham"], orc.run_options(orc.METROP_DIFFUSE, tau, orc.OBS_ENERGY, nan_reject=1), cfgs, SEED0, 20, 10)
        assert np.array_equal(got["accept"], ref["accept"])
        assert np.allclose(got["energy"], ref["energy"], rtol=1e-8, atol=1e-8)
        acc = ens.acc_get()
        assert acc.n_samples == W * 10 and acc.n_blocks == W and acc.n_moves == W * 20 * c["ne"]


def test_empty_schedules(mole):
    wf = mole.HeliumAtomWaveFunction(1.69)
    op = mole.ElectronicHamiltonian.from_ions([[0, 0, 0]], [2])
    m = mole.MetropolisDiffuse(0.1, SEED0)
    ens = mole.Ensemble(16, 2, SEED0)
    ens.init_uniform()
    before = ens.get_configs()
    ens.sweep(wf, m, op, n_sweeps=0)                       # nothing to do
    assert np.array_equal(ens.get_configs(), before) and ens.acc_get().n_samples == 0 and ens.step == 0
    ens.sweep(wf, m, op, n_sweeps=5, n_discard=5)          # moved but never sampled (pure equilibration)
    acc = ens.acc_get()
    assert acc.n_samples == 0 and acc.n_moves == 16 * 5 * 2 and ens.step == 5
    with pytest.raises(mole.MoleError):
        mole.acc_finalize(acc)                             # DataAccessError: no "Energy" samples
