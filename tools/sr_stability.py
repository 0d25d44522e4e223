"""SR stability of the bench workload (Ne Slater-Jastrow, P = 7): runs ITERS optimisation iterations exactly as
bench.py's step does for several (step, diag_scale, diag_shift) settings and prints, per iteration, the energy,
the parameters, cond(S) of the regularised matrix, |dp| and the health counters.
   python tools/sr_stability.py [W] [ITERS] [SWEEPS]
VERDICT r1 weak#1: with the reference's diag x 1.01 only (optimizers.rs:225-231) the nearly redundant Jastrow pair
(b1, b2) lets cond(S) climb to 1e10 and the loop diverges after ~20 iterations at step 0.005."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mole_b200 as m  # noqa: E402

W = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 14
ITERS = int(sys.argv[2]) if len(sys.argv) > 2 else 30
SWEEPS = int(sys.argv[3]) if len(sys.argv) > 3 else 200
SEED = bytes(32)
ctx = m.default_context()
op = m.ElectronicHamiltonian.from_ions([[0, 0, 0]], [10])
met = m.MetropolisDiffuse.from_rng(0.02, SEED)
obs = m.ffi.OBS_ENERGY | m.ffi.OBS_PGRAD | m.ffi.OBS_WFVALUE


def run(step, scale, shift, iters=ITERS, verbose=True):
    wf = m.SlaterJastrow(5, 5, (9.64, 2.88, 2.88), (0.5, 1.0, 0.0, 0.0), 1.0)
    ens = m.Ensemble(W, 10, SEED)
    ens.init_normal(0.5)
    ens.sweep(wf, m.MetropolisBox.from_rng(0.5, SEED), op, n_sweeps=200, observables=0)
    ens.sweep(wf, met, op, n_sweeps=50, observables=0)
    opt = m.StochasticReconfiguration(step, 7).set_regularization(scale, shift)
    print("== step %g  diag_scale %g  diag_shift %g  (W=%d, %d sweeps/iter)" % (step, scale, shift, W, SWEEPS), flush=True)
    out = []
    for it in range(iters):
        ens.reseed(m.derive_seed(SEED, it))
        ens.acc_reset()
        ens.sweep(wf, met, op, n_sweeps=SWEEPS, n_discard=10, block_size=10, observables=obs)
        acc = ens.acc_get()
        e, err, accp, g = m.acc_finalize(acc)
        S = opt.sr_matrix(acc)
        try:
            dp = opt.compute_parameter_update(wf.parameters(), acc)
        except m.MoleError as ex:
            print("   it %2d  E %.5f +/- %.5f  REFUSED: %s" % (it, e, err, ex))
            break
        wf.update_parameters(dp)
        p = wf.parameters()
        out.append((e, err))
        if verbose:
            print("   it %2d  E %.5f +/- %.5f  acc %.3f  cond(S) %.2e  |dp| %.2e  bad %s  p = %s" % (
                it, e, err, accp, np.linalg.cond(S), np.linalg.norm(dp), ens.health(), np.array2string(p, precision=4)), flush=True)
    return out


if __name__ == "__main__":
    for step, scale, shift in ((0.005, 1.01, 0.0), (0.005, 1.01, 1e-3), (0.02, 1.01, 1e-2), (0.05, 1.0, 1e-2), (0.005, 1.0, 1e-2)):
        run(step, scale, shift)
