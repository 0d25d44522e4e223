#!/usr/bin/env python
"""examples/dmc.rs on the GPU path: hydrogen atom, Gaussian guide optimised by VMC + SR, then DMC with SRBrancher
(examples/dmc.rs:152-227).  With its cusp-less Gaussian guide the reference algorithm has no local-energy or
weight cut-off; large populations collapse onto the nucleus sooner or later, so the default
keeps the reference's 100 walkers and --walkers scales it up at the user's risk (an STO guide, --sto, is stable)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import mole_b200 as m  # noqa: E402

# The reference's step sizes (SR 50 000 / 100 000, SD 1e-5) only make sense with its own arithmetic: the stored
# "Parameter gradient" sample is 1/(d psi/d p) because `Vector / Scalar` computes scalar / array
# (src/operator/src/traits.rs:149-150).  The examples therefore run with the reference-compatible flags; drop
# COMPAT (and use steps of order 0.05) for the intended O_k = (d psi/d p)/psi.
COMPAT = m.ffi.COMPAT_VECTOR_DIV | m.ffi.COMPAT_SR_SUBTRACT

ITERS, TOTAL_SAMPLES, BLOCK_SIZE = 100, 5000, 10


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--walkers", type=int, default=100)
    ap.add_argument("--dmc-iters", type=int, default=100_000)
    ap.add_argument("--sto", action="store_true", help="use the 1s STO guide (examples/dmc.rs:157, commented alternative)")
    a = ap.parse_args()
    ansatz = m.STO(0.8) if a.sto else m.GaussianWaveFunction(1.0)
    metrop = m.MetropolisDiffuse.from_rng(0.1, bytes(32))
    hamiltonian = m.ElectronicHamiltonian.from_ions([[0.0, 0.0, 0.0]], [1])
    obs = m.operators(**{"Energy": hamiltonian, "Parameter gradient": m.ParameterGradient,
                         "Wavefunction value": m.WavefunctionValue})
    sampler = m.Sampler.new(ansatz, metrop, obs, compat=COMPAT)
    guiding_wf, energies, errors = m.VmcRunner(sampler, m.StochasticReconfiguration(1.0, 1, compat=COMPAT)).run_optimization(
        ITERS, TOTAL_SAMPLES, BLOCK_SIZE, 4)
    print("\nVMC Energy:     %s +/- %.8f\n" % (energies[-1], errors[-1]))

    tau, dmc_block, num_eq = 0.025, 400, 10
    metrop = m.MetropolisDiffuse.from_rng(tau, bytes([1] * 32)).fix_nodes()
    dmc = m.DmcRunner.new(guiding_wf, a.walkers, float(energies[-1]), hamiltonian, metrop, m.SRBrancher.new())
    dmc_energy, dmc_errs = dmc.diffuse(tau, a.dmc_iters, dmc_block, num_eq)
    print("\nDMC Energy:   %.8f +/- %.8f   (exact -0.5)" % (dmc_energy[-1], dmc_errs[-1]))


if __name__ == "__main__":
    main()
