#!/usr/bin/env python
"""examples/hydrogen_molecule.rs on the GPU path: H2 Heitler-London VMC with stochastic reconfiguration and
steepest descent, then DMC from the SR-optimised wavefunction (flow and constants of the reference's main(),
examples/hydrogen_molecule.rs:170-231; the plots are out of scope).

    python examples/hydrogen_molecule.py                 # 2^16 walkers per VMC iteration, 2^14 DMC walkers
    python examples/hydrogen_molecule.py --faithful      # the reference's 8 workers / 100 DMC walkers
"""
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

NITERS, BLOCK_SIZE = 10, 10                                   # :171,181


def optimize_wave_function(ion_pos, wave_function, opt, nworkers, total_samples, compat):
    hamiltonian = m.ElectronicHamiltonian.from_ions(ion_pos, [1, 1])
    obs = m.operators(**{"Energy": hamiltonian, "Parameter gradient": m.ParameterGradient,
                         "Wavefunction value": m.WavefunctionValue})
    sampler = m.Sampler.new(wave_function, m.MetropolisDiffuse.from_rng(0.25, bytes(32)), obs, compat=compat)      # :250-255
    return m.VmcRunner(sampler, opt).run_optimization(NITERS, total_samples, BLOCK_SIZE, nworkers, verbose=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--faithful", action="store_true", help="reference sizes: 8 workers x 2500 sweeps, 100 DMC walkers")
    ap.add_argument("--dmc-iters", type=int, default=None)
    a = ap.parse_args()
    nworkers = 8 if a.faithful else 1 << 16
    total_samples = 20_000 if a.faithful else 250 * nworkers      # :176: 2500 sweeps per worker / 250 at ensemble scale
    ion_pos = [[-0.7, 0.0, 0.0], [0.7, 0.0, 0.0]]
    sep = ion_pos[1][0] - ion_pos[0][0]

    # --faithful: the reference's arithmetic and step sizes (:191-198); otherwise the intended O_k and steps to match
    compat = COMPAT if a.faithful else 0
    sr_step, sd_step = (50_000.0, 1e-5) if a.faithful else (0.2, 0.5)
    print("STOCHASTIC RECONFIGURATION")
    sr_wf, energies_sr, errors_sr = optimize_wave_function(ion_pos, m.HydrogenMoleculeWaveFunction(sep, [0.5]),
                                                           m.StochasticReconfiguration(sr_step, 1, compat=compat), nworkers,
                                                           total_samples, compat)
    print("\nSTEEPEST DESCENT")
    optimize_wave_function(ion_pos, m.HydrogenMoleculeWaveFunction(sep, [0.5]), m.SteepestDescent(sd_step, 1, compat=compat),
                           nworkers, total_samples, compat)
    print()

    num_walkers = 100 if a.faithful else 1024                     # :207-211
    tau, num_iters, dmc_block, num_eq = 1e-2, a.dmc_iters or (40_000 if a.faithful else 4_000), 100, 10
    hamiltonian = m.ElectronicHamiltonian.from_ions(ion_pos, [1, 1])
    metrop = m.MetropolisDiffuse.from_rng(tau, bytes(32)).fix_nodes()
    dmc = m.DmcRunner.new(sr_wf, num_walkers, float(energies_sr[-1]), hamiltonian, metrop, m.SRBrancher.new(),
                          identical_start=a.faithful)
    energies, errs = dmc.diffuse(tau, num_iters, dmc_block, num_eq)
    print("DMC Energy:   %.8f +/- %.8f   (exact -1.17447)" % (energies[-1], errs[-1]))


if __name__ == "__main__":
    main()
