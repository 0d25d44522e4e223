#!/usr/bin/env python
"""A hydrogen chain with the Hydrogen1sBasis / Orbital / SpinDeterminantProduct API the reference's LCAO tests name
(tests/helium_lcao.rs:94-101, tests/hydrogen_molecular_ion_lcao.rs:103-107) carried beyond two electrons and two
centres, times the electron-electron Jastrow of theory/jastrow.tex: every orbital coefficient and the Jastrow
parameters are variational (H8: P = 36), so the optimisation runs on the large-P path - the sweep stores the
per-sample rows (1, E_L, O_k), their Gram matrix is contracted on the fp64 tensor pipe (mma.m8n8k4) and the
stochastic-reconfiguration step of src/optimize/src/optimizers.rs:181-264 is solved from it on the host."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import mole_b200 as m  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--atoms", type=int, default=8, choices=[2, 4, 6, 8])
    ap.add_argument("--spacing", type=float, default=1.4)
    ap.add_argument("--walkers", type=int, default=4096)
    ap.add_argument("--iterations", type=int, default=10)
    ap.add_argument("--sweeps", type=int, default=60)
    a = ap.parse_args()
    n, nup = a.atoms, a.atoms // 2
    ion_pos = np.array([[a.spacing * (i - 0.5 * (n - 1)), 0.0, 0.0] for i in range(n)])
    basis = m.Hydrogen1sBasis(ion_pos, [1.0], general=True)      # more than two centres: the general LCAO kind
    # particle-in-a-box molecular orbitals as the starting coefficients: C[k][c] = sin(pi (k + 1)(c + 1) / (n + 1))
    orbitals = [m.Orbital(np.sin(np.pi * (k + 1) * (np.arange(n) + 1) / (n + 1)).reshape(n, 1), basis.clone()) for k in range(nup)]
    wave_function = m.LcaoSlaterJastrow.from_orbitals(orbitals, nup, nup, b=(0.5, 1.0, 0.0, 0.0))
    P = wave_function.num_parameters()
    hamiltonian = m.ElectronicHamiltonian.from_ions(ion_pos, [1] * n)
    seed = bytes(32)
    ens = m.Ensemble(a.walkers, n, seed)
    ens.init_uniform(-0.5 * a.spacing * n, 0.5 * a.spacing * n)
    ens.sweep(wave_function, m.MetropolisBox.from_rng(1.0, seed), hamiltonian, n_sweeps=100, observables=0)
    metrop = m.MetropolisDiffuse.from_rng(0.05, seed)
    obs = m.operators(**{"Energy": hamiltonian, "Parameter gradient": m.ParameterGradient, "Wavefunction value": m.WavefunctionValue})
    sampler = m.Sampler.with_initial_configuration(wave_function, metrop, obs, ens.get_configs(), n_walkers=a.walkers, independent=True)
    optimizer = m.StochasticReconfiguration(0.05, P).set_regularization(1.01, 1e-2)
    runner = m.VmcRunner(sampler, optimizer)
    print("H%d chain, %d up + %d down electrons, P = %d variational parameters, %d walkers" % (n, nup, nup, P, a.walkers))
    _, energies, errors = runner.run_optimization(a.iterations, a.sweeps * a.walkers, 10, a.walkers, restart_each_iter=False)
    for it, (e, de) in enumerate(zip(energies, errors)):
        print("Energy: %.6f +/- %.6f   (iteration %d)" % (e, de, it))
    print("non-finite samples skipped: %d" % runner.ensemble.health()[0])
    assert np.isfinite(energies).all() and energies[-1] < energies[0]
    return energies


if __name__ == "__main__":
    main()
