#!/usr/bin/env python
"""tests/hydrogen_molecular_ion_lcao.rs and tests/helium_lcao.rs on the GPU path, written with the
Hydrogen1sBasis / Orbital / SingleDeterminant / SpinDeterminantProduct API those tests name but keep commented
out (hydrogen_molecular_ion_lcao.rs:101-107, helium_lcao.rs:92-101), and with the tests' own samplers, run lengths
and acceptance criteria (:123-140 and :117-135)."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import mole_b200 as m  # noqa: E402


def hydrogen_molecular_ion_lcao(n_walkers):
    ion_pos = np.array([[-1.25, 0.0, 0.0], [1.25, 0.0, 0.0]])
    basis = m.Hydrogen1sBasis(ion_pos, [1.0])
    orbitals = [m.Orbital(np.array([[1.0], [1.0]]), basis.clone())]
    wave_function = m.SingleDeterminant(orbitals)
    hamiltonian = m.ElectronicHamiltonian(m.KineticEnergy(), m.IonicPotential(ion_pos, [1, 1]), m.ElectronicPotential())
    metrop = m.MetropolisBox.from_rng(1.0, bytes(32))
    sampler = m.Sampler.new(wave_function, metrop, m.operators(**{"Energy": hamiltonian}), n_walkers=n_walkers, independent=True)
    result = m.Runner(sampler).run(10000, 100)
    energy_data = result.data["Energy"]
    energy, energy_err = energy_data.mean(), energy_data.std()
    print("H2+ LCAO   Energy: %.6f  (sample std %.4f, %d walkers)" % (energy, energy_err, n_walkers))
    assert abs(energy - (-0.565)) < energy_err                       # :139-140
    return energy


def helium_lcao(n_walkers):
    optimal_width = 1.0 / 1.69
    ion_pos = np.array([[0.0, 0.0, 0.0]])
    basis = m.Hydrogen1sBasis(ion_pos, [optimal_width])
    orbitals = [m.Orbital(np.array([[1.0]]), basis.clone()), m.Orbital(np.array([[1.0]]), basis.clone())]
    wave_function = m.SpinDeterminantProduct(orbitals, 1)
    hamiltonian = m.ElectronicHamiltonian(m.KineticEnergy(), m.IonicPotential(ion_pos, [2]), m.ElectronicPotential())
    metrop = m.MetropolisDiffuse.from_rng(0.1, bytes(32))
    sampler = m.Sampler.new(wave_function, metrop, m.operators(**{"Energy": hamiltonian}), n_walkers=n_walkers, independent=True)
    result = m.Runner(sampler).run(1000, 100)
    energy_data = result.data["Energy"]
    energy, energy_err = energy_data.mean(), energy_data.std()
    exact_result = 0.5 * 1.5 ** 6 * (-0.5)                           # :131
    print("He LCAO    Energy: %.6f  (sample std %.4f, %d walkers)  reference criterion %.6f, closed form %.6f"
          % (energy, energy_err, n_walkers, exact_result, 1.69 ** 2 - 27.0 / 8.0 * 1.69))
    assert abs(energy - exact_result) < 2.0 * energy_err             # :134
    return energy


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--walkers", type=int, default=4096, help="independent chains (the reference runs one)")
    a = ap.parse_args()
    hydrogen_molecular_ion_lcao(a.walkers)
    helium_lcao(a.walkers)
