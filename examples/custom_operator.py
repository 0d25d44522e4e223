#!/usr/bin/env python
"""examples/custom_operator.rs on the GPU path: 3-d harmonic oscillator with the exact Gaussian ground state,
MetropolisBox(1.0), Runner::run(1000, 1) and a Logger printing the block energy; the local energy is exactly
1.5 for every sample (examples/custom_operator.rs:100-139)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import mole_b200 as m  # noqa: E402


class Logger:                                                       # custom_operator.rs:13-27
    def log(self, data):
        return "Energy: %s" % data["block_energy"]


def main():
    omega = 1.0
    ansatz = m.GaussianWaveFunction(np.sqrt(2.0 / omega))
    metrop = m.MetropolisBox.from_rng(1.0, bytes(32))
    hamiltonian = m.HarmonicHamiltonian(omega)
    sampler = m.Sampler.new(ansatz, metrop, m.operators(**{"Energy": hamiltonian}), n_walkers=256, independent=True)
    m.Runner(sampler, Logger()).run(20, 1, traces=False)              # logged: one line per block
    result = m.Runner(m.Sampler.new(ansatz, metrop, m.operators(**{"Energy": hamiltonian}), n_walkers=256,
                                    independent=True)).run(1000, 1)
    energy_data = result.data["Energy"]
    energy, error = energy_data.mean(), energy_data.std()
    assert abs(energy - 1.5) < 1e-14 and error < 1e-14
    print("\nEnergy:     %s +/- %.8f" % (energy, error))


if __name__ == "__main__":
    main()
