"""Shared builders: the same physical setup expressed for the oracle and for the CUDA library."""
import numpy as np

import oracle as O

SEED0 = bytes(32)            # StdRng::from_seed([0u8; 32]) in every reference test
CFG2 = np.array([[0.3, -0.2, 0.5], [-0.6, 0.1, 0.25]])   # SURVEY.md §8(c) golden configuration

# name -> (oracle wf desc, oracle ham desc, builder of (mole wf, mole op))
def cases(mole=None):
    c = {}

    def add(name, owf, oham, mk):
        c[name] = dict(owf=owf, oham=oham, make=mk, ne=owf.n_elec, np=owf.n_params)

    add("h2", O.wf_desc(O.WF_H2_HL_STO, [0.5], [1.4]),
        O.ham_desc(O.HAM_ELECTRONIC, [[-0.7, 0, 0], [0.7, 0, 0]], [1, 1]),
        lambda m: (m.HydrogenMoleculeWaveFunction(1.4, [0.5]),
                   m.ElectronicHamiltonian.from_ions([[-0.7, 0, 0], [0.7, 0, 0]], [1, 1])))
    add("he", O.wf_desc(O.WF_STO_PRODUCT, [1.69]), O.ham_desc(O.HAM_ELECTRONIC, [[0, 0, 0]], [2]),
        lambda m: (m.HeliumAtomWaveFunction(1.69), m.ElectronicHamiltonian.from_ions([[0, 0, 0]], [2])))
    h2p = O.wf_desc(O.WF_H2P_PRODUCT, [1.0], [2.5])
    h2p.n_params = 0
    add("h2p", h2p, O.ham_desc(O.HAM_ELECTRONIC, [[-1.25, 0, 0], [1.25, 0, 0]], [1, 1]),
        lambda m: (m.H2WF(2.5, 1.0), m.ElectronicHamiltonian(m.KineticEnergy(), m.IonicPotential([[-1.25, 0, 0], [1.25, 0, 0]], [1, 1]),
                                                               m.ElectronicPotential())))
    add("gauss_sho", O.wf_desc(O.WF_GAUSSIAN, [1.0]), O.ham_desc(O.HAM_HARMONIC, frequency=1.0),
        lambda m: (m.GaussianWaveFunction(1.0), m.HarmonicHamiltonian(1.0)))
    add("gauss_h", O.wf_desc(O.WF_GAUSSIAN, [1.0]), O.ham_desc(O.HAM_ELECTRONIC, [[0, 0, 0]], [1]),
        lambda m: (m.GaussianWaveFunction(1.0), m.ElectronicHamiltonian.from_ions([[0, 0, 0]], [1])))
    add("sto_h", O.wf_desc(O.WF_STO_1S, [0.8]), O.ham_desc(O.HAM_ELECTRONIC, [[0, 0, 0]], [1]),
        lambda m: (m.STO(0.8), m.ElectronicHamiltonian.from_ions([[0, 0, 0]], [1])))
    def sj(name, nup, ndn, params, kappa, Z):
        add(name, O.wf_desc(O.WF_SLATER_JASTROW, params, [kappa, nup, ndn]), O.ham_desc(O.HAM_ELECTRONIC, [[0, 0, 0]], [Z]),
            lambda m: (m.SlaterJastrow(nup, ndn, params[:3], params[3:], kappa), m.ElectronicHamiltonian.from_ions([[0, 0, 0]], [Z])))

    sj("sj_ne", 5, 5, [9.64, 2.88, 2.88, 0.5, 1.0, 0.1, -0.05], 1.0, 10)     # SURVEY.md §8(c) synthetic config 5
    sj("sj_be", 2, 2, [3.68, 0.96, 0.96, 0.5, 1.0, 0.2, 0.1], 1.0, 4)
    sj("sj_li", 2, 1, [2.69, 0.64, 0.64, 0.4, 0.8, 0.0, 0.0], 1.5, 3)
    return c


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)) if a.size else 0.0


def random_cfgs(W, ne, seed=1, scale=1.0):
    rng = np.random.default_rng(seed)
    return rng.normal(0.0, scale, size=(W, ne, 3))
