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
    # LCAO determinants over a hydrogen-1s basis (the API named at tests/helium_lcao.rs:94-101 and
    # tests/hydrogen_molecular_ion_lcao.rs:103-107), same numbers as tests/golden/make_golden.py
    def lcao(name, kind, coeff, width_inv, pos, charges, mk):
        geom = [1 if name.endswith("triplet") else 0, width_inv] + list(np.asarray(pos, dtype=float).reshape(-1)) + [0.0] * (6 - 3 * len(pos))
        add(name, O.wf_desc(kind, coeff, geom), O.ham_desc(O.HAM_ELECTRONIC, pos, charges),
            lambda m: (mk(m), m.ElectronicHamiltonian.from_ions(pos, charges)))

    def orbs(m, pos, width, rows):
        basis = m.Hydrogen1sBasis(pos, [width])
        return [m.Orbital(np.array(r).reshape(-1, 1), basis.clone()) for r in rows]

    p2 = [[-1.25, 0, 0], [1.25, 0, 0]]
    lcao("lcao_h2p", O.WF_LCAO_1E_2C, [1.0, 1.0], 1.0, p2, [1, 1], lambda m: m.SingleDeterminant(orbs(m, p2, 1.0, [[1.0, 1.0]])))
    lcao("lcao_he", O.WF_LCAO_2E_1C, [1.0, 1.0], 1.0 / (1.0 / 1.69), [[0, 0, 0]], [2],
         lambda m: m.SpinDeterminantProduct(orbs(m, [[0, 0, 0]], 1.0 / 1.69, [[1.0], [1.0]]), 1))
    ps = [[-0.7, 0, 0], [0.7, 0, 0]]
    lcao("lcao_h2_singlet", O.WF_LCAO_2E_2C, [1.0, 0.9, 0.8, 1.1], 1.0 / 0.85, ps, [1, 1],
         lambda m: m.SpinDeterminantProduct(orbs(m, ps, 0.85, [[1.0, 0.9], [0.8, 1.1]]), 1))
    pt = [[-0.7, 0.1, 0], [0.7, -0.1, 0.2]]
    lcao("lcao_h2_triplet", O.WF_LCAO_2E_2C, [1.0, 1.0, 1.0, -1.0], 1.0 / 0.85, pt, [1, 1],
         lambda m: m.SingleDeterminant(orbs(m, pt, 0.85, [[1.0, 1.0], [1.0, -1.0]])))
    # general LCAO Slater-Jastrow, same numbers as tests/golden/make_golden.py
    def lsj(name, nup, ndn, pos, alphas, C, b, kappa=1.0):
        geom = [kappa, nup, ndn, len(pos), 0, 0, 0, 0]
        for i, R in enumerate(pos):
            geom += list(R) + [alphas[i]]
        params = [v for row in C for v in row] + list(b)
        z = [1] * len(pos)
        add(name, O.wf_desc(O.WF_LCAO_SJ, params, geom), O.ham_desc(O.HAM_ELECTRONIC, pos, z),
            lambda m: (m.LcaoSlaterJastrow(nup, ndn, pos, [1.0 / a for a in alphas], C, b, kappa),
                       m.ElectronicHamiltonian.from_ions(pos, z)))

    h4 = [[-2.1, 0, 0], [-0.7, 0, 0], [0.7, 0, 0], [2.1, 0, 0]]
    lsj("lsj_h4", 2, 2, h4, [1.0, 1.1, 1.1, 1.0], [[1, 1, 1, 1], [1, 0.5, -0.5, -1]], [0.5, 1.0, 0.1, -0.05])
    lsj("lsj_h3", 2, 1, [[-1.0, 0, 0], [0.6, 0.8, 0], [0.5, -0.7, 0.4]], [1.2, 0.9, 1.0], [[1, 0.9, 0.8], [1, -0.4, -0.7]],
        [0.4, 0.8, 0.0, 0.0], kappa=1.5)
    h8 = [[1.4 * (i - 3.5), 0.3 * ((-1) ** i), 0.1 * i] for i in range(8)]
    h8c = [[1.0, 1.1, 1.2, 1.3, 1.3, 1.2, 1.1, 1.0], [1.0, 0.8, 0.5, 0.2, -0.2, -0.5, -0.8, -1.0],
           [1.0, 0.3, -0.6, -0.9, -0.9, -0.6, 0.3, 1.0], [0.7, -0.5, -0.9, 0.4, -0.4, 0.9, 0.5, -0.7]]
    lsj("lsj_h8", 4, 4, h8, [1.0] * 8, h8c, [0.5, 1.0, 0.05, 0.02])
    return c


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)) if a.size else 0.0


def random_cfgs(W, ne, seed=1, scale=1.0):
    rng = np.random.default_rng(seed)
    return rng.normal(0.0, scale, size=(W, ne, 3))
