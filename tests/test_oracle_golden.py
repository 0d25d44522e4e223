"""Pins the CPU oracle (oracle/) before it is trusted as the parity checker:
golden vectors of SURVEY.md §8(c), 40-digit mpmath fixtures (tests/golden/wf_golden.json, made by
tests/golden/make_golden.py), and the Random123 Philox4x32-10 known-answer vectors."""
import json
import os

import numpy as np
import pytest

from common import CFG2, SEED0, cases, rel_err

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "wf_golden.json")))


def _oracle_case(orc, name):
    g = GOLD[name]
    kind = {"h2": orc.WF_H2_HL_STO, "he": orc.WF_STO_PRODUCT, "h2p": orc.WF_H2P_PRODUCT, "gauss_sho": orc.WF_GAUSSIAN,
            "gauss_h": orc.WF_GAUSSIAN, "sto_h": orc.WF_STO_1S, "sj_ne": orc.WF_SLATER_JASTROW,
            "sj_be": orc.WF_SLATER_JASTROW, "sj_li": orc.WF_SLATER_JASTROW, "lcao_h2p": orc.WF_LCAO_1E_2C,
            "lcao_he": orc.WF_LCAO_2E_1C, "lcao_h2_singlet": orc.WF_LCAO_2E_2C, "lcao_h2_triplet": orc.WF_LCAO_2E_2C,
            "lsj_h4": orc.WF_LCAO_SJ, "lsj_h3": orc.WF_LCAO_SJ, "lsj_h8": orc.WF_LCAO_SJ}[name]
    wf = orc.wf_desc(kind, g["params"], g["geom"], n_elec=g["n_elec"])
    wf.n_params = g["n_params"]
    if g["ham"][0] == "electronic":
        ham = orc.ham_desc(orc.HAM_ELECTRONIC, g["ham"][1], g["ham"][2])
    else:
        ham = orc.ham_desc(orc.HAM_HARMONIC, frequency=g["ham"][1])
    return g, wf, ham


def test_philox_known_answers(orc):
    # Random123 v1.09 kat_vectors, philox4x32 10 rounds
    assert orc.philox([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert orc.philox([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert orc.philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_stream_contract(orc):
    # key folding, uniform range, Box-Muller moments
    seed = bytes((37 * i * i + 11 * i + 5) % 256 for i in range(32))
    w = [int.from_bytes(seed[4 * i:4 * i + 4], "little") for i in range(8)]
    assert orc.key_from_seed(seed) == [w[0] ^ w[2] ^ w[4] ^ w[6], w[1] ^ w[3] ^ w[5] ^ w[7]]
    u = np.array([orc.draw_move(0, SEED0, w, 3, orc.DOM_MOVE, 1) for w in range(4000)])
    assert u.min() >= 0.0 and u.max() < 1.0
    assert abs(u.mean() - 0.5) < 0.01
    n = np.array([orc.draw_move(1, SEED0, w, 7, orc.DOM_MOVE, 0) for w in range(6000)])[:, :3]
    assert abs(n.mean()) < 0.03 and abs(n.var() - 1.0) < 0.05
    assert abs(np.corrcoef(n[:, 0], n[:, 1])[0, 1]) < 0.05 and abs(np.corrcoef(n[:, 0], n[:, 2])[0, 1]) < 0.05
    # draws are a pure function of the coordinates
    assert orc.draw_move(1, SEED0, 5, 9, orc.DOM_MOVE, 1) == orc.draw_move(1, SEED0, 5, 9, orc.DOM_MOVE, 1)
    assert orc.draw_move(1, SEED0, 5, 9, orc.DOM_MOVE, 1) != orc.draw_move(1, SEED0, 5, 9, orc.DOM_MOVE, 0)


def test_survey_golden_vectors(orc):
    """SURVEY.md §8(c) table (computed there from the reference's closed forms at 40 digits)."""
    c = cases()
    tol = 1e-14
    h2, ham = c["h2"]["owf"], c["h2"]["oham"]
    assert rel_err(orc.wf_value(h2, CFG2), 0.91118742899867764155) < tol
    assert rel_err(orc.wf_gradient(h2, CFG2), [[0.056207885784915448003, 0.11802438549810129697, -0.29506096374525324243],
                                               [0.03504617929555360101, -0.11881133930312649258, -0.29702834825781623145]]) < 5e-14
    assert rel_err(orc.wf_laplacian(h2, CFG2), -3.1008769265442040005) < tol
    assert rel_err(orc.wf_parameter_gradient(h2, CFG2), [-1.3122155514819913489]) < tol
    assert rel_err(orc.ionic_potential(ham, CFG2), -5.8916738733084355192) < tol
    assert rel_err(orc.electronic_potential(CFG2), 1.0192943828752510893) < tol
    assert rel_err(orc.local_energy(ham, h2, CFG2), -3.1708212665934303038) < tol
    he, hamhe = c["he"]["owf"], c["he"]["oham"]
    assert rel_err(orc.wf_value(he, CFG2), 0.11611085002507127485) < tol
    assert rel_err(orc.wf_laplacian(he, CFG2), -0.57015267582460947911) < tol
    assert rel_err(orc.wf_parameter_gradient(he, CFG2), [-0.14793552454144074405]) < tol
    assert rel_err(orc.local_energy(hamhe, he, CFG2), -2.8110692938173301994) < tol
    p, hp = c["h2p"]["owf"], c["h2p"]["oham"]
    assert rel_err(orc.wf_value(p, CFG2[:1]), 0.065030402248653389637) < tol
    assert rel_err(orc.wf_laplacian(p, CFG2[:1]), -0.15413362830353094852) < tol
    assert rel_err(orc.local_energy(hp, p, CFG2[:1]), 0.059924268309862600711) < 1e-13
    g, sho = c["gauss_sho"]["owf"], c["gauss_sho"]["oham"]
    assert rel_err(orc.wf_value(g, CFG2[:1]), 0.68386140921235585827) < tol
    assert rel_err(orc.wf_laplacian(g, CFG2[:1]), -3.0636991132713542451) < tol
    assert rel_err(orc.local_energy(sho, g, CFG2[:1]), 2.43) < tol
    assert rel_err(orc.wf_parameter_gradient(g, CFG2[:1]), [0.51973467100139045229]) < tol
    assert rel_err(orc.local_energy(c["gauss_h"]["oham"], g, CFG2[:1]), 0.61778578869237461836) < tol
    # custom_operator.rs:104,135-136: a = sqrt(2) is an exact eigenstate, E_L = 1.5 for any cfg
    g2 = orc.wf_desc(orc.WF_GAUSSIAN, [2 ** 0.5])
    for x in np.random.default_rng(0).normal(size=(20, 1, 3)):
        assert abs(orc.local_energy(sho, g2, x) - 1.5) < 2e-15


def test_survey_golden_diffusion_move(orc):
    """H2 diffuse move of electron 0, tau=0.25, given noise xi."""
    c = cases()
    h2 = c["h2"]["owf"]
    tau, xi = 0.25, np.array([0.11, -0.07, 0.05])
    psi, grad = orc.wf_value(h2, CFG2), orc.wf_gradient(h2, CFG2)
    xp = CFG2.copy()
    xp[0] = xp[0] + grad[0] / psi * tau + xi
    assert rel_err(xp[0], [0.4254216037217182184, -0.23761797250983788003, 0.46904493127459470007]) < 1e-14
    psin, gradn = orc.wf_value(h2, xp), orc.wf_gradient(h2, xp)
    assert rel_err(psin, 0.92058661795762400912) < 1e-14
    th = np.exp(-np.sum(((CFG2 - xp) - gradn / psin * tau) ** 2) / (2 * tau))
    tl = np.exp(-np.sum(((xp - CFG2) - grad / psi * tau) ** 2) / (2 * tau))
    assert rel_err(th, 0.9258321576472455333) < 1e-13 and rel_err(tl, 0.94687045442740148382) < 1e-13
    assert rel_err(th * psin ** 2 / (tl * psi ** 2), 0.99805752218436320147) < 1e-13


@pytest.mark.parametrize("name", sorted(GOLD.keys()))
def test_mpmath_golden(orc, name):
    g, wf, ham = _oracle_case(orc, name)
    sj = name.startswith(("sj", "lsj"))
    tol = 2e-10 if sj else 1e-12   # row-replacement determinants lose a few digits near nodes
    for e in g["entries"]:
        cfg = np.array(e["cfg"]).reshape(-1, 3)
        psi = float(e["psi"])
        assert rel_err(orc.wf_value(wf, cfg), psi) < tol
        gref = np.array([float(t) for t in e["grad"]]).reshape(-1, 3)
        assert np.max(np.abs(orc.wf_gradient(wf, cfg) - gref)) < tol * max(1.0, np.max(np.abs(gref)))
        assert abs(orc.wf_laplacian(wf, cfg) - float(e["lap"])) < tol * max(1.0, abs(float(e["lap"])))
        if g["n_params"]:
            pref = np.array([float(t) for t in e["pgrad"]])
            assert np.max(np.abs(orc.wf_parameter_gradient(wf, cfg) - pref)) < tol * max(1.0, np.max(np.abs(pref)))
        assert abs(orc.local_energy(ham, wf, cfg) - float(e["eloc"])) < tol * max(1.0, abs(float(e["eloc"])))


def test_lcao_reduces_to_the_reference_closed_forms(orc):
    # tests/helium_lcao.rs:94-101: the commented-out SpinDeterminantProduct of two 1s orbitals of width 1/1.69 IS the
    # HeliumAtomWaveFunction the test runs instead (:102) -> same psi, grad, lap, E_L as the STO-product kind
    he = orc.wf_desc(orc.WF_STO_PRODUCT, [1.69])
    lc = orc.wf_desc(orc.WF_LCAO_2E_1C, [1.0, 1.0], [0, 1.69, 0, 0, 0, 0, 0, 0])
    ham = orc.ham_desc(orc.HAM_ELECTRONIC, [[0, 0, 0]], [2])
    # one orbital on one of two centres = the 1-electron STO of examples/dmc.rs:106-125 shifted to that centre
    sto = orc.wf_desc(orc.WF_STO_1S, [0.8])
    l1 = orc.wf_desc(orc.WF_LCAO_1E_2C, [0.0, 1.0], [0, 0.8, 5.0, 5.0, 5.0, 0.25, -0.5, 0.125])
    for x in np.random.default_rng(3).normal(size=(10, 2, 3)):
        assert rel_err(orc.wf_value(lc, x), orc.wf_value(he, x)) < 1e-14
        assert rel_err(orc.wf_gradient(lc, x), orc.wf_gradient(he, x)) < 1e-13
        assert rel_err(orc.wf_laplacian(lc, x), orc.wf_laplacian(he, x)) < 1e-13
        assert rel_err(orc.local_energy(ham, lc, x), orc.local_energy(ham, he, x)) < 1e-13
        y = x[:1]
        ys = y - np.array([[0.25, -0.5, 0.125]])
        assert rel_err(orc.wf_value(l1, y), orc.wf_value(sto, ys)) < 1e-14
        assert rel_err(orc.wf_gradient(l1, y), orc.wf_gradient(sto, ys)) < 1e-13
        assert rel_err(orc.wf_laplacian(l1, y), orc.wf_laplacian(sto, ys)) < 1e-13
    # the triplet determinant is antisymmetric under exchange and vanishes at coincidence
    g = GOLD["lcao_h2_triplet"]
    tr = orc.wf_desc(orc.WF_LCAO_2E_2C, g["params"], g["geom"])
    x = np.array([[0.3, -0.2, 0.5], [-0.6, 0.1, 0.25]])
    assert rel_err(orc.wf_value(tr, x[::-1]), -orc.wf_value(tr, x)) < 1e-14
    assert orc.wf_value(tr, np.array([x[0], x[0]])) == 0.0


def test_lcao_h2plus_vmc_reproduces_the_closed_form_energy(orc):
    # tests/hydrogen_molecular_ion_lcao.rs:101-140 with the LCAO function it names: MetropolisBox(1.0), block 100; the
    # energy of 1s_A + 1s_B at R = 2.5 follows from the overlap, Coulomb and exchange integrals (-0.56483; the test's -0.565)
    from common import SEED0, cases
    c = cases()["lcao_h2p"]
    W = 256
    cfgs = np.array([orc.init_uniform(SEED0, w, 1) for w in range(W)])
    r = orc.ensemble_run(c["owf"], c["oham"], orc.run_options(orc.METROP_BOX, 1.0, orc.OBS_ENERGY), cfgs, SEED0, 4000, 100,
                         trace=False)["energy"]
    R = 2.5
    S, J, K = np.exp(-R) * (1 + R + R * R / 3), -1 / R + np.exp(-2 * R) * (1 + 1 / R), -np.exp(-R) * (1 + R)
    exact = -0.5 + (J + K) / (1 + S) + 1 / R
    err = r.mean(axis=1).std() / np.sqrt(W)
    assert abs(r.mean() - exact) < 5 * err and err < 2e-3


def test_analytic_checks(orc):
    # 1-electron STO alpha=1 on hydrogen: E_L == -0.5 everywhere
    sto = orc.wf_desc(orc.WF_STO_1S, [1.0])
    ham = orc.ham_desc(orc.HAM_ELECTRONIC, [[0, 0, 0]], [1])
    for x in np.random.default_rng(1).normal(size=(10, 1, 3)):
        assert abs(orc.local_energy(ham, sto, x) + 0.5) < 1e-14
    # WaveFunctionMock: laplacian is 1.0, gradient unimplemented (metrop.rs:248-254)
    mock = orc.wf_desc(orc.WF_CONSTANT, [], [1.0])
    assert orc.wf_value(mock, [[0.1, 0.2, 0.3]]) == 1.0 and orc.wf_laplacian(mock, [[0.1, 0.2, 0.3]]) == 1.0
    with pytest.raises(RuntimeError):
        orc.wf_gradient(mock, [[0.1, 0.2, 0.3]])
