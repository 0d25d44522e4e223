#!/usr/bin/env python
"""Generates tests/golden/wf_golden.json: 40-digit evaluations of the reference's trial
wavefunctions.  psi is restated from the reference's closed forms (file:line below); every
derivative (grad psi, lap psi, d psi/d p) is obtained by mpmath numerical differentiation of psi,
so the fixtures do not share any derivative formula with the oracle or the CUDA kernels.

  H2      examples/hydrogen_molecule.rs:91-100
  He      examples/helium_atom_singlet.rs:63-70
  H2+     tests/hydrogen_molecular_ion_lcao.rs:69-73
  Gauss   examples/dmc.rs:47-50
  STO     examples/dmc.rs:108-111
  SJ      SURVEY.md §8(c) synthetic config 5; Jastrow f_ee of theory/jastrow.tex:23-26
  LSJ     the same LCAO API in general (N_c <= 8 centres, n_up, n_dn <= 5, orbitals shared by the spins) times the
          Jastrow of theory/jastrow.tex: geom = [kappa, n_up, n_dn, N_c, -, -, -, -, (R_c, alpha_c)...]
  LCAO    Hydrogen1sBasis / Orbital / SingleDeterminant / SpinDeterminantProduct as named (commented out) at
          tests/helium_lcao.rs:94-101 and tests/hydrogen_molecular_ion_lcao.rs:103-107: phi_k = sum_c C[k][c]
          exp(-|r - R_c| / width); geom = [mode, 1/width, R_0, R_1], params = C[k][c]
Potentials: src/operator/src/operator.rs:25-36,80-90; SHO examples/custom_operator.rs:56-58.

Run from the repo root:  python tests/golden/make_golden.py            (everything; the SJ cases take minutes)
                         python tests/golden/make_golden.py --only lcao  (keeps the other entries of the JSON)
"""
import json
import os
import random

import mpmath as mp

mp.mp.dps = 40


def norm(v):
    return mp.sqrt(sum(x * x for x in v))


def sto(alpha, v):
    return mp.exp(-alpha * norm(v))


def psi_h2(x, p, g):
    h = g[0] / 2
    x1, x2 = x[0:3], x[3:6]
    a1 = [x1[0] - h, x1[1], x1[2]]; b1 = [x1[0] + h, x1[1], x1[2]]
    a2 = [x2[0] - h, x2[1], x2[2]]; b2 = [x2[0] + h, x2[1], x2[2]]
    return sto(p[0], a1) * sto(p[0], b2) + sto(p[0], b1) * sto(p[0], a2)


def psi_he(x, p, g):
    return mp.exp(-p[0] * (norm(x[0:3]) + norm(x[3:6])))


def psi_h2p(x, p, g):
    h = g[0] / 2
    return sto(p[0], [x[0] - h, x[1], x[2]]) * sto(p[0], [x[0] + h, x[1], x[2]])


def psi_gauss(x, p, g):
    return mp.exp(-(norm(x) / p[0]) ** 2)


def psi_sto(x, p, g):
    return mp.exp(-p[0] * norm(x))


def psi_lcao(ne, nc):
    def psi(x, p, g):
        mode, alpha = g[0], g[1]

        def phi(k, r):
            return sum(p[k * nc + c] * sto(alpha, [r[q] - g[2 + 3 * c + q] for q in range(3)]) for c in range(nc))
        if ne == 1:
            return phi(0, x[0:3])
        direct = phi(0, x[0:3]) * phi(1, x[3:6])
        return direct - phi(0, x[3:6]) * phi(1, x[0:3]) if mode else direct
    return psi


def det(m):
    return mp.det(mp.matrix(m)) if len(m) else mp.mpf(1)


def psi_sj(x, p, g):
    kappa, nup, ndn = g[0], int(g[1]), int(g[2])
    z1, z2, z3 = p[0], p[1], p[2]
    b = p[3:7]

    def orbs(r3, n):
        r = norm(r3)
        full = [mp.exp(-z1 * r), r * mp.exp(-z2 * r), r3[0] * mp.exp(-z3 * r), r3[1] * mp.exp(-z3 * r),
                r3[2] * mp.exp(-z3 * r)]
        return full[:n]

    up = [orbs(x[3 * i:3 * i + 3], nup) for i in range(nup)]
    dn = [orbs(x[3 * (nup + i):3 * (nup + i) + 3], ndn) for i in range(ndn)]
    f = mp.mpf(0)
    ne = nup + ndn
    for i in range(ne):
        for j in range(i + 1, ne):
            r = norm([x[3 * i + c] - x[3 * j + c] for c in range(3)])
            R = (1 - mp.exp(-kappa * r)) / kappa
            f += b[0] * R / (1 + b[1] * R) + b[2] * R ** 2 + b[3] * R ** 3
    return det(up) * det(dn) * mp.exp(f)


def psi_lsj(x, p, g):
    """General LCAO Slater-Jastrow (MOLE_WF_LCAO_SJ): SpinDeterminantProduct of n_orb = max(n_up, n_dn) orbitals
    phi_k = sum_c C[k][c] exp(-alpha_c |r - R_c|) shared by both spins, times exp(f_ee) of theory/jastrow.tex:23-26.
    geom = [kappa, n_up, n_dn, N_c, 0, 0, 0, 0, (R_c, alpha_c) ...]; params = C[k][c] at k N_c + c, then b1..b4."""
    kappa, nup, ndn, nc = g[0], int(g[1]), int(g[2]), int(g[3])
    norb = max(nup, ndn)
    b = p[norb * nc:norb * nc + 4]

    def phi(k, r):
        return sum(p[k * nc + c] * sto(g[8 + 4 * c + 3], [r[q] - g[8 + 4 * c + q] for q in range(3)]) for c in range(nc))

    up = [[phi(k, x[3 * i:3 * i + 3]) for k in range(nup)] for i in range(nup)]
    dn = [[phi(k, x[3 * (nup + i):3 * (nup + i) + 3]) for k in range(ndn)] for i in range(ndn)]
    f = mp.mpf(0)
    ne = nup + ndn
    for i in range(ne):
        for j in range(i + 1, ne):
            r = norm([x[3 * i + c] - x[3 * j + c] for c in range(3)])
            R = (1 - mp.exp(-kappa * r)) / kappa
            f += b[0] * R / (1 + b[1] * R) + b[2] * R ** 2 + b[3] * R ** 3
    return det(up) * det(dn) * mp.exp(f)


def lsj_case(nup, ndn, pos, alphas, C, b, kappa=1.0):
    geom = [kappa, nup, ndn, len(pos), 0, 0, 0, 0]
    for c, R in enumerate(pos):
        geom += list(R) + [alphas[c]]
    params = [v for row in C for v in row] + list(b)
    return (psi_lsj, params, geom, nup + ndn, len(params), ("electronic", [list(R) for R in pos], [1] * len(pos)))


H4_POS = [[-2.1, 0, 0], [-0.7, 0, 0], [0.7, 0, 0], [2.1, 0, 0]]
H8_POS = [[1.4 * (i - 3.5), 0.3 * ((-1) ** i), 0.1 * i] for i in range(8)]
H8_C = [[1.0, 1.1, 1.2, 1.3, 1.3, 1.2, 1.1, 1.0],
        [1.0, 0.8, 0.5, 0.2, -0.2, -0.5, -0.8, -1.0],
        [1.0, 0.3, -0.6, -0.9, -0.9, -0.6, 0.3, 1.0],
        [0.7, -0.5, -0.9, 0.4, -0.4, 0.9, 0.5, -0.7]]


def v_ion(x, ions, z):
    ne = len(x) // 3
    pot = mp.mpf(0)
    for I, R in enumerate(ions):
        for j in range(ne):
            pot -= z[I] / norm([x[3 * j + c] - R[c] for c in range(3)])
    for I in range(len(ions)):
        for J in range(I + 1, len(ions)):
            pot += z[I] * z[J] / norm([ions[J][c] - ions[I][c] for c in range(3)])
    return pot


def v_ee(x):
    ne = len(x) // 3
    pot = mp.mpf(0)
    for i in range(ne):
        for j in range(i + 1, ne):
            pot += 1 / norm([x[3 * i + c] - x[3 * j + c] for c in range(3)])
    return pot


def derivs(psi, x, p, g):
    n = len(x)
    x = [mp.mpf(v) for v in x]
    p = [mp.mpf(v) for v in p]
    val = psi(x, p, g)
    grad, lap = [], mp.mpf(0)
    for i in range(n):
        def f(t, i=i):
            y = list(x); y[i] = t
            return psi(y, p, g)
        grad.append(mp.diff(f, x[i]))
        lap += mp.diff(f, x[i], 2)
    pg = []
    for k in range(len(p)):
        def f(t, k=k):
            q = list(p); q[k] = t
            return psi(x, q, g)
        pg.append(mp.diff(f, p[k]))
    return val, grad, lap, pg


CASES = {
    # name: (psi, params, geom, n_elec, n_opt_params, ham)
    "h2": (psi_h2, [0.5], [1.4], 2, 1, ("electronic", [[-0.7, 0, 0], [0.7, 0, 0]], [1, 1])),
    "he": (psi_he, [1.69], [], 2, 1, ("electronic", [[0, 0, 0]], [2])),
    "h2p": (psi_h2p, [1.0], [2.5], 1, 0, ("electronic", [[-1.25, 0, 0], [1.25, 0, 0]], [1, 1])),
    "gauss_sho": (psi_gauss, [1.0], [], 1, 1, ("harmonic", 1.0)),
    "gauss_h": (psi_gauss, [1.0], [], 1, 1, ("electronic", [[0, 0, 0]], [1])),
    "sto_h": (psi_sto, [0.8], [], 1, 1, ("electronic", [[0, 0, 0]], [1])),
    "sj_ne": (psi_sj, [9.64, 2.88, 2.88, 0.5, 1.0, 0.1, -0.05], [1.0, 5, 5], 10, 7, ("electronic", [[0, 0, 0]], [10])),
    "sj_be": (psi_sj, [3.68, 0.96, 0.96, 0.5, 1.0, 0.2, 0.1], [1.0, 2, 2], 4, 7, ("electronic", [[0, 0, 0]], [4])),
    "sj_li": (psi_sj, [2.69, 0.64, 0.64, 0.4, 0.8, 0.0, 0.0], [1.5, 2, 1], 3, 7, ("electronic", [[0, 0, 0]], [3])),
    # H2+ LCAO of tests/hydrogen_molecular_ion_lcao.rs:101-107 (ion_pos +-1.25, width 1, coefficients [[1], [1]])
    "lcao_h2p": (psi_lcao(1, 2), [1.0, 1.0], [0, 1.0, -1.25, 0, 0, 1.25, 0, 0], 1, 2,
                 ("electronic", [[-1.25, 0, 0], [1.25, 0, 0]], [1, 1])),
    # He LCAO of tests/helium_lcao.rs:92-101 (width 1/1.69, two orbitals [[1]], n_up = 1)
    "lcao_he": (psi_lcao(2, 1), [1.0, 1.0], [0, 1.69, 0, 0, 0, 0, 0, 0], 2, 2, ("electronic", [[0, 0, 0]], [2])),
    # H2 molecular orbitals: singlet sigma_g(1) sigma_g'(2) with unequal coefficients, and the triplet determinant
    "lcao_h2_singlet": (psi_lcao(2, 2), [1.0, 0.9, 0.8, 1.1], [0, 1.0 / 0.85, -0.7, 0, 0, 0.7, 0, 0], 2, 4,
                        ("electronic", [[-0.7, 0, 0], [0.7, 0, 0]], [1, 1])),
    "lcao_h2_triplet": (psi_lcao(2, 2), [1.0, 1.0, 1.0, -1.0], [1, 1.0 / 0.85, -0.7, 0.1, 0, 0.7, -0.1, 0.2], 2, 4,
                        ("electronic", [[-0.7, 0.1, 0], [0.7, -0.1, 0.2]], [1, 1])),
    # general LCAO Slater-Jastrow: H4 chain (2 up, 2 dn, 4 centres, P = 12), an open-shell H3 (2 up, 1 dn, widths differ,
    # P = 10) and a zig-zag H8 chain (4 up, 4 dn, 8 centres, P = 36: the large-P case of SURVEY.md 8(f)3)
    "lsj_h4": lsj_case(2, 2, H4_POS, [1.0, 1.1, 1.1, 1.0], [[1, 1, 1, 1], [1, 0.5, -0.5, -1]], [0.5, 1.0, 0.1, -0.05]),
    "lsj_h3": lsj_case(2, 1, [[-1.0, 0, 0], [0.6, 0.8, 0], [0.5, -0.7, 0.4]], [1.2, 0.9, 1.0], [[1, 0.9, 0.8], [1, -0.4, -0.7]],
                       [0.4, 0.8, 0.0, 0.0], kappa=1.5),
    "lsj_h8": lsj_case(4, 4, H8_POS, [1.0] * 8, H8_C, [0.5, 1.0, 0.05, 0.02]),
}


def main():
    import sys
    only = sys.argv[2] if len(sys.argv) > 2 and sys.argv[1] == "--only" else None
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "wf_golden.json")
    rng_main = random.Random(20261017)
    out = json.load(open(path)) if only else {}
    for name, (psi, p, g, ne, nopt, ham) in CASES.items():
        if only and not name.startswith(only):
            continue
        # the LCAO cases were added later and draw from their own stream so that the earlier entries stay reproducible
        rng = random.Random("mole-b200 " + name) if name.startswith(("lcao", "lsj")) else rng_main
        entries = []
        ncfg = 3 if name.startswith("sj") else (2 if name == "lsj_h8" else (3 if name.startswith("lsj") else 6))
        for c in range(ncfg):
            if c == 0 and ne <= 2:
                x = [0.3, -0.2, 0.5, -0.6, 0.1, 0.25][:3 * ne]      # SURVEY.md §8(c) configuration
            else:
                scale = 0.6 if name.startswith("sj") else (1.5 if name.startswith("lsj") else 1.0)
                x = [rng.gauss(0.0, scale) for _ in range(3 * ne)]
            val, grad, lap, pg = derivs(psi, x, p, [mp.mpf(v) for v in g])
            xm = [mp.mpf(v) for v in x]
            if ham[0] == "electronic":
                v = v_ion(xm, ham[1], ham[2]) + v_ee(xm)
            else:
                v = mp.mpf(ham[1]) ** 2 * sum(t * t for t in xm) / 2
            eloc = -lap / (2 * val) + v
            entries.append(dict(cfg=x, psi=mp.nstr(val, 25), grad=[mp.nstr(t, 25) for t in grad], lap=mp.nstr(lap, 25),
                                pgrad=[mp.nstr(t, 25) for t in pg[:nopt]], eloc=mp.nstr(eloc, 25), v=mp.nstr(v, 25)))
        out[name] = dict(params=p, geom=g, n_elec=ne, n_params=nopt, ham=list(ham), entries=entries)
        print(name, "done")
    with open(path, "w") as fh:
        json.dump(out, fh, indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()
