// Device-side trial-wavefunction descriptors for the closed-form kinds with N_e <= 2
// (thread-per-walker, walker state and cached intermediates live in registers).
//
// Each kind mirrors one user impl of Function / Differentiate / Optimize in the reference's
// examples/ and tests/ (file:line next to each struct).  Unlike the reference, which recomputes
// every norm and exponential on every call, a kind keeps the per-electron
// radial quantities (r, 1/r, exp(-alpha r)) in a State so that a single-electron move re-evaluates
// only what changed and no formula needs a division.  Square roots, reciprocals and exponentials go
// through the branch-free batched math of mole_math.cuh (<= 2 ulp, far inside the 1e-10 parity bar).
// value/gradient/laplacian keep the reference's meaning: UN-normalised psi, grad psi, sum_i lap_i psi.
#pragma once
#include "mole_internal.h"
#include "mole_math.cuh"

enum { K_STO_1S = 0, K_GAUSSIAN = 1, K_STO_PRODUCT = 2, K_H2_HL_STO = 3, K_H2P_PRODUCT = 4, K_SLATER_JASTROW = 5,
       K_CONSTANT = 6, K_LCAO_1E_2C = 7, K_LCAO_2E_1C = 8, K_LCAO_2E_2C = 9, K_LCAO_SJ = 10 };

template <int KIND> struct WfDev;

MOLE_D double mole_norm2_3(double a, double b, double c) { return fma(c, c, fma(b, b, a * a)); }

// ---- 1-electron STO, psi = exp(-alpha |x|)                 examples/dmc.rs:106-143
template <> struct WfDev<K_STO_1S> {
  static constexpr int NE = 1, NP = 1;
  struct State { double x[3]; double r, ir, e; };
  MOLE_D static void init(const WfParams& p, State& s) {
    s.r = m_sqrt_rsqrt(mole_norm2_3(s.x[0], s.x[1], s.x[2]), s.ir);
    s.e = m_exp(-p.p[0] * s.r);
  }
  MOLE_D static void move(const WfParams& p, State& s, int, const double xn[3]) {
    s.x[0] = xn[0]; s.x[1] = xn[1]; s.x[2] = xn[2];
    init(p, s);
  }
  MOLE_D static double psi(const WfParams&, const State& s) { return s.e; }
  MOLE_D static void grad(const WfParams& p, const State& s, double* g) {
    const double c = -p.p[0] * s.e * s.ir;                    // dmc.rs:116-119
    g[0] = c * s.x[0]; g[1] = c * s.x[1]; g[2] = c * s.x[2];
  }
  MOLE_D static double lap(const WfParams& p, const State& s) {
    return p.p[0] * s.e * s.ir * (p.p[0] * s.r - 2.0);        // dmc.rs:121-124
  }
  MOLE_D static void pgrad(const WfParams&, const State& s, double* o) { o[0] = -s.r * s.e; }  // dmc.rs:128-130
};

// ---- 1-electron Gaussian, psi = exp(-(|x|/a)^2)            examples/dmc.rs:44-92, tests/sho_optimize.rs:63-111
template <> struct WfDev<K_GAUSSIAN> {
  static constexpr int NE = 1, NP = 1;
  struct State { double x[3]; double n2, e; };
  MOLE_D static void init(const WfParams& p, State& s) {
    s.n2 = mole_norm2_3(s.x[0], s.x[1], s.x[2]);
    s.e = m_exp(-s.n2 * m_rcp(p.p[0] * p.p[0]));
  }
  MOLE_D static void move(const WfParams& p, State& s, int, const double xn[3]) {
    s.x[0] = xn[0]; s.x[1] = xn[1]; s.x[2] = xn[2];
    init(p, s);
  }
  MOLE_D static double psi(const WfParams&, const State& s) { return s.e; }
  MOLE_D static void grad(const WfParams& p, const State& s, double* g) {
    const double c = -2.0 * s.e * m_rcp(p.p[0] * p.p[0]);     // dmc.rs:56-59
    g[0] = c * s.x[0]; g[1] = c * s.x[1]; g[2] = c * s.x[2];
  }
  MOLE_D static double lap(const WfParams& p, const State& s) {
    const double a2 = p.p[0] * p.p[0];
    return s.e * (4.0 * s.n2 - 6.0 * a2) * m_rcp(a2 * a2);    // dmc.rs:61-64
  }
  MOLE_D static void pgrad(const WfParams& p, const State& s, double* o) {
    o[0] = s.e * 2.0 * s.n2 * m_rcp(p.p[0] * p.p[0] * p.p[0]);   // dmc.rs:74-79
  }
};

// ---- He singlet, psi = exp(-alpha (r1 + r2))               examples/helium_atom_singlet.rs:60-118
template <> struct WfDev<K_STO_PRODUCT> {
  static constexpr int NE = 2, NP = 1;
  struct State { double x[6]; double r[2], ir[2], e; };
  MOLE_D static void init(const WfParams& p, State& s) {
    const double n2[2] = {mole_norm2_3(s.x[0], s.x[1], s.x[2]), mole_norm2_3(s.x[3], s.x[4], s.x[5])};
    m_sqrt_rsqrt_n<2>(n2, s.r, s.ir);
    s.e = m_exp(-p.p[0] * (s.r[0] + s.r[1]));
  }
  MOLE_D static void move(const WfParams& p, State& s, int e, const double xn[3]) {
    s.x[3 * e] = xn[0]; s.x[3 * e + 1] = xn[1]; s.x[3 * e + 2] = xn[2];
    s.r[e] = m_sqrt_rsqrt(mole_norm2_3(xn[0], xn[1], xn[2]), s.ir[e]);
    s.e = m_exp(-p.p[0] * (s.r[0] + s.r[1]));
  }
  MOLE_D static double psi(const WfParams&, const State& s) { return s.e; }
  MOLE_D static void grad(const WfParams& p, const State& s, double* g) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const double c = -p.p[0] * s.ir[e] * s.e;               // helium_atom_singlet.rs:80-81
      g[3 * e] = c * s.x[3 * e]; g[3 * e + 1] = c * s.x[3 * e + 1]; g[3 * e + 2] = c * s.x[3 * e + 2];
    }
  }
  MOLE_D static double lap(const WfParams& p, const State& s) {
    const double al = p.p[0];                                 // helium_atom_singlet.rs:96-97
    return al * s.ir[0] * s.e * (al * s.r[0] - 2.0) + al * s.ir[1] * s.e * (al * s.r[1] - 2.0);
  }
  MOLE_D static void pgrad(const WfParams&, const State& s, double* o) { o[0] = -s.e * (s.r[0] + s.r[1]); }  // :102-105
};

// ---- H2 Heitler-London, psi = phi(x1-R/2)phi(x2+R/2) + phi(x1+R/2)phi(x2-R/2)   examples/hydrogen_molecule.rs:88-168
template <> struct WfDev<K_H2_HL_STO> {
  static constexpr int NE = 2, NP = 1;
  // per electron e: r[e][0] = |x_e - R/2|, r[e][1] = |x_e + R/2|, their reciprocals and the STO values
  struct State { double x[6]; double r[2][2], ir[2][2], ex[2][2]; };
  MOLE_D static void one(const WfParams& p, State& s, int e) {
    const double h = 0.5 * p.geom[0];
    const double yz = fma(s.x[3 * e + 2], s.x[3 * e + 2], s.x[3 * e + 1] * s.x[3 * e + 1]);
    const double xa = s.x[3 * e] - h, xb = s.x[3 * e] + h;
    const double n2[2] = {fma(xa, xa, yz), fma(xb, xb, yz)};
    m_sqrt_rsqrt_n<2>(n2, s.r[e], s.ir[e]);
    const double a[2] = {-p.p[0] * s.r[e][0], -p.p[0] * s.r[e][1]};
    m_exp_n<2, true>(a, s.ex[e]);
  }
  MOLE_D static void init(const WfParams& p, State& s) { one(p, s, 0); one(p, s, 1); }
  MOLE_D static void move(const WfParams& p, State& s, int e, const double xn[3]) {
    s.x[3 * e] = xn[0]; s.x[3 * e + 1] = xn[1]; s.x[3 * e + 2] = xn[2];
    one(p, s, e);
  }
  MOLE_D static double psi(const WfParams&, const State& s) {
    return fma(s.ex[0][0], s.ex[1][1], s.ex[0][1] * s.ex[1][0]);   // hydrogen_molecule.rs:96-99
  }
  MOLE_D static void grad(const WfParams& p, const State& s, double* g) {
    const double h = 0.5 * p.geom[0];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int o = 1 - e;
      // grad_e psi = phi(x_o + R/2) grad phi(x_e - R/2) + phi(x_o - R/2) grad phi(x_e + R/2)   (:114-119)
      const double ca = -p.p[0] * s.ex[e][0] * s.ir[e][0] * s.ex[o][1];
      const double cb = -p.p[0] * s.ex[e][1] * s.ir[e][1] * s.ex[o][0];
      g[3 * e] = fma(ca, s.x[3 * e] - h, cb * (s.x[3 * e] + h));
      g[3 * e + 1] = (ca + cb) * s.x[3 * e + 1];
      g[3 * e + 2] = (ca + cb) * s.x[3 * e + 2];
    }
  }
  MOLE_D static double sto_lap(double al, double r, double ir, double e) { return al * e * ir * (al * r - 2.0); }  // :51-53
  MOLE_D static double lap(const WfParams& p, const State& s) {
    const double al = p.p[0];                                  // :129-133
    return (s.ex[1][1] * sto_lap(al, s.r[0][0], s.ir[0][0], s.ex[0][0]) + s.ex[1][0] * sto_lap(al, s.r[0][1], s.ir[0][1], s.ex[0][1])) +
           (s.ex[0][1] * sto_lap(al, s.r[1][0], s.ir[1][0], s.ex[1][0]) + s.ex[0][0] * sto_lap(al, s.r[1][1], s.ir[1][1], s.ex[1][1]));
  }
  MOLE_D static void pgrad(const WfParams&, const State& s, double* o) {
    // d/dalpha of each STO is -r e (:55-57); product rule over the two terms (:144-153)
    o[0] = s.ex[0][0] * (-s.r[1][1] * s.ex[1][1]) + (-s.r[0][0] * s.ex[0][0]) * s.ex[1][1] +
           s.ex[0][1] * (-s.r[1][0] * s.ex[1][0]) + (-s.r[0][1] * s.ex[0][1]) * s.ex[1][0];
  }
};

// ---- H2+ product ansatz, psi = phi(x - R/2) phi(x + R/2)   tests/hydrogen_molecular_ion_lcao.rs:66-92
template <> struct WfDev<K_H2P_PRODUCT> {
  static constexpr int NE = 1, NP = 0;
  struct State { double x[3]; double r[2], ir[2], ex[2]; };
  MOLE_D static void init(const WfParams& p, State& s) {
    const double h = 0.5 * p.geom[0];
    const double yz = fma(s.x[2], s.x[2], s.x[1] * s.x[1]);
    const double xa = s.x[0] - h, xb = s.x[0] + h;
    const double n2[2] = {fma(xa, xa, yz), fma(xb, xb, yz)};
    m_sqrt_rsqrt_n<2>(n2, s.r, s.ir);
    const double a[2] = {-p.p[0] * s.r[0], -p.p[0] * s.r[1]};
    m_exp_n<2, true>(a, s.ex);
  }
  MOLE_D static void move(const WfParams& p, State& s, int, const double xn[3]) {
    s.x[0] = xn[0]; s.x[1] = xn[1]; s.x[2] = xn[2];
    init(p, s);
  }
  MOLE_D static double psi(const WfParams&, const State& s) { return s.ex[0] * s.ex[1]; }
  MOLE_D static void grad(const WfParams& p, const State& s, double* g) {
    const double h = 0.5 * p.geom[0];
    const double c1 = -p.p[0] * s.ex[0] * s.ir[0] * s.ex[1];   // phi(r2) grad phi(r1)
    const double c2 = -p.p[0] * s.ex[1] * s.ir[1] * s.ex[0];   // phi(r1) grad phi(r2)   (:82)
    g[0] = fma(c1, s.x[0] - h, c2 * (s.x[0] + h));
    g[1] = (c1 + c2) * s.x[1];
    g[2] = (c1 + c2) * s.x[2];
  }
  MOLE_D static double lap(const WfParams& p, const State& s) {
    const double al = p.p[0], h = 0.5 * p.geom[0];
    const double g1 = -al * s.ex[0] * s.ir[0], g2 = -al * s.ex[1] * s.ir[1];
    const double l1 = -g1 * (al * s.r[0] - 2.0), l2 = -g2 * (al * s.r[1] - 2.0);
    const double xa = s.x[0] - h, xb = s.x[0] + h;
    const double dot = (g1 * xa) * (g2 * xb) + (g1 * s.x[1]) * (g2 * s.x[1]) + (g1 * s.x[2]) * (g2 * s.x[2]);
    return s.ex[0] * l2 + s.ex[1] * l1 + 2.0 * dot;           // :88-90
  }
  MOLE_D static void pgrad(const WfParams&, const State&, double*) {}
};

// ---- LCAO determinants over a hydrogen-1s basis: the API the reference's tests name but keep commented out
// (Hydrogen1sBasis / Orbital / SingleDeterminant / SpinDeterminantProduct, tests/helium_lcao.rs:94-101,
// tests/hydrogen_molecular_ion_lcao.rs:103-107).
//   chi_c(r) = exp(-alpha |r - R_c|), alpha = 1 / width;   phi_k(r) = sum_c C[k][c] chi_c(r)
//   NE = 1: psi = phi_0(x_0)                                            SingleDeterminant([phi_0])
//   NE = 2, mode 0: psi = phi_0(x_0) phi_1(x_1)                         SpinDeterminantProduct([phi_0, phi_1], n_up = 1)
//   NE = 2, mode 1: psi = phi_0(x_0) phi_1(x_1) - phi_0(x_1) phi_1(x_0) SingleDeterminant([phi_0, phi_1])
// geom: [0] mode, [1] alpha, [2 + 3c ..] centre R_c;  params: C[k][c] at k * NC + c (P = NK * NC, all variational).
// The State caches r, 1/r and chi per (electron, centre): a single-electron move re-evaluates NC exponentials.
template <int NE_, int NC>
struct LcaoDev {
  static constexpr int NE = NE_, NK = NE_, NP = NK * NC;
  struct State { double x[3 * NE_]; double r[NE_][NC], ir[NE_][NC], ch[NE_][NC]; };
  MOLE_D static void one(const WfParams& p, State& s, int e) {
    double n2[NC], a[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c)
      n2[c] = mole_norm2_3(s.x[3 * e] - p.geom[2 + 3 * c], s.x[3 * e + 1] - p.geom[3 + 3 * c], s.x[3 * e + 2] - p.geom[4 + 3 * c]);
    m_sqrt_rsqrt_n<NC>(n2, s.r[e], s.ir[e]);
#pragma unroll
    for (int c = 0; c < NC; ++c) a[c] = -p.geom[1] * s.r[e][c];
    m_exp_n<NC, true>(a, s.ch[e]);
  }
  MOLE_D static void init(const WfParams& p, State& s) {
#pragma unroll
    for (int e = 0; e < NE; ++e) one(p, s, e);
  }
  MOLE_D static void move(const WfParams& p, State& s, int e, const double xn[3]) {
    s.x[3 * e] = xn[0]; s.x[3 * e + 1] = xn[1]; s.x[3 * e + 2] = xn[2];
    one(p, s, e);
  }
  // orbital k at electron e: value, gradient, laplacian
  MOLE_D static double phi(const WfParams& p, const State& s, int k, int e) {
    double v = 0.0;
#pragma unroll
    for (int c = 0; c < NC; ++c) v = fma(p.p[k * NC + c], s.ch[e][c], v);
    return v;
  }
  MOLE_D static void gphi(const WfParams& p, const State& s, int k, int e, double* g) {
    g[0] = g[1] = g[2] = 0.0;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const double f = -p.geom[1] * p.p[k * NC + c] * s.ch[e][c] * s.ir[e][c];   // grad chi = -alpha chi (r - R_c) / |r - R_c|
      g[0] = fma(f, s.x[3 * e] - p.geom[2 + 3 * c], g[0]);
      g[1] = fma(f, s.x[3 * e + 1] - p.geom[3 + 3 * c], g[1]);
      g[2] = fma(f, s.x[3 * e + 2] - p.geom[4 + 3 * c], g[2]);
    }
  }
  MOLE_D static double lphi(const WfParams& p, const State& s, int k, int e) {
    double v = 0.0;
#pragma unroll
    for (int c = 0; c < NC; ++c)                                                // lap chi = alpha chi (alpha r - 2) / r
      v = fma(p.p[k * NC + c] * (p.geom[1] * s.ch[e][c] * s.ir[e][c]), fma(p.geom[1], s.r[e][c], -2.0), v);
    return v;
  }
  MOLE_D static double psi(const WfParams& p, const State& s) {
    if (NE == 1) return phi(p, s, 0, 0);
    const double m = p.geom[0] != 0.0 ? 1.0 : 0.0;
    return phi(p, s, 0, 0) * phi(p, s, NK - 1, NE - 1) - m * (phi(p, s, 0, NE - 1) * phi(p, s, NK - 1, 0));
  }
  MOLE_D static void grad(const WfParams& p, const State& s, double* g) {
    if (NE == 1) { gphi(p, s, 0, 0, g); return; }
    const double m = p.geom[0] != 0.0 ? 1.0 : 0.0;
    double a[3], b[3];
    // electron 0: grad phi_0(x_0) phi_1(x_1) - m phi_0(x_1) grad phi_1(x_0)
    gphi(p, s, 0, 0, a); gphi(p, s, NK - 1, 0, b);
    const double p11 = phi(p, s, NK - 1, NE - 1), p01 = phi(p, s, 0, NE - 1);
#pragma unroll
    for (int q = 0; q < 3; ++q) g[q] = a[q] * p11 - m * (p01 * b[q]);
    // electron 1: phi_0(x_0) grad phi_1(x_1) - m grad phi_0(x_1) phi_1(x_0)
    gphi(p, s, NK - 1, NE - 1, a); gphi(p, s, 0, NE - 1, b);
    const double p00 = phi(p, s, 0, 0), p10 = phi(p, s, NK - 1, 0);
#pragma unroll
    for (int q = 0; q < 3; ++q) g[3 * (NE - 1) + q] = p00 * a[q] - m * (b[q] * p10);
  }
  MOLE_D static double lap(const WfParams& p, const State& s) {
    if (NE == 1) return lphi(p, s, 0, 0);
    const double m = p.geom[0] != 0.0 ? 1.0 : 0.0;
    const double p00 = phi(p, s, 0, 0), p11 = phi(p, s, NK - 1, NE - 1), p01 = phi(p, s, 0, NE - 1), p10 = phi(p, s, NK - 1, 0);
    const double direct = lphi(p, s, 0, 0) * p11 + p00 * lphi(p, s, NK - 1, NE - 1);
    const double exch = p01 * lphi(p, s, NK - 1, 0) + lphi(p, s, 0, NE - 1) * p10;
    return direct - m * exch;
  }
  MOLE_D static void pgrad(const WfParams& p, const State& s, double* o) {
    if (NE == 1) {
#pragma unroll
      for (int c = 0; c < NC; ++c) o[c] = s.ch[0][c];
      return;
    }
    const double m = p.geom[0] != 0.0 ? 1.0 : 0.0;
    const double p00 = phi(p, s, 0, 0), p11 = phi(p, s, NK - 1, NE - 1), p01 = phi(p, s, 0, NE - 1), p10 = phi(p, s, NK - 1, 0);
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      o[c] = s.ch[0][c] * p11 - m * (s.ch[NE - 1][c] * p10);                    // d / d C[0][c]
      o[(NK - 1) * NC + c] = p00 * s.ch[NE - 1][c] - m * (p01 * s.ch[0][c]);    // d / d C[1][c]
    }
  }
};
template <> struct WfDev<K_LCAO_1E_2C> : LcaoDev<1, 2> {};   // H2+ LCAO, tests/hydrogen_molecular_ion_lcao.rs:103-107
template <> struct WfDev<K_LCAO_2E_1C> : LcaoDev<2, 1> {};   // He LCAO, tests/helium_lcao.rs:94-101
template <> struct WfDev<K_LCAO_2E_2C> : LcaoDev<2, 2> {};   // H2 molecular orbitals (singlet product or triplet determinant)

// ---- WaveFunctionMock, psi = const                         src/metropolis/src/metrop.rs:225-255
template <> struct WfDev<K_CONSTANT> {
  static constexpr int NE = 1, NP = 0;
  struct State { double x[3]; };
  MOLE_D static void init(const WfParams&, State&) {}
  MOLE_D static void move(const WfParams&, State& s, int, const double xn[3]) { s.x[0] = xn[0]; s.x[1] = xn[1]; s.x[2] = xn[2]; }
  MOLE_D static double psi(const WfParams& p, const State&) { return p.geom[0]; }
  MOLE_D static void grad(const WfParams&, const State&, double* g) { g[0] = g[1] = g[2] = 0.0; }  // unimplemented!() upstream; rejected at the API
  MOLE_D static double lap(const WfParams&, const State&) { return 1.0; }   // metrop.rs:252-254
  MOLE_D static void pgrad(const WfParams&, const State&, double*) {}
};

// ---- potentials (src/operator/src/operator.rs) --------------------------------------------------
// Returns V such that the potential part of act_on is V*psi; kinetic part is handled by the caller.
template <int NE>
MOLE_D double mole_potential(const HamParams& h, const double* x) {
  double v = 0.0;
  if (h.kind == MOLE_OP_IONIC_POT || h.kind == MOLE_OP_IONIC || h.kind == MOLE_OP_ELECTRONIC) {
    double pot = 0.0;                                         // IonicPotential::value, operator.rs:25-36
    for (int i = 0; i < h.n_ions; ++i) {
      double d2[NE], ri[NE];
#pragma unroll
      for (int j = 0; j < NE; ++j)
        d2[j] = mole_norm2_3(x[3 * j] - h.ion_pos[3 * i], x[3 * j + 1] - h.ion_pos[3 * i + 1], x[3 * j + 2] - h.ion_pos[3 * i + 2]);
      m_rsqrt_n<NE>(d2, ri);
#pragma unroll
      for (int j = 0; j < NE; ++j) pot = fma(-h.ion_z[i], ri[j], pot);
    }
    v = pot + h.ionic_repulsion;
  }
  if (h.kind == MOLE_OP_ELEC_POT || h.kind == MOLE_OP_ELECTRONIC) {
    double pot = 0.0;                                         // ElectronicPotential::value, operator.rs:80-90
#pragma unroll
    for (int i = 0; i < NE; ++i)
#pragma unroll
      for (int j = i + 1; j < NE; ++j)
        pot += m_rsqrt(mole_norm2_3(x[3 * i] - x[3 * j], x[3 * i + 1] - x[3 * j + 1], x[3 * i + 2] - x[3 * j + 2]));
    v += pot;
  }
  if (h.kind == MOLE_OP_HARMONIC) {                           // custom_operator.rs:56-58
    double n2 = 0.0;
#pragma unroll
    for (int c = 0; c < 3 * NE; ++c) n2 = fma(x[c], x[c], n2);
    v = 0.5 * (h.frequency * h.frequency) * n2;
  }
  return v;
}

MOLE_D bool mole_op_has_kinetic(int kind) {
  return kind == MOLE_OP_KINETIC || kind == MOLE_OP_IONIC || kind == MOLE_OP_ELECTRONIC || kind == MOLE_OP_HARMONIC;
}
