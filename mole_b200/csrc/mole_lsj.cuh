// General LCAO Slater-Jastrow kind (MOLE_WF_LCAO_SJ): the Hydrogen1sBasis / Orbital /
// SpinDeterminantProduct API that tests/helium_lcao.rs:94-101 and tests/hydrogen_molecular_ion_lcao.rs:103-107 name
// (commented out upstream), for up to 5 + 5 electrons over up to 8 centres, times the e-e Jastrow of theory/jastrow.tex:
//   psi = det[phi_k(r_i)]_up det[phi_k(r_i)]_dn exp(f_ee),  phi_k(r) = sum_c C[k][c] exp(-alpha_c |r - R_c|)
//   parameters: C[k][c] (k < n_orb = max(n_up, n_dn), all variational, shared by the spins) and b1..b4: P <= 44.
// This is the widening step with a genuinely large parameter count: the P (P + 1) / 2 SR moments no longer fit a
// walker's registers or shared memory, so the sweep writes the per-sample rows (1, E_L, O_1 .. O_P) and
// mole_gram.cuh contracts them into the Gram matrix (S = O^T O as an fp64 SYRK on the tensor cores).
//
// First correct CUDA path for this kind: ONE THREAD PER WALKER, state in local memory (positions, both inverse
// Slater matrices, the orbital gradients of every electron, grad f), runtime loop bounds.  A move re-evaluates the
// moved electron's orbital row, 2 x 9 Jastrow pairs and a Sherman-Morrison copy of one inverse; the inverses are
// rebuilt from scratch every LSJ_REFRESH_EVERY sweeps.  Semantics as everywhere: metrop.rs:60-96 / :150-212 with
// the Frobenius norm over ALL electrons' drift, samplers.rs:81-117, montecarlo.rs:24-46.
#pragma once
#include "mole_internal.h"
#include "mole_rng.cuh"
#include "mole_math.cuh"

constexpr int LSJ_MAXC = 8;
constexpr int LSJ_MAXN = 5;
constexpr int LSJ_MAXE = 10;
constexpr int LSJ_REFRESH_EVERY = 16;
constexpr int LSJ_THREADS = 64;
// the big per-walker routines are real calls (ABI stack frames): with everything inlined into one 3.5 KB frame ptxas's
// local-memory slot sharing handed two live arrays the same slots (first electron's chi / grad phi overwritten; found
// against the oracle in round 2, independent of -O level and --register-usage-level)
#if defined(MOLE_EMU)
#define LSJ_FN inline
#else
#define LSJ_FN __device__ __noinline__
#endif

// geometry and parameters as the kernels use them (from WfParams: geom = [kappa, n_up, n_dn, N_c, -, -, -, -,
// (R_c, alpha_c) ...], p = C[k][c] at k N_c + c, then b1..b4)
struct LsjConst {
  int nup, ndn, nc, norb, ne, np;
  double kappa, ikappa;
  const double* C;      // into WfParams
  const double* b;
  const double* cen;    // (x, y, z, alpha) per centre
};

MOLE_D LsjConst lsj_const(const WfParams& p) {
  LsjConst c;
  c.kappa = p.geom[0]; c.ikappa = 1.0 / p.geom[0];
  c.nup = (int)p.geom[1]; c.ndn = (int)p.geom[2]; c.nc = (int)p.geom[3];
  c.norb = c.nup > c.ndn ? c.nup : c.ndn;
  c.ne = c.nup + c.ndn;
  c.np = c.norb * c.nc + 4;
  c.C = p.p; c.b = p.p + c.norb * c.nc; c.cen = p.geom + 8;
  return c;
}

struct LsjWalker {
  double x[3 * LSJ_MAXE];
  double minv[2][LSJ_MAXN * LSJ_MAXN];   // Minv[s][k * 5 + i]: orbital k, i-th electron of spin s
  double dphi[LSJ_MAXE][3 * LSJ_MAXN];   // grad phi_k at electron e: [e][3 k + q]
  double gf[LSJ_MAXE][3];                // grad_e f
  double psi;                            // D_up D_dn exp(f), carried
};

MOLE_D int lsj_spin(const LsjConst& c, int e) { return e < c.nup ? 0 : 1; }
MOLE_D int lsj_idx(const LsjConst& c, int e) { return e < c.nup ? e : e - c.nup; }
MOLE_D int lsj_n(const LsjConst& c, int s) { return s == 0 ? c.nup : c.ndn; }

// orbital values, gradients and Laplacians at a point for the orbitals k < n, centre by centre:
// chi_q = exp(-alpha_q |r - R_q|) (the reference's STO, hydrogen_molecular_ion_lcao.rs:25-49, at centre q),
// grad chi_q = -alpha_q chi_q (r - R_q) / |r - R_q|, lap chi_q = alpha_q chi_q (alpha_q - 2 / |r - R_q|)
LSJ_FN void lsj_orbitals(const LsjConst& c, const double* r, int n, double* phi, double* dphi, double* lap, double* chi_out) {
  for (int k = 0; k < n; ++k) { phi[k] = 0.0; dphi[3 * k] = 0.0; dphi[3 * k + 1] = 0.0; dphi[3 * k + 2] = 0.0; lap[k] = 0.0; }
  for (int q = 0; q < c.nc; ++q) {
    const double* g = c.cen + 4 * q;
    const double al = g[3];
    const double dx = r[0] - g[0], dy = r[1] - g[1], dz = r[2] - g[2];
    double irc;
    const double rr = m_sqrt_rsqrt(fma(dz, dz, fma(dy, dy, dx * dx)), irc);
    const double chi = m_exp(-al * rr);
    const double t = -al * chi * irc, l = al * chi * (al - 2.0 * irc);
    chi_out[q] = chi;
    for (int k = 0; k < n; ++k) {
      const double ck = c.C[k * c.nc + q];
      phi[k] = fma(ck, chi, phi[k]);
      dphi[3 * k] = fma(ck * t, dx, dphi[3 * k]);
      dphi[3 * k + 1] = fma(ck * t, dy, dphi[3 * k + 1]);
      dphi[3 * k + 2] = fma(ck * t, dz, dphi[3 * k + 2]);
      lap[k] = fma(ck, l, lap[k]);
    }
  }
}

// Jastrow pair from the squared distance: u, g/r (and div(rhat g), R if wanted)
struct LsjPair { double u, gr, lt, ir, R, iden; };
MOLE_D LsjPair lsj_pair(const LsjConst& c, double r2, bool full) {
  LsjPair o;
  const double r = m_sqrt_rsqrt(r2, o.ir);
  const double E = m_exp(-c.kappa * r);
  o.R = (1.0 - E) * c.ikappa;
  o.iden = m_rcp(fma(c.b[1], o.R, 1.0));
  const double R2 = o.R * o.R, id2 = o.iden * o.iden;
  o.u = fma(R2, fma(c.b[3], o.R, c.b[2]), (c.b[0] * o.R) * o.iden);
  const double du = fma(c.b[0], id2, fma(3.0 * c.b[3], R2, 2.0 * c.b[2] * o.R));
  const double g = E * du;
  o.gr = g * o.ir;
  o.lt = 0.0;
  if (full) {
    const double d2u = fma(-2.0 * c.b[0] * c.b[1], id2 * o.iden, fma(6.0 * c.b[3], o.R, 2.0 * c.b[2]));
    o.lt = fma(2.0, o.gr, fma(E * E, d2u, -c.kappa * g));
  }
  return o;
}

// Gauss-Jordan inverse with partial pivoting of the n x n matrix A[i][k] (row i = electron); writes Minv[k][i] and
// returns the determinant
LSJ_FN double lsj_invert(const double* A, int n, double* minv) {
  double M[LSJ_MAXN][2 * LSJ_MAXN];
  for (int i = 0; i < n; ++i)
    for (int k = 0; k < n; ++k) { M[i][k] = A[i * LSJ_MAXN + k]; M[i][n + k] = (i == k) ? 1.0 : 0.0; }
  double det = 1.0;
  for (int col = 0; col < n; ++col) {
    int piv = col;
    for (int r = col + 1; r < n; ++r)
      if (fabs(M[r][col]) > fabs(M[piv][col])) piv = r;
    if (piv != col) {
      for (int k = 0; k < 2 * n; ++k) { const double t = M[col][k]; M[col][k] = M[piv][k]; M[piv][k] = t; }
      det = -det;
    }
    det *= M[col][col];
    const double ip = m_rcp(M[col][col]);
    for (int k = 0; k < 2 * n; ++k) M[col][k] *= ip;
    for (int r = 0; r < n; ++r) {
      if (r == col) continue;
      const double f = M[r][col];
      for (int k = 0; k < 2 * n; ++k) M[r][k] = fma(-f, M[col][k], M[r][k]);
    }
  }
  for (int k = 0; k < n; ++k)
    for (int i = 0; i < n; ++i) minv[k * LSJ_MAXN + i] = M[k][n + i];   // (A^-1)[k][i]
  return n == 0 ? 1.0 : det;
}

// (re)build everything from the positions
LSJ_FN void lsj_init(const LsjConst& c, LsjWalker& w) {
  double dets[2] = {1.0, 1.0};
  for (int s = 0; s < 2; ++s) {
    const int n = lsj_n(c, s), first = s == 0 ? 0 : c.nup;
    double A[LSJ_MAXN * LSJ_MAXN];
    for (int i = 0; i < n; ++i) {
      double phi[LSJ_MAXN], lap[LSJ_MAXN], chi[LSJ_MAXC];
      lsj_orbitals(c, w.x + 3 * (first + i), n, phi, w.dphi[first + i], lap, chi);
      for (int k = 0; k < n; ++k) A[i * LSJ_MAXN + k] = phi[k];
    }
    dets[s] = lsj_invert(A, n, w.minv[s]);
  }
  double f = 0.0;
  for (int e = 0; e < c.ne; ++e) { w.gf[e][0] = 0.0; w.gf[e][1] = 0.0; w.gf[e][2] = 0.0; }
  for (int i = 0; i < c.ne; ++i)
    for (int j = i + 1; j < c.ne; ++j) {
      const double dx = w.x[3 * i] - w.x[3 * j], dy = w.x[3 * i + 1] - w.x[3 * j + 1], dz = w.x[3 * i + 2] - w.x[3 * j + 2];
      const LsjPair P = lsj_pair(c, fma(dz, dz, fma(dy, dy, dx * dx)), false);
      f += P.u;
      w.gf[i][0] = fma(P.gr, dx, w.gf[i][0]); w.gf[i][1] = fma(P.gr, dy, w.gf[i][1]); w.gf[i][2] = fma(P.gr, dz, w.gf[i][2]);
      w.gf[j][0] = fma(-P.gr, dx, w.gf[j][0]); w.gf[j][1] = fma(-P.gr, dy, w.gf[j][1]); w.gf[j][2] = fma(-P.gr, dz, w.gf[j][2]);
    }
  w.psi = dets[0] * dets[1] * m_exp(f);
}

// grad_e ln D = sum_k grad phi_k(r_e) Minv[k][i] with the given inverse and orbital gradients
MOLE_D void lsj_gradlnD(const double* dphi_e, const double* minv, int n, int i, double* G) {
  double gx = 0.0, gy = 0.0, gz = 0.0;
  for (int k = 0; k < n; ++k) {
    const double m = minv[k * LSJ_MAXN + i];
    gx = fma(dphi_e[3 * k], m, gx); gy = fma(dphi_e[3 * k + 1], m, gy); gz = fma(dphi_e[3 * k + 2], m, gz);
  }
  G[0] = gx; G[1] = gy; G[2] = gz;
}

MOLE_D double lsj_clamp_acceptance(double a, uint32_t compat) {
  if (isnan(a)) return (compat & MOLE_COMPAT_NAN_ACCEPT) ? 1.0 : 0.0;
  return fmin(a, 1.0);
}

// Metropolis::move_state for electron e
template <int METROP>
LSJ_FN bool lsj_move(const LsjConst& c, LsjWalker& w, int e, double param, double sd, RngKey key, uint64_t wid,
                     uint32_t step, uint32_t compat) {
  const int s = lsj_spin(c, e), i = lsj_idx(c, e), n = lsj_n(c, s), first = s == 0 ? 0 : c.nup;
  double xn[3], G[3];
  double u_acc;
  if (METROP == MOLE_METROP_BOX) {
    const MoveDraw d = mole_draw_uniform4(key, wid, step, DOM_MOVE, (uint32_t)e);
    const double lo = -0.5 * param, scale = 0.5 * param - lo;
    xn[0] = w.x[3 * e] + (lo + scale * d.a); xn[1] = w.x[3 * e + 1] + (lo + scale * d.b); xn[2] = w.x[3 * e + 2] + (lo + scale * d.c);
    u_acc = d.u;
  } else {
    const MoveDraw d = mole_draw_normal3_uniform1(key, wid, step, DOM_MOVE, (uint32_t)e);
    lsj_gradlnD(w.dphi[e], w.minv[s], n, i, G);
    xn[0] = (w.x[3 * e] + (G[0] + w.gf[e][0]) * param) + sd * d.a;        // metrop.rs:155-160
    xn[1] = (w.x[3 * e + 1] + (G[1] + w.gf[e][1]) * param) + sd * d.b;
    xn[2] = (w.x[3 * e + 2] + (G[2] + w.gf[e][2]) * param) + sd * d.c;
    u_acc = d.u;
  }
  // orbital row at the trial point, determinant ratio, Sherman-Morrison copy of this spin's inverse
  double phin[LSJ_MAXN], dphin[3 * LSJ_MAXN], lapn[LSJ_MAXN], chin[LSJ_MAXC], mt[LSJ_MAXN * LSJ_MAXN];
  lsj_orbitals(c, xn, n, phin, dphin, lapn, chin);
  double ratio = 0.0;
  for (int k = 0; k < n; ++k) ratio = fma(phin[k], w.minv[s][k * LSJ_MAXN + i], ratio);
  const double inv_ratio = m_rcp(ratio);
  for (int j = 0; j < n; ++j) {
    double v = 0.0;
    for (int k = 0; k < n; ++k) v = fma(phin[k], w.minv[s][k * LSJ_MAXN + j], v);
    const double vr = (j == i) ? -inv_ratio : v * inv_ratio;
    for (int k = 0; k < n; ++k) {
      const double ce = w.minv[s][k * LSJ_MAXN + i];
      mt[k * LSJ_MAXN + j] = fma(-ce, vr, (j == i) ? 0.0 : w.minv[s][k * LSJ_MAXN + j]);
    }
  }
  // Jastrow: nine pairs at the old and at the new position
  double df = 0.0, gfn[LSJ_MAXE][3];
  gfn[e][0] = 0.0; gfn[e][1] = 0.0; gfn[e][2] = 0.0;
  for (int b = 0; b < c.ne; ++b) {
    if (b == e) continue;
    const double ox = w.x[3 * b] - w.x[3 * e], oy = w.x[3 * b + 1] - w.x[3 * e + 1], oz = w.x[3 * b + 2] - w.x[3 * e + 2];
    const double nx = w.x[3 * b] - xn[0], ny = w.x[3 * b + 1] - xn[1], nz = w.x[3 * b + 2] - xn[2];
    const LsjPair Po = lsj_pair(c, fma(oz, oz, fma(oy, oy, ox * ox)), false);
    const LsjPair Pn = lsj_pair(c, fma(nz, nz, fma(ny, ny, nx * nx)), false);
    df += Pn.u - Po.u;
    gfn[b][0] = fma(Pn.gr, nx, fma(-Po.gr, ox, w.gf[b][0]));
    gfn[b][1] = fma(Pn.gr, ny, fma(-Po.gr, oy, w.gf[b][1]));
    gfn[b][2] = fma(Pn.gr, nz, fma(-Po.gr, oz, w.gf[b][2]));
    gfn[e][0] = fma(-Pn.gr, nx, gfn[e][0]); gfn[e][1] = fma(-Pn.gr, ny, gfn[e][1]); gfn[e][2] = fma(-Pn.gr, nz, gfn[e][2]);
  }
  const double q = ratio * m_exp(df);               // psi' / psi
  bool acc;
  if (METROP == MOLE_METROP_BOX) {
    acc = lsj_clamp_acceptance(q * q, compat) > u_acc;                       // metrop.rs:80
  } else {
    // Frobenius norms over ALL electrons' drift at both points (metrop.rs:182-193)
    double sh = 0.0, sl = 0.0;
    for (int j = 0; j < c.ne; ++j) {
      const int sj = lsj_spin(c, j), ij = lsj_idx(c, j), nj = lsj_n(c, sj);
      double Go[3], Gn[3];
      lsj_gradlnD(w.dphi[j], w.minv[sj], nj, ij, Go);
      if (sj == s) lsj_gradlnD(j == e ? dphin : w.dphi[j], mt, nj, ij, Gn);
      else { Gn[0] = Go[0]; Gn[1] = Go[1]; Gn[2] = Go[2]; }
      for (int t = 0; t < 3; ++t) {
        const double dx = (j == e) ? w.x[3 * e + t] - xn[t] : 0.0;
        const double a = dx - (Gn[t] + gfn[j][t]) * param, bb = -dx - (Go[t] + w.gf[j][t]) * param;
        sh = fma(a, a, sh);
        sl = fma(bb, bb, sl);
      }
    }
    const double i2t = 0.5 / param;
    const double shs = sh * i2t, sls = sl * i2t;                             // -ln t_high, -ln t_low
    const bool node = !(ratio > 0.0);                                        // :178-180 (NaN rejects)
    const double ao = fabs(w.psi), an = fabs(w.psi * q);
    double A;
    if (fmax(shs, sls) < 200.0 && ao > 1e-40 && ao < 1e40 && an > 1e-40 && an < 1e40) {
      A = lsj_clamp_acceptance((q * q) * m_exp(sls - shs), compat);          // one exponential, same real number
    } else {                                                                 // the reference's own operation sequence
      const double targ[2] = {-shs, -sls};
      double tv[2];
      m_exp_n<2, true>(targ, tv);
      const double pn = w.psi * q;
      A = lsj_clamp_acceptance(tv[0] * (pn * pn) / (tv[1] * (w.psi * w.psi)), compat);   // :195
    }
    acc = !node && (A > u_acc);
  }
  if (acc) {
    w.x[3 * e] = xn[0]; w.x[3 * e + 1] = xn[1]; w.x[3 * e + 2] = xn[2];
    for (int k = 0; k < 3 * n; ++k) w.dphi[e][k] = dphin[k];
    for (int k = 0; k < n; ++k)
      for (int j = 0; j < n; ++j) w.minv[s][k * LSJ_MAXN + j] = mt[k * LSJ_MAXN + j];
    for (int b = 0; b < c.ne; ++b) { w.gf[b][0] = gfn[b][0]; w.gf[b][1] = gfn[b][1]; w.gf[b][2] = gfn[b][2]; }
    w.psi *= q;
  }
  (void)first;
  return acc;
}

// local quantities of the current configuration: kinetic part -0.5 sum lap psi / psi, potential, and (O != nullptr)
// O_k = d ln psi / d p_k; grad (optional, [ne][3]) = grad ln psi
LSJ_FN void lsj_measure(const LsjConst& c, const HamParams& h, const LsjWalker& w, double& kin, double& pot, double* O,
                        double* grad) {
  double ksum = 0.0;
  if (O)
    for (int k = 0; k < c.np; ++k) O[k] = 0.0;
  for (int e = 0; e < c.ne; ++e) {
    const int s = lsj_spin(c, e), i = lsj_idx(c, e), n = lsj_n(c, s);
    double phi[LSJ_MAXN], dphi[3 * LSJ_MAXN], lap[LSJ_MAXN], chi[LSJ_MAXC];
    lsj_orbitals(c, w.x + 3 * e, n, phi, dphi, lap, chi);
    double lapD = 0.0, G[3];
    for (int k = 0; k < n; ++k) lapD = fma(lap[k], w.minv[s][k * LSJ_MAXN + i], lapD);
    lsj_gradlnD(dphi, w.minv[s], n, i, G);
    const double gg = G[0] * w.gf[e][0] + G[1] * w.gf[e][1] + G[2] * w.gf[e][2];
    const double ff = w.gf[e][0] * w.gf[e][0] + w.gf[e][1] * w.gf[e][1] + w.gf[e][2] * w.gf[e][2];
    ksum += lapD + (2.0 * gg + ff);
    if (grad) { grad[3 * e] = G[0] + w.gf[e][0]; grad[3 * e + 1] = G[1] + w.gf[e][1]; grad[3 * e + 2] = G[2] + w.gf[e][2]; }
    if (O)                                                         // d ln D_s / d C[k][q] = sum_i Minv[k][i] chi_q(r_i)
      for (int k = 0; k < n; ++k)
        for (int qc = 0; qc < c.nc; ++qc) O[k * c.nc + qc] = fma(w.minv[s][k * LSJ_MAXN + i], chi[qc], O[k * c.nc + qc]);
  }
  double vee = 0.0, lts = 0.0, db[4] = {0.0, 0.0, 0.0, 0.0};
  for (int i = 0; i < c.ne; ++i)
    for (int j = i + 1; j < c.ne; ++j) {
      const double dx = w.x[3 * i] - w.x[3 * j], dy = w.x[3 * i + 1] - w.x[3 * j + 1], dz = w.x[3 * i + 2] - w.x[3 * j + 2];
      const LsjPair P = lsj_pair(c, fma(dz, dz, fma(dy, dy, dx * dx)), true);
      vee += P.ir;
      lts += P.lt;
      const double qd = P.R * P.iden;
      db[0] += qd; db[1] = fma(qd, qd, db[1]); db[2] = fma(P.R, P.R, db[2]); db[3] = fma(P.R * P.R, P.R, db[3]);
    }
  ksum = fma(2.0, lts, ksum);                                       // sum_i lap_i f = 2 sum_pairs div(rhat g)
  if (O) { O[c.np - 4] = db[0]; O[c.np - 3] = -c.b[0] * db[1]; O[c.np - 2] = db[2]; O[c.np - 1] = db[3]; }
  const bool has_kin = h.kind == MOLE_OP_KINETIC || h.kind == MOLE_OP_IONIC || h.kind == MOLE_OP_ELECTRONIC || h.kind == MOLE_OP_HARMONIC;
  kin = has_kin ? -0.5 * ksum : 0.0;
  double v = 0.0;
  if (h.kind == MOLE_OP_IONIC_POT || h.kind == MOLE_OP_IONIC || h.kind == MOLE_OP_ELECTRONIC) {
    for (int a = 0; a < h.n_ions; ++a)
      for (int e = 0; e < c.ne; ++e) {
        const double dx = w.x[3 * e] - h.ion_pos[3 * a], dy = w.x[3 * e + 1] - h.ion_pos[3 * a + 1], dz = w.x[3 * e + 2] - h.ion_pos[3 * a + 2];
        v = fma(-h.ion_z[a], m_rsqrt(fma(dz, dz, fma(dy, dy, dx * dx))), v);
      }
    v += h.ionic_repulsion;
  }
  if (h.kind == MOLE_OP_ELEC_POT || h.kind == MOLE_OP_ELECTRONIC) v += vee;
  if (h.kind == MOLE_OP_HARMONIC) {
    double r2 = 0.0;
    for (int k = 0; k < 3 * c.ne; ++k) r2 = fma(w.x[k], w.x[k], r2);
    v = fma(0.5 * h.frequency * h.frequency, r2, v);
  }
  pot = (h.kind == MOLE_OP_KINETIC) ? 0.0 : v;
}

// ------------------------------------------------------------------ batched evaluation (parity entry point)
__global__ void __launch_bounds__(LSJ_THREADS) lsj_eval_kernel(const double* __restrict__ x, int64_t W, WfParams p, HamParams h,
                                                               int have_ham, double* psi, double* grad, double* lap,
                                                               double* hpsi, double* pgrad) {
  mole_math_smem_init();
  const LsjConst c = lsj_const(p);
  const int64_t wi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (wi >= W) return;
  LsjWalker w;
  for (int k = 0; k < 3 * c.ne; ++k) w.x[k] = x[(size_t)k * W + wi];
  lsj_init(c, w);
  HamParams hk = h;
  hk.kind = MOLE_OP_KINETIC;
  double kin, pot, O[MOLE_WF_MAX_PARAMS], g[3 * LSJ_MAXE];
  lsj_measure(c, hk, w, kin, pot, O, g);
  if (psi) psi[wi] = w.psi;
  if (lap) lap[wi] = -2.0 * kin * w.psi;
  if (grad)
    for (int k = 0; k < 3 * c.ne; ++k) grad[(size_t)wi * 3 * c.ne + k] = w.psi * g[k];
  if (pgrad)
    for (int k = 0; k < c.np; ++k) pgrad[(size_t)wi * c.np + k] = w.psi * O[k];
  if (hpsi && have_ham) {
    double kh, ph;
    lsj_measure(c, h, w, kh, ph, nullptr, nullptr);
    hpsi[wi] = kh * w.psi + ph * w.psi;
  }
}

// ------------------------------------------------------------------ fused sweep
// One thread per walker (grid-stride).  The ten scalar sums go through the block-tree reduction of mole_kernels.cuh;
// with OPT every sample also writes its row (1, E_L, O_1 .. O_P) to sp.osamp[(sample * cols + col) * W + w]
// (cols = P + 2; a non-finite sample writes a row of zeros, so that it drops out of the Gram matrix, and is counted).
template <int METROP, bool OPT>
__global__ void __launch_bounds__(LSJ_THREADS) lsj_sweep_kernel(const SweepParams sp) {
  mole_math_smem_init();
  using A = Acc<0>;
  A acc;
  acc.zero(nullptr);
  const LsjConst c = lsj_const(sp.wf);
  const double sd = sqrt(sp.metrop_param);
  const bool want_e = (sp.observables & MOLE_OBS_ENERGY) != 0;
  const int64_t W = sp.W;
  const int cols = c.np + 2;
  for (int64_t wi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; wi < W; wi += (int64_t)gridDim.x * blockDim.x) {
    LsjWalker w;
    for (int k = 0; k < 3 * c.ne; ++k) w.x[k] = sp.x[(size_t)k * W + wi];
    lsj_init(c, w);
    const uint64_t wid = sp.walker_offset + (uint64_t)wi;
    double blk = sp.blk[wi];
    int fill = sp.blk_fill, nbad = 0;
    for (int s = 0; s < sp.n_sweeps; ++s) {
      const uint32_t step = sp.step0 + (uint32_t)s;
      if (s > 0 && (s % LSJ_REFRESH_EVERY) == 0) lsj_init(c, w);
      for (int e = 0; e < c.ne; ++e) {                            // Sampler::move_state, samplers.rs:106-117
        const bool ok = lsj_move<METROP>(c, w, e, sp.metrop_param, sd, sp.key, wid, step, sp.compat);
        acc.v[ACC_NACC] += ok ? 1.0 : 0.0;
        acc.v[ACC_NMOVE] += 1.0;
        if (sp.tr_accept) sp.tr_accept[((size_t)s * c.ne + e) * W + wi] = ok ? 1 : 0;
      }
      if (s < sp.n_discard) continue;                             // block 0 = equilibration, montecarlo.rs:36
      const int64_t si = s - sp.n_discard;
      double kin = 0.0, pot = 0.0, O[MOLE_WF_MAX_PARAMS];
      if (want_e || OPT || (sp.observables & MOLE_OBS_KINETIC)) lsj_measure(c, sp.ham, w, kin, pot, OPT ? O : nullptr, nullptr);
      const double el = kin + pot;
      bool bad = want_e && !isfinite(el);
      if (OPT) {
        const bool quirk = (sp.compat & MOLE_COMPAT_VECTOR_DIV) != 0;
        for (int k = 0; k < c.np; ++k) {
          if (quirk) O[k] = 1.0 / (w.psi * w.psi * O[k]);          // stored sample 1/d_k psi (operator/src/traits.rs:149-150)
          bad = bad || !isfinite(O[k]);
          if (sp.tr_pgrad) sp.tr_pgrad[((size_t)si * c.np + k) * W + wi] = w.psi * O[k];
        }
      }
      if (bad) atomicAdd(sp.acc + ACC_BAD, 1.0);
      if (want_e) {
        if (!bad) {
          acc.v[ACC_N] += 1.0;
          acc.v[ACC_E] += el;
          acc.v[ACC_E2] = fma(el, el, acc.v[ACC_E2]);
          acc.v[ACC_T] += kin;
          blk += el;
        } else {
          ++nbad;
        }
        if (++fill == sp.block_size) {                            // block means, vmc.rs:158-164
          if (nbad < sp.block_size) {
            const double bm = blk / (double)(sp.block_size - nbad);
            acc.v[ACC_B] += bm;
            acc.v[ACC_B2] = fma(bm, bm, acc.v[ACC_B2]);
            acc.v[ACC_NB] += 1.0;
          }
          blk = 0.0; fill = 0; nbad = 0;
        }
        if (sp.tr_energy) sp.tr_energy[(size_t)si * W + wi] = el;
      }
      if (sp.tr_kinetic && (sp.observables & MOLE_OBS_KINETIC)) sp.tr_kinetic[(size_t)si * W + wi] = kin;
      if (!bad) acc.v[ACC_PSI] += w.psi;
      if (sp.tr_wfvalue) sp.tr_wfvalue[(size_t)si * W + wi] = w.psi;
      if (OPT && sp.osamp) {
        double* row = sp.osamp + (size_t)si * cols * W + wi;
        row[0] = bad ? 0.0 : 1.0;
        row[(size_t)W] = bad ? 0.0 : el;
        for (int k = 0; k < c.np; ++k) row[(size_t)(2 + k) * W] = bad ? 0.0 : O[k];
      }
    }
    for (int k = 0; k < 3 * c.ne; ++k) sp.x[(size_t)k * W + wi] = w.x[k];
    sp.blk[wi] = blk;
  }
  double all[A::LEN];
  acc.gather(all);
  __syncthreads();
  mole_block_reduce_to_global<A::LEN>(all, sp.partials, sp.acc, sp.ticket, [](int i) { return A::slot(i); });
}

// ------------------------------------------------------------------ DMC time step (dmc.rs:87-130)
__global__ void __launch_bounds__(LSJ_THREADS) lsj_dmc_kernel(const DmcParams dp) {
  mole_math_smem_init();
  const LsjConst c = lsj_const(dp.wf);
  const double sd = sqrt(dp.tau_move);
  double s_we = 0.0, s_w = 0.0, s_wn = 0.0, m_wn = 0.0;
  const int64_t W = dp.W;
  for (int64_t wi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; wi < W; wi += (int64_t)gridDim.x * blockDim.x) {
    LsjWalker w;
    for (int k = 0; k < 3 * c.ne; ++k) w.x[k] = dp.x[(size_t)k * W + wi];
    lsj_init(c, w);
    double kin, pot, e_old;
    if (dp.el_cached) e_old = dp.el[wi];
    else { lsj_measure(c, dp.ham, w, kin, pot, nullptr, nullptr); e_old = kin + pot; }
    const uint64_t wid = dp.walker_offset + (uint64_t)wi;
    for (int e = 0; e < c.ne; ++e) lsj_move<MOLE_METROP_DIFFUSE>(c, w, e, dp.tau_move, sd, dp.key, wid, dp.step, dp.compat);
    lsj_init(c, w);                                               // E_new from a fresh inverse, like the next step's E_old
    lsj_measure(c, dp.ham, w, kin, pot, nullptr, nullptr);
    const double e_new = kin + pot;
    const double w_in = dp.w[wi];
    const double w_up = w_in * exp(-dp.tau_weight * ((e_old + e_new) / 2.0 - dp.e_ref));
    const bool bad = !isfinite(e_old) || !isfinite(e_new) || !isfinite(w_up);
    if (bad) atomicAdd(dp.health, 1.0);
    const double wt = bad ? 0.0 : w_in;
    s_we = fma(wt, bad ? 0.0 : e_old, s_we);
    s_w += wt;
    const double wn = bad ? 0.0 : w_up;
    s_wn += wn;
    m_wn = fmax(m_wn, wn);
    dp.w[wi] = wn;
    dp.el[wi] = bad ? 0.0 : e_new;
    for (int k = 0; k < 3 * c.ne; ++k) dp.x[(size_t)k * W + wi] = w.x[k];
  }
  __shared__ double sm[32][4];
  __shared__ bool is_last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s_we += __shfl_xor_sync(0xffffffffu, s_we, o);
    s_w += __shfl_xor_sync(0xffffffffu, s_w, o);
    s_wn += __shfl_xor_sync(0xffffffffu, s_wn, o);
    m_wn = fmax(m_wn, __shfl_xor_sync(0xffffffffu, m_wn, o));
  }
  if (lane == 0) { sm[warp][0] = s_we; sm[warp][1] = s_w; sm[warp][2] = s_wn; sm[warp][3] = m_wn; }
  __syncthreads();
  if (threadIdx.x < 4) {
    double s = 0.0;
    for (int q = 0; q < nwarp; ++q) s = (threadIdx.x == 3) ? fmax(s, sm[q][3]) : s + sm[q][threadIdx.x];
    dp.partials[(size_t)blockIdx.x * 4 + threadIdx.x] = s;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = (atomicInc(dp.ticket, gridDim.x - 1) == gridDim.x - 1);
  __syncthreads();
  if (is_last) {
    __threadfence();
    mole_dmc_fold_partials(dp.partials, gridDim.x, dp.red, sm);
  }
}
