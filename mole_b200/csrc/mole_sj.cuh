// Slater-Jastrow kind (MOLE_WF_SLATER_JASTROW, SURVEY.md §8(c) synthetic config 5):
//   psi = det[phi_k(r_i)]_up * det[phi_k(r_i)]_dn * exp(f_ee),   orbitals 1s, 2s, 2p_{x,y,z} (STO),
//   f_ee = sum_{i<j} u(R_ij), u(R) = b1 R/(1+b2 R) + b3 R^2 + b4 R^3, R = (1-exp(-kappa r))/kappa
//   (theory/jastrow.tex:23-31 with the erratum of SURVEY.md §8(c)).
//
// Layout: FIVE lanes cooperate on one walker (six walkers per warp, lanes 30/31 idle).  Lane gl owns
// the up electron gl and the down electron gl: position, cached orbital exponentials, its column of
// each inverse Slater matrix, grad f and grad ln D live in registers; the pair cache (u, g/r,
// laplacian term, R, 1/(1+b2 R), 1/r for the 45 pairs), the orbital gradients and a copy of the
// positions live in shared memory.  A single-electron move re-evaluates only what changed: one
// orbital row, a Sherman-Morrison update of one 5x5 inverse (refreshed from scratch every sweep),
// nine Jastrow pairs.  Reductions over the five lanes are warp shuffles in a fixed order.
// The Metropolis semantics are the reference's (src/metropolis/src/metrop.rs:60-96,150-212),
// including the Frobenius norm over ALL electrons' drift in t_high / t_low.
#pragma once
#include <cuda_runtime.h>
#include "mole_internal.h"
#include "mole_rng.cuh"

constexpr int SJ_WPW = 6;                       // walkers per warp
constexpr int SJ_WARPS = 4;
constexpr int SJ_THREADS = 32 * SJ_WARPS;
constexpr int SJ_WPB = SJ_WPW * SJ_WARPS;       // walkers per CTA
constexpr int SJ_NPAIR = 45;
// shared memory per walker, in doubles
constexpr int SJ_PC = 6 * SJ_NPAIR;             // pair cache: u, g/r, lap term, R, 1/(1+b2R), 1/r
constexpr int SJ_GPH = 2 * 15 * 5;              // grad phi_k at every electron, [slot][3k+c][lane]
constexpr int SJ_XS = 30;                       // positions by slot id
constexpr int SJ_TSC = 25;                      // transpose scratch for the inverse
constexpr int SJ_OS = 8;                        // O_k of the current sample
constexpr int SJ_SMEM_PER_WALKER = SJ_PC + SJ_GPH + SJ_XS + SJ_TSC + SJ_OS;
constexpr size_t SJ_SMEM_BYTES = (size_t)SJ_SMEM_PER_WALKER * SJ_WPB * sizeof(double);
constexpr int SJ_NP = 7;
constexpr int SJ_NACC = 10 + 2 * SJ_NP + SJ_NP * (SJ_NP + 1) / 2;   // 52 compact accumulator entries
constexpr int SJ_ACC_PER_LANE = (SJ_NACC + 4) / 5;                   // 11
constexpr unsigned SJ_FULL = 0xffffffffu;

struct SjConst {
  int nup, ndn;
  double kappa, ikappa, z1, z2, z3, b1, b2, b3, b4;
};

__constant__ signed char c_sj_pair_a[SJ_NPAIR], c_sj_pair_b[SJ_NPAIR];   // slot ids of pair p (a<b)
__constant__ signed char c_sj_oo_k[28], c_sj_oo_l[28];                   // (k<=l) of packed O_k O_l entry q

MOLE_D int sj_pidx(int a, int b) {              // unordered pair of slot ids (0..9) -> 0..44
  const int lo = a < b ? a : b, hi = a < b ? b : a;
  return lo * (19 - lo) / 2 + (hi - lo - 1);
}

struct SjShared {
  double* pc; double* gph; double* xs; double* tsc; double* os;
};

struct SjLane {
  int lane, gl, base;
  bool act;            // this 5-lane group holds a real walker
  bool wr;             // lanes 30/31 alias group 0's shared memory and must never store to it
  bool val[2];         // slot validity: gl < n_up / gl < n_dn
  double x[2][3];      // own electrons
  double ec[2][3];     // exp(-zeta_m r) at own electrons
  double minv[2][5];   // column gl of the inverse Slater matrices
  double gf[2][3];     // grad_i f
  double G[2][3];      // grad_i ln D
  double psi, f;       // replicated over the group
  double det[2];
};

// sum over the five lanes of a group, fixed order ((v0+v4)+v2)+(v1+v3), result replicated
MOLE_D double sj_gsum(double v, const SjLane& L) {
  double y = __shfl_sync(SJ_FULL, v, L.lane + 4);
  if (L.gl == 0) v += y;
  y = __shfl_sync(SJ_FULL, v, L.lane + 2);
  if (L.gl < 2) v += y;
  y = __shfl_sync(SJ_FULL, v, L.lane + 1);
  if (L.gl == 0) v += y;
  return __shfl_sync(SJ_FULL, v, L.base);
}

struct SjPair { double u, gr, lt, R, iden, ir; };

// acceptance.min(1.0) with the NaN policy of include/mole_b200.h (MOLE_COMPAT_NAN_ACCEPT)
MOLE_D double sj_clamp_acceptance(double a, uint32_t compat) {
  if (isnan(a)) return (compat & MOLE_COMPAT_NAN_ACCEPT) ? 1.0 : 0.0;
  return fmin(a, 1.0);
}

// pair function from the squared distance (theory/jastrow.tex:23-31,45-48,68-71,82-97)
MOLE_D SjPair sj_pair(const SjConst& c, double r2) {
  SjPair o;
  const double r = sqrt(r2);
  const double E = exp(-c.kappa * r);
  o.R = (1.0 - E) * c.ikappa;
  const double den = fma(c.b2, o.R, 1.0);
  const double inv = 1.0 / (den * r);
  o.iden = inv * r;
  o.ir = inv * den;
  const double R2 = o.R * o.R;
  o.u = fma(c.b1 * o.R, o.iden, fma(c.b4 * R2, o.R, c.b3 * R2));
  const double id2 = o.iden * o.iden;
  const double du = fma(c.b1, id2, fma(3.0 * c.b4, R2, 2.0 * c.b3 * o.R));
  const double d2u = fma(-2.0 * c.b1 * c.b2, id2 * o.iden, fma(6.0 * c.b4, o.R, 2.0 * c.b3));
  const double g = E * du;
  o.gr = g * o.ir;
  o.lt = fma(2.0, o.gr, fma(E * E, d2u, -c.kappa * g));   // div(rhat g) = 2 g/r + dg/dr
  return o;
}

// orbital values at a point from r and the three exponentials; entries k >= n are zero (padding)
MOLE_D void sj_phi(const double* x, double r, const double* e, int n, double* phi) {
  phi[0] = n > 0 ? e[0] : 0.0;
  phi[1] = n > 1 ? r * e[1] : 0.0;
  phi[2] = n > 2 ? x[0] * e[2] : 0.0;
  phi[3] = n > 3 ? x[1] * e[2] : 0.0;
  phi[4] = n > 4 ? x[2] * e[2] : 0.0;
}
// grad phi_k, laid out [3k+c]
MOLE_D void sj_gphi(const SjConst& c, const double* x, double r, double ir, const double* e, double* g) {
  const double c0 = -c.z1 * e[0] * ir, c1 = (1.0 - c.z2 * r) * e[1] * ir, c2 = -c.z3 * ir;
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    g[q] = c0 * x[q];
    g[3 + q] = c1 * x[q];
  }
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int q = 0; q < 3; ++q) g[6 + 3 * a + q] = e[2] * (c2 * x[a] * x[q] + (a == q ? 1.0 : 0.0));
}

// Gauss-Jordan inverse of the 5x5 matrix whose row gl is M[] (row = lane), partial pivoting by
// row selection; returns row gl of the inverse and the determinant (replicated).
MOLE_D void sj_invert(double* M, double* out, double& det, const SjLane& L) {
  double I[5];
#pragma unroll
  for (int q = 0; q < 5; ++q) I[q] = (q == L.gl) ? 1.0 : 0.0;
  bool used = false;
  int prow = 0;
  double d = 1.0;
#pragma unroll
  for (int c = 0; c < 5; ++c) {
    const double a = used ? -1.0 : fabs(M[c]);
    double best = -2.0;
    int who = 0;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      const double ai = __shfl_sync(SJ_FULL, a, L.base + i);
      if (ai > best) { best = ai; who = i; }
    }
    const double pv = __shfl_sync(SJ_FULL, M[c], L.base + who);
    d *= pv;
    const double ipv = 1.0 / pv;
    const bool me = (L.gl == who);
    const double fac = M[c];
#pragma unroll
    for (int q = 0; q < 5; ++q) {
      const double pm = __shfl_sync(SJ_FULL, M[q], L.base + who) * ipv;
      const double pi = __shfl_sync(SJ_FULL, I[q], L.base + who) * ipv;
      M[q] = me ? pm : fma(-fac, pm, M[q]);
      I[q] = me ? pi : fma(-fac, pi, I[q]);
    }
    if (me) { used = true; prow = c; }
  }
  int src = 0, p[5];
#pragma unroll
  for (int i = 0; i < 5; ++i) {
    p[i] = __shfl_sync(SJ_FULL, prow, L.base + i);
    if (p[i] == L.gl) src = i;
  }
  int inv = 0;
#pragma unroll
  for (int i = 0; i < 5; ++i)
#pragma unroll
    for (int j = i + 1; j < 5; ++j) inv += (p[i] > p[j]) ? 1 : 0;
#pragma unroll
  for (int q = 0; q < 5; ++q) out[q] = __shfl_sync(SJ_FULL, I[q], L.base + src);
  det = (inv & 1) ? -d : d;
}

// inverse Slater matrix, determinant and grad ln D of spin slot T from the cached exponentials
template <int T>
MOLE_D void sj_refresh_spin(const SjConst& c, SjLane& L, const SjShared& sm) {
  const int n = T == 0 ? c.nup : c.ndn;
  double phi[5];
  const double r = sqrt(fma(L.x[T][2], L.x[T][2], fma(L.x[T][1], L.x[T][1], L.x[T][0] * L.x[T][0])));
  sj_phi(L.x[T], r, L.ec[T], n, phi);
#pragma unroll
  for (int k = 0; k < 5; ++k)
    if (L.wr) sm.tsc[L.gl * 5 + k] = L.val[T] ? phi[k] : (k == L.gl ? 1.0 : 0.0);   // A[i=gl][k], identity padding
  __syncwarp();
  double M[5];
#pragma unroll
  for (int i = 0; i < 5; ++i) M[i] = sm.tsc[i * 5 + L.gl];   // row gl of A^T
  __syncwarp();
  sj_invert(M, L.minv[T], L.det[T], L);
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < 5; ++k) s = fma(sm.gph[(T * 15 + 3 * k + q) * 5 + L.gl], L.minv[T][k], s);
    L.G[T][q] = s;
  }
}

MOLE_D bool sj_slot_valid(const SjConst& c, int sid) { return sid < 5 ? sid < c.nup : (sid - 5) < c.ndn; }

// full (re)initialisation of the cooperative state from the positions in L.x
MOLE_D void sj_init(const SjConst& c, SjLane& L, const SjShared& sm) {
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    const double r = sqrt(fma(L.x[t][2], L.x[t][2], fma(L.x[t][1], L.x[t][1], L.x[t][0] * L.x[t][0])));
    const double rs = L.val[t] ? r : 1.0;
    const double ir = 1.0 / rs;
    L.ec[t][0] = exp(-c.z1 * rs);
    L.ec[t][1] = exp(-c.z2 * rs);
    L.ec[t][2] = exp(-c.z3 * rs);
    double g[15];
    sj_gphi(c, L.x[t], rs, ir, L.ec[t], g);
#pragma unroll
    for (int q = 0; q < 15; ++q)
      if (L.wr) sm.gph[(t * 15 + q) * 5 + L.gl] = g[q];
#pragma unroll
    for (int q = 0; q < 3; ++q)
      if (L.wr) sm.xs[(t * 5 + L.gl) * 3 + q] = L.x[t][q];
  }
  __syncwarp();
  sj_refresh_spin<0>(c, L, sm);
  sj_refresh_spin<1>(c, L, sm);
  // Jastrow from scratch: every lane sums over the partners of its own electrons
  double fl = 0.0;
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    const int a = t * 5 + L.gl;
    double gx = 0.0, gy = 0.0, gz = 0.0;
    for (int b = 0; b < 10; ++b) {
      const bool pv = L.val[t] && b != a && sj_slot_valid(c, b);
      const double dx = L.x[t][0] - sm.xs[b * 3], dy = L.x[t][1] - sm.xs[b * 3 + 1], dz = L.x[t][2] - sm.xs[b * 3 + 2];
      const double r2 = pv ? fma(dz, dz, fma(dy, dy, dx * dx)) : 1.0;
      const SjPair P = sj_pair(c, r2);
      if (pv) {
        gx = fma(P.gr, dx, gx); gy = fma(P.gr, dy, gy); gz = fma(P.gr, dz, gz);
        if (a < b && L.wr) {
          const int p = sj_pidx(a, b);
          sm.pc[p] = P.u; sm.pc[SJ_NPAIR + p] = P.gr; sm.pc[2 * SJ_NPAIR + p] = P.lt;
          sm.pc[3 * SJ_NPAIR + p] = P.R; sm.pc[4 * SJ_NPAIR + p] = P.iden; sm.pc[5 * SJ_NPAIR + p] = P.ir;
          fl += P.u;
        }
      }
    }
    L.gf[t][0] = gx; L.gf[t][1] = gy; L.gf[t][2] = gz;
  }
  __syncwarp();
  L.f = sj_gsum(fl, L);
  L.psi = L.det[0] * L.det[1] * exp(L.f);
}

// once per sweep: rebuild the inverses from scratch (bounds the Sherman-Morrison round-off) and
// re-sum the Jastrow exponent from the pair cache
MOLE_D void sj_refresh(const SjConst& c, SjLane& L, const SjShared& sm) {
  sj_refresh_spin<0>(c, L, sm);
  sj_refresh_spin<1>(c, L, sm);
  double fl = 0.0;
  for (int p = L.gl; p < SJ_NPAIR; p += 5)
    if (sj_slot_valid(c, c_sj_pair_a[p]) && sj_slot_valid(c, c_sj_pair_b[p])) fl += sm.pc[p];
  L.f = sj_gsum(fl, L);
  L.psi = L.det[0] * L.det[1] * exp(L.f);
}

// Metropolis::move_state for electron `el` of spin slot S.  d = the pre-generated draws of THIS lane's
// slot-S electron (only the owner's are used).  Returns the accept decision (uniform over the group).
template <int S, int METROP>
MOLE_D bool sj_move(const SjConst& c, SjLane& L, const SjShared& sm, int el, const MoveDraw& d, double param, double sd,
                    uint32_t compat) {
  const int n = S == 0 ? c.nup : c.ndn;
  const int own = L.base + el;
  const bool isown = (L.gl == el);
  const int sid_e = S * 5 + el;
  double xo[3], xn[3];
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    xo[q] = L.x[S][q];
    if (METROP == MOLE_METROP_BOX) {
      const double lo = -0.5 * param, scale = 0.5 * param - lo;
      const double u = q == 0 ? d.a : (q == 1 ? d.b : d.c);
      xn[q] = xo[q] + (lo + scale * u);                                   // metrop.rs:63-68
    } else {
      const double xi = q == 0 ? d.a : (q == 1 ? d.b : d.c);
      xn[q] = (xo[q] + (L.G[S][q] + L.gf[S][q]) * param) + sd * xi;       // metrop.rs:155-160
    }
    xn[q] = __shfl_sync(SJ_FULL, xn[q], own);
    xo[q] = __shfl_sync(SJ_FULL, xo[q], own);
  }
  const double u_acc = __shfl_sync(SJ_FULL, d.u, own);
  // orbitals of the moved electron at the trial position (one exponential per lane, then shared)
  const double rn = sqrt(fma(xn[2], xn[2], fma(xn[1], xn[1], xn[0] * xn[0])));
  const double irn = 1.0 / rn;
  const double zm = L.gl == 0 ? c.z1 : (L.gl == 1 ? c.z2 : c.z3);
  const double ex = exp(-zm * rn);
  double en[3];
  en[0] = __shfl_sync(SJ_FULL, ex, L.base);
  en[1] = __shfl_sync(SJ_FULL, ex, L.base + 1);
  en[2] = __shfl_sync(SJ_FULL, ex, L.base + 2);
  double phin[5];
  sj_phi(xn, rn, en, n, phin);
  // determinant ratio and Sherman-Morrison update of this lane's column
  double v = 0.0;
#pragma unroll
  for (int k = 0; k < 5; ++k) v = fma(phin[k], L.minv[S][k], v);
  const double ratio = __shfl_sync(SJ_FULL, v, own);
  const double inv_ratio = 1.0 / ratio;
  const double vr = v * inv_ratio;
  double mt[5];
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    const double ce = __shfl_sync(SJ_FULL, L.minv[S][k], own);
    mt[k] = isown ? ce * inv_ratio : fma(-ce, vr, L.minv[S][k]);
  }
  // trial grad ln D of this lane's slot-S electron
  double gn[15];
  sj_gphi(c, xn, rn, irn, en, gn);
  double Gt[3] = {0.0, 0.0, 0.0};
  if (METROP == MOLE_METROP_DIFFUSE) {
#pragma unroll
    for (int k = 0; k < 5; ++k)
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        const double gk = isown ? gn[3 * k + q] : sm.gph[(S * 15 + 3 * k + q) * 5 + L.gl];
        Gt[q] = fma(gk, mt[k], Gt[q]);
      }
  }
  // nine Jastrow pairs of the moved electron, two per lane
  double dfl = 0.0, ge[3] = {0.0, 0.0, 0.0}, gft[2][3];
  SjPair P[2];
  int pid[2];
  bool pv[2];
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    const int b = t * 5 + L.gl;
    pv[t] = L.val[t] && !(t == S && isown);
    pid[t] = pv[t] ? sj_pidx(sid_e, b) : 0;
    const double dnx = xn[0] - L.x[t][0], dny = xn[1] - L.x[t][1], dnz = xn[2] - L.x[t][2];
    const double r2 = pv[t] ? fma(dnz, dnz, fma(dny, dny, dnx * dnx)) : 1.0;
    P[t] = sj_pair(c, r2);
    const double u_old = sm.pc[pid[t]], gr_old = sm.pc[SJ_NPAIR + pid[t]];
    const double m = pv[t] ? 1.0 : 0.0;
    dfl = fma(m, P[t].u - u_old, dfl);
    const double gn_ = m * P[t].gr, go_ = m * gr_old;
    // grad_b f changes by the (b,e) term: -(x_e - x_b) g/r
    gft[t][0] = L.gf[t][0] + go_ * (xo[0] - L.x[t][0]) - gn_ * dnx;
    gft[t][1] = L.gf[t][1] + go_ * (xo[1] - L.x[t][1]) - gn_ * dny;
    gft[t][2] = L.gf[t][2] + go_ * (xo[2] - L.x[t][2]) - gn_ * dnz;
    ge[0] = fma(gn_, dnx, ge[0]); ge[1] = fma(gn_, dny, ge[1]); ge[2] = fma(gn_, dnz, ge[2]);
  }
  const double df = sj_gsum(dfl, L);
  bool node;
  double A;
  if (METROP == MOLE_METROP_DIFFUSE) {
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      const double t = sj_gsum(ge[q], L);           // grad_e f at the trial position, summed from scratch
      if (isown) gft[S][q] = t;
    }
    // Frobenius norms over ALL electrons' drift (metrop.rs:182-193)
    double sh = 0.0, sl = 0.0;
#pragma unroll
    for (int t = 0; t < 2; ++t)
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        const double vn = (t == S ? Gt[q] : L.G[t][q]) + gft[t][q];
        const double vo = L.G[t][q] + L.gf[t][q];
        const double dx = (t == S && isown) ? xo[q] - xn[q] : 0.0;
        const double a = dx - vn * param, b = -dx - vo * param;
        const double m = L.val[t] ? 1.0 : 0.0;
        sh = fma(m * a, a, sh);
        sl = fma(m * b, b, sl);
      }
    sh = sj_gsum(sh, L);
    sl = sj_gsum(sl, L);
    // exp(df), t_high, t_low: one exponential per lane
    const double arg = L.gl == 0 ? df : (L.gl == 1 ? -sh / (2.0 * param) : -sl / (2.0 * param));
    const double e3 = exp(arg);
    const double ef = __shfl_sync(SJ_FULL, e3, L.base);
    const double th = __shfl_sync(SJ_FULL, e3, L.base + 1);
    const double tl = __shfl_sync(SJ_FULL, e3, L.base + 2);
    const double q = ratio * ef;                               // psi'/psi
    node = !(ratio > 0.0);                                     // signum(psi') != signum(psi) or NaN, :178-180
    A = sj_clamp_acceptance(th * (q * q) / tl, compat);        // :195
    if (!node && A > u_acc) L.psi *= q;
  } else {
    const double q = ratio * exp(df);
    node = false;
    A = sj_clamp_acceptance(q * q, compat);                    // metrop.rs:80
    if (A > u_acc) L.psi *= q;
  }
  const bool acc = !node && (A > u_acc);
  if (acc) {
    if (isown) {
#pragma unroll
      for (int q = 0; q < 3; ++q) { L.x[S][q] = xn[q]; L.ec[S][q] = en[q]; }
      if (L.act) {
#pragma unroll
        for (int q = 0; q < 3; ++q) sm.xs[sid_e * 3 + q] = xn[q];
#pragma unroll
        for (int q = 0; q < 15; ++q) sm.gph[(S * 15 + q) * 5 + L.gl] = gn[q];
      }
    }
#pragma unroll
    for (int k = 0; k < 5; ++k) L.minv[S][k] = mt[k];
    L.det[S] *= ratio;
    L.f += df;
    if (METROP == MOLE_METROP_DIFFUSE) {
#pragma unroll
      for (int q = 0; q < 3; ++q) { L.G[S][q] = Gt[q]; L.gf[0][q] = gft[0][q]; L.gf[1][q] = gft[1][q]; }
    }
#pragma unroll
    for (int t = 0; t < 2; ++t)
      if (pv[t] && L.act) {
        const int p = pid[t];
        sm.pc[p] = P[t].u; sm.pc[SJ_NPAIR + p] = P[t].gr; sm.pc[2 * SJ_NPAIR + p] = P[t].lt;
        sm.pc[3 * SJ_NPAIR + p] = P[t].R; sm.pc[4 * SJ_NPAIR + p] = P[t].iden; sm.pc[5 * SJ_NPAIR + p] = P[t].ir;
      }
  }
  __syncwarp();
  return acc;
}

// For the box sampler grad f / grad ln D are not maintained by the moves; rebuild them before sampling.
MOLE_D void sj_rebuild_gradients(const SjConst& c, SjLane& L, const SjShared& sm) {
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    const int a = t * 5 + L.gl;
    double gx = 0.0, gy = 0.0, gz = 0.0;
    for (int b = 0; b < 10; ++b) {
      const bool pv = L.val[t] && b != a && sj_slot_valid(c, b);
      const double gr = pv ? sm.pc[SJ_NPAIR + sj_pidx(a, b)] : 0.0;
      gx = fma(gr, L.x[t][0] - sm.xs[b * 3], gx);
      gy = fma(gr, L.x[t][1] - sm.xs[b * 3 + 1], gy);
      gz = fma(gr, L.x[t][2] - sm.xs[b * 3 + 2], gz);
    }
    L.gf[t][0] = gx; L.gf[t][1] = gy; L.gf[t][2] = gz;
  }
}

// Local quantities of the current configuration:
//   kin = -0.5 sum_i lap_i psi / psi,  pot = V (so that E_L = kin + pot),  O[k] = d ln psi / d p_k
template <bool OPT>
MOLE_D void sj_measure(const SjConst& c, const HamParams& h, SjLane& L, const SjShared& sm, double& kin, double& pot,
                       double* O) {
  double kl = 0.0, vl = 0.0, dz[3] = {0.0, 0.0, 0.0};
  const bool want_ion = h.kind == MOLE_OP_IONIC_POT || h.kind == MOLE_OP_IONIC || h.kind == MOLE_OP_ELECTRONIC;
  const bool want_ee = h.kind == MOLE_OP_ELEC_POT || h.kind == MOLE_OP_ELECTRONIC;
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    const int a = t * 5 + L.gl;
    const int n = t == 0 ? c.nup : c.ndn;
    const double* x = L.x[t];
    const double r2 = fma(x[2], x[2], fma(x[1], x[1], x[0] * x[0]));
    const double r = L.val[t] ? sqrt(r2) : 1.0, ir = 1.0 / r;
    const double* e = L.ec[t];
    // lap phi_k
    double lp[5];
    lp[0] = c.z1 * e[0] * (c.z1 - 2.0 * ir);
    lp[1] = (c.z2 * c.z2 * r - 4.0 * c.z2 + 2.0 * ir) * e[1];
    const double cp = e[2] * (c.z3 * c.z3 - 4.0 * c.z3 * ir);
    lp[2] = x[0] * cp; lp[3] = x[1] * cp; lp[4] = x[2] * cp;
    double lapD = 0.0;
#pragma unroll
    for (int k = 0; k < 5; ++k) lapD = fma(k < n ? lp[k] : 0.0, L.minv[t][k], lapD);
    double Lf = 0.0;
    for (int b = 0; b < 10; ++b) {
      const bool pv = b != a && sj_slot_valid(c, b);
      Lf += pv ? sm.pc[2 * SJ_NPAIR + sj_pidx(a, b)] : 0.0;
    }
    const double gg = L.G[t][0] * L.gf[t][0] + L.G[t][1] * L.gf[t][1] + L.G[t][2] * L.gf[t][2];
    const double ff = L.gf[t][0] * L.gf[t][0] + L.gf[t][1] * L.gf[t][1] + L.gf[t][2] * L.gf[t][2];
    const double m = L.val[t] ? 1.0 : 0.0;
    kl = fma(m, lapD + 2.0 * gg + Lf + ff, kl);
    if (want_ion) {                                              // IonicPotential::value, operator.rs:25-36
      double p = 0.0;
      for (int i = 0; i < h.n_ions; ++i) {
        const double dx = x[0] - h.ion_pos[3 * i], dy = x[1] - h.ion_pos[3 * i + 1], dzz = x[2] - h.ion_pos[3 * i + 2];
        p -= h.ion_z[i] / sqrt(fma(dzz, dzz, fma(dy, dy, dx * dx)));
      }
      vl = fma(m, p, vl);
    }
    if (h.kind == MOLE_OP_HARMONIC) vl = fma(m * 0.5 * h.frequency * h.frequency, r2, vl);
    if (OPT) {                                                   // d ln D / d zeta = tr(A^-1 dA)
      const double d0 = -r * e[0], d1 = -r * r * e[1], dp = -r * e[2];
      dz[0] = fma(m * (0 < n ? d0 : 0.0), L.minv[t][0], dz[0]);
      dz[1] = fma(m * (1 < n ? d1 : 0.0), L.minv[t][1], dz[1]);
      dz[2] = fma(m * (2 < n ? dp * x[0] : 0.0), L.minv[t][2], dz[2]);
      dz[2] = fma(m * (3 < n ? dp * x[1] : 0.0), L.minv[t][3], dz[2]);
      dz[2] = fma(m * (4 < n ? dp * x[2] : 0.0), L.minv[t][4], dz[2]);
    }
  }
  // pair sums: V_ee and df/db
  double db[4] = {0.0, 0.0, 0.0, 0.0};
  for (int p = L.gl; p < SJ_NPAIR; p += 5) {
    if (!(sj_slot_valid(c, c_sj_pair_a[p]) && sj_slot_valid(c, c_sj_pair_b[p]))) continue;
    if (want_ee) vl += sm.pc[5 * SJ_NPAIR + p];                  // ElectronicPotential::value, operator.rs:80-90
    if (OPT) {
      const double R = sm.pc[3 * SJ_NPAIR + p], id = sm.pc[4 * SJ_NPAIR + p];
      db[0] = fma(R, id, db[0]);                                 // jastrow.tex:109-119
      db[1] = fma(-c.b1 * R * R, id * id, db[1]);
      db[2] = fma(R, R, db[2]);
      db[3] = fma(R * R, R, db[3]);
    }
  }
  kin = -0.5 * sj_gsum(kl, L);
  pot = sj_gsum(vl, L);
  if (want_ion) pot += h.ionic_repulsion;
  const bool has_kin = h.kind == MOLE_OP_KINETIC || h.kind == MOLE_OP_IONIC || h.kind == MOLE_OP_ELECTRONIC ||
                       h.kind == MOLE_OP_HARMONIC;
  if (!has_kin) kin = 0.0;
  if (h.kind == MOLE_OP_KINETIC) pot = 0.0;
  if (OPT) {
#pragma unroll
    for (int k = 0; k < 3; ++k) O[k] = sj_gsum(dz[k], L);
#pragma unroll
    for (int k = 0; k < 4; ++k) O[3 + k] = sj_gsum(db[k], L);
  }
}

MOLE_D SjConst sj_const(const WfParams& p) {
  SjConst c;
  c.kappa = p.geom[0]; c.ikappa = 1.0 / p.geom[0];
  c.nup = (int)p.geom[1]; c.ndn = (int)p.geom[2];
  c.z1 = p.p[0]; c.z2 = p.p[1]; c.z3 = p.p[2];
  c.b1 = p.p[3]; c.b2 = p.p[4]; c.b3 = p.p[5]; c.b4 = p.p[6];
  return c;
}

MOLE_D void sj_lane_setup(SjLane& L, const SjConst& c, SjShared& sm, double* smem, int64_t w, int64_t W) {
  L.lane = threadIdx.x & 31;
  const int g = L.lane / 5;
  L.gl = L.lane - 5 * g;
  L.base = 5 * g;
  L.act = (g < SJ_WPW) && (w < W);
  L.wr = g < SJ_WPW;
  L.val[0] = L.gl < c.nup;
  L.val[1] = L.gl < c.ndn;
  const int slot = (threadIdx.x >> 5) * SJ_WPW + (g < SJ_WPW ? g : 0);   // idle lanes alias group 0 (loads only)
  double* b = smem + (size_t)slot * SJ_SMEM_PER_WALKER;
  sm.pc = b; sm.gph = b + SJ_PC; sm.xs = sm.gph + SJ_GPH; sm.tsc = sm.xs + SJ_XS; sm.os = sm.tsc + SJ_TSC;
}

MOLE_D void sj_load(SjLane& L, const SjConst& c, const double* x, int64_t w, int64_t W) {
  const int64_t wc = w < W ? w : W - 1;
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    const int cfg_e = t == 0 ? L.gl : c.nup + L.gl;
#pragma unroll
    for (int q = 0; q < 3; ++q)
      L.x[t][q] = L.val[t] ? x[(size_t)(3 * cfg_e + q) * W + wc] : (double)(1 + L.gl + 7 * t + q);   // phantom slots: harmless finite values
  }
}
MOLE_D void sj_store(const SjLane& L, const SjConst& c, double* x, int64_t w, int64_t W) {
  if (!L.act) return;
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    const int cfg_e = t == 0 ? L.gl : c.nup + L.gl;
    if (L.val[t])
#pragma unroll
      for (int q = 0; q < 3; ++q) x[(size_t)(3 * cfg_e + q) * W + w] = L.x[t][q];
  }
}

// ------------------------------------------------------------------ batched evaluation (parity entry point)
__global__ void __launch_bounds__(SJ_THREADS) sj_eval_kernel(const double* __restrict__ x, int64_t W, WfParams p, HamParams h,
                                                              int have_ham, double* psi, double* grad, double* lap,
                                                              double* hpsi, double* pgrad) {
  extern __shared__ double sj_smem[];
  const SjConst c = sj_const(p);
  const int g = (threadIdx.x & 31) / 5;
  const int64_t w = ((int64_t)blockIdx.x * SJ_WARPS + (threadIdx.x >> 5)) * SJ_WPW + g;
  SjLane L;
  SjShared sm;
  sj_lane_setup(L, c, sm, sj_smem, w, W);
  sj_load(L, c, x, w, W);
  sj_init(c, L, sm);
  // kinetic-only pass gives lap psi and O_k; a second pass with the caller's operator gives H psi
  HamParams hk = h;
  hk.kind = MOLE_OP_KINETIC;
  double kin, pot, O[SJ_NP];
  sj_measure<true>(c, hk, L, sm, kin, pot, O);
  double kin_h = 0.0, pot_h = 0.0, O2[SJ_NP];
  if (have_ham) sj_measure<false>(c, h, L, sm, kin_h, pot_h, O2);
  if (!L.act) return;
  if (grad)
#pragma unroll
    for (int t = 0; t < 2; ++t)
      if (L.val[t]) {
        const int cfg_e = t == 0 ? L.gl : c.nup + L.gl;
#pragma unroll
        for (int q = 0; q < 3; ++q) grad[((size_t)w * p.ne + cfg_e) * 3 + q] = L.psi * (L.G[t][q] + L.gf[t][q]);
      }
  if (L.gl == 0) {
    if (psi) psi[w] = L.psi;
    if (lap) lap[w] = -2.0 * kin * L.psi;
    if (hpsi && have_ham) hpsi[w] = kin_h * L.psi + pot_h * L.psi;
    if (pgrad)
#pragma unroll
      for (int k = 0; k < SJ_NP; ++k) pgrad[(size_t)w * SJ_NP + k] = L.psi * O[k];
  }
}

// ------------------------------------------------------------------ fused sweep
template <int S, int METROP>
MOLE_D void sj_sweep_spin(const SjConst& c, SjLane& L, const SjShared& sm, const SweepParams& sp, uint64_t wid,
                          uint32_t step, int s_local, int64_t w, double sd, double* accv) {
  const int n = S == 0 ? c.nup : c.ndn;
  const int cfg_e = S == 0 ? L.gl : c.nup + L.gl;
  // draws of this lane's slot-S electron (keyed by the electron's index in the configuration)
  MoveDraw d;
  if (METROP == MOLE_METROP_BOX) d = mole_draw_uniform4(sp.key, wid, step, DOM_MOVE, (uint32_t)cfg_e);
  else d = mole_draw_normal3_uniform1(sp.key, wid, step, DOM_MOVE, (uint32_t)cfg_e);
  for (int el = 0; el < n; ++el) {                                   // Sampler::move_state, samplers.rs:106-117
    const bool ok = sj_move<S, METROP>(c, L, sm, el, d, sp.metrop_param, sd, sp.compat);
    if (L.act) {
      if (L.gl == 1) accv[1] += ok ? 1.0 : 0.0;                      // ACC_NACC = 6 -> lane 1, idx 1
      if (L.gl == 2) accv[1] += 1.0;                                 // ACC_NMOVE = 7 -> lane 2, idx 1
      if (sp.tr_accept && L.gl == 0)
        sp.tr_accept[((size_t)s_local * (c.nup + c.ndn) + (S == 0 ? el : c.nup + el)) * sp.W + w] = ok ? 1 : 0;
    }
  }
}

template <int METROP, bool OPT>
__global__ void __launch_bounds__(SJ_THREADS, 2) sj_sweep_kernel(const SweepParams sp) {
  extern __shared__ double sj_smem[];
  const SjConst c = sj_const(sp.wf);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane / 5;
  const double sd = sqrt(sp.metrop_param);
  const bool want_e = (sp.observables & MOLE_OBS_ENERGY) != 0;
  const int64_t W = sp.W;
  double accv[SJ_ACC_PER_LANE];
#pragma unroll
  for (int i = 0; i < SJ_ACC_PER_LANE; ++i) accv[i] = 0.0;
  SjLane L;
  SjShared sm;

  const int64_t n_chunks = (W + SJ_WPW - 1) / SJ_WPW;
  for (int64_t chunk = (int64_t)blockIdx.x * SJ_WARPS + warp; chunk < n_chunks; chunk += (int64_t)gridDim.x * SJ_WARPS) {
    const int64_t w = chunk * SJ_WPW + g;
    sj_lane_setup(L, c, sm, sj_smem, w, W);
    sj_load(L, c, sp.x, w, W);
    sj_init(c, L, sm);
    const uint64_t wid = sp.walker_offset + (uint64_t)(w < W ? w : W - 1);
    double blk = L.act ? sp.blk[w] : 0.0;
    int fill = sp.blk_fill;
    for (int s = 0; s < sp.n_sweeps; ++s) {
      const uint32_t step = sp.step0 + (uint32_t)s;
      if (s > 0) sj_refresh(c, L, sm);
      sj_sweep_spin<0, METROP>(c, L, sm, sp, wid, step, s, w, sd, accv);
      sj_sweep_spin<1, METROP>(c, L, sm, sp, wid, step, s, w, sd, accv);
      if (s < sp.n_discard) continue;                                 // montecarlo.rs:36
      const int64_t si = s - sp.n_discard;
      if (METROP == MOLE_METROP_BOX) {
        sj_rebuild_gradients(c, L, sm);
        sj_refresh_spin<0>(c, L, sm);
        sj_refresh_spin<1>(c, L, sm);
      }
      double kin = 0.0, pot = 0.0, O[SJ_NP];
      if (want_e || OPT || (sp.observables & MOLE_OBS_KINETIC)) sj_measure<OPT>(c, sp.ham, L, sm, kin, pot, O);
      const double el = kin + pot;
      double bm = 0.0;
      bool closed = false;
      if (want_e) {
        blk += el;
        if (++fill == sp.block_size) { bm = blk / (double)sp.block_size; blk = 0.0; fill = 0; closed = true; }   // vmc.rs:158-164
      }
      if (L.act) {
        if (want_e) {
          // compact entries 0..4 -> idx 0 of lane gl; 5..9 -> idx 1
          accv[0] += L.gl == 0 ? 1.0 : (L.gl == 1 ? el : (L.gl == 2 ? el * el : (closed ? (L.gl == 3 ? bm : bm * bm) : 0.0)));
          if (L.gl == 0 && closed) accv[1] += 1.0;                    // ACC_NB
          if (L.gl == 3) accv[1] += kin;                              // ACC_T
          if (sp.tr_energy && L.gl == 0) sp.tr_energy[(size_t)si * W + w] = el;
        }
        if (L.gl == 4) accv[1] += L.psi;                              // ACC_PSI
        if (L.gl == 0) {
          if (sp.tr_kinetic && (sp.observables & MOLE_OBS_KINETIC)) sp.tr_kinetic[(size_t)si * W + w] = kin;
          if (sp.tr_wfvalue) sp.tr_wfvalue[(size_t)si * W + w] = L.psi;
        }
      }
      if (OPT) {
        const bool quirk = (sp.compat & MOLE_COMPAT_VECTOR_DIV) != 0;
        if (quirk) {
          // stored sample 1/d_k psi (operator/src/traits.rs:149-150) => O_k = 1/(psi^2 O_k^intended)
#pragma unroll
          for (int k = 0; k < SJ_NP; ++k) O[k] = 1.0 / (L.psi * L.psi * O[k]);
        }
        if (L.act && L.gl == 0) {
#pragma unroll
          for (int k = 0; k < SJ_NP; ++k) {
            sm.os[k] = O[k];
            if (sp.tr_pgrad) sp.tr_pgrad[((size_t)si * SJ_NP + k) * W + w] = L.psi * O[k];   // d_k psi, or 1/d_k psi under the quirk
          }
        }
        __syncwarp();
        if (L.act) {
#pragma unroll
          for (int idx = 2; idx < SJ_ACC_PER_LANE; ++idx) {
            const int j = L.gl + 5 * idx - 10;                        // 0 .. 2P+NOO-1
            if (j < SJ_NP) accv[idx] += sm.os[j];
            else if (j < 2 * SJ_NP) accv[idx] = fma(sm.os[j - SJ_NP], el, accv[idx]);
            else if (j < 2 * SJ_NP + 28) {
              const int q = j - 2 * SJ_NP;
              accv[idx] = fma(sm.os[c_sj_oo_k[q]], sm.os[c_sj_oo_l[q]], accv[idx]);
            }
          }
        }
        __syncwarp();
      }
    }
    sj_store(L, c, sp.x, w, W);
    if (L.act && L.gl == 0) sp.blk[w] = blk;
  }

  // block-tree reduction of the lane-distributed accumulators (compact entry i lives in lanes gl == i%5)
  __syncthreads();
  double* red = sj_smem;                                               // [SJ_WARPS][SJ_NACC]
  __shared__ bool is_last;
  const int gl = lane % 5;
  const bool counted = lane < 5 * SJ_WPW;
  constexpr int LEN = OPT ? SJ_NACC : 10;
#pragma unroll
  for (int i = 0; i < LEN; ++i) {
    double v = (counted && (i % 5) == gl) ? accv[i / 5] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(SJ_FULL, v, o);
    if (lane == 0) red[warp * SJ_NACC + i] = v;
  }
  __syncthreads();
  auto slot = [](int i) {
    if (i < 10) return i;
    if (i < 10 + SJ_NP) return (int)ACC_O + (i - 10);
    if (i < 10 + 2 * SJ_NP) return (int)ACC_OE + (i - 10 - SJ_NP);
    return (int)ACC_OO + (i - 10 - 2 * SJ_NP);
  };
  if (threadIdx.x < LEN) {
    double s = 0.0;
    for (int q = 0; q < SJ_WARPS; ++q) s += red[q * SJ_NACC + threadIdx.x];
    sp.partials[(size_t)blockIdx.x * ACC_LEN + slot(threadIdx.x)] = s;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = (atomicInc(sp.ticket, gridDim.x - 1) == gridDim.x - 1);
  __syncthreads();
  if (is_last) {
    __threadfence();
    if (threadIdx.x < LEN) {
      const int sl = slot(threadIdx.x);
      double s = 0.0;
      for (unsigned b = 0; b < gridDim.x; ++b) s += sp.partials[(size_t)b * ACC_LEN + sl];
      sp.acc[sl] += s;
    }
  }
}

// ------------------------------------------------------------------ DMC time step (dmc.rs:87-130)
__global__ void __launch_bounds__(SJ_THREADS, 2) sj_dmc_kernel(const DmcParams dp) {
  extern __shared__ double sj_smem[];
  const SjConst c = sj_const(dp.wf);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane / 5;
  const double sd = sqrt(dp.tau_move);
  const int64_t W = dp.W;
  double s_we = 0.0, s_w = 0.0, s_wn = 0.0, m_wn = 0.0;
  SjLane L;
  SjShared sm;
  const int64_t n_chunks = (W + SJ_WPW - 1) / SJ_WPW;
  for (int64_t chunk = (int64_t)blockIdx.x * SJ_WARPS + warp; chunk < n_chunks; chunk += (int64_t)gridDim.x * SJ_WARPS) {
    const int64_t w = chunk * SJ_WPW + g;
    sj_lane_setup(L, c, sm, sj_smem, w, W);
    sj_load(L, c, dp.x, w, W);
    sj_init(c, L, sm);
    const uint64_t wid = dp.walker_offset + (uint64_t)(w < W ? w : W - 1);
    double kin, pot, O[SJ_NP];
    double e_old;
    if (dp.el_cached) e_old = L.act ? dp.el[w] : 0.0;
    else { sj_measure<false>(c, dp.ham, L, sm, kin, pot, O); e_old = kin + pot; }
#pragma unroll
    for (int S = 0; S < 2; ++S) {
      const int n = S == 0 ? c.nup : c.ndn;
      const int cfg_e = S == 0 ? L.gl : c.nup + L.gl;
      const MoveDraw d = mole_draw_normal3_uniform1(dp.key, wid, dp.step, DOM_MOVE, (uint32_t)cfg_e);
      for (int el = 0; el < n; ++el) {
        if (S == 0) sj_move<0, MOLE_METROP_DIFFUSE>(c, L, sm, el, d, dp.tau_move, sd, dp.compat);
        else sj_move<1, MOLE_METROP_DIFFUSE>(c, L, sm, el, d, dp.tau_move, sd, dp.compat);
      }
    }
    sj_refresh(c, L, sm);
    sj_measure<false>(c, dp.ham, L, sm, kin, pot, O);
    const double e_new = kin + pot;
    if (L.act && L.gl == 0) {
      const double wt = dp.w[w];
      s_we = fma(wt, e_old, s_we);
      s_w += wt;
      const double wn = wt * exp(-dp.tau_weight * ((e_old + e_new) / 2.0 - dp.e_ref));
      s_wn += wn;
      m_wn = fmax(m_wn, wn);
      dp.w[w] = wn;
      dp.el[w] = e_new;
    }
    sj_store(L, c, dp.x, w, W);
  }
  __shared__ double smr[SJ_WARPS][4];
  __shared__ bool is_last;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s_we += __shfl_xor_sync(SJ_FULL, s_we, o);
    s_w += __shfl_xor_sync(SJ_FULL, s_w, o);
    s_wn += __shfl_xor_sync(SJ_FULL, s_wn, o);
    m_wn = fmax(m_wn, __shfl_xor_sync(SJ_FULL, m_wn, o));
  }
  if (lane == 0) { smr[warp][0] = s_we; smr[warp][1] = s_w; smr[warp][2] = s_wn; smr[warp][3] = m_wn; }
  __syncthreads();
  if (threadIdx.x < 4) {
    double s = 0.0;
    for (int q = 0; q < SJ_WARPS; ++q) s = (threadIdx.x == 3) ? fmax(s, smr[q][3]) : s + smr[q][threadIdx.x];
    dp.partials[(size_t)blockIdx.x * 4 + threadIdx.x] = s;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = (atomicInc(dp.ticket, gridDim.x - 1) == gridDim.x - 1);
  __syncthreads();
  if (is_last) {
    __threadfence();
    if (threadIdx.x < 4) {
      double s = 0.0;
      for (unsigned b = 0; b < gridDim.x; ++b)
        s = (threadIdx.x == 3) ? fmax(s, dp.partials[(size_t)b * 4 + 3]) : s + dp.partials[(size_t)b * 4 + threadIdx.x];
      dp.red[threadIdx.x] = s;
    }
  }
}

// ------------------------------------------------------------------ host launchers
static inline cudaError_t sj_upload_tables() {
  static bool done_dev[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  bool& done = done_dev[dev & 63];
  if (done) return cudaSuccess;
  signed char pa[SJ_NPAIR], pb[SJ_NPAIR], ok[28], ol[28];
  int p = 0;
  for (int a = 0; a < 10; ++a)
    for (int b = a + 1; b < 10; ++b, ++p) { pa[p] = (signed char)a; pb[p] = (signed char)b; }
  int q = 0;
  for (int k = 0; k < SJ_NP; ++k)
    for (int l = k; l < SJ_NP; ++l, ++q) { ok[q] = (signed char)k; ol[q] = (signed char)l; }
  cudaError_t e;
  if ((e = cudaMemcpyToSymbol(c_sj_pair_a, pa, sizeof(pa))) != cudaSuccess) return e;
  if ((e = cudaMemcpyToSymbol(c_sj_pair_b, pb, sizeof(pb))) != cudaSuccess) return e;
  if ((e = cudaMemcpyToSymbol(c_sj_oo_k, ok, sizeof(ok))) != cudaSuccess) return e;
  if ((e = cudaMemcpyToSymbol(c_sj_oo_l, ol, sizeof(ol))) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(sj_eval_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SJ_SMEM_BYTES)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(sj_dmc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SJ_SMEM_BYTES)) != cudaSuccess) return e;
#define SJ_ATTR(M, O) \
  if ((e = cudaFuncSetAttribute(sj_sweep_kernel<M, O>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SJ_SMEM_BYTES)) != cudaSuccess) return e;
  SJ_ATTR(MOLE_METROP_BOX, false) SJ_ATTR(MOLE_METROP_BOX, true) SJ_ATTR(MOLE_METROP_DIFFUSE, false) SJ_ATTR(MOLE_METROP_DIFFUSE, true)
#undef SJ_ATTR
  done = true;
  return cudaSuccess;
}

static inline int sj_grid(mole_ctx_s* ctx, int64_t W, int rows) {
  const int64_t ctas = (W + SJ_WPB - 1) / SJ_WPB;
  const int64_t cap = std::min<int64_t>(rows, (int64_t)ctx->sm_count * 2);
  return (int)std::max<int64_t>(1, std::min<int64_t>(ctas, cap));
}

static inline cudaError_t sj_eval_launch(cudaStream_t st, const double* x, int64_t W, const WfParams& wp, const HamParams& h,
                                         int have_ham, double* psi, double* grad, double* lap, double* hpsi, double* pgrad) {
  cudaError_t e = sj_upload_tables();
  if (e != cudaSuccess) return e;
  const int blocks = (int)((W + SJ_WPB - 1) / SJ_WPB);
  sj_eval_kernel<<<blocks, SJ_THREADS, SJ_SMEM_BYTES, st>>>(x, W, wp, h, have_ham, psi, grad, lap, hpsi, pgrad);
  return cudaSuccess;
}

static inline int32_t sj_sweep_launch(mole_ctx_s* ctx, mole_ens_s* e, const SweepParams& sp, int metrop, bool opt) {
  cudaError_t ce = sj_upload_tables();
  if (ce != cudaSuccess) return mole_set_error(ctx, MOLE_ERR_CUDA, std::string("sj tables: ") + cudaGetErrorString(ce));
  const int blocks = sj_grid(ctx, e->W, e->partial_rows);
  cudaStream_t st = (cudaStream_t)ctx->stream;
  if (metrop == MOLE_METROP_BOX) {
    if (opt) sj_sweep_kernel<MOLE_METROP_BOX, true><<<blocks, SJ_THREADS, SJ_SMEM_BYTES, st>>>(sp);
    else sj_sweep_kernel<MOLE_METROP_BOX, false><<<blocks, SJ_THREADS, SJ_SMEM_BYTES, st>>>(sp);
  } else {
    if (opt) sj_sweep_kernel<MOLE_METROP_DIFFUSE, true><<<blocks, SJ_THREADS, SJ_SMEM_BYTES, st>>>(sp);
    else sj_sweep_kernel<MOLE_METROP_DIFFUSE, false><<<blocks, SJ_THREADS, SJ_SMEM_BYTES, st>>>(sp);
  }
  return MOLE_OK;
}

static inline int32_t sj_dmc_launch(mole_ctx_s* ctx, mole_ens_s* e, const DmcParams& dp) {
  cudaError_t ce = sj_upload_tables();
  if (ce != cudaSuccess) return mole_set_error(ctx, MOLE_ERR_CUDA, std::string("sj tables: ") + cudaGetErrorString(ce));
  const int blocks = sj_grid(ctx, e->W, e->partial_rows);
  sj_dmc_kernel<<<blocks, SJ_THREADS, SJ_SMEM_BYTES, (cudaStream_t)ctx->stream>>>(dp);
  return MOLE_OK;
}
