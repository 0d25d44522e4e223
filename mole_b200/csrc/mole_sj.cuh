// Slater-Jastrow kind (MOLE_WF_SLATER_JASTROW, SURVEY.md §8(c) synthetic config 5):
//   psi = det[phi_k(r_i)]_up * det[phi_k(r_i)]_dn * exp(f_ee),   orbitals 1s, 2s, 2p_{x,y,z} (STO),
//   f_ee = sum_{i<j} u(R_ij), u(R) = b1 R/(1+b2 R) + b3 R^2 + b4 R^3, R = (1-exp(-kappa r))/kappa
//   (theory/jastrow.tex:23-31 with the erratum of SURVEY.md §8(c)).
//
// Layout: FIVE lanes cooperate on one walker, six walkers per warp (lanes 30/31 idle).  Lane gl
// owns one electron of each spin.  Registers hold the positions, the radial cache (r, 1/r and the three
// orbital exponentials) and grad f of the two own electrons; shared memory holds, per walker, the pair
// cache (u, g/r, laplacian term, 1/r, R for the 45 pairs), both inverse Slater matrices, a copy of
// the positions, a mailbox for the intra-walker exchanges and the 42 optimisation moments.  The spin being moved always sits in
// register slot 0 (the slots are swapped between the two halves of a sweep), so there is one copy of
// the move code.  A single-electron move re-evaluates only what changed: one orbital row, a
// Sherman-Morrison update of one 5x5 inverse (rebuilt from scratch every SJ_REFRESH_EVERY sweeps), nine
// Jastrow pairs; grad ln D of the own electrons is carried in registers and updated on accepted moves.
// All reductions over the five lanes go through the mailbox in a fixed order (deterministic).
// The per-walker shared-memory stride is 5 (mod 16) doubles, so the unit-stride-in-lane accesses of a
// warp's 30 active lanes fall on distinct 8-byte banks (for the two idle lanes see sj_lane_setup).
// Metropolis semantics are the reference's (src/metropolis/src/metrop.rs:60-96,150-212), including
// the Frobenius norm over ALL electrons' drift in t_high / t_low.
#pragma once
#include "mole_internal.h"
#if !defined(MOLE_EMU)
#include <cuda_runtime.h>
#endif
#include "mole_rng.cuh"
#include "mole_math.cuh"

constexpr int SJ_WPW = 6;                      // walkers per warp
#ifndef MOLE_SJ_IDLE_SHIFT
#define MOLE_SJ_IDLE_SHIFT 0    // see sj_lane_setup
#endif
#ifndef MOLE_SJ_WARPS
#define MOLE_SJ_WARPS 4
#endif
#ifndef MOLE_SJ_MIN_CTAS
#define MOLE_SJ_MIN_CTAS 2
#endif
constexpr int SJ_WARPS = MOLE_SJ_WARPS;
constexpr int SJ_THREADS = 32 * SJ_WARPS;
constexpr int SJ_WPB = SJ_WPW * SJ_WARPS;       // walkers per CTA
constexpr int SJ_MIN_CTAS = MOLE_SJ_MIN_CTAS;   // CTAs per SM the register budget is tuned for
// ptxas derives the register cap of __launch_bounds__(threads, ctas) as if the CTA had a multiple of four
// warps; -DMOLE_SJ_MAXNREG=n states the cap directly (e.g. 3 warps x 3 CTAs at 224 registers = 9 warps/SM)
#ifdef MOLE_SJ_MAXNREG
#define SJ_BOUNDS __maxnreg__(MOLE_SJ_MAXNREG)
#else
#define SJ_BOUNDS __launch_bounds__(SJ_THREADS, SJ_MIN_CTAS)
#endif
constexpr int SJ_NPAIR = 45;
constexpr int SJ_PCV = 5;                       // cached values per pair: u, g/r, lap term, 1/r, R
// shared memory per walker (offsets in doubles)
constexpr int SJ_OFF_PC = 0;                               // [SJ_PCV][45]
constexpr int SJ_OFF_MINV = SJ_OFF_PC + SJ_PCV * SJ_NPAIR; // [2][5][5] inverse Slater matrices, (spin, k, j)
constexpr int SJ_OFF_MB = SJ_OFF_MINV + 50;                // mailbox, 56 doubles
constexpr int SJ_MB = 56;
constexpr int SJ_OFF_ACC = SJ_OFF_MB + SJ_MB;              // [42] sum O_k, sum O_k E_L, sum O_k O_l of this walker slot
constexpr int SJ_NMOM = 42;
constexpr int SJ_STRIDE = 373;                             // >= SJ_OFF_ACC + SJ_NMOM and == 5 (mod 16)
static_assert(SJ_STRIDE >= SJ_OFF_ACC + SJ_NMOM && SJ_STRIDE % 16 == 5, "per-walker stride");
#ifndef MOLE_SJ_REFRESH_EVERY
#define MOLE_SJ_REFRESH_EVERY 16
#endif
constexpr int SJ_REFRESH_EVERY = MOLE_SJ_REFRESH_EVERY;    // sweeps between from-scratch rebuilds of the inverses
constexpr size_t SJ_SMEM_BYTES = (size_t)SJ_STRIDE * SJ_WPB * sizeof(double);
// mailbox slots
constexpr int MB_XN = 0;      // [0..2] trial position, [3] accept uniform, [4..6] old position of the moved electron
constexpr int MB_E = 8;       // [3] orbital exponentials at the trial position
constexpr int MB_RIN = 12;    // [6][5] reduction inputs (also the 5x5 transpose scratch, 25 doubles)
// measurement (sj_measure): [9][5] per-lane partial sums at 0..44, results at MB_OUT
constexpr int MB_OUT = 45;    // [0..6] O_k, [7] 1.0, [8] E_L, [9..10] spare: every moment is a product out[k] * out[l]
constexpr int SJ_NP = 7;
constexpr int SJ_NACC = 10 + 2 * SJ_NP + SJ_NP * (SJ_NP + 1) / 2;   // 52 compact accumulator entries
constexpr int SJ_ACC_PER_LANE = (SJ_NACC + 4) / 5;                   // 11
constexpr unsigned SJ_FULL = 0xffffffffu;

struct SjConst {
  int nup, ndn;
  double kappa, ikappa, z1, z2, z3, b1, b2, b3, b4;
};

// operand indices (k << 4 | l) into the MB_OUT block of the 42 per-walker optimisation moments:
// sum O_k = out[k] * 1, sum O_k E_L = out[k] * E_L, sum O_k O_l (k <= l, packed by rows).  Shared memory,
// because the index is lane-divergent (a divergent index into the constant bank serialises).
__shared__ unsigned char s_sj_mom[48];
__device__ __forceinline__ void sj_mom_table_init() {
  for (int j = threadIdx.x; j < 48; j += SJ_THREADS) {
    int k = 7, l = 7;                                    // padding entries: 1.0 * 1.0, never accumulated
    if (j < SJ_NP) { k = j; l = 7; }
    else if (j < 2 * SJ_NP) { k = j - SJ_NP; l = 8; }
    else if (j < SJ_NMOM) {
      int q = j - 2 * SJ_NP;
      k = 0;
      while (q >= SJ_NP - k) { q -= SJ_NP - k; ++k; }
      l = k + q;
    }
    s_sj_mom[j] = (unsigned char)(k << 4 | l);
  }
  __syncthreads();
}

MOLE_D int sj_pidx(int a, int b) {              // unordered pair of slot ids (0..9) -> 0..44
  const int lo = a < b ? a : b, hi = a < b ? b : a;
  return lo * (19 - lo) / 2 + (hi - lo - 1);
}

struct SjLane {
  int lane, gl, base;
  int ph;              // spin held in register slot 0 (slot t holds spin t ^ ph)
  bool act;            // this 5-lane group holds a real walker
  bool wr;             // lanes 30/31 alias group 0's shared memory and must never store to it
  bool val[2];         // slot validity (lane index < number of electrons of that spin)
  double x[2][3];      // own electrons
  double orb[2][5];    // r, 1/r, exp(-z1 r), exp(-z2 r), exp(-z3 r) at own electrons
  double gf[2][3];     // grad_i f
  double G[2][3];      // grad_i ln D
  double psi, fj;      // psi = L.psi * exp(L.fj): accepted moves multiply psi by the determinant ratio and add the
                       // Jastrow change to fj; sj_fold_psi / sj_refresh bring fj back to 0.  Replicated over the group
  double* sm;          // this walker's shared-memory region
};

MOLE_D int sj_spin_n(const SjConst& c, int spin) { return spin == 0 ? c.nup : c.ndn; }
MOLE_D bool sj_slot_valid(const SjConst& c, int sid) { return sid < 5 ? sid < c.nup : (sid - 5) < c.ndn; }
#if defined(MOLE_EMU)
MOLE_D void sj_sync(int line = MOLE_CALLER_LINE) { __syncwarp(0xffffffffu, line); }   // the watchdog reports the caller
#else
MOLE_D void sj_sync() { __syncwarp(); }
#endif

// acceptance.min(1.0) with the NaN policy of include/mole_b200.h (MOLE_COMPAT_NAN_ACCEPT)
MOLE_D double sj_clamp_acceptance(double a, uint32_t compat) {
  if (isnan(a)) return (compat & MOLE_COMPAT_NAN_ACCEPT) ? 1.0 : 0.0;
  return fmin(a, 1.0);
}

// sum over the five lanes of a group through the mailbox, lane order 0..4, result replicated
MOLE_D double sj_gsum(double v, const SjLane& L) {
  if (L.wr) L.sm[SJ_OFF_MB + MB_RIN + L.gl] = v;
  sj_sync();
  double s = L.sm[SJ_OFF_MB + MB_RIN];
#pragma unroll
  for (int i = 1; i < 5; ++i) s += L.sm[SJ_OFF_MB + MB_RIN + i];
  sj_sync();
  return s;
}

struct SjPair { double u, gr, lt, ir, R; };

// pair function from the squared distance (theory/jastrow.tex:23-31,45-48,68-71,82-97)
MOLE_D SjPair sj_pair(const SjConst& c, double r2) {
  SjPair o;
  const double r = m_sqrt_rsqrt(r2, o.ir);
  const double E = m_exp(-c.kappa * r);
  o.R = (1.0 - E) * c.ikappa;
  const double iden = m_rcp(fma(c.b2, o.R, 1.0));
  const double R2 = o.R * o.R;
  o.u = fma(R2, fma(c.b4, o.R, c.b3), (c.b1 * o.R) * iden);
  const double id2 = iden * iden;
  const double du = fma(c.b1, id2, fma(3.0 * c.b4, R2, 2.0 * c.b3 * o.R));
  const double d2u = fma(-2.0 * c.b1 * c.b2, id2 * iden, fma(6.0 * c.b4, o.R, 2.0 * c.b3));
  const double g = E * du;
  o.gr = g * o.ir;
  o.lt = fma(2.0, o.gr, fma(E * E, d2u, -c.kappa * g));   // div(rhat g) = 2 g/r + dg/dr
  return o;
}

// orbital values from the radial cache; entries k >= n are zero (identity padding of the Slater matrix)
MOLE_D void sj_phi(const double* x, const double* o, int n, double* phi) {
  phi[0] = n > 0 ? o[2] : 0.0;
  phi[1] = n > 1 ? o[0] * o[3] : 0.0;
  phi[2] = n > 2 ? x[0] * o[4] : 0.0;
  phi[3] = n > 3 ? x[1] * o[4] : 0.0;
  phi[4] = n > 4 ? x[2] * o[4] : 0.0;
}

// grad_i ln D = sum_k grad phi_k(r_i) Minv[k][i] from the radial cache and the electron's column m[]:
//   grad phi_0 = -z1 e1 x/r,  grad phi_1 = (1 - z2 r) e2 x/r,  grad phi_{2+a} = e3 (e_a - z3 x_a x/r)
MOLE_D void sj_gradlnD(const SjConst& c, const double* x, const double* o, const double* m, double* G) {
  // s = (A - B xm) / r with A, B independent of xm, evaluated as fma(-B/r, xm, A/r): one dependent op after xm
  const double xm = fma(x[2], m[4], fma(x[1], m[3], x[0] * m[2]));
  const double A = fma(-c.z1 * o[2], m[0], (1.0 - c.z2 * o[0]) * o[3] * m[1]);
  const double s = fma(-(c.z3 * o[4]) * o[1], xm, A * o[1]);
  G[0] = fma(s, x[0], o[4] * m[2]);
  G[1] = fma(s, x[1], o[4] * m[3]);
  G[2] = fma(s, x[2], o[4] * m[4]);
}

MOLE_D void sj_radial(const SjConst& c, const double* x, bool valid, double* o) {
  const double r2 = fma(x[2], x[2], fma(x[1], x[1], x[0] * x[0]));
  o[0] = m_sqrt_rsqrt(valid ? r2 : 1.0, o[1]);
  o[2] = m_exp(-c.z1 * o[0]);
  o[3] = m_exp(-c.z2 * o[0]);
  o[4] = m_exp(-c.z3 * o[0]);
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
    const double ipv = m_rcp(pv);
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

// rebuild the inverse Slater matrix of register slot t from scratch; returns the determinant
MOLE_D double sj_refresh_slot(const SjConst& c, SjLane& L, int t) {
  const int spin = t ^ L.ph;
  const int n = sj_spin_n(c, spin);
  double phi[5];
  sj_phi(L.x[t], L.orb[t], n, phi);
  double* tsc = L.sm + SJ_OFF_MB + MB_RIN;
#pragma unroll
  for (int k = 0; k < 5; ++k)
    if (L.wr) tsc[L.gl * 5 + k] = L.val[t] ? phi[k] : (k == L.gl ? 1.0 : 0.0);   // A[i=gl][k], identity padding
  sj_sync();
  double M[5], out[5], det;
#pragma unroll
  for (int i = 0; i < 5; ++i) M[i] = tsc[i * 5 + L.gl];   // row gl of A^T
  sj_sync();
  sj_invert(M, out, det, L);
#pragma unroll
  for (int k = 0; k < 5; ++k)
    if (L.wr) L.sm[SJ_OFF_MINV + spin * 25 + k * 5 + L.gl] = out[k];          // Minv[k][j = gl]
  sj_gradlnD(c, L.x[t], L.orb[t], out, L.G[t]);
  sj_sync();
  return det;
}

// once per sweep: both inverses from scratch (bounds the Sherman-Morrison round-off), psi re-derived
MOLE_D void sj_refresh(const SjConst& c, SjLane& L) {
  const double d0 = sj_refresh_slot(c, L, 0);
  const double d1 = sj_refresh_slot(c, L, 1);
  double fl = 0.0;
#pragma unroll
  for (int i = 0; i < 9; ++i) fl += L.sm[SJ_OFF_PC + L.gl + 5 * i];   // cache slots of absent pairs hold zeros
  L.psi = d0 * d1 * m_exp(sj_gsum(fl, L));
  L.fj = 0.0;
}

// full initialisation of the cooperative state from the positions in L.x (slot t = spin t)
MOLE_D void sj_init(const SjConst& c, SjLane& L) {
  L.ph = 0;
  for (int p = L.gl; p < SJ_PCV * SJ_NPAIR; p += 5)          // cache slots of absent pairs must hold finite values
    if (L.wr) L.sm[SJ_OFF_PC + p] = 0.0;
  double* const xs = L.sm + SJ_OFF_MB;                       // all 10 positions, staged in the mailbox for the pair loop
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    sj_radial(c, L.x[t], L.val[t], L.orb[t]);
#pragma unroll
    for (int q = 0; q < 3; ++q)
      if (L.wr) xs[(t * 5 + L.gl) * 3 + q] = L.x[t][q];
  }
  sj_sync();
  // Jastrow from scratch: every lane sums over the partners of its own electrons
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    const int a = t * 5 + L.gl;
    double gx = 0.0, gy = 0.0, gz = 0.0;
    for (int b = 0; b < 10; ++b) {
      const bool pv = L.val[t] && b != a && sj_slot_valid(c, b);
      const double* xb = xs + b * 3;
      const double dx = L.x[t][0] - xb[0], dy = L.x[t][1] - xb[1], dz = L.x[t][2] - xb[2];
      const double r2 = pv ? fma(dz, dz, fma(dy, dy, dx * dx)) : 1.0;
      const SjPair P = sj_pair(c, r2);
      if (pv) {
        gx = fma(P.gr, dx, gx); gy = fma(P.gr, dy, gy); gz = fma(P.gr, dz, gz);
        if (a < b && L.wr) {
          double* pc = L.sm + SJ_OFF_PC + sj_pidx(a, b);
          pc[0] = P.u; pc[SJ_NPAIR] = P.gr; pc[2 * SJ_NPAIR] = P.lt; pc[3 * SJ_NPAIR] = P.ir; pc[4 * SJ_NPAIR] = P.R;
        }
      }
    }
    L.gf[t][0] = gx; L.gf[t][1] = gy; L.gf[t][2] = gz;
  }
  sj_sync();
  sj_refresh(c, L);
}

// psi of the current configuration from the carried determinant part and Jastrow change
MOLE_D void sj_fold_psi(SjLane& L) {
  L.psi *= m_exp(L.fj);
  L.fj = 0.0;
}

// exchange the two register slots (the spin to be moved must sit in slot 0)
MOLE_D void sj_swap_slots(SjLane& L) {
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    double t = L.x[0][q]; L.x[0][q] = L.x[1][q]; L.x[1][q] = t;
    t = L.gf[0][q]; L.gf[0][q] = L.gf[1][q]; L.gf[1][q] = t;
    t = L.G[0][q]; L.G[0][q] = L.G[1][q]; L.G[1][q] = t;
  }
#pragma unroll
  for (int q = 0; q < 5; ++q) { const double t = L.orb[0][q]; L.orb[0][q] = L.orb[1][q]; L.orb[1][q] = t; }
  const bool v = L.val[0]; L.val[0] = L.val[1]; L.val[1] = v;
  L.ph ^= 1;
}

// Metropolis::move_state for electron `el` of the spin in slot 0.  d = the pre-generated draws of THIS
// lane's slot-0 electron (only the owner's are used).  Returns the accept decision (uniform over the group).

#include "mole_sj_move.cuh"

// Local quantities of the current configuration:
//   kin = -0.5 sum_i lap_i psi / psi,  pot = V (so that E_L = kin + pot), replicated over the group; with OPT the
//   mailbox block MB_OUT holds O[k] = d ln psi / d p_k (k < 7), 1.0 and E_L on return (visible to all lanes).
// Gout (optional): grad ln D of the two own electrons.
// Straight-line: the nine pairs of this lane (gl, gl+5, ..) are summed without validity tests (cache slots of
// absent pairs hold zeros, so they contribute nothing), their reciprocals go through one 9-wide batch, and the
// nine group reductions share two warp syncs: every lane stores its partials as rows of the mailbox, every lane
// sums rows 0 and 1 (kinetic, potential), lane gl sums rows 2+gl and 7+gl into MB_OUT.
MOLE_D double sj_row5(const double* r) { return ((r[0] + r[1]) + (r[2] + r[3])) + r[4]; }   // fixed order, lanes 0..4

template <bool OPT>
MOLE_D void sj_measure(const SjConst& c, const HamParams& h, SjLane& L, double& kin, double& pot, uint32_t compat = 0,
                       double (*Gout)[3] = nullptr) {
  double kl = 0.0, vl = 0.0, dz[3] = {0.0, 0.0, 0.0};
  const bool want_ion = h.kind == MOLE_OP_IONIC_POT || h.kind == MOLE_OP_IONIC || h.kind == MOLE_OP_ELECTRONIC;
  const bool want_ee = h.kind == MOLE_OP_ELEC_POT || h.kind == MOLE_OP_ELECTRONIC;
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    const int spin = t ^ L.ph;
    const int n = sj_spin_n(c, spin);
    const double* x = L.x[t];
    const double* o = L.orb[t];
    const double r = o[0], ir = o[1];
    double m[5];
    const double* G = L.G[t];
#pragma unroll
    for (int k = 0; k < 5; ++k) m[k] = L.sm[SJ_OFF_MINV + spin * 25 + k * 5 + L.gl];
    if (Gout) { Gout[t][0] = G[0]; Gout[t][1] = G[1]; Gout[t][2] = G[2]; }
    // lap phi_k
    const double cp = o[4] * (c.z3 * c.z3 - 4.0 * c.z3 * ir);
    double lapD = (0 < n ? c.z1 * o[2] * (c.z1 - 2.0 * ir) : 0.0) * m[0];
    lapD = fma(1 < n ? (c.z2 * c.z2 * r - 4.0 * c.z2 + 2.0 * ir) * o[3] : 0.0, m[1], lapD);
    const double lapP = fma(4 < n ? x[2] * cp : 0.0, m[4], fma(3 < n ? x[1] * cp : 0.0, m[3], (2 < n ? x[0] * cp : 0.0) * m[2]));
    const double gg = G[0] * L.gf[t][0] + G[1] * L.gf[t][1] + G[2] * L.gf[t][2];
    const double ff = L.gf[t][0] * L.gf[t][0] + L.gf[t][1] * L.gf[t][1] + L.gf[t][2] * L.gf[t][2];
    const double mk = L.val[t] ? 1.0 : 0.0;
    kl = fma(mk, (lapD + lapP) + (2.0 * gg + ff), kl);
    if (want_ion) {                                              // IonicPotential::value, operator.rs:25-36
      double p = 0.0;
      for (int i = 0; i < h.n_ions; ++i) {
        const double dx = x[0] - h.ion_pos[3 * i], dy = x[1] - h.ion_pos[3 * i + 1], dzz = x[2] - h.ion_pos[3 * i + 2];
        p = fma(-h.ion_z[i], m_rsqrt(fma(dzz, dzz, fma(dy, dy, dx * dx))), p);
      }
      vl = fma(mk, p, vl);
    }
    if (h.kind == MOLE_OP_HARMONIC) vl = fma(mk * 0.5 * h.frequency * h.frequency, r * r, vl);
    if (OPT) {                                                   // d ln D / d zeta = tr(A^-1 dA)
      const double d0 = -r * o[2], d1 = -r * r * o[3], dp = -r * o[4];
      dz[0] = fma(mk * (0 < n ? d0 : 0.0), m[0], dz[0]);
      dz[1] = fma(mk * (1 < n ? d1 : 0.0), m[1], dz[1]);
      const double s2 = fma(4 < n ? x[2] : 0.0, m[4], fma(3 < n ? x[1] : 0.0, m[3], (2 < n ? x[0] : 0.0) * m[2]));
      dz[2] = fma(mk * dp, s2, dz[2]);
    }
  }
  // pair sums: sum_i lap_i f = 2 sum_pairs div(rhat g), V_ee = sum 1/r (ElectronicPotential::value,
  // operator.rs:80-90) and df/db (jastrow.tex:109-119)
  const double* pc = L.sm + SJ_OFF_PC + L.gl;
  double lt[9], pir[9], R[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    lt[i] = pc[2 * SJ_NPAIR + 5 * i];
    pir[i] = pc[3 * SJ_NPAIR + 5 * i];
    R[i] = pc[4 * SJ_NPAIR + 5 * i];
  }
  const double lts = (((lt[0] + lt[1]) + (lt[2] + lt[3])) + ((lt[4] + lt[5]) + (lt[6] + lt[7]))) + lt[8];
  kl = fma(2.0, lts, kl);
  if (want_ee) vl += (((pir[0] + pir[1]) + (pir[2] + pir[3])) + ((pir[4] + pir[5]) + (pir[6] + pir[7]))) + pir[8];
  double db[4] = {0.0, 0.0, 0.0, 0.0};
  if (OPT) {
    double den[9], id[9], a0[3] = {0.0, 0.0, 0.0}, a1[3] = {0.0, 0.0, 0.0}, a2[3] = {0.0, 0.0, 0.0}, a3[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int i = 0; i < 9; ++i) den[i] = fma(c.b2, R[i], 1.0);
    m_rcp_n<9>(den, id);
#pragma unroll
    for (int i = 0; i < 9; ++i) {                                // three interleaved partial sums per quantity
      const double q = R[i] * id[i], R2 = R[i] * R[i];
      a0[i % 3] += q;
      a1[i % 3] = fma(q, q, a1[i % 3]);
      a2[i % 3] += R2;
      a3[i % 3] = fma(R2, R[i], a3[i % 3]);
    }
    db[0] = (a0[0] + a0[1]) + a0[2];
    db[1] = -c.b1 * ((a1[0] + a1[1]) + a1[2]);
    db[2] = (a2[0] + a2[1]) + a2[2];
    db[3] = (a3[0] + a3[1]) + a3[2];
  }
  double* const mb = L.sm + SJ_OFF_MB;
  if (L.wr) {
    mb[L.gl] = kl;
    mb[5 + L.gl] = vl;
    if (OPT) {
#pragma unroll
      for (int k = 0; k < 3; ++k) mb[10 + 5 * k + L.gl] = dz[k];
#pragma unroll
      for (int k = 0; k < 4; ++k) mb[25 + 5 * k + L.gl] = db[k];
    }
  }
  sj_sync();
  kin = -0.5 * sj_row5(mb);
  pot = sj_row5(mb + 5);
  if (want_ion) pot += h.ionic_repulsion;
  const bool has_kin = h.kind == MOLE_OP_KINETIC || h.kind == MOLE_OP_IONIC || h.kind == MOLE_OP_ELECTRONIC ||
                       h.kind == MOLE_OP_HARMONIC;
  if (!has_kin) kin = 0.0;
  if (h.kind == MOLE_OP_KINETIC) pot = 0.0;
  if (OPT) {
    double oa = sj_row5(mb + 10 + 5 * L.gl);                       // O[gl]
    double ob = sj_row5(mb + 35 + 5 * (L.gl < 2 ? L.gl : 0));      // O[5 + gl], lanes 0 and 1
    if (compat & MOLE_COMPAT_VECTOR_DIV) {
      // stored sample 1/d_k psi (operator/src/traits.rs:149-150) => O_k = 1/(psi^2 O_k^intended)
      oa = 1.0 / (L.psi * L.psi * oa);
      ob = 1.0 / (L.psi * L.psi * ob);
    }
    if (L.wr) {
      mb[MB_OUT + L.gl] = oa;
      if (L.gl < 2) mb[MB_OUT + 5 + L.gl] = ob;
      if (L.gl == 2) mb[MB_OUT + 7] = 1.0;
      if (L.gl == 3) mb[MB_OUT + 8] = kin + pot;
    }
  }
  sj_sync();
}

MOLE_D SjConst sj_const(const WfParams& p) {
  SjConst c;
  c.kappa = p.geom[0]; c.ikappa = 1.0 / p.geom[0];
  c.nup = (int)p.geom[1]; c.ndn = (int)p.geom[2];
  c.z1 = p.p[0]; c.z2 = p.p[1]; c.z3 = p.p[2];
  c.b1 = p.p[3]; c.b2 = p.p[4]; c.b3 = p.p[5]; c.b4 = p.p[6];
  return c;
}

MOLE_D void sj_lane_setup(SjLane& L, const SjConst& c, double* smem, int64_t w, int64_t W) {
  L.lane = threadIdx.x & 31;
  const int g = L.lane / 5;
  L.gl = L.lane - 5 * g;
  L.base = 5 * g;
  L.ph = 0;
  L.act = (g < SJ_WPW) && (w < W);
  L.wr = g < SJ_WPW;
  L.val[0] = L.gl < c.nup;
  L.val[1] = L.gl < c.ndn;
  // idle lanes (30, 31) only ever load and alias walker slot 0 of this warp.  That costs one shared-memory conflict
  // wavefront on every unit-stride access of the second half-warp (l1tex__data_bank_conflicts 2.2e9 per launch, VERDICT
  // r1 weak#3iv).  Round 2 measured the cure: a window shifted by 14 doubles (MOLE_SJ_IDLE_SHIFT=14) puts the two lanes
  // on the free bank pairs 14/15 and removes the conflicts, but is not faster (63.1 vs 62.3 ms): the LSU pipe is 26 % busy
  // and the conflicts hide behind the FP64 dependency stalls - and lanes that compute on garbage must then be kept out of
  // every data-dependent branch (unguarded they drag the warp into the rare accept branch on every move: 79.7 ms).
  const int slot = (threadIdx.x >> 5) * SJ_WPW + (g < SJ_WPW ? g : 0);
  L.sm = smem + (size_t)slot * SJ_STRIDE + (g < SJ_WPW ? 0 : MOLE_SJ_IDLE_SHIFT);
}

// global <-> registers; slot t must hold spin t (L.ph == 0)
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

// one sweep: N_e single-electron moves, spin up then spin down (Sampler::move_state, samplers.rs:106-117)
// returns the number of accepted moves (uniform over the group); tr_accept is non-null only in the lane that
// writes the trace (so that nothing about the walker's identity is tested per move)
template <int METROP>
MOLE_D int sj_sweep_moves(const SjConst& c, SjLane& L, RngKey key, uint64_t wid, uint32_t step, double param, double sd,
                          double inv2tau, uint32_t compat, uint8_t* tr_accept, size_t tr_stride) {
  int n_acc = 0;
#pragma unroll 1
  for (int half = 0; half < 2; ++half) {
    const int spin = L.ph;
    const int n = sj_spin_n(c, spin);
    const int cfg_e = spin == 0 ? L.gl : c.nup + L.gl;
    // draws of this lane's slot-0 electron, keyed by the electron's index in the configuration
    MoveDraw d;
    if (METROP == MOLE_METROP_BOX) d = mole_draw_uniform4(key, wid, step, DOM_MOVE, (uint32_t)cfg_e);
    else d = mole_draw_normal3_uniform1(key, wid, step, DOM_MOVE, (uint32_t)cfg_e);
#pragma unroll 1
    for (int el = 0; el < n; ++el) {
      const bool ok = sj_move<METROP>(c, L, el, d, param, sd, inv2tau, compat);
      n_acc += ok ? 1 : 0;
      if (tr_accept) tr_accept[(size_t)(spin == 0 ? el : c.nup + el) * tr_stride] = ok ? 1 : 0;
    }
    sj_swap_slots(L);
  }
  return n_acc;
}

// ------------------------------------------------------------------ batched evaluation (parity entry point)
__global__ void __launch_bounds__(SJ_THREADS) sj_eval_kernel(const double* __restrict__ x, int64_t W, WfParams p, HamParams h,
                                                              int have_ham, double* psi, double* grad, double* lap,
                                                              double* hpsi, double* pgrad) {
  MOLE_DYN_SMEM(double, sj_smem);
  mole_math_smem_init();
  const SjConst c = sj_const(p);
  const int g = (threadIdx.x & 31) / 5;
  const int64_t w = ((int64_t)blockIdx.x * SJ_WARPS + (threadIdx.x >> 5)) * SJ_WPW + g;
  SjLane L;
  sj_lane_setup(L, c, sj_smem, w, W);
  sj_load(L, c, x, w, W);
  sj_init(c, L);
  // kinetic-only pass gives lap psi and O_k; a second pass with the caller's operator gives H psi
  HamParams hk = h;
  hk.kind = MOLE_OP_KINETIC;
  double kin, pot, O[SJ_NP], G[2][3];
  sj_measure<true>(c, hk, L, kin, pot, 0, G);
#pragma unroll
  for (int k = 0; k < SJ_NP; ++k) O[k] = L.sm[SJ_OFF_MB + MB_OUT + k];
  double kin_h = 0.0, pot_h = 0.0;
  if (have_ham) sj_measure<false>(c, h, L, kin_h, pot_h);
  if (!L.act) return;
  if (grad)
#pragma unroll
    for (int t = 0; t < 2; ++t)
      if (L.val[t]) {
        const int cfg_e = t == 0 ? L.gl : c.nup + L.gl;
#pragma unroll
        for (int q = 0; q < 3; ++q) grad[((size_t)w * p.ne + cfg_e) * 3 + q] = L.psi * (G[t][q] + L.gf[t][q]);
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
template <int METROP, bool OPT>
__global__ void SJ_BOUNDS sj_sweep_kernel(const SweepParams sp) {
  MOLE_DYN_SMEM(double, sj_smem);
  mole_math_smem_init();
  if (OPT) sj_mom_table_init();
  const SjConst c = sj_const(sp.wf);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane / 5;
  const double sd = sqrt(sp.metrop_param), inv2tau = 1.0 / (2.0 * sp.metrop_param);
  const bool want_e = (sp.observables & MOLE_OBS_ENERGY) != 0;
  const int64_t W = sp.W;
  const int ne = c.nup + c.ndn;
  // the ten scalar accumulators live in two registers per lane (compact entry i in lanes gl == i%5);
  // the 42 optimisation moments live in this walker slot's shared memory
  double accv[2] = {0.0, 0.0};
  SjLane L;
  if (OPT) {
    double* am = sj_smem + (size_t)(warp * SJ_WPW + (g < SJ_WPW ? g : 0)) * SJ_STRIDE + SJ_OFF_ACC;
    if (g < SJ_WPW)
      for (int j = lane - 5 * g; j < SJ_NMOM; j += 5) am[j] = 0.0;
    __syncwarp();
  }

  const int64_t n_chunks = (W + SJ_WPW - 1) / SJ_WPW;
  for (int64_t chunk = (int64_t)blockIdx.x * SJ_WARPS + warp; chunk < n_chunks; chunk += (int64_t)gridDim.x * SJ_WARPS) {
    const int64_t w = chunk * SJ_WPW + g;
    sj_lane_setup(L, c, sj_smem, w, W);
    sj_load(L, c, sp.x, w, W);
    sj_init(c, L);
    const uint64_t wid = sp.walker_offset + (uint64_t)(w < W ? w : W - 1);
    double blk = L.act ? sp.blk[w] : 0.0;
    int fill = sp.blk_fill, nbad = 0;   // nbad: skipped samples of the open block (not carried across launches)
#pragma unroll 1
    for (int s = 0; s < sp.n_sweeps; ++s) {
      const uint32_t step = sp.step0 + (uint32_t)s;
      if (s > 0 && (s % SJ_REFRESH_EVERY) == 0) sj_refresh(c, L);
      uint8_t* tra = (sp.tr_accept && L.act && L.gl == 0) ? sp.tr_accept + (size_t)s * ne * W + w : nullptr;   // lane 0 of a real walker
      const int n_acc = sj_sweep_moves<METROP>(c, L, sp.key, wid, step, sp.metrop_param, sd, inv2tau, sp.compat, tra, (size_t)W);
      if (L.act) {
        if (L.gl == 1) accv[1] += (double)n_acc;                      // ACC_NACC = 6 -> lane 1, idx 1
        if (L.gl == 2) accv[1] += (double)ne;                         // ACC_NMOVE = 7 -> lane 2, idx 1
      }
      if (s < sp.n_discard) continue;                                 // montecarlo.rs:36
      const int64_t si = s - sp.n_discard;
      double kin = 0.0, pot = 0.0;
      sj_fold_psi(L);
      if (want_e || OPT || (sp.observables & MOLE_OBS_KINETIC)) sj_measure<OPT>(c, sp.ham, L, kin, pot, sp.compat);
      const double el = kin + pot;
      // a sample whose E_L or O_k is not finite goes to the traces as it is, but stays out of every sum and is
      // counted in acc[ACC_BAD] (mole_ensemble_health); the decision is uniform over the walker's five lanes
      bool bad = !isfinite(el);
      if (OPT) {
        const double* os = L.sm + SJ_OFF_MB + MB_OUT;
        const bool mine = !isfinite(os[L.gl]) || (L.gl < 2 && !isfinite(os[5 + L.gl]));
        const unsigned votes = __ballot_sync(SJ_FULL, mine);        // unconditionally: every lane of the warp must vote
        bad = bad || ((votes >> L.base) & 31u) != 0u;
      }
      const bool good = L.act && !bad;
      if (bad && L.act && L.gl == 0) atomicAdd(sp.acc + ACC_BAD, 1.0);
      double bm = 0.0;
      bool closed = false;
      if (want_e) {
        if (bad) ++nbad; else blk += el;
        if (++fill == sp.block_size) {                                 // vmc.rs:158-164
          closed = nbad < sp.block_size;
          bm = blk / (double)(sp.block_size - nbad);
          blk = 0.0; fill = 0; nbad = 0;
        }
      }
      if (L.act) {
        if (want_e) {
          // compact entries 0..4 -> idx 0 of lane gl; 5..9 -> idx 1
          const double sv = L.gl == 0 ? 1.0 : (L.gl == 1 ? el : el * el);
          accv[0] += L.gl < 3 ? (good ? sv : 0.0) : (closed ? (L.gl == 3 ? bm : bm * bm) : 0.0);
          if (L.gl == 0 && closed) accv[1] += 1.0;                    // ACC_NB
          if (L.gl == 3 && good) accv[1] += kin;                      // ACC_T
          if (sp.tr_energy && L.gl == 0) sp.tr_energy[(size_t)si * W + w] = el;
        }
        if (L.gl == 4 && good) accv[1] += L.psi;                      // ACC_PSI
        if (L.gl == 0) {
          if (sp.tr_kinetic && (sp.observables & MOLE_OBS_KINETIC)) sp.tr_kinetic[(size_t)si * W + w] = kin;
          if (sp.tr_wfvalue) sp.tr_wfvalue[(size_t)si * W + w] = L.psi;
        }
      }
      if (OPT) {
        // sum O_k, sum O_k E_L, sum O_k O_l: entry j of this walker slot is out[k_j] * out[l_j]; lane gl owns j == gl (mod 5)
        const double* os = L.sm + SJ_OFF_MB + MB_OUT;
        if (sp.tr_pgrad && L.act && L.gl == 0) {
#pragma unroll
          for (int k = 0; k < SJ_NP; ++k) sp.tr_pgrad[((size_t)si * SJ_NP + k) * W + w] = L.psi * os[k];   // d_k psi, or 1/d_k psi under the quirk
        }
        if (good) {
          // all loads first: the nine read-modify-writes would otherwise serialise (the compiler cannot tell that
          // the stores to am[] do not alias the os[] loads of the next entry)
          double* am = L.sm + SJ_OFF_ACC;
          double fa[9], fb[9], fm[9];
#pragma unroll
          for (int i = 0; i < 9; ++i) {
            const int j = L.gl + 5 * i;
            const int jj = (i < 8 || j < SJ_NMOM) ? j : L.gl;      // entries 42, 43 do not exist (lanes 2..4, i = 8)
            const unsigned id = s_sj_mom[jj];
            fa[i] = os[id >> 4]; fb[i] = os[id & 15]; fm[i] = am[jj];
          }
#pragma unroll
          for (int i = 0; i < 9; ++i) fm[i] = fma(fa[i], fb[i], fm[i]);
#pragma unroll
          for (int i = 0; i < 9; ++i) {
            const int j = L.gl + 5 * i;
            if (i < 8 || j < SJ_NMOM) am[j] = fm[i];
          }
        }
        sj_sync();
      }
    }
    sj_store(L, c, sp.x, w, W);
    if (L.act && L.gl == 0) sp.blk[w] = blk;
  }

  // block-tree reduction of the lane-distributed accumulators (compact entry i lives in lanes gl == i%5)
  __syncthreads();
  double* red = sj_smem;                                               // [SJ_WARPS][16] scalars (pair-cache area of slot 0)
  __shared__ bool is_last;
  const int gl = lane % 5;
  const bool counted = lane < 5 * SJ_WPW;
  constexpr int LEN = OPT ? SJ_NACC : 10;
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    double v = (counted && (i % 5) == gl) ? accv[i / 5] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(SJ_FULL, v, o);
    if (lane == 0) red[warp * 16 + i] = v;
  }
  __syncthreads();
  auto slot = [](int i) {
    if (i < 10) return i;
    if (i < 10 + SJ_NP) return (int)ACC_O + (i - 10);
    if (i < 10 + 2 * SJ_NP) return (int)ACC_OE + (i - 10 - SJ_NP);
    return (int)ACC_OO + (i - 10 - 2 * SJ_NP);
  };
  for (int i = threadIdx.x; i < LEN; i += SJ_THREADS) {                 // (a one-warp CTA has fewer threads than entries)
    double s = 0.0;
    if (i < 10) {
      for (int q = 0; q < SJ_WARPS; ++q) s += red[q * 16 + i];
    } else {
      for (int q = 0; q < SJ_WPB; ++q) s += sj_smem[(size_t)q * SJ_STRIDE + SJ_OFF_ACC + (i - 10)];
    }
    sp.partials[(size_t)blockIdx.x * ACC_LEN + slot(i)] = s;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = (atomicInc(sp.ticket, gridDim.x - 1) == gridDim.x - 1);
  __syncthreads();
  if (is_last) {
    __threadfence();
    for (int i = threadIdx.x; i < LEN; i += SJ_THREADS) {
      const int sl = slot(i);
      double s = 0.0;
      for (unsigned b = 0; b < gridDim.x; ++b) s += sp.partials[(size_t)b * ACC_LEN + sl];
      sp.acc[sl] += s;
    }
  }
}

// ------------------------------------------------------------------ DMC time step (dmc.rs:87-130)
__global__ void SJ_BOUNDS sj_dmc_kernel(const DmcParams dp) {
  MOLE_DYN_SMEM(double, sj_smem);
  mole_math_smem_init();
  const SjConst c = sj_const(dp.wf);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane / 5;
  const double sd = sqrt(dp.tau_move), inv2tau = 1.0 / (2.0 * dp.tau_move);
  const int64_t W = dp.W;
  double s_we = 0.0, s_w = 0.0, s_wn = 0.0, m_wn = 0.0;
  SjLane L;
  const int64_t n_chunks = (W + SJ_WPW - 1) / SJ_WPW;
  for (int64_t chunk = (int64_t)blockIdx.x * SJ_WARPS + warp; chunk < n_chunks; chunk += (int64_t)gridDim.x * SJ_WARPS) {
    const int64_t w = chunk * SJ_WPW + g;
    sj_lane_setup(L, c, sj_smem, w, W);
    sj_load(L, c, dp.x, w, W);
    sj_init(c, L);
    const uint64_t wid = dp.walker_offset + (uint64_t)(w < W ? w : W - 1);
    double kin, pot;
    double e_old;
    if (dp.el_cached) e_old = L.act ? dp.el[w] : 0.0;
    else { sj_measure<false>(c, dp.ham, L, kin, pot); e_old = kin + pot; }
    sj_sweep_moves<MOLE_METROP_DIFFUSE>(c, L, dp.key, wid, dp.step, dp.tau_move, sd, inv2tau, dp.compat, nullptr, 0);
    sj_refresh(c, L);
    sj_measure<false>(c, dp.ham, L, kin, pot);
    const double e_new = kin + pot;
    if (L.act && L.gl == 0) {
      const double w_in = dp.w[w];
      const double w_up = w_in * exp(-dp.tau_weight * ((e_old + e_new) / 2.0 - dp.e_ref));
      const bool bad = !isfinite(e_old) || !isfinite(e_new) || !isfinite(w_up);   // counted, and the walker dies (see dmc_step_kernel)
      if (bad) atomicAdd(dp.health, 1.0);
      const double wt = bad ? 0.0 : w_in;
      s_we = fma(wt, bad ? 0.0 : e_old, s_we);
      s_w += wt;
      const double wn = bad ? 0.0 : w_up;
      s_wn += wn;
      m_wn = fmax(m_wn, wn);
      dp.w[w] = wn;
      dp.el[w] = bad ? 0.0 : e_new;
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
    mole_dmc_fold_partials(dp.partials, gridDim.x, dp.red, smr);
  }
}

// ------------------------------------------------------------------ host launchers
#if !defined(MOLE_EMU)
// dynamic shared-memory opt-in of every SJ kernel on the CURRENT device (mole_ctx_create)
static inline cudaError_t sj_set_kernel_attributes() {
  cudaError_t e;
  if ((e = cudaFuncSetAttribute(sj_eval_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SJ_SMEM_BYTES)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(sj_dmc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SJ_SMEM_BYTES)) != cudaSuccess) return e;
#define SJ_ATTR(M, O) \
  if ((e = cudaFuncSetAttribute(sj_sweep_kernel<M, O>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SJ_SMEM_BYTES)) != cudaSuccess) return e;
  SJ_ATTR(MOLE_METROP_BOX, false) SJ_ATTR(MOLE_METROP_BOX, true) SJ_ATTR(MOLE_METROP_DIFFUSE, false) SJ_ATTR(MOLE_METROP_DIFFUSE, true)
#undef SJ_ATTR
  return cudaSuccess;
}

static inline int sj_grid(mole_ctx_s* ctx, int64_t W, int rows) {
  const int64_t ctas = (W + SJ_WPB - 1) / SJ_WPB;
  const int64_t cap = std::min<int64_t>(rows, (int64_t)ctx->sm_count * SJ_MIN_CTAS);
  return (int)std::max<int64_t>(1, std::min<int64_t>(ctas, cap));
}

static inline cudaError_t sj_eval_launch(cudaStream_t st, const double* x, int64_t W, const WfParams& wp, const HamParams& h,
                                         int have_ham, double* psi, double* grad, double* lap, double* hpsi, double* pgrad) {
  const int blocks = (int)((W + SJ_WPB - 1) / SJ_WPB);
  sj_eval_kernel<<<blocks, SJ_THREADS, SJ_SMEM_BYTES, st>>>(x, W, wp, h, have_ham, psi, grad, lap, hpsi, pgrad);
  return cudaSuccess;
}

static inline int32_t sj_sweep_launch(mole_ctx_s* ctx, mole_ens_s* e, const SweepParams& sp, int metrop, bool opt) {
  // launch errors are picked up by the caller's KERNEL_CHECK (cudaGetLastError) right after this returns
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
  const int blocks = sj_grid(ctx, e->W, e->partial_rows);
  sj_dmc_kernel<<<blocks, SJ_THREADS, SJ_SMEM_BYTES, (cudaStream_t)ctx->stream>>>(dp);
  return MOLE_OK;
}
#endif  // !MOLE_EMU
