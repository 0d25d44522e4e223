// Branch-free, batched fp64 elementary functions for the hot kernels.
//
// Why: (1) CUDA's exp(), sqrt() and '/' each carry a rarely-taken slow path whose branch splits the
// basic block; (2) ptxas schedules a Horner / Newton chain as one serial run of DFMAs, and a dependent
// DFMA issues only every ~9.4 cycles on B200 (measured, tools/ubench/dfma.cu) while the FP64 pipe
// accepts one per 2 cycles per SM sub-partition.  With the 3 warps per scheduler the Slater-Jastrow
// kernel affords, a serial chain leaves the pipe ~75% idle.  The functions below are straight-line
// code, take N independent arguments at once and are written stage-major, so the N chains (and the
// Estrin sub-terms inside exp) are adjacent independent instructions.
// Accuracy ~1-2 ulp (not correctly rounded), far inside the 1e-10 parity bar; checked on the device
// by tests/test_math_device.py.
#pragma once
#include "mole_internal.h"

#if defined(MOLE_DEVICE_CODE)

// MUFU seeds (20 mantissa bits)
#if defined(MOLE_EMU)
MOLE_D double m_seed_rcp(double x) { return mole_emu_rcp_seed(x); }
MOLE_D double m_seed_rsqrt(double x) { return mole_emu_rsqrt_seed(x); }
#else
MOLE_D double m_seed_rcp(double x) { double y; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x)); return y; }
MOLE_D double m_seed_rsqrt(double x) { double y; asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x)); return y; }
#endif

// 1/x for finite normal x != 0.  The MUFU.RCP64H seed carries 20 mantissa bits (the low word is zero), so
// ONE cubic Newton step y (1 + e + e^2), e = 1 - x y exact in the fma, leaves a truncation error e^3 < 2^-57
// and a single final rounding: <= 0.6 ulp.  (-DMOLE_NEWTON2 adds the former second, quadratic step.)
template <int N>
MOLE_D void m_rcp_n(const double (&x)[N], double (&y)[N]) {
  double e[N];
#pragma unroll
  for (int i = 0; i < N; ++i) y[i] = m_seed_rcp(x[i]);
#pragma unroll
  for (int i = 0; i < N; ++i) e[i] = fma(-x[i], y[i], 1.0);
#pragma unroll
  for (int i = 0; i < N; ++i) e[i] = fma(e[i], e[i], e[i]);
#pragma unroll
  for (int i = 0; i < N; ++i) y[i] = fma(y[i], e[i], y[i]);          // y (1 + e + e^2)
#ifdef MOLE_NEWTON2
#pragma unroll
  for (int i = 0; i < N; ++i) e[i] = fma(-x[i], y[i], 1.0);
#pragma unroll
  for (int i = 0; i < N; ++i) y[i] = fma(y[i], e[i], y[i]);
#endif
}
MOLE_D double m_rcp(double x) {
  const double a[1] = {x};
  double y[1];
  m_rcp_n<1>(a, y);
  return y[0];
}

// 1/sqrt(x) for finite normal x > 0.  MUFU.RSQ64H seed (20 mantissa bits), then ONE cubic Newton step
// y (1 + e + 1.5 e^2) with e = (1 - x y^2)/2, written on d = 2e (no halving of x): truncation 2.5 e^3 < 2^-56;
// the rounding of x y inside d and the final rounding give ~1 ulp.  (-DMOLE_NEWTON2 adds the former second, quadratic step: ~0.6 ulp.)
template <int N>
MOLE_D void m_rsqrt_n(const double (&x)[N], double (&y)[N]) {
  double hx[N], e[N];
#pragma unroll
  for (int i = 0; i < N; ++i) y[i] = m_seed_rsqrt(x[i]);
  // same step on d = 1 - x y^2 = 2e: y + (y d)(1/2 + 3/8 d), five FP64 instructions instead of six
#pragma unroll
  for (int i = 0; i < N; ++i) hx[i] = x[i] * y[i];
#pragma unroll
  for (int i = 0; i < N; ++i) e[i] = fma(-hx[i], y[i], 1.0);
#pragma unroll
  for (int i = 0; i < N; ++i) y[i] = fma(y[i] * e[i], fma(0.375, e[i], 0.5), y[i]);
#ifdef MOLE_NEWTON2
#pragma unroll
  for (int i = 0; i < N; ++i) e[i] = fma(-(0.5 * x[i]) * y[i], y[i], 0.5);
#pragma unroll
  for (int i = 0; i < N; ++i) y[i] = fma(y[i], e[i], y[i]);
#endif
}
MOLE_D double m_rsqrt(double x) {
  const double a[1] = {x};
  double y[1];
  m_rsqrt_n<1>(a, y);
  return y[0];
}

// sqrt(x) and 1/sqrt(x) together, x > 0 finite normal
template <int N>
MOLE_D void m_sqrt_rsqrt_n(const double (&x)[N], double (&s)[N], double (&rinv)[N]) {
  m_rsqrt_n<N>(x, rinv);
#pragma unroll
  for (int i = 0; i < N; ++i) s[i] = x[i] * rinv[i];
#ifdef MOLE_SQRT_CORR   // one Heron correction: <= 0.6 ulp instead of <= 2 ulp, two more dependent FMAs per value
#pragma unroll
  for (int i = 0; i < N; ++i) s[i] = fma(fma(-s[i], s[i], x[i]), 0.5 * rinv[i], s[i]);
#endif
}
MOLE_D double m_sqrt_rsqrt(double x, double& rinv) {
  const double a[1] = {x};
  double s[1], r[1];
  m_sqrt_rsqrt_n<1>(a, s, r);
  rinv = r[0];
  return s[0];
}

// exp(x) on the whole real line, branch-free: 0 below -745.2, +inf above 709.8, NaN propagates.
// exp(r) = 1 + r + r^2 q(r) on |r| <= ln2/2 with q a degree-9 near-minimax polynomial (max relative
// error 1.6e-17) evaluated in Estrin form (dependent depth 5 instead of 10).
// The argument is clamped to [-1400, 710] with data selects (NOT a select on the result: the compiler
// turns that into a branch around the whole evaluation, which serialises a batch again); the two-step
// 2^n scaling then underflows to 0 / overflows to +inf by itself.  NEG_ONLY: caller guarantees x <= 0.
// Constants live in the constant bank so that DFMA takes them as c[bank][offset] operands instead of
// two IMAD.MOV/UMOV per 64-bit literal (15% of the executed instructions of the v3 kernel).
//   [0] log2 e  [1] 1.5*2^52  [2] -ln2 hi (32 trailing zero bits)  [3] -ln2 lo  [4..13] q coefficients c0..c9
__constant__ double c_mexp[14] = {
    1.4426950408889634, 6755399441055744.0, -6.93147180369123816490e-01, -1.90821492927058770002e-10,
    0.5000000000000001, 0.16666666666666669, 0.04166666666662413, 0.008333333333330062, 0.0013888888917213717,
    0.00019841269863053618, 2.4801521295954376e-05, 2.7557268459997064e-06, 2.7620088445409746e-07, 2.510038549551032e-08};

#if !defined(MOLE_EXP_POLY) && !defined(MOLE_EXP_TABLE)
#define MOLE_EXP_TABLE 1   // default; -DMOLE_EXP_POLY selects the table-free polynomial below (A/B: SJ sweep 84.4 vs 80.4 ms)
#endif
#ifdef MOLE_EXP_TABLE
// Table-driven variant: exp(x) = 2^k 2^(j/64) e^r with n = round(64 x / ln 2) = 64 k + j, |r| <= ln2/128, so a
// degree-5 polynomial suffices (r^6/720 < 3.5e-17): 12 FP64 instructions per value instead of 20 and a
// dependent depth of 8 instead of 11.  2^(j/64) comes from a 64-entry table in shared memory (a lane-divergent
// index into the constant bank would serialise); every kernel that evaluates exp calls mole_math_smem_init().
//   [0] 64 log2 e  [1] 1.5*2^52  [2] -ln2/64 hi  [3] -ln2/64 lo  [4..7] 1/2, 1/6, 1/24, 1/120
__constant__ double c_mexpt[8] = {92.33248261689366, 6755399441055744.0, -0x1.62e42fee00000p-7, -0x1.a39ef35793c76p-39,
                                  0.5, 0.16666666666666666, 0.041666666666666664, 0.008333333333333333};
__constant__ double c_exp_tab[64] = {
    0x1.0000000000000p+0, 0x1.02c9a3e778061p+0, 0x1.059b0d3158574p+0, 0x1.0874518759bc8p+0,
    0x1.0b5586cf9890fp+0, 0x1.0e3ec32d3d1a2p+0, 0x1.11301d0125b51p+0, 0x1.1429aaea92de0p+0,
    0x1.172b83c7d517bp+0, 0x1.1a35beb6fcb75p+0, 0x1.1d4873168b9aap+0, 0x1.2063b88628cd6p+0,
    0x1.2387a6e756238p+0, 0x1.26b4565e27cddp+0, 0x1.29e9df51fdee1p+0, 0x1.2d285a6e4030bp+0,
    0x1.306fe0a31b715p+0, 0x1.33c08b26416ffp+0, 0x1.371a7373aa9cbp+0, 0x1.3a7db34e59ff7p+0,
    0x1.3dea64c123422p+0, 0x1.4160a21f72e2ap+0, 0x1.44e086061892dp+0, 0x1.486a2b5c13cd0p+0,
    0x1.4bfdad5362a27p+0, 0x1.4f9b2769d2ca7p+0, 0x1.5342b569d4f82p+0, 0x1.56f4736b527dap+0,
    0x1.5ab07dd485429p+0, 0x1.5e76f15ad2148p+0, 0x1.6247eb03a5585p+0, 0x1.6623882552225p+0,
    0x1.6a09e667f3bcdp+0, 0x1.6dfb23c651a2fp+0, 0x1.71f75e8ec5f74p+0, 0x1.75feb564267c9p+0,
    0x1.7a11473eb0187p+0, 0x1.7e2f336cf4e62p+0, 0x1.82589994cce13p+0, 0x1.868d99b4492edp+0,
    0x1.8ace5422aa0dbp+0, 0x1.8f1ae99157736p+0, 0x1.93737b0cdc5e5p+0, 0x1.97d829fde4e50p+0,
    0x1.9c49182a3f090p+0, 0x1.a0c667b5de565p+0, 0x1.a5503b23e255dp+0, 0x1.a9e6b5579fdbfp+0,
    0x1.ae89f995ad3adp+0, 0x1.b33a2b84f15fbp+0, 0x1.b7f76f2fb5e47p+0, 0x1.bcc1e904bc1d2p+0,
    0x1.c199bdd85529cp+0, 0x1.c67f12e57d14bp+0, 0x1.cb720dcef9069p+0, 0x1.d072d4a07897cp+0,
    0x1.d5818dcfba487p+0, 0x1.da9e603db3285p+0, 0x1.dfc97337b9b5fp+0, 0x1.e502ee78b3ff6p+0,
    0x1.ea4afa2a490dap+0, 0x1.efa1bee615a27p+0, 0x1.f50765b6e4540p+0, 0x1.fa7c1819e90d8p+0};
__shared__ double s_exp_tab[64];
MOLE_D void mole_math_smem_init() {
  for (int i = threadIdx.x; i < 64; i += blockDim.x) s_exp_tab[i] = c_exp_tab[i];
  __syncthreads();
}

template <int N, bool NEG_ONLY = false>
MOLE_D void m_exp_n(const double (&xin)[N], double (&y)[N]) {
  double x[N], t[N], r[N], r2[N], b0[N], b1[N], T[N];
  int n[N];
#pragma unroll
  for (int i = 0; i < N; ++i) {
    x[i] = xin[i] < -1400.0 ? -1400.0 : xin[i];
    if (!NEG_ONLY) x[i] = x[i] > 710.0 ? 710.0 : x[i];
  }
#pragma unroll
  for (int i = 0; i < N; ++i) t[i] = fma(x[i], c_mexpt[0], c_mexpt[1]);                    // low word = round(64 x log2 e)
#pragma unroll
  for (int i = 0; i < N; ++i) { n[i] = __double2loint(t[i]); t[i] -= c_mexpt[1]; }
#pragma unroll
  for (int i = 0; i < N; ++i) T[i] = s_exp_tab[n[i] & 63];
#pragma unroll
  for (int i = 0; i < N; ++i) r[i] = fma(t[i], c_mexpt[2], x[i]);
#pragma unroll
  for (int i = 0; i < N; ++i) r[i] = fma(t[i], c_mexpt[3], r[i]);
#pragma unroll
  for (int i = 0; i < N; ++i) {
    r2[i] = r[i] * r[i];
    b0[i] = fma(c_mexpt[5], r[i], c_mexpt[4]);
    b1[i] = fma(c_mexpt[7], r[i], c_mexpt[6]);
  }
#pragma unroll
  for (int i = 0; i < N; ++i) b0[i] = fma(b1[i], r2[i], b0[i]);
#pragma unroll
  for (int i = 0; i < N; ++i) b0[i] = fma(r2[i], b0[i], r[i]);                             // e^r - 1
  // two-step 2^k scaling (covers denormals / overflow); the first factor is folded into the table value, off
  // the polynomial's dependency chain
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const int k = n[i] >> 6;
    const int n1 = k >> 1, n2 = k - n1;
    const double s1 = __hiloint2double((n1 + 1023) << 20, 0), s2 = __hiloint2double((n2 + 1023) << 20, 0);
    const double Ts = T[i] * s1;
    y[i] = fma(Ts, b0[i], Ts) * s2;
  }
}
#else
MOLE_D void mole_math_smem_init() {}
template <int N, bool NEG_ONLY = false>
MOLE_D void m_exp_n(const double (&xin)[N], double (&y)[N]) {
  double x[N], t[N], r[N], r2[N], a0[N], a1[N], a2[N], a3[N], a4[N];
  int n[N];
#pragma unroll
  for (int i = 0; i < N; ++i) {
    x[i] = xin[i] < -1400.0 ? -1400.0 : xin[i];
    if (!NEG_ONLY) x[i] = x[i] > 710.0 ? 710.0 : x[i];
  }
#pragma unroll
  for (int i = 0; i < N; ++i) t[i] = fma(x[i], c_mexp[0], c_mexp[1]);                      // low word = round(x log2 e)
#pragma unroll
  for (int i = 0; i < N; ++i) { n[i] = __double2loint(t[i]); t[i] -= c_mexp[1]; }
#pragma unroll
  for (int i = 0; i < N; ++i) r[i] = fma(t[i], c_mexp[2], x[i]);
#pragma unroll
  for (int i = 0; i < N; ++i) r[i] = fma(t[i], c_mexp[3], r[i]);
#pragma unroll
  for (int i = 0; i < N; ++i) {
    r2[i] = r[i] * r[i];
    a0[i] = fma(c_mexp[5], r[i], c_mexp[4]);
    a1[i] = fma(c_mexp[7], r[i], c_mexp[6]);
    a2[i] = fma(c_mexp[9], r[i], c_mexp[8]);
    a3[i] = fma(c_mexp[11], r[i], c_mexp[10]);
    a4[i] = fma(c_mexp[13], r[i], c_mexp[12]);
  }
#pragma unroll
  for (int i = 0; i < N; ++i) {
    a0[i] = fma(a1[i], r2[i], a0[i]);
    a2[i] = fma(a3[i], r2[i], a2[i]);
    t[i] = r2[i] * r2[i];                                                                  // r^4
  }
#pragma unroll
  for (int i = 0; i < N; ++i) {
    a0[i] = fma(a2[i], t[i], a0[i]);
    t[i] = t[i] * t[i];                                                                    // r^8
  }
#pragma unroll
  for (int i = 0; i < N; ++i) a0[i] = fma(a4[i], t[i], a0[i]);
#pragma unroll
  for (int i = 0; i < N; ++i) a0[i] = fma(r2[i], a0[i], r[i]) + 1.0;
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const int n1 = n[i] >> 1, n2 = n[i] - n1;                                              // two-step scaling covers denormals / overflow
    const double s1 = __hiloint2double((n1 + 1023) << 20, 0), s2 = __hiloint2double((n2 + 1023) << 20, 0);
    y[i] = (a0[i] * s1) * s2;
  }
}
#endif
MOLE_D double m_exp(double x) {
  const double a[1] = {x};
  double y[1];
  m_exp_n<1>(a, y);
  return y[0];
}

// ln(x) for finite normal x > 0.  x = m 2^e, m in [sqrt(1/2), sqrt 2); ln m = 2 atanh(s), s = (m-1)/(m+1),
// atanh(s)/s = 1 + z P(z), z = s^2 <= 0.0295, P degree 8 (max relative error 5e-19 before rounding).
//   [0] ln2 hi  [1] ln2 lo  [2..10] P coefficients
__constant__ double c_mlog[11] = {6.93147180369123816490e-01, 1.90821492927058770002e-10,
                                  0.3333333333333333, 0.19999999999999996, 0.14285714285717704, 0.11111111109922116,
                                  0.09090909297884532, 0.07692287483230453, 0.06667822920974689, 0.05843981498771078,
                                  0.0594172619705393};
template <int N>
MOLE_D void m_log_n(const double (&x)[N], double (&y)[N]) {
  double m[N], ef[N], den[N], rden[N], s[N], z[N], p[N];
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const int hi = __double2hiint(x[i]), lo = __double2loint(x[i]);
    int e = (hi >> 20) - 1023;
    m[i] = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, lo);
    const bool big = m[i] > 1.4142135623730951;
    m[i] = big ? 0.5 * m[i] : m[i];
    e += big ? 1 : 0;
    ef[i] = (double)e;
    den[i] = m[i] + 1.0;
  }
  m_rcp_n<N>(den, rden);
#pragma unroll
  for (int i = 0; i < N; ++i) s[i] = (m[i] - 1.0) * rden[i];
#pragma unroll
  for (int i = 0; i < N; ++i) s[i] = fma(fma(-s[i], den[i], m[i] - 1.0), rden[i], s[i]);
#pragma unroll
  for (int i = 0; i < N; ++i) { z[i] = s[i] * s[i]; p[i] = c_mlog[10]; }
#pragma unroll
  for (int k = 9; k >= 2; --k)
#pragma unroll
    for (int i = 0; i < N; ++i) p[i] = fma(p[i], z[i], c_mlog[k]);
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const double s2 = s[i] + s[i];
    const double lm = fma(s2 * z[i], p[i], s2);
    y[i] = fma(ef[i], c_mlog[0], fma(ef[i], c_mlog[1], lm));
  }
}

// sin and cos of a fraction of a full turn, a in [0, 1]: quarter-turn index k = rint(4a), g = 4a - k in
// [-1/2, 1/2], sin((pi/2) g) = g S(g^2), cos((pi/2) g) = C(g^2) (max relative errors 5e-17 / 3e-17), then
// the quadrant rotation with integer sign flips - no branches, exact argument reduction.
__constant__ double c_msin[8] = {1.5707963267948966, -0.6459640975062463, 0.07969262624616692, -0.004681754135314819,
                                 0.0001604411847265753, -3.5988427165242496e-06, 5.691927654951105e-08,
                                 -6.62760087835982e-10};
__constant__ double c_mcos[9] = {1.0, -1.2337005501361697, 0.25366950790104803, -0.020863480763352916,
                                 0.0009192602748385314, -2.52020423627168e-05, 4.710874076696347e-07,
                                 -6.386325241874813e-09, 6.506630612891103e-11};
template <int N>
MOLE_D void m_sincos_turn_n(const double (&a)[N], double (&sn)[N], double (&cs)[N]) {
  double g[N], w[N], ps[N], pc[N];
  int k[N];
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const double t = fma(a[i], 4.0, 6755399441055744.0);
    k[i] = __double2loint(t);
    g[i] = fma(a[i], 4.0, -(t - 6755399441055744.0));
    w[i] = g[i] * g[i];
    ps[i] = c_msin[7];
    pc[i] = c_mcos[8];
  }
#pragma unroll
  for (int q = 6; q >= 0; --q)
#pragma unroll
    for (int i = 0; i < N; ++i) ps[i] = fma(ps[i], w[i], c_msin[q]);
#pragma unroll
  for (int q = 7; q >= 0; --q)
#pragma unroll
    for (int i = 0; i < N; ++i) pc[i] = fma(pc[i], w[i], c_mcos[q]);
#pragma unroll
  for (int i = 0; i < N; ++i) {
    ps[i] *= g[i];
    const bool odd = (k[i] & 1) != 0;
    const double s0 = odd ? pc[i] : ps[i], c0 = odd ? ps[i] : pc[i];
    sn[i] = __hiloint2double(__double2hiint(s0) ^ ((k[i] & 2) << 30), __double2loint(s0));
    cs[i] = __hiloint2double(__double2hiint(c0) ^ (((k[i] + 1) & 2) << 30), __double2loint(c0));
  }
}

#endif  // MOLE_DEVICE_CODE
