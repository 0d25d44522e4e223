// Branch-free, batched fp64 elementary functions for the hot kernels.
//
// Why: (1) CUDA's exp(), sqrt() and '/' each carry a rarely-taken slow path whose branch splits the
// basic block; (2) ptxas schedules a Horner / Newton chain as one serial run of DFMAs, and a dependent
// DFMA issues only every ~9.4 cycles on B200 (measured, scratch/ubench/dfma.cu) while the FP64 pipe
// accepts one per 2 cycles per SM sub-partition.  With the 3 warps per scheduler the Slater-Jastrow
// kernel affords, a serial chain leaves the pipe ~75% idle.  The functions below are straight-line
// code, take N independent arguments at once and are written stage-major, so the N chains (and the
// Estrin sub-terms inside exp) are adjacent independent instructions.
// Accuracy ~1-2 ulp (not correctly rounded), far inside the 1e-10 parity bar; checked on the device
// by tests/test_math_device.py.
#pragma once
#include "mole_internal.h"

#if defined(__CUDACC__)

// 1/x for finite normal x != 0  (MUFU.RCP64H seed, cubic + quadratic Newton step)
template <int N>
MOLE_D void m_rcp_n(const double (&x)[N], double (&y)[N]) {
  double e[N];
#pragma unroll
  for (int i = 0; i < N; ++i) asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y[i]) : "d"(x[i]));
#pragma unroll
  for (int i = 0; i < N; ++i) e[i] = fma(-x[i], y[i], 1.0);
#pragma unroll
  for (int i = 0; i < N; ++i) e[i] = fma(e[i], e[i], e[i]);
#pragma unroll
  for (int i = 0; i < N; ++i) y[i] = fma(y[i], e[i], y[i]);          // y (1 + e + e^2)
#pragma unroll
  for (int i = 0; i < N; ++i) e[i] = fma(-x[i], y[i], 1.0);
#pragma unroll
  for (int i = 0; i < N; ++i) y[i] = fma(y[i], e[i], y[i]);
}
MOLE_D double m_rcp(double x) {
  const double a[1] = {x};
  double y[1];
  m_rcp_n<1>(a, y);
  return y[0];
}

// 1/sqrt(x) for finite normal x > 0  (MUFU.RSQ64H seed, cubic + quadratic Newton step)
template <int N>
MOLE_D void m_rsqrt_n(const double (&x)[N], double (&y)[N]) {
  double hx[N], e[N];
#pragma unroll
  for (int i = 0; i < N; ++i) asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y[i]) : "d"(x[i]));
#pragma unroll
  for (int i = 0; i < N; ++i) hx[i] = 0.5 * x[i];
#pragma unroll
  for (int i = 0; i < N; ++i) e[i] = fma(-hx[i] * y[i], y[i], 0.5);   // (1 - x y^2)/2
#pragma unroll
  for (int i = 0; i < N; ++i) y[i] = fma(y[i], fma(1.5 * e[i], e[i], e[i]), y[i]);
#pragma unroll
  for (int i = 0; i < N; ++i) e[i] = fma(-hx[i] * y[i], y[i], 0.5);
#pragma unroll
  for (int i = 0; i < N; ++i) y[i] = fma(y[i], e[i], y[i]);
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
#pragma unroll
  for (int i = 0; i < N; ++i) s[i] = fma(fma(-s[i], s[i], x[i]), 0.5 * rinv[i], s[i]);
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

#endif  // __CUDACC__
