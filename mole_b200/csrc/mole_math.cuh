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

#endif  // __CUDACC__
