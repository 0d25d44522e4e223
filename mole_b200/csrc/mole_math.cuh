// Branch-free fp64 elementary functions for the hot kernels.
//
// CUDA's exp(), sqrt() and '/' each carry a rarely-taken slow path; the branch splits the basic block,
// so independent chains (the two Jastrow pairs, the orbital exponential, the accept exponentials of one
// Metropolis move) cannot be interleaved by the scheduler and the FP64 pipe idles on dependent-issue
// latency.  The versions below are straight-line code: a MUFU seed refined by Newton steps in DFMA, and
// a degree-9 polynomial for exp.  Accuracy is ~1 ulp (not correctly rounded), far inside the 1e-10
// parity bar; domains are stated per function.  tests/test_math_device.py checks them on the device.
#pragma once
#include "mole_internal.h"

#if defined(__CUDACC__)

// 1/x for finite normal x != 0  (MUFU.RCP64H seed, two Newton steps, ~1 ulp)
MOLE_D double m_rcp(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = fma(-x, y, 1.0);
  y = fma(y, fma(e, e, e), y);          // cubic step: y (1 + e + e^2)
  e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  return y;
}

// a/b for finite normal b != 0
MOLE_D double m_div(double a, double b) {
  const double y = m_rcp(b);
  const double q = a * y;
  return fma(fma(-b, q, a), y, q);
}

// 1/sqrt(x) for finite normal x > 0  (MUFU.RSQ64H seed, two Newton steps)
MOLE_D double m_rsqrt(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double hx = 0.5 * x;
  double e = fma(-hx * y, y, 0.5);      // (1 - x y^2)/2
  y = fma(y, fma(1.5 * e, e, e), y);    // cubic step: y (1 + e' + 1.5 e'^2), e' = e
  e = fma(-hx * y, y, 0.5);
  y = fma(y, e, y);
  return y;
}

// sqrt(x) and 1/sqrt(x) together, x > 0 finite normal
MOLE_D double m_sqrt_rsqrt(double x, double& rinv) {
  rinv = m_rsqrt(x);
  const double s = x * rinv;
  return fma(fma(-s, s, x), 0.5 * rinv, s);
}

// exp(x) on the whole real line, branch-free: 0 below -745.2, +inf above 709.8, NaN propagates
MOLE_D double m_exp(double x) {
  const double xc = fmin(fmax(x, -746.0), 710.0);
  const double t = fma(xc, 1.4426950408889634, 6755399441055744.0);     // 1.5 * 2^52: low word = round(x log2 e)
  const int n = __double2loint(t);
  const double nf = t - 6755399441055744.0;
  double r = fma(nf, -6.93147180369123816490e-01, xc);                  // ln2 hi (32 trailing zero bits)
  r = fma(nf, -1.90821492927058770002e-10, r);                          // ln2 lo
  // exp(r) = 1 + r + r^2 q(r), |r| <= ln2/2; q = degree-9 near-minimax (max rel. error 1.6e-17)
  double q = 2.510038549551032e-08;
  q = fma(q, r, 2.7620088445409746e-07);
  q = fma(q, r, 2.7557268459997064e-06);
  q = fma(q, r, 2.4801521295954376e-05);
  q = fma(q, r, 0.00019841269863053618);
  q = fma(q, r, 0.0013888888917213717);
  q = fma(q, r, 0.008333333333330062);
  q = fma(q, r, 0.04166666666662413);
  q = fma(q, r, 0.16666666666666669);
  q = fma(q, r, 0.5000000000000001);
  const double p = fma(r * r, q, r) + 1.0;
  const int n1 = n >> 1, n2 = n - n1;                                   // two-step scaling covers denormals / overflow
  const double s1 = __hiloint2double((n1 + 1023) << 20, 0);
  const double s2 = __hiloint2double((n2 + 1023) << 20, 0);
  const double res = (p * s1) * s2;
  return isnan(x) ? x : res;
}

#endif  // __CUDACC__
