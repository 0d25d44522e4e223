// Counter-based Philox4x32-10 stream (DESIGN.md "Philox stream contract").
// Replaces the reference's `StdRng` plumbing (src/metropolis/src/traits.rs:30-37,
// src/metropolis/src/metrop.rs:56-58,98-100,214-216): a draw is a pure function of
// (seed key, global walker id, sweep counter, domain, electron, slot), so results do not
// depend on the launch geometry or on how walkers are sharded over GPUs.
#pragma once
#include "mole_internal.h"

struct Philox4 { uint32_t a, b, c, d; };

MOLE_HD uint32_t mole_mulhi32(uint32_t x, uint32_t y) {
#if defined(__CUDA_ARCH__)
  return __umulhi(x, y);
#else
  return (uint32_t)(((uint64_t)x * y) >> 32);
#endif
}

MOLE_HD Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = mole_mulhi32(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = mole_mulhi32(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ k0;
    c1 = lo1;
    c2 = hi0 ^ c3 ^ k1;
    c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return Philox4{c0, c1, c2, c3};
}

enum { DOM_MOVE = 0, DOM_INIT = 1, DOM_BRANCH = 2, DOM_SEED = 3 };

MOLE_HD Philox4 mole_draw(RngKey k, uint64_t walker, uint32_t step, uint32_t dom, uint32_t elec, uint32_t slot) {
  return philox4x32_10((uint32_t)walker, (uint32_t)(walker >> 32), step, (dom << 28) | (elec << 4) | slot, k.k0, k.k1);
}

// 53-bit uniform in [0,1)
MOLE_HD double mole_u53(uint32_t lo, uint32_t hi) {
  const uint64_t x = ((uint64_t)hi << 32) | lo;
  return (double)(x >> 11) * 0x1.0p-53;
}
// 53-bit uniform in (0,1]
MOLE_HD double mole_u53_open(uint32_t lo, uint32_t hi) {
  const uint64_t x = ((uint64_t)hi << 32) | lo;
  return (double)((x >> 11) + 1) * 0x1.0p-53;
}
MOLE_HD uint64_t mole_u64(const Philox4& p) { return ((uint64_t)p.b << 32) | p.a; }

struct MoveDraw { double a, b, c, u; };

// four uniforms: MetropolisBox (metrop.rs:64-68,81) and Sampler::new (samplers.rs:46)
MOLE_HD MoveDraw mole_draw_uniform4(RngKey k, uint64_t walker, uint32_t step, uint32_t dom, uint32_t elec) {
  const Philox4 p = mole_draw(k, walker, step, dom, elec, 0);
  const Philox4 q = mole_draw(k, walker, step, dom, elec, 1);
  return MoveDraw{mole_u53(p.a, p.b), mole_u53(p.c, p.d), mole_u53(q.a, q.b), mole_u53(q.c, q.d)};
}

#if defined(MOLE_DEVICE_CODE) && !defined(MOLE_HOST_ONLY)
#include "mole_math.cuh"
// three standard normals (Box-Muller) + one uniform: MetropolisDiffuse (metrop.rs:160,197), DmcRunner::new (dmc.rs:52-56)
// r = sqrt(-2 ln u), angle = 2 pi (w / 2^32); the two logs, two square roots and two sin/cos go through
// the branch-free batched math (exact argument reduction for the angles).
__device__ __forceinline__ MoveDraw mole_draw_normal3_uniform1(RngKey k, uint64_t walker, uint32_t step, uint32_t dom,
                                                               uint32_t elec) {
  const Philox4 p = mole_draw(k, walker, step, dom, elec, 0);
  const Philox4 q = mole_draw(k, walker, step, dom, elec, 1);
  const double u[2] = {mole_u53_open(p.a, p.b), mole_u53_open(q.a, q.b)};
  double lg[2], arg[2], r[2], ri[2];
  m_log_n<2>(u, lg);
  arg[0] = fmax(-2.0 * lg[0], 1e-300);                   // u = 1 gives ln u = 0: keep the rsqrt seed finite (r ~ 1e-150)
  arg[1] = fmax(-2.0 * lg[1], 1e-300);
  m_sqrt_rsqrt_n<2>(arg, r, ri);
  const double a[2] = {(double)p.c * 0x1.0p-32, (double)p.d * 0x1.0p-32};
  double sn[2], cs[2];
  m_sincos_turn_n<2>(a, sn, cs);
  return MoveDraw{r[0] * cs[0], r[0] * sn[0], r[1] * cs[1], mole_u53(q.c, q.d)};
}
#endif
