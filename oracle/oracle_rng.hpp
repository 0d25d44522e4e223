// ORACLE — TEST INFRASTRUCTURE ONLY.  Never imported by the product path (mole_b200/).
//
// Counter-based RNG of the stream contract (DESIGN.md §"Philox stream contract").
// The reference draws from rand 0.5 `StdRng` (HC-128), an un-vendored crate
// (`src/metropolis/src/metrop.rs:4-6,43`, `Cargo.toml:33`); north_star replaces it
// with Philox keyed by (walker, step), so this file restates the *published*
// Philox4x32-10 algorithm (Salmon et al., SC'11, Random123 v1.09) independently
// of the CUDA implementation in mole_b200/csrc/.  Pinned by the Random123
// known-answer vectors in tests/test_oracle_rng.py.
#pragma once
#include <cstdint>
#include <cmath>
#include <cstring>

namespace orc {

struct Philox4 { uint32_t w[4]; };

inline Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                             uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
  const uint32_t W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
  for (int round = 0; round < 10; ++round) {
    const uint64_t p0 = (uint64_t)M0 * c0;
    const uint64_t p1 = (uint64_t)M1 * c2;
    const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
    const uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
    const uint32_t n0 = hi1 ^ c1 ^ k0;
    const uint32_t n1 = lo1;
    const uint32_t n2 = hi0 ^ c3 ^ k1;
    const uint32_t n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += W0; k1 += W1;
  }
  return Philox4{{c0, c1, c2, c3}};
}

// ---- stream contract -------------------------------------------------------
enum Domain : uint32_t { DOM_MOVE = 0, DOM_INIT = 1, DOM_BRANCH = 2, DOM_SEED = 3 };

struct Key { uint32_t k0, k1; };

// 32-byte seed (the reference's `[u8; 32]`, metropolis/src/traits.rs:33-36)
// folded to the 64-bit Philox key: XOR of the even / odd little-endian words.
inline Key key_from_seed(const uint8_t seed[32]) {
  uint32_t s[8];
  for (int i = 0; i < 8; ++i)
    s[i] = (uint32_t)seed[4 * i] | ((uint32_t)seed[4 * i + 1] << 8) |
           ((uint32_t)seed[4 * i + 2] << 16) | ((uint32_t)seed[4 * i + 3] << 24);
  return Key{s[0] ^ s[2] ^ s[4] ^ s[6], s[1] ^ s[3] ^ s[5] ^ s[7]};
}

inline Philox4 draw(Key k, uint64_t walker, uint32_t step, Domain dom, uint32_t elec,
                    uint32_t slot) {
  const uint32_t c3 = ((uint32_t)dom << 28) | (elec << 4) | slot;
  return philox4x32_10((uint32_t)walker, (uint32_t)(walker >> 32), step, c3, k.k0, k.k1);
}

// 53-bit uniform in [0,1)
inline double u53(uint32_t lo, uint32_t hi) {
  const uint64_t x = ((uint64_t)hi << 32) | lo;
  return (double)(x >> 11) * 0x1.0p-53;
}
// 53-bit uniform in (0,1]  (argument of the Box-Muller log)
inline double u53_open(uint32_t lo, uint32_t hi) {
  const uint64_t x = ((uint64_t)hi << 32) | lo;
  return (double)((x >> 11) + 1) * 0x1.0p-53;
}
// 32-bit fraction of a full turn, [0,1)
inline double turn32(uint32_t w) { return (double)w * 0x1.0p-32; }

// `generate_seed` (metropolis/src/traits.rs:35-37): n-th 32-byte seed derived from a master key.
inline void derive_seed(const uint8_t master[32], uint32_t n, uint8_t out[32]) {
  const Key k = key_from_seed(master);
  for (uint32_t half = 0; half < 2; ++half) {
    const Philox4 r = draw(k, n, 0, DOM_SEED, 0, half);
    for (int i = 0; i < 4; ++i)
      for (int b = 0; b < 4; ++b) out[16 * half + 4 * i + b] = (uint8_t)(r.w[i] >> (8 * b));
  }
}

// Three standard normals + one uniform for a diffusion move, or four uniforms for a
// box move: two Philox calls (slots 0 and 1) per (walker, step, electron).
struct MoveDraw { double a, b, c, u; };

inline MoveDraw draw_uniform4(Key k, uint64_t walker, uint32_t step, Domain dom, uint32_t elec) {
  const Philox4 p = draw(k, walker, step, dom, elec, 0);
  const Philox4 q = draw(k, walker, step, dom, elec, 1);
  return MoveDraw{u53(p.w[0], p.w[1]), u53(p.w[2], p.w[3]), u53(q.w[0], q.w[1]),
                  u53(q.w[2], q.w[3])};
}

inline MoveDraw draw_normal3_uniform1(Key k, uint64_t walker, uint32_t step, Domain dom,
                                      uint32_t elec) {
  const Philox4 p = draw(k, walker, step, dom, elec, 0);
  const Philox4 q = draw(k, walker, step, dom, elec, 1);
  const double two_pi = 6.283185307179586476925286766559;
  const double r1 = std::sqrt(-2.0 * std::log(u53_open(p.w[0], p.w[1])));
  const double r2 = std::sqrt(-2.0 * std::log(u53_open(q.w[0], q.w[1])));
  const double t1 = two_pi * turn32(p.w[2]);
  const double t2 = two_pi * turn32(p.w[3]);
  return MoveDraw{r1 * std::cos(t1), r1 * std::sin(t1), r2 * std::cos(t2), u53(q.w[2], q.w[3])};
}

}  // namespace orc
