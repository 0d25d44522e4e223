// Two-level weighted pick of SRBrancher (src/dmc/src/branching.rs:27-36: WeightedChoice = first index whose
// inclusive cumulative weight exceeds the draw) on TILE-LOCAL inclusive prefix sums plus exclusive tile offsets.
// Plain C++ (no CUDA types) so that the same function is compiled into sr_pick_gather_tiled_kernel and into the
// host harness of tests/test_search_host.py, which compares it with std::upper_bound on the global prefix sums.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define MOLE_SEARCH_HD __host__ __device__ __forceinline__
#else
#define MOLE_SEARCH_HD inline
#endif

// cum[i]  : inclusive prefix sum of the integer weights INSIDE the tile of walker i (tile = i / tile_len)
// ts[t]   : exclusive offset of tile t, ts[n_tiles] = grand total; requires 0 <= u < ts[n_tiles]
// returns the first walker whose global inclusive prefix sum is > u (never a zero-weight walker)
// `ld` reads one prefix sum: a plain load, or a load through L2 where another CTA of the same launch wrote the array
template <class Load>
MOLE_SEARCH_HD int64_t mole_pick_tiled_ld(const unsigned long long* cum, const unsigned long long* ts, int n_tiles, int64_t W,
                                          int tile_len, unsigned long long u, Load ld) {
  // last tile whose exclusive offset is <= u, skipping empty tiles by construction (offset[t+1] > u)
  int tl = 0, th = n_tiles - 1;
  while (tl < th) {
    const int mid = (tl + th) >> 1;
    if (ts[mid + 1] > u) th = mid; else tl = mid + 1;
  }
  const unsigned long long ul = u - ts[tl];
  int64_t lo = (int64_t)tl * tile_len, hi = lo + tile_len;
  hi = (hi < W ? hi : W) - 1;
  // 4-ary rounds: three independent probes per dependent round trip.  Invariant: the answer is in [lo, hi].
  while (hi - lo >= 4) {
    const int64_t q = (hi - lo) >> 2, m1 = lo + q, m2 = m1 + q, m3 = m2 + q;   // lo < m1 < m2 < m3 < hi
    const unsigned long long c1 = ld(cum + m1), c2 = ld(cum + m2), c3 = ld(cum + m3);
    if (c1 > ul) hi = m1;
    else if (c2 > ul) { lo = m1 + 1; hi = m2; }
    else if (c3 > ul) { lo = m2 + 1; hi = m3; }
    else lo = m3 + 1;
  }
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (ld(cum + mid) > ul) hi = mid; else lo = mid + 1;
  }
  return lo;
}

struct MolePlainLoad {
  MOLE_SEARCH_HD unsigned long long operator()(const unsigned long long* p) const { return *p; }
};
MOLE_SEARCH_HD int64_t mole_pick_tiled(const unsigned long long* cum, const unsigned long long* ts, int n_tiles, int64_t W,
                                       int tile_len, unsigned long long u) {
  return mole_pick_tiled_ld(cum, ts, n_tiles, W, tile_len, u, MolePlainLoad());
}
