// DMC branching / population control as GPU scan + search + gather
// (BranchingAlgorithm::branch, src/dmc/src/traits.rs:4-7; SRBrancher src/dmc/src/branching.rs:15-40;
//  SimpleBranching branching.rs:50-91).
#pragma once
#include "mole_internal.h"
#include "mole_rng.cuh"
#include "mole_search.h"

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 4;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

// inclusive scan of one tile held as ITEMS consecutive values per thread (THREADS threads); returns tile total.
// Integer sums: the result does not depend on the CTA shape.
template <int THREADS, int ITEMS>
MOLE_D unsigned long long mole_tile_scan_t(unsigned long long v[ITEMS]) {
  __shared__ unsigned long long warp_tot[THREADS / 32];
  __shared__ unsigned long long tile_total;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 1; i < ITEMS; ++i) v[i] += v[i - 1];
  unsigned long long run = v[ITEMS - 1];
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long t = __shfl_up_sync(0xffffffffu, run, o);
    if (lane >= o) run += t;
  }
  if (lane == 31) warp_tot[warp] = run;
  __syncthreads();
  if (warp == 0) {
    unsigned long long t = lane < THREADS / 32 ? warp_tot[lane] : 0ull;
#pragma unroll
    for (int o = 1; o < THREADS / 32; o <<= 1) {
      const unsigned long long u = __shfl_up_sync(0xffffffffu, t, o);
      if (lane >= o) t += u;
    }
    if (lane < THREADS / 32) warp_tot[lane] = t;
    if (lane == THREADS / 32 - 1) tile_total = t;
  }
  __syncthreads();
  const unsigned long long excl = (run - v[ITEMS - 1]) + (warp > 0 ? warp_tot[warp - 1] : 0ull);
#pragma unroll
  for (int i = 0; i < ITEMS; ++i) v[i] += excl;
  return tile_total;
}
MOLE_D unsigned long long mole_tile_scan(unsigned long long v[SCAN_ITEMS]) { return mole_tile_scan_t<SCAN_THREADS, SCAN_ITEMS>(v); }

// pass 1 (SR): integer weights k_i = trunc(w_i * N/w_max) (branching.rs:24-30, `as u32` saturates),
// tile-local inclusive scan -> cum, tile totals -> tile_sums
// exclusive scan of the tile totals in place by ONE CTA (sequential carry over chunks);
// tile_sums[n_tiles] receives the grand total
MOLE_D void mole_scan_tile_sums(unsigned long long* tile_sums, int n_tiles) {
  __shared__ unsigned long long carry;
  if (threadIdx.x == 0) carry = 0ull;
  __syncthreads();
  for (int base = 0; base < n_tiles; base += SCAN_TILE) {
    unsigned long long v[SCAN_ITEMS], orig[SCAN_ITEMS];
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
      const int idx = base + threadIdx.x * SCAN_ITEMS + i;
      orig[i] = v[i] = idx < n_tiles ? __ldcg(tile_sums + idx) : 0ull;   // written by other CTAs: read through L2
    }
    const unsigned long long tot = mole_tile_scan(v);
    const unsigned long long c = carry;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
      const int idx = base + threadIdx.x * SCAN_ITEMS + i;
      if (idx < n_tiles) tile_sums[idx] = c + v[i] - orig[i];
    }
    __syncthreads();
    if (threadIdx.x == 0) carry = c + tot;
    __syncthreads();
  }
  if (threadIdx.x == 0) tile_sums[n_tiles] = carry;
}

// Fused SR pass 1+2 for the launch-bound DMC loop: as sr_weights_scan_kernel, but the normalisation
// N / w_max is read from the device (max_ptr, written by the step kernel's reduction or an NCCL max)
// and the last CTA to finish (ticket) scans the tile totals, so no second launch and no host read.
// cum stays TILE-LOCAL; sr_pick_gather_kernel searches tiles first.
__global__ void __launch_bounds__(SCAN_THREADS) sr_weights_scan_fused_kernel(const double* __restrict__ w, int64_t W,
                                                                             double norm_factor, const double* red_rows,
                                                                             int n_rows, double gcount,
                                                                             unsigned long long* cum,
                                                                             unsigned long long* tile_sums, int n_tiles,
                                                                             unsigned int* ticket) {
  __shared__ bool is_last;
  if (red_rows) {                                               // rows of {sum w E, sum w, sum w', max w'}, one per rank
    double mx = red_rows[3];
    for (int r = 1; r < n_rows; ++r) mx = fmax(mx, red_rows[4 * r + 3]);
    norm_factor = gcount / mx;                                  // branching.rs:24 (global N, global w_max)
  }
  unsigned long long v[SCAN_ITEMS];
  const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    unsigned long long k = 0;
    if (base + i < W) {
      const double s = w[base + i] * norm_factor;
      k = (s >= 4294967295.0) ? 4294967295ull : (s > 0.0 ? (unsigned long long)(uint32_t)s : 0ull);
    }
    v[i] = k;
  }
  const unsigned long long tot = mole_tile_scan(v);
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i)
    if (base + i < W) cum[base + i] = v[i];
  if (threadIdx.x == 0) {
    tile_sums[blockIdx.x] = tot;
    __threadfence();
    const unsigned int t = atomicAdd(ticket, 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    mole_scan_tile_sums(tile_sums, n_tiles);
    if (threadIdx.x == 0) *ticket = 0u;
  }
}

__global__ void __launch_bounds__(SCAN_THREADS) sr_weights_scan_kernel(const double* __restrict__ w, int64_t W,
                                                                       double norm_factor, unsigned long long* cum,
                                                                       unsigned long long* tile_sums) {
  unsigned long long v[SCAN_ITEMS];
  const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    unsigned long long k = 0;
    if (base + i < W) {
      const double s = w[base + i] * norm_factor;
      k = (s >= 4294967295.0) ? 4294967295ull : (s > 0.0 ? (unsigned long long)(uint32_t)s : 0ull);
    }
    v[i] = k;
  }
  const unsigned long long tot = mole_tile_scan(v);
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i)
    if (base + i < W) cum[base + i] = v[i];
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = tot;
}

// pass 2: exclusive scan of the tile totals in place (single CTA, sequential carry over chunks);
// tile_sums[n_tiles] receives the grand total
__global__ void __launch_bounds__(SCAN_THREADS) scan_tile_sums_kernel(unsigned long long* tile_sums, int n_tiles) {
  mole_scan_tile_sums(tile_sums, n_tiles);
}

// pass 3: cum[i] += offset of its tile
__global__ void add_tile_offsets_kernel(unsigned long long* cum, int64_t W, const unsigned long long* tile_sums) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < W) cum[i] += tile_sums[i / SCAN_TILE];
}

// pass 4 (SR): N draws from WeightedChoice (first index with cumulative weight > u), gather the
// configuration + cached E_L, every new walker gets the mean weight (branching.rs:32-37)
__global__ void sr_pick_gather_kernel(const unsigned long long* __restrict__ cum, const unsigned long long* tile_sums,
                                      int n_tiles, int64_t W, int n, uint64_t walker_offset, RngKey key, uint32_t step,
                                      const double* __restrict__ x, double* x2, const double* __restrict__ el, double* el2,
                                      double* w2, double new_weight, int32_t* src) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= W) return;
  const unsigned long long total = tile_sums[n_tiles];
  const Philox4 p = mole_draw(key, walker_offset + (uint64_t)j, step, DOM_BRANCH, 0, 0);
  const unsigned long long u = __umul64hi(mole_u64(p), total);     // uniform integer in [0,total)
  int64_t lo = 0, hi = W - 1;                                       // upper_bound: first cum > u
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (cum[mid] > u) hi = mid; else lo = mid + 1;
  }
  src[j] = (int32_t)lo;
  for (int c = 0; c < n; ++c) x2[(size_t)c * W + j] = x[(size_t)c * W + lo];
  el2[j] = el[lo];
  w2[j] = new_weight;
}

// pass 4 (SR) on TILE-LOCAL prefix sums: the search goes over the exclusive tile offsets first and then
// inside the tile - the same walker as the upper bound on the global prefix sums, without the pass that
// adds the offsets.  The new weight and the step's ensemble energy are formed from device-side sums
// (local_sum_ptr = sum of this rank's post-update weights; red_rows = {sum w E, sum w, ..} rows), exactly as
// the host does on the synchronous path, and the step's {sum w E, sum w} go to step_e_out[0..1].
constexpr int PICK_SMEM_TILES = 1024;   // tile offsets of up to 2^20 walkers are searched in shared memory
__global__ void sr_pick_gather_tiled_kernel(const unsigned long long* __restrict__ cum, const unsigned long long* tile_sums,
                                            int n_tiles, int64_t W, int n, uint64_t walker_offset, RngKey key, uint32_t step,
                                            const double* __restrict__ x, double* x2, const double* __restrict__ el,
                                            double* el2, double* w2, double new_weight, const double* local_sum_ptr,
                                            const double* red_rows, int n_rows, double* step_e_out, int32_t* src) {
  // The step is bound by dependent-load latency, not bandwidth (measured: 10 us for this kernel at 4096 and at 32768
  // walkers alike).  The tile offsets are staged in shared memory once per CTA, and the search inside the tile is 4-ary:
  // three independent probes per round, 4 + 2 dependent L2 round trips instead of 10.  The picked index is the first
  // one whose inclusive prefix sum exceeds the draw, which is unique, so the result is that of the binary search.
  __shared__ unsigned long long s_tiles[PICK_SMEM_TILES + 1];
  const bool staged = n_tiles <= PICK_SMEM_TILES;
  if (staged) {
    for (int i = threadIdx.x; i <= n_tiles; i += blockDim.x) s_tiles[i] = tile_sums[i];
    __syncthreads();
  }
  const unsigned long long* ts = staged ? s_tiles : tile_sums;
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j == 0 && step_e_out) {                                     // this step's {sum w E, sum w} (dmc.rs:112-113,133): the division
    double swe = red_rows[0], sw = red_rows[1];                   // happens on the host, after ONE gather per block over the ranks
    for (int r = 1; r < n_rows; ++r) { swe += red_rows[4 * r]; sw += red_rows[4 * r + 1]; }
    step_e_out[0] = swe;
    step_e_out[1] = sw;
  }
  if (j >= W) return;
  if (local_sum_ptr) new_weight = *local_sum_ptr / (double)W;     // branching.rs:21
  const unsigned long long total = ts[n_tiles];
  const Philox4 p = mole_draw(key, walker_offset + (uint64_t)j, step, DOM_BRANCH, 0, 0);
  const unsigned long long u = __umul64hi(mole_u64(p), total);     // uniform integer in [0,total)
  const int64_t lo = mole_pick_tiled(cum, ts, n_tiles, W, SCAN_TILE, u);   // mole_search.h (host-tested)
  src[j] = (int32_t)lo;
  for (int c = 0; c < n; ++c) x2[(size_t)c * W + j] = x[(size_t)c * W + lo];
  el2[j] = el[lo];
  w2[j] = new_weight;
}

// ---------------------------------------------------------------------------------------------------
// SimpleBranching, branching.rs:50-91.
// pass 1: copies_i = min(trunc(w_i + u_i), 3) (:60); scan survivors (copies>0) and births (copies-1)
// packed as (survivors << 32 | births) in one 64-bit scan.
__global__ void __launch_bounds__(SCAN_THREADS) simple_copies_scan_kernel(const double* __restrict__ w, int64_t W,
                                                                          uint64_t walker_offset, RngKey key, uint32_t step,
                                                                          unsigned long long* cum,
                                                                          unsigned long long* tile_sums) {
  unsigned long long v[SCAN_ITEMS];
  const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    unsigned long long k = 0;
    if (base + i < W) {
      const Philox4 p = mole_draw(key, walker_offset + (uint64_t)(base + i), step, DOM_BRANCH, 0, 0);
      const double t = w[base + i] + mole_u53(p.a, p.b);
      unsigned long long copies = t > 0.0 ? (unsigned long long)t : 0ull;
      copies = copies < 3ull ? copies : 3ull;
      if (copies > 0) k = (1ull << 32) | (copies - 1);
    }
    v[i] = k;
  }
  const unsigned long long tot = mole_tile_scan(v);
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i)
    if (base + i < W) cum[base + i] = v[i];
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = tot;
}

// pass 4a: build the source-index list of the un-clamped new population:
// survivors in walker order, then births in walker order (:69-72); excess clones appended (:75-82).
__global__ void simple_build_list_kernel(const unsigned long long* __restrict__ cum, const unsigned long long* tile_sums,
                                         int n_tiles, int64_t W, uint64_t walker_offset, RngKey key, uint32_t step,
                                         int32_t* list /* capacity 3W */) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned long long total = tile_sums[n_tiles];
  const int64_t n_surv = (int64_t)(total >> 32), n_birth = (int64_t)(total & 0xffffffffull);
  const int64_t n_new = n_surv + n_birth;
  if (i < W) {
    const unsigned long long incl = cum[i];
    const unsigned long long prev = i > 0 ? cum[i - 1] : 0ull;
    const int64_t s_incl = (int64_t)(incl >> 32), s_prev = (int64_t)(prev >> 32);
    const int64_t b_incl = (int64_t)(incl & 0xffffffffull), b_prev = (int64_t)(prev & 0xffffffffull);
    if (s_incl > s_prev) list[s_prev] = (int32_t)i;
    for (int64_t b = b_prev; b < b_incl; ++b) list[n_surv + b] = (int32_t)i;
  }
  if (n_new < W && i < W - n_new) {                                 // clone random OLD walkers
    const Philox4 p = mole_draw(key, walker_offset + (uint64_t)i, step, DOM_BRANCH, 1, 0);
    list[n_new + i] = (int32_t)__umul64hi(mole_u64(p), (unsigned long long)W);
  }
}

// pass 4b: remove `excess` random entries one at a time, each index drawn from the CURRENT list
// length (:83-89).  Sequential by definition; done by one CTA with an alive-bitmask and a Fenwick
// tree of per-word popcounts (shared memory when they fit, else global scratch); the draws themselves
// are generated in parallel beforehand.  Writes the compacted list to list_out.
__global__ void __launch_bounds__(1024) simple_remove_kernel(int32_t* list, const unsigned long long* tile_sums,
                                                             int n_tiles, int64_t W, uint64_t walker_offset, RngKey key,
                                                             uint32_t step, uint32_t* mask_g /* words */, int32_t* fen_g /* words+1 */,
                                                             int64_t* draws /* 2W */, int smem_words, int32_t* list_out) {
  extern __shared__ uint32_t sb_smem[];
  const unsigned long long total = tile_sums[n_tiles];
  const int64_t n_new = (int64_t)(total >> 32) + (int64_t)(total & 0xffffffffull);
  if (n_new <= W) {
    for (int64_t i = threadIdx.x; i < W; i += blockDim.x) list_out[i] = list[i];
    return;
  }
  const int64_t words = (n_new + 31) / 32;
  // the bitmask and the Fenwick tree sit in shared memory whenever they fit (the walk below is a chain of
  // dependent loads: ~30 cycles each there against ~600 in global memory)
  const bool in_smem = words + 1 <= (int64_t)smem_words;
  uint32_t* mask = in_smem ? sb_smem : mask_g;
  int32_t* fen = in_smem ? (int32_t*)(sb_smem + smem_words) : fen_g;
  const int64_t excess = n_new - W;
  // the j-th removal draws from the CURRENT length n_new - j, which is known in advance: all draws in parallel
  for (int64_t j = threadIdx.x; j < excess; j += blockDim.x) {
    const Philox4 p = mole_draw(key, walker_offset + (uint64_t)j, step, DOM_BRANCH, 2, 0);
    draws[j] = (int64_t)__umul64hi(mole_u64(p), (unsigned long long)(n_new - j));
  }
  for (int64_t q = threadIdx.x; q < words; q += blockDim.x) {
    const int64_t rem = n_new - q * 32;
    mask[q] = rem >= 32 ? 0xffffffffu : ((1u << rem) - 1u);
  }
  __syncthreads();
  // Fenwick tree over per-word popcounts (1-based)
  for (int64_t q = threadIdx.x; q < words; q += blockDim.x) fen[q + 1] = __popc(mask[q]);
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int64_t q = 1; q <= words; ++q) {
      const int64_t parent = q + (q & -q);
      if (parent <= words) fen[parent] += fen[q];
    }
    int64_t top = 1;
    while (top * 2 <= words) top *= 2;
    for (int64_t j = 0; j < excess; ++j) {
      int64_t k = draws[j];                                            // k-th (0-based) alive entry
      int64_t pos = 0;
      for (int64_t bit = top; bit > 0; bit >>= 1) {
        const int64_t nxt = pos + bit;
        if (nxt <= words && fen[nxt] <= k) { pos = nxt; k -= fen[nxt]; }
      }
      uint32_t m = mask[pos];                                          // word `pos` (0-based) holds the target
      uint32_t mm = m;
      for (int64_t t = 0; t < k; ++t) mm &= mm - 1;                    // drop k lowest set bits
      const int b = __ffs(mm) - 1;
      mask[pos] = m & ~(1u << b);
      for (int64_t q = pos + 1; q <= words; q += q & -q) fen[q] -= 1;
    }
  }
  __syncthreads();
  // compaction: rank of each alive entry
  __shared__ int64_t chunk_base;
  if (threadIdx.x == 0) chunk_base = 0;
  __syncthreads();
  for (int64_t q0 = 0; q0 < words; q0 += blockDim.x) {
    const int64_t q = q0 + threadIdx.x;
    const uint32_t m = q < words ? mask[q] : 0u;
    // block-wide exclusive scan of popcounts
    __shared__ int32_t sc[1024];
    sc[threadIdx.x] = __popc(m);
    __syncthreads();
    for (int o = 1; o < (int)blockDim.x; o <<= 1) {
      const int32_t t = threadIdx.x >= (unsigned)o ? sc[threadIdx.x - o] : 0;
      __syncthreads();
      sc[threadIdx.x] += t;
      __syncthreads();
    }
    int64_t out = chunk_base + sc[threadIdx.x] - __popc(m);
    uint32_t mm = m;
    while (mm) {
      const int b = __ffs(mm) - 1;
      mm &= mm - 1;
      list_out[out++] = list[q * 32 + b];
    }
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) chunk_base += sc[threadIdx.x];
    __syncthreads();
  }
}

// pass 5: gather by source list; births/clones keep the parent's weight (:65, :80)
__global__ void gather_by_list_kernel(const int32_t* __restrict__ list, int64_t W, int n, const double* __restrict__ x,
                                      double* x2, const double* __restrict__ el, double* el2, const double* __restrict__ w,
                                      double* w2, int32_t* src) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= W) return;
  const int32_t s = list[j];
  src[j] = s;
  for (int c = 0; c < n; ++c) x2[(size_t)c * W + j] = x[(size_t)c * W + s];
  el2[j] = el[s];
  w2[j] = w[s];
}

// ---------------------------------------------------------------------------------------------------
// Cross-rank population rebalancing (mole_rebalance): every walker of the rank carries the same weight; the rank's
// walkers fill `share` slots of the global population by systematic resampling, copy j <- walker
// floor((j + v) count / share) mod count.  Copies j < keep stay (gathered into x2 / el2), copies j >= keep are
// packed as rows [3 n_e coordinates, E_L] for the exchange.
__global__ void rebalance_gather_kernel(const double* __restrict__ x, double* x2, const double* __restrict__ el, double* el2,
                                        int64_t W, int n, int64_t share, double v, int64_t keep, double* rows) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= share) return;
  int64_t src = (int64_t)(((double)j + v) * (double)W / (double)share);
  src = src < 0 ? 0 : (src >= W ? src % W : src);
  if (j < keep) {
    for (int c = 0; c < n; ++c) x2[(size_t)c * W + j] = x[(size_t)c * W + src];
    el2[j] = el[src];
  } else {
    double* r = rows + (size_t)(j - keep) * (n + 1);
    for (int c = 0; c < n; ++c) r[c] = x[(size_t)c * W + src];
    r[n] = el[src];
  }
}
// received rows fill the slots first .. first + count - 1
__global__ void rebalance_scatter_kernel(const double* __restrict__ rows, double* x2, double* el2, int64_t W, int n, int64_t first,
                                         int64_t count) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= count) return;
  const double* r = rows + (size_t)j * (n + 1);
  for (int c = 0; c < n; ++c) x2[(size_t)c * W + first + j] = r[c];
  el2[first + j] = r[n];
}
