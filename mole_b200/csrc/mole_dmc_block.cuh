// One launch for a whole block of DMC time steps with SRBrancher (DmcRunner::diffuse's inner loop,
// src/dmc/src/dmc.rs:84-141; SRBrancher::branch src/dmc/src/branching.rs:15-40), thread-per-walker kinds.
//
// The three kernels of a step (time step + reduction, weight scan, pick + gather) are latency-bound at the BASELINE
// population (2^15 walkers per GPU: 10 + 6 + 12 us, each mostly launch ramp, dependent L2 round trips and drain).
// Here they are the three phases of one persistent, cooperatively launched kernel separated by two grid barriers per
// step; the state stays in the same global (L2-resident) arrays, so every phase is the code of the per-step kernels:
//   1  mole_dmc_walker_step per walker, CTA reduction -> one row of `partials` per 128 walkers
//      -- barrier --
//   2  every CTA folds the partial rows (mole_dmc_fold_partials, the per-step kernels' order), forms N / w_max, the
//      integer weights of its walkers and their prefix sums inside the virtual block (tile = 128 walkers), plus a
//      coarse copy of the prefix at the end of every 16 walkers
//      -- barrier --
//   3  every CTA stages the exclusive tile offsets and the coarse prefix sums in shared memory, draws, searches
//      (tile and 16-walker group in shared memory, then ONE 128-byte line of prefix sums from L2) and gathers its
//      walkers into the other buffer set; no barrier is needed before the next step's phase 1, which touches
//      only the CTA's own walkers of that set.
// Measured at 2^15 walkers (B200, -DMOLE_DMCB_PROF prints CTA 0's phase times; ns per step):
//                                           phase 1  barrier  phase 2  barrier  phase 3   step
//   per-step kernels (3 launches)                                                         25 950
//   first fused version (1024-walker scan tiles, 4-ary search: 6-7 dependent L2 round trips per pick)
//                                             2 500    1 650    3 950    2 030    8 100   18 400
//   this version                              2 410    1 790    2 130    2 090    5 240   14 000
// Population dependence (us per time step, this kernel / per-step launches): 4 096 walkers 10.4 / 24.9, 16 384: 11.4 /
// 24.9, 32 768: 13.7 / 25.0, 65 536: 22.7 / 28.9, 98 304: 43.9 / 31.1, 131 072: 46.6 / 35.6, 262 144: 81.8 / 51.6.  Beyond
// the co-resident grid (one walker per thread: ~75 000 one-electron walkers on 148 SMs) a CTA walks several virtual
// blocks per phase, their L2 round trips do not overlap, and the per-step kernels (16 CTAs per SM in flight) win: the
// default (mole_dmc_block_select 0) uses this kernel only while every CTA owns exactly one virtual block.
// What is left is ~7 dependent L2 round trips per step (own walker, partial rows, staged sums, prefix line, gather and
// one per barrier) at ~1.2 us each under this access pattern, plus ~3 us of arithmetic.
#pragma once
#include "mole_kernels.cuh"
#include "mole_branch.cuh"

constexpr int DMCB_SUB = 16;                                   // walkers per coarse prefix entry: one 128-byte line of `cum`
constexpr int DMCB_SUBS = SWEEP_THREADS / DMCB_SUB;            // 8 coarse entries per virtual block
constexpr int DMCB_STAGE_BYTES = 64 * 1024;                    // coarse prefix sums are staged in shared memory up to this size

struct DmcBlockParams {
  DmcParams dp;                       // x / w / el: the buffer set that holds the walkers at entry
  double* x2; double* w2; double* el2;
  unsigned long long* cum;            // [W] inclusive prefix sums of the integer weights inside the virtual block
  unsigned long long* tile_sums;      // [n_vb] totals of the virtual blocks
  unsigned long long* coarse;         // [n_vb * 8] prefix (inside the virtual block) at the end of every 16 walkers
  int32_t* src;
  double* step_e;                     // [n_steps][2] {sum w E, sum w}
  unsigned int* bar;                  // [0] arrival counter (zero at launch), [1] barrier time-out flag
  int n_steps, n, stage_coarse;
};

// Monotonic-counter grid barrier.  All CTAs are co-resident (cooperative launch), so the spin terminates; the bound
// turns a protocol error into an error status (bar[1]) instead of a hung device.
MOLE_D void mole_grid_barrier(unsigned int* bar, unsigned int target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(bar, 1u);
    unsigned int spins = 0;
    while (*(volatile unsigned int*)bar < target) {
      if (++spins > (1u << 26) || *(volatile unsigned int*)(bar + 1) != 0u) { atomicExch(bar + 1, 1u); break; }
    }
    __threadfence();
  }
  __syncthreads();
}

// inclusive scan of one value per thread over the CTA (SWEEP_THREADS threads); *total = sum over the CTA
MOLE_D unsigned long long mole_cta_scan_u64(unsigned long long v, unsigned long long* s_wt, unsigned long long* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned long long run = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long t = __shfl_up_sync(0xffffffffu, run, o);
    if (lane >= o) run += t;
  }
  __syncthreads();                                              // s_wt may still be read from the previous call
  if (lane == 31) s_wt[warp] = run;
  __syncthreads();
  unsigned long long off = 0ull, tot = 0ull;
#pragma unroll
  for (int q = 0; q < SWEEP_THREADS / 32; ++q) {
    const unsigned long long t = s_wt[q];
    if (q < warp) off += t;
    tot += t;
  }
  *total = tot;
  return run + off;
}

template <int KIND>
__global__ void __launch_bounds__(SWEEP_THREADS) dmc_block_kernel(const DmcBlockParams bp) {
  mole_math_smem_init();
  extern __shared__ unsigned long long s_dyn[];
  __shared__ double sm[32][4];
  __shared__ double s_red[4];
  __shared__ unsigned long long s_wt[SWEEP_THREADS / 32];
  const DmcParams& dp = bp.dp;
  double *x = dp.x, *w = dp.w, *el = dp.el, *x2 = bp.x2, *w2 = bp.w2, *el2 = bp.el2;
  const int64_t W = dp.W;
  const int n_vb = (int)((W + SWEEP_THREADS - 1) / SWEEP_THREADS);
  const int n_coarse = n_vb * DMCB_SUBS;
  unsigned long long* const s_tiles = s_dyn;                    // [n_vb + 1] exclusive offsets of the virtual blocks, grand total
  unsigned long long* const s_coarse = s_dyn + n_vb + 1;        // [n_coarse] when staged
  const double sd = sqrt(dp.tau_move);
  // a CTA that owns exactly one virtual block keeps its walkers in registers from the gather into the next time step
  constexpr int NE = WfDev<KIND>::NE;
  const bool single = (int)gridDim.x == n_vb;
  double xr[3 * NE], el_r = 0.0, w_r = 0.0;
#pragma unroll
  for (int c = 0; c < 3 * NE; ++c) xr[c] = 0.0;
  unsigned int target = 0;
#ifdef MOLE_DMCB_PROF
  unsigned long long tprof[6] = {0, 0, 0, 0, 0, 0}, t0 = 0, t1 = 0;
#define DMCB_T(i) do { if (blockIdx.x == 0 && threadIdx.x == 0) { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1)); if (t0) tprof[i] += t1 - t0; t0 = t1; } } while (0)
#else
#define DMCB_T(i) do { } while (0)
#endif
  for (int j = 0; j < bp.n_steps; ++j) {
    DMCB_T(5);
    const uint32_t step = dp.step + (uint32_t)j;
    // ---- 1: time step of the CTA's walkers (dmc.rs:87-130)
    for (int vb = blockIdx.x; vb < n_vb; vb += gridDim.x) {
      double s_we = 0.0, s_w = 0.0, s_wn = 0.0, m_wn = 0.0;
      const int64_t wi = (int64_t)vb * SWEEP_THREADS + threadIdx.x;
      if (wi < W) {
        if (single && j > 0) {                                  // the walker gathered in the previous step is still in registers
          const double wn = mole_dmc_walker_core<KIND>(dp, xr, w_r, el_r, 1, wi, step, sd, s_we, s_w, s_wn, m_wn);
          w[wi] = wn;
          el[wi] = el_r;
#pragma unroll
          for (int c = 0; c < 3 * NE; ++c) x[(size_t)c * W + wi] = xr[c];
          w_r = wn;
        } else {
          mole_dmc_walker_step<KIND>(dp, x, w, el, wi, j == 0 ? dp.el_cached : 1, step, sd, s_we, s_w, s_wn, m_wn);
        }
      }
      __syncthreads();                                          // sm is reused from the previous virtual block
      mole_dmc_cta_reduce(s_we, s_w, s_wn, m_wn, sm, dp.partials + (size_t)vb * 4);
    }
    DMCB_T(0);
    target += gridDim.x;
    mole_grid_barrier(bp.bar, target);
    DMCB_T(1);
    // ---- 2: fold of the partial rows, integer weights k_i = trunc(w_i N / w_max) (branching.rs:24-30), prefix sums
    unsigned long long draw0;
    {
      const int64_t w0 = (int64_t)blockIdx.x * SWEEP_THREADS + threadIdx.x;
      double wv = (single && j > 0) ? w_r : (w0 < W ? __ldcg(w + w0) : 0.0);   // in flight during the fold
      // the branching variate depends on (walker, step) only: drawn here, in the shadow of the barrier
      draw0 = mole_u64(mole_draw(dp.key, dp.walker_offset + (uint64_t)w0, step, DOM_BRANCH, 0, 0));
      mole_dmc_fold_partials(dp.partials, (unsigned)n_vb, s_red, sm);
      __syncthreads();
      if (blockIdx.x == 0 && threadIdx.x < 4) dp.red[threadIdx.x] = s_red[threadIdx.x];
      const double norm_factor = (double)W / s_red[3];          // branching.rs:24
      for (int vb = blockIdx.x; vb < n_vb; vb += gridDim.x) {
        const int64_t wi = (int64_t)vb * SWEEP_THREADS + threadIdx.x;
        if (vb != (int)blockIdx.x) wv = wi < W ? __ldcg(w + wi) : 0.0;
        unsigned long long k = 0ull;
        if (wi < W) {
          const double s = wv * norm_factor;
          k = (s >= 4294967295.0) ? 4294967295ull : (s > 0.0 ? (unsigned long long)(uint32_t)s : 0ull);
        }
        unsigned long long tot;
        const unsigned long long incl = mole_cta_scan_u64(k, s_wt, &tot);
        if (wi < W) bp.cum[wi] = incl;
        if ((threadIdx.x & (DMCB_SUB - 1)) == DMCB_SUB - 1) bp.coarse[(size_t)vb * DMCB_SUBS + (threadIdx.x / DMCB_SUB)] = incl;
        if (threadIdx.x == 0) bp.tile_sums[vb] = tot;
      }
    }
    DMCB_T(2);
    target += gridDim.x;
    mole_grid_barrier(bp.bar, target);
    DMCB_T(3);
    // ---- 3: exclusive tile offsets (and the coarse prefix sums) in shared memory, N weighted draws, gather (branching.rs:32-37)
    {
      if (bp.stage_coarse)
        for (int i = threadIdx.x; i < n_coarse; i += SWEEP_THREADS) s_coarse[i] = __ldcg(bp.coarse + i);
      // CTA-wide exclusive scan of the n_vb totals: thread t takes the consecutive entries [t c, (t + 1) c)
      const int c = (n_vb + SWEEP_THREADS - 1) / SWEEP_THREADS;
      unsigned long long tv[4] = {0ull, 0ull, 0ull, 0ull}, mine = 0ull;
      if (c <= 4) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int idx = threadIdx.x * c + i;
          if (i < c && idx < n_vb) tv[i] = __ldcg(bp.tile_sums + idx);
          mine += tv[i];
        }
      } else {
        for (int i = 0; i < c; ++i) {
          const int idx = threadIdx.x * c + i;
          if (idx < n_vb) mine += __ldcg(bp.tile_sums + idx);
        }
      }
      unsigned long long total;
      unsigned long long run = mole_cta_scan_u64(mine, s_wt, &total) - mine;
      if (c <= 4) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int idx = threadIdx.x * c + i;
          if (i < c && idx < n_vb) { s_tiles[idx] = run; run += tv[i]; }
        }
      } else {
        for (int i = 0; i < c; ++i) {
          const int idx = threadIdx.x * c + i;
          if (idx < n_vb) { s_tiles[idx] = run; run += __ldcg(bp.tile_sums + idx); }
        }
      }
      if (threadIdx.x == 0) s_tiles[n_vb] = total;
      __syncthreads();
      const double new_weight = s_red[2] / (double)W;           // branching.rs:21
      if (blockIdx.x == 0 && threadIdx.x == 0) {                // dmc.rs:112-113,133: the division happens on the host
        bp.step_e[2 * j] = s_red[0];
        bp.step_e[2 * j + 1] = s_red[1];
      }
      for (int vb = blockIdx.x; vb < n_vb; vb += gridDim.x) {
        const int64_t jw = (int64_t)vb * SWEEP_THREADS + threadIdx.x;
        if (jw < W) {
          const unsigned long long draw = vb == (int)blockIdx.x ? draw0 : mole_u64(mole_draw(dp.key, dp.walker_offset + (uint64_t)jw, step, DOM_BRANCH, 0, 0));
          const unsigned long long u = __umul64hi(draw, total);            // uniform integer in [0,total)
          // WeightedChoice: the first walker whose inclusive prefix sum exceeds u.  Virtual block: last one whose
          // exclusive offset is <= u (empty blocks are skipped by construction: offset[t + 1] > u)
          int tl = 0, th = n_vb - 1;
          while (tl < th) {
            const int mid = (tl + th) >> 1;
            if (s_tiles[mid + 1] > u) th = mid; else tl = mid + 1;
          }
          const unsigned long long ul = u - s_tiles[tl];
          // group of 16 walkers inside the block: first coarse prefix > ul (the last group if none: all-zero weights)
          int sub = 0;
          if (bp.stage_coarse) {
            const unsigned long long* cs = s_coarse + (size_t)tl * DMCB_SUBS;
#pragma unroll
            for (int q = DMCB_SUBS - 2; q >= 0; --q) sub += (cs[q] <= ul) ? 1 : 0;   // prefix sums are non-decreasing
          } else {
            const unsigned long long* cs = bp.coarse + (size_t)tl * DMCB_SUBS;
            unsigned long long cv[DMCB_SUBS - 1];
#pragma unroll
            for (int q = 0; q < DMCB_SUBS - 1; ++q) cv[q] = __ldcg(cs + q);
#pragma unroll
            for (int q = 0; q < DMCB_SUBS - 1; ++q) sub += (cv[q] <= ul) ? 1 : 0;
          }
          // one 128-byte line of prefix sums: first entry > ul
          const int64_t base = (int64_t)tl * SWEEP_THREADS + (int64_t)sub * DMCB_SUB;
          unsigned long long line[DMCB_SUB];
#pragma unroll
          for (int q = 0; q < DMCB_SUB; q += 2) {
            if (base + q + 1 < W) {
              const ulonglong2 t = __ldcg(reinterpret_cast<const ulonglong2*>(bp.cum + base + q));
              line[q] = t.x; line[q + 1] = t.y;
            } else {
              line[q] = base + q < W ? __ldcg(bp.cum + base + q) : ~0ull;
              line[q + 1] = ~0ull;
            }
          }
          int off = 0;
#pragma unroll
          for (int q = 0; q < DMCB_SUB - 1; ++q) off += (line[q] <= ul) ? 1 : 0;
          int64_t lo = base + off;
          lo = lo < W ? lo : W - 1;
          bp.src[jw] = (int32_t)lo;
#pragma unroll
          for (int cc = 0; cc < 3 * NE; ++cc) { xr[cc] = __ldcg(x + (size_t)cc * W + lo); x2[(size_t)cc * W + jw] = xr[cc]; }
          el_r = __ldcg(el + lo);
          el2[jw] = el_r;
          w_r = new_weight;
          w2[jw] = new_weight;
        }
      }
    }
    DMCB_T(4);
    double* t;
    t = x; x = x2; x2 = t;
    t = w; w = w2; w2 = t;
    t = el; el = el2; el2 = t;
  }
#ifdef MOLE_DMCB_PROF
  if (blockIdx.x == 0 && threadIdx.x == 0)
    printf("dmc_block_kernel CTA0 ns/step: phase1 %.0f  barrier1 %.0f  phase2 %.0f  barrier2 %.0f  phase3 %.0f  loop %.0f  (grid %d, %d steps)\n",
           (double)tprof[0] / bp.n_steps, (double)tprof[1] / bp.n_steps, (double)tprof[2] / bp.n_steps, (double)tprof[3] / bp.n_steps,
           (double)tprof[4] / bp.n_steps, (double)tprof[5] / bp.n_steps, (int)gridDim.x, bp.n_steps);
#endif
}
