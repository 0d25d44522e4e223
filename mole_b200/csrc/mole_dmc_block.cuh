// One launch for a whole block of DMC time steps with SRBrancher (DmcRunner::diffuse's inner loop,
// src/dmc/src/dmc.rs:84-141; SRBrancher::branch src/dmc/src/branching.rs:15-40), thread-per-walker kinds.
//
// The three kernels of a step (time step + reduction, weight scan, pick + gather) are latency-bound at the BASELINE
// population (2^15 walkers per GPU: 10 + 6 + 12 us, each mostly launch ramp, dependent L2 round trips and drain).
// Here they are the three phases of one persistent, cooperatively launched kernel separated by two grid barriers per
// step; the state stays in the same global (L2-resident) arrays, so every phase is the code of the per-step kernels:
//   1  mole_dmc_walker_step per walker, CTA reduction -> one row of `partials` per 128 walkers
//      -- barrier --
//   2  the CTAs that own a scan tile fold the partial rows (mole_dmc_fold_partials, the per-step kernels' order),
//      form N / w_max, the integer weights and the tile-local prefix sums
//      -- barrier --
//   3  every CTA scans the tile totals in shared memory, then draws, searches (mole_pick_tiled_ld) and gathers
//      its walkers into the other buffer set; no barrier is needed before the next step's phase 1, which touches
//      only the CTA's own walkers of that set.
// A CTA walks "virtual blocks" of 128 walkers (vb = blockIdx.x, + gridDim.x, ..), so the partial rows, their fold
// and therefore every result are bit-identical to mole_dmc_step + mole_branch whatever the co-resident grid size.
// Arrays another CTA wrote earlier in the launch are read through L2 (__ldcg): L1 is not coherent across SMs.
#pragma once
#include "mole_kernels.cuh"
#include "mole_branch.cuh"

constexpr int DMCB_ITEMS = SCAN_TILE / SWEEP_THREADS;          // 8 weights per thread and scan tile
constexpr int DMCB_MAX_TILES = 320;                            // >= partial_rows * SWEEP_THREADS / SCAN_TILE (296 on 148 SMs)

struct DmcBlockParams {
  DmcParams dp;                       // x / w / el: the buffer set that holds the walkers at entry
  double* x2; double* w2; double* el2;
  unsigned long long* cum; unsigned long long* tile_sums;
  int32_t* src;
  double* step_e;                     // [n_steps][2] {sum w E, sum w}
  unsigned int* bar;                  // [0] arrival counter (zero at launch), [1] barrier time-out flag
  int n_tiles, n_steps, n;
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

struct MoleCgLoad {
  MOLE_D unsigned long long operator()(const unsigned long long* p) const { return __ldcg(p); }
};

template <int KIND>
__global__ void __launch_bounds__(SWEEP_THREADS) dmc_block_kernel(const DmcBlockParams bp) {
  mole_math_smem_init();
  __shared__ double sm[32][4];
  __shared__ double s_red[4];
  __shared__ unsigned long long s_tiles[DMCB_MAX_TILES + 1];
  __shared__ unsigned long long s_carry;
  const DmcParams& dp = bp.dp;
  double *x = dp.x, *w = dp.w, *el = dp.el, *x2 = bp.x2, *w2 = bp.w2, *el2 = bp.el2;
  const int64_t W = dp.W;
  const int n_vb = (int)((W + SWEEP_THREADS - 1) / SWEEP_THREADS);
  const int n_tiles = bp.n_tiles;
  const double sd = sqrt(dp.tau_move);
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
      if (wi < W) mole_dmc_walker_step<KIND>(dp, x, w, el, wi, j == 0 ? dp.el_cached : 1, step, sd, s_we, s_w, s_wn, m_wn);
      __syncthreads();                                          // sm is reused from the previous virtual block
      mole_dmc_cta_reduce(s_we, s_w, s_wn, m_wn, sm, dp.partials + (size_t)vb * 4);
    }
    DMCB_T(0);
    target += gridDim.x;
    mole_grid_barrier(bp.bar, target);
    DMCB_T(1);
    // ---- 2: fold of the partial rows, integer weights k_i = trunc(w_i N / w_max) (branching.rs:24-30), tile scans
    if ((int)blockIdx.x < n_tiles || blockIdx.x == 0) {
      mole_dmc_fold_partials(dp.partials, (unsigned)n_vb, s_red, sm);
      __syncthreads();
      if (blockIdx.x == 0 && threadIdx.x < 4) dp.red[threadIdx.x] = s_red[threadIdx.x];
      const double norm_factor = (double)W / s_red[3];          // branching.rs:24
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        unsigned long long v[DMCB_ITEMS];
        const int64_t base = (int64_t)tile * SCAN_TILE + (int64_t)threadIdx.x * DMCB_ITEMS;
#pragma unroll
        for (int i = 0; i < DMCB_ITEMS; ++i) {
          unsigned long long k = 0;
          if (base + i < W) {
            const double s = __ldcg(w + base + i) * norm_factor;
            k = (s >= 4294967295.0) ? 4294967295ull : (s > 0.0 ? (unsigned long long)(uint32_t)s : 0ull);
          }
          v[i] = k;
        }
        __syncthreads();
        const unsigned long long tot = mole_tile_scan_t<SWEEP_THREADS, DMCB_ITEMS>(v);
#pragma unroll
        for (int i = 0; i < DMCB_ITEMS; ++i)
          if (base + i < W) bp.cum[base + i] = v[i];
        if (threadIdx.x == 0) bp.tile_sums[tile] = tot;
      }
    }
    DMCB_T(2);
    target += gridDim.x;
    mole_grid_barrier(bp.bar, target);
    DMCB_T(3);
    // ---- 3: exclusive tile offsets in shared memory, N weighted draws, gather (branching.rs:32-37)
    if (threadIdx.x == 0) s_carry = 0ull;
    __syncthreads();
    if (threadIdx.x < 32) {
      const int lane = threadIdx.x;
      unsigned long long carry = 0ull;
      for (int b = 0; b < n_tiles; b += 32) {
        const unsigned long long v = (b + lane < n_tiles) ? __ldcg(bp.tile_sums + b + lane) : 0ull;
        unsigned long long run = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const unsigned long long t = __shfl_up_sync(0xffffffffu, run, o);
          if (lane >= o) run += t;
        }
        if (b + lane < n_tiles) s_tiles[b + lane] = carry + run - v;
        carry += __shfl_sync(0xffffffffu, run, 31);
      }
      if (lane == 0) s_tiles[n_tiles] = carry;
    }
    __syncthreads();
    const double new_weight = __ldcg(dp.red + 2) / (double)W;   // branching.rs:21
    if (blockIdx.x == 0 && threadIdx.x == 0) {                  // dmc.rs:112-113,133: the division happens on the host
      bp.step_e[2 * j] = __ldcg(dp.red);
      bp.step_e[2 * j + 1] = __ldcg(dp.red + 1);
    }
    const unsigned long long total = s_tiles[n_tiles];
    for (int vb = blockIdx.x; vb < n_vb; vb += gridDim.x) {
      const int64_t jw = (int64_t)vb * SWEEP_THREADS + threadIdx.x;
      if (jw < W) {
        const Philox4 p = mole_draw(dp.key, dp.walker_offset + (uint64_t)jw, step, DOM_BRANCH, 0, 0);
        const unsigned long long u = __umul64hi(mole_u64(p), total);     // uniform integer in [0,total)
        const int64_t lo = mole_pick_tiled_ld(bp.cum, s_tiles, n_tiles, W, SCAN_TILE, u, MoleCgLoad());
        bp.src[jw] = (int32_t)lo;
        for (int c = 0; c < bp.n; ++c) x2[(size_t)c * W + jw] = __ldcg(x + (size_t)c * W + lo);
        el2[jw] = __ldcg(el + lo);
        w2[jw] = new_weight;
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
