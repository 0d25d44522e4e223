// Gram matrix of the per-sample rows v = (1, E_L, O_1 .. O_P): G = sum_samples v v^T, the one contraction that
// holds every optimisation moment of a large-P kind (n, sum E, sum E^2, sum O_k, sum O_k E, sum O_k O_l;
// compute_energy_gradient src/optimize/src/util.rs:6-46, construct_sr_matrix src/optimize/src/optimizers.rs:191-233).
// north_star (3): "the S = O^T O contraction on tensor cores only once the parameter count makes it a genuinely dense
// GEMM" - here it is: M = walkers x samples rows (1e6 .. 1e8), N = K = P + 2 <= 48 columns, fp64.
//   gram_dmma_kernel   mma.sync.aligned.m8n8k4.f64 (DMMA): a warp takes 4 rows at a time, keeps the upper triangle of
//                      8 x 8 tiles in registers; one 8-byte load per column tile and lane, 4-row segments of 32 bytes
//   gram_fma_kernel    the same contraction on the FP64 vector pipe (shared-memory row tiles), for the A/B the
//                      survey asks for (SURVEY 7 hard part 4) and as the check of the DMMA fragment layout
// Both are HBM-streaming: every row is read once, 8 (P + 2) bytes for (P + 2)(P + 3) flops, i.e. ~ (P + 3) / 8 flop/B.
// Per-CTA partial matrices are folded in a fixed order (deterministic for a given geometry), no float atomics.
#pragma once
#include "mole_internal.h"

constexpr int GRAM_PAD = 48;                    // columns padded to 6 tiles of 8
constexpr int GRAM_TILES = GRAM_PAD / 8;
constexpr int GRAM_THREADS = 256;
constexpr int GRAM_NPAIR = GRAM_TILES * (GRAM_TILES + 1) / 2;   // 21 tile pairs (upper triangle)

// rows are addressed as data[(s * cols + col) * W + w] for sample s, walker w
__global__ void __launch_bounds__(GRAM_THREADS) gram_dmma_kernel(const double* __restrict__ data, int64_t W, int64_t n_samples, int cols,
                                                                 double* partials) {
#if defined(MOLE_EMU)
  (void)data; (void)W; (void)n_samples; (void)cols; (void)partials;
#else
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = GRAM_THREADS / 32;
  const int kr = lane & 3, mc = lane >> 2;       // row inside the 4-row group, column inside the 8-column tile
  const int ntile = (cols + 7) / 8;
  double acc[GRAM_NPAIR][2];
#pragma unroll
  for (int p = 0; p < GRAM_NPAIR; ++p) { acc[p][0] = 0.0; acc[p][1] = 0.0; }
  // 16 rows per step when W is even (16-byte aligned pairs): lane (kr, mc) loads rows 2 kr, 2 kr + 1 of two 8-row
  // blocks with one LDG.128 each; the four values feed four MMAs (any four rows form a valid k-group), so 12 independent
  // 16-byte loads are in flight per lane before the first MMA.  Odd W: 4 rows per step, 8-byte loads.
  if ((W & 1) == 0) {
    const int64_t groups_per_sample = (W + 15) / 16;
    const int64_t n_groups = groups_per_sample * n_samples;
    for (int64_t g = (int64_t)blockIdx.x * nwarp + warp; g < n_groups; g += (int64_t)gridDim.x * nwarp) {
      const int64_t s = g / groups_per_sample, w0 = (g - s * groups_per_sample) * 16 + 2 * kr;
      double2 fa[GRAM_TILES], fb[GRAM_TILES];
#pragma unroll
      for (int t = 0; t < GRAM_TILES; ++t) {
        const int col = 8 * t + mc;
        const bool okc = t < ntile && col < cols;
        const double* row = data + ((size_t)s * cols + (okc ? col : 0)) * W;
        fa[t] = (okc && w0 < W) ? *reinterpret_cast<const double2*>(row + w0) : make_double2(0.0, 0.0);
        fb[t] = (okc && w0 + 8 < W) ? *reinterpret_cast<const double2*>(row + w0 + 8) : make_double2(0.0, 0.0);
      }
      // pair-major (the four MMAs of a tile pair back to back) measured faster than value-major (2296 vs 1932 GB/s at
      // 38 columns): the kernel is bound by the DMMA pipe (~14.5 TFLOP/s executed with the padding), not by the loads
      int p = 0;
#pragma unroll
      for (int ta = 0; ta < GRAM_TILES; ++ta)
#pragma unroll
        for (int tb = ta; tb < GRAM_TILES; ++tb, ++p)
          if (tb < ntile) {
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(acc[p][0]), "+d"(acc[p][1]) : "d"(fa[ta].x), "d"(fa[tb].x));
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(acc[p][0]), "+d"(acc[p][1]) : "d"(fa[ta].y), "d"(fa[tb].y));
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(acc[p][0]), "+d"(acc[p][1]) : "d"(fb[ta].x), "d"(fb[tb].x));
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(acc[p][0]), "+d"(acc[p][1]) : "d"(fb[ta].y), "d"(fb[tb].y));
          }
    }
  } else {
    const int64_t groups_per_sample = (W + 3) / 4;
    const int64_t n_groups = groups_per_sample * n_samples;
    for (int64_t g = (int64_t)blockIdx.x * nwarp + warp; g < n_groups; g += (int64_t)gridDim.x * nwarp) {
      const int64_t s = g / groups_per_sample, w = (g - s * groups_per_sample) * 4 + kr;
      double frag[GRAM_TILES];
#pragma unroll
      for (int t = 0; t < GRAM_TILES; ++t) {
        const int col = 8 * t + mc;
        frag[t] = (t < ntile && col < cols && w < W) ? data[((size_t)s * cols + col) * W + w] : 0.0;
      }
      int p = 0;
#pragma unroll
      for (int ta = 0; ta < GRAM_TILES; ++ta)
#pragma unroll
        for (int tb = ta; tb < GRAM_TILES; ++tb, ++p)
          if (tb < ntile)                     // A = X^T tile (8 x 4), B = X tile (4 x 8): the same register serves as both
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(acc[p][0]), "+d"(acc[p][1]) : "d"(frag[ta]), "d"(frag[tb]));
    }
  }
  // C fragment: row = lane / 4, columns 2 (lane % 4) + {0, 1} of the tile
  __shared__ double sm[GRAM_PAD * GRAM_PAD];
  for (int i = threadIdx.x; i < GRAM_PAD * GRAM_PAD; i += GRAM_THREADS) sm[i] = 0.0;
  __syncthreads();
  for (int wq = 0; wq < nwarp; ++wq) {          // fixed warp order
    if (warp == wq) {
      int p = 0;
#pragma unroll
      for (int ta = 0; ta < GRAM_TILES; ++ta)
#pragma unroll
        for (int tb = ta; tb < GRAM_TILES; ++tb, ++p) {
          const int r = 8 * ta + mc, c0 = 8 * tb + 2 * kr;
          sm[r * GRAM_PAD + c0] += acc[p][0];
          sm[r * GRAM_PAD + c0 + 1] += acc[p][1];
        }
    }
    __syncthreads();
  }
  for (int i = threadIdx.x; i < GRAM_PAD * GRAM_PAD; i += GRAM_THREADS) partials[(size_t)blockIdx.x * GRAM_PAD * GRAM_PAD + i] = sm[i];
#endif
}

// the same on the vector pipe: a CTA stages GRAM_ROWS rows in shared memory, thread t owns the (a, b) pairs
// t, t + T, .. of the upper triangle and runs the dot products over the staged rows
constexpr int GRAM_ROWS = 64;
__global__ void __launch_bounds__(GRAM_THREADS) gram_fma_kernel(const double* __restrict__ data, int64_t W, int64_t n_samples, int cols,
                                                                double* partials) {
  __shared__ double tile[GRAM_PAD][GRAM_ROWS + 1];
  constexpr int MAXP = (GRAM_PAD * (GRAM_PAD + 1) / 2 + GRAM_THREADS - 1) / GRAM_THREADS;   // 5
  const int npair = cols * (cols + 1) / 2;
  double acc[MAXP];
  int pa[MAXP], pb[MAXP];
#pragma unroll
  for (int q = 0; q < MAXP; ++q) {
    acc[q] = 0.0;
    int idx = threadIdx.x + q * GRAM_THREADS, a = 0;
    if (idx >= npair) { pa[q] = -1; pb[q] = 0; continue; }
    while (idx >= cols - a) { idx -= cols - a; ++a; }
    pa[q] = a; pb[q] = a + idx;
  }
  const int64_t chunks_per_sample = (W + GRAM_ROWS - 1) / GRAM_ROWS;
  const int64_t n_chunks = chunks_per_sample * n_samples;
  for (int64_t ch = blockIdx.x; ch < n_chunks; ch += gridDim.x) {
    const int64_t s = ch / chunks_per_sample, w0 = (ch - s * chunks_per_sample) * GRAM_ROWS;
    __syncthreads();
    for (int i = threadIdx.x; i < cols * GRAM_ROWS; i += GRAM_THREADS) {
      const int col = i / GRAM_ROWS, r = i - col * GRAM_ROWS;
      tile[col][r] = (w0 + r < W) ? data[((size_t)s * cols + col) * W + w0 + r] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < MAXP; ++q) {
      if (pa[q] < 0) continue;
      double s0 = 0.0, s1 = 0.0;
#pragma unroll 8
      for (int r = 0; r < GRAM_ROWS; r += 2) {
        s0 = fma(tile[pa[q]][r], tile[pb[q]][r], s0);
        s1 = fma(tile[pa[q]][r + 1], tile[pb[q]][r + 1], s1);
      }
      acc[q] += s0 + s1;
    }
  }
  double* out = partials + (size_t)blockIdx.x * GRAM_PAD * GRAM_PAD;
  for (int i = threadIdx.x; i < GRAM_PAD * GRAM_PAD; i += GRAM_THREADS) out[i] = 0.0;
  __syncthreads();
#pragma unroll
  for (int q = 0; q < MAXP; ++q)
    if (pa[q] >= 0) out[pa[q] * GRAM_PAD + pb[q]] = acc[q];
}

// fold of the per-CTA partial matrices in CTA order, added to gram (GRAM_PAD x GRAM_PAD, upper triangle meaningful)
__global__ void gram_fold_kernel(const double* __restrict__ partials, int n_blocks, double* gram) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= GRAM_PAD * GRAM_PAD) return;
  double s = 0.0;
  for (int b = 0; b < n_blocks; ++b) s += partials[(size_t)b * GRAM_PAD * GRAM_PAD + i];
  gram[i] += s;
}

// synthetic rows for mole_bench_gram: a cheap hash of the index mapped to (-1, 1)
__global__ void gram_fill_kernel(double* data, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    unsigned long long z = i * 0x9E3779B97F4A7C15ull + 0xD1B54A32D192ED03ull;
    z ^= z >> 29; z *= 0xBF58476D1CE4E5B9ull; z ^= z >> 32;
    data[i] = (double)(long long)(z >> 11) * (1.0 / 4503599627370496.0) - 1.0;
  }
}
