// Gram matrix of the per-sample rows v = (1, E_L, O_1 .. O_P): G = sum_samples v v^T, the one contraction that
// holds every optimisation moment of a large-P kind (n, sum E, sum E^2, sum O_k, sum O_k E, sum O_k O_l;
// compute_energy_gradient src/optimize/src/util.rs:6-46, construct_sr_matrix src/optimize/src/optimizers.rs:191-233).
// north_star (3): "the S = O^T O contraction on tensor cores only once the parameter count makes it a genuinely dense
// GEMM" - here it is: M = walkers x samples rows (1e6 .. 1e8), N = K = P + 2 <= 48 columns, fp64.
//   gram_dmma_kernel   mma.sync.aligned.m8n8k4.f64 (DMMA): a warp takes 16 rows at a time (4 when W is odd), keeps the
//                      upper triangle of 8 x 8 tiles in registers; two 16-byte loads per column tile and lane
//   gram_fma_kernel    the same contraction on the FP64 vector pipe (shared-memory row tiles), for the A/B the
//                      survey asks for (SURVEY 7 hard part 4) and as the check of the DMMA fragment layout
// Both are HBM-streaming: every row is read once, 8 (P + 2) bytes for (P + 2)(P + 3) flops, i.e. ~ (P + 3) / 8 flop/B.
// Per-CTA partial matrices are folded in a fixed order (deterministic for a given geometry), no float atomics.
#pragma once
#include "mole_internal.h"

// Build-time variants kept for the A/B (B200, 65536 walkers x 190 samples, ms per launch at 38 / 46 / 22 columns):
//   ORDER 1 PIPE 1 MINB 1   0.989 / 1.226 / 0.434      value-major MMAs, next row group prefetched into a second register set
//   ORDER 1 PIPE 0 MINB 2   0.819 / 1.262 / 0.463
//   ORDER 0 PIPE 0 MINB 2   0.818 / 1.150 / 0.462      <- default: two CTAs (16 warps) per SM hide the load latency
//   ORDER 1 PIPE 1 MINB 2   1.671 / 3.678 / 0.426      (spills)
// 0.818 ms = 4.63 TB/s (0.71 of the measured HBM copy rate) and 29 TFLOP/s of executed DMMA (0.79 of the 37.0 the
// m8n8k4 probe reaches, mole_bench_dmma_peak): at 38+ columns the tensor pipe, not HBM, is the nearer ceiling.
#ifndef GRAM_ORDER
#define GRAM_ORDER 0
#endif
#ifndef GRAM_PIPE
#define GRAM_PIPE 0
#endif
#ifndef GRAM_MINB
#define GRAM_MINB 2
#endif
constexpr int GRAM_PAD = 48;                    // columns padded to 6 tiles of 8
constexpr int GRAM_TILES = GRAM_PAD / 8;
constexpr int GRAM_THREADS = 256;
constexpr int GRAM_NPAIR = GRAM_TILES * (GRAM_TILES + 1) / 2;   // 21 tile pairs (upper triangle)

// rows are addressed as data[(s * cols + col) * W + w] for sample s, walker w
// NT = number of 8-column tiles actually present, a template parameter: a predicated-off DMMA still occupies its issue
// slot and the accumulator chain (measured: 14 and 46 columns took the same time per row group with a run-time bound)
template <int NT>
__global__ void __launch_bounds__(GRAM_THREADS, GRAM_MINB) gram_dmma_kernel(const double* __restrict__ data, int64_t W, int64_t n_samples, int cols,
                                                                 double* partials) {
#if defined(MOLE_EMU)
  (void)data; (void)W; (void)n_samples; (void)cols; (void)partials;
#else
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = GRAM_THREADS / 32;
  const int kr = lane & 3, mc = lane >> 2;       // row inside the 4-row group, column inside the 8-column tile
  constexpr int NPAIR = NT * (NT + 1) / 2;
  double acc[NPAIR][2];
#pragma unroll
  for (int p = 0; p < NPAIR; ++p) { acc[p][0] = 0.0; acc[p][1] = 0.0; }
  // 16 rows per step when W is even (16-byte aligned pairs): lane (kr, mc) loads rows 2 kr, 2 kr + 1 of two 8-row
  // blocks with one LDG.128 each; the four values feed four MMAs (any four rows form a valid k-group), so 12 independent
  // 16-byte loads are in flight per lane before the first MMA.  Odd W: 4 rows per step, 8-byte loads.
  if ((W & 1) == 0) {
    const int64_t groups_per_sample = (W + 15) / 16;
    const int64_t n_groups = groups_per_sample * n_samples;
    // GRAM_PIPE: software pipeline, the fragments of the warp's next row group in flight while the MMAs of the current
    // one issue (two register sets).  Measured slower than simply running two CTAs per SM (table above).
    const int64_t stride = (int64_t)gridDim.x * nwarp;
#if GRAM_PIPE
    double2 fa[NT], fb[NT], na[NT], nb[NT];
#else
    double2 fa[NT], fb[NT];
#endif
    auto load = [&](int64_t g, double2* a, double2* b) {
      const int64_t s = g / groups_per_sample, w0 = (g - s * groups_per_sample) * 16 + 2 * kr;
#pragma unroll
      for (int t = 0; t < NT; ++t) {
        const int col = 8 * t + mc;
        const bool okc = g < n_groups && col < cols;
        const double* row = data + ((size_t)s * cols + (okc ? col : 0)) * W;
        a[t] = (okc && w0 < W) ? *reinterpret_cast<const double2*>(row + w0) : make_double2(0.0, 0.0);
        b[t] = (okc && w0 + 8 < W) ? *reinterpret_cast<const double2*>(row + w0 + 8) : make_double2(0.0, 0.0);
      }
    };
    int64_t g = (int64_t)blockIdx.x * nwarp + warp;
#if GRAM_PIPE
    load(g, fa, fb);
#endif
    for (; g < n_groups; g += stride) {
#if GRAM_PIPE
      load(g + stride, na, nb);
#else
      load(g, fa, fb);
#endif
#if GRAM_ORDER == 1
      // value-major: the NT (NT + 1) / 2 accumulators of one k-group are independent, so a single warp keeps the
      // DMMA pipe busy (dependent m8n8k4 issue ~27 cycles apart, independent ones 16: mole_bench_dmma_peak)
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        int p = 0;
#pragma unroll
        for (int ta = 0; ta < NT; ++ta)
#pragma unroll
          for (int tb = ta; tb < NT; ++tb, ++p) {
            const double xa = v == 0 ? fa[ta].x : v == 1 ? fa[ta].y : v == 2 ? fb[ta].x : fb[ta].y;
            const double xb = v == 0 ? fa[tb].x : v == 1 ? fa[tb].y : v == 2 ? fb[tb].x : fb[tb].y;
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(acc[p][0]), "+d"(acc[p][1]) : "d"(xa), "d"(xb));
          }
      }
#else
      int p = 0;
#pragma unroll
      for (int ta = 0; ta < NT; ++ta)
#pragma unroll
        for (int tb = ta; tb < NT; ++tb, ++p)
          {
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(acc[p][0]), "+d"(acc[p][1]) : "d"(fa[ta].x), "d"(fa[tb].x));
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(acc[p][0]), "+d"(acc[p][1]) : "d"(fa[ta].y), "d"(fa[tb].y));
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(acc[p][0]), "+d"(acc[p][1]) : "d"(fb[ta].x), "d"(fb[tb].x));
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(acc[p][0]), "+d"(acc[p][1]) : "d"(fb[ta].y), "d"(fb[tb].y));
          }
#endif
#if GRAM_PIPE
#pragma unroll
      for (int t = 0; t < NT; ++t) { fa[t] = na[t]; fb[t] = nb[t]; }
#endif
    }
  } else {
    const int64_t groups_per_sample = (W + 3) / 4;
    const int64_t n_groups = groups_per_sample * n_samples;
    for (int64_t g = (int64_t)blockIdx.x * nwarp + warp; g < n_groups; g += (int64_t)gridDim.x * nwarp) {
      const int64_t s = g / groups_per_sample, w = (g - s * groups_per_sample) * 4 + kr;
      double frag[NT];
#pragma unroll
      for (int t = 0; t < NT; ++t) {
        const int col = 8 * t + mc;
        frag[t] = (col < cols && w < W) ? data[((size_t)s * cols + col) * W + w] : 0.0;
      }
      int p = 0;
#pragma unroll
      for (int ta = 0; ta < NT; ++ta)
#pragma unroll
        for (int tb = ta; tb < NT; ++tb, ++p)
          // A = X^T tile (8 x 4), B = X tile (4 x 8): the same register serves as both
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
      for (int ta = 0; ta < NT; ++ta)
#pragma unroll
        for (int tb = ta; tb < NT; ++tb, ++p) {
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

#if !defined(MOLE_EMU)
inline void gram_dmma_launch(int rows, cudaStream_t st, const double* data, int64_t W, int64_t n_samples, int cols, double* partials) {
  switch ((cols + 7) / 8) {
    case 1: gram_dmma_kernel<1><<<rows, GRAM_THREADS, 0, st>>>(data, W, n_samples, cols, partials); break;
    case 2: gram_dmma_kernel<2><<<rows, GRAM_THREADS, 0, st>>>(data, W, n_samples, cols, partials); break;
    case 3: gram_dmma_kernel<3><<<rows, GRAM_THREADS, 0, st>>>(data, W, n_samples, cols, partials); break;
    case 4: gram_dmma_kernel<4><<<rows, GRAM_THREADS, 0, st>>>(data, W, n_samples, cols, partials); break;
    case 5: gram_dmma_kernel<5><<<rows, GRAM_THREADS, 0, st>>>(data, W, n_samples, cols, partials); break;
    default: gram_dmma_kernel<6><<<rows, GRAM_THREADS, 0, st>>>(data, W, n_samples, cols, partials); break;
  }
}
#endif

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

// DMMA issue-rate probe: `CH` independent accumulator chains per warp, register-resident operands
template <int CH>
__global__ void dmma_peak_kernel(double* out, int iters, double seed) {
#if !defined(MOLE_EMU)
  double acc[CH][2];
#pragma unroll
  for (int c = 0; c < CH; ++c) { acc[c][0] = seed + c; acc[c][1] = seed - c; }
  const double a = 1e-3 * (threadIdx.x & 7), b = 1e-3 * (threadIdx.x >> 3 & 3);
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int c = 0; c < CH; ++c)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(acc[c][0]), "+d"(acc[c][1]) : "d"(a), "d"(b));
  }
  double s = 0.0;
#pragma unroll
  for (int c = 0; c < CH; ++c) s += acc[c][0] + acc[c][1];
  if (s == 123.456) out[blockIdx.x] = s;
#endif
}
