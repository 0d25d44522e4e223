// Per-walker series statistics on the device: the reference's offline analysis tool
// scripts/statfor.rs (mean :17-19, variance :23-26, correlation :31-54, blocking :58-81) applied to
// every walker's E_L series, series[s * W + w], s = 0..n-1.  Summation orders follow the reference
// (sequential in s), so the per-walker results agree with the oracle's restatement to rounding of
// fused multiply-adds only.
#pragma once
#include "mole_internal.h"

constexpr int STAT_WX = 16;          // walkers per CTA (x): 128-byte rows
constexpr int STAT_LY = 16;          // lag groups per CTA (y)
constexpr int STAT_MAX_LAG = 200;    // MAX_STEPS, statfor.rs:32
constexpr int STAT_NSTAT = 5;        // average, variance, tcorr, n_eff, sigma

// mean (sum / len) and variance with ddof = 1: sum((average - x)^2) / (len - 1)
__global__ void series_moments_kernel(const double* __restrict__ series, int64_t W, int64_t n, double* __restrict__ mean,
                                      double* __restrict__ var) {
  const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= W) return;
  double s = 0.0;
  for (int64_t t = 0; t < n; ++t) s += series[t * W + w];
  const double av = s / (double)n;
  double v = 0.0;
  for (int64_t t = 0; t < n; ++t) {
    const double d = av - series[t * W + w];
    v += d * d;
  }
  mean[w] = av;
  var[w] = v / (double)(n - 1);
}

// autocorrelation corr_i = sum_{s < n-i} (x_s - mean)(x_{s+i} - mean) / (variance (n - i)), i = 1..lags,
// then tcorr = max(1, 1 + 2 sum_i corr_i f) with f dropping to 0 at the first negative corr_i,
// n_eff = n / tcorr, sigma = sqrt(variance tcorr / n).  CTA = 16 walkers x 16 lag groups; the corr_i
// of the CTA's walkers sit in shared memory for the sequential tcorr fold, and their sums over the
// 16 walkers go to corr_part[blockIdx.x][lags] / stat_part[blockIdx.x][5] for the walker mean.
__global__ void __launch_bounds__(STAT_WX* STAT_LY) series_corr_kernel(const double* __restrict__ series, int64_t W, int64_t n,
                                                                       int lags, const double* __restrict__ mean,
                                                                       const double* __restrict__ var,
                                                                       double* __restrict__ stats /* [W][5] */,
                                                                       double* __restrict__ corr_out /* [lags][W] or null */,
                                                                       double* __restrict__ corr_part, double* __restrict__ stat_part) {
  __shared__ double sc[STAT_MAX_LAG][STAT_WX + 1];
  __shared__ double st[STAT_NSTAT][STAT_WX + 1];
  const int wx = threadIdx.x, ty = threadIdx.y;
  const int64_t w = (int64_t)blockIdx.x * STAT_WX + wx;
  const bool live = w < W;
  const double av = live ? mean[w] : 0.0, vr = live ? var[w] : 1.0;
  for (int i = 1 + ty; i <= lags; i += STAT_LY) {
    double c = 0.0;
    if (live) {
      const double* p = series + w;
      for (int64_t s = 0; s < n - i; ++s) c += (p[s * W] - av) * (p[(s + i) * W] - av);
      c /= vr * (double)(n - i);
    }
    sc[i - 1][wx] = c;
    if (live && corr_out) corr_out[(size_t)(i - 1) * W + w] = c;
  }
  __syncthreads();
  if (ty == 0) {
    double tc = 1.0, f = 1.0;
    for (int i = 0; i < lags; ++i) {
      const double c = sc[i][wx];
      if (c < 0.0) f = 0.0;
      tc += 2.0 * c * f;
    }
    tc = fmax(tc, 1.0);                       // f64::max: a NaN tcorr becomes 1.0
    const double neff = (double)n / tc, sg = sqrt(vr * tc / (double)n);
    if (live) {
      double* o = stats + (size_t)w * STAT_NSTAT;
      o[0] = av; o[1] = vr; o[2] = tc; o[3] = neff; o[4] = sg;
    }
    st[0][wx] = live ? av : 0.0; st[1][wx] = live ? vr : 0.0; st[2][wx] = live ? tc : 0.0;
    st[3][wx] = live ? neff : 0.0; st[4][wx] = live ? sg : 0.0;
  }
  __syncthreads();
  // fixed-order sums over the CTA's walkers
  const int tid = ty * STAT_WX + wx;
  const int nlive = (int)min((int64_t)STAT_WX, W - (int64_t)blockIdx.x * STAT_WX);
  for (int i = tid; i < lags; i += STAT_WX * STAT_LY) {
    double s = 0.0;
    for (int k = 0; k < nlive; ++k) s += sc[i][k];
    corr_part[(size_t)blockIdx.x * lags + i] = s;
  }
  if (tid < STAT_NSTAT) {
    double s = 0.0;
    for (int k = 0; k < nlive; ++k) s += st[tid][k];
    stat_part[(size_t)blockIdx.x * STAT_NSTAT + tid] = s;
  }
}

// blocking error at one block size (statfor.rs:69-78): data.chunks(size) INCLUDES the ragged last
// chunk in both means, while the divisor uses nblocks = n / size.
__global__ void series_blocking_kernel(const double* __restrict__ series, int64_t W, int64_t n, int n_sizes,
                                       const int32_t* __restrict__ sizes, double* __restrict__ err /* [n_sizes][W] */) {
  const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int k = blockIdx.y;
  if (w >= W || k >= n_sizes) return;
  const int64_t size = sizes[k];
  const int64_t nblocks = n / size;
  const double* p = series + w;
  double sm = 0.0, sm2 = 0.0;
  int64_t nch = 0;
  for (int64_t b = 0; b < n; b += size) {
    const int64_t e = min(b + size, n);
    double s = 0.0;
    for (int64_t t = b; t < e; ++t) s += p[t * W];
    const double m = s / (double)(e - b);
    sm += m;
    sm2 += m * m;
    ++nch;
  }
  const double ave = sm / (double)nch, ave_sq = sm2 / (double)nch;
  err[(size_t)k * W + w] = sqrt((ave_sq - ave * ave) / (double)(nblocks - 1));
}

// out[j] = (sum over rows r of part[r][j]) / denom, rows summed in order (deterministic)
__global__ void stat_fold_rows_kernel(const double* __restrict__ part, int rows, int cols, double denom, double* __restrict__ out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= cols) return;
  double s = 0.0;
  for (int r = 0; r < rows; ++r) s += part[(size_t)r * cols + j];
  out[j] = s / denom;
}

// out[k] = mean over walkers of err[k][.]: one CTA per size, fixed-order tree
__global__ void __launch_bounds__(256) stat_row_mean_kernel(const double* __restrict__ a, int64_t W, double* __restrict__ out) {
  __shared__ double sh[256];
  const double* row = a + (size_t)blockIdx.x * W;
  double s = 0.0;
  for (int64_t i = threadIdx.x; i < W; i += 256) s += row[i];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[blockIdx.x] = sh[0] / (double)W;
}

// copies one walker's series (stride W) into a contiguous buffer
__global__ void series_extract_kernel(const double* __restrict__ series, int64_t W, int64_t n, int64_t w, double* __restrict__ out) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) out[t] = series[t * W + w];
}
