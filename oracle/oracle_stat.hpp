// ORACLE - TEST INFRASTRUCTURE ONLY.
// Restatement of the reference's offline series analysis, scripts/statfor.rs (the same correlation()
// is in scripts/statfor.py:21-40, which tests/golden/make_statfor_golden.py imports to pin this file).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <vector>

namespace orc {

// scripts/statfor.rs:17-19
inline double stat_mean(const double* d, int64_t n) {
  double s = 0.0;
  for (int64_t i = 0; i < n; ++i) s += d[i];
  return s / (double)n;
}

// scripts/statfor.rs:23-26
inline double stat_variance(const double* d, int64_t n, int ddof) {
  const double average = stat_mean(d, n);
  double s = 0.0;
  for (int64_t i = 0; i < n; ++i) s += std::pow(average - d[i], 2);
  return s / (double)(n - ddof);
}

// scripts/statfor.rs:31-54; corr (nullable) receives corr_1..corr_max_i; returns max_i
inline int stat_correlation(const double* d, int64_t nsteps, double mean, double variance, double* corr_out,
                            double* tcorr_out, double* neff_out, double* sigma_out) {
  const int64_t MAX_STEPS = 200;
  double tcorr = 1.0;
  const int64_t max_i = std::min(MAX_STEPS, nsteps - 1);
  double f = 1.0;
  for (int64_t i = 1; i <= max_i; ++i) {
    double corr = 0.0;
    for (int64_t s = 0; s < nsteps - i; ++s) corr += (d[s] - mean) * (d[s + i] - mean);
    corr /= variance * (double)(nsteps - i);
    if (corr < 0.0) f = 0.0;
    tcorr += 2.0 * corr * f;
    if (corr_out) corr_out[i - 1] = corr;
  }
  tcorr = std::isnan(tcorr) ? 1.0 : std::max(tcorr, 1.0);   // f64::max drops a NaN operand
  *tcorr_out = tcorr;
  *neff_out = (double)nsteps / tcorr;
  *sigma_out = std::sqrt(variance * tcorr / (double)nsteps);
  return (int)max_i;
}

// block-size schedule, scripts/statfor.rs:59-66
inline std::vector<int> stat_block_sizes(int64_t ndata) {
  const int64_t MIN_LEFT = 20, NSIZES = 100;
  const int64_t large = ndata / MIN_LEFT;
  const int64_t step_size = std::max<int64_t>(large / NSIZES, 1);
  std::vector<int> out;
  for (int64_t size = 1; size <= large; size += step_size) out.push_back((int)size);
  return out;
}

// estimated error at one block size, scripts/statfor.rs:67-76: data.chunks(size) keeps the ragged last
// chunk, nblocks = ndata / size
inline double stat_block_error(const double* d, int64_t ndata, int64_t size) {
  const int64_t nblocks = ndata / size;
  std::vector<double> ave_blks;
  for (int64_t b = 0; b < ndata; b += size) ave_blks.push_back(stat_mean(d + b, std::min(size, ndata - b)));
  std::vector<double> sq(ave_blks.size());
  for (size_t i = 0; i < sq.size(); ++i) sq[i] = std::pow(ave_blks[i], 2);
  const double ave_sq = stat_mean(sq.data(), (int64_t)sq.size());
  const double ave = stat_mean(ave_blks.data(), (int64_t)ave_blks.size());
  return std::sqrt((ave_sq - std::pow(ave, 2)) / (double)(nblocks - 1));
}

}  // namespace orc
