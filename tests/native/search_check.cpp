// Host harness for mole_b200/csrc/mole_search.h: the two-level 4-ary pick the SR branching kernel runs must return
// the walker std::upper_bound finds on the GLOBAL inclusive prefix sums (= rand 0.5's WeightedChoice as
// src/dmc/src/branching.rs:27-36 uses it).  Built and run by tests/test_search_host.py; prints "ok <cases> <draws>".
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <random>
#include <vector>

#include "mole_search.h"

static long check(int64_t W, int tile, const std::vector<unsigned long long>& k, std::mt19937_64& rng, int extra) {
  const int n_tiles = (int)((W + tile - 1) / tile);
  std::vector<unsigned long long> cum(W), glob(W), ts(n_tiles + 1, 0);
  unsigned long long run = 0;
  for (int t = 0; t < n_tiles; ++t) {
    ts[t] = run;
    unsigned long long loc = 0;
    for (int64_t i = (int64_t)t * tile; i < std::min<int64_t>((int64_t)(t + 1) * tile, W); ++i) {
      loc += k[i];
      cum[i] = loc;
      glob[i] = run + loc;
    }
    run += loc;
  }
  ts[n_tiles] = run;
  if (run == 0) return 0;   // the library refuses an all-zero ensemble before the kernel runs
  std::vector<unsigned long long> draws;
  // every boundary of the cumulative distribution and its neighbours, plus random draws
  for (int64_t i = 0; i < W && (int64_t)draws.size() < 60000; ++i)
    for (long d = -1; d <= 1; ++d) {
      const unsigned long long u = glob[i] + (unsigned long long)d;
      if (u < run) draws.push_back(u);
    }
  draws.push_back(0);
  draws.push_back(run - 1);
  std::uniform_int_distribution<unsigned long long> U(0, run - 1);
  for (int i = 0; i < extra; ++i) draws.push_back(U(rng));
  for (unsigned long long u : draws) {
    const int64_t want = std::upper_bound(glob.begin(), glob.end(), u) - glob.begin();
    const int64_t got = mole_pick_tiled(cum.data(), ts.data(), n_tiles, W, tile, u);
    if (got != want || k[got] == 0) {
      std::printf("MISMATCH W=%lld tile=%d u=%llu got=%lld want=%lld\n", (long long)W, tile, u, (long long)got, (long long)want);
      std::exit(1);
    }
  }
  return (long)draws.size();
}

int main() {
  std::mt19937_64 rng(20261017);
  long cases = 0, draws = 0;
  const int64_t sizes[] = {1, 2, 3, 4, 5, 7, 100, 1023, 1024, 1025, 2048, 4097, 32768, 100003};
  const int tiles[] = {1024, 4, 5, 16};   // 1024 = SCAN_TILE of the kernels; small tiles exercise the tile search
  for (int64_t W : sizes)
    for (int tile : tiles) {
      if ((W + tile - 1) / tile > 40000) continue;
      for (int pattern = 0; pattern < 6; ++pattern) {
        std::vector<unsigned long long> k(W);
        for (int64_t i = 0; i < W; ++i) {
          const unsigned long long r = rng();
          switch (pattern) {
            case 0: k[i] = 1 + r % W; break;                                    // SRBrancher range, all alive
            case 1: k[i] = (r & 3) ? 0 : 1 + (r >> 8) % 1000; break;            // 3/4 dead walkers
            case 2: k[i] = ((i / tile) & 1) ? 0 : r % 7; break;                 // whole tiles empty
            case 3: k[i] = i == (int64_t)(r % W) || i == W - 1 ? (unsigned long long)W : 0; break;   // almost everything dead
            case 4: k[i] = (unsigned long long)W; break;                        // N^2 totals (2^36 at 2^18: u64 needed)
            default: k[i] = i == 0 ? 1 : 0; break;                              // one survivor, first walker
          }
        }
        draws += check(W, tile, k, rng, 2000);
        ++cases;
      }
    }
  std::printf("ok %ld %ld\n", cases, draws);
  return 0;
}
