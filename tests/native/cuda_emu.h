// Host emulation of the CUDA execution model, just enough to run this repo's cooperative kernels UNCHANGED on
// the CPU (test infrastructure only: the product has no CPU path).  One std::thread per CUDA thread of ONE block at
// a time; __syncwarp / __syncthreads are barriers, warp shuffles and ballots go through a per-warp scratch row.
// This checks the LOGIC of a kernel (indexing, mailbox protocol, synchronisation placement, reductions) against the
// oracle before GPU minutes are spent; bit patterns differ from the device only through the reciprocal / rsqrt seeds.
//
// Usage: #define MOLE_EMU, include this header BEFORE the kernel headers, then
//   emu::launch(grid, block, dyn_smem_bytes, [&] { kernel(args); });
#pragma once
#include <algorithm>
#include <atomic>
#include <cmath>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <functional>
#include <mutex>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <unistd.h>
#include <vector>

#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __restrict__
#define __shared__ static
#define __constant__ static
#define __launch_bounds__(...)
#define __maxnreg__(...)

namespace emu {
struct dim3 { unsigned x = 1, y = 1, z = 1; };

class Barrier {   // reusable counting barrier (generation counter)
 public:
  void reset(int n) { n_ = n; count_ = 0; gen_ = 0; }
  void wait() {
    std::unique_lock<std::mutex> lk(m_);
    const unsigned g = gen_;
    if (++count_ == n_) { count_ = 0; ++gen_; cv_.notify_all(); }
    else cv_.wait(lk, [&] { return gen_ != g; });
  }
 private:
  std::mutex m_;
  std::condition_variable cv_;
  int n_ = 1, count_ = 0;
  unsigned gen_ = 0;
};

struct Block {
  Barrier block_bar;
  std::vector<Barrier> warp_bar;
  std::vector<uint64_t> shfl;   // [warp][32]
  std::vector<unsigned char> dyn_smem;
};
inline Block& blk() { static Block b; return b; }
}  // namespace emu

static thread_local emu::dim3 threadIdx, blockIdx;
static emu::dim3 blockDim, gridDim;
static void* mole_emu_dyn_smem = nullptr;

// last synchronisation site of every emulated thread (source line of the caller), for the deadlock watchdog
static int emu_site[1024];
static int emu_hist[1024][12];
static int emu_nsync[1024];
inline void emu_note(int line) {
  const unsigned t = threadIdx.x;
  emu_site[t] = line;
  emu_hist[t][emu_nsync[t] % 12] = line;
  ++emu_nsync[t];
}
#define MOLE_CALLER_LINE __builtin_LINE()
inline void __syncthreads(int line = __builtin_LINE()) { emu_note(-line); emu::blk().block_bar.wait(); }
inline void __syncwarp(unsigned = 0xffffffffu, int line = __builtin_LINE()) {
  emu_note(line);
  emu::blk().warp_bar[threadIdx.x >> 5].wait();
}
inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }

template <class T>
inline T emu_exchange(T v, int src_lane) {
  static_assert(sizeof(T) <= 8, "shuffle payload");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint64_t* row = emu::blk().shfl.data() + (size_t)warp * 32;
  uint64_t bits = 0;
  std::memcpy(&bits, &v, sizeof(T));
  row[lane] = bits;
  __syncwarp();
  const uint64_t got = row[src_lane & 31];
  __syncwarp();
  T out;
  std::memcpy(&out, &got, sizeof(T));
  return out;
}
template <class T> inline T __shfl_sync(unsigned, T v, int src) { return emu_exchange(v, src); }
template <class T> inline T __shfl_xor_sync(unsigned, T v, int m) { return emu_exchange(v, (int)(threadIdx.x & 31) ^ m); }
template <class T> inline T __shfl_up_sync(unsigned, T v, int d) {
  const int lane = threadIdx.x & 31;
  const T got = emu_exchange(v, lane >= d ? lane - d : lane);
  return got;
}
inline unsigned __ballot_sync(unsigned, bool p) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint64_t* row = emu::blk().shfl.data() + (size_t)warp * 32;
  row[lane] = p ? 1u : 0u;
  __syncwarp();
  unsigned m = 0;
  const int n = (int)std::min<unsigned>(32u, blockDim.x - 32u * warp);
  for (int i = 0; i < n; ++i) m |= (unsigned)(row[i] & 1u) << i;
  __syncwarp();
  return m;
}
inline bool __any_sync(unsigned m, bool p) { return __ballot_sync(m, p) != 0u; }

inline int __double2loint(double x) { uint64_t b; std::memcpy(&b, &x, 8); return (int)(uint32_t)b; }
inline int __double2hiint(double x) { uint64_t b; std::memcpy(&b, &x, 8); return (int)(uint32_t)(b >> 32); }
inline double __hiloint2double(int hi, int lo) {
  const uint64_t b = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo;
  double x; std::memcpy(&x, &b, 8); return x;
}
inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
inline unsigned long long __umul64hi(unsigned long long a, unsigned long long b) { return (unsigned long long)(((unsigned __int128)a * b) >> 64); }
template <class T> inline T __ldcg(const T* p) { return *p; }
using std::fma; using std::fmax; using std::fmin; using std::isfinite; using std::isnan; using std::signbit; using std::sqrt; using std::exp; using std::fabs;
inline int max(int a, int b) { return a > b ? a : b; }
inline int min(int a, int b) { return a < b ? a : b; }

static std::mutex emu_atomic_mutex;
inline unsigned atomicInc(unsigned* p, unsigned lim) { std::lock_guard<std::mutex> g(emu_atomic_mutex); const unsigned o = *p; *p = (o >= lim) ? 0u : o + 1u; return o; }
inline unsigned atomicAdd(unsigned* p, unsigned v) { std::lock_guard<std::mutex> g(emu_atomic_mutex); const unsigned o = *p; *p = o + v; return o; }
inline double atomicAdd(double* p, double v) { std::lock_guard<std::mutex> g(emu_atomic_mutex); const double o = *p; *p = o + v; return o; }

// MUFU.RCP64H / MUFU.RSQ64H stand-ins: ~20 good mantissa bits (the low word is zero on the device)
inline double emu_seed20(double y) {
  uint64_t b; std::memcpy(&b, &y, 8);
  b &= 0xffffffff00000000ull;
  std::memcpy(&y, &b, 8);
  return y;
}
inline double mole_emu_rcp_seed(double x) { return emu_seed20(1.0 / x); }
inline double mole_emu_rsqrt_seed(double x) { return emu_seed20(1.0 / std::sqrt(x)); }

namespace emu {
// run `body` as every thread of every block of a (grid, block) launch; blocks run one after the other
inline void launch(unsigned grid, unsigned block, size_t dyn_smem, const std::function<void()>& body) {
  Block& b = blk();
  const int nwarp = (int)((block + 31) / 32);
  ::blockDim.x = block;
  ::gridDim.x = grid;
  b.dyn_smem.assign(dyn_smem + 16, 0xff);   // NaN-ish fill: a kernel must initialise what it reads
  mole_emu_dyn_smem = b.dyn_smem.data();
  b.shfl.assign((size_t)nwarp * 32, 0);
  for (unsigned bx = 0; bx < grid; ++bx) {
    b.block_bar.reset((int)block);
    b.warp_bar = std::vector<Barrier>(nwarp);
    for (int w = 0; w < nwarp; ++w) b.warp_bar[w].reset((int)std::min<unsigned>(32u, block - 32u * w));
    std::vector<std::thread> th;
    th.reserve(block);
    std::atomic<bool> done{false};
    std::thread dog([&] {   // deadlock watchdog: MOLE_EMU_WATCHDOG seconds (default 60), then the threads' last sync sites
      const char* w = getenv("MOLE_EMU_WATCHDOG");
      const int limit = w ? atoi(w) : 60;
      for (int t = 0; t < limit * 10 && !done.load(); ++t) std::this_thread::sleep_for(std::chrono::milliseconds(100));
      if (done.load()) return;
      fprintf(stderr, "cuda_emu: block %u did not finish in %d s; last sync site (source line, <0: __syncthreads) per thread:\n", bx, limit);
      for (unsigned t = 0; t < block; ++t) fprintf(stderr, "%s%d", t % 16 ? " " : "\n  ", emu_site[t]);
      fprintf(stderr, "\nsync counts and the last 12 sites of threads 0 and %u:\n", block / 2);
      for (unsigned t : {0u, block / 2}) {
        fprintf(stderr, "  thread %u: %d syncs:", t, emu_nsync[t]);
        for (int i = 0; i < 12; ++i) fprintf(stderr, " %d", emu_hist[t][(emu_nsync[t] + i) % 12]);
        fprintf(stderr, "\n");
      }
      _exit(3);
    });
    for (unsigned t = 0; t < block; ++t)
      th.emplace_back([&, t, bx] {
        ::threadIdx.x = t;
        ::blockIdx.x = bx;
        body();
      });
    for (auto& x : th) x.join();
    done.store(true);
    dog.join();
  }
}
}  // namespace emu
