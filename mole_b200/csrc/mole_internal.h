// Internal structures shared by the CUDA translation units and the host-side C++.
// Product code: never includes anything from oracle/.
#pragma once
#include <cstdint>
#include <string>
#include <vector>
#include "../../include/mole_b200.h"

#if defined(__CUDACC__)
#define MOLE_HD __host__ __device__ __forceinline__
#define MOLE_D __device__ __forceinline__
#else
#define MOLE_HD inline
#define MOLE_D inline
#endif
// dynamic shared memory of a kernel; MOLE_EMU = the host emulation of tests/native/cuda_emu.h (test infrastructure:
// the kernels are compiled unchanged by g++ and run thread-per-lane against the oracle, no GPU needed)
#if !defined(MOLE_CALLER_LINE)
#define MOLE_CALLER_LINE 0
#endif
#if defined(MOLE_EMU)
#define MOLE_DYN_SMEM(T, name) T* name = (T*)mole_emu_dyn_smem
#define MOLE_DEVICE_CODE 1
#else
#define MOLE_DYN_SMEM(T, name) extern __shared__ T name[]
#if defined(__CUDACC__)
#define MOLE_DEVICE_CODE 1
#endif
#endif

// NVTX ranges around the host entry points that enqueue device work (sweep, reduce, DMC step, branch, collectives),
// so that nsys / ncu --nvtx captures are attributable; compiled in with -DMOLE_NVTX (build.sh: MOLE_NVCC_EXTRA)
#if defined(MOLE_NVTX) && !defined(MOLE_EMU)
#include <nvtx3/nvToolsExt.h>
struct MoleRange {
  explicit MoleRange(const char* name) { nvtxRangePushA(name); }
  ~MoleRange() { nvtxRangePop(); }
};
#define MOLE_RANGE(name) MoleRange mole_range_guard_(name)
#else
#define MOLE_RANGE(name) do { } while (0)
#endif

// ---- packed accumulator layout (doubles); mirrors mole_acc_host field order -----------------
enum {
  ACC_N = 0, ACC_E = 1, ACC_E2 = 2, ACC_B = 3, ACC_B2 = 4, ACC_NB = 5, ACC_NACC = 6, ACC_NMOVE = 7,
  ACC_T = 8, ACC_PSI = 9,
  ACC_O = 10,                                  // 8 slots
  ACC_OE = ACC_O + MOLE_ACC_MAX_PARAMS,        // 8 slots
  ACC_OO = ACC_OE + MOLE_ACC_MAX_PARAMS,       // 36 slots, (k<=l) packed row-major over the ACTUAL P
  ACC_LEN = ACC_OO + MOLE_ACC_MAX_PARAMS * (MOLE_ACC_MAX_PARAMS + 1) / 2,
  // health counters behind the packed moments (device vector only; mole_ensemble_health reads them): samples whose
  // E_L or O_k was not finite are counted and kept OUT of every sum, DMC walkers with a non-finite E_L are killed
  ACC_BAD = ACC_LEN, ACC_BAD_DMC = ACC_LEN + 1,
  ACC_DEV_LEN = ACC_LEN + 2
};

// ---- POD parameter blocks passed to kernels by value -------------------------------------------
struct WfParams {
  int32_t kind, ne, np, pad;
  double p[MOLE_WF_MAX_PARAMS];
  double geom[MOLE_WF_MAX_GEOM];
};

struct HamParams {
  int32_t kind, n_ions;
  double ion_pos[MOLE_OP_MAX_IONS * 3];
  double ion_z[MOLE_OP_MAX_IONS];
  double ionic_repulsion;  // IonicPotential::new precomputes it (operator.rs:40-55)
  double frequency;
};

struct RngKey { uint32_t k0, k1; };

struct SweepParams {
  // ensemble
  double* x;          // [3*ne][W]
  double* blk;        // [W] partial block sums
  double* acc;        // [ACC_LEN] global accumulators
  double* partials;   // [grid][ACC_LEN]
  unsigned int* ticket;
  int64_t W;
  uint64_t walker_offset;
  RngKey key;
  uint32_t step0;
  // schedule
  int32_t n_sweeps, n_discard, block_size, blk_fill;
  uint32_t observables, compat;
  double metrop_param;   // box side | tau
  // traces (device pointers, nullable)
  double* tr_energy; double* tr_wfvalue; double* tr_kinetic; double* tr_pgrad; uint8_t* tr_accept;
  double* osamp;      // kinds with P > MOLE_ACC_MAX_PARAMS: per-sample rows (1, E_L, O_1..O_P), [sample][col][W]
  WfParams wf;
  HamParams ham;
};

struct DmcParams {
  double* x; double* w; double* el; uint8_t* el_valid_flag;
  double* red;        // [4] sum_w_e, sum_w, sum_w_new, max_w_new (atomics-free: partials + ticket)
  double* partials; unsigned int* ticket;
  double* health;     // &acc[ACC_BAD_DMC]
  int64_t W; uint64_t walker_offset; RngKey key; uint32_t step;
  double tau_move, tau_weight, e_ref;
  int32_t el_cached;
  uint32_t compat;
  WfParams wf; HamParams ham;
};

// ---- host-side objects behind the opaque handles ----------------------------------------------
struct mole_ctx_s {
  int device = -1;
  void* stream = nullptr;   // cudaStream_t
  int sm_count = 0;
  std::string last_error;
  int64_t launches = 0;
  // NCCL (dlopen'ed lazily, see mole_comm.cpp)
  void* nccl_comm = nullptr;
  double* comm_scratch = nullptr;  // device, 16 doubles
  int nranks = 1, rank = 0;
  int p2p_warm = 0;                // every send/recv pair of the communicator has been opened (mole_comm_warm_p2p)
  // ensembles keep a pointer to their context: a context destroyed while ensembles are alive is only marked and
  // is freed with the last of them (bindings with garbage-collected wrappers destroy in arbitrary order)
  int live_ens = 0;
  bool closing = false;
};
// FNV-1a over the parameter blocks: the cached local energy of a DMC ensemble is only reused for the same
// wavefunction parameters and operator (dmc.rs:89-96 recomputes it every step)
uint64_t mole_el_signature(const WfParams& w, const HamParams& h);

struct mole_wf_s { mole_ctx_s* ctx; WfParams p; };
struct mole_op_s { mole_ctx_s* ctx; HamParams p; };
struct mole_metrop_s { int32_t kind; double param; uint32_t compat; };

struct mole_ens_s {
  mole_ctx_s* ctx;
  int64_t W; int32_t ne; uint64_t walker_offset;
  RngKey key; uint32_t step;
  double* x = nullptr;      // [3*ne][W]
  double* x2 = nullptr;     // gather target for branching (swapped with x)
  double* x0 = nullptr;     // snapshot (lazily allocated)
  double* w = nullptr; double* w2 = nullptr;
  double* el = nullptr; double* el2 = nullptr; int el_cached = 0;
  uint64_t el_sig = 0;       // signature of the (wavefunction parameters, operator) the cached E_L was computed with
  int wstats_valid = 0;      // red[2..3] hold sum/max of the CURRENT weights
  int w_uniform = 1;         // every walker of this rank carries the same weight (after creation / SR branching)
  double* xchg = nullptr; size_t xchg_cap = 0;   // send | receive rows of mole_rebalance
  double* blk = nullptr; int32_t blk_fill = 0; int32_t blk_size = 0;
  double* acc = nullptr;    // [ACC_LEN]
  double* partials = nullptr; int partial_rows = 0;
  unsigned int* ticket = nullptr;
  double* red = nullptr;    // [8] dmc reductions
  unsigned long long* cum = nullptr;  // [W] branching prefix sums
  unsigned long long* blocksums = nullptr; int n_scan_blocks = 0;
  int32_t* src = nullptr;   // [W] branching source indices
  int32_t np_last = 0;      // P of the last sweep that touched acc
  double* series = nullptr; // [series_n][W] E_L samples kept on the device (MOLE_SWEEP_KEEP_SERIES)
  int64_t series_n = 0, series_cap = 0;
  double* step_e = nullptr; // [step_e_cap] per-step ensemble energies of a DMC block (mole_dmc_block)
  int64_t step_e_cap = 0;
  double* gath = nullptr;   // [nranks][4] all-gathered per-step DMC reductions (multi-rank block loop)
  int32_t* sb_list = nullptr; int32_t* sb_fen = nullptr; uint32_t* sb_mask = nullptr;   // SimpleBranching scratch
  int64_t* sb_draws = nullptr;
  // kinds with P > MOLE_ACC_MAX_PARAMS: per-sample rows (1, E_L, O_k) and their Gram matrix (mole_gram.cuh)
  double* osamp = nullptr; size_t osamp_cap = 0;       // doubles
  double* gram = nullptr;                               // [48 x 48], upper triangle
  double* gram_partials = nullptr; int gram_rows = 0;   // per-CTA partial matrices
  int32_t gram_cols = 0;                                // P + 2 of the last sweep that touched gram
  unsigned int* bar = nullptr;                          // [2] grid-barrier counter + time-out flag of dmc_block_kernel
  unsigned long long* vb_sums = nullptr;                // [n_vb + 1] / [8 n_vb]: prefix-sum levels of dmc_block_kernel
  unsigned long long* vb_coarse = nullptr;
  std::vector<double> island_cnt;                       // walker counts of all ranks (gathered once)
  std::vector<double> island_sw;                        // per-rank sum of weights at the last step of the last SR block
  int32_t dmc_block_impl = 0;                           // 0: one persistent launch per block where eligible, 1: per-step launches
  int32_t gram_impl = 0;                                // 0: DMMA (tensor cores), 1: FP64 vector pipe
};

// ---- host-only helpers (mole_host.cpp) ---------------------------------------------------------
RngKey mole_key_from_seed(const uint8_t seed[32]);
int mole_set_error(mole_ctx_s* ctx, int code, const std::string& msg);
int mole_oo_index(int P, int k, int l);  // k<=l
// mole_comm.cpp: sum/max allreduce of a few host scalars over the ctx communicator (no-op for one rank)
int32_t mole_comm_allreduce_host(mole_ctx_s* ctx, double* sum_vals, int n_sum, double* max_vals, int n_max);
// all-gather of n device doubles per rank on the context stream, without a host synchronisation
int32_t mole_comm_allgather_device(mole_ctx_s* ctx, const double* send_dev, double* recv_dev, int n);
int32_t mole_comm_allgather_host(mole_ctx_s* ctx, const double* mine, int n, double* all);
struct MoleMove { int src, dst; int64_t send_first, recv_first, count; };
int32_t mole_comm_exchange_rows(mole_ctx_s* ctx, const double* send_dev, double* recv_dev, int row_len, const std::vector<MoleMove>& moves);
// population shares and the transfer plan of mole_rebalance (pure host arithmetic, identical on every rank)
void mole_rebalance_shares(int nranks, const double* totals, const int64_t* counts, double u, int64_t* shares);
// one 8-byte exchange between every pair of ranks: NCCL connects peers lazily, ~0.15 s for the first send/recv of a pair
int32_t mole_comm_warm_p2p(mole_ctx_s* ctx);
std::vector<MoleMove> mole_rebalance_moves(int nranks, const int64_t* counts, const int64_t* shares);
