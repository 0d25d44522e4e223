// C ABI (include/mole_b200.h): device-touching entry points.  Host-only logic (optimizers,
// finaliser, drivers) lives in mole_host.cpp.  There is no CPU fallback anywhere in this file:
// every entry point either launches the sm_100a kernels or returns an error status.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <algorithm>
#include <string>
#include <vector>
#include "mole_internal.h"
#include "mole_kernels.cuh"
#include "mole_branch.cuh"
#include "mole_sj.cuh"
#include "mole_lsj.cuh"
#include "mole_gram.cuh"
#include "mole_dmc_block.cuh"
#include "mole_stats.cuh"

// errors raised without a context (NULL handles, mole_ctx_create); per thread, because different contexts may be
// driven from different threads (include/mole_b200.h)
static thread_local std::string g_last_error;

int mole_set_error(mole_ctx_s* ctx, int code, const std::string& msg) {
  if (ctx) ctx->last_error = msg;
  g_last_error = msg;
  return code;
}

#define CU(ctx, call)                                                                                   \
  do {                                                                                                  \
    cudaError_t e__ = (call);                                                                           \
    if (e__ != cudaSuccess)                                                                             \
      return mole_set_error(ctx, e__ == cudaErrorNoDevice || e__ == cudaErrorInsufficientDriver         \
                                     ? MOLE_ERR_NO_DEVICE : MOLE_ERR_CUDA,                              \
                            std::string(#call) + ": " + cudaGetErrorString(e__));                       \
  } while (0)

#define STREAM(ctx) ((cudaStream_t)(ctx)->stream)
#define KERNEL_CHECK(ctx)                 \
  do {                                    \
    (ctx)->launches++;                    \
    CU(ctx, cudaGetLastError());          \
  } while (0)

static inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

extern "C" {

int32_t mole_version(void) { return 100; }

const char* mole_last_error_string(mole_ctx_t ctx) { return ctx ? ctx->last_error.c_str() : g_last_error.c_str(); }

// ------------------------------------------------------------------ context
int32_t mole_ctx_create(int32_t device, mole_ctx_t* out) {
  if (!out) return mole_set_error(nullptr, MOLE_ERR_INVALID_ARG, "ctx out pointer is NULL");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    return mole_set_error(nullptr, MOLE_ERR_NO_DEVICE,
                          std::string("no CUDA device: the mole_b200 hot path has no CPU fallback (") +
                              cudaGetErrorString(e) + ")");
  if (device < 0 || device >= n) return mole_set_error(nullptr, MOLE_ERR_INVALID_ARG, "bad device index");
  CU(nullptr, cudaSetDevice(device));
  cudaDeviceProp prop;
  CU(nullptr, cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return mole_set_error(nullptr, MOLE_ERR_NO_DEVICE,
                          "device is not sm_100 (B200): this library ships sm_100a code only");
  cudaStream_t s;
  CU(nullptr, cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  // opt-ins to more than 48 KB of dynamic shared memory are per device: set them when the context is made, not
  // behind process-wide flags (a second context on another GPU, or on another thread, needs them too)
  CU(nullptr, sj_set_kernel_attributes());
  CU(nullptr, cudaFuncSetAttribute(simple_remove_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  mole_ctx_s* c = new mole_ctx_s();
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  c->stream = (void*)s;
  *out = c;
  return MOLE_OK;
}

static void ctx_free(mole_ctx_s* ctx) {
  cudaSetDevice(ctx->device);
  mole_comm_destroy(ctx);                       // communicator and its scratch, if mole_comm_init was called
  if (ctx->stream) cudaStreamDestroy(STREAM(ctx));
  delete ctx;
}

int32_t mole_ctx_destroy(mole_ctx_t ctx) {
  if (!ctx) return MOLE_OK;
  if (ctx->live_ens > 0) { ctx->closing = true; return MOLE_OK; }   // freed by the last mole_ensemble_destroy
  ctx_free(ctx);
  return MOLE_OK;
}

int32_t mole_ctx_synchronize(mole_ctx_t ctx) {
  if (!ctx) return MOLE_ERR_INVALID_ARG;
  CU(ctx, cudaSetDevice(ctx->device));
  CU(ctx, cudaStreamSynchronize(STREAM(ctx)));
  return MOLE_OK;
}

int32_t mole_ctx_stream(mole_ctx_t ctx, void** stream) {
  if (!ctx || !stream) return MOLE_ERR_INVALID_ARG;
  *stream = ctx->stream;
  return MOLE_OK;
}

int32_t mole_ctx_launch_count(mole_ctx_t ctx, int64_t* n) {
  if (!ctx || !n) return MOLE_ERR_INVALID_ARG;
  *n = ctx->launches;
  return MOLE_OK;
}

// ------------------------------------------------------------------ wavefunction descriptors
static int wf_kind_ne(const mole_wf_desc* d) {
  switch (d->kind) {
    case MOLE_WF_STO_1S: case MOLE_WF_GAUSSIAN: case MOLE_WF_H2P_PRODUCT: case MOLE_WF_CONSTANT: case MOLE_WF_LCAO_1E_2C: return 1;
    case MOLE_WF_STO_PRODUCT: case MOLE_WF_H2_HL_STO: case MOLE_WF_LCAO_2E_1C: case MOLE_WF_LCAO_2E_2C: return 2;
    case MOLE_WF_SLATER_JASTROW: case MOLE_WF_LCAO_SJ: return (int)d->geom[1] + (int)d->geom[2];
  }
  return -1;
}
static int wf_kind_np(const mole_wf_desc* d) {
  const int kind = d->kind;
  switch (kind) {
    case MOLE_WF_STO_1S: case MOLE_WF_GAUSSIAN: case MOLE_WF_STO_PRODUCT: case MOLE_WF_H2_HL_STO: return 1;
    case MOLE_WF_H2P_PRODUCT: case MOLE_WF_CONSTANT: return 0;
    case MOLE_WF_SLATER_JASTROW: return 7;
    case MOLE_WF_LCAO_1E_2C: case MOLE_WF_LCAO_2E_1C: return 2;   // coefficients C[k][c]
    case MOLE_WF_LCAO_2E_2C: return 4;
    case MOLE_WF_LCAO_SJ: return std::max((int)d->geom[1], (int)d->geom[2]) * (int)d->geom[3] + 4;
  }
  return -1;
}

int32_t mole_wf_create(mole_ctx_t ctx, const mole_wf_desc* d, mole_wf_t* out) {
  if (!ctx || !d || !out) return mole_set_error(ctx, MOLE_ERR_INVALID_ARG, "mole_wf_create: NULL argument");
  const int ne = wf_kind_ne(d), np = wf_kind_np(d);
  if (ne < 0) return mole_set_error(ctx, MOLE_ERR_INVALID_ARG, "unknown wavefunction kind");
  if (d->n_elec != ne) return mole_set_error(ctx, MOLE_ERR_SHAPE, "n_elec does not match the wavefunction kind");
  if (d->n_params != np) return mole_set_error(ctx, MOLE_ERR_SHAPE, "n_params does not match the wavefunction kind");
  if (d->kind == MOLE_WF_SLATER_JASTROW) {
    const int nu = (int)d->geom[1], nd = (int)d->geom[2];
    if (nu < 0 || nd < 0 || nu > 5 || nd > 5 || nu + nd < 1 || !(d->geom[0] > 0.0))
      return mole_set_error(ctx, MOLE_ERR_INVALID_ARG, "Slater-Jastrow needs kappa>0 and 0<=n_up,n_dn<=5");
  }
  if (d->kind == MOLE_WF_LCAO_SJ) {
    const int nu = (int)d->geom[1], nd = (int)d->geom[2], nc = (int)d->geom[3];
    bool ok = nu >= 0 && nd >= 0 && nu <= 5 && nd <= 5 && nu + nd >= 1 && nc >= 1 && nc <= 8 && d->geom[0] > 0.0;
    for (int c = 0; ok && c < nc; ++c) ok = d->geom[8 + 4 * c + 3] > 0.0;
    if (!ok) return mole_set_error(ctx, MOLE_ERR_INVALID_ARG, "LCAO Slater-Jastrow needs kappa>0, 0<=n_up,n_dn<=5, 1<=N_c<=8 and alpha_c = 1/width_c > 0");
  }
  if (d->kind >= MOLE_WF_LCAO_1E_2C && d->kind <= MOLE_WF_LCAO_2E_2C) {
    if (!(d->geom[1] > 0.0) || (d->geom[0] != 0.0 && d->geom[0] != 1.0))
      return mole_set_error(ctx, MOLE_ERR_INVALID_ARG, "LCAO needs alpha = 1/width > 0 and mode 0 (spin product) or 1 (determinant)");
  }
  mole_wf_s* w = new mole_wf_s();
  w->ctx = ctx;
  w->p.kind = d->kind; w->p.ne = ne; w->p.np = np; w->p.pad = 0;
  for (int i = 0; i < MOLE_WF_MAX_PARAMS; ++i) w->p.p[i] = d->params[i];
  for (int i = 0; i < MOLE_WF_MAX_GEOM; ++i) w->p.geom[i] = d->geom[i];
  *out = w;
  return MOLE_OK;
}
int32_t mole_wf_destroy(mole_wf_t wf) { delete wf; return MOLE_OK; }
int32_t mole_wf_num_electrons(mole_wf_t wf, int32_t* n) { if (!wf || !n) return MOLE_ERR_INVALID_ARG; *n = wf->p.ne; return MOLE_OK; }
int32_t mole_wf_num_parameters(mole_wf_t wf, int32_t* n) { if (!wf || !n) return MOLE_ERR_INVALID_ARG; *n = wf->p.np; return MOLE_OK; }
int32_t mole_wf_get_parameters(mole_wf_t wf, double* p) {
  if (!wf || !p) return MOLE_ERR_INVALID_ARG;
  // H2P carries alpha in params[0] although it exposes no Optimize impl
  for (int i = 0; i < wf->p.np; ++i) p[i] = wf->p.p[i];
  return MOLE_OK;
}
int32_t mole_wf_update_parameters(mole_wf_t wf, const double* dp) {
  if (!wf || !dp) return MOLE_ERR_INVALID_ARG;
  for (int i = 0; i < wf->p.np; ++i) wf->p.p[i] += dp[i];
  return MOLE_OK;
}
int32_t mole_wf_set_parameters(mole_wf_t wf, const double* p) {
  if (!wf || !p) return MOLE_ERR_INVALID_ARG;
  for (int i = 0; i < wf->p.np; ++i) wf->p.p[i] = p[i];
  return MOLE_OK;
}

// ------------------------------------------------------------------ operators
int32_t mole_op_create(mole_ctx_t ctx, const mole_op_desc* d, mole_op_t* out) {
  if (!ctx || !d || !out) return mole_set_error(ctx, MOLE_ERR_INVALID_ARG, "mole_op_create: NULL argument");
  if (d->kind < MOLE_OP_KINETIC || d->kind > MOLE_OP_HARMONIC) return mole_set_error(ctx, MOLE_ERR_INVALID_ARG, "unknown operator kind");
  if (d->n_ions < 0 || d->n_ions > MOLE_OP_MAX_IONS) return mole_set_error(ctx, MOLE_ERR_SHAPE, "too many ions");
  mole_op_s* o = new mole_op_s();
  o->ctx = ctx;
  o->p.kind = d->kind; o->p.n_ions = d->n_ions; o->p.frequency = d->frequency;
  for (int i = 0; i < MOLE_OP_MAX_IONS * 3; ++i) o->p.ion_pos[i] = d->ion_pos[i];
  for (int i = 0; i < MOLE_OP_MAX_IONS; ++i) o->p.ion_z[i] = (double)d->ion_charge[i];
  double pot = 0.0;  // IonicPotential::new, operator.rs:40-55
  for (int i = 0; i < d->n_ions; ++i)
    for (int j = i + 1; j < d->n_ions; ++j) {
      double s2 = 0.0;
      for (int k = 0; k < 3; ++k) { const double t = d->ion_pos[3 * j + k] - d->ion_pos[3 * i + k]; s2 += t * t; }
      pot += (double)(d->ion_charge[i] * d->ion_charge[j]) / std::sqrt(s2);
    }
  o->p.ionic_repulsion = pot;
  *out = o;
  return MOLE_OK;
}
int32_t mole_op_destroy(mole_op_t op) { delete op; return MOLE_OK; }

// ------------------------------------------------------------------ ensemble
static int ens_grid_rows(mole_ctx_s* ctx) { return ctx->sm_count * 16; }

int32_t mole_ensemble_create(mole_ctx_t ctx, int64_t W, int32_t ne, const uint8_t seed[32], uint64_t walker_offset,
                             mole_ens_t* out) {
  if (!ctx || !out || !seed) return mole_set_error(ctx, MOLE_ERR_INVALID_ARG, "mole_ensemble_create: NULL argument");
  if (W < 1 || W > (int64_t)1 << 30 || ne < 1 || ne > MOLE_WF_MAX_ELEC)
    return mole_set_error(ctx, MOLE_ERR_SHAPE, "ensemble shape out of range");
  CU(ctx, cudaSetDevice(ctx->device));
  mole_ens_s* e = new mole_ens_s();
  e->ctx = ctx; e->W = W; e->ne = ne; e->walker_offset = walker_offset;
  e->key = mole_key_from_seed(seed); e->step = 0;
  const size_t nx = (size_t)3 * ne * W;
  CU(ctx, cudaMalloc(&e->x, nx * sizeof(double)));
  CU(ctx, cudaMalloc(&e->x2, nx * sizeof(double)));
  CU(ctx, cudaMalloc(&e->w, W * sizeof(double)));
  CU(ctx, cudaMalloc(&e->w2, W * sizeof(double)));
  CU(ctx, cudaMalloc(&e->el, W * sizeof(double)));
  CU(ctx, cudaMalloc(&e->el2, W * sizeof(double)));
  CU(ctx, cudaMalloc(&e->blk, W * sizeof(double)));
  CU(ctx, cudaMalloc(&e->acc, ACC_DEV_LEN * sizeof(double)));
  e->partial_rows = ens_grid_rows(ctx);
  CU(ctx, cudaMalloc(&e->partials, (size_t)e->partial_rows * ACC_LEN * sizeof(double)));
  CU(ctx, cudaMalloc(&e->ticket, 2 * sizeof(unsigned int)));   // [0] accumulator reduction, [1] branching scan
  CU(ctx, cudaMalloc(&e->red, 8 * sizeof(double)));
  e->n_scan_blocks = cdiv(W, SCAN_TILE);
  CU(ctx, cudaMalloc(&e->cum, (size_t)W * sizeof(unsigned long long)));
  CU(ctx, cudaMalloc(&e->blocksums, (size_t)(e->n_scan_blocks + 1) * sizeof(unsigned long long)));
  CU(ctx, cudaMalloc(&e->src, (size_t)W * sizeof(int32_t)));
  CU(ctx, cudaMemsetAsync(e->x, 0, nx * sizeof(double), STREAM(ctx)));
  CU(ctx, cudaMemsetAsync(e->blk, 0, W * sizeof(double), STREAM(ctx)));
  CU(ctx, cudaMemsetAsync(e->acc, 0, ACC_DEV_LEN * sizeof(double), STREAM(ctx)));
  CU(ctx, cudaMemsetAsync(e->ticket, 0, 2 * sizeof(unsigned int), STREAM(ctx)));
  fill_kernel<<<cdiv(W, 256), 256, 0, STREAM(ctx)>>>(e->w, W, 1.0);   // dmc.rs:51 initial weight 1.0
  KERNEL_CHECK(ctx);
  ++ctx->live_ens;
  *out = e;
  return MOLE_OK;
}

int32_t mole_ensemble_destroy(mole_ens_t e) {
  if (!e) return MOLE_OK;
  cudaSetDevice(e->ctx->device);
  cudaStreamSynchronize(STREAM(e->ctx));
  cudaFree(e->x0);
  cudaFree(e->x); cudaFree(e->x2); cudaFree(e->w); cudaFree(e->w2); cudaFree(e->el); cudaFree(e->el2);
  cudaFree(e->blk); cudaFree(e->acc); cudaFree(e->partials); cudaFree(e->ticket); cudaFree(e->red);
  cudaFree(e->cum); cudaFree(e->blocksums); cudaFree(e->src); cudaFree(e->series); cudaFree(e->step_e); cudaFree(e->gath); cudaFree(e->bar); cudaFree(e->vb_sums); cudaFree(e->vb_coarse);
  cudaFree(e->sb_list); cudaFree(e->sb_mask); cudaFree(e->sb_fen); cudaFree(e->sb_draws);
  cudaFree(e->osamp); cudaFree(e->gram); cudaFree(e->gram_partials); cudaFree(e->xchg);
  mole_ctx_s* ctx = e->ctx;
  delete e;
  if (--ctx->live_ens == 0 && ctx->closing) ctx_free(ctx);
  return MOLE_OK;
}

int32_t mole_ensemble_num_walkers(mole_ens_t e, int64_t* n) { if (!e || !n) return MOLE_ERR_INVALID_ARG; *n = e->W; return MOLE_OK; }

static int32_t ens_init(mole_ens_t e, int normal, double a, double b, int broadcast) {
  if (!e) return MOLE_ERR_INVALID_ARG;
  mole_ctx_s* ctx = e->ctx;
  CU(ctx, cudaSetDevice(ctx->device));
  init_kernel<<<cdiv(e->W, 256), 256, 0, STREAM(ctx)>>>(e->x, e->W, e->ne, e->walker_offset, e->key, normal, a, b, broadcast);
  KERNEL_CHECK(ctx);
  e->el_cached = 0;
  return MOLE_OK;
}
int32_t mole_ensemble_init_uniform(mole_ens_t e, double lo, double hi, int32_t bc) { return ens_init(e, 0, lo, hi, bc); }
int32_t mole_ensemble_init_normal(mole_ens_t e, double sigma, int32_t bc) { return ens_init(e, 1, sigma, 0.0, bc); }

static int32_t ens_set(mole_ens_t e, const double* host, int broadcast) {
  if (!e || !host) return MOLE_ERR_INVALID_ARG;
  mole_ctx_s* ctx = e->ctx;
  CU(ctx, cudaSetDevice(ctx->device));
  const int n = 3 * e->ne;
  const size_t cnt = broadcast ? (size_t)n : (size_t)n * e->W;
  // stage through x2 (AoS), transpose into x (SoA)
  CU(ctx, cudaMemcpyAsync(e->x2, host, cnt * sizeof(double), cudaMemcpyHostToDevice, STREAM(ctx)));
  aos_to_soa_kernel<<<cdiv((int64_t)n * e->W, 256), 256, 0, STREAM(ctx)>>>(e->x2, e->x, e->W, n, broadcast);
  KERNEL_CHECK(ctx);
  CU(ctx, cudaStreamSynchronize(STREAM(ctx)));
  e->el_cached = 0;
  return MOLE_OK;
}
int32_t mole_ensemble_set_configs(mole_ens_t e, const double* cfgs) { return ens_set(e, cfgs, 0); }
int32_t mole_ensemble_set_configs_broadcast(mole_ens_t e, const double* cfg) { return ens_set(e, cfg, 1); }

int32_t mole_ensemble_get_configs(mole_ens_t e, double* cfgs) {
  if (!e || !cfgs) return MOLE_ERR_INVALID_ARG;
  mole_ctx_s* ctx = e->ctx;
  CU(ctx, cudaSetDevice(ctx->device));
  const int n = 3 * e->ne;
  soa_to_aos_kernel<<<cdiv((int64_t)n * e->W, 256), 256, 0, STREAM(ctx)>>>(e->x, e->x2, e->W, n);
  KERNEL_CHECK(ctx);
  CU(ctx, cudaMemcpyAsync(cfgs, e->x2, (size_t)n * e->W * sizeof(double), cudaMemcpyDeviceToHost, STREAM(ctx)));
  CU(ctx, cudaStreamSynchronize(STREAM(ctx)));
  return MOLE_OK;
}
int32_t mole_ensemble_set_weights(mole_ens_t e, const double* w) {
  if (!e || !w) return MOLE_ERR_INVALID_ARG;
  CU(e->ctx, cudaSetDevice(e->ctx->device));
  CU(e->ctx, cudaMemcpyAsync(e->w, w, e->W * sizeof(double), cudaMemcpyHostToDevice, STREAM(e->ctx)));
  CU(e->ctx, cudaStreamSynchronize(STREAM(e->ctx)));
  e->wstats_valid = 0;
  e->w_uniform = 0;
  return MOLE_OK;
}
int32_t mole_ensemble_get_weights(mole_ens_t e, double* w) {
  if (!e || !w) return MOLE_ERR_INVALID_ARG;
  CU(e->ctx, cudaSetDevice(e->ctx->device));
  CU(e->ctx, cudaMemcpyAsync(w, e->w, e->W * sizeof(double), cudaMemcpyDeviceToHost, STREAM(e->ctx)));
  CU(e->ctx, cudaStreamSynchronize(STREAM(e->ctx)));
  return MOLE_OK;
}
int32_t mole_ensemble_snapshot(mole_ens_t e) {
  if (!e) return MOLE_ERR_INVALID_ARG;
  CU(e->ctx, cudaSetDevice(e->ctx->device));
  const size_t bytes = (size_t)3 * e->ne * e->W * sizeof(double);
  if (!e->x0) CU(e->ctx, cudaMalloc(&e->x0, bytes));
  CU(e->ctx, cudaMemcpyAsync(e->x0, e->x, bytes, cudaMemcpyDeviceToDevice, STREAM(e->ctx)));
  return MOLE_OK;
}
int32_t mole_ensemble_restore(mole_ens_t e) {
  if (!e) return MOLE_ERR_INVALID_ARG;
  if (!e->x0) return mole_set_error(e->ctx, MOLE_ERR_EMPTY_CACHE, "mole_ensemble_restore without a snapshot");
  CU(e->ctx, cudaSetDevice(e->ctx->device));
  const size_t bytes = (size_t)3 * e->ne * e->W * sizeof(double);
  CU(e->ctx, cudaMemcpyAsync(e->x, e->x0, bytes, cudaMemcpyDeviceToDevice, STREAM(e->ctx)));
  e->el_cached = 0;
  return MOLE_OK;
}
int32_t mole_ensemble_reseed(mole_ens_t e, const uint8_t seed[32]) {
  if (!e || !seed) return MOLE_ERR_INVALID_ARG;
  e->key = mole_key_from_seed(seed);
  e->step = 0;
  return MOLE_OK;
}
int32_t mole_ensemble_set_step(mole_ens_t e, uint32_t s) { if (!e) return MOLE_ERR_INVALID_ARG; e->step = s; return MOLE_OK; }
int32_t mole_ensemble_get_step(mole_ens_t e, uint32_t* s) { if (!e || !s) return MOLE_ERR_INVALID_ARG; *s = e->step; return MOLE_OK; }

// ------------------------------------------------------------------ batched evaluation
static int32_t eval_device(mole_ctx_s* ctx, const WfParams& wp, const HamParams* hp, const double* x_dev, int64_t W,
                           double* psi, double* grad, double* lap, double* hpsi, double* pgrad) {
  // outputs are HOST pointers; stage through temporaries
  const int n = 3 * wp.ne, np = wp.np;
  double *d_psi = nullptr, *d_grad = nullptr, *d_lap = nullptr, *d_h = nullptr, *d_pg = nullptr;
  struct Free {                                   // the CU() error returns below must not leak the temporaries
    double **a, **b, **c, **d, **e;
    ~Free() { cudaFree(*a); cudaFree(*b); cudaFree(*c); cudaFree(*d); cudaFree(*e); }
  } guard{&d_psi, &d_grad, &d_lap, &d_h, &d_pg};
  if (psi) CU(ctx, cudaMalloc(&d_psi, W * sizeof(double)));
  if (grad) CU(ctx, cudaMalloc(&d_grad, (size_t)W * n * sizeof(double)));
  if (lap) CU(ctx, cudaMalloc(&d_lap, W * sizeof(double)));
  if (hpsi && hp) CU(ctx, cudaMalloc(&d_h, W * sizeof(double)));
  if (pgrad && np > 0) CU(ctx, cudaMalloc(&d_pg, (size_t)W * np * sizeof(double)));
  HamParams h0;
  memset(&h0, 0, sizeof(h0));
  const HamParams& h = hp ? *hp : h0;
  const int blocks = cdiv(W, 128);
  switch (wp.kind) {
#define EV(K) case K: eval_kernel<K><<<blocks, 128, 0, STREAM(ctx)>>>(x_dev, W, wp, h, hp != nullptr, d_psi, d_grad, d_lap, d_h, d_pg); break;
    EV(K_STO_1S) EV(K_GAUSSIAN) EV(K_STO_PRODUCT) EV(K_H2_HL_STO) EV(K_H2P_PRODUCT) EV(K_CONSTANT) EV(K_LCAO_1E_2C) EV(K_LCAO_2E_1C) EV(K_LCAO_2E_2C)
#undef EV
    case K_SLATER_JASTROW:
      sj_eval_launch(STREAM(ctx), x_dev, W, wp, h, hp != nullptr, d_psi, d_grad, d_lap, d_h, d_pg);
      break;
    case K_LCAO_SJ:
      lsj_eval_kernel<<<cdiv(W, LSJ_THREADS), LSJ_THREADS, 0, STREAM(ctx)>>>(x_dev, W, wp, h, hp != nullptr, d_psi, d_grad, d_lap, d_h, d_pg);
      break;
    default: return mole_set_error(ctx, MOLE_ERR_INVALID_ARG, "unknown wavefunction kind");
  }
  KERNEL_CHECK(ctx);
  if (psi) CU(ctx, cudaMemcpyAsync(psi, d_psi, W * sizeof(double), cudaMemcpyDeviceToHost, STREAM(ctx)));
  if (grad) CU(ctx, cudaMemcpyAsync(grad, d_grad, (size_t)W * n * sizeof(double), cudaMemcpyDeviceToHost, STREAM(ctx)));
  if (lap) CU(ctx, cudaMemcpyAsync(lap, d_lap, W * sizeof(double), cudaMemcpyDeviceToHost, STREAM(ctx)));
  if (d_h) CU(ctx, cudaMemcpyAsync(hpsi, d_h, W * sizeof(double), cudaMemcpyDeviceToHost, STREAM(ctx)));
  if (d_pg) CU(ctx, cudaMemcpyAsync(pgrad, d_pg, (size_t)W * np * sizeof(double), cudaMemcpyDeviceToHost, STREAM(ctx)));
  CU(ctx, cudaStreamSynchronize(STREAM(ctx)));
  return MOLE_OK;
}

int32_t mole_eval_vgl(mole_ens_t e, mole_wf_t wf, mole_op_t op, double* psi, double* grad, double* lap, double* hpsi,
                      double* pgrad) {
  if (!e || !wf) return MOLE_ERR_INVALID_ARG;
  mole_ctx_s* ctx = e->ctx;
  if (wf->p.ne != e->ne) return mole_set_error(ctx, MOLE_ERR_SHAPE, "wavefunction / ensemble electron count mismatch");
  if (grad && wf->p.kind == MOLE_WF_CONSTANT) return mole_set_error(ctx, MOLE_ERR_FUNC, "WaveFunctionMock::gradient is unimplemented");
  CU(ctx, cudaSetDevice(ctx->device));
  return eval_device(ctx, wf->p, op ? &op->p : nullptr, e->x, e->W, psi, grad, lap, hpsi, pgrad);
}

// single-configuration trait calls (Function::value etc.), evaluated on the device
static int32_t eval_one(mole_wf_t wf, mole_op_t op, const double* cfg, double* psi, double* grad, double* lap, double* hpsi,
                        double* pgrad) {
  if (!wf || !cfg) return MOLE_ERR_INVALID_ARG;
  mole_ctx_s* ctx = wf->ctx;
  CU(ctx, cudaSetDevice(ctx->device));
  const int n = 3 * wf->p.ne;
  double* d_x = nullptr;
  CU(ctx, cudaMalloc(&d_x, n * sizeof(double)));
  CU(ctx, cudaMemcpyAsync(d_x, cfg, n * sizeof(double), cudaMemcpyHostToDevice, STREAM(ctx)));  // W=1: AoS == SoA
  const int32_t rc = eval_device(ctx, wf->p, op ? &op->p : nullptr, d_x, 1, psi, grad, lap, hpsi, pgrad);
  cudaFree(d_x);
  return rc;
}
int32_t mole_wf_value(mole_wf_t wf, const double* cfg, double* out) { return eval_one(wf, nullptr, cfg, out, nullptr, nullptr, nullptr, nullptr); }
int32_t mole_wf_gradient(mole_wf_t wf, const double* cfg, double* out) {
  if (wf && wf->p.kind == MOLE_WF_CONSTANT) return mole_set_error(wf->ctx, MOLE_ERR_FUNC, "WaveFunctionMock::gradient is unimplemented");
  return eval_one(wf, nullptr, cfg, nullptr, out, nullptr, nullptr, nullptr);
}
int32_t mole_wf_laplacian(mole_wf_t wf, const double* cfg, double* out) { return eval_one(wf, nullptr, cfg, nullptr, nullptr, out, nullptr, nullptr); }
int32_t mole_wf_parameter_gradient(mole_wf_t wf, const double* cfg, double* out) {
  if (wf && wf->p.np == 0) return mole_set_error(wf->ctx, MOLE_ERR_FUNC, "this wavefunction kind has no Optimize impl");
  return eval_one(wf, nullptr, cfg, nullptr, nullptr, nullptr, nullptr, out);
}
int32_t mole_op_act_on(mole_op_t op, mole_wf_t wf, const double* cfg, double* out) {
  if (!op) return MOLE_ERR_INVALID_ARG;
  return eval_one(wf, op, cfg, nullptr, nullptr, nullptr, out, nullptr);
}

// ------------------------------------------------------------------ Metropolis objects
int32_t mole_metropolis_create(int32_t kind, double param, mole_metrop_t* out) {
  if (!out || (kind != MOLE_METROP_BOX && kind != MOLE_METROP_DIFFUSE) || !(param > 0.0))
    return mole_set_error(nullptr, MOLE_ERR_INVALID_ARG, "mole_metropolis_create: bad kind or parameter");
  *out = new mole_metrop_s{kind, param, 0u};
  return MOLE_OK;
}
int32_t mole_metropolis_set_compat(mole_metrop_t m, uint32_t compat) {
  if (!m) return MOLE_ERR_INVALID_ARG;
  m->compat = compat;
  return MOLE_OK;
}
int32_t mole_metropolis_destroy(mole_metrop_t m) { delete m; return MOLE_OK; }

}  // extern "C"

// ------------------------------------------------------------------ fused sweep
template <int KIND>
static void launch_sweep_kind(cudaStream_t st, int blocks, const SweepParams& sp, int metrop, bool opt) {
  if (metrop == MOLE_METROP_BOX) {
    if (opt) sweep_kernel<KIND, MOLE_METROP_BOX, true><<<blocks, SWEEP_THREADS, 0, st>>>(sp);
    else sweep_kernel<KIND, MOLE_METROP_BOX, false><<<blocks, SWEEP_THREADS, 0, st>>>(sp);
  } else {
    if (opt) sweep_kernel<KIND, MOLE_METROP_DIFFUSE, true><<<blocks, SWEEP_THREADS, 0, st>>>(sp);
    else sweep_kernel<KIND, MOLE_METROP_DIFFUSE, false><<<blocks, SWEEP_THREADS, 0, st>>>(sp);
  }
}

extern "C" {

int32_t mole_sweep(mole_ens_t e, mole_wf_t wf, mole_metrop_t m, mole_op_t op, const mole_sweep_args* a) {
  if (!e || !wf || !m || !a) return mole_set_error(e ? e->ctx : nullptr, MOLE_ERR_INVALID_ARG, "mole_sweep: NULL argument");
  mole_ctx_s* ctx = e->ctx;
  if (wf->p.ne != e->ne) return mole_set_error(ctx, MOLE_ERR_SHAPE, "wavefunction / ensemble electron count mismatch");
  if (a->n_sweeps < 0 || a->n_discard < 0 || a->n_discard > a->n_sweeps || a->block_size < 1)
    return mole_set_error(ctx, MOLE_ERR_INVALID_ARG, "mole_sweep: bad sweep schedule");
  if ((a->observables & (MOLE_OBS_ENERGY | MOLE_OBS_KINETIC)) && !op)
    return mole_set_error(ctx, MOLE_ERR_DATA_ACCESS, "observable needs an operator");
  const bool opt = (a->observables & MOLE_OBS_PGRAD) != 0;
  if (opt && wf->p.np == 0) return mole_set_error(ctx, MOLE_ERR_FUNC, "\"Parameter gradient\" needs a wavefunction with an Optimize impl");
  if (opt && !(a->observables & MOLE_OBS_ENERGY))
    return mole_set_error(ctx, MOLE_ERR_DATA_ACCESS, "\"Parameter gradient\" moments need the \"Energy\" observable (util.rs:24-27)");
  if (wf->p.kind == MOLE_WF_CONSTANT && m->kind == MOLE_METROP_DIFFUSE)
    return mole_set_error(ctx, MOLE_ERR_FUNC, "WaveFunctionMock::gradient is unimplemented");
  if (a->n_sweeps == 0) return MOLE_OK;
  const bool big_p = opt && wf->p.np > MOLE_ACC_MAX_PARAMS;     // moments = Gram matrix of the per-sample rows
  if (big_p && wf->p.kind != K_LCAO_SJ) return mole_set_error(ctx, MOLE_ERR_SHAPE, "more than MOLE_ACC_MAX_PARAMS parameters");
  if (big_p) {
    // the rows of one launch must fit the sample buffer (<= 1 GiB): longer schedules run as several launches, which
    // the counter-based streams make bit-identical to one
    const int64_t ns_all = a->n_sweeps - a->n_discard;
    const size_t row_bytes = (size_t)(wf->p.np + 2) * e->W * sizeof(double);
    const int64_t cap = std::max<int64_t>(1, (int64_t)(((size_t)1 << 30) / row_bytes));
    if (ns_all > cap) {
      if (a->energy_trace || a->wfvalue_trace || a->kinetic_trace || a->pgrad_trace || a->accept_trace)
        return mole_set_error(ctx, MOLE_ERR_INVALID_ARG, "traces of a large-P sweep need a schedule that fits one launch");
      mole_sweep_args part = *a;
      int64_t left = a->n_sweeps;
      bool first = true;
      while (left > 0) {
        const int64_t disc = first ? a->n_discard : 0;
        const int64_t take = std::min<int64_t>(left, disc + cap);
        part.n_sweeps = (int32_t)take; part.n_discard = (int32_t)std::min<int64_t>(disc, take);
        if (!first && (a->flags & MOLE_SWEEP_KEEP_SERIES)) part.flags = (a->flags & ~MOLE_SWEEP_KEEP_SERIES) | MOLE_SWEEP_APPEND_SERIES;
        const int32_t rc = mole_sweep(e, wf, m, op, &part);
        if (rc != MOLE_OK) return rc;
        left -= take;
        first = false;
      }
      return MOLE_OK;
    }
  }
  MOLE_RANGE("mole_sweep");
  CU(ctx, cudaSetDevice(ctx->device));

  const int64_t W = e->W;
  const int ne = e->ne, np = wf->p.np;
  const int64_t nsamp = a->n_sweeps - a->n_discard;
  if (e->blk_size != a->block_size) {   // a new blocking analysis starts
    CU(ctx, cudaMemsetAsync(e->blk, 0, W * sizeof(double), STREAM(ctx)));
    e->blk_size = a->block_size;
    e->blk_fill = 0;
  }

  SweepParams sp;
  memset(&sp, 0, sizeof(sp));
  sp.x = e->x; sp.blk = e->blk; sp.acc = e->acc; sp.partials = e->partials; sp.ticket = e->ticket;
  sp.W = W; sp.walker_offset = e->walker_offset; sp.key = e->key; sp.step0 = e->step;
  sp.n_sweeps = a->n_sweeps; sp.n_discard = a->n_discard; sp.block_size = a->block_size; sp.blk_fill = e->blk_fill;
  sp.observables = a->observables; sp.compat = a->compat | m->compat; sp.metrop_param = m->param;
  sp.wf = wf->p;
  if (op) sp.ham = op->p;

  // E_L series kept on the device for mole_series_analyze: the kernel writes straight into it
  const bool keep_series = (a->flags & (MOLE_SWEEP_KEEP_SERIES | MOLE_SWEEP_APPEND_SERIES)) && nsamp > 0 &&
                           (a->observables & MOLE_OBS_ENERGY);
  if (keep_series) {
    if (!(a->flags & MOLE_SWEEP_APPEND_SERIES)) e->series_n = 0;
    const int64_t need = e->series_n + nsamp;
    if (need > e->series_cap) {
      const int64_t cap = std::max(need, 2 * e->series_cap);
      double* grown = nullptr;
      CU(ctx, cudaMalloc(&grown, (size_t)cap * W * sizeof(double)));
      if (e->series_n > 0)
        CU(ctx, cudaMemcpyAsync(grown, e->series, (size_t)e->series_n * W * sizeof(double), cudaMemcpyDeviceToDevice, STREAM(ctx)));
      CU(ctx, cudaStreamSynchronize(STREAM(ctx)));
      cudaFree(e->series);
      e->series = grown;
      e->series_cap = cap;
    }
    sp.tr_energy = e->series + (size_t)e->series_n * W;
    e->series_n = need;
  }
  // device staging for the optional traces
  if (a->energy_trace && nsamp > 0 && !keep_series) CU(ctx, cudaMalloc(&sp.tr_energy, (size_t)W * nsamp * sizeof(double)));
  if (a->wfvalue_trace && nsamp > 0) CU(ctx, cudaMalloc(&sp.tr_wfvalue, (size_t)W * nsamp * sizeof(double)));
  if (a->kinetic_trace && nsamp > 0) CU(ctx, cudaMalloc(&sp.tr_kinetic, (size_t)W * nsamp * sizeof(double)));
  if (a->pgrad_trace && nsamp > 0 && opt) CU(ctx, cudaMalloc(&sp.tr_pgrad, (size_t)W * nsamp * np * sizeof(double)));
  if (a->accept_trace) CU(ctx, cudaMalloc(&sp.tr_accept, (size_t)W * a->n_sweeps * ne));

  if (wf->p.kind == K_SLATER_JASTROW) {
    const int32_t rc = sj_sweep_launch(ctx, e, sp, m->kind, opt);
    if (rc != MOLE_OK) return rc;
  } else if (wf->p.kind == K_LCAO_SJ) {
    const int cols = np + 2;
    if (opt && nsamp > 0) {
      const size_t need = (size_t)nsamp * cols * W;
      if (need > e->osamp_cap) {
        CU(ctx, cudaStreamSynchronize(STREAM(ctx)));
        cudaFree(e->osamp);
        e->osamp = nullptr; e->osamp_cap = 0;
        CU(ctx, cudaMalloc(&e->osamp, need * sizeof(double)));
        e->osamp_cap = need;
      }
      if (!e->gram) {
        CU(ctx, cudaMalloc(&e->gram, GRAM_PAD * GRAM_PAD * sizeof(double)));
        CU(ctx, cudaMemsetAsync(e->gram, 0, GRAM_PAD * GRAM_PAD * sizeof(double), STREAM(ctx)));
        e->gram_rows = ctx->sm_count * 2;
        CU(ctx, cudaMalloc(&e->gram_partials, (size_t)e->gram_rows * GRAM_PAD * GRAM_PAD * sizeof(double)));
      }
      sp.osamp = e->osamp;
    }
    // all walkers resident (up to 16 CTAs of 64 threads per SM): the kernel streams its local-memory frames through
    // DRAM at 2^16 walkers (ncu: 30 GB per 20 sweeps, frames 275 MB > L2), yet fewer resident CTAs so that the frames
    // fit L2 measured slower at every setting (H8, 2^16 walkers, 50 sweeps: 16 CTAs/SM 35.4 ms, 8 34.7, 4 43.9, 2 54.1)
    const int blocks = std::min(cdiv(W, LSJ_THREADS), e->partial_rows);
    if (m->kind == MOLE_METROP_BOX) {
      if (opt) lsj_sweep_kernel<MOLE_METROP_BOX, true><<<blocks, LSJ_THREADS, 0, STREAM(ctx)>>>(sp);
      else lsj_sweep_kernel<MOLE_METROP_BOX, false><<<blocks, LSJ_THREADS, 0, STREAM(ctx)>>>(sp);
    } else {
      if (opt) lsj_sweep_kernel<MOLE_METROP_DIFFUSE, true><<<blocks, LSJ_THREADS, 0, STREAM(ctx)>>>(sp);
      else lsj_sweep_kernel<MOLE_METROP_DIFFUSE, false><<<blocks, LSJ_THREADS, 0, STREAM(ctx)>>>(sp);
    }
    if (sp.osamp) {                                          // S = O^T O (and every other moment): the Gram matrix of the rows
      KERNEL_CHECK(ctx);
      MOLE_RANGE("mole_gram");
      if (e->gram_impl == 0) gram_dmma_launch(e->gram_rows, STREAM(ctx), e->osamp, W, nsamp, cols, e->gram_partials);
      else gram_fma_kernel<<<e->gram_rows, GRAM_THREADS, 0, STREAM(ctx)>>>(e->osamp, W, nsamp, cols, e->gram_partials);
      KERNEL_CHECK(ctx);
      gram_fold_kernel<<<cdiv(GRAM_PAD * GRAM_PAD, 256), 256, 0, STREAM(ctx)>>>(e->gram_partials, e->gram_rows, e->gram);
      e->gram_cols = cols;
    }
  } else {
    const int blocks = std::min(cdiv(W, SWEEP_THREADS), e->partial_rows);
    switch (wf->p.kind) {
#define SW(K) case K: launch_sweep_kind<K>(STREAM(ctx), blocks, sp, m->kind, opt); break;
      SW(K_STO_1S) SW(K_GAUSSIAN) SW(K_STO_PRODUCT) SW(K_H2_HL_STO) SW(K_H2P_PRODUCT) SW(K_CONSTANT) SW(K_LCAO_1E_2C) SW(K_LCAO_2E_1C) SW(K_LCAO_2E_2C)
#undef SW
      default: return mole_set_error(ctx, MOLE_ERR_INVALID_ARG, "unknown wavefunction kind");
    }
  }
  KERNEL_CHECK(ctx);
  e->step += (uint32_t)a->n_sweeps;
  if (a->observables & MOLE_OBS_ENERGY) e->blk_fill = (int32_t)((e->blk_fill + nsamp) % a->block_size);
  if (opt && !big_p) e->np_last = np;
  e->el_cached = 0;

  const bool any_trace = (sp.tr_energy && a->energy_trace) || sp.tr_wfvalue || sp.tr_kinetic || sp.tr_pgrad || sp.tr_accept;
  if (any_trace) {
    if (sp.tr_energy && a->energy_trace) CU(ctx, cudaMemcpyAsync(a->energy_trace, sp.tr_energy, (size_t)W * nsamp * sizeof(double), cudaMemcpyDeviceToHost, STREAM(ctx)));
    if (sp.tr_wfvalue) CU(ctx, cudaMemcpyAsync(a->wfvalue_trace, sp.tr_wfvalue, (size_t)W * nsamp * sizeof(double), cudaMemcpyDeviceToHost, STREAM(ctx)));
    if (sp.tr_kinetic) CU(ctx, cudaMemcpyAsync(a->kinetic_trace, sp.tr_kinetic, (size_t)W * nsamp * sizeof(double), cudaMemcpyDeviceToHost, STREAM(ctx)));
    if (sp.tr_pgrad) CU(ctx, cudaMemcpyAsync(a->pgrad_trace, sp.tr_pgrad, (size_t)W * nsamp * np * sizeof(double), cudaMemcpyDeviceToHost, STREAM(ctx)));
    if (sp.tr_accept) CU(ctx, cudaMemcpyAsync(a->accept_trace, sp.tr_accept, (size_t)W * a->n_sweeps * ne, cudaMemcpyDeviceToHost, STREAM(ctx)));
    CU(ctx, cudaStreamSynchronize(STREAM(ctx)));
    if (!keep_series) cudaFree(sp.tr_energy);
    cudaFree(sp.tr_wfvalue); cudaFree(sp.tr_kinetic); cudaFree(sp.tr_pgrad); cudaFree(sp.tr_accept);
  }
  return MOLE_OK;
}

// ------------------------------------------------------------------ accumulators
int32_t mole_acc_reset(mole_ens_t e) {
  if (!e) return MOLE_ERR_INVALID_ARG;
  CU(e->ctx, cudaSetDevice(e->ctx->device));
  CU(e->ctx, cudaMemsetAsync(e->acc, 0, ACC_DEV_LEN * sizeof(double), STREAM(e->ctx)));
  CU(e->ctx, cudaMemsetAsync(e->blk, 0, e->W * sizeof(double), STREAM(e->ctx)));
  if (e->gram) CU(e->ctx, cudaMemsetAsync(e->gram, 0, GRAM_PAD * GRAM_PAD * sizeof(double), STREAM(e->ctx)));
  e->blk_fill = 0;
  return MOLE_OK;
}

int32_t mole_acc_get(mole_ens_t e, mole_acc_host* out) {
  if (!e || !out) return MOLE_ERR_INVALID_ARG;
  MOLE_RANGE("mole_acc_get");
  static_assert(sizeof(mole_acc_host) == ACC_LEN * sizeof(double) + 2 * sizeof(int32_t), "acc layout");
  CU(e->ctx, cudaSetDevice(e->ctx->device));
  CU(e->ctx, cudaMemcpyAsync(out, e->acc, ACC_LEN * sizeof(double), cudaMemcpyDeviceToHost, STREAM(e->ctx)));
  CU(e->ctx, cudaStreamSynchronize(STREAM(e->ctx)));
  out->n_params = e->np_last;
  out->reserved = 0;
  return MOLE_OK;
}

int32_t mole_ensemble_health(mole_ens_t e, mole_ens_health* out) {
  if (!e || !out) return MOLE_ERR_INVALID_ARG;
  double h[2];
  CU(e->ctx, cudaSetDevice(e->ctx->device));
  CU(e->ctx, cudaMemcpyAsync(h, e->acc + ACC_BAD, 2 * sizeof(double), cudaMemcpyDeviceToHost, STREAM(e->ctx)));
  CU(e->ctx, cudaStreamSynchronize(STREAM(e->ctx)));
  memset(out, 0, sizeof(*out));
  out->nonfinite_samples = (int64_t)h[0];
  out->nonfinite_dmc_walkers = (int64_t)h[1];
  return MOLE_OK;
}

int32_t mole_acc_device_ptr(mole_ens_t e, void** p, int32_t* n) {
  if (!e || !p || !n) return MOLE_ERR_INVALID_ARG;
  *p = e->acc;
  *n = ACC_DEV_LEN;   // the packed moments followed by the two health counters (all summed over ranks)
  return MOLE_OK;
}

// ------------------------------------------------------------------ Gram matrix of the per-sample rows (large-P kinds)
int32_t mole_gram_get(mole_ens_t e, int32_t* n_cols, double* out) {
  if (!e || !n_cols) return MOLE_ERR_INVALID_ARG;
  *n_cols = e->gram_cols;
  if (!out) return MOLE_OK;
  if (!e->gram || e->gram_cols <= 0) return mole_set_error(e->ctx, MOLE_ERR_DATA_ACCESS, "no large-P \"Parameter gradient\" samples accumulated");
  std::vector<double> h(GRAM_PAD * GRAM_PAD);
  CU(e->ctx, cudaSetDevice(e->ctx->device));
  CU(e->ctx, cudaMemcpyAsync(h.data(), e->gram, h.size() * sizeof(double), cudaMemcpyDeviceToHost, STREAM(e->ctx)));
  CU(e->ctx, cudaStreamSynchronize(STREAM(e->ctx)));
  const int n = e->gram_cols;
  for (int a = 0; a < n; ++a)
    for (int b = 0; b < n; ++b) out[a * n + b] = h[std::min(a, b) * GRAM_PAD + std::max(a, b)];   // symmetric from the upper triangle
  return MOLE_OK;
}

int32_t mole_gram_device_ptr(mole_ens_t e, void** p, int32_t* n_doubles) {
  if (!e || !p || !n_doubles) return MOLE_ERR_INVALID_ARG;
  *p = e->gram;
  *n_doubles = e->gram ? GRAM_PAD * GRAM_PAD : 0;
  return MOLE_OK;
}

// Ratio of the largest to the smallest island weight PER WALKER at the last step of the last SRBrancher block (1 on a
// single rank, or before any block): the same number on every rank, formed from rows that block gathered anyway.
int32_t mole_dmc_island_imbalance(mole_ens_t e, double* ratio) {
  if (!e || !ratio) return MOLE_ERR_INVALID_ARG;
  *ratio = 1.0;
  if (e->island_sw.size() < 2) return MOLE_OK;
  // every island has the rank's own walker count; counts are equal under the bench's sharding, and in general the
  // per-walker mean is what rebalancing equalises - the counts come from the communicator
  if (e->island_cnt.size() != e->island_sw.size()) {            // constant for the life of the ensemble: gathered once
    e->island_cnt.assign(e->island_sw.size(), (double)e->W);
    if (e->ctx->nranks > 1) {
      double mine = (double)e->W;
      int32_t rc = mole_comm_allgather_host(e->ctx, &mine, 1, e->island_cnt.data());
      if (rc != MOLE_OK) { e->island_cnt.clear(); return rc; }
    }
  }
  const std::vector<double>& cnt = e->island_cnt;
  double lo = 1e300, hi = 0.0;
  for (size_t r = 0; r < e->island_sw.size(); ++r) {
    const double m = e->island_sw[r] / cnt[r];
    lo = std::min(lo, m); hi = std::max(hi, m);
  }
  *ratio = (lo > 0.0 && std::isfinite(hi)) ? hi / lo : HUGE_VAL;
  return MOLE_OK;
}

int32_t mole_dmc_block_select(mole_ens_t e, int32_t impl) {
  if (!e || impl < 0 || impl > 2) return MOLE_ERR_INVALID_ARG;
  e->dmc_block_impl = impl;
  return MOLE_OK;
}

int32_t mole_gram_select(mole_ens_t e, int32_t impl) {
  if (!e || (impl != 0 && impl != 1)) return MOLE_ERR_INVALID_ARG;
  e->gram_impl = impl;
  return MOLE_OK;
}

// ------------------------------------------------------------------ DMC
}  // extern "C"

// one time step for all walkers, enqueued on the stream; the reduction leaves
// red[0..3] = {sum w E_old, sum w, sum w', max w'} on the device
static int32_t dmc_params_fill(mole_ens_t e, mole_wf_t wf, mole_metrop_t m, mole_op_t op, double time_step, double e_ref, DmcParams& dp) {
  if (!e || !wf || !m || !op) return mole_set_error(e ? e->ctx : nullptr, MOLE_ERR_INVALID_ARG, "mole_dmc_step: NULL argument");
  mole_ctx_s* ctx = e->ctx;
  if (m->kind != MOLE_METROP_DIFFUSE) return mole_set_error(ctx, MOLE_ERR_INVALID_ARG, "DmcRunner takes a MetropolisDiffuse (dmc.rs:28)");
  if (wf->p.ne != e->ne) return mole_set_error(ctx, MOLE_ERR_SHAPE, "wavefunction / ensemble electron count mismatch");
  if (wf->p.kind == MOLE_WF_CONSTANT) return mole_set_error(ctx, MOLE_ERR_FUNC, "WaveFunctionMock::gradient is unimplemented");
  CU(ctx, cudaSetDevice(ctx->device));
  memset(&dp, 0, sizeof(dp));
  dp.x = e->x; dp.w = e->w; dp.el = e->el; dp.red = e->red; dp.partials = e->partials; dp.ticket = e->ticket; dp.health = e->acc + ACC_BAD_DMC;
  dp.W = e->W; dp.walker_offset = e->walker_offset; dp.key = e->key; dp.step = e->step;
  // E_old is carried from the previous step (it equals that step's E_new) - only for the same trial function and operator
  const uint64_t sig = mole_el_signature(wf->p, op->p);
  if (e->el_sig != sig) e->el_cached = 0;
  e->el_sig = sig;
  dp.tau_move = m->param; dp.tau_weight = time_step; dp.e_ref = e_ref; dp.el_cached = e->el_cached; dp.compat = m->compat;
  dp.wf = wf->p; dp.ham = op->p;
  return MOLE_OK;
}

static int32_t dmc_step_launch(mole_ens_t e, mole_wf_t wf, mole_metrop_t m, mole_op_t op, double time_step, double e_ref) {
  DmcParams dp;
  int32_t rcf = dmc_params_fill(e, wf, m, op, time_step, e_ref, dp);
  if (rcf != MOLE_OK) return rcf;
  mole_ctx_s* ctx = e->ctx;
  MOLE_RANGE("mole_dmc_step");
  if (wf->p.kind == K_SLATER_JASTROW) {
    const int32_t rc = sj_dmc_launch(ctx, e, dp);
    if (rc != MOLE_OK) return rc;
  } else if (wf->p.kind == K_LCAO_SJ) {
    const int blocks = std::min(cdiv(e->W, LSJ_THREADS), e->partial_rows);
    lsj_dmc_kernel<<<blocks, LSJ_THREADS, 0, STREAM(ctx)>>>(dp);
  } else {
    const int blocks = std::min(cdiv(e->W, SWEEP_THREADS), e->partial_rows);
    switch (wf->p.kind) {
#define DM(K) case K: dmc_step_kernel<K><<<blocks, SWEEP_THREADS, 0, STREAM(ctx)>>>(dp); break;
      DM(K_STO_1S) DM(K_GAUSSIAN) DM(K_STO_PRODUCT) DM(K_H2_HL_STO) DM(K_H2P_PRODUCT) DM(K_LCAO_1E_2C) DM(K_LCAO_2E_1C) DM(K_LCAO_2E_2C)
#undef DM
      default: return mole_set_error(ctx, MOLE_ERR_INVALID_ARG, "unknown wavefunction kind");
    }
  }
  KERNEL_CHECK(ctx);
  e->el_cached = 1;
  e->wstats_valid = 1;
  e->w_uniform = 0;
  return MOLE_OK;
}

// SRBrancher::branch enqueued on the stream: fused weight scan (+ tile-total scan by the last CTA) and
// the pick + gather on tile-local prefix sums.  Either the scalars (norm_factor, new_weight) come from
// the host, or - max_ptr / local_sum_ptr non-null - they are formed on the device from the step
// kernel's reduction, with the same arithmetic.
static int32_t sr_branch_launch(mole_ens_t e, double norm_factor, double new_weight, const double* red_rows, int n_rows,
                                double gcount, const double* local_sum_ptr, double* step_e_out) {
  mole_ctx_s* ctx = e->ctx;
  MOLE_RANGE("mole_branch_sr");
  const int64_t W = e->W;
  const int n = 3 * e->ne;
  const int tiles = e->n_scan_blocks;
  cudaStream_t st = STREAM(ctx);
  if (!e->el_cached) CU(ctx, cudaMemsetAsync(e->el, 0, W * sizeof(double), st));
  sr_weights_scan_fused_kernel<<<tiles, SCAN_THREADS, 0, st>>>(e->w, W, norm_factor, red_rows, n_rows, gcount, e->cum,
                                                               e->blocksums, tiles, e->ticket + 1);
  KERNEL_CHECK(ctx);
  sr_pick_gather_tiled_kernel<<<cdiv(W, 128), 128, 0, st>>>(e->cum, e->blocksums, tiles, W, n, e->walker_offset, e->key, e->step,
                                                            e->x, e->x2, e->el, e->el2, e->w2, new_weight, local_sum_ptr,
                                                            red_rows, n_rows, step_e_out, e->src);
  KERNEL_CHECK(ctx);
  std::swap(e->x, e->x2);
  std::swap(e->w, e->w2);
  std::swap(e->el, e->el2);
  e->wstats_valid = 0;
  e->w_uniform = 1;
  e->step += 1;
  return MOLE_OK;
}

// A whole block of SRBrancher time steps as ONE cooperative launch (mole_dmc_block.cuh).  Returns MOLE_OK with
// *done = 0 when the combination is not eligible (cooperative kinds, population too large for a co-resident grid,
// per-step launches selected): the caller then enqueues the per-step kernels, with identical results.
template <int KIND>
static int32_t dmc_block_launch_kind(mole_ctx_s* ctx, const DmcBlockParams& bp, int n_vb, size_t smem, bool one_block_per_cta,
                                     int* grid_out) {
  if (smem > 48 * 1024)
    CU(ctx, cudaFuncSetAttribute(dmc_block_kernel<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 0;
  CU(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, dmc_block_kernel<KIND>, SWEEP_THREADS, smem));
  int grid = std::min(n_vb, occ * ctx->sm_count);
  // Default policy: only populations that fit the co-resident grid with ONE virtual block per CTA.  Beyond it a CTA
  // walks several blocks per phase and their L2 round trips do not overlap; measured (H atom, us per time step, this
  // kernel / per-step launches): 65 536 walkers 22.7 / 28.9, 98 304: 43.9 / 31.1, 131 072: 46.6 / 35.6, 262 144: 81.8 / 51.6
  if (one_block_per_cta && grid < n_vb) grid = 0;
  *grid_out = grid;
  if (grid < 1) return MOLE_OK;
  void* args[] = {(void*)&bp};
  CU(ctx, cudaLaunchCooperativeKernel((const void*)dmc_block_kernel<KIND>, dim3(grid), dim3(SWEEP_THREADS), args, smem, STREAM(ctx)));
  return MOLE_OK;
}

static int32_t dmc_block_fused(mole_ens_t e, mole_wf_t wf, mole_metrop_t m, mole_op_t op, double time_step, double e_ref,
                               int n_steps, int* done) {
  *done = 0;
  mole_ctx_s* ctx = e->ctx;
  const int kind = wf->p.kind;
  const int n_vb = cdiv(e->W, SWEEP_THREADS);
  if (e->dmc_block_impl == 1 || kind == K_SLATER_JASTROW || kind == K_LCAO_SJ) return MOLE_OK;
  if (n_vb > e->partial_rows) return MOLE_OK;
  int coop = 0;
  CU(ctx, cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, ctx->device));
  if (!coop) return MOLE_OK;
  DmcBlockParams bp;
  memset(&bp, 0, sizeof(bp));
  int32_t rc = dmc_params_fill(e, wf, m, op, time_step, e_ref, bp.dp);
  if (rc != MOLE_OK) return rc;
  MOLE_RANGE("mole_dmc_block_fused");
  cudaStream_t st = STREAM(ctx);
  if (!e->bar) {
    CU(ctx, cudaMalloc(&e->bar, 2 * sizeof(unsigned int)));
    CU(ctx, cudaMalloc(&e->vb_sums, (size_t)(n_vb + 1) * sizeof(unsigned long long)));
    CU(ctx, cudaMalloc(&e->vb_coarse, (size_t)n_vb * DMCB_SUBS * sizeof(unsigned long long)));
  }
  // shared memory: exclusive offsets of the virtual blocks, plus the coarse prefix sums when they fit
  size_t smem = (size_t)(n_vb + 1) * sizeof(unsigned long long);
  const size_t with_coarse = smem + (size_t)n_vb * DMCB_SUBS * sizeof(unsigned long long);
  const int stage_coarse = with_coarse <= (size_t)DMCB_STAGE_BYTES;
  if (stage_coarse) smem = with_coarse;
  // the arrival counter counts 2 * grid per step: chunks keep it far below 2^32
  const int max_chunk = 1 << 14;
  for (int first = 0; first < n_steps; first += max_chunk) {
    const int chunk = std::min(max_chunk, n_steps - first);
    CU(ctx, cudaMemsetAsync(e->bar, 0, 2 * sizeof(unsigned int), st));
    bp.dp.x = e->x; bp.dp.w = e->w; bp.dp.el = e->el; bp.dp.step = e->step; bp.dp.el_cached = e->el_cached;
    bp.x2 = e->x2; bp.w2 = e->w2; bp.el2 = e->el2;
    bp.cum = e->cum; bp.tile_sums = e->vb_sums; bp.coarse = e->vb_coarse; bp.src = e->src;
    bp.step_e = e->step_e + 2 * (size_t)first; bp.bar = e->bar;
    bp.n_steps = chunk; bp.n = 3 * e->ne; bp.stage_coarse = stage_coarse;
    int grid = 0;
    switch (kind) {
#define DB(K) case K: rc = dmc_block_launch_kind<K>(ctx, bp, n_vb, smem, e->dmc_block_impl == 0, &grid); break;
      DB(K_STO_1S) DB(K_GAUSSIAN) DB(K_STO_PRODUCT) DB(K_H2_HL_STO) DB(K_H2P_PRODUCT) DB(K_LCAO_1E_2C) DB(K_LCAO_2E_1C) DB(K_LCAO_2E_2C)
#undef DB
      default: return mole_set_error(ctx, MOLE_ERR_INVALID_ARG, "unknown wavefunction kind");
    }
    if (rc != MOLE_OK) return rc;
    if (grid < 1) {
      if (first == 0) return MOLE_OK;                          // nothing launched yet: per-step path
      return mole_set_error(ctx, MOLE_ERR_CUDA, "dmc_block_kernel: no co-resident grid");
    }
    KERNEL_CHECK(ctx);
    if (chunk & 1) { std::swap(e->x, e->x2); std::swap(e->w, e->w2); std::swap(e->el, e->el2); }
    e->step += (uint32_t)chunk;
    e->el_cached = 1;
    e->wstats_valid = 0;
    e->w_uniform = 1;
    unsigned int flag[2] = {0u, 0u};
    if (first + chunk < n_steps) {                             // (the last chunk's flag is read with the block's rows)
      CU(ctx, cudaMemcpyAsync(flag, e->bar, sizeof(flag), cudaMemcpyDeviceToHost, st));
      CU(ctx, cudaStreamSynchronize(st));
      if (flag[1]) return mole_set_error(ctx, MOLE_ERR_CUDA, "dmc_block_kernel: grid barrier timed out");
    }
  }
  *done = 1;
  return MOLE_OK;
}

extern "C" {

int32_t mole_dmc_step(mole_ens_t e, mole_wf_t wf, mole_metrop_t m, mole_op_t op, double time_step, double e_ref,
                      double* sum_w_e, double* sum_w) {
  const int32_t rc = dmc_step_launch(e, wf, m, op, time_step, e_ref);
  if (rc != MOLE_OK) return rc;
  mole_ctx_s* ctx = e->ctx;
  double red[4];
  CU(ctx, cudaMemcpyAsync(red, e->red, 4 * sizeof(double), cudaMemcpyDeviceToHost, STREAM(ctx)));
  CU(ctx, cudaStreamSynchronize(STREAM(ctx)));
  if (sum_w_e) *sum_w_e = red[0];
  if (sum_w) *sum_w = red[1];
  return MOLE_OK;
}

// One block of DmcRunner::diffuse's inner loop (dmc.rs:84-141), SRBrancher: n_steps x (time step, ensemble
// energy, branch) enqueued without any host read - the reference energy only changes between blocks
// (dmc.rs:143-145) - then ONE copy of the per-step {sum w E, sum w} rows.  Multi-rank: ranks are population islands
// within a block and the rows are all-gathered once per block.  Results are identical to the step-by-step entry
// points (mole_dmc_step + mole_branch), which the other branchers and the parity tests use.
int32_t mole_dmc_block(mole_ens_t e, mole_wf_t wf, mole_metrop_t m, mole_op_t op, int32_t branch_kind, double time_step,
                       double e_ref, int32_t n_steps, double* step_energies) {
  if (!e || !step_energies || n_steps < 0) return MOLE_ERR_INVALID_ARG;
  mole_ctx_s* ctx = e->ctx;
  int32_t rc;
  if (branch_kind != MOLE_BRANCH_SR) {
    for (int j = 0; j < n_steps; ++j) {
      double swe, sw;
      if ((rc = mole_dmc_step(e, wf, m, op, time_step, e_ref, &swe, &sw)) != MOLE_OK) return rc;
      double s[2] = {swe, sw};
      if ((rc = mole_comm_allreduce_host(ctx, s, 2, nullptr, 0)) != MOLE_OK) return rc;
      step_energies[j] = s[0] / s[1];   // dmc.rs:133
      if ((rc = mole_branch(e, branch_kind)) != MOLE_OK) return rc;
    }
    return MOLE_OK;
  }
  if (n_steps == 0) return MOLE_OK;
  MOLE_RANGE("mole_dmc_block");
  CU(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = STREAM(ctx);
  const bool multi = ctx->nranks > 1;
  const int nr = multi ? ctx->nranks : 1;
  if (e->step_e_cap < n_steps) {
    CU(ctx, cudaStreamSynchronize(st));
    cudaFree(e->step_e);
    cudaFree(e->gath);
    e->step_e = nullptr;
    e->gath = nullptr;
    CU(ctx, cudaMalloc(&e->step_e, (size_t)2 * n_steps * sizeof(double)));
    e->step_e_cap = n_steps;
  }
  if (multi && !e->gath) CU(ctx, cudaMalloc(&e->gath, (size_t)nr * 2 * e->step_e_cap * sizeof(double)));
  // Every rank is a population island inside a block: SRBrancher normalises with the rank's own N / w_max and resets
  // to the rank's own mean weight (stratified resampling, unbiased), so NO collective sits in the step loop; the
  // per-step {sum w E, sum w} rows of all ranks are gathered ONCE per block and every rank forms the same energies.
  int fused = 0;
  if ((rc = dmc_block_fused(e, wf, m, op, time_step, e_ref, n_steps, &fused)) != MOLE_OK) return rc;
  for (int j = 0; j < n_steps && !fused; ++j) {
    if ((rc = dmc_step_launch(e, wf, m, op, time_step, e_ref)) != MOLE_OK) return rc;
    if ((rc = sr_branch_launch(e, 0.0, 0.0, e->red, 1, (double)e->W, e->red + 2, e->step_e + 2 * j)) != MOLE_OK) return rc;
  }
  std::vector<double> rows((size_t)nr * 2 * n_steps);
  if (multi) {
    if ((rc = mole_comm_allgather_device(ctx, e->step_e, e->gath, 2 * n_steps)) != MOLE_OK) return rc;
    CU(ctx, cudaMemcpyAsync(rows.data(), e->gath, rows.size() * sizeof(double), cudaMemcpyDeviceToHost, st));
  } else {
    CU(ctx, cudaMemcpyAsync(rows.data(), e->step_e, rows.size() * sizeof(double), cudaMemcpyDeviceToHost, st));
  }
  unsigned int bar_flag[2] = {0u, 0u};
  if (fused) CU(ctx, cudaMemcpyAsync(bar_flag, e->bar, sizeof(bar_flag), cudaMemcpyDeviceToHost, st));
  CU(ctx, cudaStreamSynchronize(st));
  if (bar_flag[1]) return mole_set_error(ctx, MOLE_ERR_CUDA, "dmc_block_kernel: grid barrier timed out");
  e->island_sw.assign(nr, 0.0);                                    // island totals, for mole_dmc_island_imbalance
  for (int r = 0; r < nr; ++r) e->island_sw[r] = rows[((size_t)r * n_steps + n_steps - 1) * 2 + 1];
  for (int j = 0; j < n_steps; ++j) {
    double swe = 0.0, sw = 0.0;                                    // rank order: identical on every rank
    for (int r = 0; r < nr; ++r) { swe += rows[((size_t)r * n_steps + j) * 2]; sw += rows[((size_t)r * n_steps + j) * 2 + 1]; }
    step_energies[j] = swe / sw;                                   // dmc.rs:133
  }
  return MOLE_OK;
}

int32_t mole_branch(mole_ens_t e, int32_t kind) {
  if (!e) return MOLE_ERR_INVALID_ARG;
  mole_ctx_s* ctx = e->ctx;
  if (kind != MOLE_BRANCH_SR && kind != MOLE_BRANCH_SIMPLE) return mole_set_error(ctx, MOLE_ERR_INVALID_ARG, "unknown brancher");
  CU(ctx, cudaSetDevice(ctx->device));
  const int64_t W = e->W;
  const int n = 3 * e->ne;
  const int tiles = e->n_scan_blocks;
  cudaStream_t st = STREAM(ctx);
  if (!e->el_cached) CU(ctx, cudaMemsetAsync(e->el, 0, W * sizeof(double), st));
  if (kind == MOLE_BRANCH_SR) {
    // post-update sum and max of the weights come from the last mole_dmc_step reduction (red[2], red[3]);
    // when the weights were set by hand they are recomputed here.
    double red[4];
    if (e->wstats_valid) {
      CU(ctx, cudaMemcpyAsync(red, e->red, 4 * sizeof(double), cudaMemcpyDeviceToHost, st));
      CU(ctx, cudaStreamSynchronize(st));
    } else {
      std::vector<double> hw(W);
      CU(ctx, cudaMemcpyAsync(hw.data(), e->w, W * sizeof(double), cudaMemcpyDeviceToHost, st));
      CU(ctx, cudaStreamSynchronize(st));
      red[2] = 0.0; red[3] = 0.0;
      for (int64_t i = 0; i < W; ++i) { red[2] = red[2] + hw[i]; red[3] = std::max(red[3], hw[i]); }
    }
    // stratified per rank (see mole_dmc_block): the rank's own N, w_max and mean weight, no collective
    const double norm_factor = (double)W / red[3];            // branching.rs:24
    const double new_weight = red[2] / (double)W;             // branching.rs:21
    return sr_branch_launch(e, norm_factor, new_weight, nullptr, 0, 0.0, nullptr, nullptr);
  } else {
    // scratch of the clone list (<= 3 copies per walker, branching.rs:60) lives with the ensemble: no
    // allocation or synchronisation per time step
    const int64_t cap = 3 * W;
    if (!e->sb_list) {
      CU(ctx, cudaMalloc(&e->sb_list, cap * sizeof(int32_t)));
      CU(ctx, cudaMalloc(&e->sb_mask, (cap / 32 + 2) * sizeof(uint32_t)));
      CU(ctx, cudaMalloc(&e->sb_fen, (cap / 32 + 3) * sizeof(int32_t)));
      CU(ctx, cudaMalloc(&e->sb_draws, 2 * W * sizeof(int64_t)));
    }
    int32_t* list = e->sb_list; int32_t* fen = e->sb_fen; uint32_t* mask = e->sb_mask;
    simple_copies_scan_kernel<<<tiles, SCAN_THREADS, 0, st>>>(e->w, W, e->walker_offset, e->key, e->step, e->cum, e->blocksums);
    KERNEL_CHECK(ctx);
    scan_tile_sums_kernel<<<1, SCAN_THREADS, 0, st>>>(e->blocksums, tiles);
    KERNEL_CHECK(ctx);
    add_tile_offsets_kernel<<<cdiv(W, 256), 256, 0, st>>>(e->cum, W, e->blocksums);
    KERNEL_CHECK(ctx);
    simple_build_list_kernel<<<cdiv(W, 256), 256, 0, st>>>(e->cum, e->blocksums, tiles, W, e->walker_offset, e->key, e->step, list);
    KERNEL_CHECK(ctx);
    // bitmask + Fenwick tree in shared memory when 3W/32 words fit (W <= ~2.6e5), else in the global scratch
    int smem_words = (int)(cap / 32 + 3);
    if ((size_t)smem_words * 8 > 200 * 1024) smem_words = 0;
    simple_remove_kernel<<<1, 1024, (size_t)smem_words * 8, st>>>(list, e->blocksums, tiles, W, e->walker_offset, e->key, e->step,
                                                                 mask, fen, e->sb_draws, smem_words, e->src);
    KERNEL_CHECK(ctx);
    gather_by_list_kernel<<<cdiv(W, 128), 128, 0, st>>>(e->src, W, n, e->x, e->x2, e->el, e->el2, e->w, e->w2, e->src);
    KERNEL_CHECK(ctx);
  }
  std::swap(e->x, e->x2);
  std::swap(e->w, e->w2);
  std::swap(e->el, e->el2);
  e->wstats_valid = 0;
  e->step += 1;
  return MOLE_OK;
}

// Cross-rank population rebalancing (north_star (5)).  Ranks are population islands inside a
// DMC block (mole_dmc_block); their total weights drift apart.  Rebalancing turns the islands back into ONE
// equal-weight population: rank r's walkers fill share_r of the N_total slots (proportional to the rank's total
// weight, systematic rounding with one shared Philox draw), each rank resamples its own walkers systematically to
// share_r copies, surplus copies travel to the ranks with free slots in one grouped ncclSend / ncclRecv exchange
// (3 N_e + 1 doubles per walker: configuration and cached E_L), and every walker ends with weight W_total / N_total.
// Walker counts per rank do not change.  Unbiased: E[copies of a walker] = its weight / mean weight.
int32_t mole_rebalance(mole_ens_t e) {
  if (!e) return MOLE_ERR_INVALID_ARG;
  mole_ctx_s* ctx = e->ctx;
  MOLE_RANGE("mole_rebalance");
  CU(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = STREAM(ctx);
  int32_t rc;
  const int64_t W = e->W;
  if ((rc = mole_comm_warm_p2p(ctx)) != MOLE_OK) return rc;   // first call on a communicator: opens every peer pair once
  if (!e->w_uniform) {   // weights set by hand or by a time step: resample inside the rank only if they really differ
    std::vector<double> hw(W);
    CU(ctx, cudaMemcpyAsync(hw.data(), e->w, W * sizeof(double), cudaMemcpyDeviceToHost, st));
    CU(ctx, cudaStreamSynchronize(st));
    bool same = true;
    for (int64_t i = 1; i < W && same; ++i) same = hw[i] == hw[0];
    if (!same && (rc = mole_branch(e, MOLE_BRANCH_SR)) != MOLE_OK) return rc;
  }
  const int n = 3 * e->ne, nr = ctx->nranks;
  double w0 = 0.0;
  CU(ctx, cudaMemcpyAsync(&w0, e->w, sizeof(double), cudaMemcpyDeviceToHost, st));
  CU(ctx, cudaStreamSynchronize(st));
  const double mine[2] = {w0 * (double)W, (double)W};
  std::vector<double> all(2 * (size_t)nr);
  if ((rc = mole_comm_allgather_host(ctx, mine, 2, all.data())) != MOLE_OK) return rc;
  std::vector<double> totals(nr);
  std::vector<int64_t> counts(nr), shares(nr);
  double T = 0.0;
  int64_t N = 0;
  for (int r = 0; r < nr; ++r) { totals[r] = all[2 * r]; counts[r] = (int64_t)all[2 * r + 1]; T += totals[r]; N += counts[r]; }
  if (!(T > 0.0) || !std::isfinite(T)) return mole_set_error(ctx, MOLE_ERR_DATA_ACCESS, "mole_rebalance: the ensemble has no finite weight left");
  // shared draws: the same on every rank (same key, same step counter); v per rank
  const Philox4 pu = mole_draw(e->key, ~0ull, e->step, DOM_BRANCH, 2, 0);
  const double u = mole_u53(pu.a, pu.b);
  const Philox4 pv = mole_draw(e->key, ~0ull - 1 - (uint64_t)ctx->rank, e->step, DOM_BRANCH, 2, 1);
  const double v = mole_u53(pv.a, pv.b);
  mole_rebalance_shares(nr, totals.data(), counts.data(), u, shares.data());
  const std::vector<MoleMove> moves = mole_rebalance_moves(nr, counts.data(), shares.data());
  const int64_t share = shares[ctx->rank], keep = std::min(share, W);
  const int64_t n_send = std::max<int64_t>(share - W, 0), n_recv = std::max<int64_t>(W - share, 0);
  const size_t need = (size_t)(n_send + n_recv) * (n + 1);
  if (need > e->xchg_cap) {
    cudaFree(e->xchg);
    e->xchg = nullptr; e->xchg_cap = 0;
    CU(ctx, cudaMalloc(&e->xchg, need * sizeof(double)));
    e->xchg_cap = need;
  }
  double* send = e->xchg;
  double* recv = e->xchg + (size_t)n_send * (n + 1);
  if (!e->el_cached) CU(ctx, cudaMemsetAsync(e->el, 0, W * sizeof(double), st));
  if (share > 0) {
    rebalance_gather_kernel<<<cdiv(share, 128), 128, 0, st>>>(e->x, e->x2, e->el, e->el2, W, n, share, v, keep, send);
    KERNEL_CHECK(ctx);
  }
  if ((rc = mole_comm_exchange_rows(ctx, send, recv, n + 1, moves)) != MOLE_OK) return rc;
  if (n_recv > 0) {
    rebalance_scatter_kernel<<<cdiv(n_recv, 128), 128, 0, st>>>(recv, e->x2, e->el2, W, n, share, n_recv);
    KERNEL_CHECK(ctx);
  }
  std::swap(e->x, e->x2);
  std::swap(e->el, e->el2);
  fill_kernel<<<cdiv(W, 256), 256, 0, st>>>(e->w, W, T / (double)N);
  KERNEL_CHECK(ctx);
  CU(ctx, cudaStreamSynchronize(st));
  e->wstats_valid = 0;
  e->w_uniform = 1;
  return MOLE_OK;
}

int32_t mole_branch_sources(mole_ens_t e, int32_t* src) {
  if (!e || !src) return MOLE_ERR_INVALID_ARG;
  CU(e->ctx, cudaSetDevice(e->ctx->device));
  CU(e->ctx, cudaMemcpyAsync(src, e->src, e->W * sizeof(int32_t), cudaMemcpyDeviceToHost, STREAM(e->ctx)));
  CU(e->ctx, cudaStreamSynchronize(STREAM(e->ctx)));
  return MOLE_OK;
}

// ------------------------------------------------------------------ math probe (accuracy tests)
int32_t mole_math_probe(mole_ctx_t ctx, int32_t which, const double* in, int64_t n, double* out) {
  if (!ctx || !in || !out || n < 1 || which < 0 || which > 6) return MOLE_ERR_INVALID_ARG;
  CU(ctx, cudaSetDevice(ctx->device));
  double *d_in = nullptr, *d_out = nullptr;
  CU(ctx, cudaMalloc(&d_in, n * sizeof(double)));
  CU(ctx, cudaMalloc(&d_out, n * sizeof(double)));
  CU(ctx, cudaMemcpyAsync(d_in, in, n * sizeof(double), cudaMemcpyHostToDevice, STREAM(ctx)));
  math_probe_kernel<<<cdiv(n, 256), 256, 0, STREAM(ctx)>>>(which, d_in, n, d_out);
  KERNEL_CHECK(ctx);
  CU(ctx, cudaMemcpyAsync(out, d_out, n * sizeof(double), cudaMemcpyDeviceToHost, STREAM(ctx)));
  CU(ctx, cudaStreamSynchronize(STREAM(ctx)));
  cudaFree(d_in); cudaFree(d_out);
  return MOLE_OK;
}

// ------------------------------------------------------------------ FP64 peak probe
int32_t mole_bench_fp64_peak(mole_ctx_t ctx, double* tflops) {
  if (!ctx || !tflops) return MOLE_ERR_INVALID_ARG;
  CU(ctx, cudaSetDevice(ctx->device));
  // 32 warps/SM x 8 independent chains: the configuration that reached 36.9 TFLOP/s in the DFMA sweep
  // (64 warps/SM reach only 34.8)
  const int blocks = ctx->sm_count * 4, threads = 256, iters = 1 << 16;
  double* out = nullptr;
  CU(ctx, cudaMalloc(&out, blocks * sizeof(double)));
  cudaEvent_t a, b;
  CU(ctx, cudaEventCreate(&a));
  CU(ctx, cudaEventCreate(&b));
  double best = 0.0;
  for (int rep = 0; rep < 5; ++rep) {
    CU(ctx, cudaEventRecord(a, STREAM(ctx)));
    dfma_peak_kernel<<<blocks, threads, 0, STREAM(ctx)>>>(out, iters, 1.0 + rep);
    KERNEL_CHECK(ctx);
    CU(ctx, cudaEventRecord(b, STREAM(ctx)));
    CU(ctx, cudaEventSynchronize(b));
    float ms = 0.f;
    CU(ctx, cudaEventElapsedTime(&ms, a, b));
    const double fl = (double)blocks * threads * (double)iters * 8.0 * 2.0;
    if (rep > 0) best = std::max(best, fl / (ms * 1e-3) / 1e12);
  }
  cudaEventDestroy(a); cudaEventDestroy(b); cudaFree(out);
  *tflops = best;
  return MOLE_OK;
}

// DMMA (m8n8k4 fp64) issue rate: `chains` independent accumulators per warp, `warps_per_sm` resident warps
int32_t mole_bench_dmma_peak(mole_ctx_t ctx, int32_t chains, int32_t warps_per_sm, double* tflops) {
  if (!ctx || !tflops || warps_per_sm < 1 || warps_per_sm > 32) return MOLE_ERR_INVALID_ARG;
  if (chains != 1 && chains != 2 && chains != 4 && chains != 8 && chains != 16) return MOLE_ERR_INVALID_ARG;
  CU(ctx, cudaSetDevice(ctx->device));
  const int blocks = ctx->sm_count, threads = 32 * warps_per_sm, iters = 1 << 14;
  double* out = nullptr;
  CU(ctx, cudaMalloc(&out, blocks * sizeof(double)));
  cudaEvent_t a, b;
  CU(ctx, cudaEventCreate(&a));
  CU(ctx, cudaEventCreate(&b));
  double best = 0.0;
  for (int rep = 0; rep < 4; ++rep) {
    CU(ctx, cudaEventRecord(a, STREAM(ctx)));
    switch (chains) {
      case 1: dmma_peak_kernel<1><<<blocks, threads, 0, STREAM(ctx)>>>(out, iters, 1.0 + rep); break;
      case 2: dmma_peak_kernel<2><<<blocks, threads, 0, STREAM(ctx)>>>(out, iters, 1.0 + rep); break;
      case 4: dmma_peak_kernel<4><<<blocks, threads, 0, STREAM(ctx)>>>(out, iters, 1.0 + rep); break;
      case 8: dmma_peak_kernel<8><<<blocks, threads, 0, STREAM(ctx)>>>(out, iters, 1.0 + rep); break;
      default: dmma_peak_kernel<16><<<blocks, threads, 0, STREAM(ctx)>>>(out, iters, 1.0 + rep); break;
    }
    KERNEL_CHECK(ctx);
    CU(ctx, cudaEventRecord(b, STREAM(ctx)));
    CU(ctx, cudaEventSynchronize(b));
    float ms = 0.f;
    CU(ctx, cudaEventElapsedTime(&ms, a, b));
    const double fl = (double)blocks * warps_per_sm * (double)iters * chains * 512.0;   // 8 x 8 x 4 x 2 flops per MMA
    if (rep > 0) best = std::max(best, fl / (ms * 1e-3) / 1e12);
  }
  cudaEventDestroy(a); cudaEventDestroy(b); cudaFree(out);
  *tflops = best;
  return MOLE_OK;
}

// ------------------------------------------------------------------ Gram contraction timed alone (roofline of the large-P path)
// rows of (1, E, O_k)-like synthetic values, W walkers x n_samples samples x cols columns; returns the mean time of
// `reps` launches (CUDA events on the context's stream) of the chosen implementation (0 DMMA, 1 FP64 vector pipe)
int32_t mole_bench_gram(mole_ctx_t ctx, int64_t W, int64_t n_samples, int32_t cols, int32_t impl, int32_t reps, double* ms_out,
                        double* checksum) {
  if (!ctx || !ms_out || W < 1 || n_samples < 1 || cols < 3 || cols > GRAM_PAD || reps < 1) return MOLE_ERR_INVALID_ARG;
  CU(ctx, cudaSetDevice(ctx->device));
  const size_t n = (size_t)W * n_samples * cols;
  double *data = nullptr, *partials = nullptr, *gram = nullptr;
  const int rows = ctx->sm_count * 2;
  CU(ctx, cudaMalloc(&data, n * sizeof(double)));
  CU(ctx, cudaMalloc(&partials, (size_t)rows * GRAM_PAD * GRAM_PAD * sizeof(double)));
  CU(ctx, cudaMalloc(&gram, GRAM_PAD * GRAM_PAD * sizeof(double)));
  CU(ctx, cudaMemsetAsync(gram, 0, GRAM_PAD * GRAM_PAD * sizeof(double), STREAM(ctx)));
  gram_fill_kernel<<<ctx->sm_count * 8, 256, 0, STREAM(ctx)>>>(data, n);
  KERNEL_CHECK(ctx);
  cudaEvent_t a, b;
  CU(ctx, cudaEventCreate(&a));
  CU(ctx, cudaEventCreate(&b));
  float total = 0.f;
  for (int r = 0; r < reps + 1; ++r) {
    CU(ctx, cudaEventRecord(a, STREAM(ctx)));
    if (impl == 0) gram_dmma_launch(rows, STREAM(ctx), data, W, n_samples, cols, partials);
    else gram_fma_kernel<<<rows, GRAM_THREADS, 0, STREAM(ctx)>>>(data, W, n_samples, cols, partials);
    KERNEL_CHECK(ctx);
    gram_fold_kernel<<<cdiv(GRAM_PAD * GRAM_PAD, 256), 256, 0, STREAM(ctx)>>>(partials, rows, gram);
    KERNEL_CHECK(ctx);
    CU(ctx, cudaEventRecord(b, STREAM(ctx)));
    CU(ctx, cudaEventSynchronize(b));
    float ms = 0.f;
    CU(ctx, cudaEventElapsedTime(&ms, a, b));
    if (r > 0) total += ms;
  }
  if (checksum) {
    std::vector<double> h(GRAM_PAD * GRAM_PAD);
    CU(ctx, cudaMemcpy(h.data(), gram, h.size() * sizeof(double), cudaMemcpyDeviceToHost));
    double s = 0.0;
    for (int i = 0; i < cols; ++i)
      for (int j = i; j < cols; ++j) s += h[i * GRAM_PAD + j];
    *checksum = s / (reps + 1);
  }
  cudaEventDestroy(a); cudaEventDestroy(b); cudaFree(data); cudaFree(partials); cudaFree(gram);
  *ms_out = total / reps;
  return MOLE_OK;
}

#include "mole_api_series.inc"

}  // extern "C"
