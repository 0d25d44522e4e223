// Slater-Jastrow kind (MOLE_WF_SLATER_JASTROW): cooperative sub-warp-per-walker kernels.
// PLACEHOLDER while the thread-per-walker kinds are brought up: every entry point reports an error
// (there is deliberately no CPU fallback).
#pragma once
#include <cuda_runtime.h>
#include "mole_internal.h"

static inline void sj_eval_launch(cudaStream_t, const double*, int64_t, const WfParams&, const HamParams&, int, double*,
                                  double*, double*, double*, double*) {}
static inline int32_t sj_sweep_launch(mole_ctx_s* ctx, mole_ens_s*, const SweepParams&, int, bool) {
  return mole_set_error(ctx, MOLE_ERR_INVALID_ARG, "Slater-Jastrow sweep kernel not built yet");
}
static inline int32_t sj_dmc_launch(mole_ctx_s* ctx, mole_ens_s*, const DmcParams&) {
  return mole_set_error(ctx, MOLE_ERR_INVALID_ARG, "Slater-Jastrow DMC kernel not built yet");
}
