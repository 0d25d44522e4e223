// Kernels of the walker-ensemble hot path for the thread-per-walker kinds (N_e <= 2).
//   init_kernel       Sampler::new / DmcRunner::new configuration draws
//   eval_kernel       batched psi, grad psi, lap psi, H psi, d psi/d p  (parity entry point)
//   sweep_kernel      fused Metropolis sweep + local energy + optimisation moments, K sweeps/launch
//   dmc_step_kernel   one DMC time step (dmc.rs:87-130) with block-tree reductions
// Roofline: FP64 pipe.  Walker state is loaded once per launch, lives in
// registers for all K sweeps, and is stored once; HBM traffic per walker-step is 2*8*(3N_e+1)/K bytes.
#pragma once
#include "mole_internal.h"
#include "mole_rng.cuh"
#include "mole_math.cuh"
#include "mole_wf.cuh"

constexpr int SWEEP_THREADS = 128;

// ------------------------------------------------------------------ per-thread accumulators
// The ten scalar sums always live in registers.  The 2 NP + NP (NP + 1) / 2 optimisation moments live in registers
// too, except when SMEM is set (kinds with NP >= 4: the 18 moments of the two-centre LCAO kind next to its walker
// state spilled, profiles/r01d): then each thread owns one column of a [moments][SWEEP_THREADS] shared-memory block
// (thread index fastest: conflict-free), the way sj_sweep_kernel keeps its 42 moments.
template <int NP, bool SMEM = false>
struct Acc {
  static constexpr int NOO = NP * (NP + 1) / 2;
  static constexpr int NMOM = 2 * NP + NOO;
  static constexpr int LEN = 10 + NMOM;
  double v[SMEM ? 10 : LEN];
  double* col;   // SMEM: this thread's column
  MOLE_D void zero(double* smem_block) {
#pragma unroll
    for (int i = 0; i < (SMEM ? 10 : LEN); ++i) v[i] = 0.0;
    col = smem_block + threadIdx.x;
    if (SMEM)
#pragma unroll
      for (int i = 0; i < NMOM; ++i) col[i * SWEEP_THREADS] = 0.0;
  }
  // moment i (0 .. NMOM-1): sum O_k at k, sum O_k E at NP + k, sum O_k O_l packed at 2 NP + q
  MOLE_D void add(int i, double a, double b) {
    if (SMEM) col[i * SWEEP_THREADS] = fma(a, b, col[i * SWEEP_THREADS]);
    else v[10 + i] = fma(a, b, v[10 + i]);
  }
  MOLE_D void gather(double* all) const {
#pragma unroll
    for (int i = 0; i < 10; ++i) all[i] = v[i];
#pragma unroll
    for (int i = 0; i < NMOM; ++i) all[10 + i] = SMEM ? col[i * SWEEP_THREADS] : v[SMEM ? 0 : 10 + i];
  }
  // compact index -> slot in the packed ACC_* layout
  MOLE_D static int slot(int i) {
    if (i < 10) return i;
    if (i < 10 + NP) return ACC_O + (i - 10);
    if (i < 10 + 2 * NP) return ACC_OE + (i - 10 - NP);
    return ACC_OO + (i - 10 - 2 * NP);
  }
};

// Block-tree reduction: warp shuffles -> shared memory -> one row of `partials` per CTA ->
// the last CTA to arrive (ticket) folds all rows in a fixed order into the global accumulators.
// Deterministic for a given launch geometry; no floating-point atomics.
template <int LEN, class SlotFn>
MOLE_D void mole_block_reduce_to_global(double* v, double* partials, double* acc, unsigned int* ticket, SlotFn slot) {
  __shared__ double sm[32][LEN];
  __shared__ bool is_last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int i = 0; i < LEN; ++i) {
    double x = v[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if (lane == 0) sm[warp][i] = x;
  }
  __syncthreads();
  if (threadIdx.x < LEN) {
    double s = 0.0;
    for (int w = 0; w < nwarp; ++w) s += sm[w][threadIdx.x];
    partials[(size_t)blockIdx.x * ACC_LEN + slot(threadIdx.x)] = s;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = (atomicInc(ticket, gridDim.x - 1) == gridDim.x - 1);
  __syncthreads();
  if (is_last) {
    __threadfence();
    if (threadIdx.x < LEN) {
      const int sl = slot(threadIdx.x);
      double s = 0.0;
      for (unsigned b = 0; b < gridDim.x; ++b) s += partials[(size_t)b * ACC_LEN + sl];
      acc[sl] += s;
    }
  }
}

// ------------------------------------------------------------------ walker = state + cached psi, grad psi
template <class WF>
struct Walker {
  typename WF::State st;
  double psi;
  double g[3 * WF::NE];
  MOLE_D void refresh(const WfParams& p) {
    psi = WF::psi(p, st);
    WF::grad(p, st, g);
  }
};

// `acceptance.min(1.0)` (metrop.rs:80,195).  Rust's f64::min drops a NaN operand, so upstream a NaN
// ratio becomes 1.0 and is accepted; that is reproduced only under MOLE_COMPAT_NAN_ACCEPT, by default a
// NaN ratio is rejected (include/mole_b200.h).
MOLE_D double mole_clamp_acceptance(double a, uint32_t compat) {
  if (isnan(a)) return (compat & MOLE_COMPAT_NAN_ACCEPT) ? 1.0 : 0.0;
  return fmin(a, 1.0);
}

// Metropolis::move_state for electron e.  BOX: metrop.rs:60-96; DIFFUSE: metrop.rs:150-212.
// psi(x), grad psi(x) are cached in the walker instead of being re-evaluated (the reference
// evaluates them three times per diffusion move); the values are identical.
template <class WF, int METROP>
MOLE_D bool mole_move_state(const WfParams& p, Walker<WF>& wk, int e, double param, double sd, RngKey key,
                            uint64_t wid, uint32_t step, uint32_t compat) {
  constexpr int NE = WF::NE;
  double xn[3];
  if (METROP == MOLE_METROP_BOX) {
    const MoveDraw d = mole_draw_uniform4(key, wid, step, DOM_MOVE, (uint32_t)e);
    const double lo = -0.5 * param, scale = 0.5 * param - lo;   // Range::new(-b/2, b/2)
    xn[0] = wk.st.x[3 * e] + (lo + scale * d.a);
    xn[1] = wk.st.x[3 * e + 1] + (lo + scale * d.b);
    xn[2] = wk.st.x[3 * e + 2] + (lo + scale * d.c);
    typename WF::State tr = wk.st;
    WF::move(p, tr, e, xn);
    const double pn = WF::psi(p, tr);
    const double A = mole_clamp_acceptance((pn * pn) / (wk.psi * wk.psi), compat);   // metrop.rs:80
    if (A > d.u) {
      wk.st = tr;
      wk.psi = pn;
      return true;
    }
    return false;
  } else {
    const MoveDraw d = mole_draw_normal3_uniform1(key, wid, step, DOM_MOVE, (uint32_t)e);
    const double inv_o = m_rcp(wk.psi);
    xn[0] = (wk.st.x[3 * e] + wk.g[3 * e] * inv_o * param) + sd * d.a;           // metrop.rs:155-160
    xn[1] = (wk.st.x[3 * e + 1] + wk.g[3 * e + 1] * inv_o * param) + sd * d.b;
    xn[2] = (wk.st.x[3 * e + 2] + wk.g[3 * e + 2] * inv_o * param) + sd * d.c;
    typename WF::State tr = wk.st;
    WF::move(p, tr, e, xn);
    const double pn = WF::psi(p, tr);
    double gn[3 * NE];
    WF::grad(p, tr, gn);
    // node test, metrop.rs:178-180 (signum of NaN is NaN and NaN != NaN rejects)
    if (isnan(pn) || isnan(wk.psi) || (signbit(pn) != signbit(wk.psi))) return false;
    const double inv_n = m_rcp(pn);
    double sh = 0.0, sl = 0.0;                                   // Frobenius norms over ALL electrons, :182-193
#pragma unroll
    for (int i = 0; i < 3 * NE; ++i) {
      const double dx = wk.st.x[i] - tr.x[i];
      const double a = dx - gn[i] * inv_n * param;
      const double b = -dx - wk.g[i] * inv_o * param;
      sh = fma(a, a, sh);
      sl = fma(b, b, sl);
    }
    const double i2t = 0.5 / param;
    const double shs = sh * i2t, sls = sl * i2t;                 // -ln t_high, -ln t_low
    const double ao = fabs(wk.psi), an = fabs(pn);
    double A;
    if (fmax(shs, sls) < 200.0 && ao > 1e-40 && ao < 1e40 && an > 1e-40 && an < 1e40) {
      // none of the reference's intermediate products (t >= e^-200, psi^2 within 1e-80 .. 1e80) can leave the normal
      // range: (t_high psi'^2) / (t_low psi^2) = (psi'/psi)^2 exp((s_low - s_high)/2tau), one exponential and no
      // division, the same real number to ~1e-16
      const double q = pn * inv_o;
      A = mole_clamp_acceptance((q * q) * m_exp(sls - shs), compat);
    } else {
      // next to a node (t denormal or 0, t psi^2 underflowing, 0/0): the reference's own sequence of operations
      const double targ[2] = {-shs, -sls};
      double tv[2];
      m_exp_n<2, true>(targ, tv);
      A = mole_clamp_acceptance(tv[0] * (pn * pn) / (tv[1] * (wk.psi * wk.psi)), compat);           // :195
    }
    if (A > d.u) {
      wk.st = tr;
      wk.psi = pn;
#pragma unroll
      for (int i = 0; i < 3 * NE; ++i) wk.g[i] = gn[i];
      return true;
    }
    return false;
  }
}

// act_on/psi for the Hamiltonian kinds (operator.rs:59-61,94-96,122-124,146-148,181-183)
template <class WF>
MOLE_D double mole_local_energy(const WfParams& p, const HamParams& h, const typename WF::State& st, double psi,
                                double& hpsi_out, double& kin_psi) {
  double hpsi = 0.0;
  kin_psi = 0.0;
  if (mole_op_has_kinetic(h.kind)) {
    kin_psi = -0.5 * WF::lap(p, st);
    hpsi = kin_psi;
  }
  if (h.kind != MOLE_OP_KINETIC) hpsi = fma(mole_potential<WF::NE>(h, st.x), psi, hpsi);
  hpsi_out = hpsi;
  return hpsi * m_rcp(psi);
}

// ------------------------------------------------------------------ init
__global__ void init_kernel(double* x, int64_t W, int ne, uint64_t walker_offset, RngKey key, int normal, double a,
                            double b, int broadcast) {
  const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= W) return;
  const uint64_t wid = broadcast ? 0ull : walker_offset + (uint64_t)w;
  for (int e = 0; e < ne; ++e) {
    double v[3];
    if (normal) {                                               // dmc.rs:52-56
      const MoveDraw d = mole_draw_normal3_uniform1(key, wid, 0, DOM_INIT, (uint32_t)e);
      v[0] = a * d.a; v[1] = a * d.b; v[2] = a * d.c;
    } else {                                                    // samplers.rs:46
      const MoveDraw d = mole_draw_uniform4(key, wid, 0, DOM_INIT, (uint32_t)e);
      const double scale = b - a;
      v[0] = a + scale * d.a; v[1] = a + scale * d.b; v[2] = a + scale * d.c;
    }
    for (int c = 0; c < 3; ++c) x[(size_t)(3 * e + c) * W + w] = v[c];
  }
}

// AoS (W,ne,3) <-> SoA [3ne][W]
__global__ void aos_to_soa_kernel(const double* aos, double* soa, int64_t W, int n, int broadcast) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= W * n) return;
  const int64_t w = i % W;
  const int c = (int)(i / W);
  soa[i] = aos[(broadcast ? 0 : w * n) + c];
}
__global__ void soa_to_aos_kernel(const double* soa, double* aos, int64_t W, int n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= W * n) return;
  const int64_t w = i / n;
  const int c = (int)(i % n);
  aos[i] = soa[(size_t)c * W + w];
}
__global__ void fill_kernel(double* p, int64_t n, double v) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// ------------------------------------------------------------------ batched evaluation
template <int KIND>
__global__ void eval_kernel(const double* __restrict__ x, int64_t W, WfParams p, HamParams h, int have_ham,
                            double* psi, double* grad, double* lap, double* hpsi, double* pgrad) {
  mole_math_smem_init();
  using WF = WfDev<KIND>;
  constexpr int NE = WF::NE;
  const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= W) return;
  typename WF::State st;
#pragma unroll
  for (int c = 0; c < 3 * NE; ++c) st.x[c] = x[(size_t)c * W + w];
  WF::init(p, st);
  const double ps = WF::psi(p, st);
  if (psi) psi[w] = ps;
  if (grad) {
    double g[3 * NE];
    WF::grad(p, st, g);
#pragma unroll
    for (int c = 0; c < 3 * NE; ++c) grad[(size_t)w * 3 * NE + c] = g[c];
  }
  if (lap) lap[w] = WF::lap(p, st);
  if (hpsi && have_ham) {
    double hp, kp;
    mole_local_energy<WF>(p, h, st, ps, hp, kp);
    hpsi[w] = hp;
  }
  if (pgrad && WF::NP > 0) {
    double pg[WF::NP > 0 ? WF::NP : 1];
    WF::pgrad(p, st, pg);
#pragma unroll
    for (int k = 0; k < WF::NP; ++k) pgrad[(size_t)w * WF::NP + k] = pg[k];
  }
}

// ------------------------------------------------------------------ fused sweep
// One thread per walker (grid-stride).  Per launch: load state -> n_sweeps x (N_e Metropolis moves
// [+ sample E_L, O_k, moments]) -> store state -> block-tree reduction of the accumulators.
// Occupancy target: 16 warps/SM (128 registers) for every kind.  Measured (B200, SR sweeps with moments): +3..4 % for the
// one- and two-electron STO / Gaussian kinds at 2^20 walkers and +10 % at 2^16.  The H2 Heitler-London and two-electron
// LCAO kinds spill 100-200 bytes per thread at 128 registers and ran at 12 warps/SM (156 registers, no spill) until the
// second round; timed side by side they are equal at 2^18 and 2^20 walkers (H2 13.81 / 13.76 ms for 2^20 x 200 sweeps,
// LCAO singlet 17.12 / 17.09) and the 16-warp build is 10-12 % faster at 2^16 walkers x 2500 sweeps (H2 14.48 -> 13.18 ms,
// LCAO singlet 18.05 -> 16.09), where 512 CTAs fit one wave of 592 instead of one and a tail of 444.
#ifndef MOLE_SWEEP_MIN_CTAS
#define MOLE_SWEEP_MIN_CTAS(KIND) 4
#endif
template <int KIND, int METROP, bool OPT>
__global__ void __launch_bounds__(SWEEP_THREADS, MOLE_SWEEP_MIN_CTAS(KIND)) sweep_kernel(const SweepParams sp) {
  mole_math_smem_init();
  using WF = WfDev<KIND>;
  constexpr int NE = WF::NE;
  constexpr int NP = OPT ? WF::NP : 0;
  constexpr bool MOM_SMEM = NP >= 4;
  using A = Acc<NP, MOM_SMEM>;
  __shared__ double s_mom[MOM_SMEM ? A::NMOM * SWEEP_THREADS : 1];
  A acc;
  acc.zero(s_mom);
  const WfParams& p = sp.wf;
  const double sd = sqrt(sp.metrop_param);
  const bool want_e = (sp.observables & MOLE_OBS_ENERGY) != 0;
  const int64_t W = sp.W;
  const int64_t nsamp = sp.n_sweeps - sp.n_discard;

  for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < W; w += (int64_t)gridDim.x * blockDim.x) {
    Walker<WF> wk;
#pragma unroll
    for (int c = 0; c < 3 * NE; ++c) wk.st.x[c] = sp.x[(size_t)c * W + w];
    WF::init(p, wk.st);
    wk.refresh(p);
    const uint64_t wid = sp.walker_offset + (uint64_t)w;
    double blk = sp.blk[w];
    int fill = sp.blk_fill, nbad = 0;   // nbad: skipped samples of the open block (not carried across launches)

    for (int s = 0; s < sp.n_sweeps; ++s) {
      const uint32_t step = sp.step0 + (uint32_t)s;
#pragma unroll
      for (int e = 0; e < NE; ++e) {                            // Sampler::move_state, samplers.rs:106-117
        const bool ok = mole_move_state<WF, METROP>(p, wk, e, sp.metrop_param, sd, sp.key, wid, step, sp.compat);
        acc.v[ACC_NACC] += ok ? 1.0 : 0.0;
        acc.v[ACC_NMOVE] += 1.0;
        if (sp.tr_accept) sp.tr_accept[((size_t)s * NE + e) * W + w] = ok ? 1 : 0;
      }
      if (s < sp.n_discard) continue;                           // block 0 = equilibration, montecarlo.rs:36
      const int64_t si = s - sp.n_discard;
      // Sampler::sample, samplers.rs:81-104.  A sample whose E_L or O_k is not finite (psi underflowed to 0, two
      // particles on top of each other) is written to the traces as it is but kept out of every sum and counted
      // in acc[ACC_BAD] (mole_ensemble_health): upstream it would turn the mean, the gradient and S into NaN.
      double el = 0.0, hpsi, kin_psi = 0.0;
      const double inv = m_rcp(wk.psi);
      double pg[NP > 0 ? NP : 1], o[NP > 0 ? NP : 1];
      const bool quirk = (sp.compat & MOLE_COMPAT_VECTOR_DIV) != 0;
      bool bad = false;
      if (want_e) {
        el = mole_local_energy<WF>(p, sp.ham, wk.st, wk.psi, hpsi, kin_psi);
        bad = !isfinite(el);
        if (sp.tr_energy) sp.tr_energy[(size_t)si * W + w] = el;
      }
      if (OPT && NP > 0) {
        WF::pgrad(p, wk.st, pg);
#pragma unroll
        for (int k = 0; k < NP; ++k) {
          // intended O_k = d_k psi / psi; MOLE_COMPAT_VECTOR_DIV: stored sample 1/d_k psi, O_k = 1/(psi d_k psi)
          o[k] = quirk ? inv / pg[k] : pg[k] * inv;
          bad = bad || !isfinite(o[k]);
          if (sp.tr_pgrad) sp.tr_pgrad[((size_t)si * NP + k) * W + w] = quirk ? 1.0 / pg[k] : pg[k];
        }
      }
      if (bad) atomicAdd(sp.acc + ACC_BAD, 1.0);                // rare; integer-valued, so the order does not matter
      if (want_e) {
        if (!bad) {
          acc.v[ACC_N] += 1.0;
          acc.v[ACC_E] += el;
          acc.v[ACC_E2] = fma(el, el, acc.v[ACC_E2]);
          acc.v[ACC_T] += kin_psi * inv;
          blk += el;
        } else {
          ++nbad;
        }
        if (++fill == sp.block_size) {                          // block means, vmc.rs:158-164
          if (nbad < sp.block_size) {
            const double bm = blk / (double)(sp.block_size - nbad);
            acc.v[ACC_B] += bm;
            acc.v[ACC_B2] = fma(bm, bm, acc.v[ACC_B2]);
            acc.v[ACC_NB] += 1.0;
          }
          blk = 0.0;
          fill = 0;
          nbad = 0;
        }
      }
      if (sp.observables & MOLE_OBS_KINETIC) {
        const double k = want_e ? kin_psi * inv : -0.5 * WF::lap(p, wk.st) * inv;
        if (sp.tr_kinetic) sp.tr_kinetic[(size_t)si * W + w] = k;
      }
      if (!bad) acc.v[ACC_PSI] += wk.psi;
      if (sp.tr_wfvalue) sp.tr_wfvalue[(size_t)si * W + w] = wk.psi;   // psi^2/psi, operators.rs:22 + samplers.rs:90
      if (OPT && NP > 0 && !bad) {
        int q = 0;
#pragma unroll
        for (int k = 0; k < NP; ++k) {
          acc.add(k, o[k], 1.0);
          acc.add(NP + k, o[k], el);
#pragma unroll
          for (int l = k; l < NP; ++l, ++q) acc.add(2 * NP + q, o[k], o[l]);
        }
      }
    }
#pragma unroll
    for (int c = 0; c < 3 * NE; ++c) sp.x[(size_t)c * W + w] = wk.st.x[c];
    sp.blk[w] = blk;
  }
  double all[A::LEN];
  acc.gather(all);
  __syncthreads();
  mole_block_reduce_to_global<A::LEN>(all, sp.partials, sp.acc, sp.ticket, [](int i) { return A::slot(i); });
}

// Last-CTA fold of the per-CTA DMC partial rows {sum w E, sum w, sum w', max w'} into red[0..3] by the
// whole CTA: thread t takes rows t, t+T, ... in order, then warp shuffles and a per-warp pass in a fixed
// order - deterministic for a given launch geometry (a serial fold over ~2000 rows cost 100 us per step).
MOLE_D void mole_dmc_fold_partials(const double* partials, unsigned rows, double* red, double (*sm)[4]) {
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
  for (unsigned b = threadIdx.x; b < rows; b += blockDim.x) {
    const double* r = partials + (size_t)b * 4;
    a0 += __ldcg(r); a1 += __ldcg(r + 1); a2 += __ldcg(r + 2); a3 = fmax(a3, __ldcg(r + 3));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a0 += __shfl_xor_sync(0xffffffffu, a0, o);
    a1 += __shfl_xor_sync(0xffffffffu, a1, o);
    a2 += __shfl_xor_sync(0xffffffffu, a2, o);
    a3 = fmax(a3, __shfl_xor_sync(0xffffffffu, a3, o));
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (lane == 0) { sm[warp][0] = a0; sm[warp][1] = a1; sm[warp][2] = a2; sm[warp][3] = a3; }
  __syncthreads();
  if (threadIdx.x < 4) {
    double s = 0.0;
    for (int q = 0; q < nwarp; ++q) s = (threadIdx.x == 3) ? fmax(s, sm[q][3]) : s + sm[q][threadIdx.x];
    red[threadIdx.x] = s;
  }
}

// ------------------------------------------------------------------ DMC time step (dmc.rs:87-130)
// one walker's time step on values held in registers: drift-diffusion move of every electron, weight update,
// contribution to the step's sums.  In: xr = configuration, e_io = E_L before the move when el_cached, w_in = weight.
// Out: xr = configuration after the move, e_io = E_L after it (0 for a killed walker), returns the new weight.
template <int KIND>
MOLE_D double mole_dmc_walker_core(const DmcParams& dp, double* xr, double w_in, double& e_io, int el_cached, int64_t w, uint32_t step,
                                   double sd, double& s_we, double& s_w, double& s_wn, double& m_wn) {
  using WF = WfDev<KIND>;
  constexpr int NE = WF::NE;
  const WfParams& p = dp.wf;
  Walker<WF> wk;
#pragma unroll
  for (int c = 0; c < 3 * NE; ++c) wk.st.x[c] = xr[c];
  WF::init(p, wk.st);
  wk.refresh(p);
  double hp, kp;
  // E_L before the move (dmc.rs:89-96).  It equals the post-move E_L of the previous step for the
  // same configuration, so it is carried in `el` (and gathered by branching) instead of recomputed.
  const double e_old = el_cached ? e_io : mole_local_energy<WF>(p, dp.ham, wk.st, wk.psi, hp, kp);
  const uint64_t wid = dp.walker_offset + (uint64_t)w;
#pragma unroll
  for (int e = 0; e < NE; ++e) mole_move_state<WF, MOLE_METROP_DIFFUSE>(p, wk, e, dp.tau_move, sd, dp.key, wid, step, dp.compat);
  const double e_new = mole_local_energy<WF>(p, dp.ham, wk.st, wk.psi, hp, kp);   // :115-124
  // a walker whose local energy is not finite (upstream: NaN ensemble energy from here on) is counted
  // (acc[ACC_BAD_DMC]) and dies: weight 0 before and after the step, never picked by the brancher
  const double w_up = w_in * exp(-dp.tau_weight * ((e_old + e_new) / 2.0 - dp.e_ref));           // :126-128
  const bool bad = !isfinite(e_old) || !isfinite(e_new) || !isfinite(w_up);
  if (bad) atomicAdd(dp.health, 1.0);
  const double wt = bad ? 0.0 : w_in;
  s_we = fma(wt, bad ? 0.0 : e_old, s_we);                    // dmc.rs:112-113
  s_w += wt;
  const double wn = bad ? 0.0 : w_up;
  s_wn += wn;
  m_wn = fmax(m_wn, wn);
  e_io = bad ? 0.0 : e_new;
#pragma unroll
  for (int c = 0; c < 3 * NE; ++c) xr[c] = wk.st.x[c];
  return wn;
}

// the same from / to the ensemble arrays.  x / wgt / el are read through L2 (__ldcg): inside dmc_block_kernel other
// CTAs wrote them earlier in the same launch.
template <int KIND>
MOLE_D void mole_dmc_walker_step(const DmcParams& dp, double* x, double* wgt, double* el, int64_t w, int el_cached, uint32_t step,
                                 double sd, double& s_we, double& s_w, double& s_wn, double& m_wn) {
  constexpr int NE = WfDev<KIND>::NE;
  const int64_t W = dp.W;
  double xr[3 * NE];
#pragma unroll
  for (int c = 0; c < 3 * NE; ++c) xr[c] = __ldcg(x + (size_t)c * W + w);
  double e_io = el_cached ? __ldcg(el + w) : 0.0;
  const double wn = mole_dmc_walker_core<KIND>(dp, xr, __ldcg(wgt + w), e_io, el_cached, w, step, sd, s_we, s_w, s_wn, m_wn);
  wgt[w] = wn;
  el[w] = e_io;
#pragma unroll
  for (int c = 0; c < 3 * NE; ++c) x[(size_t)c * W + w] = xr[c];
}

// CTA reduction of the step's sums (3 sums + 1 max) into one row of `partials`; fixed order for a given CTA shape
MOLE_D void mole_dmc_cta_reduce(double s_we, double s_w, double s_wn, double m_wn, double (*sm)[4], double* row) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s_we += __shfl_xor_sync(0xffffffffu, s_we, o);
    s_w += __shfl_xor_sync(0xffffffffu, s_w, o);
    s_wn += __shfl_xor_sync(0xffffffffu, s_wn, o);
    m_wn = fmax(m_wn, __shfl_xor_sync(0xffffffffu, m_wn, o));
  }
  if (lane == 0) { sm[warp][0] = s_we; sm[warp][1] = s_w; sm[warp][2] = s_wn; sm[warp][3] = m_wn; }
  __syncthreads();
  if (threadIdx.x < 4) {
    double s = 0.0;
    for (int q = 0; q < nwarp; ++q) s = (threadIdx.x == 3) ? fmax(s, sm[q][3]) : s + sm[q][threadIdx.x];
    row[threadIdx.x] = s;
  }
}

// red[0] += sum w E_old, red[1] += sum w (pre-update), red[2] = sum w (post-update), red[3] = max w (post-update)
template <int KIND>
__global__ void __launch_bounds__(SWEEP_THREADS) dmc_step_kernel(const DmcParams dp) {
  mole_math_smem_init();
  const double sd = sqrt(dp.tau_move);
  double s_we = 0.0, s_w = 0.0, s_wn = 0.0, m_wn = 0.0;
  const int64_t W = dp.W;
  for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < W; w += (int64_t)gridDim.x * blockDim.x)
    mole_dmc_walker_step<KIND>(dp, dp.x, dp.w, dp.el, w, dp.el_cached, dp.step, sd, s_we, s_w, s_wn, m_wn);
  __shared__ double sm[32][4];
  __shared__ bool is_last;
  mole_dmc_cta_reduce(s_we, s_w, s_wn, m_wn, sm, dp.partials + (size_t)blockIdx.x * 4);
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = (atomicInc(dp.ticket, gridDim.x - 1) == gridDim.x - 1);
  __syncthreads();
  if (is_last) {
    __threadfence();
    mole_dmc_fold_partials(dp.partials, gridDim.x, dp.red, sm);
  }
}

// ------------------------------------------------------------------ math probe
__global__ void math_probe_kernel(int which, const double* __restrict__ in, int64_t n, double* out) {
  mole_math_smem_init();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double x = in[i];
  double r, y;
  switch (which) {
    case 0: y = m_exp(x); break;
    case 1: y = m_rcp(x); break;
    case 2: y = m_rsqrt(x); break;
    case 3: y = m_sqrt_rsqrt(x, r); break;
    case 4: { const double a[1] = {x}; double o[1]; m_log_n<1>(a, o); y = o[0]; break; }
    default: {
      const double a[1] = {x};
      double s[1], c[1];
      m_sincos_turn_n<1>(a, s, c);
      y = which == 5 ? s[0] : c[0];
      break;
    }
  }
  out[i] = y;
}

// ------------------------------------------------------------------ FP64 peak probe
__global__ void dfma_peak_kernel(double* out, int iters, double seed) {
  double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double m = 1.0000001, c = 1e-9;
#pragma unroll 8
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
    a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
  }
  const double s = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
  if (s == 123.456) out[blockIdx.x] = s;  // never true; keeps the chain alive
}
