/*
 * mole_b200.h — C ABI of the B200-native walker-ensemble VMC/DMC hot path of Jvanrhijn/mole.
 *
 * The reference has no FFI: its boundary is the Rust trait surface re-exported by
 * `mole::prelude` (src/lib.rs:10-21).  Every entry point below names the reference item it
 * replaces (paths relative to the reference tree).  A Rust `-sys` crate binds this header
 * verbatim (see INTEGRATION.md and rust/mole-b200-sys/); the safe wrapper implements the
 * original traits by delegating to these symbols.
 *
 * Conventions
 *  - every call returns an int32 status (MOLE_OK == 0); nothing unwinds across the boundary;
 *  - all pointers are caller-owned HOST memory unless the name ends in `_dev`;
 *  - configurations are row-major (N_e, 3) fp64 exactly like the reference's `Array2<f64>`;
 *    ensembles are passed as (W, N_e, 3); on the device they are stored structure-of-arrays;
 *  - one mole_ctx per GPU; calls on one ctx are not re-entrant, different ctxs may be driven
 *    from different threads (the analogue of the `Send + Sync` bounds, src/vmc/src/vmc.rs:29-32);
 *  - there is NO CPU fallback: with no usable sm_100 device every device entry point fails with
 *    MOLE_ERR_CUDA / MOLE_ERR_NO_DEVICE.
 */
#ifndef MOLE_B200_H
#define MOLE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status codes: 1..6 mirror `errors::Error` one to one (src/errors/src/lib.rs:8-15) ---- */
enum {
  MOLE_OK = 0,
  MOLE_ERR_LINALG = 1,                /* Error::LinalgError   (singular SR matrix, optimizers.rs:251) */
  MOLE_ERR_SHAPE = 2,                 /* Error::ShapeError */
  MOLE_ERR_FUNC = 3,                  /* Error::FuncError     (e.g. gradient of WaveFunctionMock) */
  MOLE_ERR_OPERATOR_VALUE_ACCESS = 4, /* Error::OperatorValueAccessError */
  MOLE_ERR_DATA_ACCESS = 5,           /* Error::DataAccessError (observable missing, util.rs:12-27) */
  MOLE_ERR_EMPTY_CACHE = 6,           /* Error::EmptyCacheError */
  MOLE_ERR_CUDA = 100,
  MOLE_ERR_NCCL = 101,
  MOLE_ERR_INVALID_ARG = 102,
  MOLE_ERR_NO_DEVICE = 103,
  MOLE_ERR_ASSERT = 104               /* reference `assert!` failures (montecarlo.rs:29) */
};

typedef struct mole_ctx_s* mole_ctx_t;
typedef struct mole_wf_s* mole_wf_t;
typedef struct mole_op_s* mole_op_t;
typedef struct mole_ens_s* mole_ens_t;
typedef struct mole_metrop_s* mole_metrop_t;
typedef struct mole_opt_s* mole_opt_t;

/* ---- context ---------------------------------------------------------------------------- */
int32_t mole_ctx_create(int32_t device, mole_ctx_t* ctx);
/* A context destroyed while ensembles created on it are alive is freed with the last of them. */
int32_t mole_ctx_destroy(mole_ctx_t ctx);
int32_t mole_ctx_synchronize(mole_ctx_t ctx);
const char* mole_last_error_string(mole_ctx_t ctx); /* ctx may be NULL: last global error */
/* the ctx's CUDA stream (cudaStream_t) so that callers can record their own events on it */
int32_t mole_ctx_stream(mole_ctx_t ctx, void** stream);
int32_t mole_version(void);

/* ---- trial wavefunction descriptors -------------------------------------------------------
 * Replaces user impls of Function<f64,D=Ix2> / Differentiate / WaveFunction / Optimize
 * (src/wavefunction_traits/src/lib.rs:7-24, src/optimize/src/traits.rs:8-16).  Arbitrary Rust
 * impls cannot run on the device, so the closed set of kinds used by the reference's
 * examples/ and tests/ is provided. */
enum {
  MOLE_WF_STO_1S = 0,         /* examples/dmc.rs:95-149                    geom: -            P=1 (alpha) */
  MOLE_WF_GAUSSIAN = 1,       /* examples/dmc.rs:34-92, tests/sho_optimize.rs:52-111          P=1 (a)     */
  MOLE_WF_STO_PRODUCT = 2,    /* examples/helium_atom_singlet.rs:35-118    geom: -            P=1 (alpha) */
  MOLE_WF_H2_HL_STO = 3,      /* examples/hydrogen_molecule.rs:65-168      geom: [R]          P=1 (alpha) */
  MOLE_WF_H2P_PRODUCT = 4,    /* tests/hydrogen_molecular_ion_lcao.rs:51-98 geom: [R], params[0]=alpha, P=0 */
  MOLE_WF_SLATER_JASTROW = 5, /* SURVEY.md §8(c) config 5 + theory/jastrow.tex; geom: [kappa,n_up,n_dn];
                                 params: zeta1,zeta2,zeta3,b1,b2,b3,b4 (P=7) */
  MOLE_WF_CONSTANT = 6,       /* src/metropolis/src/metrop.rs:225-255 WaveFunctionMock; geom: [value] */
  /* LCAO determinants over a hydrogen-1s basis chi_c(r) = exp(-|r - R_c| / width): the Hydrogen1sBasis / Orbital /
     SingleDeterminant / SpinDeterminantProduct API that tests/helium_lcao.rs:94-101 and
     tests/hydrogen_molecular_ion_lcao.rs:103-107 name (commented out upstream).  geom: [mode, 1/width, R_0 (3), R_1 (3)];
     params: orbital coefficients C[k][c] at k * n_centres + c, all variational.  mode 0: spin product
     phi_0(x_0) phi_1(x_1) (SpinDeterminantProduct, n_up = 1); mode 1: 2x2 determinant (SingleDeterminant). */
  MOLE_WF_LCAO_1E_2C = 7,     /* one electron, two centres (H2+):   psi = phi_0(x_0)             P=2 */
  MOLE_WF_LCAO_2E_1C = 8,     /* two electrons, one centre (He):    two orbitals                 P=2 */
  MOLE_WF_LCAO_2E_2C = 9,     /* two electrons, two centres (H2 MO): two orbitals                P=4 */
  /* The same API in general: SpinDeterminantProduct of n_orb = max(n_up, n_dn) LCAO orbitals over
     N_c <= 8 centres, n_up, n_dn <= 5, both spins sharing the orbitals, times the e-e Jastrow of theory/jastrow.tex
     (b = 0 switches it off).  geom: [kappa, n_up, n_dn, N_c, 0, 0, 0, 0, (R_c x, y, z, alpha_c = 1/width_c) per centre];
     params: C[k][c] at k N_c + c, then b1..b4: P = n_orb N_c + 4 <= 44.  With P > MOLE_ACC_MAX_PARAMS the optimisation
     moments are the Gram matrix of the per-sample rows (mole_gram_*), not mole_acc_host. */
  MOLE_WF_LCAO_SJ = 10
};
#define MOLE_WF_MAX_PARAMS 48
#define MOLE_WF_MAX_GEOM 40
#define MOLE_WF_MAX_ELEC 10

typedef struct {
  int32_t kind;
  int32_t n_elec;
  int32_t n_params;
  int32_t reserved;
  double params[MOLE_WF_MAX_PARAMS]; /* Optimize::parameters */
  double geom[MOLE_WF_MAX_GEOM];     /* fixed constants, see the kind list */
} mole_wf_desc;

int32_t mole_wf_create(mole_ctx_t ctx, const mole_wf_desc* desc, mole_wf_t* wf);
int32_t mole_wf_destroy(mole_wf_t wf);
int32_t mole_wf_num_electrons(mole_wf_t wf, int32_t* n);            /* WaveFunction::num_electrons */
int32_t mole_wf_num_parameters(mole_wf_t wf, int32_t* n);           /* Optimize::num_parameters   */
int32_t mole_wf_get_parameters(mole_wf_t wf, double* params);       /* Optimize::parameters       */
int32_t mole_wf_update_parameters(mole_wf_t wf, const double* deltap); /* Optimize::update_parameters */
int32_t mole_wf_set_parameters(mole_wf_t wf, const double* params);
/* pointwise entry points, evaluated ON THE DEVICE for one configuration (parity tests):        */
int32_t mole_wf_value(mole_wf_t wf, const double* cfg, double* out);              /* Function::value          */
int32_t mole_wf_gradient(mole_wf_t wf, const double* cfg, double* out);           /* Differentiate::gradient  */
int32_t mole_wf_laplacian(mole_wf_t wf, const double* cfg, double* out);          /* Differentiate::laplacian */
int32_t mole_wf_parameter_gradient(mole_wf_t wf, const double* cfg, double* out); /* Optimize::parameter_gradient */

/* ---- local operators (src/operator/src/operator.rs) ----------------------------------------- */
enum {
  MOLE_OP_KINETIC = 0,    /* KineticEnergy          operator.rs:103-125 */
  MOLE_OP_IONIC_POT = 1,  /* IonicPotential         operator.rs:16-62   */
  MOLE_OP_ELEC_POT = 2,   /* ElectronicPotential    operator.rs:68-97   */
  MOLE_OP_IONIC = 3,      /* IonicHamiltonian       operator.rs:130-149 */
  MOLE_OP_ELECTRONIC = 4, /* ElectronicHamiltonian  operator.rs:156-184 */
  MOLE_OP_HARMONIC = 5    /* HarmonicHamiltonian    examples/custom_operator.rs:30-61 */
};
#define MOLE_OP_MAX_IONS 8
typedef struct {
  int32_t kind;
  int32_t n_ions;
  double ion_pos[MOLE_OP_MAX_IONS * 3]; /* (N_n,3) row-major */
  int32_t ion_charge[MOLE_OP_MAX_IONS]; /* Array1<i32>, operator.rs:19 */
  double frequency;                     /* HarmonicHamiltonian::frequency */
} mole_op_desc;

int32_t mole_op_create(mole_ctx_t ctx, const mole_op_desc* desc, mole_op_t* op);
int32_t mole_op_destroy(mole_op_t op);
/* LocalOperator::act_on (src/operator/src/traits.rs:263-265): returns H psi, NOT divided by psi */
int32_t mole_op_act_on(mole_op_t op, mole_wf_t wf, const double* cfg, double* out);

/* ---- walker ensemble -------------------------------------------------------------------------
 * Structure-of-arrays fp64 ensemble resident in HBM.  Walker w of this ensemble has the global id
 * walker_offset + w, which (with the sweep counter) keys its Philox stream: results do not depend
 * on how walkers are sharded over GPUs.  Replaces Sampler::{new,with_initial_configuration}
 * (src/montecarlo/src/samplers.rs:40-71), the per-worker clones of src/vmc/src/vmc.rs:56 and
 * DmcRunner::new's walker vector (src/dmc/src/dmc.rs:49-58). */
int32_t mole_ensemble_create(mole_ctx_t ctx, int64_t n_walkers, int32_t n_elec, const uint8_t seed[32],
                             uint64_t walker_offset, mole_ens_t* ens);
int32_t mole_ensemble_destroy(mole_ens_t ens);
int32_t mole_ensemble_num_walkers(mole_ens_t ens, int64_t* n);
/* cfg ~ U(lo,hi): samplers.rs:45-46.  broadcast_walker0 != 0 gives every walker the configuration
 * of global walker 0 (the reference clones ONE sampler, vmc.rs:56). */
int32_t mole_ensemble_init_uniform(mole_ens_t ens, double lo, double hi, int32_t broadcast_walker0);
/* cfg ~ N(0,sigma): dmc.rs:49-58 (`vec![x; n]` clones one draw: broadcast_walker0 = 1 is faithful) */
int32_t mole_ensemble_init_normal(mole_ens_t ens, double sigma, int32_t broadcast_walker0);
int32_t mole_ensemble_set_configs(mole_ens_t ens, const double* cfgs /* (W,N_e,3) */);
int32_t mole_ensemble_set_configs_broadcast(mole_ens_t ens, const double* cfg /* (N_e,3) */);
int32_t mole_ensemble_get_configs(mole_ens_t ens, double* cfgs /* (W,N_e,3) */);
int32_t mole_ensemble_set_weights(mole_ens_t ens, const double* w);
int32_t mole_ensemble_get_weights(mole_ens_t ens, double* w);
/* device-side copy of the current configurations / restore from it (the master sampler whose
 * configuration every per-iteration clone starts from, vmc.rs:56) */
int32_t mole_ensemble_snapshot(mole_ens_t ens);
int32_t mole_ensemble_restore(mole_ens_t ens);
/* Metropolis::reseed_rng (src/metropolis/src/traits.rs:33): new key, sweep counter back to 0 */
int32_t mole_ensemble_reseed(mole_ens_t ens, const uint8_t seed[32]);
int32_t mole_ensemble_set_step(mole_ens_t ens, uint32_t step);
int32_t mole_ensemble_get_step(mole_ens_t ens, uint32_t* step);
/* Metropolis::generate_seed (traits.rs:35-37): the n-th 32-byte seed derived from `master` */
int32_t mole_derive_seed(const uint8_t master[32], uint32_t n, uint8_t out[32]);

/* ---- batched evaluation on identical walker configurations (parity entry point) ---------------
 * psi[W], grad[W*N_e*3] (un-normalised), lap[W], hpsi[W] (= act_on, un-normalised; needs op),
 * pgrad[W*P] (un-normalised d psi/d p).  Any output may be NULL. */
int32_t mole_eval_vgl(mole_ens_t ens, mole_wf_t wf, mole_op_t op, double* psi, double* grad, double* lap,
                      double* hpsi, double* pgrad);

/* ---- Metropolis samplers (src/metropolis/src/metrop.rs) ---------------------------------------- */
enum { MOLE_METROP_BOX = 0 /* metrop.rs:22-101 */, MOLE_METROP_DIFFUSE = 1 /* metrop.rs:103-217 */ };
int32_t mole_metropolis_create(int32_t kind, double param /* box_side | time_step */, mole_metrop_t* m);
int32_t mole_metropolis_destroy(mole_metrop_t m);
/* MOLE_COMPAT_* bits that concern the sampler itself (MOLE_COMPAT_NAN_ACCEPT); OR-ed with the
 * per-call compat of mole_sweep and used by mole_dmc_step */
int32_t mole_metropolis_set_compat(mole_metrop_t m, uint32_t compat);

/* ---- fused sweep: Sampler::move_state + Sampler::sample + Runner::run's block loop -------------
 * (samplers.rs:81-117, montecarlo.rs:24-46) for all walkers, n_sweeps sweeps in ONE launch. */
enum {
  MOLE_OBS_ENERGY = 1,  /* "Energy" => op                                                        */
  MOLE_OBS_PGRAD = 2,   /* "Parameter gradient" => ParameterGradient   src/vmc/src/operators.rs:7-14  */
  MOLE_OBS_WFVALUE = 4, /* "Wavefunction value" => WavefunctionValue  src/vmc/src/operators.rs:16-24 */
  MOLE_OBS_KINETIC = 8  /* "Kin. Energy" => KineticEnergy             examples/helium_atom_singlet.rs:162 */
};
/* compat flags */
enum {
  /* reproduce `Vector / Scalar == scalar / array` (src/operator/src/traits.rs:149-150): the stored
   * "Parameter gradient" sample becomes 1/(d psi/d p) and O_k = 1/(psi d psi/d p).  Default (0) is
   * the intended O_k = (d psi/d p)/psi. */
  MOLE_COMPAT_VECTOR_DIV = 1,
  /* reproduce optimizers.rs:219-224, which subtracts <O_i><O_j> from EVERY element of S */
  MOLE_COMPAT_SR_SUBTRACT = 2,
  /* reproduce `acceptance.min(1.0)` with Rust's NaN-dropping f64::min (metrop.rs:80,195): a NaN
   * acceptance ratio (0/0 when t_high and t_low both underflow next to a node, or psi = 0) becomes 1.0
   * and the move is ACCEPTED.  Default (0): a NaN ratio is rejected, which is what keeps 2^20-walker
   * ensembles of nodal wavefunctions finite (DESIGN.md "deliberate deviations").  Identical results
   * whenever the reference's own ratio is not NaN. */
  MOLE_COMPAT_NAN_ACCEPT = 4
};

typedef struct {
  int32_t n_sweeps;       /* sweeps in this call                                                    */
  int32_t n_discard;      /* leading sweeps of this call that are moved but not sampled (block 0)   */
  int32_t block_size;     /* blocking-analysis block, vmc.rs:150-170 (>=1)                          */
  uint32_t observables;   /* MOLE_OBS_* mask; accumulators are updated for the enabled ones         */
  uint32_t compat;        /* MOLE_COMPAT_*                                                          */
  uint32_t flags;         /* MOLE_SWEEP_*: keep / append the E_L series on the device (see "series") */
  /* optional per-sample traces, HOST pointers, sample-major: trace[s*W + w], s over sampled sweeps */
  double* energy_trace;   /* E_L                                   (W*n_samples) */
  double* wfvalue_trace;  /* psi  (the stored "Wavefunction value") (W*n_samples) */
  double* kinetic_trace;  /* -0.5 lap/psi                           (W*n_samples) */
  double* pgrad_trace;    /* stored "Parameter gradient", [s][k][w] (W*n_samples*P) */
  uint8_t* accept_trace;  /* accept bits, [sweep][electron][w]      (W*n_sweeps*N_e) */
} mole_sweep_args;

int32_t mole_sweep(mole_ens_t ens, mole_wf_t wf, mole_metrop_t m, mole_op_t op, const mole_sweep_args* args);

/* ---- accumulators ------------------------------------------------------------------------------
 * Device-resident sums that replace the reference's raw sample vectors
 * (concatenate_worker_data vmc.rs:108-130, process_monte_carlo_results vmc.rs:133-172,
 * compute_energy_gradient util.rs:6-46, construct_sr_matrix optimizers.rs:191-233). */
#define MOLE_ACC_MAX_PARAMS 8
typedef struct {
  double n_samples;  /* number of E_L samples (all walkers)                          */
  double sum_e;      /* sum E_L                                                      */
  double sum_e2;     /* sum E_L^2                                                    */
  double sum_b;      /* sum of block means (blocks never straddle walkers)           */
  double sum_b2;     /* sum of squared block means                                   */
  double n_blocks;   /* number of complete blocks                                    */
  double n_accept;   /* accepted single-electron moves                               */
  double n_moves;    /* proposed single-electron moves                               */
  double sum_t;      /* sum of -0.5 lap/psi                                          */
  double sum_psi;    /* sum of psi                                                   */
  double sum_o[MOLE_ACC_MAX_PARAMS];                                  /* sum O_k        */
  double sum_oe[MOLE_ACC_MAX_PARAMS];                                 /* sum O_k E_L    */
  double sum_oo[MOLE_ACC_MAX_PARAMS * (MOLE_ACC_MAX_PARAMS + 1) / 2]; /* sum O_k O_l, k<=l packed row-major */
  int32_t n_params;
  int32_t reserved;
} mole_acc_host;

int32_t mole_acc_reset(mole_ens_t ens);
int32_t mole_acc_get(mole_ens_t ens, mole_acc_host* out);
/* Health of the ensemble since the last mole_acc_reset.  The reference has no such notion: a NaN/Inf local
 * energy or parameter gradient silently poisons its sample vectors (vmc.rs:133-172) and the acceptance test
 * silently rejects a non-finite psi (metrop.rs:81,197).  Here a sample whose E_L or O_k is not finite is
 * kept OUT of every accumulator (it still appears in the traces) and counted; a DMC walker whose E_L is not
 * finite gets weight 0 (it is never picked by the brancher) and is counted.  After mole_acc_allreduce the
 * counters are sums over all ranks. */
typedef struct {
  int64_t nonfinite_samples;     /* VMC samples skipped (per walker and sweep)      */
  int64_t nonfinite_dmc_walkers; /* DMC walker-steps whose weight was zeroed        */
  int64_t reserved[2];
} mole_ens_health;
int32_t mole_ensemble_health(mole_ens_t ens, mole_ens_health* out);
/* sum the device accumulators over all ranks of the communicator (NCCL allreduce, fp64 sum) */
int32_t mole_acc_allreduce(mole_ens_t ens);
/* device pointer + length (doubles) of the packed accumulator vector, for callers that bring their
 * own collective (e.g. torch.distributed) */
int32_t mole_acc_device_ptr(mole_ens_t ens, void** ptr_dev, int32_t* n_doubles);
/* host finaliser: mean energy, blocking error (vmc.rs:150-170), acceptance, energy gradient
 * g_k = 2(<O_k E> - <O_k><E>) (util.rs:6-46).  grad may be NULL. */
int32_t mole_acc_finalize(const mole_acc_host* acc, double* energy, double* error, double* acceptance_per_sweep,
                          double* grad);

/* ---- optimisation moments of a large-P kind (P > MOLE_ACC_MAX_PARAMS) -----------------
 * A sweep with MOLE_OBS_PGRAD writes every sample's row v = (1, E_L, O_1 .. O_P) and contracts the rows into the
 * Gram matrix G = sum v v^T on the device (fp64 SYRK on the tensor cores, mma.sync m8n8k4; mole_gram_select(ens, 1)
 * runs the same contraction on the FP64 vector pipe).  G holds every moment the optimizers read: G[0][0] = n,
 * G[0][1] = sum E, G[1][1] = sum E^2, G[0][2+k] = sum O_k, G[1][2+k] = sum O_k E, G[2+k][2+l] = sum O_k O_l
 * (construct_sr_matrix optimizers.rs:191-233, compute_energy_gradient util.rs:6-46).  mole_acc_reset clears it. */
int32_t mole_gram_get(mole_ens_t ens, int32_t* n_cols, double* gram /* (P+2)^2 row-major; NULL: query n_cols */);
int32_t mole_gram_allreduce(mole_ens_t ens);
int32_t mole_gram_device_ptr(mole_ens_t ens, void** ptr_dev, int32_t* n_doubles);
int32_t mole_gram_select(mole_ens_t ens, int32_t impl /* 0: DMMA (default), 1: FP64 vector pipe */);
int32_t mole_gram_finalize(int32_t n_cols, const double* gram, double* energy, double* grad /* P, nullable */);

/* ---- NCCL communicator (walkers shard across the GPUs of one box) ------------------------------ */
#define MOLE_NCCL_UNIQUE_ID_BYTES 128
int32_t mole_comm_get_unique_id(uint8_t id[MOLE_NCCL_UNIQUE_ID_BYTES]);
int32_t mole_comm_init(mole_ctx_t ctx, int32_t nranks, int32_t rank, const uint8_t id[MOLE_NCCL_UNIQUE_ID_BYTES]);
int32_t mole_comm_destroy(mole_ctx_t ctx);

/* ---- optimizers (host side; src/optimize/src/optimizers.rs) ------------------------------------ */
enum {
  MOLE_OPT_SD = 0,       /* SteepestDescent            optimizers.rs:9-30    */
  MOLE_OPT_MOMENTUM = 1, /* MomentumDescent            optimizers.rs:32-60   */
  MOLE_OPT_NESTEROV = 2, /* NesterovMomentum           optimizers.rs:62-94   */
  MOLE_OPT_LBFGS = 3,    /* OnlineLbfgs                optimizers.rs:96-179  */
  MOLE_OPT_SR = 4        /* StochasticReconfiguration  optimizers.rs:181-253 */
};
int32_t mole_opt_create(int32_t kind, int32_t n_params, double step_size, double momentum_parameter,
                        int32_t history, uint32_t compat, mole_opt_t* opt);
int32_t mole_opt_destroy(mole_opt_t opt);
/* Optimizer::compute_parameter_update (optimize/src/traits.rs:18-25) from the reduced moments.  Non-finite moments
 * (MOLE_ERR_DATA_ACCESS), a pivot below n eps max|S| or a non-finite update (MOLE_ERR_LINALG) are refused and
 * deltap is left untouched. */
int32_t mole_opt_step(mole_opt_t opt, const double* pars, const mole_acc_host* acc, double* deltap);
/* StochasticReconfiguration only: S_kk <- S_kk * diag_scale + diag_shift before the solve.  Default (1.01, 0) is the
 * reference's hard-coded regularisation (optimizers.rs:225-231); an absolute shift is the usual stabiliser when
 * two parameters are nearly redundant (the Jastrow b1/b2 pair of the Slater-Jastrow kind). */
int32_t mole_opt_set_sr_regularization(mole_opt_t opt, double diag_scale, double diag_shift);
/* the same from the Gram matrix of a large-P kind (n_cols = P + 2) */
int32_t mole_opt_step_gram(mole_opt_t opt, const double* pars, int32_t n_cols, const double* gram, double* deltap);
int32_t mole_opt_sr_matrix_gram(mole_opt_t opt, int32_t n_cols, const double* gram, double* S);
/* the regularised SR matrix the solve uses (P*P row-major), for parity tests */
int32_t mole_opt_sr_matrix(mole_opt_t opt, const mole_acc_host* acc, double* S);

/* ---- drivers ------------------------------------------------------------------------------------ */
/* Runner::run (montecarlo.rs:24-46): steps sweeps in blocks of block_size, block 0 discarded.
 * Accumulators are NOT reset.  Trace pointers as in mole_sweep_args (may be NULL). */
int32_t mole_runner_run(mole_ens_t ens, mole_wf_t wf, mole_metrop_t m, mole_op_t op, uint32_t observables,
                        uint32_t compat, int32_t steps, int32_t block_size, double* energy_trace,
                        double* wfvalue_trace, double* kinetic_trace, double* pgrad_trace, uint8_t* accept_trace);

/* VmcRunner::run_optimization (vmc.rs:43-106).  n_walkers of `ens` plays the role of nworkers:
 * every walker runs total_samples/nworkers sweeps; flags bit0: restart every iteration from the
 * ensemble's configurations at entry (reference behaviour, vmc.rs:56); 0 = carry walkers over.
 * Outputs: energies[iters], errors[iters], acceptance[iters] (nullable), param_history[iters*P]. */
enum { MOLE_VMC_RESTART_EACH_ITER = 1 };
int32_t mole_vmc_run_optimization(mole_ens_t ens, mole_wf_t wf, mole_metrop_t m, mole_op_t op, mole_opt_t opt,
                                  const uint8_t master_seed[32], int32_t iters, int64_t total_samples,
                                  int32_t block_size, uint32_t compat, uint32_t flags, double* energies,
                                  double* errors, double* acceptance, double* param_history);

/* ---- DMC (src/dmc/src/dmc.rs, branching.rs) ------------------------------------------------------ */
enum { MOLE_BRANCH_SR = 0 /* SRBrancher branching.rs:7-40 */, MOLE_BRANCH_SIMPLE = 1 /* SimpleBranching :42-92 */ };
/* one time step of DmcRunner::diffuse's walker loop (dmc.rs:87-130) for all walkers; returns the
 * LOCAL (this rank) sums: sum_w_e = sum w_i E_L,i(old), sum_w = sum w_i (pre-update weights). */
int32_t mole_dmc_step(mole_ens_t ens, mole_wf_t wf, mole_metrop_t m, mole_op_t op, double time_step,
                      double reference_energy, double* sum_w_e, double* sum_w);
/* BranchingAlgorithm::branch (dmc/src/traits.rs:4-7) on the device (scan + search + gather).
 * Branching is the last operation of a time step (dmc.rs:139-140): it draws with the current step
 * counter and then advances it by one. */
int32_t mole_branch(mole_ens_t ens, int32_t kind);
/* Cross-rank population rebalancing: the ranks' islands (see mole_dmc_block) become one equal-weight population
 * again.  Rank r's walkers fill a share of the N_total slots proportional to the rank's total weight; surplus copies
 * migrate in one grouped ncclSend / ncclRecv exchange (configuration + cached E_L); every walker ends with the global
 * mean weight; walker counts per rank are unchanged.  Call it between blocks.  mole_rebalance_plan is the host
 * arithmetic alone (shares[r], moves[src][dst]) for a given shared draw u in [0, 1). */
int32_t mole_rebalance(mole_ens_t ens);
/* largest / smallest island weight per walker at the last step of the last SRBrancher block (1.0 on one rank; the same
 * value on every rank, no extra collective beyond the walker counts): mole_dmc_diffuse rebalances between blocks when
 * it exceeds MOLE_REBALANCE_RATIO (measured: two islands of 2^15 H-atom walkers drift 3 % apart in 2800 time steps) */
#define MOLE_REBALANCE_RATIO 1.05
int32_t mole_dmc_island_imbalance(mole_ens_t ens, double* ratio);
int32_t mole_rebalance_plan(int32_t nranks, const double* totals, const int64_t* counts, double u, int64_t* shares,
                            int64_t* moves /* nranks x nranks, nullable */);
/* source walker index of every walker after the last mole_branch (parity tests) */
int32_t mole_branch_sources(mole_ens_t ens, int32_t* src);
/* One block of DmcRunner::diffuse's inner loop (dmc.rs:84-141): n_steps x (time step, ensemble energy
 * sum w E / sum w over ALL ranks, branch).  With SRBrancher the block is enqueued without host reads
 * (the reference energy is constant within a block, dmc.rs:143-145).  Thread-per-walker kinds whose population
 * fits a co-resident grid at one walker per thread run the whole block as ONE cooperative launch, the
 * three phases of a step (time step + reduction, weight scan, pick + gather) separated by two grid barriers; other
 * cases enqueue 3 launches per time step.  Either way the branching normalisation is formed on the device and the
 * per-step {sum w E, sum w} rows come back with one copy, and the results are bit-identical.  Multi-rank: every rank is a population island (SRBrancher resamples against the rank's own N, w_max
 * and mean weight - stratified, unbiased), so no collective sits in the step loop; the rows of all ranks are
 * all-gathered once per block and every rank forms the same step_energies[n_steps].  Identical results to
 * n_steps x (mole_dmc_step, mole_branch). */
int32_t mole_dmc_block(mole_ens_t ens, mole_wf_t wf, mole_metrop_t m, mole_op_t op, int32_t branch_kind,
                       double time_step, double reference_energy, int32_t n_steps, double* step_energies);
/* impl 0 (default): one persistent launch per block when the population fits the co-resident grid at one walker per
 * thread (about 75 000 one-electron walkers on 148 SMs; faster there, slower beyond), per-step launches otherwise;
 * 1: always per-step launches; 2: the persistent launch for every population it can serve (tests, A/B) */
int32_t mole_dmc_block_select(mole_ens_t ens, int32_t impl);
/* DmcRunner::diffuse (dmc.rs:69-153): returns n_out running energies and errors */
int32_t mole_dmc_diffuse(mole_ens_t ens, mole_wf_t wf, mole_metrop_t m, mole_op_t op, int32_t branch_kind,
                         double time_step, double* reference_energy /* in/out */, int32_t num_iterations,
                         int32_t block_size, int32_t num_eq_blocks, double* energies, double* errors,
                         int32_t* n_out, double* step_energies /* nullable, num_iterations */);

/* ---- series statistics on the device ---------------------------------------------
 * Port of the reference's offline analysis tool scripts/statfor.rs (mean :17-19, variance :23-26,
 * correlation :31-54, blocking :58-81; scripts/statfor.py:21-40 is the same correlation()) applied to
 * every walker's local-energy series without the series leaving the GPU.  A sweep called with
 * MOLE_SWEEP_KEEP_SERIES stores its E_L samples in a device buffer owned by the ensemble
 * (MOLE_SWEEP_APPEND_SERIES appends to it); mole_series_analyze then runs statfor per walker and
 * returns walker-averaged results (and, optionally, the per-walker ones). */
enum { MOLE_SWEEP_KEEP_SERIES = 1, MOLE_SWEEP_APPEND_SERIES = 2 };
#define MOLE_SERIES_MAX_LAG 200 /* MAX_STEPS, statfor.rs:32 */
typedef struct {
  double average;  /* statfor.rs:17-19              */
  double variance; /* ddof = 1, statfor.rs:23-26,87 */
  double tcorr;    /* statfor.rs:35-52              */
  double n_eff;    /* nsteps / tcorr, :53           */
  double sigma;    /* sqrt(variance*tcorr/nsteps)   */
} mole_series_stats;
/* number of samples per walker currently held */
int32_t mole_series_length(mole_ens_t ens, int64_t* n);
int32_t mole_series_clear(mole_ens_t ens);
/* block-size schedule of statfor.rs:59-66 for a series of n samples: 1, 1+step, ... <= n/20.
 * sizes may be NULL to query the count. */
int32_t mole_series_block_sizes(int64_t n, int32_t* sizes, int32_t* n_sizes);
/* statfor per walker.  drop_last != 0 analyses data[..len-1] like statfor.rs:85.
 * mean_stats: the five quantities averaged over walkers; corr[lags]: walker-mean autocorrelation,
 * lags = min(MOLE_SERIES_MAX_LAG, n-1) (nullable); block_errors[n_sizes]: walker-mean blocking error
 * for the given block sizes (nullable); per_walker[W], per_walker_corr[lags*W] (lag-major) and
 * per_walker_block_errors[n_sizes*W]: host buffers for the unreduced results (nullable; parity tests). */
int32_t mole_series_analyze(mole_ens_t ens, int32_t drop_last, mole_series_stats* mean_stats, double* corr,
                            int32_t n_sizes, const int32_t* block_sizes, double* block_errors,
                            mole_series_stats* per_walker, double* per_walker_corr, double* per_walker_block_errors);
/* one walker's series to the host, or as the text format scripts/statfor.py:17-19 and statfor.rs:7-13
 * read (one float per line, 17 significant digits) */
int32_t mole_series_get(mole_ens_t ens, int64_t walker, double* out);
int32_t mole_series_write_text(mole_ens_t ens, int64_t walker, const char* path);

/* ---- Log (montecarlo/src/traits.rs:44-47) ----------------------------------------------------------
 * Runner::run calls logger.log(data) after every sample and prints its output once per block
 * (montecarlo.rs:31-43).  The ensemble equivalent is a callback per block fed with block-level
 * reductions over all walkers of this rank; one launch per block. */
typedef struct {
  int32_t block_nr;       /* 1.. (block 0 is the discarded equilibration block) */
  int32_t block_size;
  double n_samples;       /* samples in this block, all walkers                 */
  double block_energy;    /* mean E_L over this block                           */
  double running_energy;  /* mean E_L over all sampled blocks so far            */
  double block_kinetic;   /* mean -0.5 lap/psi (MOLE_OBS_KINETIC)               */
  double block_wfvalue;   /* mean psi (MOLE_OBS_WFVALUE)                        */
  double acceptance;      /* accepted / proposed single-electron moves, block   */
} mole_block_log;
typedef void (*mole_log_fn)(void* user, const mole_block_log* data);
int32_t mole_runner_run_logged(mole_ens_t ens, mole_wf_t wf, mole_metrop_t m, mole_op_t op, uint32_t observables,
                               uint32_t compat, int32_t steps, int32_t block_size, uint32_t sweep_flags,
                               mole_log_fn log, void* user);

/* ---- checkpoint / restart ----------------------------------------------------------
 * Binary dump of the ensemble: configurations, weights, block partial sums, accumulators, cached E_L,
 * step counter and Philox key.  The random streams are counter-based, so a restored ensemble continues
 * bit-identically. */
int32_t mole_ensemble_save(mole_ens_t ens, const char* path);
int32_t mole_ensemble_load(mole_ens_t ens, const char* path); /* W and N_e must match */

/* ---- measurement helpers ---------------------------------------------------------------------- */
/* sustained DFMA throughput of the device (TFLOP/s) measured with a register-resident FMA chain */
int32_t mole_bench_fp64_peak(mole_ctx_t ctx, double* tflops);
/* fp64 tensor-core (mma.m8n8k4, DMMA) throughput with `chains` (1,2,4,8,16) independent accumulators per warp and
 * `warps_per_sm` (1..32) resident warps: the compute ceiling of the Gram contraction */
int32_t mole_bench_dmma_peak(mole_ctx_t ctx, int32_t chains, int32_t warps_per_sm, double* tflops);
/* evaluates the kernels' branch-free fp64 elementary functions on the device (accuracy tests):
 * which = 0 exp, 1 reciprocal, 2 reciprocal square root, 3 square root, 4 natural log,
 * 5 / 6 sine / cosine of a fraction of a full turn */
int32_t mole_math_probe(mole_ctx_t ctx, int32_t which, const double* in, int64_t n, double* out);
/* the Gram contraction of the large-P path timed alone on synthetic rows (W walkers x n_samples x cols columns),
 * impl 0: DMMA, 1: FP64 vector pipe; mean milliseconds of `reps` launches (CUDA events); checksum: sum of the upper
 * triangle, to compare the two implementations.  Algorithmic bytes = 8 W n_samples cols (every row read once). */
int32_t mole_bench_gram(mole_ctx_t ctx, int64_t n_walkers, int64_t n_samples, int32_t cols, int32_t impl, int32_t reps,
                        double* ms, double* checksum);
/* number of kernels this library has launched on ctx since creation */
int32_t mole_ctx_launch_count(mole_ctx_t ctx, int64_t* n);

#ifdef __cplusplus
}
#endif
#endif /* MOLE_B200_H */
