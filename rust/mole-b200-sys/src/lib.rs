//! Raw FFI binding of include/mole_b200.h (generated from the header by rust/gen in this repo's history;
//! one `pub fn` per exported symbol).  NOT COMPILED in this image (no rustc) - see rust/README.md.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_void};

#[repr(C)] pub struct mole_ctx_s { _private: [u8; 0] }
#[repr(C)] pub struct mole_wf_s { _private: [u8; 0] }
#[repr(C)] pub struct mole_op_s { _private: [u8; 0] }
#[repr(C)] pub struct mole_ens_s { _private: [u8; 0] }
#[repr(C)] pub struct mole_metrop_s { _private: [u8; 0] }
#[repr(C)] pub struct mole_opt_s { _private: [u8; 0] }

pub const MOLE_OK: i32 = 0;
pub const MOLE_ERR_LINALG: i32 = 1;
pub const MOLE_ERR_SHAPE: i32 = 2;
pub const MOLE_ERR_FUNC: i32 = 3;
pub const MOLE_ERR_OPERATOR_VALUE_ACCESS: i32 = 4;
pub const MOLE_ERR_DATA_ACCESS: i32 = 5;
pub const MOLE_ERR_EMPTY_CACHE: i32 = 6;
pub const MOLE_ERR_CUDA: i32 = 100;
pub const MOLE_ERR_NCCL: i32 = 101;
pub const MOLE_ERR_INVALID_ARG: i32 = 102;
pub const MOLE_ERR_NO_DEVICE: i32 = 103;
pub const MOLE_ERR_ASSERT: i32 = 104;

#[repr(C)] #[derive(Clone, Copy, Default)] pub struct mole_ens_health { pub nonfinite_samples: i64, pub nonfinite_dmc_walkers: i64, pub reserved: [i64; 2] }
#[repr(C)] #[derive(Clone, Copy)] pub struct mole_wf_desc { pub kind: i32, pub n_elec: i32, pub n_params: i32, pub reserved: i32, pub params: [f64; 48], pub geom: [f64; 40] }   // MOLE_WF_MAX_PARAMS, MOLE_WF_MAX_GEOM
#[repr(C)] #[derive(Clone, Copy)] pub struct mole_op_desc { pub kind: i32, pub n_ions: i32, pub ion_pos: [f64; 24], pub ion_charge: [i32; 8], pub frequency: f64 }
#[repr(C)] pub struct mole_sweep_args { pub n_sweeps: i32, pub n_discard: i32, pub block_size: i32, pub observables: u32, pub compat: u32, pub flags: u32,
    pub energy_trace: *mut f64, pub wfvalue_trace: *mut f64, pub kinetic_trace: *mut f64, pub pgrad_trace: *mut f64, pub accept_trace: *mut u8 }
pub const MOLE_SWEEP_KEEP_SERIES: u32 = 1;
pub const MOLE_SWEEP_APPEND_SERIES: u32 = 2;
pub const MOLE_SERIES_MAX_LAG: usize = 200;
#[repr(C)] #[derive(Clone, Copy, Default)] pub struct mole_series_stats { pub average: f64, pub variance: f64, pub tcorr: f64, pub n_eff: f64, pub sigma: f64 }
#[repr(C)] #[derive(Clone, Copy)] pub struct mole_block_log { pub block_nr: i32, pub block_size: i32, pub n_samples: f64, pub block_energy: f64,
    pub running_energy: f64, pub block_kinetic: f64, pub block_wfvalue: f64, pub acceptance: f64 }
pub type mole_log_fn = Option<unsafe extern "C" fn(user: *mut core::ffi::c_void, data: *const mole_block_log)>;
#[repr(C)] #[derive(Clone, Copy)] pub struct mole_acc_host { pub n_samples: f64, pub sum_e: f64, pub sum_e2: f64, pub sum_b: f64, pub sum_b2: f64, pub n_blocks: f64,
    pub n_accept: f64, pub n_moves: f64, pub sum_t: f64, pub sum_psi: f64, pub sum_o: [f64; 8], pub sum_oe: [f64; 8], pub sum_oo: [f64; 36], pub n_params: i32, pub reserved: i32 }

extern "C" {
    pub fn mole_ctx_create(device: i32, ctx: *mut *mut mole_ctx_s) -> i32;
    pub fn mole_ctx_destroy(ctx: *mut mole_ctx_s) -> i32;
    pub fn mole_ctx_synchronize(ctx: *mut mole_ctx_s) -> i32;
    pub fn mole_last_error_string(ctx: *mut mole_ctx_s) -> *const c_char;
    pub fn mole_ctx_stream(ctx: *mut mole_ctx_s, stream: *mut *mut c_void) -> i32;
    pub fn mole_version() -> i32;
    pub fn mole_wf_create(ctx: *mut mole_ctx_s, desc: *const mole_wf_desc, wf: *mut *mut mole_wf_s) -> i32;
    pub fn mole_wf_destroy(wf: *mut mole_wf_s) -> i32;
    pub fn mole_wf_num_electrons(wf: *mut mole_wf_s, n: *mut i32) -> i32;
    pub fn mole_wf_num_parameters(wf: *mut mole_wf_s, n: *mut i32) -> i32;
    pub fn mole_wf_get_parameters(wf: *mut mole_wf_s, params: *mut f64) -> i32;
    pub fn mole_wf_update_parameters(wf: *mut mole_wf_s, deltap: *const f64) -> i32;
    pub fn mole_wf_set_parameters(wf: *mut mole_wf_s, params: *const f64) -> i32;
    pub fn mole_wf_value(wf: *mut mole_wf_s, cfg: *const f64, out: *mut f64) -> i32;
    pub fn mole_wf_gradient(wf: *mut mole_wf_s, cfg: *const f64, out: *mut f64) -> i32;
    pub fn mole_wf_laplacian(wf: *mut mole_wf_s, cfg: *const f64, out: *mut f64) -> i32;
    pub fn mole_wf_parameter_gradient(wf: *mut mole_wf_s, cfg: *const f64, out: *mut f64) -> i32;
    pub fn mole_op_create(ctx: *mut mole_ctx_s, desc: *const mole_op_desc, op: *mut *mut mole_op_s) -> i32;
    pub fn mole_op_destroy(op: *mut mole_op_s) -> i32;
    pub fn mole_op_act_on(op: *mut mole_op_s, wf: *mut mole_wf_s, cfg: *const f64, out: *mut f64) -> i32;
    pub fn mole_ensemble_create(ctx: *mut mole_ctx_s, n_walkers: i64, n_elec: i32, seed: *const u8, walker_offset: u64, ens: *mut *mut mole_ens_s) -> i32;
    pub fn mole_ensemble_destroy(ens: *mut mole_ens_s) -> i32;
    pub fn mole_ensemble_num_walkers(ens: *mut mole_ens_s, n: *mut i64) -> i32;
    pub fn mole_ensemble_init_uniform(ens: *mut mole_ens_s, lo: f64, hi: f64, broadcast_walker0: i32) -> i32;
    pub fn mole_ensemble_init_normal(ens: *mut mole_ens_s, sigma: f64, broadcast_walker0: i32) -> i32;
    pub fn mole_ensemble_set_configs(ens: *mut mole_ens_s, cfgs: *const f64) -> i32;
    pub fn mole_ensemble_set_configs_broadcast(ens: *mut mole_ens_s, cfg: *const f64) -> i32;
    pub fn mole_ensemble_get_configs(ens: *mut mole_ens_s, cfgs: *mut f64) -> i32;
    pub fn mole_ensemble_set_weights(ens: *mut mole_ens_s, w: *const f64) -> i32;
    pub fn mole_ensemble_get_weights(ens: *mut mole_ens_s, w: *mut f64) -> i32;
    pub fn mole_ensemble_snapshot(ens: *mut mole_ens_s) -> i32;
    pub fn mole_ensemble_restore(ens: *mut mole_ens_s) -> i32;
    pub fn mole_ensemble_reseed(ens: *mut mole_ens_s, seed: *const u8) -> i32;
    pub fn mole_ensemble_set_step(ens: *mut mole_ens_s, step: u32) -> i32;
    pub fn mole_ensemble_get_step(ens: *mut mole_ens_s, step: *mut u32) -> i32;
    pub fn mole_derive_seed(master: *const u8, n: u32, out: *mut u8) -> i32;
    pub fn mole_eval_vgl(ens: *mut mole_ens_s, wf: *mut mole_wf_s, op: *mut mole_op_s, psi: *mut f64, grad: *mut f64, lap: *mut f64, hpsi: *mut f64, pgrad: *mut f64) -> i32;
    pub fn mole_metropolis_create(kind: i32, param: f64, m: *mut *mut mole_metrop_s) -> i32;
    pub fn mole_metropolis_destroy(m: *mut mole_metrop_s) -> i32;
    pub fn mole_metropolis_set_compat(m: *mut mole_metrop_s, compat: u32) -> i32;
    pub fn mole_sweep(ens: *mut mole_ens_s, wf: *mut mole_wf_s, m: *mut mole_metrop_s, op: *mut mole_op_s, args: *const mole_sweep_args) -> i32;
    pub fn mole_acc_reset(ens: *mut mole_ens_s) -> i32;
    pub fn mole_acc_get(ens: *mut mole_ens_s, out: *mut mole_acc_host) -> i32;
    pub fn mole_acc_allreduce(ens: *mut mole_ens_s) -> i32;
    pub fn mole_acc_device_ptr(ens: *mut mole_ens_s, ptr_dev: *mut *mut c_void, n_doubles: *mut i32) -> i32;
    pub fn mole_acc_finalize(acc: *const mole_acc_host, energy: *mut f64, error: *mut f64, acceptance_per_sweep: *mut f64, grad: *mut f64) -> i32;
    pub fn mole_comm_get_unique_id(id: *mut u8) -> i32;
    pub fn mole_comm_init(ctx: *mut mole_ctx_s, nranks: i32, rank: i32, id: *const u8) -> i32;
    pub fn mole_comm_destroy(ctx: *mut mole_ctx_s) -> i32;
    pub fn mole_opt_create(kind: i32, n_params: i32, step_size: f64, momentum_parameter: f64, history: i32, compat: u32, opt: *mut *mut mole_opt_s) -> i32;
    pub fn mole_opt_destroy(opt: *mut mole_opt_s) -> i32;
    pub fn mole_opt_step(opt: *mut mole_opt_s, pars: *const f64, acc: *const mole_acc_host, deltap: *mut f64) -> i32;
    pub fn mole_opt_sr_matrix(opt: *mut mole_opt_s, acc: *const mole_acc_host, S: *mut f64) -> i32;
    pub fn mole_runner_run(ens: *mut mole_ens_s, wf: *mut mole_wf_s, m: *mut mole_metrop_s, op: *mut mole_op_s, observables: u32, compat: u32, steps: i32, block_size: i32, energy_trace: *mut f64, wfvalue_trace: *mut f64, kinetic_trace: *mut f64, pgrad_trace: *mut f64, accept_trace: *mut u8) -> i32;
    pub fn mole_vmc_run_optimization(ens: *mut mole_ens_s, wf: *mut mole_wf_s, m: *mut mole_metrop_s, op: *mut mole_op_s, opt: *mut mole_opt_s, master_seed: *const u8, iters: i32, total_samples: i64, block_size: i32, compat: u32, flags: u32, energies: *mut f64, errors: *mut f64, acceptance: *mut f64, param_history: *mut f64) -> i32;
    pub fn mole_dmc_step(ens: *mut mole_ens_s, wf: *mut mole_wf_s, m: *mut mole_metrop_s, op: *mut mole_op_s, time_step: f64, reference_energy: f64, sum_w_e: *mut f64, sum_w: *mut f64) -> i32;
    pub fn mole_branch(ens: *mut mole_ens_s, kind: i32) -> i32;
    pub fn mole_branch_sources(ens: *mut mole_ens_s, src: *mut i32) -> i32;
    pub fn mole_dmc_diffuse(ens: *mut mole_ens_s, wf: *mut mole_wf_s, m: *mut mole_metrop_s, op: *mut mole_op_s, branch_kind: i32, time_step: f64, reference_energy: *mut f64, num_iterations: i32, block_size: i32, num_eq_blocks: i32, energies: *mut f64, errors: *mut f64, n_out: *mut i32, step_energies: *mut f64) -> i32;
    pub fn mole_bench_fp64_peak(ctx: *mut mole_ctx_s, tflops: *mut f64) -> i32;
    pub fn mole_math_probe(ctx: *mut mole_ctx_s, which: i32, in_: *const f64, n: i64, out: *mut f64) -> i32;
    pub fn mole_ctx_launch_count(ctx: *mut mole_ctx_s, n: *mut i64) -> i32;
    // series statistics (scripts/statfor.rs), Log callback (montecarlo/src/traits.rs:44-47), checkpoint / restart
    pub fn mole_series_length(ens: *mut mole_ens_s, n: *mut i64) -> i32;
    pub fn mole_series_clear(ens: *mut mole_ens_s) -> i32;
    pub fn mole_series_block_sizes(n: i64, sizes: *mut i32, n_sizes: *mut i32) -> i32;
    pub fn mole_series_analyze(ens: *mut mole_ens_s, drop_last: i32, mean_stats: *mut mole_series_stats, corr: *mut f64, n_sizes: i32, block_sizes: *const i32, block_errors: *mut f64, per_walker: *mut mole_series_stats, per_walker_corr: *mut f64, per_walker_block_errors: *mut f64) -> i32;
    pub fn mole_series_get(ens: *mut mole_ens_s, walker: i64, out: *mut f64) -> i32;
    pub fn mole_series_write_text(ens: *mut mole_ens_s, walker: i64, path: *const core::ffi::c_char) -> i32;
    pub fn mole_runner_run_logged(ens: *mut mole_ens_s, wf: *mut mole_wf_s, m: *mut mole_metrop_s, op: *mut mole_op_s, observables: u32, compat: u32, steps: i32, block_size: i32, sweep_flags: u32, log: mole_log_fn, user: *mut core::ffi::c_void) -> i32;
    pub fn mole_ensemble_save(ens: *mut mole_ens_s, path: *const core::ffi::c_char) -> i32;
    pub fn mole_ensemble_load(ens: *mut mole_ens_s, path: *const core::ffi::c_char) -> i32;
    // second round: health counters, SR regularisation, large-P Gram path, DMC block selection, rebalancing, probes
    pub fn mole_ensemble_health(ens: *mut mole_ens_s, out: *mut mole_ens_health) -> i32;
    pub fn mole_opt_set_sr_regularization(opt: *mut mole_opt_s, diag_scale: f64, diag_shift: f64) -> i32;
    pub fn mole_gram_get(ens: *mut mole_ens_s, n_cols: *mut i32, gram: *mut f64) -> i32;
    pub fn mole_gram_allreduce(ens: *mut mole_ens_s) -> i32;
    pub fn mole_gram_device_ptr(ens: *mut mole_ens_s, ptr_dev: *mut *mut core::ffi::c_void, n_doubles: *mut i32) -> i32;
    pub fn mole_gram_select(ens: *mut mole_ens_s, impl_: i32) -> i32;
    pub fn mole_gram_finalize(n_cols: i32, gram: *const f64, energy: *mut f64, grad: *mut f64) -> i32;
    pub fn mole_opt_step_gram(opt: *mut mole_opt_s, pars: *const f64, n_cols: i32, gram: *const f64, deltap: *mut f64) -> i32;
    pub fn mole_opt_sr_matrix_gram(opt: *mut mole_opt_s, n_cols: i32, gram: *const f64, s: *mut f64) -> i32;
    pub fn mole_dmc_block_select(ens: *mut mole_ens_s, impl_: i32) -> i32;
    pub fn mole_rebalance(ens: *mut mole_ens_s) -> i32;
    pub fn mole_dmc_island_imbalance(ens: *mut mole_ens_s, ratio: *mut f64) -> i32;
    pub fn mole_rebalance_plan(nranks: i32, totals: *const f64, counts: *const i64, u: f64, shares: *mut i64, moves: *mut i64) -> i32;
    pub fn mole_bench_gram(ctx: *mut mole_ctx_s, n_walkers: i64, n_samples: i64, cols: i32, impl_: i32, reps: i32, ms: *mut f64, checksum: *mut f64) -> i32;
    pub fn mole_bench_dmma_peak(ctx: *mut mole_ctx_s, chains: i32, warps_per_sm: i32, tflops: *mut f64) -> i32;
}

// enumerations of include/mole_b200.h
pub const MOLE_METROP_BOX: i32 = 0;
pub const MOLE_METROP_DIFFUSE: i32 = 1;
pub const MOLE_OBS_ENERGY: u32 = 1;
pub const MOLE_OBS_PGRAD: u32 = 2;
pub const MOLE_OBS_WFVALUE: u32 = 4;
pub const MOLE_OBS_KINETIC: u32 = 8;
pub const MOLE_OPT_SD: i32 = 0;
pub const MOLE_OPT_MOMENTUM: i32 = 1;
pub const MOLE_OPT_NESTEROV: i32 = 2;
pub const MOLE_OPT_LBFGS: i32 = 3;
pub const MOLE_OPT_SR: i32 = 4;
pub const MOLE_BRANCH_SR: i32 = 0;
pub const MOLE_BRANCH_SIMPLE: i32 = 1;
pub const MOLE_VMC_RESTART_EACH_ITER: u32 = 1;
extern "C" {
    pub fn mole_dmc_block(ens: *mut mole_ens_s, wf: *mut mole_wf_s, m: *mut mole_metrop_s, op: *mut mole_op_s, branch_kind: i32, time_step: f64, reference_energy: f64, n_steps: i32, step_energies: *mut f64) -> i32;
}
