//! Samplers and drivers over the C ABI: Metropolis kernels, Sampler/Runner with a `Log`, optimizers,
//! DmcRunner, series statistics and checkpoints.
//! SOURCES ONLY - never compiled in this image (no rustc); see rust/README.md and INTEGRATION.md.
use crate::{check, Context, GpuHamiltonian, GpuWaveFunction, Result};
use mole_b200_sys as sys;
use ndarray::Array1;
use std::ffi::{c_void, CString};
use std::ptr;

/// `MetropolisBox::{new, from_rng}` / `MetropolisDiffuse::{new, from_rng, fix_nodes}`
/// (src/metropolis/src/metrop.rs:34-45,113-135).  The 32-byte seed keys the Philox streams
/// (`Metropolis::reseed_rng`, `generate_seed`: src/metropolis/src/traits.rs:33-37).
pub struct GpuMetropolis { pub(crate) raw: *mut sys::mole_metrop_s, pub seed: [u8; 32], generated: u32 }
impl GpuMetropolis {
    fn create(kind: i32, param: f64, seed: [u8; 32]) -> Result<Self> {
        let mut raw = ptr::null_mut();
        check(unsafe { sys::mole_metropolis_create(kind, param, &mut raw) })?;
        Ok(Self { raw, seed, generated: 0 })
    }
    pub fn box_from_seed(box_side: f64, seed: [u8; 32]) -> Result<Self> { Self::create(sys::MOLE_METROP_BOX, box_side, seed) }
    pub fn diffuse_from_seed(time_step: f64, seed: [u8; 32]) -> Result<Self> { Self::create(sys::MOLE_METROP_DIFFUSE, time_step, seed) }
    /// metrop.rs:128-135; the node test is unconditional upstream (`fixed_node` is never read)
    pub fn fix_nodes(self) -> Self { self }
    pub fn reseed_rng(&mut self, seed: [u8; 32]) { self.seed = seed; self.generated = 0; }
    pub fn generate_seed(&mut self) -> [u8; 32] {
        let mut out = [0u8; 32];
        unsafe { sys::mole_derive_seed(self.seed.as_ptr(), self.generated, out.as_mut_ptr()); }
        self.generated += 1;
        out
    }
}
impl Drop for GpuMetropolis { fn drop(&mut self) { unsafe { sys::mole_metropolis_destroy(self.raw); } } }

/// Block-level view handed to a logger (the ensemble form of `Log::log`'s data map,
/// src/montecarlo/src/traits.rs:44-47).
pub type BlockLog = sys::mole_block_log;
pub trait GpuLog { fn log(&mut self, data: &BlockLog) -> String; }
pub struct EmptyLogger;                                                      // src/vmc/src/vmc.rs:14-19
impl GpuLog for EmptyLogger { fn log(&mut self, _d: &BlockLog) -> String { String::new() } }

unsafe extern "C" fn log_trampoline<L: GpuLog>(user: *mut c_void, data: *const BlockLog) {
    let logger = &mut *(user as *mut L);
    let out = logger.log(&*data);
    if !out.is_empty() { println!("{}", out); }                              // montecarlo.rs:40-42
}

/// `Sampler::{new, with_initial_configuration}` (src/montecarlo/src/samplers.rs:40-71) for an ensemble of
/// `n_walkers` independent chains, and `Runner::{new, run}` (src/montecarlo/src/montecarlo.rs:20-46).
pub struct GpuSampler { pub(crate) ens: *mut sys::mole_ens_s, pub metrop: GpuMetropolis, pub observables: u32, n_elec: usize }
impl GpuSampler {
    pub fn new(ctx: &Context, wf: &GpuWaveFunction, metrop: GpuMetropolis, observables: u32, n_walkers: usize) -> Result<Self> {
        use wavefunction_traits::WaveFunction;
        let mut ens = ptr::null_mut();
        let ne = wf.num_electrons();
        check(unsafe { sys::mole_ensemble_create(ctx.raw, n_walkers as i64, ne as i32, metrop.seed.as_ptr(), 0, &mut ens) })?;
        check(unsafe { sys::mole_ensemble_init_uniform(ens, -1.0, 1.0, 0) })?;         // samplers.rs:45-47
        Ok(Self { ens, metrop, observables, n_elec: ne })
    }
    /// Runner::run: `steps` sweeps in blocks of `block_size`, block 0 discarded; returns
    /// (mean energy, blocking error, acceptance) from the device-reduced moments.
    pub fn run<L: GpuLog>(&mut self, wf: &GpuWaveFunction, h: &GpuHamiltonian, steps: usize, block_size: usize,
                          logger: &mut L) -> Result<(f64, f64, f64)> {
        check(unsafe { sys::mole_acc_reset(self.ens) })?;
        check(unsafe { sys::mole_runner_run_logged(self.ens, wf.raw(), self.metrop.raw, h.raw, self.observables, 0,
                                                   steps as i32, block_size as i32, 0, Some(log_trampoline::<L>),
                                                   logger as *mut L as *mut c_void) })?;
        let mut acc: sys::mole_acc_host = unsafe { std::mem::zeroed() };
        check(unsafe { sys::mole_acc_get(self.ens, &mut acc) })?;
        let (mut e, mut err, mut a) = (0.0, 0.0, 0.0);
        check(unsafe { sys::mole_acc_finalize(&acc, &mut e, &mut err, &mut a, ptr::null_mut()) })?;
        Ok((e, err, a))
    }
    /// scripts/statfor.rs per walker on the device-resident E_L series (keep it with
    /// `MOLE_SWEEP_KEEP_SERIES`); returns the walker means and the autocorrelation function.
    pub fn series_analyze(&self) -> Result<(sys::mole_series_stats, Vec<f64>)> {
        let mut n = 0i64;
        check(unsafe { sys::mole_series_length(self.ens, &mut n) })?;
        let lags = std::cmp::min(sys::MOLE_SERIES_MAX_LAG as i64, n - 1).max(0) as usize;
        let mut st = sys::mole_series_stats::default();
        let mut corr = vec![0.0; lags];
        check(unsafe { sys::mole_series_analyze(self.ens, 0, &mut st, corr.as_mut_ptr(), 0, ptr::null(), ptr::null_mut(),
                                                ptr::null_mut(), ptr::null_mut(), ptr::null_mut()) })?;
        Ok((st, corr))
    }
    pub fn save(&self, path: &str) -> Result<()> {
        let p = CString::new(path).expect("path");
        check(unsafe { sys::mole_ensemble_save(self.ens, p.as_ptr()) })
    }
    pub fn load(&mut self, path: &str) -> Result<()> {
        let p = CString::new(path).expect("path");
        check(unsafe { sys::mole_ensemble_load(self.ens, p.as_ptr()) })
    }
    pub fn num_electrons(&self) -> usize { self.n_elec }
}
impl Drop for GpuSampler { fn drop(&mut self) { unsafe { sys::mole_ensemble_destroy(self.ens); } } }

/// The five optimizers of src/optimize/src/optimizers.rs:9-253 consuming reduced moments.
pub struct GpuOptimizer { pub(crate) raw: *mut sys::mole_opt_s }
impl GpuOptimizer {
    fn create(kind: i32, n_params: usize, step: f64, momentum: f64, history: usize) -> Result<Self> {
        let mut raw = ptr::null_mut();
        check(unsafe { sys::mole_opt_create(kind, n_params as i32, step, momentum, history as i32, 0, &mut raw) })?;
        Ok(Self { raw })
    }
    pub fn steepest_descent(n: usize, step: f64) -> Result<Self> { Self::create(sys::MOLE_OPT_SD, n, step, 0.0, 0) }
    pub fn momentum_descent(n: usize, step: f64, mu: f64) -> Result<Self> { Self::create(sys::MOLE_OPT_MOMENTUM, n, step, mu, 0) }
    pub fn nesterov_momentum(n: usize, step: f64, mu: f64) -> Result<Self> { Self::create(sys::MOLE_OPT_NESTEROV, n, step, mu, 0) }
    pub fn online_lbfgs(n: usize, step: f64, history: usize) -> Result<Self> { Self::create(sys::MOLE_OPT_LBFGS, n, step, 0.0, history) }
    pub fn stochastic_reconfiguration(n: usize, step: f64) -> Result<Self> { Self::create(sys::MOLE_OPT_SR, n, step, 0.0, 0) }
}
impl Drop for GpuOptimizer { fn drop(&mut self) { unsafe { sys::mole_opt_destroy(self.raw); } } }

/// `DmcRunner::{new, diffuse}` (src/dmc/src/dmc.rs:40-153) with `SRBrancher` / `SimpleBranching`
/// (src/dmc/src/branching.rs:7-92).
pub struct GpuDmcRunner { ens: *mut sys::mole_ens_s, metrop: GpuMetropolis, branch_kind: i32, pub reference_energy: f64 }
impl GpuDmcRunner {
    pub fn new(ctx: &Context, wf: &GpuWaveFunction, num_walkers: usize, reference_energy: f64, metrop: GpuMetropolis,
               sr_brancher: bool) -> Result<Self> {
        use wavefunction_traits::WaveFunction;
        let mut ens = ptr::null_mut();
        check(unsafe { sys::mole_ensemble_create(ctx.raw, num_walkers as i64, wf.num_electrons() as i32, metrop.seed.as_ptr(), 0, &mut ens) })?;
        check(unsafe { sys::mole_ensemble_init_normal(ens, 1.0, 1) })?;     // dmc.rs:49-58: vec![(1.0, cfg); n]
        Ok(Self { ens, metrop, branch_kind: if sr_brancher { sys::MOLE_BRANCH_SR } else { sys::MOLE_BRANCH_SIMPLE }, reference_energy })
    }
    pub fn diffuse(&mut self, wf: &GpuWaveFunction, h: &GpuHamiltonian, time_step: f64, num_iterations: usize,
                   block_size: usize, num_eq_blocks: usize) -> Result<(Array1<f64>, Array1<f64>)> {
        let nb = num_iterations / block_size;
        let (mut e, mut err, mut n) = (vec![0.0; nb.max(1)], vec![0.0; nb.max(1)], 0i32);
        check(unsafe { sys::mole_dmc_diffuse(self.ens, wf.raw(), self.metrop.raw, h.raw, self.branch_kind, time_step,
                                             &mut self.reference_energy, num_iterations as i32, block_size as i32,
                                             num_eq_blocks as i32, e.as_mut_ptr(), err.as_mut_ptr(), &mut n, ptr::null_mut()) })?;
        e.truncate(n as usize);
        err.truncate(n as usize);
        Ok((Array1::from_vec(e), Array1::from_vec(err)))
    }
}
impl Drop for GpuDmcRunner { fn drop(&mut self) { unsafe { sys::mole_ensemble_destroy(self.ens); } } }
