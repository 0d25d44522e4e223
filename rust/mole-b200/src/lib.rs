//! Safe wrapper: the reference's traits implemented over libmole_b200.so.
//! SOURCES ONLY - never compiled in this image (no rustc); see rust/README.md and INTEGRATION.md.
//!
//! Every impl cites the reference item it stands in for (paths inside Jvanrhijn/mole).
use errors::Error;
use mole_b200_sys as sys;
use ndarray::{Array1, Array2, Ix2};
use operator::{LocalOperator, OperatorValue};
use optimize::Optimize;
use std::ptr;
use wavefunction_traits::{Differentiate, Function, WaveFunction};

pub(crate) type Result<T> = std::result::Result<T, Error>;

mod drivers;
pub use drivers::{BlockLog, EmptyLogger, GpuDmcRunner, GpuLog, GpuMetropolis, GpuOptimizer, GpuSampler};

/// status code -> errors::Error (src/errors/src/lib.rs:8-15); codes >= 100 are CUDA/NCCL/argument failures
pub(crate) fn check(rc: i32) -> Result<()> {
    match rc {
        sys::MOLE_OK => Ok(()),
        sys::MOLE_ERR_FUNC => Err(Error::FuncError),
        sys::MOLE_ERR_OPERATOR_VALUE_ACCESS => Err(Error::OperatorValueAccessError),
        sys::MOLE_ERR_DATA_ACCESS => Err(Error::DataAccessError),
        sys::MOLE_ERR_EMPTY_CACHE => Err(Error::EmptyCacheError),
        other => panic!("mole_b200 error {}", other), // LinalgError/ShapeError carry payload types of ndarray(-linalg)
    }
}

pub struct Context { raw: *mut sys::mole_ctx_s }
impl Context {
    pub fn new(device: i32) -> Result<Self> {
        let mut raw = ptr::null_mut();
        check(unsafe { sys::mole_ctx_create(device, &mut raw) })?;
        Ok(Self { raw })
    }
}
impl Drop for Context { fn drop(&mut self) { unsafe { sys::mole_ctx_destroy(self.raw); } } }

/// Device-side trial wavefunction descriptor.  Stands in for the user structs of the reference's examples:
/// `HydrogenMoleculeWaveFunction` (examples/hydrogen_molecule.rs:65-168), `HeliumAtomWaveFunction`
/// (examples/helium_atom_singlet.rs:35-118), `GaussianWaveFunction`/`STO` (examples/dmc.rs:34-149), `H2WF`
/// (tests/hydrogen_molecular_ion_lcao.rs:51-98).
pub struct GpuWaveFunction { raw: *mut sys::mole_wf_s, params: Array1<f64>, n_elec: usize }

impl GpuWaveFunction {
    fn create(ctx: &Context, kind: i32, n_elec: i32, params: &[f64], geom: &[f64], optimizable: bool) -> Result<Self> {
        let mut d = sys::mole_wf_desc { kind, n_elec, n_params: if optimizable { params.len() as i32 } else { 0 },
                                        reserved: 0, params: [0.0; 8], geom: [0.0; 8] };
        d.params[..params.len()].copy_from_slice(params);
        d.geom[..geom.len()].copy_from_slice(geom);
        let mut raw = ptr::null_mut();
        check(unsafe { sys::mole_wf_create(ctx.raw, &d, &mut raw) })?;
        Ok(Self { raw, params: Array1::from_vec(params.to_vec()), n_elec: n_elec as usize })
    }
    pub fn h2_heitler_london(ctx: &Context, nuclear_separation: f64, alpha: f64) -> Result<Self> {
        Self::create(ctx, 3, 2, &[alpha], &[nuclear_separation], true)
    }
    pub fn helium_sto_product(ctx: &Context, alpha: f64) -> Result<Self> { Self::create(ctx, 2, 2, &[alpha], &[], true) }
    pub fn gaussian(ctx: &Context, a: f64) -> Result<Self> { Self::create(ctx, 1, 1, &[a], &[], true) }
    pub fn sto_1s(ctx: &Context, alpha: f64) -> Result<Self> { Self::create(ctx, 0, 1, &[alpha], &[], true) }
    pub fn h2_plus_product(ctx: &Context, r: f64, alpha: f64) -> Result<Self> { Self::create(ctx, 4, 1, &[alpha], &[r], false) }
    pub fn slater_jastrow(ctx: &Context, n_up: usize, n_dn: usize, zeta: [f64; 3], b: [f64; 4], kappa: f64) -> Result<Self> {
        let p = [zeta[0], zeta[1], zeta[2], b[0], b[1], b[2], b[3]];
        Self::create(ctx, 5, (n_up + n_dn) as i32, &p, &[kappa, n_up as f64, n_dn as f64], true)
    }
    /// LCAO determinants over a hydrogen-1s basis (kinds 7-9 of include/mole_b200.h): the `Hydrogen1sBasis` /
    /// `Orbital` / `SingleDeterminant` / `SpinDeterminantProduct` API named at tests/helium_lcao.rs:94-101 and
    /// tests/hydrogen_molecular_ion_lcao.rs:103-107.  `ion_pos` has one or two rows, `coefficients[k][c]` is
    /// orbital k's weight on centre c, `determinant` selects the 2x2 determinant instead of the spin product.
    pub fn lcao(ctx: &Context, ion_pos: &Array2<f64>, width: f64, coefficients: &Array2<f64>, determinant: bool) -> Result<Self> {
        let (nc, ne) = (ion_pos.nrows(), coefficients.nrows());
        let bad_shape = || Error::ShapeError(ndarray::ShapeError::from_kind(ndarray::ErrorKind::IncompatibleShape));
        let kind = match (ne, nc) { (1, 2) => 7, (2, 1) => 8, (2, 2) => 9, _ => return Err(bad_shape()) };
        if coefficients.ncols() != nc || ion_pos.ncols() != 3 { return Err(bad_shape()); }
        let mut geom = [0.0f64; 8];
        geom[0] = if determinant && ne == 2 { 1.0 } else { 0.0 };
        geom[1] = 1.0 / width;
        for (i, v) in ion_pos.iter().enumerate() { geom[2 + i] = *v; }
        let p: Vec<f64> = coefficients.iter().cloned().collect();
        Self::create(ctx, kind, ne as i32, &p, &geom, true)
    }
    pub(crate) fn raw(&self) -> *mut sys::mole_wf_s { self.raw }
}
impl Drop for GpuWaveFunction { fn drop(&mut self) { unsafe { sys::mole_wf_destroy(self.raw); } } }

/// src/wavefunction_traits/src/lib.rs:7-11
impl Function<f64> for GpuWaveFunction {
    type D = Ix2;
    fn value(&self, cfg: &Array2<f64>) -> Result<f64> {
        let mut out = 0.0;
        check(unsafe { sys::mole_wf_value(self.raw, cfg.as_ptr(), &mut out) })?;
        Ok(out)
    }
}
/// src/wavefunction_traits/src/lib.rs:14-20
impl Differentiate for GpuWaveFunction {
    type D = Ix2;
    fn gradient(&self, cfg: &Array2<f64>) -> Result<Array2<f64>> {
        let mut out = Array2::<f64>::zeros((self.n_elec, 3));
        check(unsafe { sys::mole_wf_gradient(self.raw, cfg.as_ptr(), out.as_mut_ptr()) })?;
        Ok(out)
    }
    fn laplacian(&self, cfg: &Array2<f64>) -> Result<f64> {
        let mut out = 0.0;
        check(unsafe { sys::mole_wf_laplacian(self.raw, cfg.as_ptr(), &mut out) })?;
        Ok(out)
    }
}
/// src/wavefunction_traits/src/lib.rs:22-24
impl WaveFunction for GpuWaveFunction { fn num_electrons(&self) -> usize { self.n_elec } }
/// src/optimize/src/traits.rs:8-16
impl Optimize for GpuWaveFunction {
    fn parameter_gradient(&self, cfg: &Array2<f64>) -> Result<Array1<f64>> {
        let mut out = Array1::<f64>::zeros(self.params.len());
        check(unsafe { sys::mole_wf_parameter_gradient(self.raw, cfg.as_ptr(), out.as_mut_ptr()) })?;
        Ok(out)
    }
    fn update_parameters(&mut self, deltap: &Array1<f64>) {
        self.params += deltap;
        unsafe { sys::mole_wf_update_parameters(self.raw, deltap.as_ptr()); }
    }
    fn parameters(&self) -> &Array1<f64> { &self.params }
    fn num_parameters(&self) -> usize { self.params.len() }
}

/// `ElectronicHamiltonian`, `IonicHamiltonian`, `KineticEnergy`, `IonicPotential`, `ElectronicPotential`
/// (src/operator/src/operator.rs:16-184) and the SHO operator of examples/custom_operator.rs:30-61.
pub struct GpuHamiltonian { raw: *mut sys::mole_op_s }
impl GpuHamiltonian {
    pub fn from_ions(ctx: &Context, ion_pos: &Array2<f64>, ion_charge: &Array1<i32>) -> Result<Self> {
        let mut d = sys::mole_op_desc { kind: 4, n_ions: ion_charge.len() as i32, ion_pos: [0.0; 24], ion_charge: [0; 8], frequency: 0.0 };
        for (i, x) in ion_pos.iter().enumerate() { d.ion_pos[i] = *x; }
        for (i, z) in ion_charge.iter().enumerate() { d.ion_charge[i] = *z; }
        let mut raw = ptr::null_mut();
        check(unsafe { sys::mole_op_create(ctx.raw, &d, &mut raw) })?;
        Ok(Self { raw })
    }
}
unsafe impl Send for GpuHamiltonian {}
unsafe impl Sync for GpuHamiltonian {}
/// src/operator/src/traits.rs:263-265
impl LocalOperator<GpuWaveFunction> for GpuHamiltonian {
    fn act_on(&self, wf: &GpuWaveFunction, cfg: &Array2<f64>) -> Result<OperatorValue> {
        let mut out = 0.0;
        check(unsafe { sys::mole_op_act_on(self.raw, wf.raw(), cfg.as_ptr(), &mut out) })?;
        Ok(OperatorValue::Scalar(out))
    }
}

/// `VmcRunner::run_optimization` (src/vmc/src/vmc.rs:43-106) over one ensemble of `nworkers` walkers.
pub struct GpuVmcRunner { pub ens: *mut sys::mole_ens_s, pub metrop: *mut sys::mole_metrop_s, pub opt: *mut sys::mole_opt_s,
                          pub master_seed: [u8; 32] }
impl GpuVmcRunner {
    pub fn run_optimization(&mut self, wf: &mut GpuWaveFunction, h: &GpuHamiltonian, iters: usize, total_samples: usize,
                            block_size: usize, _nworkers: usize) -> Result<(Array1<f64>, Array1<f64>)> {
        let mut e = vec![0.0; iters];
        let mut err = vec![0.0; iters];
        let mut acc = vec![0.0; iters];
        check(unsafe { sys::mole_vmc_run_optimization(self.ens, wf.raw(), self.metrop, h.raw, self.opt, self.master_seed.as_ptr(),
                                                      iters as i32, total_samples as i64, block_size as i32, 0, 1,
                                                      e.as_mut_ptr(), err.as_mut_ptr(), acc.as_mut_ptr(), ptr::null_mut()) })?;
        for i in 0..iters {
            println!("Energy:      {:.8} +/- {:.9}    accept: {:.8}", e[i], err[i], acc[i]); // vmc.rs:93-98
        }
        let mut p = vec![0.0; wf.params.len()];
        unsafe { sys::mole_wf_get_parameters(wf.raw(), p.as_mut_ptr()); }
        wf.params = Array1::from_vec(p);
        Ok((Array1::from_vec(e), Array1::from_vec(err)))
    }
}
