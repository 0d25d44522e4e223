"""Host-side mirror of the reference's public surface (`mole::prelude`, src/lib.rs:10-21) over the
C ABI.  Names, argument meaning and error behaviour follow the Rust items cited in each docstring;
all arithmetic happens in libmole_b200.so (CUDA, sm_100a).  Nothing here computes physics on the CPU.
"""
import ctypes as C

import numpy as np

from . import ffi
from .ffi import MoleError, check, lib, seed32

_default_ctx = None


def _destroy(fn_name, obj):
    """Destructor helper: safe during interpreter shutdown, when module globals are already cleared."""
    h = getattr(obj, "handle", None)
    l = getattr(ffi, "_lib", None) if ffi is not None else None
    if h and l is not None:
        getattr(l, fn_name)(h)
    obj.handle = None


def _h(obj):
    return obj.handle if obj is not None else None


def _dp(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Context:
    """One GPU (`mole_ctx`).  device defaults to LOCAL_RANK / cuda:0."""

    def __init__(self, device=0):
        self.handle = C.c_void_p()
        check(lib().mole_ctx_create(C.c_int32(device), C.byref(self.handle)))
        self.device = device
        self.nranks, self.rank = 1, 0

    def synchronize(self):
        check(lib().mole_ctx_synchronize(self.handle), self.handle)

    @property
    def stream(self):
        s = C.c_void_p()
        check(lib().mole_ctx_stream(self.handle, C.byref(s)), self.handle)
        return s.value

    def launch_count(self):
        n = C.c_int64()
        check(lib().mole_ctx_launch_count(self.handle, C.byref(n)), self.handle)
        return n.value

    def fp64_peak_tflops(self):
        t = C.c_double()
        check(lib().mole_bench_fp64_peak(self.handle, C.byref(t)), self.handle)
        return t.value

    def dmma_peak_tflops(self, chains=8, warps_per_sm=8):
        t = C.c_double()
        check(lib().mole_bench_dmma_peak(self.handle, C.c_int32(chains), C.c_int32(warps_per_sm), C.byref(t)), self.handle)
        return t.value

    def bench_gram(self, n_walkers, n_samples, cols, impl=0, reps=5):
        """(ms per launch, GB/s of algorithmic bytes, checksum) of the Gram contraction alone on synthetic rows."""
        ms, cs = C.c_double(), C.c_double()
        check(lib().mole_bench_gram(self.handle, C.c_int64(n_walkers), C.c_int64(n_samples), C.c_int32(cols), C.c_int32(impl),
                                    C.c_int32(reps), C.byref(ms), C.byref(cs)), self.handle)
        return ms.value, 8.0 * n_walkers * n_samples * cols / (ms.value * 1e-3) / 1e9, cs.value

    def math_probe(self, which, x):
        """Device evaluation of the kernels' branch-free exp / rcp / rsqrt / sqrt (which = 0..3)."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        out = np.empty_like(x)
        check(lib().mole_math_probe(self.handle, C.c_int32(which), _dp(x), C.c_int64(x.size), _dp(out)), self.handle)
        return out

    def comm_init(self, nranks, rank, unique_id):
        buf = (C.c_uint8 * ffi.NCCL_UNIQUE_ID_BYTES)(*bytes(unique_id))
        check(lib().mole_comm_init(self.handle, C.c_int32(nranks), C.c_int32(rank), buf), self.handle)
        self.nranks, self.rank = nranks, rank

    def close(self):
        if self.handle:
            lib().mole_comm_destroy(self.handle)
            lib().mole_ctx_destroy(self.handle)
            self.handle = None


def comm_unique_id():
    buf = (C.c_uint8 * ffi.NCCL_UNIQUE_ID_BYTES)()
    check(lib().mole_comm_get_unique_id(buf))
    return bytes(buf)


def default_context():
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(0)
    return _default_ctx


def derive_seed(master, n):
    """Metropolis::generate_seed (src/metropolis/src/traits.rs:35-37): n-th seed derived from master."""
    out = (C.c_uint8 * 32)()
    check(lib().mole_derive_seed(seed32(master), C.c_uint32(n), out))
    return bytes(out)


# ---------------------------------------------------------------------------------------------------
# wavefunction descriptors: Function / Differentiate / WaveFunction / Optimize
# ---------------------------------------------------------------------------------------------------
class WaveFunction:
    """Device-side descriptor implementing Function<f64,D=Ix2>, Differentiate, WaveFunction and
    Optimize (src/wavefunction_traits/src/lib.rs:7-24, src/optimize/src/traits.rs:8-16)."""

    KIND = None

    def __init__(self, params=(), geom=(), n_elec=None, ctx=None):
        self.ctx = ctx or default_context()
        d = ffi.WfDesc()
        d.kind, d.n_elec, d.n_params = self.KIND, n_elec, len(params) if self._optimizable() else 0
        for i, p in enumerate(params):
            d.params[i] = p
        for i, g in enumerate(geom):
            d.geom[i] = g
        self.desc = d
        self.handle = C.c_void_p()
        check(lib().mole_wf_create(self.ctx.handle, C.byref(d), C.byref(self.handle)), self.ctx.handle)

    def _optimizable(self):
        return True

    def _cfg(self, cfg):
        cfg = np.ascontiguousarray(cfg, dtype=np.float64)
        if cfg.size != 3 * self.num_electrons():
            raise MoleError(ffi.ERR_SHAPE, "configuration must be (%d, 3)" % self.num_electrons())
        return cfg

    # Function::value
    def value(self, cfg):
        out = C.c_double()
        check(lib().mole_wf_value(self.handle, _dp(self._cfg(cfg)), C.byref(out)), self.ctx.handle)
        return out.value

    # Differentiate::gradient (un-normalised grad psi)
    def gradient(self, cfg):
        out = np.empty((self.num_electrons(), 3))
        check(lib().mole_wf_gradient(self.handle, _dp(self._cfg(cfg)), _dp(out)), self.ctx.handle)
        return out

    # Differentiate::laplacian (un-normalised, summed over electrons)
    def laplacian(self, cfg):
        out = C.c_double()
        check(lib().mole_wf_laplacian(self.handle, _dp(self._cfg(cfg)), C.byref(out)), self.ctx.handle)
        return out.value

    # WaveFunction::num_electrons
    def num_electrons(self):
        n = C.c_int32()
        check(lib().mole_wf_num_electrons(self.handle, C.byref(n)))
        return n.value

    # Optimize::*
    def parameter_gradient(self, cfg):
        out = np.zeros(max(self.num_parameters(), 1))
        check(lib().mole_wf_parameter_gradient(self.handle, _dp(self._cfg(cfg)), _dp(out)), self.ctx.handle)
        return out[:self.num_parameters()]

    def num_parameters(self):
        n = C.c_int32()
        check(lib().mole_wf_num_parameters(self.handle, C.byref(n)))
        return n.value

    def parameters(self):
        out = np.zeros(max(self.num_parameters(), 1))
        check(lib().mole_wf_get_parameters(self.handle, _dp(out)))
        return out[:self.num_parameters()]

    def update_parameters(self, deltap):
        dp = np.ascontiguousarray(deltap, dtype=np.float64)
        check(lib().mole_wf_update_parameters(self.handle, _dp(dp)))

    def set_parameters(self, p):
        p = np.ascontiguousarray(p, dtype=np.float64)
        check(lib().mole_wf_set_parameters(self.handle, _dp(p)))

    def clone(self):
        new = object.__new__(type(self))
        new.__dict__.update({k: v for k, v in self.__dict__.items() if k not in ("handle", "desc")})
        d = ffi.WfDesc()
        C.memmove(C.byref(d), C.byref(self.desc), C.sizeof(d))
        for i, p in enumerate(self.parameters()):
            d.params[i] = p
        new.desc = d
        new.handle = C.c_void_p()
        check(lib().mole_wf_create(self.ctx.handle, C.byref(d), C.byref(new.handle)), self.ctx.handle)
        return new

    def __del__(self):
        _destroy("mole_wf_destroy", self)


class STO(WaveFunction):
    """1-electron Slater-type orbital exp(-alpha |x|)  (examples/dmc.rs:95-149)."""
    KIND = ffi.WF_STO_1S

    def __init__(self, alpha, ctx=None):
        super().__init__([alpha], n_elec=1, ctx=ctx)


class GaussianWaveFunction(WaveFunction):
    """exp(-(|x|/a)^2)  (examples/dmc.rs:34-92, tests/sho_optimize.rs:52-111, examples/custom_operator.rs:63-98)."""
    KIND = ffi.WF_GAUSSIAN

    def __init__(self, a, ctx=None):
        super().__init__([a], n_elec=1, ctx=ctx)


class HeliumAtomWaveFunction(WaveFunction):
    """exp(-alpha (r1 + r2))  (examples/helium_atom_singlet.rs:35-118, tests/helium_lcao.rs:23-87)."""
    KIND = ffi.WF_STO_PRODUCT

    def __init__(self, alpha, ctx=None):
        super().__init__([alpha], n_elec=2, ctx=ctx)


class HydrogenMoleculeWaveFunction(WaveFunction):
    """Heitler-London STO product sum (examples/hydrogen_molecule.rs:65-168).
    new(nuclear_separation, params) keeps the reference's argument order."""
    KIND = ffi.WF_H2_HL_STO

    def __init__(self, nuclear_separation, params, ctx=None):
        super().__init__([float(np.asarray(params).reshape(-1)[0])], [nuclear_separation], n_elec=2, ctx=ctx)


class H2WF(WaveFunction):
    """phi(x - R/2) phi(x + R/2) for H2+ (tests/hydrogen_molecular_ion_lcao.rs:51-98); no Optimize impl."""
    KIND = ffi.WF_H2P_PRODUCT

    def __init__(self, r, alpha, ctx=None):
        super().__init__([alpha], [r], n_elec=1, ctx=ctx)

    def _optimizable(self):
        return False


# ---- LCAO determinants over a hydrogen-1s basis: the API the reference's tests name (commented out upstream:
# tests/helium_lcao.rs:94-101, tests/hydrogen_molecular_ion_lcao.rs:103-107), SURVEY.md 8(f) row 3
class Hydrogen1sBasis:
    """Hydrogen1sBasis::new(ion_pos, widths): one function exp(-|r - R_c| / width) per centre.
    The device kinds take one width and one or two centres (closed set, like every other kind)."""

    def __init__(self, ion_pos, widths, general=False):
        self.ion_pos = np.asarray(ion_pos, dtype=np.float64).reshape(-1, 3)
        self.widths = [float(w) for w in widths]
        if general:                                              # for LcaoSlaterJastrow.from_orbitals: up to 8 centres
            if not 1 <= len(self.ion_pos) <= 8 or len(self.widths) not in (1, len(self.ion_pos)):
                raise MoleError(ffi.ERR_SHAPE, "general Hydrogen1sBasis: 1..8 centres, one width or one per centre")
        elif len(self.widths) != 1 or not 1 <= len(self.ion_pos) <= 2:
            raise MoleError(ffi.ERR_SHAPE, "Hydrogen1sBasis on the device: one width, one or two centres")

    def clone(self):
        return self


class Orbital:
    """Orbital::new(coefficients (n_centres x n_widths), basis): phi(r) = sum_c C[c][0] chi_c(r)."""

    def __init__(self, coefficients, basis):
        self.basis = basis
        self.coefficients = np.asarray(coefficients, dtype=np.float64).reshape(-1)
        if self.coefficients.size != len(basis.ion_pos):
            raise MoleError(ffi.ERR_SHAPE, "Orbital: one coefficient per (centre, width)")


class _LcaoWaveFunction(WaveFunction):
    def __init__(self, orbitals, n_elec, mode, ctx=None):
        basis = orbitals[0].basis
        if any(o.basis is not basis and (o.basis.widths != basis.widths or not np.array_equal(o.basis.ion_pos, basis.ion_pos))
               for o in orbitals):
            raise MoleError(ffi.ERR_SHAPE, "all orbitals must share one basis")
        nc = len(basis.ion_pos)
        kind = {(1, 2): ffi.WF_LCAO_1E_2C, (2, 1): ffi.WF_LCAO_2E_1C, (2, 2): ffi.WF_LCAO_2E_2C}.get((n_elec, nc))
        if kind is None or len(orbitals) != n_elec:
            raise MoleError(ffi.ERR_SHAPE, "LCAO on the device: (electrons, centres) in {(1, 2), (2, 1), (2, 2)}, one orbital per electron")
        self.KIND = kind
        pos = np.zeros(6)
        pos[:3 * nc] = basis.ion_pos.reshape(-1)
        coeff = np.concatenate([o.coefficients for o in orbitals])
        super().__init__(list(coeff), [float(mode), 1.0 / basis.widths[0]] + list(pos), n_elec=n_elec, ctx=ctx)


class SingleDeterminant(_LcaoWaveFunction):
    """SingleDeterminant::new(orbitals): det[phi_k(x_i)], all electrons of one spin
    (tests/hydrogen_molecular_ion_lcao.rs:107)."""

    def __init__(self, orbitals, ctx=None):
        super().__init__(orbitals, len(orbitals), 1 if len(orbitals) == 2 else 0, ctx=ctx)


class SpinDeterminantProduct(_LcaoWaveFunction):
    """SpinDeterminantProduct::new(orbitals, n_up): det_up * det_dn (tests/helium_lcao.rs:101); on the device
    two electrons with n_up = 1: psi = phi_0(x_0) phi_1(x_1)."""

    def __init__(self, orbitals, n_up, ctx=None):
        if len(orbitals) != 2 or n_up != 1:
            raise MoleError(ffi.ERR_SHAPE, "SpinDeterminantProduct on the device: two orbitals, n_up = 1")
        super().__init__(orbitals, 2, 0, ctx=ctx)


class SlaterJastrow(WaveFunction):
    """det_up * det_dn * exp(f_ee) over STO 1s/2s/2p orbitals with the Pade+polynomial Jastrow of
    theory/jastrow.tex (SURVEY.md §8(c) synthetic config 5).  params = (zeta1,zeta2,zeta3,b1,b2,b3,b4)."""
    KIND = ffi.WF_SLATER_JASTROW

    def __init__(self, n_up=5, n_dn=5, zeta=(9.64, 2.88, 2.88), b=(0.5, 1.0, 0.0, 0.0), kappa=1.0, ctx=None):
        super().__init__(list(zeta) + list(b), [kappa, n_up, n_dn], n_elec=n_up + n_dn, ctx=ctx)


class LcaoSlaterJastrow(WaveFunction):
    """SpinDeterminantProduct of LCAO orbitals over a Hydrogen1sBasis in general (tests/helium_lcao.rs:94-101: up to
    5 + 5 electrons, up to 8 centres, both spins sharing the n_orb = max(n_up, n_dn) orbitals) times the e-e Jastrow
    of theory/jastrow.tex.  coefficients[k][c] and b are the P = n_orb N_c + 4 variational parameters."""
    KIND = ffi.WF_LCAO_SJ

    def __init__(self, n_up, n_dn, ion_pos, widths, coefficients, b=(0.0, 1.0, 0.0, 0.0), kappa=1.0, ctx=None):
        pos = np.asarray(ion_pos, dtype=np.float64).reshape(-1, 3)
        nc, norb = len(pos), max(n_up, n_dn)
        coef = np.asarray(coefficients, dtype=np.float64).reshape(norb, nc)
        w = np.broadcast_to(np.asarray(widths, dtype=np.float64), (nc,))
        if nc > 8 or norb > 5:
            raise MoleError(ffi.ERR_SHAPE, "at most 8 centres and 5 orbitals per spin")
        geom = [kappa, n_up, n_dn, nc, 0, 0, 0, 0]
        for c in range(nc):
            geom += list(pos[c]) + [1.0 / w[c]]
        super().__init__(list(coef.reshape(-1)) + list(b), geom, n_elec=n_up + n_dn, ctx=ctx)

    @classmethod
    def from_orbitals(cls, orbitals, n_up, n_dn, b=(0.0, 1.0, 0.0, 0.0), kappa=1.0, ctx=None):
        """SpinDeterminantProduct::new(orbitals, n_up) over Orbital::new(coefficients, basis) objects."""
        basis = orbitals[0].basis
        return cls(n_up, n_dn, basis.ion_pos, basis.widths, [o.coefficients.reshape(-1) for o in orbitals], b, kappa, ctx)


class WaveFunctionMock(WaveFunction):
    """Constant psi (src/metropolis/src/metrop.rs:225-255); gradient is unimplemented!() upstream."""
    KIND = ffi.WF_CONSTANT

    def __init__(self, value=1.0, ctx=None):
        super().__init__([], [value], n_elec=1, ctx=ctx)

    def _optimizable(self):
        return False


# ---------------------------------------------------------------------------------------------------
# LocalOperator<T>  (src/operator/src/traits.rs:263-265, src/operator/src/operator.rs)
# ---------------------------------------------------------------------------------------------------
class LocalOperator:
    KIND = None

    def __init__(self, ion_pos=(), ion_charge=(), frequency=0.0, ctx=None):
        self.ctx = ctx or default_context()
        d = ffi.OpDesc()
        pos = np.asarray(ion_pos, dtype=np.float64).reshape(-1)
        if pos.size != 3 * len(ion_charge):
            raise MoleError(ffi.ERR_SHAPE, "ion_pos must be (N_n, 3) and match ion_charge")
        d.kind, d.n_ions, d.frequency = self.KIND, len(ion_charge), frequency
        for i, x in enumerate(pos):
            d.ion_pos[i] = x
        for i, z in enumerate(ion_charge):
            d.ion_charge[i] = int(z)
        self.desc = d
        self.handle = C.c_void_p()
        check(lib().mole_op_create(self.ctx.handle, C.byref(d), C.byref(self.handle)), self.ctx.handle)

    def act_on(self, wf, cfg):
        """LocalOperator::act_on -> H psi (not divided by psi)."""
        out = C.c_double()
        check(lib().mole_op_act_on(self.handle, wf.handle, _dp(wf._cfg(cfg)), C.byref(out)), self.ctx.handle)
        return out.value

    def __del__(self):
        _destroy("mole_op_destroy", self)


class KineticEnergy(LocalOperator):
    KIND = ffi.OP_KINETIC

    def __init__(self, ctx=None):
        super().__init__(ctx=ctx)


class IonicPotential(LocalOperator):
    KIND = ffi.OP_IONIC_POT

    def __init__(self, ion_positions, ion_charge, ctx=None):
        super().__init__(ion_positions, ion_charge, ctx=ctx)


class ElectronicPotential(LocalOperator):
    KIND = ffi.OP_ELEC_POT

    def __init__(self, ctx=None):
        super().__init__(ctx=ctx)


class IonicHamiltonian(LocalOperator):
    """IonicHamiltonian::new(t, v) (operator.rs:137)."""
    KIND = ffi.OP_IONIC

    def __init__(self, t=None, v=None, ctx=None):
        super().__init__(np.array(v.desc.ion_pos)[:3 * v.desc.n_ions], list(v.desc.ion_charge)[:v.desc.n_ions], ctx=ctx or v.ctx)


class ElectronicHamiltonian(LocalOperator):
    """ElectronicHamiltonian::{new, from_ions} (operator.rs:164-175)."""
    KIND = ffi.OP_ELECTRONIC

    def __init__(self, t=None, vion=None, velec=None, ctx=None):
        super().__init__(np.array(vion.desc.ion_pos)[:3 * vion.desc.n_ions],
                         list(vion.desc.ion_charge)[:vion.desc.n_ions], ctx=ctx or vion.ctx)

    @classmethod
    def from_ions(cls, ion_pos, ion_charge, ctx=None):
        self = object.__new__(cls)
        LocalOperator.__init__(self, ion_pos, ion_charge, ctx=ctx)
        return self


class HarmonicHamiltonian(LocalOperator):
    """examples/custom_operator.rs:30-61: T + 0.5 w^2 |x|^2."""
    KIND = ffi.OP_HARMONIC

    def __init__(self, frequency, ctx=None):
        super().__init__(frequency=frequency, ctx=ctx)


class ParameterGradient:
    """src/vmc/src/operators.rs:7-14 (observable marker)."""
    MASK = ffi.OBS_PGRAD


class WavefunctionValue:
    """src/vmc/src/operators.rs:16-24 (observable marker)."""
    MASK = ffi.OBS_WFVALUE


def operators(**named):
    """`operators!{ "Energy" => h, ... }` (src/util/src/lib.rs:13-22)."""
    return dict(named)


# ---------------------------------------------------------------------------------------------------
# Metropolis samplers  (src/metropolis/src/metrop.rs)
# ---------------------------------------------------------------------------------------------------
class _Metropolis:
    KIND = None

    def __init__(self, param, seed):
        self.param = float(param)
        self.seed = bytes(seed32(seed))
        self._n_seeds = 0
        self.handle = C.c_void_p()
        check(lib().mole_metropolis_create(C.c_int32(self.KIND), C.c_double(param), C.byref(self.handle)))

    def reseed_rng(self, s):
        self.seed = bytes(seed32(s))
        self._n_seeds = 0

    def set_compat(self, compat):
        """MOLE_COMPAT_NAN_ACCEPT: reproduce Rust's NaN-dropping `acceptance.min(1.0)` (metrop.rs:80,195)."""
        check(lib().mole_metropolis_set_compat(self.handle, C.c_uint32(compat)))
        return self

    def generate_seed(self):
        s = derive_seed(self.seed, self._n_seeds)
        self._n_seeds += 1
        return s

    def __del__(self):
        _destroy("mole_metropolis_destroy", self)


class MetropolisBox(_Metropolis):
    """MetropolisBox::{new, from_rng} (metrop.rs:34-45); the `rng` is a 32-byte Philox seed."""
    KIND = ffi.METROP_BOX

    def __init__(self, box_side, seed=None):
        super().__init__(box_side, seed if seed is not None else np.random.bytes(32))

    @classmethod
    def from_rng(cls, box_side, seed):
        return cls(box_side, seed)


class MetropolisDiffuse(_Metropolis):
    """MetropolisDiffuse::{new, from_rng, fix_nodes} (metrop.rs:113-135)."""
    KIND = ffi.METROP_DIFFUSE

    def __init__(self, time_step, seed=None):
        super().__init__(time_step, seed if seed is not None else np.random.bytes(32))
        self.fixed_node = False

    @classmethod
    def from_rng(cls, time_step, seed):
        return cls(time_step, seed)

    def fix_nodes(self):
        self.fixed_node = True   # stored but never read upstream either (metrop.rs:109,122-125)
        return self


# ---------------------------------------------------------------------------------------------------
# walker ensemble
# ---------------------------------------------------------------------------------------------------
class Ensemble:
    """SoA fp64 walker ensemble resident in HBM (`mole_ens`)."""

    def __init__(self, n_walkers, n_elec, seed, walker_offset=0, ctx=None):
        self.ctx = ctx or default_context()
        self.n_walkers, self.n_elec = int(n_walkers), int(n_elec)
        self.handle = C.c_void_p()
        check(lib().mole_ensemble_create(self.ctx.handle, C.c_int64(n_walkers), C.c_int32(n_elec), seed32(seed),
                                         C.c_uint64(walker_offset), C.byref(self.handle)), self.ctx.handle)

    def _c(self, rc):
        check(rc, self.ctx.handle)

    def init_uniform(self, lo=-1.0, hi=1.0, broadcast_walker0=False):
        self._c(lib().mole_ensemble_init_uniform(self.handle, C.c_double(lo), C.c_double(hi), C.c_int32(int(broadcast_walker0))))

    def init_normal(self, sigma=1.0, broadcast_walker0=False):
        self._c(lib().mole_ensemble_init_normal(self.handle, C.c_double(sigma), C.c_int32(int(broadcast_walker0))))

    def set_configs(self, cfgs):
        cfgs = np.ascontiguousarray(cfgs, dtype=np.float64)
        if cfgs.size == 3 * self.n_elec:
            self._c(lib().mole_ensemble_set_configs_broadcast(self.handle, _dp(cfgs)))
        elif cfgs.size == 3 * self.n_elec * self.n_walkers:
            self._c(lib().mole_ensemble_set_configs(self.handle, _dp(cfgs)))
        else:
            raise MoleError(ffi.ERR_SHAPE, "configs must be (W, N_e, 3) or (N_e, 3)")

    def get_configs(self):
        out = np.empty((self.n_walkers, self.n_elec, 3))
        self._c(lib().mole_ensemble_get_configs(self.handle, _dp(out)))
        return out

    def set_weights(self, w):
        w = np.ascontiguousarray(w, dtype=np.float64)
        if w.size != self.n_walkers:
            raise MoleError(ffi.ERR_SHAPE, "weights must have one entry per walker")
        self._c(lib().mole_ensemble_set_weights(self.handle, _dp(w)))

    def get_weights(self):
        out = np.empty(self.n_walkers)
        self._c(lib().mole_ensemble_get_weights(self.handle, _dp(out)))
        return out

    def reseed(self, seed):
        self._c(lib().mole_ensemble_reseed(self.handle, seed32(seed)))

    def snapshot(self):
        self._c(lib().mole_ensemble_snapshot(self.handle))

    def restore(self):
        self._c(lib().mole_ensemble_restore(self.handle))

    @property
    def step(self):
        s = C.c_uint32()
        self._c(lib().mole_ensemble_get_step(self.handle, C.byref(s)))
        return s.value

    @step.setter
    def step(self, v):
        self._c(lib().mole_ensemble_set_step(self.handle, C.c_uint32(v)))

    def eval_vgl(self, wf, op=None, want=("psi", "grad", "lap", "hpsi", "pgrad")):
        """Batched psi, grad psi, lap psi, H psi, d psi/dp on the current configurations."""
        W, ne, P = self.n_walkers, self.n_elec, wf.num_parameters()
        out = {}
        if "psi" in want:
            out["psi"] = np.empty(W)
        if "grad" in want:
            out["grad"] = np.empty((W, ne, 3))
        if "lap" in want:
            out["lap"] = np.empty(W)
        if "hpsi" in want and op is not None:
            out["hpsi"] = np.empty(W)
        if "pgrad" in want and P > 0:
            out["pgrad"] = np.empty((W, P))
        self._c(lib().mole_eval_vgl(self.handle, wf.handle, _h(op), _dp(out.get("psi")), _dp(out.get("grad")),
                                    _dp(out.get("lap")), _dp(out.get("hpsi")), _dp(out.get("pgrad"))))
        return out

    def sweep(self, wf, metrop, op, n_sweeps, n_discard=0, block_size=1, observables=ffi.OBS_ENERGY, compat=0,
              traces=(), keep_series=False, append_series=False):
        """mole_sweep.  traces: subset of ("energy","wfvalue","kinetic","pgrad","accept").
        Returned traces are indexed [walker, sample(, k)] / accept [walker, sweep, electron].
        keep_series / append_series: keep the E_L samples on the device for series_analyze()."""
        W, ne, P = self.n_walkers, self.n_elec, wf.num_parameters()
        ns = n_sweeps - n_discard
        a = ffi.SweepArgs()
        a.n_sweeps, a.n_discard, a.block_size, a.observables, a.compat = n_sweeps, n_discard, block_size, observables, compat
        a.flags = (ffi.SWEEP_KEEP_SERIES if keep_series else 0) | (ffi.SWEEP_APPEND_SERIES if append_series else 0)
        bufs = {}
        if "energy" in traces and ns > 0:
            bufs["energy"] = np.empty((ns, W)); a.energy_trace = bufs["energy"].ctypes.data
        if "wfvalue" in traces and ns > 0:
            bufs["wfvalue"] = np.empty((ns, W)); a.wfvalue_trace = bufs["wfvalue"].ctypes.data
        if "kinetic" in traces and ns > 0:
            bufs["kinetic"] = np.empty((ns, W)); a.kinetic_trace = bufs["kinetic"].ctypes.data
        if "pgrad" in traces and ns > 0 and P > 0:
            bufs["pgrad"] = np.empty((ns, P, W)); a.pgrad_trace = bufs["pgrad"].ctypes.data
        if "accept" in traces and n_sweeps > 0:
            bufs["accept"] = np.empty((n_sweeps, ne, W), dtype=np.uint8); a.accept_trace = bufs["accept"].ctypes.data
        self._c(lib().mole_sweep(self.handle, wf.handle, metrop.handle, _h(op), C.byref(a)))
        out = {}
        for k, v in bufs.items():
            out[k] = np.ascontiguousarray(np.moveaxis(v, -1, 0))   # walker-major
        return out

    # ---- series statistics on the device (scripts/statfor.rs) ----
    def series_length(self):
        n = C.c_int64()
        self._c(lib().mole_series_length(self.handle, C.byref(n)))
        return n.value

    def series_clear(self):
        self._c(lib().mole_series_clear(self.handle))

    def series_analyze(self, block_sizes=None, drop_last=False, per_walker=False):
        """statfor per walker on the kept E_L series; returns walker means (and per-walker arrays).
        block_sizes=None uses the reference's schedule (statfor.rs:59-66)."""
        n = self.series_length() - (1 if drop_last else 0)
        if block_sizes is None:
            block_sizes = series_block_sizes(n)
        bs = np.ascontiguousarray(block_sizes, dtype=np.int32)
        lags = max(min(ffi.SERIES_MAX_LAG, n - 1), 0)
        W = self.n_walkers
        ms = ffi.SeriesStats()
        corr, berr = np.empty(lags), np.empty(bs.size)
        pw = np.empty((W, 5)) if per_walker else None
        pwc = np.empty((lags, W)) if per_walker else None
        pwb = np.empty((bs.size, W)) if per_walker else None
        ptr = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None and a.size else None
        self._c(lib().mole_series_analyze(self.handle, C.c_int32(1 if drop_last else 0), C.byref(ms), ptr(corr),
                                          C.c_int32(bs.size), ptr(bs), ptr(berr), ptr(pw), ptr(pwc), ptr(pwb)))
        out = dict(average=ms.average, variance=ms.variance, tcorr=ms.tcorr, n_eff=ms.n_eff, sigma=ms.sigma,
                   corr=corr, block_sizes=bs, block_errors=berr)
        if per_walker:
            out.update(per_walker=pw, per_walker_corr=pwc.T.copy(), per_walker_block_errors=pwb.T.copy())
        return out

    def series_get(self, walker):
        out = np.empty(self.series_length())
        self._c(lib().mole_series_get(self.handle, C.c_int64(walker), out.ctypes.data_as(C.c_void_p)))
        return out

    def series_write_text(self, walker, path):
        """One float per line: the input format of scripts/statfor.py:17-19 / statfor.rs:7-13."""
        self._c(lib().mole_series_write_text(self.handle, C.c_int64(walker), str(path).encode()))

    # ---- checkpoint / restart ----
    def save(self, path):
        self._c(lib().mole_ensemble_save(self.handle, str(path).encode()))

    def load(self, path):
        self._c(lib().mole_ensemble_load(self.handle, str(path).encode()))

    def run_logged(self, wf, metrop, op, steps, block_size, log=None, observables=ffi.OBS_ENERGY, compat=0,
                   keep_series=False):
        """Runner::run with a Log (montecarlo.rs:24-46): `log(dict) -> str` is called once per sampled
        block with block-level reductions; non-empty output is printed like the reference does."""
        def _cb(_user, d):
            d = d.contents
            s = log({f: getattr(d, f) for f, _ in ffi.BlockLog._fields_})
            if s:
                print(s)
        cb = ffi.LOG_FN(_cb) if log is not None else C.cast(None, ffi.LOG_FN)
        self._c(lib().mole_runner_run_logged(self.handle, wf.handle, metrop.handle, _h(op), C.c_uint32(observables),
                                             C.c_uint32(compat), C.c_int32(steps), C.c_int32(block_size),
                                             C.c_uint32(ffi.SWEEP_KEEP_SERIES if keep_series else 0), cb, None))

    def acc_reset(self):
        self._c(lib().mole_acc_reset(self.handle))

    def acc_get(self):
        a = ffi.AccHost()
        self._c(lib().mole_acc_get(self.handle, C.byref(a)))
        return a

    def acc_allreduce(self):
        self._c(lib().mole_acc_allreduce(self.handle))

    # ---- optimisation moments of a large-P kind: Gram matrix of the per-sample rows (1, E_L, O_k) ----
    def gram_get(self):
        n = C.c_int32()
        self._c(lib().mole_gram_get(self.handle, C.byref(n), None))
        g = np.empty((n.value, n.value))
        self._c(lib().mole_gram_get(self.handle, C.byref(n), _dp(g)))
        return g

    def gram_allreduce(self):
        self._c(lib().mole_gram_allreduce(self.handle))

    def dmc_block_select(self, impl):
        """0: one persistent cooperative launch per DMC block when the population fits one walker per resident thread
        (default); 1: per-step launches; 2: the persistent launch for every population it can serve."""
        self._c(lib().mole_dmc_block_select(self.handle, C.c_int32(impl)))

    def gram_select(self, impl):
        """0: DMMA (tensor cores, default); 1: FP64 vector pipe."""
        self._c(lib().mole_gram_select(self.handle, C.c_int32(impl)))

    def health(self):
        """(non-finite VMC samples skipped, DMC walker-steps killed) since the last acc_reset."""
        h = ffi.EnsHealth()
        self._c(lib().mole_ensemble_health(self.handle, C.byref(h)))
        return int(h.nonfinite_samples), int(h.nonfinite_dmc_walkers)

    def acc_device_ptr(self):
        p, n = C.c_void_p(), C.c_int32()
        self._c(lib().mole_acc_device_ptr(self.handle, C.byref(p), C.byref(n)))
        return p.value, n.value

    def dmc_step(self, wf, metrop, op, time_step, reference_energy):
        swe, sw = C.c_double(), C.c_double()
        self._c(lib().mole_dmc_step(self.handle, wf.handle, metrop.handle, op.handle, C.c_double(time_step),
                                    C.c_double(reference_energy), C.byref(swe), C.byref(sw)))
        return swe.value, sw.value

    def dmc_block(self, wf, metrop, op, branch_kind, time_step, reference_energy, n_steps):
        """n_steps x (time step, ensemble energy, branch) without host reads; returns the step energies."""
        out = np.empty(n_steps)
        self._c(lib().mole_dmc_block(self.handle, wf.handle, metrop.handle, op.handle, C.c_int32(branch_kind),
                                     C.c_double(time_step), C.c_double(reference_energy), C.c_int32(n_steps),
                                     out.ctypes.data_as(C.c_void_p)))
        return out

    def branch(self, kind):
        self._c(lib().mole_branch(self.handle, C.c_int32(kind)))

    def island_imbalance(self):
        """largest / smallest island weight per walker after the last SRBrancher block (1.0 on a single rank)."""
        r = C.c_double()
        self._c(lib().mole_dmc_island_imbalance(self.handle, C.byref(r)))
        return r.value

    def rebalance(self):
        """cross-rank population rebalancing (mole_rebalance): equal weights everywhere, surplus walkers migrate"""
        self._c(lib().mole_rebalance(self.handle))

    def branch_sources(self):
        out = np.empty(self.n_walkers, dtype=np.int32)
        self._c(lib().mole_branch_sources(self.handle, out.ctypes.data_as(C.c_void_p)))
        return out

    def __del__(self):
        _destroy("mole_ensemble_destroy", self)


def series_block_sizes(n):
    """Block-size schedule of the reference's blocking analysis (scripts/statfor.rs:59-66)."""
    k = C.c_int32()
    ffi.check(lib().mole_series_block_sizes(C.c_int64(n), None, C.byref(k)))
    out = np.empty(k.value, dtype=np.int32)
    if k.value:
        ffi.check(lib().mole_series_block_sizes(C.c_int64(n), out.ctypes.data_as(C.c_void_p), C.byref(k)))
    return out


def acc_finalize(acc):
    """(mean energy, blocking error, acceptance, energy gradient) from reduced moments."""
    e, err, ac = C.c_double(), C.c_double(), C.c_double()
    g = np.zeros(max(acc.n_params, 1))
    check(lib().mole_acc_finalize(C.byref(acc), C.byref(e), C.byref(err), C.byref(ac), _dp(g)))
    return e.value, err.value, ac.value, g[:acc.n_params]


def rebalance_plan(totals, counts, u):
    """(shares[r], moves[src][dst]) of mole_rebalance for the ranks' total weights and walker counts and a shared draw u."""
    t = np.ascontiguousarray(totals, dtype=np.float64)
    c = np.ascontiguousarray(counts, dtype=np.int64)
    shares = np.zeros(t.size, dtype=np.int64)
    moves = np.zeros((t.size, t.size), dtype=np.int64)
    check(lib().mole_rebalance_plan(C.c_int32(t.size), _dp(t), c.ctypes.data_as(C.c_void_p), C.c_double(u),
                                    shares.ctypes.data_as(C.c_void_p), moves.ctypes.data_as(C.c_void_p)))
    return shares, moves


def gram_finalize(gram):
    """(mean energy, energy gradient) from the Gram matrix of a large-P kind (Ensemble.gram_get)."""
    g = np.ascontiguousarray(gram, dtype=np.float64)
    e = C.c_double()
    grad = np.zeros(g.shape[0] - 2)
    check(lib().mole_gram_finalize(C.c_int32(g.shape[0]), _dp(g), C.byref(e), _dp(grad)))
    return e.value, grad


# ---------------------------------------------------------------------------------------------------
# Optimizers  (src/optimize/src/optimizers.rs)
# ---------------------------------------------------------------------------------------------------
class Optimizer:
    KIND = None

    def __init__(self, nparm, step_size, momentum_parameter=0.0, history=5, compat=0):
        self.nparm = nparm
        self.handle = C.c_void_p()
        check(lib().mole_opt_create(C.c_int32(self.KIND), C.c_int32(nparm), C.c_double(step_size),
                                    C.c_double(momentum_parameter), C.c_int32(history), C.c_uint32(compat),
                                    C.byref(self.handle)))

    def compute_parameter_update(self, pars, acc):
        """Optimizer::compute_parameter_update (optimize/src/traits.rs:18-25) from the reduced moments."""
        pars = np.ascontiguousarray(pars, dtype=np.float64)
        dp = np.empty(self.nparm)
        if isinstance(acc, np.ndarray):                          # Gram matrix of a large-P kind (Ensemble.gram_get)
            g = np.ascontiguousarray(acc, dtype=np.float64)
            check(lib().mole_opt_step_gram(self.handle, _dp(pars), C.c_int32(g.shape[0]), _dp(g), _dp(dp)))
        else:
            check(lib().mole_opt_step(self.handle, _dp(pars), C.byref(acc), _dp(dp)))
        return dp

    def sr_matrix(self, acc):
        S = np.empty((self.nparm, self.nparm))
        if isinstance(acc, np.ndarray):
            g = np.ascontiguousarray(acc, dtype=np.float64)
            check(lib().mole_opt_sr_matrix_gram(self.handle, C.c_int32(g.shape[0]), _dp(g), _dp(S)))
        else:
            check(lib().mole_opt_sr_matrix(self.handle, C.byref(acc), _dp(S)))
        return S

    def __del__(self):
        _destroy("mole_opt_destroy", self)


class SteepestDescent(Optimizer):
    KIND = ffi.OPT_SD

    def __init__(self, step_size, nparm=1, compat=0):
        super().__init__(nparm, step_size, compat=compat)


class MomentumDescent(Optimizer):
    KIND = ffi.OPT_MOMENTUM

    def __init__(self, step_size, momentum_parameter, nparm, compat=0):
        super().__init__(nparm, step_size, momentum_parameter, compat=compat)


class NesterovMomentum(Optimizer):
    KIND = ffi.OPT_NESTEROV

    def __init__(self, step_size, momentum_parameter, nparm, compat=0):
        super().__init__(nparm, step_size, momentum_parameter, compat=compat)


class OnlineLbfgs(Optimizer):
    KIND = ffi.OPT_LBFGS

    def __init__(self, step_size, history, nparm, compat=0):
        super().__init__(nparm, step_size, history=history, compat=compat)


class StochasticReconfiguration(Optimizer):
    KIND = ffi.OPT_SR

    def __init__(self, step_size, nparm=1, compat=0):
        super().__init__(nparm, step_size, compat=compat)

    def set_regularization(self, diag_scale=1.01, diag_shift=0.0):
        """S_kk <- S_kk * diag_scale + diag_shift before the solve (default: optimizers.rs:225-231)."""
        check(lib().mole_opt_set_sr_regularization(self.handle, C.c_double(diag_scale), C.c_double(diag_shift)))
        return self


# ---------------------------------------------------------------------------------------------------
# Sampler / Runner / VmcRunner / DmcRunner
# ---------------------------------------------------------------------------------------------------
class MonteCarloResult:
    """src/montecarlo/src/traits.rs:5-9.  data[name] is [walker, sample(, k)]."""

    def __init__(self, wave_function, acceptance, data):
        self.wave_function, self.acceptance, self.data = wave_function, acceptance, data


def _obs_mask(observables):
    mask, ham = 0, None
    for name, op in observables.items():
        if isinstance(op, (ParameterGradient, WavefunctionValue)) or op in (ParameterGradient, WavefunctionValue):
            mask |= op.MASK
        elif isinstance(op, KineticEnergy) and name != "Energy":
            mask |= ffi.OBS_KINETIC
        elif isinstance(op, LocalOperator):
            if name != "Energy":
                raise MoleError(ffi.ERR_INVALID_ARG, "the device samples one Hamiltonian, registered as \"Energy\"")
            mask |= ffi.OBS_ENERGY
            ham = op
        else:
            raise MoleError(ffi.ERR_INVALID_ARG, "user-defined LocalOperator impls cannot run on the device")
    return mask, ham


class Sampler:
    """Sampler::{new, with_initial_configuration} (src/montecarlo/src/samplers.rs:40-71).
    n_walkers > 1 turns the single chain into an ensemble of independent chains (the GPU analogue
    of the per-worker clones, vmc.rs:56); independent=False starts them all from walker 0's draw."""

    def __init__(self, wave_function, metrop, observables, n_walkers=1, independent=False, walker_offset=0,
                 initial_configuration=None, compat=0):
        self.wave_function, self.metropolis, self.observables = wave_function, metrop, observables
        self.mask, self.ham = _obs_mask(observables)
        self.compat = compat
        self.ensemble = Ensemble(n_walkers, wave_function.num_electrons(), metrop.seed, walker_offset,
                                 ctx=wave_function.ctx)
        if initial_configuration is not None:
            self.ensemble.set_configs(initial_configuration)
        else:
            self.ensemble.init_uniform(-1.0, 1.0, broadcast_walker0=not independent)   # samplers.rs:46
        self.acceptance_ = 0.0

    @classmethod
    def new(cls, wave_function, metrop, observables, **kw):
        return cls(wave_function, metrop, observables, **kw)

    @classmethod
    def with_initial_configuration(cls, wave_function, metrop, observables, cfg, **kw):
        return cls(wave_function, metrop, observables, initial_configuration=cfg, **kw)

    def reseed_rng(self, s):
        self.metropolis.reseed_rng(s)
        self.ensemble.reseed(s)

    def generate_seed(self):
        return self.metropolis.generate_seed()

    def acceptance(self):
        return self.acceptance_

    def num_observables(self):
        return len(self.observables)

    def observable_names(self):
        return list(self.observables.keys())


class Runner:
    """Runner::{new, run} (src/montecarlo/src/montecarlo.rs:20-46)."""

    def __init__(self, sampler, logger=None):
        self.sampler, self.logger = sampler, logger

    def run(self, steps, block_size, traces=True):
        s = self.sampler
        ens, wf = s.ensemble, s.wave_function
        W, ne, P = ens.n_walkers, ens.n_elec, wf.num_parameters()
        if block_size < 1 or not steps >= 2 * block_size:
            raise MoleError(ffi.ERR_ASSERT, "assertion failed: steps >= 2 * block_size")
        blocks = steps // block_size
        ns = (blocks - 1) * block_size
        bufs = {}
        if traces:
            if s.mask & ffi.OBS_ENERGY:
                bufs["Energy"] = np.empty((ns, W))
            if s.mask & ffi.OBS_WFVALUE:
                bufs["Wavefunction value"] = np.empty((ns, W))
            if s.mask & ffi.OBS_KINETIC:
                bufs["Kin. Energy"] = np.empty((ns, W))
            if s.mask & ffi.OBS_PGRAD:
                bufs["Parameter gradient"] = np.empty((ns, P, W))
        ens.acc_reset()
        if self.logger is not None and not traces:
            # Log (montecarlo/src/traits.rs:44-47): one launch per block, logger.log(block reductions) per block
            ens.run_logged(wf, s.metropolis, s.ham, steps, block_size, log=self.logger.log, observables=s.mask,
                           compat=s.compat)
            acc = ens.acc_get()
            s.acceptance_ = acc.n_accept / ne
            self.acc = acc
            return MonteCarloResult(wf, s.acceptance_, {})
        check(lib().mole_runner_run(ens.handle, wf.handle, s.metropolis.handle, _h(s.ham), C.c_uint32(s.mask),
                                    C.c_uint32(s.compat), C.c_int32(steps), C.c_int32(block_size),
                                    _dp(bufs.get("Energy")), _dp(bufs.get("Wavefunction value")),
                                    _dp(bufs.get("Kin. Energy")), _dp(bufs.get("Parameter gradient")), None),
              ens.ctx.handle)
        acc = ens.acc_get()
        s.acceptance_ = acc.n_accept / ne                       # samplers.rs:113 (summed over walkers)
        data = {}
        for name, v in bufs.items():
            user = [k for k, o in s.observables.items()
                    if (name == "Energy" and o is s.ham) or (name == "Kin. Energy" and isinstance(o, KineticEnergy) and o is not s.ham)
                    or (name == "Parameter gradient" and (o is ParameterGradient or isinstance(o, ParameterGradient)))
                    or (name == "Wavefunction value" and (o is WavefunctionValue or isinstance(o, WavefunctionValue)))]
            data[user[0] if user else name] = np.ascontiguousarray(np.moveaxis(v, -1, 0))
        self.acc = acc
        return MonteCarloResult(wf, s.acceptance_, data)


class VmcRunner:
    """VmcRunner::{new, run_optimization} (src/vmc/src/vmc.rs:34-106)."""

    def __init__(self, sampler, optimizer, logger=None):
        self.sampler, self.optimizer, self.logger = sampler, optimizer, logger

    def run_optimization(self, iters, total_samples, block_size, nworkers, restart_each_iter=True, verbose=False):
        s = self.sampler
        wf = s.wave_function
        ens = s.ensemble
        if ens.n_walkers * ens.ctx.nranks != nworkers:
            # vec![self.sampler.clone(); nworkers] (vmc.rs:56): every worker starts from the master cfg
            cfg0 = ens.get_configs()[0]
            ens = Ensemble(nworkers // ens.ctx.nranks, ens.n_elec, s.metropolis.seed,
                           walker_offset=ens.ctx.rank * (nworkers // ens.ctx.nranks), ctx=ens.ctx)
            ens.set_configs(cfg0)
        P = wf.num_parameters()
        en, er, ac = np.empty(iters), np.empty(iters), np.empty(iters)
        ph = np.empty((iters, max(P, 1)))
        flags = ffi.VMC_RESTART_EACH_ITER if restart_each_iter else 0
        check(lib().mole_vmc_run_optimization(ens.handle, wf.handle, s.metropolis.handle, _h(s.ham), self.optimizer.handle,
                                              seed32(s.metropolis.seed), C.c_int32(iters), C.c_int64(total_samples),
                                              C.c_int32(block_size), C.c_uint32(s.compat), C.c_uint32(flags), _dp(en),
                                              _dp(er), _dp(ac), _dp(ph)), ens.ctx.handle)
        if verbose:
            for i in range(iters):   # vmc.rs:93-98
                print("Energy:      %.8f +/- %.9f    accept: %.8f" % (en[i], er[i], ac[i]))
        self.acceptance, self.param_history, self.ensemble = ac, ph[:, :P], ens
        return wf, en, er


class SRBrancher:
    """src/dmc/src/branching.rs:7-40."""
    KIND = ffi.BRANCH_SR

    @classmethod
    def new(cls):
        return cls()


class SimpleBranching:
    """src/dmc/src/branching.rs:42-92."""
    KIND = ffi.BRANCH_SIMPLE

    @classmethod
    def new(cls):
        return cls()


class DmcRunner:
    """DmcRunner::{new, diffuse} (src/dmc/src/dmc.rs:40-153).
    identical_start=True reproduces `vec![(1.0, cfg); n]` (all walkers share one N(0,1) draw)."""

    def __init__(self, guiding_wave_function, num_walkers, reference_energy, hamiltonian, metropolis, branching,
                 identical_start=True, walker_offset=0):
        self.wf, self.reference_energy, self.hamiltonian = guiding_wave_function, reference_energy, hamiltonian
        self.metrop, self.branching = metropolis, branching
        self.ensemble = Ensemble(num_walkers, self.wf.num_electrons(), metropolis.seed, walker_offset, ctx=self.wf.ctx)
        self.ensemble.init_normal(1.0, broadcast_walker0=identical_start)

    @classmethod
    def new(cls, *a, **kw):
        return cls(*a, **kw)

    def diffuse(self, time_step, num_iterations, block_size, num_eq_blocks, verbose=False, want_steps=False):
        nb = max(num_iterations // block_size, 1)
        en, er = np.empty(nb), np.empty(nb)
        se = np.empty(nb * block_size) if want_steps else None
        n_out, eref = C.c_int32(), C.c_double(self.reference_energy)
        check(lib().mole_dmc_diffuse(self.ensemble.handle, self.wf.handle, self.metrop.handle, self.hamiltonian.handle,
                                     C.c_int32(self.branching.KIND), C.c_double(time_step), C.byref(eref),
                                     C.c_int32(num_iterations), C.c_int32(block_size), C.c_int32(num_eq_blocks), _dp(en),
                                     _dp(er), C.byref(n_out), _dp(se)), self.ensemble.ctx.handle)
        self.reference_energy = eref.value
        self.step_energies = se
        k = n_out.value
        if verbose and k:
            print("DMC Energy:   %.8f +/- %.8f" % (en[k - 1], er[k - 1]))
        return en[:k].copy(), er[:k].copy()
