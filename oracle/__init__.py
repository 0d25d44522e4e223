"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes front-end of oracle/_build/libmole_oracle.so, the CPU restatement of the reference's
walker-ensemble VMC/DMC hot path (see oracle_wf.hpp / oracle_mc.hpp for the file:line map).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package; mole_b200/ never does.

Parity status: the reference is Rust nightly + MKL and cannot be built in this image, so there
is no oracle/_ref.  The oracle is pinned against (a) the golden vectors of SURVEY.md §8(c),
(b) 40-digit mpmath evaluations of the reference's closed forms (tests/golden/make_golden.py),
(c) the reference tests' known-answer energies, (d) the Random123 Philox4x32-10 vectors.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libmole_oracle.so")

# enums (oracle_wf.hpp / oracle_mc.hpp)
WF_STO_1S, WF_GAUSSIAN, WF_STO_PRODUCT, WF_H2_HL_STO, WF_H2P_PRODUCT, WF_SLATER_JASTROW, WF_CONSTANT = range(7)
WF_LCAO_1E_2C, WF_LCAO_2E_1C, WF_LCAO_2E_2C, WF_LCAO_SJ = 7, 8, 9, 10
HAM_KINETIC, HAM_IONIC_POT, HAM_ELEC_POT, HAM_IONIC, HAM_ELECTRONIC, HAM_HARMONIC = range(6)
METROP_BOX, METROP_DIFFUSE = 0, 1
OBS_ENERGY, OBS_PGRAD, OBS_WFVALUE, OBS_KINETIC = 1, 2, 4, 8
OPT_SD, OPT_MOMENTUM, OPT_NESTEROV, OPT_LBFGS, OPT_SR = range(5)
BRANCH_SR, BRANCH_SIMPLE = 0, 1
DOM_MOVE, DOM_INIT, DOM_BRANCH, DOM_SEED = range(4)


class WfDesc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("n_elec", C.c_int32), ("n_params", C.c_int32), ("reserved", C.c_int32),
                ("params", C.c_double * 48), ("geom", C.c_double * 40)]


class HamDesc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("n_ions", C.c_int32), ("ion_pos", C.c_double * 24),
                ("ion_charge", C.c_int32 * 8), ("frequency", C.c_double)]


class RunOptions(C.Structure):
    _fields_ = [("metrop_kind", C.c_int32), ("metrop_param", C.c_double), ("observables", C.c_uint32),
                ("quirk_vector_div", C.c_int32), ("nan_reject", C.c_int32)]


def build(force=False):
    """Compile the oracle with the committed Makefile (no-op when the .so is up to date)."""
    if force or not os.path.exists(_LIB_PATH) or any(
            os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(_LIB_PATH)
            for f in ("mole_oracle.cpp", "oracle_rng.hpp", "oracle_wf.hpp", "oracle_mc.hpp")):
        subprocess.check_call(["make", "-C", _HERE], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_ionic_potential.restype = C.c_double
        _lib.orc_electronic_potential.restype = C.c_double
        _lib.orc_mean_fold.restype = C.c_double
        _lib.orc_blocking_error.restype = C.c_double
        _lib.orc_bench_vmc.restype = C.c_double
        _lib.orc_opt_create.restype = C.c_void_p
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def _seed(seed):
    b = bytes(seed)
    assert len(b) == 32
    return (C.c_uint8 * 32)(*b)


# ------------------------------------------------------------------ descriptors
def wf_desc(kind, params=(), geom=(), n_elec=None):
    ne = {WF_STO_1S: 1, WF_GAUSSIAN: 1, WF_STO_PRODUCT: 2, WF_H2_HL_STO: 2, WF_H2P_PRODUCT: 1,
          WF_CONSTANT: 1, WF_LCAO_1E_2C: 1, WF_LCAO_2E_1C: 2, WF_LCAO_2E_2C: 2}.get(kind)
    if kind in (WF_SLATER_JASTROW, WF_LCAO_SJ):
        ne = int(geom[1]) + int(geom[2])
    if n_elec is not None:
        ne = n_elec
    d = WfDesc()
    d.kind, d.n_elec, d.n_params = kind, ne, len(params)
    for i, p in enumerate(params):
        d.params[i] = p
    for i, g in enumerate(geom):
        d.geom[i] = g
    return d


def ham_desc(kind, ion_pos=(), ion_charge=(), frequency=0.0):
    h = HamDesc()
    ion_pos = np.asarray(ion_pos, dtype=np.float64).reshape(-1)
    h.kind, h.n_ions, h.frequency = kind, len(ion_charge), frequency
    for i, x in enumerate(ion_pos):
        h.ion_pos[i] = x
    for i, z in enumerate(ion_charge):
        h.ion_charge[i] = int(z)
    return h


def run_options(metrop_kind, metrop_param, observables=OBS_ENERGY, quirk_vector_div=0, nan_reject=0):
    return RunOptions(metrop_kind, metrop_param, observables, quirk_vector_div, nan_reject)


# ------------------------------------------------------------------ RNG
def philox(ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    lib().orc_philox4x32_10(c, k, o)
    return list(o)


def key_from_seed(seed):
    o = (C.c_uint32 * 2)()
    lib().orc_key_from_seed(_seed(seed), o)
    return list(o)


def derive_seed(master, n):
    o = (C.c_uint8 * 32)()
    lib().orc_derive_seed(_seed(master), C.c_uint32(n), o)
    return bytes(o)


def draw_move(kind, seed, walker, step, domain, elec):
    o = (C.c_double * 4)()
    lib().orc_draw_move(kind, _seed(seed), C.c_uint64(walker), C.c_uint32(step), C.c_uint32(domain),
                        C.c_uint32(elec), o)
    return list(o)


def init_uniform(seed, walker, ne, lo=-1.0, hi=1.0):
    cfg = np.empty((ne, 3))
    lib().orc_init_uniform(_seed(seed), C.c_uint64(walker), ne, C.c_double(lo), C.c_double(hi), _dp(cfg))
    return cfg


def init_normal(seed, walker, ne, sigma=1.0):
    cfg = np.empty((ne, 3))
    lib().orc_init_normal(_seed(seed), C.c_uint64(walker), ne, C.c_double(sigma), _dp(cfg))
    return cfg


# ------------------------------------------------------------------ pointwise
def _chk(rc):
    if rc != 0:
        raise RuntimeError("oracle error code %d" % rc)


def wf_value(d, cfg):
    cfg = np.ascontiguousarray(cfg, dtype=np.float64)
    o = C.c_double()
    _chk(lib().orc_wf_value(C.byref(d), _dp(cfg), C.byref(o)))
    return o.value


def wf_gradient(d, cfg):
    cfg = np.ascontiguousarray(cfg, dtype=np.float64)
    out = np.empty((d.n_elec, 3))
    _chk(lib().orc_wf_gradient(C.byref(d), _dp(cfg), _dp(out)))
    return out


def wf_laplacian(d, cfg):
    cfg = np.ascontiguousarray(cfg, dtype=np.float64)
    o = C.c_double()
    _chk(lib().orc_wf_laplacian(C.byref(d), _dp(cfg), C.byref(o)))
    return o.value


def wf_parameter_gradient(d, cfg):
    cfg = np.ascontiguousarray(cfg, dtype=np.float64)
    out = np.zeros(max(d.n_params, 1))
    _chk(lib().orc_wf_parameter_gradient(C.byref(d), _dp(cfg), _dp(out)))
    return out[:d.n_params]


def ham_act_on(h, d, cfg):
    cfg = np.ascontiguousarray(cfg, dtype=np.float64)
    o = C.c_double()
    _chk(lib().orc_ham_act_on(C.byref(h), C.byref(d), _dp(cfg), C.byref(o)))
    return o.value


def local_energy(h, d, cfg):
    return ham_act_on(h, d, cfg) / wf_value(d, cfg)


def ionic_potential(h, cfg):
    cfg = np.ascontiguousarray(cfg, dtype=np.float64)
    return lib().orc_ionic_potential(C.byref(h), _dp(cfg), cfg.size // 3)


def electronic_potential(cfg):
    cfg = np.ascontiguousarray(cfg, dtype=np.float64)
    return lib().orc_electronic_potential(_dp(cfg), cfg.size // 3)


def eval_batch(d, h, cfgs, want_pgrad=True):
    cfgs = np.ascontiguousarray(cfgs, dtype=np.float64)
    W = cfgs.shape[0]
    n = 3 * d.n_elec
    psi, lap, hpsi = np.empty(W), np.empty(W), np.empty(W)
    grad = np.empty((W, d.n_elec, 3))
    pg = np.zeros((W, max(d.n_params, 1))) if want_pgrad and d.n_params else None
    _chk(lib().orc_eval_batch(C.byref(d), C.byref(h) if h is not None else None, _dp(cfgs), C.c_int64(W),
                              _dp(psi), _dp(grad), _dp(lap), _dp(hpsi) if h is not None else None, _dp(pg)))
    return dict(psi=psi, grad=grad, lap=lap, hpsi=hpsi if h is not None else None,
                pgrad=pg[:, :d.n_params] if pg is not None else None)


def move_state(d, metrop_kind, param, cfg, idx, seed, walker, step):
    cfg = np.array(cfg, dtype=np.float64, copy=True)
    ratio = C.c_double(float("nan"))
    rc = lib().orc_move_state(C.byref(d), metrop_kind, C.c_double(param), _dp(cfg), idx, _seed(seed),
                              C.c_uint64(walker), C.c_uint32(step), C.byref(ratio))
    if rc < 0:
        raise RuntimeError("oracle move_state error")
    return bool(rc), cfg, ratio.value


def ensemble_run(d, h, opts, cfgs, seed, steps, block_size, walker_offset=0, want=("energy",), trace=True):
    """Runner::run for W independent chains.  Returns dict with final cfgs, per-sample series,
    accept bits [W, steps_eff, N_e] and acceptance[W]."""
    cfgs = np.array(cfgs, dtype=np.float64, copy=True)
    W = cfgs.shape[0]
    ne, P = d.n_elec, d.n_params
    ns = (steps // block_size - 1) * block_size
    se = (steps // block_size) * block_size
    energy = np.empty((W, ns)) if (opts.observables & OBS_ENERGY) else None
    wfv = np.empty((W, ns)) if (opts.observables & OBS_WFVALUE) else None
    kin = np.empty((W, ns)) if (opts.observables & OBS_KINETIC) else None
    pg = np.empty((W, ns, max(P, 1))) if (opts.observables & OBS_PGRAD) else None
    acc = np.empty((W, se, ne), dtype=np.uint8) if trace else None
    acceptance = np.empty(W)
    _chk(lib().orc_ensemble_run(C.byref(d), C.byref(h), C.byref(opts), _dp(cfgs), _seed(seed),
                                C.c_uint64(walker_offset), C.c_int64(W), steps, block_size, _dp(energy),
                                _dp(wfv), _dp(kin), _dp(pg),
                                acc.ctypes.data_as(C.POINTER(C.c_uint8)) if trace else None, _dp(acceptance)))
    return dict(cfgs=cfgs, energy=energy, wfvalue=wfv, kinetic=kin, pgrad=pg, accept=acc,
                acceptance=acceptance)


# ------------------------------------------------------------------ statistics / optimizers
def mean_fold(v):
    v = np.ascontiguousarray(v, dtype=np.float64)
    return lib().orc_mean_fold(_dp(v), C.c_int64(v.size))


def blocking_error(v, block_size, mean):
    v = np.ascontiguousarray(v, dtype=np.float64)
    return lib().orc_blocking_error(_dp(v), C.c_int64(v.size), C.c_int64(block_size), C.c_double(mean))


class Optimizer:
    def __init__(self, kind, np_, step, momentum=0.0, history=5, quirk_sr_subtract=0):
        self.np = np_
        self.h = C.c_void_p(lib().orc_opt_create(kind, np_, C.c_double(step), C.c_double(momentum), history,
                                                 quirk_sr_subtract))

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_opt_destroy(self.h)
            self.h = None

    def step(self, pars, energy_avg, wfv, pg, en):
        pars = np.ascontiguousarray(pars, dtype=np.float64)
        wfv = np.ascontiguousarray(wfv, dtype=np.float64).reshape(-1)
        pg = np.ascontiguousarray(pg, dtype=np.float64).reshape(wfv.size, self.np)
        en = np.ascontiguousarray(en, dtype=np.float64).reshape(-1)
        dp = np.empty(self.np)
        rc = lib().orc_opt_step(self.h, _dp(pars), C.c_double(energy_avg), _dp(wfv), _dp(pg), _dp(en),
                                C.c_int64(wfv.size), _dp(dp))
        if rc:
            raise RuntimeError("LinalgError")
        return dp

    def sr_matrix(self, wfv, pg):
        wfv = np.ascontiguousarray(wfv, dtype=np.float64).reshape(-1)
        pg = np.ascontiguousarray(pg, dtype=np.float64).reshape(wfv.size, self.np)
        S = np.empty((self.np, self.np))
        lib().orc_sr_matrix(self.h, _dp(wfv), _dp(pg), C.c_int64(wfv.size), _dp(S))
        return S


def energy_gradient(wfv, pg, en, energy):
    wfv = np.ascontiguousarray(wfv, dtype=np.float64).reshape(-1)
    en = np.ascontiguousarray(en, dtype=np.float64).reshape(-1)
    pg = np.ascontiguousarray(pg, dtype=np.float64).reshape(wfv.size, -1)
    g = np.empty(pg.shape[1])
    lib().orc_energy_gradient(_dp(wfv), _dp(pg), _dp(en), C.c_int64(wfv.size), pg.shape[1], C.c_double(energy),
                              _dp(g))
    return g


def vmc_run_optimization(d, h, opts, opt, master_seed, cfg0, iters, total_samples, block_size, nworkers):
    cfg0 = np.ascontiguousarray(cfg0, dtype=np.float64)
    P = d.n_params
    en, er, ac = np.empty(iters), np.empty(iters), np.empty(iters)
    ph, fp = np.empty((iters, P)), np.empty(P)
    _chk(lib().orc_vmc_run_optimization(C.byref(d), C.byref(h), C.byref(opts), opt.h, _seed(master_seed),
                                        _dp(cfg0), iters, total_samples, block_size, nworkers, _dp(en), _dp(er),
                                        _dp(ac), _dp(ph), _dp(fp)))
    return dict(energies=en, errors=er, acceptance=ac, param_history=ph, params=fp)


# ------------------------------------------------------------------ DMC
def dmc_step(d, h, weights, cfgs, metrop_tau, time_step, e_ref, seed, step):
    weights = np.array(weights, dtype=np.float64, copy=True)
    cfgs = np.array(cfgs, dtype=np.float64, copy=True)
    e, tw = C.c_double(), C.c_double()
    _chk(lib().orc_dmc_step(C.byref(d), C.byref(h), _dp(weights), _dp(cfgs), C.c_int64(weights.size),
                            C.c_double(metrop_tau), C.c_double(time_step), C.c_double(e_ref), _seed(seed),
                            C.c_uint32(step), C.byref(e), C.byref(tw)))
    return e.value, tw.value, weights, cfgs


def branch(kind, ne, weights, cfgs, seed, step):
    weights = np.ascontiguousarray(weights, dtype=np.float64)
    cfgs = np.ascontiguousarray(cfgs, dtype=np.float64)
    wo, co = np.empty_like(weights), np.empty_like(cfgs)
    _chk(lib().orc_branch(kind, ne, _dp(weights), _dp(cfgs), C.c_int64(weights.size), _seed(seed),
                          C.c_uint32(step), _dp(wo), _dp(co)))
    return wo, co


def dmc_diffuse(d, h, weights, cfgs, metrop_tau, e_ref, branch_kind, seed, time_step, num_iterations,
                block_size, num_eq_blocks):
    weights = np.array(weights, dtype=np.float64, copy=True)
    cfgs = np.array(cfgs, dtype=np.float64, copy=True)
    nb = num_iterations // block_size
    en, er = np.empty(max(nb, 1)), np.empty(max(nb, 1))
    se = np.empty(nb * block_size)
    n_out, eref = C.c_int(), C.c_double()
    _chk(lib().orc_dmc_diffuse(C.byref(d), C.byref(h), _dp(weights), _dp(cfgs), C.c_int64(weights.size),
                               C.c_double(metrop_tau), C.c_double(e_ref), branch_kind, _seed(seed),
                               C.c_double(time_step), num_iterations, block_size, num_eq_blocks, _dp(en), _dp(er),
                               C.byref(n_out), _dp(se), C.byref(eref)))
    k = n_out.value
    return dict(energies=en[:k], errors=er[:k], step_energies=se, reference_energy=eref.value,
                weights=weights, cfgs=cfgs)


def count_ops_faithful(d, h, opts, cfg):
    cfg = np.ascontiguousarray(cfg, dtype=np.float64)
    out = (C.c_uint64 * 7)()
    _chk(lib().orc_count_ops_faithful(C.byref(d), C.byref(h), C.byref(opts), _dp(cfg), out))
    return dict(zip(("add", "mul", "div", "sqrt", "exp", "log", "cmp"), list(out)))


def bench_vmc(d, h, opts, cfgs, steps, block_size, seed):
    cfgs = np.ascontiguousarray(cfgs, dtype=np.float64)
    es, th = C.c_double(), C.c_int()
    secs = lib().orc_bench_vmc(C.byref(d), C.byref(h), C.byref(opts), _dp(cfgs), C.c_int64(cfgs.shape[0]), steps,
                               block_size, _seed(seed), C.byref(es), C.byref(th))
    return secs, es.value, th.value


def statfor_block_sizes(n):
    """Block-size schedule of scripts/statfor.rs:59-66."""
    k = lib().orc_statfor_block_sizes(C.c_int64(n), None)
    out = np.empty(k, dtype=np.int32)
    if k:
        lib().orc_statfor_block_sizes(C.c_int64(n), out.ctypes.data_as(C.c_void_p))
    return out


def statfor(series, block_sizes=None):
    """scripts/statfor.rs on one series (mean, variance, correlation, blocking)."""
    x = np.ascontiguousarray(series, dtype=np.float64)
    n = x.size
    bs = statfor_block_sizes(n) if block_sizes is None else np.ascontiguousarray(block_sizes, dtype=np.int32)
    lags = max(min(200, n - 1), 0)
    st, corr, errs = np.empty(5), np.empty(lags), np.empty(bs.size)
    lib().orc_statfor(_dp(x), C.c_int64(n), _dp(st), _dp(corr), C.c_int(bs.size), bs.ctypes.data_as(C.c_void_p), _dp(errs))
    return dict(average=st[0], variance=st[1], tcorr=st[2], n_eff=st[3], sigma=st[4], corr=corr, block_sizes=bs,
                block_errors=errs)
