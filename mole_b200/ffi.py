"""ctypes binding of libmole_b200.so (include/mole_b200.h).

The shared library is built in-tree by mole_b200/csrc/build.sh (see __graft_entry__.build()).
There is no fallback: if the library is missing the import fails loudly, and on a machine
without a B200 every device entry point returns MOLE_ERR_NO_DEVICE, surfaced as MoleError.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# MOLE_B200_LIB: a variant build of the same library (tools/ab_sj.py times several against each other)
LIB_PATH = os.environ.get("MOLE_B200_LIB") or os.path.join(_HERE, "libmole_b200.so")

# ---- status codes (include/mole_b200.h) ----
OK = 0
ERR_LINALG, ERR_SHAPE, ERR_FUNC, ERR_OPERATOR_VALUE_ACCESS, ERR_DATA_ACCESS, ERR_EMPTY_CACHE = 1, 2, 3, 4, 5, 6
ERR_CUDA, ERR_NCCL, ERR_INVALID_ARG, ERR_NO_DEVICE, ERR_ASSERT = 100, 101, 102, 103, 104
_ERR_NAMES = {1: "LinalgError", 2: "ShapeError", 3: "FuncError", 4: "OperatorValueAccessError",
              5: "DataAccessError", 6: "EmptyCacheError", 100: "CudaError", 101: "NcclError",
              102: "InvalidArgument", 103: "NoDevice", 104: "AssertionFailed"}

WF_STO_1S, WF_GAUSSIAN, WF_STO_PRODUCT, WF_H2_HL_STO, WF_H2P_PRODUCT, WF_SLATER_JASTROW, WF_CONSTANT = range(7)
WF_LCAO_1E_2C, WF_LCAO_2E_1C, WF_LCAO_2E_2C, WF_LCAO_SJ = 7, 8, 9, 10
WF_MAX_PARAMS, WF_MAX_GEOM = 48, 40
OP_KINETIC, OP_IONIC_POT, OP_ELEC_POT, OP_IONIC, OP_ELECTRONIC, OP_HARMONIC = range(6)
METROP_BOX, METROP_DIFFUSE = 0, 1
OBS_ENERGY, OBS_PGRAD, OBS_WFVALUE, OBS_KINETIC = 1, 2, 4, 8
COMPAT_VECTOR_DIV, COMPAT_SR_SUBTRACT, COMPAT_NAN_ACCEPT = 1, 2, 4
OPT_SD, OPT_MOMENTUM, OPT_NESTEROV, OPT_LBFGS, OPT_SR = range(5)
BRANCH_SR, BRANCH_SIMPLE = 0, 1
VMC_RESTART_EACH_ITER = 1
ACC_MAX_PARAMS = 8
NCCL_UNIQUE_ID_BYTES = 128


class MoleError(RuntimeError):
    """Mirror of `errors::Error` (src/errors/src/lib.rs:8-15) plus the CUDA/NCCL/argument codes."""

    def __init__(self, code, msg=""):
        self.code = code
        self.kind = _ERR_NAMES.get(code, "Error%d" % code)
        super().__init__("%s (%d): %s" % (self.kind, code, msg))


class WfDesc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("n_elec", C.c_int32), ("n_params", C.c_int32), ("reserved", C.c_int32),
                ("params", C.c_double * 48), ("geom", C.c_double * 40)]


class OpDesc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("n_ions", C.c_int32), ("ion_pos", C.c_double * 24),
                ("ion_charge", C.c_int32 * 8), ("frequency", C.c_double)]


class SweepArgs(C.Structure):
    _fields_ = [("n_sweeps", C.c_int32), ("n_discard", C.c_int32), ("block_size", C.c_int32),
                ("observables", C.c_uint32), ("compat", C.c_uint32), ("flags", C.c_uint32),
                ("energy_trace", C.c_void_p), ("wfvalue_trace", C.c_void_p), ("kinetic_trace", C.c_void_p),
                ("pgrad_trace", C.c_void_p), ("accept_trace", C.c_void_p)]


class SeriesStats(C.Structure):
    _fields_ = [("average", C.c_double), ("variance", C.c_double), ("tcorr", C.c_double), ("n_eff", C.c_double),
                ("sigma", C.c_double)]


class BlockLog(C.Structure):
    _fields_ = [("block_nr", C.c_int32), ("block_size", C.c_int32), ("n_samples", C.c_double),
                ("block_energy", C.c_double), ("running_energy", C.c_double), ("block_kinetic", C.c_double),
                ("block_wfvalue", C.c_double), ("acceptance", C.c_double)]


class EnsHealth(C.Structure):
    _fields_ = [("nonfinite_samples", C.c_int64), ("nonfinite_dmc_walkers", C.c_int64), ("reserved", C.c_int64 * 2)]


LOG_FN = C.CFUNCTYPE(None, C.c_void_p, C.POINTER(BlockLog))
SWEEP_KEEP_SERIES, SWEEP_APPEND_SERIES = 1, 2
SERIES_MAX_LAG = 200


class AccHost(C.Structure):
    _fields_ = [("n_samples", C.c_double), ("sum_e", C.c_double), ("sum_e2", C.c_double), ("sum_b", C.c_double),
                ("sum_b2", C.c_double), ("n_blocks", C.c_double), ("n_accept", C.c_double), ("n_moves", C.c_double),
                ("sum_t", C.c_double), ("sum_psi", C.c_double), ("sum_o", C.c_double * 8), ("sum_oe", C.c_double * 8),
                ("sum_oo", C.c_double * 36), ("n_params", C.c_int32), ("reserved", C.c_int32)]

    N_DOUBLES = 62

    def oo(self, k, l):
        P = self.n_params
        k, l = min(k, l), max(k, l)
        return self.sum_oo[k * P - k * (k - 1) // 2 + (l - k)]


# every symbol include/mole_b200.h declares (checked by tests/test_abi.py)
SYMBOLS = [
    "mole_ctx_create", "mole_ctx_destroy", "mole_ctx_synchronize", "mole_last_error_string", "mole_ctx_stream",
    "mole_version", "mole_wf_create", "mole_wf_destroy", "mole_wf_num_electrons", "mole_wf_num_parameters",
    "mole_wf_get_parameters", "mole_wf_update_parameters", "mole_wf_set_parameters", "mole_wf_value",
    "mole_wf_gradient", "mole_wf_laplacian", "mole_wf_parameter_gradient", "mole_op_create", "mole_op_destroy",
    "mole_op_act_on", "mole_ensemble_create", "mole_ensemble_destroy", "mole_ensemble_num_walkers",
    "mole_ensemble_init_uniform", "mole_ensemble_init_normal", "mole_ensemble_set_configs",
    "mole_ensemble_set_configs_broadcast", "mole_ensemble_get_configs", "mole_ensemble_set_weights",
    "mole_ensemble_get_weights", "mole_ensemble_snapshot", "mole_ensemble_restore", "mole_ensemble_reseed",
    "mole_ensemble_set_step", "mole_ensemble_get_step", "mole_derive_seed", "mole_eval_vgl",
    "mole_metropolis_create", "mole_metropolis_destroy", "mole_metropolis_set_compat", "mole_sweep", "mole_acc_reset", "mole_acc_get",
    "mole_acc_allreduce", "mole_acc_device_ptr", "mole_acc_finalize", "mole_comm_get_unique_id", "mole_comm_init",
    "mole_comm_destroy", "mole_opt_create", "mole_opt_destroy", "mole_opt_step", "mole_opt_sr_matrix",
    "mole_runner_run", "mole_vmc_run_optimization", "mole_dmc_step", "mole_branch", "mole_branch_sources",
    "mole_dmc_diffuse", "mole_bench_fp64_peak", "mole_ctx_launch_count", "mole_math_probe",
    "mole_series_length", "mole_series_clear", "mole_series_block_sizes", "mole_series_analyze", "mole_series_get",
    "mole_series_write_text", "mole_runner_run_logged", "mole_ensemble_save", "mole_ensemble_load", "mole_dmc_block",
    "mole_ensemble_health", "mole_opt_set_sr_regularization", "mole_gram_get", "mole_gram_allreduce",
    "mole_gram_device_ptr", "mole_gram_select", "mole_gram_finalize", "mole_opt_step_gram", "mole_opt_sr_matrix_gram", "mole_bench_gram", "mole_bench_dmma_peak", "mole_dmc_block_select", "mole_dmc_island_imbalance", "mole_rebalance", "mole_rebalance_plan",
]

_lib = None


def lib():
    """Load libmole_b200.so; raises if the CUDA extension has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("mole_b200: %s is missing - run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(or mole_b200/csrc/build.sh); there is no CPU fallback" % LIB_PATH)
        _lib = C.CDLL(LIB_PATH)
        _lib.mole_last_error_string.restype = C.c_char_p
        _lib.mole_last_error_string.argtypes = [C.c_void_p]
        for name in SYMBOLS:
            try:
                fn = getattr(_lib, name)
            except AttributeError:
                # tools/ab_sj.py times libraries built from older commits against the current one (MOLE_B200_LIB);
                # anywhere else a missing export is a broken build and must fail loudly
                if os.environ.get("MOLE_B200_AB_OLD_LIB") == "1" and os.environ.get("MOLE_B200_LIB"):
                    continue
                raise
            if name != "mole_last_error_string":
                fn.restype = C.c_int32
    return _lib


def check(rc, ctx=None):
    if rc != OK:
        msg = lib().mole_last_error_string(ctx)
        raise MoleError(rc, msg.decode() if msg else "")


def seed32(seed):
    """The reference's `[u8; 32]` seeds: accepts bytes/list of 32 ints, or one int byte to repeat."""
    if isinstance(seed, int):
        seed = [seed] * 32
    b = bytes(seed)
    if len(b) != 32:
        raise MoleError(ERR_INVALID_ARG, "seed must be 32 bytes")
    return (C.c_uint8 * 32)(*b)
