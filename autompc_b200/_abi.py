"""ctypes binding of libampc_b200.so (include/ampc_b200.h).

The product path fails loudly here when the CUDA library is missing: there is
no CPU fallback anywhere in ``autompc_b200``.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libampc_b200.so")

AMPC_OK, AMPC_ERR_INVALID, AMPC_ERR_UNSUPPORTED, AMPC_ERR_CUDA, AMPC_ERR_NOMEM = 0, -1, -2, -3, -4
ACT_CODES = {"relu": 0, "tanh": 1, "sigmoid": 2, "selu": 3}
PREC_CODES = {"fp32": 0, "bf16": 1, "fp16": 2}
MAX_LAYERS = 5

_dp = C.POINTER(C.c_double)
_fp = C.POINTER(C.c_float)


class MlpDesc(C.Structure):
    _fields_ = [("n_layers", C.c_int32), ("dims", C.POINTER(C.c_int32)), ("W", C.POINTER(_dp)),
                ("b", C.POINTER(_dp)), ("act", C.c_int32), ("xu_mean", _dp), ("xu_std", _dp),
                ("dy_mean", _dp), ("dy_std", _dp)]


class QuadCost(C.Structure):
    _fields_ = [("Q", _dp), ("R", _dp), ("F", _dp), ("goal", _dp), ("umin", _dp), ("umax", _dp), ("goal_term", _dp)]


class MppiCfg(C.Structure):
    _fields_ = [("K", C.c_int32), ("H", C.c_int32), ("nx", C.c_int32), ("nu", C.c_int32),
                ("sigma", C.c_double), ("lmda", C.c_double), ("terminal_mode", C.c_int32),
                ("precision", C.c_int32), ("k_offset", C.c_int32), ("K_global", C.c_int32),
                ("device", C.c_int32)]


class IlqrCfg(C.Structure):
    _fields_ = [("H", C.c_int32), ("nx", C.c_int32), ("nu", C.c_int32), ("dt", C.c_double),
                ("bounded", C.c_int32), ("max_iter", C.c_int32), ("ls_max_iter", C.c_int32),
                ("ls_discount", C.c_double), ("ls_cost_threshold", C.c_double), ("u_threshold", C.c_double),
                ("device", C.c_int32)]


EXPORTS = {
    "ampc_mppi_create": [C.POINTER(C.c_void_p), C.POINTER(MppiCfg), C.POINTER(MlpDesc), C.POINTER(QuadCost)],
    "ampc_mppi_destroy": [C.c_void_p],
    "ampc_mppi_set_act_seq": [C.c_void_p, _dp],
    "ampc_mppi_get_act_seq": [C.c_void_p, _dp],
    "ampc_mppi_solve_host": [C.c_void_p, _dp, _dp, C.c_uint64, C.c_uint64, _dp],
    "ampc_mppi_solve": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p],
    "ampc_mppi_get_costs": [C.c_void_p, _dp, _dp],
    "ampc_mppi_set_box_costs": [C.c_void_p, C.c_int32, _dp, _dp, _dp],
    "ampc_mppi_set_eval_cost": [C.c_void_p, C.POINTER(QuadCost), C.c_int32, _dp, _dp, _dp],
    "ampc_mppi_get_noise": [C.c_void_p, C.c_uint64, C.c_uint64, _fp],
    "ampc_mppi_record_floats": [C.c_void_p],
    "ampc_mppi_rollout_partial": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p],
    "ampc_mppi_merge": [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p],
    "ampc_mppi_mailbox_ipc": [C.c_void_p, C.c_int32, C.c_void_p],
    "ampc_mppi_connect_peers_ipc": [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p],
    "ampc_mppi_connect_peers_local": [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_void_p)],
    "ampc_mppi_solve_fused": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p],
    "ampc_mppi_solve_fused_host": [C.c_void_p, _dp, _dp, C.c_uint64, C.c_uint64, _dp],
    "ampc_mppi_closed_loop_start": [C.c_void_p, C.c_void_p, _dp, C.c_int32, C.c_uint64, C.c_uint64],
    "ampc_mppi_closed_loop_finish": [C.c_void_p, C.c_int32, _dp, _dp, _dp],
    "ampc_mppi_debug_trace": [C.c_void_p, C.POINTER(C.c_uint64), C.c_int32],
    "ampc_mppi_debug_tc_mode": [C.c_void_p],
    "ampc_mlp_create": [C.POINTER(C.c_void_p), C.POINTER(MlpDesc), C.c_int32, C.c_int32, C.c_int32],
    "ampc_mlp_destroy": [C.c_void_p],
    "ampc_mlp_pred_batch": [C.c_void_p, C.c_int32, _dp, _dp, _dp],
    "ampc_mlp_rollout_batch": [C.c_void_p, C.c_int32, C.c_int32, _dp, _dp, _dp],
    "ampc_mlp_pred_diff_batch": [C.c_void_p, C.c_int32, _dp, _dp, _dp, _dp, _dp],
    "ampc_mlp_nmpc_constraint": [C.c_void_p, C.c_int32, _dp, _dp],
    "ampc_mlp_nmpc_jacobian": [C.c_void_p, C.c_int32, _dp, _dp],
    "ampc_mlp_debug_last_kernel_ms": [C.c_void_p, C.POINTER(C.c_float)],
    "ampc_linear_create": [C.POINTER(C.c_void_p), C.c_int32, C.c_int32, _dp, _dp, C.c_int32],
    "ampc_linear_destroy": [C.c_void_p],
    "ampc_linear_pred_batch": [C.c_void_p, C.c_int32, _dp, _dp, _dp],
    "ampc_ilqr_create": [C.POINTER(C.c_void_p), C.POINTER(IlqrCfg), C.POINTER(MlpDesc), C.POINTER(QuadCost)],
    "ampc_ilqr_destroy": [C.c_void_p],
    "ampc_ilqr_solve_host": [C.c_void_p, _dp, _dp, _dp, _dp, _dp, _dp, C.POINTER(C.c_int32), C.POINTER(C.c_int32)],
    "ampc_ilqr_launch": [C.c_void_p, C.c_void_p],
    "ampc_ilqr_debug_profile": [C.c_void_p, C.POINTER(C.c_uint64)],
    "ampc_last_error": [],
    "ampc_version": [],
    "ampc_launch_count": [],
}

_lib = None


def lib():
    """Loads the shared library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "libampc_b200.so is missing at %s -- build it with `python -m autompc_b200.build` "
                "(autompc_b200 has no CPU fallback)" % LIB_PATH)
        l = C.CDLL(LIB_PATH)
        for name, argtypes in EXPORTS.items():
            fn = getattr(l, name)
            fn.argtypes = argtypes
            fn.restype = C.c_int
        l.ampc_last_error.restype = C.c_char_p
        l.ampc_version.restype = C.c_char_p
        l.ampc_launch_count.restype = C.c_uint64
        _lib = l
    return _lib


def check(rc):
    """Error mapping of SURVEY.md 8(b): bad shape / unsupported -> ValueError, CUDA -> RuntimeError."""
    if rc == AMPC_OK:
        return
    msg = lib().ampc_last_error().decode("utf-8", "replace")
    if rc in (AMPC_ERR_INVALID, AMPC_ERR_UNSUPPORTED):
        raise ValueError(msg)
    if rc == AMPC_ERR_NOMEM:
        raise MemoryError(msg)
    raise RuntimeError(msg)


def launch_count():
    return int(lib().ampc_launch_count())


def dptr(a):
    return a.ctypes.data_as(_dp)


def f64(a, shape=None):
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float64))
    if shape is not None and a.shape != tuple(shape):
        raise ValueError("expected shape %s, got %s" % (tuple(shape), a.shape))
    return a


class MlpDescHolder:
    """Builds an ``ampc_mlp_desc`` and keeps the NumPy buffers alive."""

    def __init__(self, weights):
        w = weights
        n = len(w.W)
        self.keep = [f64(x) for x in w.W] + [f64(x) for x in w.b]
        self.dims = (C.c_int32 * (n + 1))(*w.dims)
        self.Wp = (_dp * n)(*[dptr(x) for x in self.keep[:n]])
        self.bp = (_dp * n)(*[dptr(x) for x in self.keep[n:]])
        self.norm = [f64(w.xu_mean), f64(w.xu_std), f64(w.dy_mean), f64(w.dy_std)]
        self.desc = MlpDesc(n, self.dims, self.Wp, self.bp, ACT_CODES[w.act], *[dptr(x) for x in self.norm])


class QuadCostHolder:
    def __init__(self, Q, R, F, goal, umin, umax, nx, nu, goal_term=None):
        self.keep = [f64(Q, (nx, nx)), f64(R, (nu, nu)), f64(F, (nx, nx)), f64(goal, (nx,)),
                     f64(umin, (nu,)), f64(umax, (nu,))]
        self.desc = QuadCost(*[dptr(x) for x in self.keep])
        if goal_term is not None:
            self.keep.append(f64(goal_term, (nx,)))
            self.desc.goal_term = dptr(self.keep[-1])
