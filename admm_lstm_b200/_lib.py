"""ctypes binding of the C ABI in include/admm_lstm_b200.h (the stub INTEGRATION.md describes).

The shared library is built in-tree by admm_lstm_b200/build.py.  There is deliberately no Python or
CPU implementation behind these entry points: if the library is missing it is an error."""
from __future__ import annotations

import ctypes as C
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libadmm_lstm_b200.so")

ADMM_MAX_O = 16
ADMM_MAX_CAND = 32
ADMM_N_METRICS = 8
ADMM_EST_CAND = 64
ADMM_FK_SLOTS = 72
VARIANT_ADMM, VARIANT_NO_DUAL_Y = 0, 1
SRC_X, SRC_H = 0, 1
TC_WEIGHTS, TC_INPUTS, TC_STATE = 1, 2, 4

fp = C.POINTER(C.c_float)
dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int32)


class Hyper(C.Structure):
    _fields_ = [("rho", C.c_float * 7), ("beta_x", C.c_float * 4), ("beta_h", C.c_float * 4), ("beta_wy", C.c_float)]


class Problem(C.Structure):
    _fields_ = [
        ("n", C.c_int64), ("n_global", C.c_int64), ("ldn", C.c_int64),
        ("T", C.c_int32), ("D", C.c_int32), ("H", C.c_int32), ("O", C.c_int32),
        ("variant", C.c_int32), ("with_dual_y", C.c_int32),
        ("hp", Hyper),
        ("x", C.c_void_p), ("y", C.c_void_p),
        ("gate", C.c_void_p * 6), ("dual", C.c_void_p * 5), ("dual_h", C.c_void_p),
        ("a", C.c_void_p), ("dual_y", C.c_void_p),
        ("wx", C.c_void_p), ("wh", C.c_void_p), ("wy", C.c_void_p),
        ("tc_ws", C.c_void_p), ("tc_ws_bytes", C.c_int64),
        ("zstore", C.c_void_p), ("wx_prev", C.c_void_p), ("z_valid", C.c_int32), ("reserved_", C.c_int32),
    ]


class ProbePlan(C.Structure):
    _fields_ = [("k0", C.c_int32 * 4), ("ncand", C.c_int32), ("proof", C.c_int32), ("moments", C.c_int32),
                ("order", C.c_int32)]


class LHyper(C.Structure):
    _fields_ = [("rho_s", C.c_float), ("rho_p", C.c_float), ("rho9", C.c_float), ("rho10", C.c_float),
                ("rho11", C.c_float), ("lam_w", C.c_float), ("lam_u", C.c_float), ("lam_y", C.c_float),
                ("n_norm", C.c_float)]


class LProblem(C.Structure):
    """admm_l_problem (ADMM-LSTM-L): `base` first, so a pointer to it is also a valid admm_problem*."""
    _fields_ = [("base", Problem), ("z", C.c_void_p * 4), ("lam_s", C.c_void_p * 4), ("lam_p", C.c_void_p * 4),
                ("lam9", C.c_void_p), ("lam10", C.c_void_p), ("hp", LHyper)]


PP = C.POINTER(Problem)
LPP = C.POINTER(LProblem)
PLAN = C.POINTER(ProbePlan)
vp = C.c_void_p

# name -> (restype, argtypes); must list every function declared in include/admm_lstm_b200.h
SIGNATURES = {
    "admm_last_error": (C.c_char_p, []),
    "admm_abi_version": (C.c_int, []),
    "admm_sizeof_problem": (C.c_int, []),
    "admm_device_ok": (C.c_int, []),
    "admm_forward_t": (C.c_int, [PP, C.c_int, vp]),
    "admm_predict": (C.c_int, [PP, vp, vp, vp]),
    "admm_wy_grad": (C.c_int, [PP, vp, vp]),
    "admm_wy_apply": (C.c_int, [PP, vp, vp]),
    "admm_weight_begin": (C.c_int, [PP, C.c_int, vp]),
    "admm_weight_grad": (C.c_int, [PP, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp]),
    "admm_weight_finish_grad": (C.c_int, [PP, C.c_int, vp, vp, vp, vp]),
    "admm_weight_probe": (C.c_int, [PP, C.c_int, C.c_int, C.c_int, vp, vp, PLAN, vp, vp, vp, vp]),
    "admm_weight_select": (C.c_int, [PP, C.c_int, vp, vp, vp, PLAN, C.c_int, vp, vp, vp]),
    "admm_weight_apply": (C.c_int, [PP, C.c_int, vp, vp, vp]),
    "admm_sweep_t": (C.c_int, [PP, C.c_int, vp, vp]),
    "admm_last_probe": (C.c_int, [PP, vp, vp]),
    "admm_last_select": (C.c_int, [PP, vp, vp, vp]),
    "admm_last_apply": (C.c_int, [PP, vp, vp, vp]),
    "admm_tc_workspace_bytes": (C.c_int64, [PP]),
    "admm_tc_refresh": (C.c_int, [PP, C.c_int, vp]),
    "admm_debug_preact": (C.c_int, [PP, C.c_int, vp, C.c_int, vp]),
    "admm_tc_overflow": (C.c_int, [PP, C.c_int, vp]),
    "admm_load_inputs": (C.c_int, [PP, vp, vp, vp]),
    "admm_launch_count": (C.c_int64, [C.c_int]),
    "admm_kernel_timing": (C.c_int, [C.c_int]),
    "admm_kernel_timing_report": (C.c_int64, [C.c_char_p, C.c_int64]),
    # ADMM-LSTM-L
    "admm_l_sizeof_problem": (C.c_int, []),
    "admm_l_forward_t": (C.c_int, [LPP, C.c_int, vp, vp, vp]),
    "admm_l_output": (C.c_int, [LPP, vp]),
    "admm_l_sums": (C.c_int, [LPP, C.c_int, C.c_int, vp, vp, vp, vp]),
    "admm_l_gram_xx": (C.c_int, [LPP, vp, vp]),
    "admm_l_sums_last": (C.c_int, [LPP, vp, vp, vp]),
    "admm_l_sweep_gates": (C.c_int, [LPP, C.c_int, vp, vp, vp, vp]),
    "admm_l_sweep_cell": (C.c_int, [LPP, C.c_int, vp, vp, vp, vp, vp]),
    "admm_l_last": (C.c_int, [LPP, vp, vp, vp, vp, vp]),
}

_lib = None


class AdmmLibraryError(RuntimeError):
    pass


def load():
    """Load libadmm_lstm_b200.so; raises (never falls back) when it is absent or incomplete."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise AdmmLibraryError(
            f"{LIB_PATH} is missing: build it with `python -m admm_lstm_b200.build` (needs nvcc). "
            "admm_lstm_b200 has no CPU/PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)      # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.admm_abi_version() != 1:
        raise AdmmLibraryError("ABI version mismatch between _lib.py and the shared library")
    if lib.admm_sizeof_problem() != C.sizeof(Problem):
        raise AdmmLibraryError("admm_problem layout mismatch between _lib.py and the shared library")
    if lib.admm_l_sizeof_problem() != C.sizeof(LProblem):
        raise AdmmLibraryError("admm_l_problem layout mismatch between _lib.py and the shared library")
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().admm_last_error().decode("utf-8", "replace")
        raise AdmmLibraryError(f"{what or 'admm call'} failed (code {rc}): {msg}")
