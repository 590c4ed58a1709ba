"""ctypes binding of libgpmpc.so (include/gpmpc.h).  There is NO fallback: if the shared library is
missing or a CUDA device is absent, every compute entry point raises."""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_double, c_int, c_longlong, c_void_p, POINTER

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgpmpc.so")

# every symbol include/gpmpc.h declares: name -> (restype, argtypes)
_P = c_void_p
SIGNATURES = {
    "gpmpc_version": (c_int, []),
    "gpmpc_create": (c_int, [c_int, c_int, c_int, POINTER(c_void_p)]),
    "gpmpc_destroy": (c_int, [_P]),
    "gpmpc_last_error": (c_char_p, [_P]),
    "gpmpc_set_stream": (c_int, [_P, _P]),
    "gpmpc_synchronize": (c_int, [_P]),
    "gpmpc_set_option": (c_int, [_P, c_char_p, c_int]),
    "gpmpc_num_train": (c_int, [_P]),
    "gpmpc_fit": (c_int, [_P, c_int, _P, _P, _P, _P, _P]),
    "gpmpc_refit_output": (c_int, [_P, c_int, _P, _P, c_double, c_double]),
    "gpmpc_append_point": (c_int, [_P, _P, _P]),
    "gpmpc_set_propagation_hypers": (c_int, [_P, _P, _P]),
    "gpmpc_get_matrix": (c_int, [_P, c_int, c_int, _P]),
    "gpmpc_kernel_matrix": (c_int, [_P, c_int, c_int, _P, _P]),
    "gpmpc_predict": (c_int, [_P, c_int, c_int, _P, _P, _P, c_int]),
    "gpmpc_kernel_matrix_ex": (c_int, [_P, c_int, c_int, _P, _P, _P]),
    "gpmpc_predict_ex": (c_int, [_P, c_int, c_int, _P, _P, _P, _P, _P, c_int]),
    "gpmpc_marginal_likelihood": (c_int, [_P, c_int, _P, _P, _P]),
    "gpmpc_moment_match": (c_int, [_P, c_int, _P, _P, c_int, _P, _P]),
    "gpmpc_moment_match_cov": (c_int, [_P, c_int, _P, _P, _P, _P]),
    "gpmpc_moment_match_raw": (c_int, [_P, c_int, c_int, _P, _P, _P, _P, _P, _P, c_double, _P, _P, _P, _P]),
    "gpmpc_covariance_raw": (c_int, [_P, c_int, c_int, _P, _P, _P, _P, _P, c_double, c_double, _P, _P,
                                     c_double, c_double, c_int, _P]),
    "gpmpc_rollout_full": (c_int, [_P, c_int, c_int, _P, _P, _P, _P]),
    "gpmpc_rollout_full_vjp": (c_int, [_P, c_int, c_int, _P, _P, _P, _P]),
    "gpmpc_rollout_cost_grad_full": (c_int, [_P, c_int, c_int] + [_P] * 13),
    "gpmpc_rollout": (c_int, [_P, c_int, c_int, _P, _P, _P, _P]),
    "gpmpc_rollout_vjp": (c_int, [_P, c_int, c_int, _P, _P, _P, _P]),
    "gpmpc_rollout_cost_grad": (c_int, [_P, c_int, c_int] + [_P] * 13),
    "gpmpc_split_export": (c_int, [_P, _P]),
    "gpmpc_split_connect": (c_int, [_P, c_int, c_int, _P]),
    "gpmpc_split_connect_local": (c_int, [_P, c_int, c_int, POINTER(c_void_p)]),
    "gpmpc_split_disconnect": (c_int, [_P]),
    "gpmpc_split_last_exchange_us": (c_int, [_P, POINTER(c_double), POINTER(c_double)]),
    "gpmpc_launch_count": (c_longlong, [_P]),
    "gpmpc_last_pair_kernel_ms": (c_int, [_P, POINTER(c_double), POINTER(c_longlong)]),
    "gpmpc_set_pair_timing": (c_int, [_P, c_int]),
    "gpmpc_measure_fp64_peak": (c_int, [_P, POINTER(c_double), POINTER(c_double)]),
}

MAT_KF, MAT_KY, MAT_KY_INV, MAT_BETA = 0, 1, 2, 3

_lib = None


class GpmpcError(RuntimeError):
    pass


def load():
    """Load libgpmpc.so (building nothing).  Raises GpmpcError if the extension has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GpmpcError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                         "(the CUDA extension is mandatory; there is no CPU fallback)")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(handle, rc, what=""):
    if rc != 0:
        msg = load().gpmpc_last_error(handle)
        raise GpmpcError(f"{what} failed with code {rc}: {msg.decode() if msg else ''}")
