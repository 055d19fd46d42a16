"""ctypes binding of liblinfa_b200.so (include/linfa_b200.h).

The library is the product; this file only declares its C ABI.  It fails loudly when the shared
library is missing or cannot be loaded -- there is no CPU fallback anywhere in this package.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "liblinfa_b200.so")

# status codes (include/linfa_b200.h)
OK, NOT_POSITIVE_DEFINITE, NOT_THIN, NOT_SQUARE, EMPTY_MATRIX, WRONG_ROWS, NON_INVERTIBLE, INVALID_ARGUMENT, UNSUPPORTED = range(9)
ERR_CUDA, ERR_ALLOC, ERR_NCCL = 100, 101, 102
UPPER, LOWER = 0, 1

_i64, _int, _vp, _dbl, _flt = C.c_int64, C.c_int, C.c_void_p, C.c_double, C.c_float

# name -> argtypes (restype is int unless listed in _RESTYPES).  "T" is replaced per dtype suffix.
_VIEW = [_vp, _i64, _i64, _i64, _i64]
SIGNATURES = {
    "lfb_create": [C.POINTER(_vp), _int],
    "lfb_destroy": [_vp],
    "lfb_last_error": [_vp],
    "lfb_set_stream": [_vp, _vp],
    "lfb_use_own_stream": [_vp],
    "lfb_synchronize": [_vp],
    "lfb_version": [],
    "lfb_launch_count": [_vp],
    "lfb_set_option": [_vp, C.c_char_p, _i64],
    "lfb_microbench_fp64": [_vp, _int, C.POINTER(_dbl)],
    "lfb_microbench_kernel": [_vp, C.c_char_p, _i64, _int, C.POINTER(_dbl)],
    "lfb_debug_panel_phases": [_vp, C.POINTER(C.c_longlong)],
    "lfb_profile_begin": [_vp],
    "lfb_profile_end": [_vp, C.POINTER(_dbl), C.POINTER(_dbl), C.POINTER(_i64)],
}
_TYPED = {
    "lfb_qr": [_vp] + _VIEW + [_vp],
    "lfb_qr_tsqr": [_vp] + _VIEW + [_vp],
    "lfb_assemble_q": [_vp] + _VIEW + [_i64, _vp, _vp, _i64, _i64],
    "lfb_qt_mul": [_vp] + _VIEW + [_vp, _vp, _i64, _i64, _i64],
    "lfb_cholesky": [_vp] + _VIEW + [_int, C.POINTER(_i64)],
    "lfb_solve_triangular": [_vp] + _VIEW + _VIEW + [_int, _vp],
    "lfb_triangular_inplace": [_vp] + _VIEW + [_int],
    "lfb_sym_tridiagonal": [_vp] + _VIEW + [_vp],
    "lfb_bidiagonal": [_vp] + _VIEW + [_vp, _vp],
    "lfb_eigh": [_vp] + _VIEW + [_vp, _vp, _i64, _i64],
    "lfb_least_squares": [_vp] + _VIEW + _VIEW + [_vp, _i64, _i64],
    "lfb_qr_solve": [_vp] + _VIEW + [_vp] + _VIEW + [_vp, _i64, _i64],
    "lfb_qr_solve_tr": [_vp] + _VIEW + [_vp] + _VIEW + [_vp, _i64, _i64],
    "lfb_solvec": [_vp] + _VIEW + [_int] + _VIEW + [C.POINTER(_i64)],
    "lfb_invc": [_vp] + _VIEW + [_vp, _i64, _i64, C.POINTER(_i64)],
    "lfb_svd": [_vp] + _VIEW + [_vp, _vp, _i64, _i64, _vp, _i64, _i64],
    "lfb_orthonormalize": [_vp] + _VIEW + [_vp, _i64, _i64, C.POINTER(_i64)],
    "lfb_apply_constraints": [_vp] + _VIEW + [_vp, _i64, _i64, _i64] + _VIEW,
    "lfb_sorted_eig": [_vp, _vp, _i64, _i64, _i64, _vp, _i64, _i64, _i64, _int, _vp, _vp, _i64, _i64],
    "lfb_qr_batched": [_vp, _vp, _i64, _i64, _i64, _vp],
    "lfb_cholesky_batched": [_vp, _vp, _i64, _i64, _int, C.POINTER(_i64), C.POINTER(_i64)],
    "lfb_qr_dev": [_vp, _vp, _i64, _i64, _i64, _vp],
    "lfb_qr_tsqr_dev": [_vp, _vp, _i64, _i64, _i64, _vp],
    "lfb_cholesky_dev": [_vp, _vp, _i64, _i64, _int, _vp],
}
for _n, _a in _TYPED.items():
    SIGNATURES[_n + "_f64"] = _a
    SIGNATURES[_n + "_f32"] = _a
SIGNATURES.update({
    "lfb_assemble_q_dev_f64": [_vp, _vp, _i64, _i64, _i64, _i64, _vp, _vp, _i64],
    "lfb_sym_tridiagonal_dev_f64": [_vp, _vp, _i64, _i64, _vp],
    "lfb_bidiagonal_dev_f64": [_vp, _vp, _i64, _i64, _i64, _vp, _vp],
    "lfb_eigh_dev_f64": [_vp, _vp, _i64, _i64, _vp, _vp, _i64],
    "lfb_svd_dev_f64": [_vp, _vp, _i64, _i64, _i64, _vp, _vp, _i64, _vp, _i64],
    "lfb_qr_batched_dev_f32": [_vp, _vp, _i64, _i64, _i64, _vp],
    "lfb_cholesky_batched_dev_f32": [_vp, _vp, _i64, _i64, _int, _vp],
    "lfb_tsqr_local_r_dev_f64": [_vp, _vp, _i64, _i64, _i64, _vp, _i64],
    "lfb_tsqr_explicit_q_dev_f64": [_vp, _vp, _i64, _i64, _i64, _vp, _i64],
    "lfb_tsqr_apply_q_dev_f64": [_vp, _vp, _i64, _i64, _i64, _vp, _i64],
    "lfb_tsqr_leaf_dev_f64": [_vp, _vp, _i64, _i64, _i64, _vp, _i64, _vp, _i64, C.POINTER(_int)],
    "lfb_hh_reconstruct_top_dev_f64": [_vp, _vp, _i64, _i64, _vp, _i64, _vp, _i64, _vp],
    "lfb_hh_reconstruct_rows_dev_f64": [_vp, _vp, _i64, _i64, _i64, _vp, _i64],
    "lfb_orthonormalize_dev_f64": [_vp, _vp, _i64, _i64, _i64, _vp, _i64, _vp],
    "lfb_sorted_eig_dev_f64": [_vp, _vp, _i64, _vp, _i64, _i64, _i64, _int, _vp, _vp, _i64],
    "lfb_apply_constraints_dev_f64": [_vp, _vp, _i64, _i64, _i64, _vp, _i64, _i64, _vp, _i64],
    "lfb_gemm_dev_f64": [_vp, _int, _int, _i64, _i64, _i64, _dbl, _vp, _i64, _vp, _i64, _dbl, _vp, _i64],
    "lfb_gemm_dev_f32": [_vp, _int, _int, _i64, _i64, _i64, _flt, _vp, _i64, _vp, _i64, _flt, _vp, _i64],
})
_PP = C.POINTER(_vp)          # array of device pointers
_PI = C.POINTER(_i64)         # array of int64
SIGNATURES.update({
    "lfb_create_multi": [C.POINTER(_vp), C.POINTER(_int), _int],
    "lfb_destroy_multi": [_vp],
    "lfb_multi_last_error": [_vp],
    "lfb_multi_device_count": [_vp],
    "lfb_multi_nccl_ranks": [_vp],
    "lfb_multi_nccl_version": [_vp],
    "lfb_multi_handle": [_vp, _int],
    "lfb_multi_set_option": [_vp, C.c_char_p, _i64],
    "lfb_multi_launch_count": [_vp],
    "lfb_multi_synchronize": [_vp],
    "lfb_multi_time_begin": [_vp],
    "lfb_multi_time_end": [_vp, C.POINTER(_dbl)],
    "lfb_qr_tsqr_multi_f64": [_vp] + _VIEW + [_vp],
    "lfb_qr_tsqr_multi_f32": [_vp] + _VIEW + [_vp],
    "lfb_tsqr_r_multi_f64": [_vp] + _VIEW + [_vp, _i64, _i64],
    "lfb_tsqr_r_multi_f32": [_vp] + _VIEW + [_vp, _i64, _i64],
    "lfb_qr_tsqr_multi_dev_f64": [_vp, _PP, _PI, _i64, _PI, _PP, _PP],
    "lfb_tsqr_r_multi_dev_f64": [_vp, _PP, _PI, _i64, _PI, _PP],
    "lfb_qr_batched_multi_f32": [_vp, _vp, _i64, _i64, _i64, _vp],
    "lfb_qr_batched_multi_f64": [_vp, _vp, _i64, _i64, _i64, _vp],
    "lfb_cholesky_batched_multi_f32": [_vp, _vp, _i64, _i64, _int, C.POINTER(_i64), C.POINTER(_i64)],
    "lfb_cholesky_batched_multi_f64": [_vp, _vp, _i64, _i64, _int, C.POINTER(_i64), C.POINTER(_i64)],
    "lfb_qr_batched_multi_dev_f32": [_vp, _PP, _PI, _i64, _i64, _PP],
})
_RESTYPES = {"lfb_last_error": C.c_char_p, "lfb_version": C.c_char_p, "lfb_launch_count": _i64,
             "lfb_multi_last_error": C.c_char_p, "lfb_multi_handle": _vp, "lfb_multi_launch_count": _i64}

_lib = None


def load() -> C.CDLL:
    """Load the shared library and declare every prototype.  Raises if the .so is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m linfa_linalg_b200.build` "
            "(or __graft_entry__.build()).  linfa_linalg_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, args in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.argtypes = args
        fn.restype = _RESTYPES.get(name, _int)
    _lib = lib
    return lib
