"""CPU-only: the C-ABI library builds, loads, and exports every symbol include/linfa_b200.h declares.
No compute call is made here (there is no GPU in the build container)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "linfa_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lfb_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_both_scalar_families():
    names = _declared()
    for base in ("lfb_qr", "lfb_assemble_q", "lfb_qt_mul", "lfb_cholesky", "lfb_solve_triangular",
                 "lfb_sym_tridiagonal", "lfb_bidiagonal", "lfb_eigh", "lfb_svd",
                 "lfb_least_squares", "lfb_qr_solve", "lfb_solvec", "lfb_invc", "lfb_cholesky_batched"):
        assert base + "_f32" in names and base + "_f64" in names


def test_library_exports_every_declared_symbol():
    from linfa_linalg_b200 import _ffi
    assert os.path.exists(_ffi.LIB_PATH), "build with __graft_entry__.build()"
    lib = C.CDLL(_ffi.LIB_PATH)
    missing = [n for n in _declared() if not hasattr(lib, n)]
    assert not missing, missing


def test_ffi_table_matches_header():
    from linfa_linalg_b200 import _ffi
    declared = set(_declared())
    bound = set(_ffi.SIGNATURES)
    assert bound <= declared, bound - declared
    _ffi.load()  # declares prototypes; raises if a bound symbol is absent


def test_no_cpu_fallback_without_device():
    """Without a CUDA device the engine must fail loudly, not compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import numpy as np
    import linfa_linalg_b200 as L
    with pytest.raises(L.LinalgError):
        L.qr(np.eye(3))


def test_shape_errors_raised_before_any_device_work():
    """NotThin / NotSquare / WrongRows mirror src/qr.rs:34-36, src/lib.rs:64-71, src/triangular.rs:103-108."""
    import numpy as np
    import linfa_linalg_b200 as L
    with pytest.raises(L.NotThin):
        L.qr_into(np.zeros((2, 3)), eng=object())
    with pytest.raises(L.NotSquare):
        L.cholesky_inplace(np.zeros((2, 3)), eng=object())
    with pytest.raises(L.NotSquare):
        L.sym_tridiagonal(np.zeros((2, 3)), eng=object())
    with pytest.raises(L.EmptyMatrix):
        L.sym_tridiagonal(np.zeros((0, 0)), eng=object())
    with pytest.raises(L.EmptyMatrix):
        L.bidiagonal(np.zeros((0, 0)), eng=object())
    with pytest.raises(L.WrongRows):
        L.solve_triangular_inplace(np.eye(2), np.zeros((1, 2)), L.UPPER, eng=object())
    with pytest.raises(L.NotSquare):                                     # eigh.rs:15 check_square
        L.eigh(np.zeros((2, 3)), eng=object())
    vals, vecs = L.eigh(np.zeros((0, 0)), eng=object())                  # eigh.rs:16-25 / :411-420 corner
    assert vals.shape == (0,) and vecs.shape == (0, 0)
    assert L.eigvalsh(np.zeros((0, 0)), eng=object()).shape == (0,)
    with pytest.raises(L.EmptyMatrix):                                   # svd.rs:23-25, :602-607
        L.svd(np.zeros((0, 1)), False, False, eng=object())
    with pytest.raises(L.NotThin):                                       # qr.rs:34-36 on the tall-skinny route too
        L.qr_tsqr_into(np.zeros((2, 3)), eng=object())
    with pytest.raises(L.NotSquare):                                     # cholesky_yy must be square (lobpcg/algorithm.rs:70)
        L.apply_constraints(np.zeros((5, 2)), np.zeros((2, 3)), np.zeros((5, 2)), eng=object())
    with pytest.raises(ValueError):                                      # y is (rows of v) x (order of cholesky_yy)
        L.apply_constraints(np.zeros((5, 2)), np.eye(3), np.zeros((4, 3)), eng=object())


def test_null_handle_is_an_argument_error_not_a_crash():
    """Every entry point validates the handle before touching CUDA (no device needed to see that)."""
    import ctypes as C
    from linfa_linalg_b200 import _ffi
    lib = _ffi.load()
    buf = (C.c_double * 16)()
    p = C.cast(buf, C.c_void_p)
    assert lib.lfb_qr_f64(None, p, 4, 4, 4, 1, p) == _ffi.INVALID_ARGUMENT
    assert lib.lfb_qr_tsqr_f64(None, p, 4, 4, 4, 1, p) == _ffi.INVALID_ARGUMENT
    assert lib.lfb_qr_tsqr_dev_f64(None, p, 4, 4, 4, p) == _ffi.INVALID_ARGUMENT
    assert lib.lfb_tsqr_explicit_q_dev_f64(None, p, 4, 4, 4, p, 4) == _ffi.INVALID_ARGUMENT
    assert lib.lfb_orthonormalize_f64(None, p, 4, 4, 4, 1, p, 4, 1, None) == _ffi.INVALID_ARGUMENT
    assert lib.lfb_apply_constraints_dev_f64(None, p, 4, 2, 4, p, 2, 2, p, 4) == _ffi.INVALID_ARGUMENT
    assert lib.lfb_set_option(None, b"qr_tsqr_auto", 1) == _ffi.INVALID_ARGUMENT
    # shape errors come before the handle is used
    assert lib.lfb_qr_tsqr_f64(None, p, 2, 3, 3, 1, p) == _ffi.NOT_THIN


def test_sort_eig_host_side():
    """eigh.rs:275-325 EigSort."""
    import numpy as np
    import linfa_linalg_b200 as L
    vals = np.array([3.0, -1.0, 2.0])
    vecs = np.arange(9.0).reshape(3, 3)
    v, q = L.sort_eig_asc((vals, vecs))
    np.testing.assert_array_equal(v, [-1, 2, 3])
    np.testing.assert_array_equal(q, vecs[:, [1, 2, 0]])
    np.testing.assert_array_equal(L.sort_eig_desc(vals), [3, 2, -1])
    with pytest.raises(ValueError):
        L.sort_eig(np.array([1.0, np.nan]))


def test_cpp_mirror_header_compiles(tmp_path):
    """include/linfa_b200.hpp (the C++ host mirror) and examples/qr_kat.cpp compile and link against the in-tree library;
    the program itself needs a GPU and runs in tests/test_gpu_parity_2048.py."""
    import shutil
    import subprocess
    from linfa_linalg_b200 import _ffi
    gxx = shutil.which("g++")
    assert gxx
    libdir = os.path.dirname(_ffi.LIB_PATH)
    subprocess.check_call([gxx, "-std=c++17", "-Wall", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "qr_kat.cpp"), "-L" + libdir, "-llinfa_b200",
                           "-Wl,-rpath," + libdir, "-o", str(tmp_path / "qr_kat")])


def test_set_option_validates_ranges():
    """lfb_set_option rejects unknown keys and out-of-range values before touching the handle's state."""
    from linfa_linalg_b200 import _ffi
    lib = _ffi.load()
    assert lib.lfb_set_option(None, b"chol_nb", 512) == _ffi.INVALID_ARGUMENT


def test_rust_shim_declares_only_exported_symbols():
    """rust/src/ffi.rs (the shim a linfa-linalg maintainer adds; not compiled here: no rustc in the image) must only name
    symbols that the header declares and the library exports."""
    from linfa_linalg_b200 import _ffi
    src = open(os.path.join(ROOT, "rust", "src", "ffi.rs")).read()
    names = set(re.findall(r"\b(lfb_[a-z0-9_]+)\b", src)) - {"lfb_handle", "lfb_multi"}
    declared = set(_declared())
    assert names and names <= declared, sorted(names - declared)
    lib = C.CDLL(_ffi.LIB_PATH)
    assert all(hasattr(lib, n) for n in names)
