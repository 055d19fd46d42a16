"""ctypes front-end of the CPU oracle (oracle/lfo_oracle.c).  TEST INFRASTRUCTURE ONLY.

Each function mirrors one reference routine (file:line in the docstring) and works
on numpy arrays of ANY strides in place, like the reference's ndarray views do.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liblfo_oracle.so")
_lib = None

UPPER, LOWER = 0, 1


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (oracle/Makefile).  Building the checker is not using it."""
    src = [os.path.join(_HERE, f) for f in ("lfo_oracle.c", "lfo_impl.inc", "Makefile")]
    stale = (not os.path.exists(_SO)) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in src)
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-B"], stdout=subprocess.DEVNULL)
    return _SO


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
    return _lib


def _sfx(a: np.ndarray) -> str:
    if a.dtype == np.float64:
        return "_f64"
    if a.dtype == np.float32:
        return "_f32"
    raise TypeError(f"oracle supports f32/f64 only, got {a.dtype}")


def _ct(a):
    return C.c_double if a.dtype == np.float64 else C.c_float


def _p(a: np.ndarray):
    return C.c_void_p(a.ctypes.data)


def _es(a: np.ndarray, ax: int) -> int:
    return a.strides[ax] // a.itemsize


def _i(x):
    return C.c_int64(int(x))


def reflection_axis(col: np.ndarray):
    """householder.rs:9-28.  In place on a 1-D view; returns the scalar or None."""
    out = _ct(col)()
    f = getattr(lib(), "lfo_reflection_axis" + _sfx(col))
    f.restype = C.c_int
    some = f(_p(col), _i(col.shape[0]), _i(_es(col, 0)), C.byref(out))
    return col.dtype.type(out.value) if some else None


def reflect_cols(axis: np.ndarray, rhs: np.ndarray, bias=0.0):
    """reflection.rs:26-32."""
    f = getattr(lib(), "lfo_reflect_cols" + _sfx(rhs))
    f.restype = None
    f(_p(axis), _i(axis.shape[0]), _i(_es(axis, 0)), _ct(rhs)(bias),
      _p(rhs), _i(rhs.shape[1]), _i(_es(rhs, 0)), _i(_es(rhs, 1)))


def reflect_rows(axis: np.ndarray, lhs: np.ndarray, bias=0.0):
    """reflection.rs:35-37."""
    reflect_cols(axis, lhs.T, bias)


def clear_column(m: np.ndarray, icol: int, shift: int):
    """householder.rs:34-51."""
    f = getattr(lib(), "lfo_clear_column" + _sfx(m))
    f.restype = _ct(m)
    return f(_p(m), _i(m.shape[0]), _i(m.shape[1]), _i(_es(m, 0)), _i(_es(m, 1)), _i(icol), _i(shift))


def clear_row(m: np.ndarray, irow: int, shift: int):
    """householder.rs:57-63."""
    return clear_column(m.T, irow, shift)


def assemble_q(m: np.ndarray, shift: int, signs: np.ndarray) -> np.ndarray:
    """householder.rs:68-93.  Returns a new row-major nrows x min(nrows,ncols) array."""
    rows, cols = m.shape
    dim = min(rows, cols)
    if shift > dim:
        raise IndexError("shift exceeds matrix dimension (the reference panics here)")
    q = np.zeros((rows, dim), dtype=m.dtype)
    signs = np.ascontiguousarray(signs, dtype=m.dtype)
    if rows and dim:
        f = getattr(lib(), "lfo_assemble_q" + _sfx(m))
        f.restype = None
        f(_p(m), _i(rows), _i(cols), _i(_es(m, 0)), _i(_es(m, 1)), _i(shift), _p(signs),
          _p(q), _i(_es(q, 0)), _i(_es(q, 1)))
    return q


def qr(a: np.ndarray) -> np.ndarray:
    """qr.rs:32-44.  Factors `a` in place (compact form), returns diag."""
    rows, cols = a.shape
    if rows < cols:
        raise ValueError("NotThin")
    diag = np.zeros(cols, dtype=a.dtype)
    if cols:
        f = getattr(lib(), "lfo_qr" + _sfx(a))
        f.restype = C.c_int
        f(_p(a), _i(rows), _i(cols), _i(_es(a, 0)), _i(_es(a, 1)), _p(diag))
    return diag


def qr_into_r(qrm: np.ndarray, diag: np.ndarray) -> np.ndarray:
    """qr.rs:91-98."""
    n = qrm.shape[1]
    r = np.triu(np.asarray(qrm[:n, :n]), 1).astype(qrm.dtype)
    r[np.arange(n), np.arange(n)] = np.abs(diag)
    return r


def generate_q(qrm: np.ndarray, diag: np.ndarray) -> np.ndarray:
    """qr.rs:86-88."""
    return assemble_q(qrm, 0, diag)


def qt_mul(qrm: np.ndarray, diag: np.ndarray, b: np.ndarray):
    """qr.rs:110-120, in place on b (rows(b) >= rows(qr))."""
    f = getattr(lib(), "lfo_qt_mul" + _sfx(qrm))
    f.restype = None
    diag = np.ascontiguousarray(diag, dtype=qrm.dtype)
    f(_p(qrm), _i(qrm.shape[0]), _i(qrm.shape[1]), _i(_es(qrm, 0)), _i(_es(qrm, 1)), _p(diag),
      _p(b), _i(b.shape[1]), _i(_es(b, 0)), _i(_es(b, 1)))


def triangular_inplace(a: np.ndarray, uplo: int):
    """triangular.rs:37-53."""
    if a.shape[0] != a.shape[1]:
        raise ValueError("NotSquare")
    f = getattr(lib(), "lfo_triangular_inplace" + _sfx(a))
    f.restype = None
    f(_p(a), _i(a.shape[0]), _i(_es(a, 0)), _i(_es(a, 1)), C.c_int(uplo))


def solve_triangular(a: np.ndarray, b: np.ndarray, uplo: int, ext_diag=None):
    """triangular.rs:95-144, in place on b."""
    if a.shape[0] != a.shape[1]:
        raise ValueError("NotSquare")
    if b.shape[0] != a.shape[0]:
        raise ValueError("WrongRows")
    f = getattr(lib(), "lfo_solve_triangular" + _sfx(a))
    f.restype = None
    dp = None
    if ext_diag is not None:
        ext_diag = np.ascontiguousarray(ext_diag, dtype=a.dtype)
        dp = _p(ext_diag)
    f(_p(a), _i(a.shape[0]), _i(_es(a, 0)), _i(_es(a, 1)),
      _p(b), _i(b.shape[1]), _i(_es(b, 0)), _i(_es(b, 1)), C.c_int(uplo), dp)


def cholesky(a: np.ndarray, clean: bool = True):
    """cholesky.rs:51-83, in place.  Returns (status, fail_index): status 1 = NotPositiveDefinite."""
    if a.shape[0] != a.shape[1]:
        raise ValueError("NotSquare")
    fail = C.c_int64(-1)
    f = getattr(lib(), "lfo_cholesky" + _sfx(a))
    f.restype = C.c_int
    st = f(_p(a), _i(a.shape[0]), _i(_es(a, 0)), _i(_es(a, 1)), C.c_int(int(clean)), C.byref(fail))
    return st, fail.value


def sym_tridiagonal(a: np.ndarray) -> np.ndarray:
    """tridiagonal.rs:31-66, in place; returns the signed off-diagonal (n-1)."""
    n = a.shape[0]
    if a.shape[0] != a.shape[1]:
        raise ValueError("NotSquare")
    if n < 1:
        raise ValueError("EmptyMatrix")
    off = np.zeros(n - 1, dtype=a.dtype)
    p = np.zeros(max(n - 1, 1), dtype=a.dtype)
    f = getattr(lib(), "lfo_sym_tridiagonal" + _sfx(a))
    f.restype = None
    f(_p(a), _i(n), _i(_es(a, 0)), _i(_es(a, 1)), _p(off), _p(p))
    return off


def bidiagonal(a: np.ndarray):
    """bidiagonal.rs:27-59, in place; returns signed (diagonal, off_diagonal)."""
    rows, cols = a.shape
    md = min(rows, cols)
    if md == 0:
        raise ValueError("EmptyMatrix")
    d = np.zeros(md, dtype=a.dtype)
    e = np.zeros(max(md - 1, 0), dtype=a.dtype)
    ebuf = e if e.size else np.zeros(1, dtype=a.dtype)
    f = getattr(lib(), "lfo_bidiagonal" + _sfx(a))
    f.restype = None
    f(_p(a), _i(rows), _i(cols), _i(_es(a, 0)), _i(_es(a, 1)), _p(d), _p(ebuf))
    return d, e


def qr_batched(a: np.ndarray) -> np.ndarray:
    """qr.rs:32-44 over a packed [batch][m][n] C-contiguous array, in place; returns diag [batch][n]."""
    assert a.flags.c_contiguous and a.ndim == 3
    batch, m, n = a.shape
    diag = np.zeros((batch, n), dtype=a.dtype)
    f = getattr(lib(), "lfo_qr_batched" + _sfx(a))
    f.restype = None
    f(_p(a), _i(batch), _i(m), _i(n), _p(diag))
    return diag


def symmetric_eig(a: np.ndarray, vectors: bool = True, eps=None):
    """eigh.rs:10-129 (symmetric_eig): `a` is consumed; returns (vals, vecs or None) in the reference's own
    (unsorted) order, eigenvectors as columns."""
    n = a.shape[0]
    if a.shape[0] != a.shape[1]:
        raise ValueError("NotSquare")
    vals = np.zeros(n, dtype=a.dtype)
    if n < 1:
        return vals, (np.zeros((0, 0), dtype=a.dtype) if vectors else None)
    q = np.zeros((n, n), dtype=a.dtype) if vectors else np.zeros((1, 1), dtype=a.dtype)
    off = np.zeros(n, dtype=a.dtype)
    work = np.zeros(n, dtype=a.dtype)
    e = np.finfo(a.dtype).eps if eps is None else eps
    f = getattr(lib(), "lfo_symmetric_eig" + _sfx(a))
    f.restype = None
    f(_p(a), _i(n), _i(_es(a, 0)), _i(_es(a, 1)), C.c_int(1 if vectors else 0), _ct(a)(e),
      _p(vals), _p(q), _i(_es(q, 0)), _i(_es(q, 1)), _p(off), _p(work))
    return vals, (q if vectors else None)


def svd(a: np.ndarray, calc_u: bool = True, calc_vt: bool = True, eps=None):
    """svd.rs:17-221 (svd), `a` is consumed; returns (u or None, s, vt or None) in the reference's own order.
    eps defaults to 5 * machine epsilon like SVDInto::svd_into (svd.rs:441)."""
    rows, cols = a.shape
    if rows == 0 or cols == 0:
        raise ValueError("EmptyMatrix")
    dim = min(rows, cols)
    s = np.zeros(dim, dtype=a.dtype)
    u = np.zeros((rows, dim), dtype=a.dtype) if calc_u else None
    vt = np.zeros((dim, cols), dtype=a.dtype) if calc_vt else None
    off = np.zeros(max(dim, 1), dtype=a.dtype)
    e = 5 * np.finfo(a.dtype).eps if eps is None else eps
    f = getattr(lib(), "lfo_svd" + _sfx(a))
    f.restype = C.c_int
    up = (_p(u), _i(_es(u, 0)), _i(_es(u, 1))) if calc_u else (None, _i(0), _i(0))
    vp = (_p(vt), _i(_es(vt, 0)), _i(_es(vt, 1))) if calc_vt else (None, _i(0), _i(0))
    st = f(_p(a), _i(rows), _i(cols), _i(_es(a, 0)), _i(_es(a, 1)), _ct(a)(e), _p(s), *up, *vp, _p(off))
    assert st == 0
    return u, s, vt


def cholesky_batched(a: np.ndarray, clean: bool = True):
    """cholesky.rs:51-83 in a loop over a packed [batch][n][n] array, in place; returns (first failing matrix or -1,
    its failing row or -1), stopping at the first failure like a caller of the reference would."""
    assert a.ndim == 3 and a.shape[1] == a.shape[2]
    for b in range(a.shape[0]):
        st, fail = cholesky(a[b], clean=clean)
        if st != 0:
            return b, fail
    return -1, -1


def lobpcg_orthonormalize(v: np.ndarray):
    """lobpcg/algorithm.rs:81-97 orthonormalize(v) -> (u, gram_vv_fac), composed from the restated routines:
    gram = v^T v (:82, ndarray `dot`: third-party GEMM, restated by definition), cholesky_into (:83, clean lower),
    u = solve_triangular_into(L, v^T, Lower)^T (:91-94).  Returns (status, fail_index, u, L); status 1 = NotPositiveDefinite."""
    gram = np.ascontiguousarray(v.T @ v)
    st, fi = cholesky(gram, clean=True)
    if st != 0:
        return st, fi, None, None
    vt = np.ascontiguousarray(v.T)
    solve_triangular(gram, vt, 1)
    return 0, -1, np.ascontiguousarray(vt.T), gram


def lobpcg_apply_constraints(v: np.ndarray, cholesky_yy: np.ndarray, y: np.ndarray):
    """lobpcg/algorithm.rs:63-76 apply_constraints, in place on v: gram_yv = y^T v (:68); u = solve_triangular_into
    (cholesky_yy, gram_yv, Lower) (:70-72); v -= y u (:75, general_mat_mul)."""
    u = np.ascontiguousarray(y.T @ v)
    solve_triangular(np.ascontiguousarray(cholesky_yy), u, 1)
    v -= y @ u
    return v
