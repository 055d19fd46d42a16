"""linfa_linalg_b200 -- host-side mirror of linfa-linalg's public traits over liblinfa_b200.so.

Names, argument meaning and error behaviour follow the reference (rust-ml/linfa-linalg v0.2.1):
`qr`/`qr_into` + `QRDecomp` (src/qr.rs), `cholesky*`/`solvec*`/`invc*` (src/cholesky.rs),
`solve_triangular*`/`into_triangular`/`is_triangular` (src/triangular.rs), `sym_tridiagonal` +
`TridiagonalDecomp` (src/tridiagonal.rs), `bidiagonal` + `BidiagonalDecomp` (src/bidiagonal.rs).
numpy arrays of any strides stand in for ndarray views; the `*_into` / `*_inplace` variants work in
place on the caller's storage.  All arithmetic runs on the GPU through the C ABI; nothing here
falls back to the CPU (importing works without a GPU, the first call does not).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _ffi
from ._ffi import LOWER, UPPER  # noqa: F401  (triangular.rs:10-13 UPLO)

__all__ = [
    "LinalgError", "NotSquare", "NotThin", "NotPositiveDefinite", "NonInvertible", "EmptyMatrix", "WrongRows",
    "Engine", "engine", "UPPER", "LOWER",
    "qr", "qr_into", "qr_tsqr", "qr_tsqr_into", "QRDecomp", "least_squares", "least_squares_into", "qr_batched", "cholesky_batched",
    "cholesky", "cholesky_dirty", "cholesky_into", "cholesky_into_dirty", "cholesky_inplace", "cholesky_inplace_dirty",
    "solvec", "solvec_into", "solvec_inplace", "invc", "invc_inplace", "orthonormalize", "apply_constraints", "generalized_eig", "sorted_eig",
    "solve_triangular", "solve_triangular_into", "solve_triangular_inplace", "triangular_inplace", "into_triangular",
    "is_triangular", "sym_tridiagonal", "TridiagonalDecomp", "bidiagonal", "BidiagonalDecomp",
    "svd", "svd_into", "sort_svd", "sort_svd_asc", "sort_svd_desc",
    "eigh", "eigh_into", "eigvalsh", "eigvalsh_into", "sort_eig", "sort_eig_asc", "sort_eig_desc", "LARGEST", "SMALLEST",
]


# ---- errors: src/lib.rs:33-60 ----------------------------------------------------------------
class LinalgError(Exception):
    pass


class NotSquare(LinalgError):
    def __init__(self, rows, cols):
        super().__init__(f"Matrix of ({rows}, {cols}) is not square")
        self.rows, self.cols = rows, cols


class NotThin(LinalgError):
    def __init__(self, rows, cols):
        super().__init__(f"Expected matrix rows({rows}) >= cols({cols})")
        self.rows, self.cols = rows, cols


class NotPositiveDefinite(LinalgError):
    def __init__(self, index=None):
        super().__init__("Matrix is not positive definite")
        self.index = index


class NonInvertible(LinalgError):
    def __init__(self):
        super().__init__("Matrix is non-invertible")


class EmptyMatrix(LinalgError):
    def __init__(self):
        super().__init__("Matrix is empty")


class WrongRows(LinalgError):
    def __init__(self, expected, actual):
        super().__init__(f"Matrix must have {expected} rows, not {actual}")
        self.expected, self.actual = expected, actual


class DeviceError(LinalgError):
    """CUDA / allocation failure (a new variant; the reference enum is #[non_exhaustive])."""


# ---- engine handle ---------------------------------------------------------------------------
def _sfx(a: np.ndarray) -> str:
    if a.dtype == np.float64:
        return "_f64"
    if a.dtype == np.float32:
        return "_f32"
    raise TypeError(f"A: NdFloat means f32 or f64, got {a.dtype}")


def _view(a: np.ndarray):
    assert a.ndim == 2
    it = a.itemsize
    return (C.c_void_p(a.ctypes.data), a.shape[0], a.shape[1], a.strides[0] // it, a.strides[1] // it)


def _vecp(v: np.ndarray):
    return C.c_void_p(v.ctypes.data)


class Engine:
    """Owns one lfb_handle (one CUDA device, one stream, one workspace pool)."""

    def __init__(self, device: int = 0):
        self.lib = _ffi.load()
        hp = C.c_void_p()
        st = self.lib.lfb_create(C.byref(hp), device)
        if st != _ffi.OK:
            raise DeviceError(f"lfb_create(device={device}) failed with status {st}: no usable CUDA device "
                              "(linfa_linalg_b200 has no CPU fallback)")
        self.h = hp
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self.lib.lfb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_option(self, key: str, value: int):
        if self.lib.lfb_set_option(self.h, key.encode(), int(value)) != _ffi.OK:
            raise ValueError(f"unknown option {key}")

    def set_stream(self, cuda_stream: int | None):
        """cuda_stream: a cudaStream_t handle as int (0 = legacy default stream); None = the engine's own stream."""
        if cuda_stream is None:
            self.lib.lfb_use_own_stream(self.h)
        else:
            self.lib.lfb_set_stream(self.h, C.c_void_p(int(cuda_stream)))

    def synchronize(self):
        self._check(self.lib.lfb_synchronize(self.h))

    @property
    def launch_count(self) -> int:
        return int(self.lib.lfb_launch_count(self.h))

    def _check(self, st: int):
        if st == _ffi.OK:
            return
        if st == _ffi.NON_INVERTIBLE:            # qr.rs:134-136
            raise NonInvertible()
        if st == _ffi.EMPTY_MATRIX:
            raise EmptyMatrix()
        msg = (self.lib.lfb_last_error(self.h) or b"").decode()
        raise DeviceError(f"liblinfa_b200 status {st}: {msg}")

    def call(self, name: str, *args) -> int:
        return getattr(self.lib, name)(self.h, *args)


_default: Engine | None = None


def engine() -> Engine:
    global _default
    if _default is None:
        _default = Engine(0)
    return _default


def _owned(a) -> np.ndarray:
    """`to_owned()`: a fresh row-major copy (qr.rs:61, cholesky.rs:107)."""
    a = np.asarray(a)
    if a.dtype not in (np.float32, np.float64):
        a = a.astype(np.float64)
    return np.array(a, order="C", copy=True)


def _check_square(a):  # lib.rs:64-71
    if a.shape[0] != a.shape[1]:
        raise NotSquare(a.shape[0], a.shape[1])
    return a.shape[0]


# ---- QR: src/qr.rs ---------------------------------------------------------------------------
class QRDecomp:
    """qr.rs:68-73: compact factor `qr` + signed `diag`."""

    def __init__(self, qr_: np.ndarray, diag: np.ndarray, eng: Engine):
        self.qr, self.diag, self._e = qr_, diag, eng

    def generate_q(self) -> np.ndarray:  # qr.rs:86-88
        return _assemble_q(self._e, self.qr, 0, self.diag)

    def into_r(self) -> np.ndarray:  # qr.rs:91-98
        """R as a NEW n x n array (strict lower zeroed, |diag| on the diagonal).  In Rust `into_r(self)` consumes the
        decomposition; a Python object cannot be moved out of, so the compact factor is left intact and the object stays
        usable (`generate_q`, `qt_mul`, `solve*` after `into_r` keep giving the reference's results)."""
        n = self.qr.shape[1]
        r = np.triu(self.qr[:n, :n], 1)
        r[np.arange(n), np.arange(n)] = np.abs(self.diag)
        return r

    def into_decomp(self):  # qr.rs:101-104
        q = self.generate_q()
        return q, self.into_r()

    def qt_mul(self, b: np.ndarray) -> None:  # qr.rs:110-120 (in place)
        st = self._e.call("lfb_qt_mul" + _sfx(self.qr), *_view(self.qr), _vecp(self.diag), *_view(b)[:1],
                          b.shape[1], *_view(b)[3:])
        self._e._check(st)

    def is_invertible(self) -> bool:  # qr.rs:194-197
        return bool(np.all(self.diag != 0))

    def solve_into(self, b: np.ndarray) -> np.ndarray:  # qr.rs:124-152
        """Q^T b and the triangular solve run back to back on the device (lfb_qr_solve); the solution lands in the
        first `ncols` rows of `b`, which is what the reference returns (`b.slice_move(s![..ncols, ..])`)."""
        if self.qr.shape[0] != b.shape[0]:
            raise WrongRows(self.qr.shape[0], b.shape[0])
        if not self.is_invertible():
            raise NonInvertible()
        n = self.qr.shape[1]
        x = b[:n, :]
        st = self._e.call("lfb_qr_solve" + _sfx(self.qr), *_view(self.qr), _vecp(self.diag), *_view(b), *_view(x)[:1], *_view(x)[3:])
        self._e._check(st)
        return x

    def solve_tr_into(self, b: np.ndarray) -> np.ndarray:  # qr.rs:156-181
        n = self.qr.shape[1]
        if n != b.shape[0]:
            raise WrongRows(n, b.shape[0])
        if not self.is_invertible():
            raise NonInvertible()
        x = np.zeros((self.qr.shape[0], b.shape[1]), dtype=self.qr.dtype)
        st = self._e.call("lfb_qr_solve_tr" + _sfx(self.qr), *_view(self.qr), _vecp(self.diag), *_view(b), *_view(x)[:1], *_view(x)[3:])
        self._e._check(st)          # R^T m = b, generate_q and Q m on the device in one round trip (:172-180)
        return x

    def solve(self, b):  # qr.rs:184-186
        return self.solve_into(_owned(b).astype(self.qr.dtype, copy=False))

    def solve_tr(self, b):  # qr.rs:189-191
        return self.solve_tr_into(_owned(b).astype(self.qr.dtype, copy=False))

    def inverse(self) -> np.ndarray:  # qr.rs:200-203
        _check_square(self.qr)
        return self.solve_into(np.eye(len(self.diag), dtype=self.qr.dtype))


def qr_into(a: np.ndarray, eng: Engine | None = None) -> QRDecomp:
    """qr.rs:29-45 QRInto::qr_into -- factors `a` in place."""
    e = eng or engine()
    rows, cols = a.shape
    if rows < cols:
        raise NotThin(rows, cols)
    diag = np.zeros(cols, dtype=a.dtype)
    st = e.call("lfb_qr" + _sfx(a), *_view(a), _vecp(diag))
    e._check(st)
    return QRDecomp(a, diag, e)


def qr(a, eng: Engine | None = None) -> QRDecomp:
    """qr.rs:57-63 QR::qr (by reference: copies first)."""
    return qr_into(_owned(a), eng)


def qr_tsqr_into(a: np.ndarray, eng: Engine | None = None) -> QRDecomp:
    """qr.rs:29-45 for a tall-skinny `a`: the same QRDecomp (same compact factor and diag, to rounding), computed
    as TSQR over row chunks + Householder reconstruction (csrc/tsqr_hr.cu).  `qr_into` takes this route by itself
    once `engine().set_option("qr_tsqr_auto", 1)` is set and the matrix is tall and skinny enough."""
    e = eng or engine()
    rows, cols = a.shape
    if rows < cols:
        raise NotThin(rows, cols)
    diag = np.zeros(cols, dtype=a.dtype)
    st = e.call("lfb_qr_tsqr" + _sfx(a), *_view(a), _vecp(diag))
    e._check(st)
    return QRDecomp(a, diag, e)


def qr_tsqr(a, eng: Engine | None = None) -> QRDecomp:
    return qr_tsqr_into(_owned(a), eng)


def least_squares_into(a: np.ndarray, b: np.ndarray, eng: Engine | None = None) -> np.ndarray:
    """qr.rs:207-229 -- factorisation, Q^T b (or the wide case's R^T solve and Q m) and the triangular solve as ONE
    library call (lfb_least_squares): a and b cross PCIe once, x comes back once."""
    e = eng or engine()
    if a.shape[0] != b.shape[0]:
        raise WrongRows(a.shape[0], b.shape[0])
    x = np.zeros((a.shape[1], b.shape[1]), dtype=a.dtype)
    st = e.call("lfb_least_squares" + _sfx(a), *_view(a), *_view(b), *_view(x)[:1], *_view(x)[3:])
    e._check(st)
    return x


def least_squares(a: np.ndarray, b, eng: Engine | None = None) -> np.ndarray:
    """qr.rs:233-248."""
    return least_squares_into(a, _owned(b).astype(a.dtype, copy=False), eng)


def _assemble_q(e: Engine, m: np.ndarray, shift: int, signs: np.ndarray) -> np.ndarray:
    """householder.rs:68-93."""
    rows, cols = m.shape
    dim = min(rows, cols)
    if shift > dim:
        raise IndexError("shift exceeds matrix dimension (the reference panics here)")
    q = np.zeros((rows, dim), dtype=m.dtype)
    signs = np.ascontiguousarray(signs, dtype=m.dtype)
    sbuf = signs if signs.size else np.zeros(1, dtype=m.dtype)
    st = e.call("lfb_assemble_q" + _sfx(m), *_view(m), shift, _vecp(sbuf), *_view(q)[:1], *_view(q)[3:])
    e._check(st)
    return q


def qr_batched(a: np.ndarray, eng: Engine | None = None) -> np.ndarray:
    """qr.rs:32-44 over a C-contiguous [batch][m][n] array, in place; returns diag [batch][n]."""
    e = eng or engine()
    assert a.ndim == 3 and a.flags.c_contiguous
    batch, m, n = a.shape
    if m < n:
        raise NotThin(m, n)
    diag = np.zeros((batch, n), dtype=a.dtype)
    st = e.call("lfb_qr_batched" + _sfx(a), C.c_void_p(a.ctypes.data), batch, m, n, _vecp(diag))
    e._check(st)
    return diag


def cholesky_batched(a: np.ndarray, clean: bool = True, eng: Engine | None = None) -> np.ndarray:
    """cholesky.rs:51-83 over a C-contiguous [batch][n][n] array (n <= 32), in place.  NotPositiveDefinite carries the
    failing row of the FIRST failing matrix (`.index`) and that matrix's position in the batch (`.matrix`)."""
    e = eng or engine()
    assert a.ndim == 3 and a.flags.c_contiguous
    batch, n, n2 = a.shape
    if n != n2:
        raise NotSquare(n, n2)
    fm, fi = C.c_int64(-1), C.c_int64(-1)
    st = e.call("lfb_cholesky_batched" + _sfx(a), C.c_void_p(a.ctypes.data), batch, n, int(clean), C.byref(fm), C.byref(fi))
    if st == _ffi.NOT_POSITIVE_DEFINITE:
        err = NotPositiveDefinite(fi.value)
        err.matrix = fm.value
        raise err
    e._check(st)
    return a


# ---- Cholesky: src/cholesky.rs -----------------------------------------------------------------
def _chol(a: np.ndarray, clean: bool, e: Engine):
    n = a.shape[0]
    if a.shape[0] != a.shape[1]:
        raise NotSquare(a.shape[0], a.shape[1])
    fail = C.c_int64(-1)
    st = e.call("lfb_cholesky" + _sfx(a), *_view(a), int(clean), C.byref(fail))
    if st == _ffi.NOT_POSITIVE_DEFINITE:
        raise NotPositiveDefinite(fail.value)
    e._check(st)
    return a


def cholesky_inplace_dirty(a, eng=None):  # cholesky.rs:51-76
    return _chol(a, False, eng or engine())


def cholesky_inplace(a, eng=None):  # cholesky.rs:78-82
    return _chol(a, True, eng or engine())


cholesky_into_dirty = cholesky_inplace_dirty  # cholesky.rs:24-31
cholesky_into = cholesky_inplace  # cholesky.rs:37-43


def cholesky_dirty(a, eng=None):  # cholesky.rs:106-109
    return _chol(_owned(a), False, eng or engine())


def cholesky(a, eng=None):  # cholesky.rs:111-114
    return _chol(_owned(a), True, eng or engine())


def _solvec(a: np.ndarray, b: np.ndarray, write_factor: bool, e: Engine) -> np.ndarray:
    _check_square(a)
    if a.shape[0] != b.shape[0]:
        raise WrongRows(a.shape[0], b.shape[0])
    fail = C.c_int64(-1)
    st = e.call("lfb_solvec" + _sfx(a), *_view(a), 1 if write_factor else 0, *_view(b), C.byref(fail))
    if st == _ffi.NOT_POSITIVE_DEFINITE:
        raise NotPositiveDefinite(fail.value)
    e._check(st)
    return b


def solvec_inplace(a: np.ndarray, b: np.ndarray, eng=None) -> np.ndarray:  # cholesky.rs:136-144
    """Cholesky + both triangular solves in one library call (lfb_solvec); `a` receives its factor (dirty)."""
    return _solvec(a, b, True, eng or engine())


def solvec_into(a, b, eng=None):  # cholesky.rs:127-130 (consumes a: the factor is not written back)
    return _solvec(a, b, False, eng or engine())


def solvec(a, b, eng=None):  # cholesky.rs:155-163
    return solvec_inplace(a, _owned(b).astype(a.dtype, copy=False), eng)


def invc_inplace(a, eng=None):  # cholesky.rs:178-182 (the identity right-hand side is generated on the device)
    e = eng or engine()
    n = _check_square(a)
    inv = np.zeros((n, n), dtype=a.dtype)
    fail = C.c_int64(-1)
    st = e.call("lfb_invc" + _sfx(a), *_view(a), *_view(inv)[:1], *_view(inv)[3:], C.byref(fail))
    if st == _ffi.NOT_POSITIVE_DEFINITE:
        raise NotPositiveDefinite(fail.value)
    e._check(st)
    return inv


def invc(a, eng=None):  # cholesky.rs:193-199
    return invc_inplace(_owned(a), eng)


# ---- the dense blocks of LOBPCG: src/lobpcg/algorithm.rs ---------------------------------------------
def orthonormalize(v: np.ndarray, eng=None):
    """lobpcg/algorithm.rs:81-97 orthonormalize(v) -> (u, gram_vv_fac): Gram matrix, its Cholesky factor and the
    triangular solve on the transposed block as ONE library call (lfb_orthonormalize); `v` is consumed (overwritten
    with u, as the reference moves it).  NotPositiveDefinite where cholesky_into fails (:83)."""
    e = eng or engine()
    rows, cols = v.shape
    l = np.zeros((cols, cols), dtype=v.dtype)
    fail = C.c_int64(-1)
    st = e.call("lfb_orthonormalize" + _sfx(v), *_view(v), *_view(l)[:1], *_view(l)[3:], C.byref(fail))
    if st == _ffi.NOT_POSITIVE_DEFINITE:
        raise NotPositiveDefinite(fail.value)
    e._check(st)
    return v, l


def apply_constraints(v: np.ndarray, cholesky_yy: np.ndarray, y: np.ndarray, eng=None) -> np.ndarray:
    """lobpcg/algorithm.rs:63-76: v -= y * solve_triangular(cholesky_yy, y^T v, Lower), in place, one library call."""
    e = eng or engine()
    m = _check_square(cholesky_yy)
    if y.shape != (v.shape[0], m):
        raise ValueError(f"y has shape {y.shape}, expected {(v.shape[0], m)}")   # ndarray panics on the mismatch
    y = np.asarray(y, dtype=v.dtype)
    cholesky_yy = np.asarray(cholesky_yy, dtype=v.dtype)
    lv = _view(cholesky_yy)
    st = e.call("lfb_apply_constraints" + _sfx(v), *_view(v), lv[0], m, lv[3], lv[4], *_view(y))
    e._check(st)
    return v


# ---- triangular: src/triangular.rs ---------------------------------------------------------------
def _host_triangular(a, uplo):
    n = a.shape[0]
    for i in range(n):
        if uplo == UPPER:
            a[i, :i] = 0
        else:
            a[i, i + 1:] = 0


def triangular_inplace(a: np.ndarray, uplo: int, eng=None):  # triangular.rs:37-53
    _check_square(a)
    e = eng or engine()
    e._check(e.call("lfb_triangular_inplace" + _sfx(a), *_view(a), uplo))
    return a


into_triangular = triangular_inplace  # triangular.rs:32-35


def is_triangular(a: np.ndarray, uplo: int) -> bool:  # triangular.rs:67-90 (a host-side predicate)
    if a.shape[0] != a.shape[1]:
        return False
    return bool(np.all(np.tril(a, -1) == 0)) if uplo == UPPER else bool(np.all(np.triu(a, 1) == 0))


def _solve_tri(e: Engine, a: np.ndarray, b: np.ndarray, uplo: int, ext_diag):
    _check_square(a)  # triangular.rs:102
    if b.shape[0] != a.shape[0]:
        raise WrongRows(a.shape[0], b.shape[0])  # :103-108
    dp = None
    if ext_diag is not None:
        ext_diag = np.ascontiguousarray(ext_diag, dtype=a.dtype)
        dp = _vecp(ext_diag)
    st = e.call("lfb_solve_triangular" + _sfx(a), *_view(a), *_view(b), uplo, dp)
    e._check(st)
    return b


def solve_triangular_inplace(a, b, uplo, eng=None):  # triangular.rs:161-168
    return _solve_tri(eng or engine(), a, b, uplo, None)


def solve_triangular_into(a, b, uplo, eng=None):  # triangular.rs:152-155
    return _solve_tri(eng or engine(), a, b, uplo, None)


def solve_triangular(a, b, uplo, eng=None):  # triangular.rs:184-186
    return _solve_tri(eng or engine(), a, _owned(b).astype(a.dtype, copy=False), uplo, None)


# ---- tridiagonal: src/tridiagonal.rs -------------------------------------------------------------
class TridiagonalDecomp:
    """tridiagonal.rs:71-77."""

    def __init__(self, diag_matrix, off_diagonal, eng):
        self.diag_matrix, self.off_diagonal, self._e = diag_matrix, off_diagonal, eng

    def generate_q(self):  # :90-92
        return _assemble_q(self._e, self.diag_matrix, 1, self.off_diagonal)

    def into_diagonals(self):  # :96-101
        return np.array(np.diag(self.diag_matrix)), np.abs(self.off_diagonal)

    def into_tridiag_matrix(self):  # :104-113
        m = self.diag_matrix
        n = m.shape[0]
        d = np.array(np.diag(m))
        m[...] = 0
        idx = np.arange(n)
        m[idx, idx] = d
        off = np.abs(self.off_diagonal)
        m[idx[1:], idx[:-1]] = off
        m[idx[:-1], idx[1:]] = off
        return m


def sym_tridiagonal(a: np.ndarray, eng=None) -> TridiagonalDecomp:
    """tridiagonal.rs:31-66 -- consumes `a` (in place)."""
    e = eng or engine()
    n = _check_square(a)
    if n < 1:
        raise EmptyMatrix()
    off = np.zeros(n - 1, dtype=a.dtype)
    obuf = off if off.size else np.zeros(1, dtype=a.dtype)
    st = e.call("lfb_sym_tridiagonal" + _sfx(a), *_view(a), _vecp(obuf))
    e._check(st)
    return TridiagonalDecomp(a, off, e)


# ---- eigh: src/eigh.rs ------------------------------------------------------------------------------
LARGEST, SMALLEST = "largest", "smallest"   # lib.rs:77-80 Order


def _eigh(a: np.ndarray, vectors: bool, e: Engine):
    n = _check_square(a)
    vals = np.zeros(n, dtype=a.dtype)
    vecs = np.zeros((n, n), dtype=a.dtype) if vectors else None
    if n == 0:                                  # eigh.rs:16-25
        return vals, vecs
    vv = (_vecp(vecs), vecs.strides[0] // vecs.itemsize, vecs.strides[1] // vecs.itemsize) if vectors else (None, 0, 0)
    st = e.call("lfb_eigh" + _sfx(a), *_view(a), _vecp(vals), *vv)
    e._check(st)
    return vals, vecs


def eigh_into(a: np.ndarray, eng=None):
    """eigh.rs:211-219 EighInto::eigh_into -> (eigenvalues in the reference's order, eigenvectors as columns)."""
    return _eigh(a, True, eng or engine())


def eigh(a, eng=None):
    """eigh.rs:231-238 Eigh::eigh (works on a copy, like `to_owned()`)."""
    return eigh_into(_owned(a), eng)


def eigvalsh_into(a: np.ndarray, eng=None) -> np.ndarray:
    """eigh.rs:249-255 EigValshInto::eigvalsh_into (no Q is formed)."""
    return _eigh(a, False, eng or engine())[0]


def eigvalsh(a, eng=None) -> np.ndarray:
    """eigh.rs:266-272 EigValsh::eigvalsh."""
    return eigvalsh_into(_owned(a), eng)


def sort_eig(res, order=SMALLEST):
    """eigh.rs:275-325 EigSort (host-side, as in the reference): a stable sort of the eigenvalues, columns of the
    eigenvector matrix permuted alike.  `res` is an eigenvalue array or an (eigenvalues, eigenvectors) pair."""
    vals, vecs = (res, None) if isinstance(res, np.ndarray) else res
    if np.isnan(vals).any():
        raise ValueError("NaN values in array")          # cmp_floats panics
    idx = np.argsort(-vals if order == LARGEST else vals, kind="stable")
    if vecs is None:
        return vals[idx]
    return vals[idx], np.ascontiguousarray(vecs[:, idx])


def sort_eig_asc(res):
    return sort_eig(res, SMALLEST)


def sort_eig_desc(res):
    return sort_eig(res, LARGEST)


# ---- lobpcg/algorithm.rs:16-44: the small dense eigenproblems of LOBPCG, ONE library call each (lfb_sorted_eig) ----
def _sorted_eig_call(a: np.ndarray, b, size: int, order_code: int, eng=None):
    e = eng or engine()
    k = _check_square(a)
    if b is not None:
        b = np.asarray(b, dtype=a.dtype)
        if b.shape != a.shape:
            raise ValueError(f"b has shape {b.shape}, expected {a.shape}")
    nout = k if order_code == 0 else min(size, k)
    vals = np.zeros(nout, dtype=a.dtype)
    vecs = np.zeros((k, nout), dtype=a.dtype)
    if k == 0:
        return vals, vecs
    it = a.itemsize
    bp = (_vecp(b), b.strides[0] // it, b.strides[1] // it) if b is not None else (None, 0, 0)
    vb, vq = (vals if nout else np.zeros(1, dtype=a.dtype)), (vecs if nout else np.zeros((k, 1), dtype=a.dtype))
    st = e.call("lfb_sorted_eig" + _sfx(a), _vecp(a), k, a.strides[0] // it, a.strides[1] // it, *bp, size, order_code,
                _vecp(vb), _vecp(vq), vq.strides[0] // it, vq.strides[1] // it)
    if st == _ffi.INVALID_ARGUMENT:
        raise ValueError("NaN values in array")          # cmp_floats panics (eigh.rs:326-328)
    e._check(st)
    return vals, vecs


def generalized_eig(a: np.ndarray, b: np.ndarray, eng=None):
    """lobpcg/algorithm.rs:16-25: pencil (A, B) -> (vals_a, vecs_b~ vecs_a), unsorted; both eigendecompositions and the
    k x k products between them run on the device in one call.  `a` and `b` are consumed, as in the reference."""
    return _sorted_eig_call(a, b, a.shape[0], 0, eng)


def sorted_eig(a: np.ndarray, b, size: int, order=LARGEST, eng=None):
    """lobpcg/algorithm.rs:28-44: full (generalized) eigenproblem, sorted by `order`, signs made deterministic by the
    first row (Rust's signum: the sign BIT), truncated to `size`.  `a` and `b` are consumed, as in the reference."""
    return _sorted_eig_call(a, b, size, 1 if order == LARGEST else 2, eng)


# ---- svd: src/svd.rs ---------------------------------------------------------------------------------
def svd_into(a: np.ndarray, calc_u: bool, calc_vt: bool, eng=None):
    """svd.rs:431-443 SVDInto::svd_into -> (u or None, sigma in the reference's order, vt or None)."""
    e = eng or engine()
    rows, cols = a.shape
    if rows == 0 or cols == 0:
        raise EmptyMatrix()                     # svd.rs:23-25
    dim = min(rows, cols)
    s = np.zeros(dim, dtype=a.dtype)
    u = np.zeros((rows, dim), dtype=a.dtype) if calc_u else None
    vt = np.zeros((dim, cols), dtype=a.dtype) if calc_vt else None
    it = a.itemsize
    up = (_vecp(u), u.strides[0] // it, u.strides[1] // it) if calc_u else (None, 0, 0)
    vp = (_vecp(vt), vt.strides[0] // it, vt.strides[1] // it) if calc_vt else (None, 0, 0)
    st = e.call("lfb_svd" + _sfx(a), *_view(a), _vecp(s), *up, *vp)
    e._check(st)
    return u, s, vt


def svd(a, calc_u: bool, calc_vt: bool, eng=None):
    """svd.rs:462-478 SVD::svd (works on a copy, like `to_owned()`)."""
    return svd_into(_owned(a), calc_u, calc_vt, eng)


def sort_svd(res, order=LARGEST):
    """svd.rs:487-527 SvdSort (host-side, as in the reference): stable sort of sigma, columns of U and rows of Vt alike."""
    u, s, vt = res
    if np.isnan(s).any():
        raise ValueError("NaN values in array")
    idx = np.argsort(-s if order == LARGEST else s, kind="stable")
    return (None if u is None else np.ascontiguousarray(u[:, idx])), s[idx], (None if vt is None else np.ascontiguousarray(vt[idx, :]))


def sort_svd_asc(res):
    return sort_svd(res, SMALLEST)


def sort_svd_desc(res):
    return sort_svd(res, LARGEST)


# ---- bidiagonal: src/bidiagonal.rs ---------------------------------------------------------------
class BidiagonalDecomp:
    """bidiagonal.rs:64-69."""

    def __init__(self, uv, diagonal, off_diagonal, upper_diag, eng):
        self.uv, self.diagonal, self.off_diagonal, self.upper_diag, self._e = uv, diagonal, off_diagonal, upper_diag, eng

    def is_upper_diag(self):  # :85-87
        return self.upper_diag

    def generate_u(self):  # :90-97
        shift = 0 if self.upper_diag else 1
        return _assemble_q(self._e, self.uv, shift, self.diagonal if self.upper_diag else self.off_diagonal)

    def generate_vt(self):  # :101-109
        shift = 1 if self.upper_diag else 0
        return _assemble_q(self._e, self.uv.T, shift, self.off_diagonal if self.upper_diag else self.diagonal).T

    def into_diagonals(self):  # :126-131
        return np.abs(self.diagonal), np.abs(self.off_diagonal)

    def into_b(self):  # :113-123
        d, e = self.into_diagonals()
        return np.diag(d) + (np.diag(e, 1) if self.upper_diag else np.diag(e, -1))


def bidiagonal(a: np.ndarray, eng=None) -> BidiagonalDecomp:
    """bidiagonal.rs:27-59 -- consumes `a` (in place)."""
    e = eng or engine()
    rows, cols = a.shape
    md = min(rows, cols)
    if md == 0:
        raise EmptyMatrix()
    d = np.zeros(md, dtype=a.dtype)
    off = np.zeros(md - 1, dtype=a.dtype)
    obuf = off if off.size else np.zeros(1, dtype=a.dtype)
    st = e.call("lfb_bidiagonal" + _sfx(a), *_view(a), _vecp(d), _vecp(obuf))
    e._check(st)
    return BidiagonalDecomp(a, d, off, rows >= cols, e)
