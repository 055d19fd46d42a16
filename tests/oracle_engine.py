"""TEST INFRASTRUCTURE: an engine object that answers the C-ABI calls of include/linfa_b200.h with the CPU oracle.

`linfa_linalg_b200`'s mirror of the reference traits reaches the device through `Engine.call(name, *ctypes_args)`.
`OracleEngine` implements that one method on top of `oracle/` (same argument lists, same in-place/strided-view
contract, same status codes), so the mirror's HOST logic -- shape checks, error variants, slicing, sorting, the by-ref /
into / inplace variants -- and the reference's own property tests (tests/*.rs) can run on a machine without a GPU, and
the very same test bodies then run against the real engine with `-m gpu`.  It lives under tests/ (the only place besides
smoke() and bench.py's CPU legs that may import `oracle`); the product never sees it.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
from numpy.lib.stride_tricks import as_strided

import oracle as O

OK, NOT_POSITIVE_DEFINITE, NOT_THIN, NOT_SQUARE, EMPTY_MATRIX, WRONG_ROWS, NON_INVERTIBLE, INVALID_ARGUMENT, UNSUPPORTED = range(9)
_DT = {"f64": np.float64, "f32": np.float32}
_CT = {"f64": C.c_double, "f32": C.c_float}


def _addr(p):
    if p is None:
        return 0
    if isinstance(p, int):
        return p
    return p.value or 0


def _mat(p, rows, cols, rs, cs, sfx):
    """Writable numpy view of the ndarray-style strided matrix (ptr, rows, cols, rs, cs); strides in elements, signed."""
    dt = np.dtype(_DT[sfx])
    if rows == 0 or cols == 0 or _addr(p) == 0:
        return np.zeros((rows, cols), dtype=dt)
    origin = np.ctypeslib.as_array((_CT[sfx] * 1).from_address(_addr(p)))
    return as_strided(origin, shape=(rows, cols), strides=(rs * dt.itemsize, cs * dt.itemsize), writeable=True)


def _vec(p, n, sfx):
    if n == 0 or _addr(p) == 0:
        return np.zeros(n, dtype=_DT[sfx])
    return np.ctypeslib.as_array((_CT[sfx] * n).from_address(_addr(p)))


def _set(ref, value):
    """Store through an out-parameter passed as ctypes.byref(x) (or None)."""
    if ref is not None:
        ref._obj.value = value


class OracleEngine:
    """Drop-in for linfa_linalg_b200.Engine in tests: `.call` runs the oracle, `._check` raises the mirror's errors."""

    def __init__(self):
        import linfa_linalg_b200 as L
        self._L = L
        self.calls = []

    # -- the parts of Engine the mirror uses --------------------------------------------------------------
    def _check(self, st: int):
        L = self._L
        if st == OK:
            return
        if st == NON_INVERTIBLE:
            raise L.NonInvertible()
        if st == EMPTY_MATRIX:
            raise L.EmptyMatrix()
        raise L.DeviceError(f"oracle engine status {st}")

    def call(self, name: str, *a) -> int:
        assert name.startswith("lfb_")
        base, sfx = name[4:].rsplit("_", 1)
        self.calls.append(name)
        return getattr(self, "_" + base)(sfx, *a)

    def set_option(self, key, value):
        pass

    def close(self):
        pass

    def _qr_solve_tr(self, sfx, p, rows, cols, rs, cs, diag, b, b_rows, bcols, b_rs, b_cs, x, x_rs, x_cs):
        """qr.rs:156-181 solve_tr_into: R^T m = b with |diag| as the diagonal, then Q m."""
        if rows < cols:
            return NOT_THIN
        if b_rows != cols:
            return WRONG_ROWS
        d = _vec(diag, cols, sfx)
        if np.any(d == 0):
            return NON_INVERTIBLE
        if rows == 0 or cols == 0 or bcols == 0:
            return OK
        qrm = _mat(p, rows, cols, rs, cs, sfx)
        w = np.array(_mat(b, cols, bcols, b_rs, b_cs, sfx))
        O.solve_triangular(qrm[:cols, :cols].T, w, O.LOWER, ext_diag=np.abs(d))
        _mat(x, rows, bcols, x_rs, x_cs, sfx)[...] = O.generate_q(qrm, d) @ w
        return OK

    # -- lobpcg/algorithm.rs:16-44 -----------------------------------------------------------------------------
    def _sorted_eig(self, sfx, a, k, a_rs, a_cs, b, b_rs, b_cs, size, order, vals, vecs, v_rs, v_cs):
        if k == 0:
            return OK
        am = np.array(_mat(a, k, k, a_rs, a_cs, sfx))
        if _addr(b):
            bm = np.array(_mat(b, k, k, b_rs, b_cs, sfx))
            vb, qb = O.symmetric_eig(bm, vectors=True)                                   # :17
            floor = am.dtype.type(np.float32(1e-10))
            qb = qb * (1.0 / np.sqrt(np.maximum(vb, floor))).astype(am.dtype)            # :18-19
            at = np.ascontiguousarray(qb.T @ (am @ qb))                                  # :20
            va, qa = O.symmetric_eig(at, vectors=True)                                   # :21
            q = qb @ qa                                                                  # :22
        else:
            va, q = O.symmetric_eig(am, vectors=True)
        nout = k
        if order != 0:
            if np.isnan(va).any():
                return INVALID_ARGUMENT
            idx = np.argsort(-va if order == 1 else va, kind="stable")
            nout = min(size, k)
            va, q = va[idx][:nout], q[:, idx][:, :nout]
            q = q * np.where(np.signbit(q[0, :]), -1.0, 1.0).astype(q.dtype)
        _vec(vals, nout, sfx)[:] = va
        _mat(vecs, k, nout, v_rs, v_cs, sfx)[...] = q
        return OK

    # -- qr.rs ----------------------------------------------------------------------------------------------
    def _qr(self, sfx, p, rows, cols, rs, cs, diag):
        if rows < cols:
            return NOT_THIN
        if cols == 0:
            return OK
        _vec(diag, cols, sfx)[:] = O.qr(_mat(p, rows, cols, rs, cs, sfx))
        return OK

    _qr_tsqr = _qr          # same contract (include/linfa_b200.h)

    def _assemble_q(self, sfx, p, rows, cols, rs, cs, shift, signs, q, q_rs, q_cs):
        dim = min(rows, cols)
        if shift > dim:
            return INVALID_ARGUMENT
        if rows == 0 or dim == 0:
            return OK
        m = _mat(p, rows, cols, rs, cs, sfx)
        _mat(q, rows, dim, q_rs, q_cs, sfx)[...] = O.assemble_q(m, shift, _vec(signs, max(dim - shift, 0), sfx))
        return OK

    def _qt_mul(self, sfx, p, rows, cols, rs, cs, diag, b, bcols, b_rs, b_cs):
        if rows < cols:
            return NOT_THIN
        if rows == 0 or cols == 0 or bcols == 0:
            return OK
        O.qt_mul(_mat(p, rows, cols, rs, cs, sfx), _vec(diag, cols, sfx), _mat(b, rows, bcols, b_rs, b_cs, sfx))
        return OK

    def _qr_solve(self, sfx, p, rows, cols, rs, cs, diag, b, b_rows, bcols, b_rs, b_cs, x, x_rs, x_cs):
        """qr.rs:124-152 solve_into: qt_mul, then the upper solve with |diag| on the first `cols` rows."""
        if rows < cols:
            return NOT_THIN
        if b_rows != rows:
            return WRONG_ROWS
        d = _vec(diag, cols, sfx)
        if np.any(d == 0):
            return NON_INVERTIBLE
        if cols == 0 or bcols == 0:
            return OK
        qrm = _mat(p, rows, cols, rs, cs, sfx)
        w = np.array(_mat(b, rows, bcols, b_rs, b_cs, sfx))
        O.qt_mul(qrm, d, w)
        top = np.ascontiguousarray(w[:cols])
        O.solve_triangular(qrm[:cols, :cols], top, O.UPPER, ext_diag=np.abs(d))
        _mat(x, cols, bcols, x_rs, x_cs, sfx)[...] = top
        return OK

    def _least_squares(self, sfx, p, rows, cols, rs, cs, b, b_rows, bcols, b_rs, b_cs, x, x_rs, x_cs):
        """qr.rs:207-229: thin -> qr_into + solve_into; wide -> QR of the transpose + solve_tr_into (:156-181)."""
        if b_rows != rows:
            return WRONG_ROWS
        a = np.array(_mat(p, rows, cols, rs, cs, sfx))
        bm = np.array(_mat(b, rows, bcols, b_rs, b_cs, sfx))
        xm = _mat(x, cols, bcols, x_rs, x_cs, sfx)
        if rows >= cols:
            d = O.qr(a)
            if np.any(d == 0):
                return NON_INVERTIBLE
            O.qt_mul(a, d, bm)
            top = np.ascontiguousarray(bm[:cols])
            O.solve_triangular(a[:cols, :cols], top, O.UPPER, ext_diag=np.abs(d))
            xm[...] = top
        else:
            at = np.ascontiguousarray(a.T)                     # cols x rows, thin
            d = O.qr(at)
            if np.any(d == 0):
                return NON_INVERTIBLE
            O.solve_triangular(at[:rows, :rows].T, bm, O.LOWER, ext_diag=np.abs(d))      # R^T m = b (:172-177)
            xm[...] = O.generate_q(at, d) @ bm                                          # Q m (:180)
        return OK

    # -- cholesky.rs / triangular.rs ---------------------------------------------------------------------------
    def _cholesky(self, sfx, p, rows, cols, rs, cs, clean, fail):
        if rows != cols:
            return NOT_SQUARE
        _set(fail, -1)
        if rows == 0:
            return OK
        st, fi = O.cholesky(_mat(p, rows, cols, rs, cs, sfx), clean=bool(clean))
        if st != 0:
            _set(fail, fi)
            return NOT_POSITIVE_DEFINITE
        return OK

    def _solve_triangular(self, sfx, a, a_rows, a_cols, a_rs, a_cs, b, b_rows, b_cols, b_rs, b_cs, uplo, ext):
        if a_rows != a_cols:
            return NOT_SQUARE
        if b_rows != a_rows:
            return WRONG_ROWS
        if a_rows == 0 or b_cols == 0:
            return OK
        ed = _vec(ext, a_rows, sfx) if _addr(ext) else None
        O.solve_triangular(_mat(a, a_rows, a_cols, a_rs, a_cs, sfx), _mat(b, b_rows, b_cols, b_rs, b_cs, sfx), uplo, ext_diag=ed)
        return OK

    def _triangular_inplace(self, sfx, p, rows, cols, rs, cs, uplo):
        if rows != cols:
            return NOT_SQUARE
        if rows:
            O.triangular_inplace(_mat(p, rows, cols, rs, cs, sfx), uplo)
        return OK

    def _solvec(self, sfx, p, rows, cols, rs, cs, write_factor, b, b_rows, bcols, b_rs, b_cs, fail):
        """cholesky.rs:136-144: dirty factorisation, L y = b, L^T x = y."""
        if rows != cols:
            return NOT_SQUARE
        if b_rows != rows:
            return WRONG_ROWS
        _set(fail, -1)
        if rows == 0:
            return OK
        a = _mat(p, rows, cols, rs, cs, sfx)
        f = a if write_factor else np.array(a)
        st, fi = O.cholesky(f, clean=False)
        if st != 0:
            _set(fail, fi)
            return NOT_POSITIVE_DEFINITE
        if bcols:
            bm = _mat(b, rows, bcols, b_rs, b_cs, sfx)
            O.solve_triangular(f, bm, O.LOWER)
            O.solve_triangular(f.T, bm, O.UPPER)
        return OK

    def _invc(self, sfx, p, rows, cols, rs, cs, inv, i_rs, i_cs, fail):
        """cholesky.rs:178-182: solvec with the identity."""
        if rows != cols:
            return NOT_SQUARE
        _set(fail, -1)
        if rows == 0:
            return OK
        f = np.array(_mat(p, rows, cols, rs, cs, sfx))
        st, fi = O.cholesky(f, clean=False)
        if st != 0:
            _set(fail, fi)
            return NOT_POSITIVE_DEFINITE
        out = _mat(inv, rows, rows, i_rs, i_cs, sfx)
        out[...] = np.eye(rows, dtype=out.dtype)
        O.solve_triangular(f, out, O.LOWER)
        O.solve_triangular(f.T, out, O.UPPER)
        return OK

    # -- tridiagonal.rs / bidiagonal.rs / eigh.rs / svd.rs -----------------------------------------------------
    def _sym_tridiagonal(self, sfx, p, rows, cols, rs, cs, off):
        if rows != cols:
            return NOT_SQUARE
        if rows < 1:
            return EMPTY_MATRIX
        _vec(off, rows - 1, sfx)[:] = O.sym_tridiagonal(_mat(p, rows, cols, rs, cs, sfx))
        return OK

    def _bidiagonal(self, sfx, p, rows, cols, rs, cs, d, e):
        md = min(rows, cols)
        if md < 1:
            return EMPTY_MATRIX
        dd, ee = O.bidiagonal(_mat(p, rows, cols, rs, cs, sfx))
        _vec(d, md, sfx)[:] = dd
        _vec(e, md - 1, sfx)[:] = ee
        return OK

    def _eigh(self, sfx, p, rows, cols, rs, cs, vals, vecs, v_rs, v_cs):
        if rows != cols:
            return NOT_SQUARE
        if rows == 0:
            return OK
        want = _addr(vecs) != 0
        v, q = O.symmetric_eig(np.array(_mat(p, rows, cols, rs, cs, sfx)), vectors=want)
        _vec(vals, rows, sfx)[:] = v
        if want:
            _mat(vecs, rows, rows, v_rs, v_cs, sfx)[...] = q
        return OK

    def _svd(self, sfx, p, rows, cols, rs, cs, sigma, u, u_rs, u_cs, vt, v_rs, v_cs):
        dim = min(rows, cols)
        if dim < 1:
            return EMPTY_MATRIX
        wu, wv = _addr(u) != 0, _addr(vt) != 0
        uu, s, vv = O.svd(np.array(_mat(p, rows, cols, rs, cs, sfx)), wu, wv)
        _vec(sigma, dim, sfx)[:] = s
        if wu:
            _mat(u, rows, dim, u_rs, u_cs, sfx)[...] = uu
        if wv:
            _mat(vt, dim, cols, v_rs, v_cs, sfx)[...] = vv
        return OK

    # -- batched, lobpcg blocks ------------------------------------------------------------------------------------
    def _qr_batched(self, sfx, p, batch, m, n, diag):
        if m < n:
            return NOT_THIN
        if batch == 0 or n == 0:
            return OK
        a = _vec(p, batch * m * n, sfx).reshape(batch, m, n)
        _vec(diag, batch * n, sfx).reshape(batch, n)[...] = O.qr_batched(a)
        return OK

    def _cholesky_batched(self, sfx, p, batch, n, clean, fm, fi):
        _set(fm, -1)
        _set(fi, -1)
        if batch == 0 or n == 0:
            return OK
        a = _vec(p, batch * n * n, sfx).reshape(batch, n, n)
        m, i = O.cholesky_batched(a, bool(clean))
        if m >= 0:
            _set(fm, m)
            _set(fi, i)
            return NOT_POSITIVE_DEFINITE
        return OK

    def _orthonormalize(self, sfx, p, rows, cols, rs, cs, l, l_rs, l_cs, fail):
        _set(fail, -1)
        if cols == 0:
            return OK
        v = _mat(p, rows, cols, rs, cs, sfx)
        st, fi, u, lf = O.lobpcg_orthonormalize(np.array(v))
        if st != 0:
            _set(fail, fi)
            return NOT_POSITIVE_DEFINITE
        v[...] = u
        if _addr(l):
            _mat(l, cols, cols, l_rs, l_cs, sfx)[...] = lf
        return OK

    def _apply_constraints(self, sfx, p, n, k, rs, cs, lyy, m, l_rs, l_cs, y, y_rows, y_cols, y_rs, y_cs):
        if y_rows != n or y_cols != m:
            return INVALID_ARGUMENT
        if n == 0 or k == 0 or m == 0:
            return OK
        v = _mat(p, n, k, rs, cs, sfx)
        w = np.array(v)
        O.lobpcg_apply_constraints(w, np.array(_mat(lyy, m, m, l_rs, l_cs, sfx)), np.array(_mat(y, n, m, y_rs, y_cs, sfx)))
        v[...] = w
        return OK
