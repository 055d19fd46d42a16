// linfa_b200.hpp -- header-only C++ mirror of linfa-linalg's public traits over the C ABI
// (include/linfa_b200.h).  Names, argument meaning and error behaviour follow the reference
// (rust-ml/linfa-linalg v0.2.1): QR/QRInto + QRDecomp (src/qr.rs), Cholesky*/SolveC*/InverseC*
// (src/cholesky.rs), SolveTriangular*/IntoTriangular (src/triangular.rs), SymmetricTridiagonal +
// TridiagonalDecomp (src/tridiagonal.rs), Bidiagonal + BidiagonalDecomp (src/bidiagonal.rs).
// `View<T>` plays the role of an ndarray ArrayBase<_, Ix2> view: pointer + shape + signed element
// strides; `Matrix<T>` is an owned row-major array (ndarray's default layout).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#include "linfa_b200.h"

namespace linfa_b200 {

// src/lib.rs:33-60
struct LinalgError : std::runtime_error {
    int status;
    LinalgError(int st, const std::string &what) : std::runtime_error(what), status(st) {}
};
struct NotSquare : LinalgError { NotSquare(int64_t r, int64_t c) : LinalgError(LFB_NOT_SQUARE, "Matrix of (" + std::to_string(r) + ", " + std::to_string(c) + ") is not square") {} };
struct NotThin : LinalgError { NotThin(int64_t r, int64_t c) : LinalgError(LFB_NOT_THIN, "Expected matrix rows(" + std::to_string(r) + ") >= cols(" + std::to_string(c) + ")") {} };
struct NotPositiveDefinite : LinalgError { int64_t index; explicit NotPositiveDefinite(int64_t i) : LinalgError(LFB_NOT_POSITIVE_DEFINITE, "Matrix is not positive definite"), index(i) {} };
struct NonInvertible : LinalgError { NonInvertible() : LinalgError(LFB_NON_INVERTIBLE, "Matrix is non-invertible") {} };
struct EmptyMatrix : LinalgError { EmptyMatrix() : LinalgError(LFB_EMPTY_MATRIX, "Matrix is empty") {} };
struct WrongRows : LinalgError { WrongRows(int64_t e, int64_t a) : LinalgError(LFB_WRONG_ROWS, "Matrix must have " + std::to_string(e) + " rows, not " + std::to_string(a)) {} };

enum class UPLO { Upper = LFB_UPPER, Lower = LFB_LOWER };   // src/triangular.rs:10-13

template <typename T>
struct View {
    T *ptr; int64_t rows, cols, rs, cs;
    T &operator()(int64_t i, int64_t j) const { return ptr[i * rs + j * cs]; }
    View t() const { return {ptr, cols, rows, cs, rs}; }                                  // reversed_axes
    View slice(int64_t r0, int64_t r1, int64_t c0, int64_t c1) const { return {ptr + r0 * rs + c0 * cs, r1 - r0, c1 - c0, rs, cs}; }
};

template <typename T>
struct Matrix {
    int64_t rows = 0, cols = 0;
    std::vector<T> data;
    Matrix() = default;
    Matrix(int64_t r, int64_t c, T fill = T(0)) : rows(r), cols(c), data((size_t)(r * c), fill) {}
    static Matrix eye(int64_t n) { Matrix m(n, n); for (int64_t i = 0; i < n; ++i) m(i, i) = T(1); return m; }
    T &operator()(int64_t i, int64_t j) { return data[(size_t)(i * cols + j)]; }
    const T &operator()(int64_t i, int64_t j) const { return data[(size_t)(i * cols + j)]; }
    View<T> view() { return {data.data(), rows, cols, cols, 1}; }
    static Matrix from(const View<const T> &v) { Matrix m(v.rows, v.cols); for (int64_t i = 0; i < v.rows; ++i) for (int64_t j = 0; j < v.cols; ++j) m(i, j) = v(i, j); return m; }
};

// One engine handle (one CUDA device, stream, workspace pool).  No CPU fallback: throws without a device.
class Engine {
  public:
    explicit Engine(int device = 0) { if (lfb_create(&h_, device) != LFB_OK) throw LinalgError(LFB_ERR_CUDA, "lfb_create failed: no usable CUDA device"); }
    ~Engine() { lfb_destroy(h_); }
    Engine(const Engine &) = delete;
    Engine &operator=(const Engine &) = delete;
    lfb_handle *handle() const { return h_; }
    void check(int st) const { if (st != LFB_OK) throw LinalgError(st, std::string("liblinfa_b200: ") + lfb_last_error(h_)); }
  private:
    lfb_handle *h_ = nullptr;
};

namespace detail {
template <typename T> constexpr bool is64 = std::is_same<T, double>::value;
#define LFB_DISPATCH(name, T, ...) (detail::is64<T> ? name##_f64(__VA_ARGS__) : name##_f32(__VA_ARGS__))
inline int qr(lfb_handle *h, double *a, int64_t r, int64_t c, int64_t rs, int64_t cs, double *d) { return lfb_qr_f64(h, a, r, c, rs, cs, d); }
inline int qr(lfb_handle *h, float *a, int64_t r, int64_t c, int64_t rs, int64_t cs, float *d) { return lfb_qr_f32(h, a, r, c, rs, cs, d); }
inline int qr_tsqr(lfb_handle *h, double *a, int64_t r, int64_t c, int64_t rs, int64_t cs, double *d) { return lfb_qr_tsqr_f64(h, a, r, c, rs, cs, d); }
inline int qr_tsqr(lfb_handle *h, float *a, int64_t r, int64_t c, int64_t rs, int64_t cs, float *d) { return lfb_qr_tsqr_f32(h, a, r, c, rs, cs, d); }
inline int assemble_q(lfb_handle *h, const double *m, int64_t r, int64_t c, int64_t rs, int64_t cs, int64_t sh, const double *s, double *q, int64_t qrs, int64_t qcs) { return lfb_assemble_q_f64(h, m, r, c, rs, cs, sh, s, q, qrs, qcs); }
inline int assemble_q(lfb_handle *h, const float *m, int64_t r, int64_t c, int64_t rs, int64_t cs, int64_t sh, const float *s, float *q, int64_t qrs, int64_t qcs) { return lfb_assemble_q_f32(h, m, r, c, rs, cs, sh, s, q, qrs, qcs); }
inline int qt_mul(lfb_handle *h, const double *m, int64_t r, int64_t c, int64_t rs, int64_t cs, const double *d, double *b, int64_t bc, int64_t brs, int64_t bcs) { return lfb_qt_mul_f64(h, m, r, c, rs, cs, d, b, bc, brs, bcs); }
inline int qt_mul(lfb_handle *h, const float *m, int64_t r, int64_t c, int64_t rs, int64_t cs, const float *d, float *b, int64_t bc, int64_t brs, int64_t bcs) { return lfb_qt_mul_f32(h, m, r, c, rs, cs, d, b, bc, brs, bcs); }
inline int cholesky(lfb_handle *h, double *a, int64_t r, int64_t c, int64_t rs, int64_t cs, int clean, int64_t *fi) { return lfb_cholesky_f64(h, a, r, c, rs, cs, clean, fi); }
inline int cholesky(lfb_handle *h, float *a, int64_t r, int64_t c, int64_t rs, int64_t cs, int clean, int64_t *fi) { return lfb_cholesky_f32(h, a, r, c, rs, cs, clean, fi); }
inline int solve_tri(lfb_handle *h, const double *a, int64_t ar, int64_t ac, int64_t ars, int64_t acs, double *b, int64_t br, int64_t bc, int64_t brs, int64_t bcs, int u, const double *d) { return lfb_solve_triangular_f64(h, a, ar, ac, ars, acs, b, br, bc, brs, bcs, u, d); }
inline int solve_tri(lfb_handle *h, const float *a, int64_t ar, int64_t ac, int64_t ars, int64_t acs, float *b, int64_t br, int64_t bc, int64_t brs, int64_t bcs, int u, const float *d) { return lfb_solve_triangular_f32(h, a, ar, ac, ars, acs, b, br, bc, brs, bcs, u, d); }
inline int tridiag(lfb_handle *h, double *a, int64_t r, int64_t c, int64_t rs, int64_t cs, double *off) { return lfb_sym_tridiagonal_f64(h, a, r, c, rs, cs, off); }
inline int tridiag(lfb_handle *h, float *a, int64_t r, int64_t c, int64_t rs, int64_t cs, float *off) { return lfb_sym_tridiagonal_f32(h, a, r, c, rs, cs, off); }
inline int bidiag(lfb_handle *h, double *a, int64_t r, int64_t c, int64_t rs, int64_t cs, double *d, double *e) { return lfb_bidiagonal_f64(h, a, r, c, rs, cs, d, e); }
inline int bidiag(lfb_handle *h, float *a, int64_t r, int64_t c, int64_t rs, int64_t cs, float *d, float *e) { return lfb_bidiagonal_f32(h, a, r, c, rs, cs, d, e); }
inline int eigh(lfb_handle *h, const double *a, int64_t r, int64_t c, int64_t rs, int64_t cs, double *v, double *q, int64_t qrs, int64_t qcs) { return lfb_eigh_f64(h, a, r, c, rs, cs, v, q, qrs, qcs); }
inline int eigh(lfb_handle *h, const float *a, int64_t r, int64_t c, int64_t rs, int64_t cs, float *v, float *q, int64_t qrs, int64_t qcs) { return lfb_eigh_f32(h, a, r, c, rs, cs, v, q, qrs, qcs); }
inline int svd(lfb_handle *h, const double *a, int64_t r, int64_t c, int64_t rs, int64_t cs, double *s, double *u, int64_t urs, int64_t ucs, double *vt, int64_t vrs, int64_t vcs) { return lfb_svd_f64(h, a, r, c, rs, cs, s, u, urs, ucs, vt, vrs, vcs); }
inline int svd(lfb_handle *h, const float *a, int64_t r, int64_t c, int64_t rs, int64_t cs, float *s, float *u, int64_t urs, int64_t ucs, float *vt, int64_t vrs, int64_t vcs) { return lfb_svd_f32(h, a, r, c, rs, cs, s, u, urs, ucs, vt, vrs, vcs); }
inline int least_squares(lfb_handle *h, const double *a, int64_t r, int64_t c, int64_t rs, int64_t cs, const double *b, int64_t br, int64_t bc, int64_t brs, int64_t bcs, double *x, int64_t xrs, int64_t xcs) { return lfb_least_squares_f64(h, a, r, c, rs, cs, b, br, bc, brs, bcs, x, xrs, xcs); }
inline int least_squares(lfb_handle *h, const float *a, int64_t r, int64_t c, int64_t rs, int64_t cs, const float *b, int64_t br, int64_t bc, int64_t brs, int64_t bcs, float *x, int64_t xrs, int64_t xcs) { return lfb_least_squares_f32(h, a, r, c, rs, cs, b, br, bc, brs, bcs, x, xrs, xcs); }
template <typename T> T signum(T x) { return std::signbit(x) ? T(-1) : T(1); }
}  // namespace detail

// householder.rs:68-93
template <typename T>
Matrix<T> assemble_q(Engine &e, View<T> m, int64_t shift, const std::vector<T> &signs) {
    const int64_t dim = std::min(m.rows, m.cols);
    if (shift > dim) throw std::out_of_range("shift exceeds matrix dimension");   // the reference panics here
    Matrix<T> q(m.rows, dim);
    if (m.rows && dim) e.check(detail::assemble_q(e.handle(), m.ptr, m.rows, m.cols, m.rs, m.cs, shift, signs.data(), q.data.data(), dim, 1));
    return q;
}

// triangular.rs:95-144 (ext_diag == nullptr: the diagonal of a)
template <typename T>
void solve_triangular_inplace(Engine &e, View<T> a, View<T> b, UPLO uplo, const T *ext_diag = nullptr) {
    if (a.rows != a.cols) throw NotSquare(a.rows, a.cols);
    if (b.rows != a.rows) throw WrongRows(a.rows, b.rows);
    e.check(detail::solve_tri(e.handle(), a.ptr, a.rows, a.cols, a.rs, a.cs, b.ptr, b.rows, b.cols, b.rs, b.cs, (int)uplo, ext_diag));
}

// qr.rs:68-203
template <typename T>
class QRDecomp {
  public:
    QRDecomp(Engine &e, View<T> qr, std::vector<T> diag) : e_(e), qr_(qr), diag_(std::move(diag)) {}
    Matrix<T> generate_q() const { return assemble_q(e_, qr_, 0, diag_); }                     // :86-88
    Matrix<T> into_r() const {                                                                 // :91-98
        const int64_t n = qr_.cols;
        Matrix<T> r(n, n);
        for (int64_t i = 0; i < n; ++i) { r(i, i) = std::abs(diag_[(size_t)i]); for (int64_t j = i + 1; j < n; ++j) r(i, j) = qr_(i, j); }
        return r;
    }
    void qt_mul(View<T> b) const { e_.check(detail::qt_mul(e_.handle(), qr_.ptr, qr_.rows, qr_.cols, qr_.rs, qr_.cs, diag_.data(), b.ptr, b.cols, b.rs, b.cs)); }   // :110-120
    bool is_invertible() const { return std::all_of(diag_.begin(), diag_.end(), [](T d) { return d != T(0); }); }                                             // :194-197
    View<T> solve_into(View<T> b) const {                                                      // :124-152
        if (qr_.rows != b.rows) throw WrongRows(qr_.rows, b.rows);
        if (!is_invertible()) throw NonInvertible();
        qt_mul(b);
        const int64_t n = qr_.cols;
        View<T> x = b.slice(0, n, 0, b.cols);
        std::vector<T> ad(diag_);
        for (auto &d : ad) d = std::abs(d);
        solve_triangular_inplace(e_, qr_.slice(0, n, 0, n), x, UPLO::Upper, ad.data());
        return x;
    }
    Matrix<T> inverse() const {                                                                // :200-203
        if (qr_.rows != qr_.cols) throw NotSquare(qr_.rows, qr_.cols);
        Matrix<T> eye = Matrix<T>::eye((int64_t)diag_.size());
        solve_into(eye.view());
        return eye;
    }
    const std::vector<T> &diag() const { return diag_; }
  private:
    Engine &e_; View<T> qr_; std::vector<T> diag_;
};

// qr.rs:29-45  QRInto::qr_into (in place on the caller's storage)
template <typename T>
QRDecomp<T> qr_into(Engine &e, View<T> a) {
    if (a.rows < a.cols) throw NotThin(a.rows, a.cols);
    std::vector<T> diag((size_t)a.cols, T(0));
    e.check(detail::qr(e.handle(), a.ptr, a.rows, a.cols, a.rs, a.cs, diag.data()));
    return QRDecomp<T>(e, a, std::move(diag));
}

// qr.rs:29-45 for a tall-skinny view: the same QRDecomp through TSQR + Householder reconstruction (csrc/tsqr_hr.cu)
template <typename T>
QRDecomp<T> qr_tsqr_into(Engine &e, View<T> a) {
    if (a.rows < a.cols) throw NotThin(a.rows, a.cols);
    std::vector<T> diag((size_t)a.cols, T(0));
    e.check(detail::qr_tsqr(e.handle(), a.ptr, a.rows, a.cols, a.rs, a.cs, diag.data()));
    return QRDecomp<T>(e, a, std::move(diag));
}

// cholesky.rs:51-83
template <typename T>
void cholesky_inplace(Engine &e, View<T> a, bool clean = true) {
    if (a.rows != a.cols) throw NotSquare(a.rows, a.cols);
    int64_t fail = -1;
    const int st = detail::cholesky(e.handle(), a.ptr, a.rows, a.cols, a.rs, a.cs, clean ? 1 : 0, &fail);
    if (st == LFB_NOT_POSITIVE_DEFINITE) throw NotPositiveDefinite(fail);
    e.check(st);
}
template <typename T> void cholesky_inplace_dirty(Engine &e, View<T> a) { cholesky_inplace(e, a, false); }

// cholesky.rs:136-144
template <typename T>
void solvec_inplace(Engine &e, View<T> a, View<T> b) {
    cholesky_inplace_dirty(e, a);
    solve_triangular_inplace(e, a, b, UPLO::Lower);
    solve_triangular_inplace(e, a.t(), b, UPLO::Upper);
}

// lobpcg/algorithm.rs:81-97  orthonormalize(v) -> (u, gram_vv_fac): v is overwritten with u, the factor is returned
template <typename T>
Matrix<T> orthonormalize(Engine &e, View<T> v) {
    Matrix<T> l(v.cols, v.cols);
    int64_t fail = -1;
    int st;
    if constexpr (detail::is64<T>) st = lfb_orthonormalize_f64(e.handle(), v.ptr, v.rows, v.cols, v.rs, v.cs, l.data.data(), l.cols, 1, &fail);
    else st = lfb_orthonormalize_f32(e.handle(), v.ptr, v.rows, v.cols, v.rs, v.cs, l.data.data(), l.cols, 1, &fail);
    if (st == LFB_NOT_POSITIVE_DEFINITE) throw NotPositiveDefinite(fail);
    e.check(st);
    return l;
}

// lobpcg/algorithm.rs:63-76  apply_constraints(v, cholesky_yy, y), in place on v
template <typename T>
void apply_constraints(Engine &e, View<T> v, View<T> cholesky_yy, View<T> y) {
    if (cholesky_yy.rows != cholesky_yy.cols) throw NotSquare(cholesky_yy.rows, cholesky_yy.cols);
    int st;
    if constexpr (detail::is64<T>)
        st = lfb_apply_constraints_f64(e.handle(), v.ptr, v.rows, v.cols, v.rs, v.cs, cholesky_yy.ptr, cholesky_yy.rows, cholesky_yy.rs,
                                       cholesky_yy.cs, y.ptr, y.rows, y.cols, y.rs, y.cs);
    else
        st = lfb_apply_constraints_f32(e.handle(), v.ptr, v.rows, v.cols, v.rs, v.cs, cholesky_yy.ptr, cholesky_yy.rows, cholesky_yy.rs,
                                       cholesky_yy.cs, y.ptr, y.rows, y.cols, y.rs, y.cs);
    e.check(st);
}

// tridiagonal.rs:31-113
template <typename T>
struct TridiagonalDecomp {
    Engine &e; View<T> diag_matrix; std::vector<T> off_diagonal;
    Matrix<T> generate_q() const { return assemble_q(e, diag_matrix, 1, off_diagonal); }
    std::pair<std::vector<T>, std::vector<T>> into_diagonals() const {
        std::vector<T> d((size_t)diag_matrix.rows), o(off_diagonal);
        for (int64_t i = 0; i < diag_matrix.rows; ++i) d[(size_t)i] = diag_matrix(i, i);
        for (auto &x : o) x = std::abs(x);
        return {d, o};
    }
};
template <typename T>
TridiagonalDecomp<T> sym_tridiagonal(Engine &e, View<T> a) {
    if (a.rows != a.cols) throw NotSquare(a.rows, a.cols);
    if (a.rows < 1) throw EmptyMatrix();
    std::vector<T> off((size_t)std::max<int64_t>(a.rows - 1, 1), T(0));
    e.check(detail::tridiag(e.handle(), a.ptr, a.rows, a.cols, a.rs, a.cs, off.data()));
    off.resize((size_t)(a.rows - 1));
    return {e, a, std::move(off)};
}

// bidiagonal.rs:27-131
template <typename T>
struct BidiagonalDecomp {
    Engine &e; View<T> uv; std::vector<T> diagonal, off_diagonal; bool upper_diag;
    bool is_upper_diag() const { return upper_diag; }
    Matrix<T> generate_u() const { return assemble_q(e, uv, upper_diag ? 0 : 1, upper_diag ? diagonal : off_diagonal); }
    Matrix<T> generate_vt() const {   // assemble_q on the transposed view, then reversed_axes
        Matrix<T> q = assemble_q(e, uv.t(), upper_diag ? 1 : 0, upper_diag ? off_diagonal : diagonal);
        Matrix<T> vt(q.cols, q.rows);
        for (int64_t i = 0; i < q.rows; ++i) for (int64_t j = 0; j < q.cols; ++j) vt(j, i) = q(i, j);
        return vt;
    }
};
template <typename T>
BidiagonalDecomp<T> bidiagonal(Engine &e, View<T> a) {
    const int64_t md = std::min(a.rows, a.cols);
    if (md == 0) throw EmptyMatrix();
    std::vector<T> d((size_t)md, T(0)), off((size_t)std::max<int64_t>(md - 1, 1), T(0));
    e.check(detail::bidiag(e.handle(), a.ptr, a.rows, a.cols, a.rs, a.cs, d.data(), off.data()));
    off.resize((size_t)(md - 1));
    return {e, a, std::move(d), std::move(off), a.rows >= a.cols};
}

// eigh.rs:202-268 EighInto / EigValshInto: eigenvalues in the reference's (unsorted) order, eigenvectors as columns.
template <typename T>
std::pair<std::vector<T>, Matrix<T>> eigh_into(Engine &e, View<T> a) {
    if (a.rows != a.cols) throw NotSquare(a.rows, a.cols);
    std::vector<T> vals((size_t)a.rows, T(0));
    Matrix<T> vecs(a.rows, a.rows);
    if (a.rows) e.check(detail::eigh(e.handle(), a.ptr, a.rows, a.cols, a.rs, a.cs, vals.data(), vecs.data.data(), a.rows, 1));
    return {std::move(vals), std::move(vecs)};
}
template <typename T>
std::vector<T> eigvalsh_into(Engine &e, View<T> a) {
    if (a.rows != a.cols) throw NotSquare(a.rows, a.cols);
    std::vector<T> vals((size_t)a.rows, T(0));
    if (a.rows) e.check(detail::eigh(e.handle(), a.ptr, a.rows, a.cols, a.rs, a.cs, vals.data(), (T *)nullptr, 0, 0));
    return vals;
}

// svd.rs:415-479 SVDInto: (u, sigma, vt); u / vt are empty matrices when not requested.
template <typename T>
struct SvdResult { Matrix<T> u; std::vector<T> sigma; Matrix<T> vt; };
template <typename T>
SvdResult<T> svd_into(Engine &e, View<T> a, bool calc_u, bool calc_vt) {
    if (a.rows == 0 || a.cols == 0) throw EmptyMatrix();
    const int64_t dim = std::min(a.rows, a.cols);
    SvdResult<T> r;
    r.sigma.assign((size_t)dim, T(0));
    if (calc_u) r.u = Matrix<T>(a.rows, dim);
    if (calc_vt) r.vt = Matrix<T>(dim, a.cols);
    e.check(detail::svd(e.handle(), a.ptr, a.rows, a.cols, a.rs, a.cs, r.sigma.data(), calc_u ? r.u.data.data() : (T *)nullptr, dim, 1,
                        calc_vt ? r.vt.data.data() : (T *)nullptr, a.cols, 1));
    return r;
}

// qr.rs:207-229 LeastSquaresQrInto (thin and wide), one library call.
template <typename T>
Matrix<T> least_squares_into(Engine &e, View<T> a, View<T> b) {
    if (a.rows != b.rows) throw WrongRows(a.rows, b.rows);
    Matrix<T> x(a.cols, b.cols);
    const int st = detail::least_squares(e.handle(), a.ptr, a.rows, a.cols, a.rs, a.cs, b.ptr, b.rows, b.cols, b.rs, b.cs, x.data.data(), b.cols, 1);
    if (st == LFB_NON_INVERTIBLE) throw NonInvertible();
    e.check(st);
    return x;
}

}  // namespace linfa_b200
