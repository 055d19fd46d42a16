// Compact singular value decomposition: src/svd.rs:17-221 (svd) end to end.
//
// Phase 1 (scale by max|a|, bidiagonalise, generate U and V) runs on the kernels of bidiag.cu / householder.cu.
// Phase 2, the implicit-shift Golub-Kahan QR iteration (svd.rs:55-208), follows the same division of labour as
// eigh.cu: the scalar recurrence on (diag, off) runs on the host in the reference's order and emits chains of
// Givens rotations -- one chain for U and one for V per sweep -- which the device applies row-parallel, 16
// sweeps per pass (rotations.cuh).  Vt is kept as V = Vt^T on the device, so "rotate rows k, k+1 of Vt"
// (givens.rs:109-119 rotate_cols) is the same column-pair kernel as for U.  The rare rotations on NON-adjacent
// column pairs (cancel_horizontal / cancel_vertical, svd.rs:293-370, taken when a diagonal entry vanishes) are
// applied one launch each, in order with the chains.
#include "rotations.cuh"

namespace lfb {

namespace {

template <typename T>
struct SvdState {
    std::vector<T> diag, off;
    RotationApplier<T> *u, *v;      // nullptr if not requested
    bool upper;
};

// GivensRotation::cancel_y (givens.rs:16-27): c = x / r, s = -y / r, r = hypot(x, y); None when y == 0.
template <typename T>
inline bool cancel_y(T x, T y, T &c, T &s, T &r) {
    if (y == T(0)) return false;
    r = h_hypot(x, y);
    c = x / r;
    s = -y / r;
    return true;
}

// svd.rs:293-328
template <typename T>
void cancel_horizontal(SvdState<T> &st, int64_t i, int64_t end) {
    T v0 = st.off[i], v1 = st.diag[i + 1];
    st.off[i] = T(0);
    for (int64_t k = i; k < end; ++k) {
        T c, s, r;
        if (!cancel_y<T>(v1, v0, c, s, r)) break;            // cancel_x(v0, v1) = cancel_y(v1, v0) ...
        s = -s;                                              // ... with s negated (givens.rs:31-36)
        st.diag[k + 1] = r;
        if (st.upper) {
            if (st.u) st.u->pair(i, k + 1, c, -s);           // rot.inverse().rotate_rows(u[.., (i, k+1)])
        } else if (st.v) {
            st.v->pair(i, k + 1, c, -s);                     // rot.rotate_cols(vt[(i, k+1), ..])
        }
        if (k + 1 != end) {
            v0 = -s * st.off[k + 1];
            v1 = st.diag[k + 2];
            st.off[k + 1] *= c;
        }
    }
}

// svd.rs:330-370
template <typename T>
void cancel_vertical(SvdState<T> &st, int64_t i) {
    T v0 = st.diag[i], v1 = st.off[i];
    st.off[i] = T(0);
    for (int64_t k = i; k >= 0; --k) {
        T c, s, r;
        if (!cancel_y<T>(v0, v1, c, s, r)) break;
        st.diag[k] = r;
        if (st.upper) {
            if (st.v) st.v->pair(k, i + 1, c, -s);           // rot.rotate_cols(vt[(k, i+1), ..])
        } else if (st.u) {
            st.u->pair(k, i + 1, c, -s);                     // rot.inverse().rotate_rows(u[.., (k, i+1)])
        }
        if (k > 0) {
            v0 = st.diag[k - 1];
            v1 = s * st.off[k - 1];
            st.off[k - 1] *= c;
        }
    }
}

// svd.rs:223-291
template <typename T>
void svd_delimit(SvdState<T> &st, int64_t end, T eps, int64_t &start, int64_t &nend) {
    auto &diag = st.diag;
    auto &off = st.off;
    int64_t n = end;
    while (n > 0) {
        const int64_t m = n - 1;
        if (off[m] == T(0) || std::fabs(off[m]) <= eps * (std::fabs(diag[n]) + std::fabs(diag[m]))) {
            off[m] = T(0);
        } else if (std::fabs(diag[m]) <= eps) {
            diag[m] = T(0);
            cancel_horizontal<T>(st, m, m + 1);
            if (m != 0) cancel_vertical<T>(st, m - 1);
        } else if (std::fabs(diag[n]) <= eps) {
            diag[n] = T(0);
            cancel_vertical<T>(st, m);
        } else {
            break;
        }
        n -= 1;
    }
    if (n == 0) { start = 0; nend = 0; return; }
    int64_t ns = n - 1;
    while (ns > 0) {
        const int64_t m = ns - 1;
        if (std::fabs(off[m]) <= eps * (std::fabs(diag[ns]) + std::fabs(diag[m]))) {
            off[m] = T(0);
            break;
        }
        if (std::fabs(diag[m]) <= eps) {
            diag[m] = T(0);
            cancel_horizontal<T>(st, m, n);
            if (m != 0) cancel_vertical<T>(st, m - 1);
            break;
        }
        ns -= 1;
    }
    start = ns; nend = n;
}

// GivensRotation::new (givens.rs:59-62): unit (c, s) and the norm; identity and 0 when the norm is 0.
template <typename T>
inline void givens_new(T c, T s, T &oc, T &os, T &norm) {
    norm = std::hypot(c, s);
    if (norm > T(0)) { oc = c / norm; os = s / norm; }
    else { oc = T(1); os = T(0); norm = T(0); }
}

}  // namespace

// svd.rs:17-221.  dA (rows x cols column-major, ld) is consumed.  sv: HOST array of min(rows, cols) singular values
// in the reference's own order.  dU: device rows x dim (ldu) or nullptr; dV: device cols x dim (ldv) holding
// V = Vt^T, or nullptr.
template <typename T>
void svd_dev(lfb_handle &h, T *dA, int64_t rows, int64_t cols, int64_t ld, T *sv, T *dU, int64_t ldu, T *dV, int64_t ldv) {
    const int64_t dim = std::min(rows, cols);
    if (dim < 1) return;
    FAST_HYPOT = h.opt.fast_hypot != 0;
    const T eps = std::numeric_limits<T>::epsilon() * T(5);                      // svd.rs:441
    const bool upper = rows >= cols;                                             // bidiagonal.rs:36
    DevBuf<T> scal(h, 2), dD(h, dim), dE(h, dim);
    LFB_CUDA(cudaMemsetAsync(scal.get(), 0, 2 * sizeof(T), h.stream));
    {
        dim3 grid((unsigned)std::min<int64_t>(cdiv(rows, 256), 64), (unsigned)std::min<int64_t>(cols, 1024));
        absmax_kernel<T><<<grid, 256, 0, h.stream>>>(dA, ld, rows, cols, scal.get());      // :29-32
        LFB_LAUNCH_CHECK(h);
        scale_div_kernel<T><<<grid, 256, 0, h.stream>>>(dA, ld, rows, cols, scal.get());   // :34-36
        LFB_LAUNCH_CHECK(h);
    }
    bidiagonal<T>(h, dA, rows, cols, ld, dD.get(), dE.get());                               // :38
    if (dU)                                                                                 // :40, bidiagonal.rs:90-97
        assemble_q<T>(h, dA, rows, cols, ld, upper ? 0 : 1, upper ? dD.get() : dE.get(), dU, ldu);
    if (dV) {                                                                               // :41, bidiagonal.rs:101-109
        const int64_t ldt = round_up(cols, 2);
        DevBuf<T> At(h, (size_t)ldt * rows);
        transpose<T>(h, dA, rows, cols, ld, At.get(), ldt);
        assemble_q<T>(h, At.get(), cols, rows, ldt, upper ? 1 : 0, upper ? dE.get() : dD.get(), dV, ldv);
    }
    SvdState<T> st;
    st.diag.resize(dim);
    st.off.resize(std::max<int64_t>(dim, 1));
    st.upper = upper;
    T amax = T(0);
    LFB_CUDA(cudaMemcpyAsync(st.diag.data(), dD.get(), sizeof(T) * dim, cudaMemcpyDeviceToHost, h.stream));
    if (dim > 1) LFB_CUDA(cudaMemcpyAsync(st.off.data(), dE.get(), sizeof(T) * (dim - 1), cudaMemcpyDeviceToHost, h.stream));
    LFB_CUDA(cudaMemcpyAsync(&amax, scal.get(), sizeof(T), cudaMemcpyDeviceToHost, h.stream));
    LFB_CUDA(cudaStreamSynchronize(h.stream));
    for (int64_t i = 0; i < dim; ++i) st.diag[i] = std::fabs(st.diag[i]);                   // bidiagonal.rs:126-131
    for (int64_t i = 0; i + 1 < dim; ++i) st.off[i] = std::fabs(st.off[i]);
    std::unique_ptr<RotationApplier<T>> au, av;
    if (dU) au.reset(new RotationApplier<T>(h, dU, ldu, rows));
    if (dV) av.reset(new RotationApplier<T>(h, dV, ldv, cols));
    st.u = au.get();
    st.v = av.get();
    auto &diag = st.diag;
    auto &off = st.off;
    const bool cu = dU != nullptr, cv = dV != nullptr;

    int64_t start, end;
    svd_delimit<T>(st, dim - 1, eps, start, end);                                           // :44-52
    while (end != start) {                                                                  // :55
        const int64_t subdim = end - start + 1;
        if (subdim > 2) {
            const int64_t m = end - 1, n = end;
            const T dm = diag[m], dn = diag[n], fm = off[m], fm1 = off[m - 1];              // :62-76
            const T tmm = dm * dm + fm1 * fm1, tmn = dm * fm, tnn = dn * dn + fm * fm;
            const T shift = wilkinson_shift<T>(tmm, tnn, tmn);
            const T ds = diag[start];
            T vec0 = ds * ds - shift, vec1 = ds * off[start];
            Chain<T> chu, chv;
            chu.p = chv.p = start;
            for (int64_t k = start; k < n; ++k) {                                           // :78
                const T m12 = (k == n - 1) ? T(0) : off[k + 1];
                T s00 = diag[k], s01 = off[k], s02 = T(0), s10 = T(0), s11 = diag[k + 1], s12 = m12;   // subm (2 x 3)
                T c1, sn1, r1;
                if (!cancel_y<T>(vec0, vec1, c1, sn1, r1)) break;                           // :98, :154-156
                {   // rot1.inverse().rotate_rows(subm[.., 0..=1])   :99-101
                    const T a0 = s00, b0 = s01, a1 = s10, b1 = s11;
                    s00 = a0 * c1 - sn1 * b0; s01 = sn1 * a0 + b0 * c1;
                    s10 = a1 * c1 - sn1 * b1; s11 = sn1 * a1 + b1 * c1;
                }
                if (k > start) off[k - 1] = r1;                                             // :105-107
                T c2 = T(1), sn2 = T(0), r2;
                T norm2;
                const bool have2 = cancel_y<T>(s00, s10, c2, sn2, r2);                      // :109-110
                if (have2) {   // rot.rotate_cols(subm[.., 1..=2])   :111
                    const T a0 = s01, b0 = s11, a1 = s02, b1 = s12;
                    s01 = a0 * c2 - sn2 * b0; s11 = sn2 * a0 + b0 * c2;
                    s02 = a1 * c2 - sn2 * b1; s12 = sn2 * a1 + b1 * c2;
                    norm2 = r2;
                } else {
                    c2 = T(1); sn2 = T(0);
                    norm2 = s00;                                                            // :115-116
                }
                s00 = norm2;                                                                // :118
                if (cv) {                                                                   // :121-129: rotate_cols on rows k, k+1 of Vt
                    if (upper) { chv.c.push_back(c1); chv.s.push_back(-sn1); }
                    else { chv.c.push_back(c2); chv.s.push_back(-sn2); }                    // identity when rot2 is None
                }
                if (cu) {                                                                   // :131-141: inverse().rotate_rows on columns k, k+1 of U
                    if (!upper) { chu.c.push_back(c1); chu.s.push_back(-sn1); }
                    else { chu.c.push_back(c2); chu.s.push_back(-sn2); }
                }
                diag[k] = s00;                                                              // :143-151
                diag[k + 1] = s11;
                off[k] = s01;
                if (k != n - 1) off[k + 1] = s12;
                vec0 = s01;
                vec1 = s02;
            }
            if (cu) au->push(std::move(chu));
            if (cv) av->push(std::move(chv));
        } else if (subdim == 2) {                                                           // :158-196
            const T m11 = diag[start], m12 = off[start], m22 = diag[start + 1];
            const bool want_u2 = (cu && upper) || (cv && !upper), want_v2 = (cv && upper) || (cu && !upper);
            // compute_2x2_uptrig_svd, svd.rs:375-412
            const T denom = std::hypot(m11 + m22, m12) + std::hypot(m11 - m22, m12);
            T v1 = m11 * m22 * T(2) / denom, v2 = denom / T(2);
            T uc = T(1), us = T(0), vc = T(1), vs = T(0);
            if (want_v2 || want_u2) {
                T sg;
                givens_new<T>(m11 * m12, v1 * v1 - m11 * m11, vc, vs, sg);
                v1 *= sg; v2 *= sg;
                givens_new<T>((m11 * vc + m12 * vs) / v1, (m22 * vs) / v1, uc, us, sg);
                v1 *= sg; v2 *= sg;
            }
            diag[start] = v1;
            diag[start + 1] = v2;
            off[start] = T(0);
            const T ruc = upper ? uc : vc, rus = upper ? us : vs, rvc = upper ? vc : uc, rvs = upper ? vs : us;   // :170-174
            if (cu) {                                                                       // rot_u.rotate_rows(u[.., start..start+2])
                Chain<T> ch; ch.p = start; ch.c.push_back(ruc); ch.s.push_back(rus);
                au->push(std::move(ch));
            }
            if (cv) {                                                                       // rot_v.inverse().rotate_cols(vt[start..start+2, ..])
                Chain<T> ch; ch.p = start; ch.c.push_back(rvc); ch.s.push_back(rvs);
                av->push(std::move(ch));
            }
            end -= 1;
        }
        svd_delimit<T>(st, end, eps, start, end);                                           // :198-208
    }
    for (int64_t i = 0; i < dim; ++i) diag[i] *= amax;                                      // :210
    for (int64_t i = 0; i < dim; ++i) {                                                     // :213-221
        const T val = diag[i];
        if (std::signbit(val)) {
            diag[i] = -val;
            if (cu) au->scale_col(i, -T(0));     // sic: the reference multiplies the column by `-A::zero()`
        }
    }
    if (cu) au->flush();
    if (cv) av->flush();
    LFB_CUDA(cudaStreamSynchronize(h.stream));
    for (int64_t i = 0; i < dim; ++i) sv[i] = diag[i];
}

template void svd_dev<float>(lfb_handle &, float *, int64_t, int64_t, int64_t, float *, float *, int64_t, float *, int64_t);
template void svd_dev<double>(lfb_handle &, double *, int64_t, int64_t, int64_t, double *, double *, int64_t, double *, int64_t);

}  // namespace lfb
