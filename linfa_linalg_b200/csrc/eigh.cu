// Symmetric eigendecomposition: src/eigh.rs:10-129 (symmetric_eig) end to end.
//
// Phase 1 (scale by max|a|, tridiagonalise, generate Q) runs on the kernels of tridiag.cu / householder.cu.
// Phase 2, the implicit symmetric QR iteration with Wilkinson shifts (eigh.rs:51-128), is O(n^2) SCALAR work
// on the tridiagonal (d, e) -- a strictly sequential bulge chase -- plus O(n^3) rotation work on Q.  Here the
// scalar recurrence runs on the host, exactly as the reference orders it, and emits "chains": runs of Givens
// rotations acting on consecutive column pairs (i, i+1), i = p .. p+cnt-1 (one chain per QR sweep, a chain of
// one for a deflated 2x2 block).  Every row of Q sees the same sequence of rotations, so the device applies
// them row-parallel.  K chains are applied in ONE pass over Q as a wavefront: chain s runs two columns
// behind chain s-1, a thread keeps a sliding window of 2K row entries in registers, each element of Q is
// read and written once per K sweeps (the reference touches it once per sweep), and the host is already
// chasing the next K sweeps while the device applies the previous ones.
#include "rotations.cuh"

namespace lfb {

namespace {

// eigh.rs:131-170
template <typename T>
inline void delimit_subproblem(const std::vector<T> &diag, std::vector<T> &off, int64_t end, T eps, int64_t &start, int64_t &nend) {
    int64_t n = end;
    while (n > 0) {
        const int64_t m = n - 1;
        if (std::fabs(off[m]) > eps * (std::fabs(diag[n]) + std::fabs(diag[m]))) break;
        n -= 1;
    }
    if (n == 0) { start = 0; nend = 0; return; }
    int64_t ns = n - 1;
    while (ns > 0) {
        const int64_t m = ns - 1;
        if (off[m] == T(0) || std::fabs(off[m]) <= eps * (std::fabs(diag[ns]) + std::fabs(diag[m]))) {
            off[m] = T(0);
            break;
        }
        ns -= 1;
    }
    start = ns; nend = n;
}

}  // namespace

// eigh.rs:10-129.  dA (n x n column-major, ld) is consumed.  vals: HOST array of n entries (reference order,
// unsorted).  dQ: device n x n (ldq) for the eigenvectors as columns, or nullptr (eigvalsh).
template <typename T>
void symmetric_eig(lfb_handle &h, T *dA, int64_t n, int64_t ld, T *vals, T *dQ, int64_t ldq) {
    if (n < 1) return;                                                          // :16-25
    FAST_HYPOT = h.opt.fast_hypot != 0;
    const T eps = std::numeric_limits<T>::epsilon();                            // :214 (A::epsilon())
    DevBuf<T> scal(h, 2), dOff(h, n), dDiag(h, n);
    LFB_CUDA(cudaMemsetAsync(scal.get(), 0, 2 * sizeof(T), h.stream));
    {
        dim3 grid((unsigned)std::min<int64_t>(cdiv(n, 256), 64), (unsigned)std::min<int64_t>(n, 1024));
        absmax_kernel<T><<<grid, 256, 0, h.stream>>>(dA, ld, n, n, scal.get());            // :27-30
        LFB_LAUNCH_CHECK(h);
        scale_div_kernel<T><<<grid, 256, 0, h.stream>>>(dA, ld, n, n, scal.get());          // :32-34
        LFB_LAUNCH_CHECK(h);
    }
    sym_tridiagonal<T>(h, dA, n, ld, dOff.get());                                             // :36
    if (dQ) assemble_q<T>(h, dA, n, n, ld, 1, dOff.get(), dQ, ldq);                           // :37-41 (tridiagonal.rs:90-92)
    extract_diag_kernel<T><<<(unsigned)cdiv(n, 256), 256, 0, h.stream>>>(dA, ld, n, dDiag.get());
    LFB_LAUNCH_CHECK(h);
    std::vector<T> diag(n), off(std::max<int64_t>(n - 1, 1));
    T amax = T(0);
    LFB_CUDA(cudaMemcpyAsync(diag.data(), dDiag.get(), sizeof(T) * n, cudaMemcpyDeviceToHost, h.stream));
    if (n > 1) LFB_CUDA(cudaMemcpyAsync(off.data(), dOff.get(), sizeof(T) * (n - 1), cudaMemcpyDeviceToHost, h.stream));
    LFB_CUDA(cudaMemcpyAsync(&amax, scal.get(), sizeof(T), cudaMemcpyDeviceToHost, h.stream));
    LFB_CUDA(cudaStreamSynchronize(h.stream));
    for (int64_t i = 0; i + 1 < n; ++i) off[i] = std::fabs(off[i]);                          // tridiagonal.rs:96-101
    if (n == 1) {                                                                            // :44-47
        vals[0] = diag[0] * amax;
        return;
    }
    std::unique_ptr<RotationApplier<T>> app;
    if (dQ) app.reset(new RotationApplier<T>(h, dQ, ldq, n));

    const auto t_loop0 = std::chrono::steady_clock::now();
    int64_t start, end;
    delimit_subproblem<T>(diag, off, n - 1, eps, start, end);                                // :49
    while (end != start) {                                                                   // :51
        const int64_t subdim = end - start + 1;
        if (subdim > 2) {                                                                    // :55
            const int64_t m = end - 1, nn = end;
            T x = diag[start] - wilkinson_shift<T>(diag[m], diag[nn], off[m]);               // :59
            T y = off[start];
            Chain<T> ch;
            ch.p = start;
            if (dQ) { ch.c.reserve(nn - start); ch.s.reserve(nn - start); }
            for (int64_t i = start; i < nn; ++i) {                                           // :62
                const int64_t j = i + 1;
                if (y == T(0)) break;                                                        // givens.rs:18 -> :96-98
                const T r = h_hypot(x, y), c = x / r, s = -y / r;                            // givens.rs:19-21
                if (i > start) off[i - 1] = r;                                               // :66-68
                const T cc = c * c, ss = s * s, cs = c * s;
                const T mii = diag[i], mjj = diag[j], mij = off[i];
                const T b = cs * mij * T(2);
                diag[i] = cc * mii + ss * mjj - b;                                           // :78-80
                diag[j] = ss * mii + cc * mjj + b;
                off[i] = cs * (mii - mjj) + mij * (cc - ss);
                if (i != nn - 1) {                                                           // :82-86
                    x = off[i];
                    y = -s * off[i + 1];
                    off[i + 1] *= c;
                }
                if (dQ) { ch.c.push_back(c); ch.s.push_back(-s); }                           // :89-94: the inverse rotation
            }
            if (dQ) app->push(std::move(ch));
            if (std::fabs(off[m]) <= eps * (std::fabs(diag[m]) + std::fabs(diag[nn]))) end -= 1;   // :100-102
        } else if (subdim == 2) {                                                            // :103
            const T h00 = diag[start], h10 = off[start], h11 = diag[start + 1];
            const T val = (h00 - h11) * T(0.5);                                              // :188-198 compute_2x2_eigvals
            const T discr = h10 * h10 + val * val;                                           // >= 0 for a symmetric block
            const T sq = std::sqrt(discr), half = (h00 + h11) * T(0.5);
            const T e0 = half + sq, e1 = half - sq;
            T b0 = e0 - diag[start + 1], b1 = off[start];                                    // :111
            if (h.opt.eigh_stable_2x2 && h00 < h11) {
                // (e0 - h11, h10) and (h10, e0 - h00) span the same line ((e0 - h11)(e0 - h00) = h10^2).  For
                // h00 < h11 the reference's e0 - h11 = h10^2 / (e0 - h00) cancels catastrophically when h10 is
                // barely above the deflation threshold: its rounding noise is then as large as h10 itself and
                // the "rotation" mixes two well-separated eigenvectors (seen with the reference's own algorithm
                // on the CPU oracle: residual 0.1 .. 1.7 instead of 3e-10 at n = 600 for ~20 % of 1e-15-sized
                // perturbations of the input).  Same direction and sign, computed without cancellation:
                b0 = std::fabs(h10);
                b1 = h_signum(h10) * (e0 - h00);
            }
            diag[start] = e0;
            diag[start + 1] = e1;
            if (dQ) {                                                                        // :116-121, givens.rs:48-57
                const T norm = std::hypot(b0, b1);
                if (norm > eps) {
                    Chain<T> ch;
                    ch.p = start;
                    ch.c.push_back(b0 / norm);
                    ch.s.push_back(b1 / norm);
                    app->push(std::move(ch));
                }
            }
            end -= 1;
        }
        delimit_subproblem<T>(diag, off, end, eps, start, end);                              // :125-127
    }
    if (dQ) {
        app->flush();
        const auto c0 = std::chrono::steady_clock::now();
        LFB_CUDA(cudaStreamSynchronize(h.stream));
        if (h.opt.trd_profile)
            fprintf(stderr, "[eigh_profile] n=%lld: host recurrence + packing %.3f s (waiting for a staging buffer %.3f s, packing + "
                            "launch %.3f s, %lld passes), final drain of the device queue %.3f s\n", (long long)n,
                    std::chrono::duration<double>(c0 - t_loop0).count(), app->t_wait, app->t_pack, (long long)app->batches,
                    std::chrono::duration<double>(std::chrono::steady_clock::now() - c0).count());
    }
    for (int64_t i = 0; i < n; ++i) vals[i] = diag[i] * amax;                                // :130
}

template void symmetric_eig<float>(lfb_handle &, float *, int64_t, int64_t, float *, float *, int64_t);
template void symmetric_eig<double>(lfb_handle &, double *, int64_t, int64_t, double *, double *, int64_t);

}  // namespace lfb
