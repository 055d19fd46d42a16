// Householder reconstruction: tall-skinny QR through TSQR, delivered in the reference's compact form.
//
// SURVEY.md 8(f) rank 2.  TSQR (householder.cu: tsqr_local_r / tsqr_explicit_q) is how a tall-skinny matrix is factored
// at the speed of the hardware -- independent row chunks, one small tree -- but what it produces is (Q, R), while
// the reference's callers hold a QRDecomp{qr, diag} (qr.rs:68-73): unit-norm reflectors v_k below the diagonal
// (householder.rs:9-28), the sign-scaled rows of R above it (householder.rs:45-48), the signed pivots in `diag`, and
// generate_q / qt_mul / solve_into (qr.rs:86-152) all read that form.  This file converts one into the other:
//
//   Q - [S; 0] = Y U   (LU without pivoting of the explicit thin Q, S = diag(s_k), s_k = -sgn of the pivot met at
//                       step k; |pivot| = 1 + |q_kk| >= 1, so no pivoting is ever needed)
//
// Y (unit lower trapezoidal) holds the Householder vectors in LAPACK's v[0] = 1 normalisation, tau_k = |U_kk|, and
// beta_k = s_k R_kk is the pivot a Householder QR of the same matrix would have produced -- exactly, not just up to
// signs -- so the reference's quantities follow from SURVEY.md 8(a)'s conversion (P_k = sgn(beta_k) = s_k):
//
//   v_ref,k = c_k y_k,  c_k = -s_{k-1} s_k sqrt(|U_kk| / 2)        (s_{-1} = 1)
//   diag_ref[k] = s_{k-1} s_k R_kk,   R_ref[k, j>k] = R[k, j]      (R has diag >= 0, qr.rs:96)
//
// The rows of Y below the top block are Q2 U^-1; scaling by c_k is folded into the solve (U' = diag(1/c) U), so the
// tall part is touched exactly once more, by one blocked right-hand TRSM (trsm.cu: trsm_right_upper).
// Validated against the CPU oracle's qr (elementwise) in NumPy before any CUDA was written.
#include "common.cuh"

namespace lfb {
namespace {

inline unsigned ycap(int64_t n) { return (unsigned)(n < 1 ? 1 : (n < 65535 ? n : 65535)); }

// In-place LU of (Q1 - S) on one CTA (n is the SKINNY dimension; n^3/3 flops, all latency).  Column-major, global
// memory (256 x 256 f64 does not fit shared memory; the block stays L1/L2 resident).  Warp w owns trailing columns
// k+1+w, k+1+w+32, ...; lanes run down the rows, so every access is coalesced.  Plain loads/stores only: the data
// is rewritten by other warps of the CTA between barriers.
template <typename T>
__global__ void __launch_bounds__(1024) hr_lu_kernel(T *Q, int64_t ld, int n, T *s) {
    __shared__ T s_piv;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int k = 0; k < n; ++k) {
        T *colk = Q + (int64_t)k * ld;
        if (tid == 0) {
            const T q = colk[k];
            const T sk = q < T(0) ? T(1) : T(-1);   // s_k = -sgn(q_kk), sgn(0) = +1
            const T p = q - sk;                     // |p| = 1 + |q_kk|
            colk[k] = p;
            s[k] = sk;
            s_piv = p;
        }
        __syncthreads();
        const T p = s_piv;
        for (int i = k + 1 + tid; i < n; i += 1024) colk[i] = colk[i] / p;
        __syncthreads();
        for (int j = k + 1 + warp; j < n; j += 32) {
            T *colj = Q + (int64_t)j * ld;
            const T ukj = colj[k];
            for (int i = k + 1 + lane; i < n; i += 32) colj[i] -= colk[i] * ukj;
        }
        __syncthreads();
    }
}

// Blocked version of the same LU (n <= 512), one CTA: 32-column panels live in shared memory while they are factored
// (three CTA barriers per pivot on shared data instead of three round trips to L2 per pivot), the block row U12 = L11^-1 A12
// is solved from shared memory, and the trailing update A22 -= L21 U12 streams the rest of the matrix through the SM once
// per PANEL instead of once per pivot: 89 MB -> ~3 MB of L2 traffic at n = 256.  1.75 ms -> ~0.2 ms
// (profiles/r2_multi_gpu.md); the arithmetic per entry is the same right-looking elimination, s_k = -sgn(q_kk) as before.
constexpr int HR_PB = 32;
template <typename T>
__global__ void __launch_bounds__(1024) hr_lu_blocked_kernel(T *Q, int64_t ld, int n, T *s) {
    extern __shared__ __align__(16) unsigned char hr_smem[];
    T *P = reinterpret_cast<T *>(hr_smem);            // panel: [n rows][HR_PB + 1]
    T *U = P + (size_t)n * (HR_PB + 1);               // block row: [HR_PB][n + 1]
    __shared__ T s_piv;
    const int tid = threadIdx.x;
    for (int k0 = 0; k0 < n; k0 += HR_PB) {
        const int kb = min(HR_PB, n - k0), m = n - k0;           // panel is m x kb
        for (int e = tid; e < m * kb; e += 1024) {
            const int r = e % m, c = e / m;
            P[r * (HR_PB + 1) + c] = Q[(k0 + r) + (int64_t)(k0 + c) * ld];
        }
        __syncthreads();
        for (int j = 0; j < kb; ++j) {
            if (tid == 0) {
                const T q = P[j * (HR_PB + 1) + j];
                const T sk = q < T(0) ? T(1) : T(-1);            // s_k = -sgn(q_kk), sgn(0) = +1
                const T p = q - sk;                              // |p| = 1 + |q_kk|
                P[j * (HR_PB + 1) + j] = p;
                s[k0 + j] = sk;
                s_piv = p;
            }
            __syncthreads();
            const T p = s_piv;
            for (int r = j + 1 + tid; r < m; r += 1024) P[r * (HR_PB + 1) + j] /= p;
            __syncthreads();
            const int w = kb - j - 1, hgt = m - j - 1;
            for (int e = tid; e < w * hgt; e += 1024) {
                const int r = j + 1 + e % hgt, c = j + 1 + e / hgt;
                P[r * (HR_PB + 1) + c] -= P[r * (HR_PB + 1) + j] * P[j * (HR_PB + 1) + c];
            }
            __syncthreads();
        }
        for (int e = tid; e < m * kb; e += 1024) {               // the finished panel goes back
            const int r = e % m, c = e / m;
            Q[(k0 + r) + (int64_t)(k0 + c) * ld] = P[r * (HR_PB + 1) + c];
        }
        const int nc = n - k0 - kb;                              // columns right of the panel
        if (nc > 0) {
            for (int e = tid; e < kb * nc; e += 1024) {          // A12 -> shared
                const int j = e % kb, c = e / kb;
                U[j * (n + 1) + c] = Q[(k0 + j) + (int64_t)(k0 + kb + c) * ld];
            }
            __syncthreads();
            for (int c = tid; c < nc; c += 1024)                 // U12 = L11^-1 A12 (unit lower), one column per thread
                for (int j = 1; j < kb; ++j) {
                    T acc = U[j * (n + 1) + c];
                    for (int i = 0; i < j; ++i) acc -= P[j * (HR_PB + 1) + i] * U[i * (n + 1) + c];
                    U[j * (n + 1) + c] = acc;
                }
            __syncthreads();
            for (int e = tid; e < kb * nc; e += 1024) {
                const int j = e % kb, c = e / kb;
                Q[(k0 + j) + (int64_t)(k0 + kb + c) * ld] = U[j * (n + 1) + c];
            }
            const int mr = m - kb;                               // rows below the panel's diagonal block
            for (int e = tid; e < mr * nc; e += 1024) {          // A22 -= L21 U12
                const int r = e % mr, c = e / mr;
                const T *l = P + (size_t)(kb + r) * (HR_PB + 1);
                T acc = T(0);
#pragma unroll 8
                for (int i = 0; i < kb; ++i) acc += l[i] * U[i * (n + 1) + c];
                Q[(k0 + kb + r) + (int64_t)(k0 + kb + c) * ld] -= acc;
            }
        }
        __syncthreads();
    }
}

// c_k and diag_ref[k] from the LU pivots, the signs and diag(R).
// internal = 1: the blocked QR's own convention instead (householder.cu: unit-norm reflectors without the running sign,
// beta_k = s_k R_kk, R rows as the reflections leave them = s_k R[k, :]); the driver's final sign pass turns that into the above.
template <typename T>
__global__ void hr_scale_kernel(const T *Q, int64_t ld, int n, const T *R, int64_t ldr, const T *s, T *c, T *diag, int internal) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const T sp = (k > 0 && !internal) ? s[k - 1] : T(1);
    T u = Q[k + (int64_t)k * ld];
    u = u < T(0) ? -u : u;
    c[k] = -sp * s[k] * sqrt(u / T(2));
    T rkk = R[k + (int64_t)k * ldr];
    rkk = fabs(rkk);                                 // |.| also clears the sign bit of a -0.0 pivot
    diag[k] = sp * s[k] * rkk;                       // R_kk = 0 keeps the sign bit of s_{k-1} s_k: Rust's signum reads it
                                                     // (householder.rs:45, :89), and Q = [I; 0] for A = 0 depends on it
}

// U' = diag(1/c) U out of the upper triangle, then the top block's final contents: c_k on the diagonal (head of
// v_ref,k), c_k y_ik below, R[k, j>k] above.  Every thread reads only its own entry of Q.
template <typename T>
__global__ void hr_finish_kernel(T *Q, int64_t ld, int n, const T *R, int64_t ldr, const T *c, T *U, int64_t ldu, const T *s_int) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    for (int j = blockIdx.y; j < n; j += gridDim.y) {
        const T q = Q[i + (int64_t)j * ld];
        if (i < j) {
            U[i + (int64_t)j * ldu] = q / c[i];
            const T r = R[i + (int64_t)j * ldr];
            Q[i + (int64_t)j * ld] = (s_int && s_int[i] < T(0)) ? -r : r;
        } else if (i == j) {
            U[i + (int64_t)j * ldu] = q / c[i];
            Q[i + (int64_t)j * ld] = c[i];
        } else {
            U[i + (int64_t)j * ldu] = T(0);
            Q[i + (int64_t)j * ld] = q * c[j];
        }
    }
}

}  // namespace

// Qtop: the first n rows of an explicit thin Q (n columns), overwritten with the top n x n block of the reference's
// compact factor; R (n x n upper, diag >= 0) its triangular factor; U (n x n) receives U' for the rows below
// (Y2' = Q2 U'^-1); diag[n] the signed pivots.
template <typename T>
void hh_reconstruct_top(lfb_handle &h, T *Qtop, int64_t n, int64_t ld, const T *R, int64_t ldr, T *U, int64_t ldu, T *diag, int internal) {
    if (n <= 0) return;
    DevBuf<T> s(h, n), c(h, n);
    const size_t smem_lu = sizeof(T) * ((size_t)n * (HR_PB + 1) + (size_t)HR_PB * (n + 1));
    if (h.opt.hr_lu_blocked && n <= 512 && smem_lu + 1024 <= h.smem_optin) {
        static DeviceOnce cfg;   // function attributes are per device
        cfg.run(h.device, [&] {
            LFB_CUDA(cudaFuncSetAttribute(hr_lu_blocked_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h.smem_optin - 1024));   // minus the static s_piv
        });
        hr_lu_blocked_kernel<T><<<1, 1024, smem_lu, h.stream>>>(Qtop, ld, (int)n, s);
    } else {
        hr_lu_kernel<T><<<1, 1024, 0, h.stream>>>(Qtop, ld, (int)n, s);
    }
    LFB_LAUNCH_CHECK(h);
    hr_scale_kernel<T><<<(unsigned)cdiv(n, 256), 256, 0, h.stream>>>(Qtop, ld, (int)n, R, ldr, s, c, diag, internal);
    LFB_LAUNCH_CHECK(h);
    dim3 grid((unsigned)cdiv(n, 256), ycap(n));
    hr_finish_kernel<T><<<grid, 256, 0, h.stream>>>(Qtop, ld, (int)n, R, ldr, c, U, ldu, internal ? s.get() : (const T *)nullptr);
    LFB_LAUNCH_CHECK(h);
}

// qr.rs:29-45 qr_into for a tall-skinny A (rows >> cols): same contract as qr_factor -- A becomes the compact factor,
// diag the signed pivots -- computed as TSQR (explicit Q) + Householder reconstruction.  3x the flops of the R-only
// TSQR (factor, assemble, combine: 2 m n^2 each) plus m n^2 for the solve, all of it chunk-parallel or tensor-core GEMM.
template <typename T>
void qr_tsqr(lfb_handle &h, T *A, int64_t rows, int64_t cols, int64_t ld, T *diag) {
    if (cols <= 0 || rows <= 0) return;
    const int64_t ldw = round_up(rows, 2), ldu = round_up(cols, 2);
    DevBuf<T> R(h, (size_t)ldu * cols), U(h, (size_t)ldu * cols);
    // Folded route on top of the Cholesky-QR leaf: the explicit Q is never formed below the top block.  With Q = A R^-1,
    //   Q_top = A_top R^-1 (n^3)  ->  LU of (Q_top - S) gives U' and diag  ->  rows below: Y2 = Q2 U'^-1 = A2 (R^-1 U'^-1),
    // so the tall part is touched by exactly ONE GEMM (m n^2 * 2 flops) after the Gram GEMM: 3.5 m n^2 flops in all against
    // 7 m n^2 for factor + assemble + combine + solve, all of it on the tensor pipe with K >= n.
    if (rows >= 4 * cols && cols <= 512) {
        DevBuf<T> Rinv(h, (size_t)ldu * cols);
        if (cholqr_factor<T>(h, A, rows, cols, ld, R.get(), ldu, Rinv.get(), ldu)) {
            DevBuf<T> Qt(h, (size_t)ldu * cols);
            gemm<T>(h, 0, 0, cols, cols, cols, T(1), A, ld, Rinv.get(), ldu, T(0), Qt.get(), ldu);       // Q_top
            hh_reconstruct_top<T>(h, Qt.get(), cols, ldu, R.get(), ldu, U.get(), ldu, diag);             // Qt <- top block of the factor
            trsm_right_upper<T>(h, cols, cols, U.get(), ldu, Rinv.get(), ldu);                           // Rinv <- R^-1 U'^-1
            tsqr_apply_q<T>(h, A + cols, rows - cols, cols, ld, Rinv.get(), ldu);                        // A2 <- A2 (R^-1 U'^-1)
            copy2d<T>(h, Qt.get(), ldu, A, ld, cols, cols);
            return;
        }
    }
    {
        DevBuf<T> Wk(h, (size_t)ldw * cols);
        tsqr_explicit_q<T>(h, A, rows, cols, ld, Wk, ldw, R, ldu);
    }
    hh_reconstruct_top<T>(h, A, cols, ld, R.get(), ldu, U.get(), ldu, diag);
    trsm_right_upper<T>(h, rows - cols, cols, U.get(), ldu, A + cols, ld);
}

#define INST(T)                                                                                                        \
    template void hh_reconstruct_top<T>(lfb_handle &, T *, int64_t, int64_t, const T *, int64_t, T *, int64_t, T *, int); \
    template void qr_tsqr<T>(lfb_handle &, T *, int64_t, int64_t, int64_t, T *);
INST(float)
INST(double)
#undef INST

}  // namespace lfb
