// Blocked Cholesky (lower, column-major) and triangular solves.
//
// Replaces src/cholesky.rs:51-83 (cholesky_inplace_dirty / cholesky_inplace) and
// src/triangular.rs:95-144 (solve_triangular_system).  The reference's row-by-row triple loop is
// restructured right-looking with panels of `chol_nb` columns:
//     L_kk = potrf(A_kk)  (recursive down to 64x64 single-CTA blocks)
//     P    = A[k+nb:, k:k+nb] L_kk^-T          (recursive TRSM: 64-wide substitutions + GEMM)
//     A22 -= P P^T  (lower triangle only)       (one large-K tensor-core GEMM per panel)
// so that ~95 % of the flops are DMMA GEMMs with K = chol_nb.  Only the lower triangle is read or
// written (the `dirty` contract, cholesky.rs:17-19).  A row-major ndarray lower factor is the
// transpose of this layout; api.cu handles that.
#include "common.cuh"

namespace lfb {
namespace {

constexpr int CB = 64;  // diagonal block size of the base kernels

template <typename T> __device__ __forceinline__ T t_sqrt(T x);
template <> __device__ __forceinline__ double t_sqrt<double>(double x) { return sqrt(x); }
template <> __device__ __forceinline__ float t_sqrt<float>(float x) { return sqrtf(x); }

// Unblocked right-looking Cholesky of an n x n (n <= 64) lower block: two warps, thread r keeps
// row r in registers, column j is broadcast through shared memory.
// info[0]: 0 = ok so far, else (failing global row + 1).  cholesky.rs:69-71: pivot <= 0 fails,
// NaN pivots are NOT reported.  On failure the block is written back partially factored.
template <typename T>
__global__ void __launch_bounds__(64) potf2_kernel(T *A, int64_t ld, int n, int64_t row0, int64_t *info) {
    __shared__ T col[CB];
    __shared__ T sdj;
    __shared__ int sfail;
    if (*info != 0) return;
    const int r = threadIdx.x;
    T a[CB];
#pragma unroll
    for (int k = 0; k < CB; ++k) a[k] = (r < n && k <= r) ? A[r + (int64_t)k * ld] : T(0);
    if (r == 0) sfail = 0;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < CB; ++j) {
        if (j < n) {
            if (r == j) {
                T d = a[j];
                if (d <= T(0)) {
                    sfail = j + 1;
                } else {
                    T dj = t_sqrt(d);
                    a[j] = dj;
                    sdj = dj;
                }
            }
            __syncthreads();
            if (sfail) break;
            T l = T(0);
            if (r > j) {
                l = a[j] / sdj;
                a[j] = l;
                col[r] = l;
            }
            __syncthreads();
            if (r > j) {
#pragma unroll
                for (int k = j + 1; k < CB; ++k)
                    if (k <= r) a[k] -= l * col[k];
            }
        }
    }
    if (r == 0 && sfail) *info = row0 + sfail;
#pragma unroll
    for (int k = 0; k < CB; ++k)
        if (r < n && k <= r) A[r + (int64_t)k * ld] = a[k];
}

// Base triangular solve for independent vectors against an nb x nb (nb <= 64) coefficient matrix,
// one vector per thread, right-looking (axpy) form so that the 63-k updates after each unknown are
// independent FMAs:   x_k = acc_k / M(k,k);  acc_i -= M(i,k) x_k  for the unsolved equations i.
//   M(j,i) = tri[j*sj + i*si]   (caller encodes lower/upper and transposition in the strides)
//   element j of vector v lives at B[v*sv + j*sb]
//   FORWARD: k ascending (uses i > k) ; else k descending (uses i < k).
template <typename T, bool FORWARD>
__global__ void __launch_bounds__(128) trsv_block_kernel(const T *__restrict__ tri, int64_t sj, int64_t si, int nb,
                                                         const T *__restrict__ ext_diag, T *B, int64_t sv, int64_t sb,
                                                         int64_t nvec, const int64_t *info) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *sMT = reinterpret_cast<T *>(smem_raw);  // [CB][CB]: sMT[k*CB + i] = M(i,k)
    T *sB = sMT + CB * CB;                     // [CB][129]
    constexpr int LDB = 129;
    if (info && *info != 0) return;
    const int tid = threadIdx.x;
    const int64_t v0 = (int64_t)blockIdx.x * 128;
    const int64_t nleft = nvec - v0;
    const int nv = nleft < 128 ? (int)nleft : 128;
    for (int e = tid; e < CB * CB; e += 128) {
        int j, i;  // equation j, unknown i
        if (si == 1) { i = e % CB; j = e / CB; } else { j = e % CB; i = e / CB; }
        T val = (i == j) ? T(1) : T(0);
        if (i < nb && j < nb) {
            bool used = FORWARD ? (i < j) : (i > j);
            if (used) val = tri[j * sj + i * si];
            else if (i == j) val = ext_diag ? ext_diag[j] : tri[j * sj + i * si];
            else val = T(0);
        }
        sMT[i * CB + j] = val;
    }
    for (int e = tid; e < CB * 128; e += 128) {
        int j, v;
        if (sb == 1) { j = e % CB; v = e / CB; } else { v = e % 128; j = e / 128; }
        T val = T(0);
        if (j < nb && v < nv) val = B[(v0 + v) * sv + j * sb];
        sB[j * LDB + v] = val;
    }
    __syncthreads();
    if (tid < nv) {
        T acc[CB];
#pragma unroll
        for (int j = 0; j < CB; ++j) acc[j] = sB[j * LDB + tid];
        if (FORWARD) {
#pragma unroll
            for (int k = 0; k < CB; ++k) {
                const T x = acc[k] / sMT[k * CB + k];
                acc[k] = x;
#pragma unroll
                for (int i = k + 1; i < CB; ++i) acc[i] -= sMT[k * CB + i] * x;
            }
        } else {
#pragma unroll
            for (int k = CB - 1; k >= 0; --k) {
                const T x = acc[k] / sMT[k * CB + k];
                acc[k] = x;
#pragma unroll
                for (int i = 0; i < k; ++i) acc[i] -= sMT[k * CB + i] * x;
            }
        }
#pragma unroll
        for (int j = 0; j < CB; ++j) sB[j * LDB + tid] = acc[j];
    }
    __syncthreads();
    for (int e = tid; e < CB * 128; e += 128) {
        int j, v;
        if (sb == 1) { j = e % CB; v = e / CB; } else { v = e % 128; j = e / 128; }
        if (j < nb && v < nv) B[(v0 + v) * sv + j * sb] = sB[j * LDB + v];
    }
}

template <typename T>
void trsv_block(lfb_handle &h, bool forward, const T *tri, int64_t sj, int64_t si, int nb, const T *ext_diag, T *B,
                int64_t sv, int64_t sb, int64_t nvec, const int64_t *info) {
    if (nvec <= 0 || nb <= 0) return;
    size_t smem = sizeof(T) * (CB * CB + CB * 129);
    static bool cfg = false;
    if (!cfg) {
        LFB_CUDA(cudaFuncSetAttribute(trsv_block_kernel<T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        LFB_CUDA(cudaFuncSetAttribute(trsv_block_kernel<T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cfg = true;
    }
    unsigned grid = (unsigned)cdiv(nvec, 128);
    if (forward) trsv_block_kernel<T, true><<<grid, 128, smem, h.stream>>>(tri, sj, si, nb, ext_diag, B, sv, sb, nvec, info);
    else trsv_block_kernel<T, false><<<grid, 128, smem, h.stream>>>(tri, sj, si, nb, ext_diag, B, sv, sb, nvec, info);
    LFB_LAUNCH_CHECK(h);
}

inline int64_t split_point(int64_t n) {  // n > CB: CB-aligned split near the middle
    int64_t n1 = round_up(n / 2, CB);
    if (n1 >= n) n1 = ((n - 1) / CB) * CB;
    return n1;
}

// B (rows x n) <- B L^-T, L lower n x n column-major.
template <typename T>
void trsm_right_lt(lfb_handle &h, int64_t rows, int64_t n, const T *L, int64_t ldl, T *B, int64_t ldb, const int64_t *info) {
    if (rows <= 0 || n <= 0) return;
    if (n <= CB) {
        // row x solves sum_{i<=j} x_i L[j][i] = b_j : M(j,i) = L[j + i*ldl]
        trsv_block<T>(h, true, L, 1, ldl, (int)n, nullptr, B, /*sv=*/1, /*sb=*/ldb, rows, info);
        return;
    }
    int64_t n1 = split_point(n), n2 = n - n1;
    trsm_right_lt<T>(h, rows, n1, L, ldl, B, ldb, info);
    gemm<T>(h, 0, 1, rows, n2, n1, T(-1), B, ldb, L + n1, ldl, T(1), B + n1 * ldb, ldb);   // B2 -= X1 L21^T
    trsm_right_lt<T>(h, rows, n2, L + n1 + n1 * ldl, ldl, B + n1 * ldb, ldb, info);
}

template <typename T>
void potrf_rec(lfb_handle &h, T *A, int64_t n, int64_t ld, int64_t row0, int64_t *info) {
    if (n <= 0) return;
    if (n <= CB) {
        potf2_kernel<T><<<1, 64, 0, h.stream>>>(A, ld, (int)n, row0, info);
        LFB_LAUNCH_CHECK(h);
        return;
    }
    int64_t n1 = split_point(n), n2 = n - n1;
    potrf_rec<T>(h, A, n1, ld, row0, info);
    trsm_right_lt<T>(h, n2, n1, A, ld, A + n1, ld, info);
    gemm<T>(h, 0, 1, n2, n2, n1, T(-1), A + n1, ld, A + n1, ld, T(1), A + n1 + n1 * ld, ld, /*lower_only=*/1);
    potrf_rec<T>(h, A + n1 + n1 * ld, n2, ld, row0 + n1, info);
}

}  // namespace

template <typename T>
void cholesky_lower(lfb_handle &h, T *A, int64_t n, int64_t ld, int clean, int64_t *d_info) {
    LFB_CUDA(cudaMemsetAsync(d_info, 0, sizeof(int64_t), h.stream));
    if (n <= 0) return;
    const int64_t NB = std::max<int64_t>(CB, round_up(h.opt.chol_nb, CB));
    for (int64_t k0 = 0; k0 < n; k0 += NB) {
        const int64_t nb = std::min<int64_t>(NB, n - k0);
        T *Akk = A + k0 + k0 * ld;
        potrf_rec<T>(h, Akk, nb, ld, k0, d_info);
        const int64_t rows = n - k0 - nb;
        if (rows > 0) {
            T *P = Akk + nb;
            trsm_right_lt<T>(h, rows, nb, Akk, ld, P, ld, d_info);
            gemm<T>(h, 0, 1, rows, rows, nb, T(-1), P, ld, P, ld, T(1), A + (k0 + nb) + (k0 + nb) * ld, ld, /*lower_only=*/1);
        }
    }
    if (clean) triangular_zero<T>(h, A, n, ld, /*keep_lower=*/1);
}

// op(A) X = B in place on B (n x nrhs), recursive with GEMM updates.
template <typename T>
void trsm_left(lfb_handle &h, int lower, int trans, int64_t n, int64_t nrhs, const T *A, int64_t lda, const T *ext_diag,
               T *B, int64_t ldb) {
    if (n <= 0 || nrhs <= 0) return;
    // equation j: sum_i op(A)[j][i] x_i = b_j ; op(A)[j][i] = trans ? A[i + j*lda] : A[j + i*lda]
    const bool forward = (lower != 0) != (trans != 0);  // lower-N and upper-T are forward substitutions
    if (n <= CB) {
        int64_t sj = trans ? lda : 1, si = trans ? 1 : lda;
        trsv_block<T>(h, forward, A, sj, si, (int)n, ext_diag, B, /*sv=*/ldb, /*sb=*/1, nrhs, nullptr);
        return;
    }
    int64_t n1 = split_point(n), n2 = n - n1;
    const T *A11 = A, *A22 = A + n1 + n1 * lda;
    const T *A21 = A + n1, *A12 = A + n1 * lda;
    T *B1 = B, *B2 = B + n1;
    const T *d1 = ext_diag, *d2 = ext_diag ? ext_diag + n1 : nullptr;
    if (forward) {
        trsm_left<T>(h, lower, trans, n1, nrhs, A11, lda, d1, B1, ldb);
        if (lower) gemm<T>(h, 0, 0, n2, nrhs, n1, T(-1), A21, lda, B1, ldb, T(1), B2, ldb);   // B2 -= A21 X1
        else gemm<T>(h, 1, 0, n2, nrhs, n1, T(-1), A12, lda, B1, ldb, T(1), B2, ldb);        // B2 -= A12^T X1
        trsm_left<T>(h, lower, trans, n2, nrhs, A22, lda, d2, B2, ldb);
    } else {
        trsm_left<T>(h, lower, trans, n2, nrhs, A22, lda, d2, B2, ldb);
        if (lower) gemm<T>(h, 1, 0, n1, nrhs, n2, T(-1), A21, lda, B2, ldb, T(1), B1, ldb);   // B1 -= A21^T X2
        else gemm<T>(h, 0, 0, n1, nrhs, n2, T(-1), A12, lda, B2, ldb, T(1), B1, ldb);        // B1 -= A12 X2
        trsm_left<T>(h, lower, trans, n1, nrhs, A11, lda, d1, B1, ldb);
    }
}

#define INST(T)                                                                                       \
    template void cholesky_lower<T>(lfb_handle &, T *, int64_t, int64_t, int, int64_t *);             \
    template void trsm_left<T>(lfb_handle &, int, int, int64_t, int64_t, const T *, int64_t, const T *, T *, int64_t);
INST(float)
INST(double)
#undef INST

}  // namespace lfb
