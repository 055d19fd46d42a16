// Triangular solves with many right-hand sides (column-major).
//
// Replaces src/triangular.rs:95-144 (solve_triangular_system): the reference's per-RHS column-axpy
// substitution becomes a recursive blocked TRSM -- 64 x 64 diagonal blocks solved by one small
// shared-memory kernel (one right-hand side per thread), everything off the diagonal a tensor-core
// GEMM.  Only the named triangle of A is read (dirty Cholesky factors and compact QR factors are
// legal inputs, cholesky.rs:140-142, qr.rs:145-150); an external diagonal is honoured.
#include "common.cuh"

namespace lfb {
namespace {

constexpr int CB = 64;  // diagonal block size of the base kernels

// Base triangular solve for independent vectors against an nb x nb (nb <= 64) coefficient matrix,
// one vector per thread, blocked 8 x 8 with everything in shared memory and dynamic outer loops so
// the code stays small (instruction-cache resident):
//     for kb:  solve the 8x8 diagonal block (registers), then  acc[ib] -= M(ib,kb) x[kb]  for the
//              unsolved blocks ib (64 independent FMAs each).
//   M(j,i) = tri[j*sj + i*si]   (caller encodes lower/upper and transposition in the strides)
//   element j of vector v lives at B[v*sv + j*sb]
//   FORWARD: blocks ascending ; else descending.  The diagonal is applied as a reciprocal
//   (x = acc * (1/d)), which differs from the reference's division by <= 1 ulp.
template <typename T, bool FORWARD>
__global__ void __launch_bounds__(128) trsv_block_kernel(const T *__restrict__ tri, int64_t sj, int64_t si, int nb,
                                                         const T *__restrict__ ext_diag, T *B, int64_t sv, int64_t sb,
                                                         int64_t nvec, const int64_t *info) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *sM = reinterpret_cast<T *>(smem_raw);   // [CB][CB]: sM[j*CB + i] = M(j,i)  (row j = equation j)
    T *sB = sM + CB * CB;                      // [CB][129]
    T *sInv = sB + CB * 129;                   // [CB] reciprocal diagonal
    constexpr int LDB = 129;
    if (info && *info != 0) return;
    const int tid = threadIdx.x;
    const int64_t v0 = (int64_t)blockIdx.x * 128;
    const int64_t nleft = nvec - v0;
    const int nv = nleft < 128 ? (int)nleft : 128;
#pragma unroll 8
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
        sM[j * CB + i] = val;
        if (i == j) sInv[i] = T(1) / val;
    }
#pragma unroll 16
    for (int e = tid; e < CB * 128; e += 128) {
        int j, v;
        if (sb == 1) { j = e % CB; v = e / CB; } else { v = e % 128; j = e / 128; }
        T val = T(0);
        if (j < nb && v < nv) val = B[(v0 + v) * sv + j * sb];
        sB[j * LDB + v] = val;
    }
    __syncthreads();
    if (tid < nv) {
        T *my = sB + tid;  // my[j*LDB] = element j of this thread's vector
        for (int step = 0; step < CB / 8; ++step) {
            const int kb = FORWARD ? step : CB / 8 - 1 - step;
            T x[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) x[q] = my[(kb * 8 + q) * LDB];
            // 8x8 diagonal block
            if (FORWARD) {
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    x[q] *= sInv[kb * 8 + q];
#pragma unroll
                    for (int t = q + 1; t < 8; ++t) x[t] -= sM[(kb * 8 + t) * CB + kb * 8 + q] * x[q];
                }
            } else {
#pragma unroll
                for (int q = 7; q >= 0; --q) {
                    x[q] *= sInv[kb * 8 + q];
#pragma unroll
                    for (int t = 0; t < q; ++t) x[t] -= sM[(kb * 8 + t) * CB + kb * 8 + q] * x[q];
                }
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) my[(kb * 8 + q) * LDB] = x[q];
            // update the unsolved blocks
            const int ib0 = FORWARD ? kb + 1 : 0, ib1 = FORWARD ? CB / 8 : kb;
            for (int ib = ib0; ib < ib1; ++ib) {
                T acc[8];
#pragma unroll
                for (int t = 0; t < 8; ++t) acc[t] = my[(ib * 8 + t) * LDB];
#pragma unroll
                for (int t = 0; t < 8; ++t) {
                    const T *mrow = sM + (ib * 8 + t) * CB + kb * 8;
#pragma unroll
                    for (int q = 0; q < 8; ++q) acc[t] -= mrow[q] * x[q];
                }
#pragma unroll
                for (int t = 0; t < 8; ++t) my[(ib * 8 + t) * LDB] = acc[t];
            }
        }
    }
    __syncthreads();
#pragma unroll 16
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
    size_t smem = sizeof(T) * (CB * CB + CB * 129 + CB);
    static DeviceOnce cfg;   // function attributes are per device
    cfg.run(h.device, [&] {
        LFB_CUDA(cudaFuncSetAttribute(trsv_block_kernel<T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        LFB_CUDA(cudaFuncSetAttribute(trsv_block_kernel<T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    });
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

}  // namespace

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

// X U = B in place on B (rows x n, column-major) with U upper triangular, given either as U itself (trans_lower = 0:
// U[i, j] = Tri[i + j*ldt], only the upper triangle is read) or as the transpose of a lower-triangular factor
// (trans_lower = 1: U = L^T, U[i, j] = Tri[j + i*ldt], only the lower triangle is read).  The right-hand solves of the
// Householder reconstruction (Y2 = Q2 U^-1, tsqr_hr.cu) and of LOBPCG's orthonormalize (V L^-T, lobpcg/algorithm.rs:91-94).
// Left-looking over 64-column blocks: the finished block columns enter by one GEMM, the diagonal block by the base
// kernel with every ROW of B as one right-hand side (element j of vector v at B[v + j*ldb]).  `info` (device, may be
// null): a non-zero value skips the diagonal-block solves (the factor is not valid).
template <typename T>
void trsm_right(lfb_handle &h, int64_t rows, int64_t n, const T *Tri, int64_t ldt, int trans_lower, T *B, int64_t ldb,
                const int64_t *info) {
    if (rows <= 0 || n <= 0) return;
    for (int64_t j0 = 0; j0 < n; j0 += CB) {
        const int nb = (int)std::min<int64_t>(CB, n - j0);
        if (j0 > 0) {
            if (!trans_lower) gemm<T>(h, 0, 0, rows, nb, j0, T(-1), B, ldb, Tri + j0 * ldt, ldt, T(1), B + j0 * ldb, ldb);
            else gemm<T>(h, 0, 1, rows, nb, j0, T(-1), B, ldb, Tri + j0, ldt, T(1), B + j0 * ldb, ldb);   // (L[j0.., ..j0])^T
        }
        // equation j: sum_{i <= j} x_i U[i, j] = b_j  ->  M(j, i) = U[i, j]; forward substitution
        const int64_t sj = trans_lower ? 1 : ldt, si = trans_lower ? ldt : 1;
        trsv_block<T>(h, true, Tri + j0 + j0 * ldt, sj, si, nb, (const T *)nullptr, B + j0 * ldb, /*sv=*/1, /*sb=*/ldb, rows, info);
    }
}

template <typename T>
void trsm_right_upper(lfb_handle &h, int64_t rows, int64_t n, const T *U, int64_t ldu, T *B, int64_t ldb) {
    trsm_right<T>(h, rows, n, U, ldu, 0, B, ldb, nullptr);
}

#define INST(T)                                                                                       \
    template void trsm_left<T>(lfb_handle &, int, int, int64_t, int64_t, const T *, int64_t, const T *, T *, int64_t); \
    template void trsm_right_upper<T>(lfb_handle &, int64_t, int64_t, const T *, int64_t, T *, int64_t); \
    template void trsm_right<T>(lfb_handle &, int64_t, int64_t, const T *, int64_t, int, T *, int64_t, const int64_t *);
INST(float)
INST(double)
#undef INST

}  // namespace lfb
