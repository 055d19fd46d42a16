// The dense blocks of LOBPCG, chained on the device.
//
// lobpcg/algorithm.rs is the biggest in-crate consumer of the Cholesky / triangular-solve path: every iteration calls
// `orthonormalize` (:81-97: Gram matrix, cholesky_into, solve_triangular_into on the transposed block) on tall-thin
// blocks and `apply_constraints` (:63-76: Gram, one triangular solve, one GEMM).  Through the per-factorisation entry
// points each of these would cross PCIe four times; here each is ONE call whose intermediates (the k x k Gram matrix,
// its factor) never leave HBM, built from kernels that are already parity-tested on their own: the tensor-core GEMM
// (split-K for the k x k x n Gram products), cholesky_lower and the blocked right-/left-hand triangular solves.
#include <algorithm>

#include "common.cuh"

namespace lfb {

// lobpcg/algorithm.rs:81-97 orthonormalize(v) -> (u, gram_vv_fac):
//   gram = v^T v (:82);  L = gram.cholesky_into() (:83, lower, strict upper zeroed -- cholesky.rs:78-82);
//   u = (L^-1 v^T)^T = v L^-T (:91-94).
// V (rows x cols) is overwritten with u, Lm (cols x cols) receives L; *d_info = 0, or the failing pivot row + 1
// (NotPositiveDefinite, cholesky.rs:69-71) -- V is then unspecified (the reference has consumed `v` by value, :81).
template <typename T>
void orthonormalize(lfb_handle &h, T *V, int64_t rows, int64_t cols, int64_t ld, T *Lm, int64_t ldl, int64_t *d_info) {
    if (cols <= 0) {
        LFB_CUDA(cudaMemsetAsync(d_info, 0, sizeof(int64_t), h.stream));
        return;
    }
    if (rows > 0) gemm<T>(h, 1, 0, cols, cols, rows, T(1), V, ld, V, ld, T(0), Lm, ldl);
    else fill<T>(h, Lm, cols, cols, ldl, T(0), T(0));
    cholesky_lower<T>(h, Lm, cols, ldl, /*clean=*/1, d_info);
    trsm_right<T>(h, rows, cols, Lm, ldl, /*trans_lower=*/1, V, ld, d_info);
}

// lobpcg/algorithm.rs:63-76 apply_constraints(v, cholesky_yy, y):
//   gram_yv = y^T v (:68);  u = cholesky_yy.solve_triangular_into(gram_yv, Lower) (:70-72);  v -= y u (:75).
// V: n x k in place, Y: n x m, Lyy: m x m (lower triangle read).
template <typename T>
void apply_constraints(lfb_handle &h, T *V, int64_t n, int64_t k, int64_t ldv, const T *Lyy, int64_t m, int64_t ldl, const T *Y,
                       int64_t ldy) {
    if (n <= 0 || k <= 0 || m <= 0) return;
    const int64_t ldg = round_up(m, 2);
    DevBuf<T> G(h, (size_t)ldg * k);
    gemm<T>(h, 1, 0, m, k, n, T(1), Y, ldy, V, ldv, T(0), G, ldg);
    trsm_left<T>(h, /*lower=*/1, /*trans=*/0, m, k, Lyy, ldl, (const T *)nullptr, G, ldg);
    gemm<T>(h, 0, 0, n, k, m, T(-1), Y, ldy, G, ldg, T(1), V, ldv);
}

namespace {

// column j of `out` (k x size) = sign * column perm[j] of `src`, sign = -1 if the first entry of that column has its sign
// BIT set (Rust's signum, algorithm.rs:40-41) and fix_sign != 0.
template <typename T>
__global__ void gather_sign_cols_kernel(const T *__restrict__ src, int64_t lds, int k, int size, const int *__restrict__ perm, int fix_sign,
                                        T *__restrict__ out, int64_t ldo) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= k) return;
    for (int j = blockIdx.y; j < size; j += gridDim.y) {
        const T *col = src + (int64_t)perm[j] * lds;
        const T sg = (fix_sign && signbit(col[0])) ? T(-1) : T(1);
        out[i + (int64_t)j * ldo] = sg * col[i];
    }
}

template <typename T>
__global__ void scale_cols_by_kernel(T *__restrict__ a, int64_t ld, int k, const T *__restrict__ s) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= k) return;
    for (int j = blockIdx.y; j < k; j += gridDim.y) a[i + (int64_t)j * ld] *= s[j];
}

}  // namespace

// lobpcg/algorithm.rs:16-44 generalized_eig / sorted_eig on device-resident k x k operands (both CONSUMED, as the
// reference moves them): eigenvalues land in host memory, eigenvectors stay on the device.
//   dB != nullptr (:16-25): (vals_b, vecs_b) = eigh(b);  vecs_b~ = vecs_b diag(1 / sqrt(max(vals_b, 1e-10f32)));
//                           a~ = vecs_b~^T (a vecs_b~);  (vals, vecs_a) = eigh(a~);  vecs = vecs_b~ vecs_a
//   order 0: the pair as eigh returns it (generalized_eig / eigh_into);  1 / 2: sort_eig(Largest / Smallest) (eigh.rs:275-325,
//   a stable sort), columns multiplied by the sign of their first entry (:40-41), truncated to `size` (:43).
// Returns false if an eigenvalue is NaN (the reference's sort panics, eigh.rs:326-328).
template <typename T>
bool sorted_eig_dev(lfb_handle &h, T *dA, int64_t lda, T *dB, int64_t ldb, int64_t k, int64_t size, int order, T *vals_host, T *dVecs,
                    int64_t ldv) {
    if (k <= 0) return true;
    const int64_t ldq = round_up(k, 2);
    DevBuf<T> Q(h, (size_t)ldq * k), Tmp(h, (size_t)ldq * k);
    std::vector<T> vals((size_t)k);
    const T *src = Q.get();
    if (dB) {
        DevBuf<T> Qb(h, (size_t)ldq * k), dRecip(h, (size_t)k);
        std::vector<T> vb((size_t)k);
        symmetric_eig<T>(h, dB, k, ldb, vb.data(), Qb.get(), ldq);                                          // :17
        const T floor_ = (T)1e-10f;                                                                         // A::from(1e-10f32)
        for (auto &x : vb) x = T(1) / std::sqrt(std::max(x, floor_));                                       // :18
        LFB_CUDA(cudaMemcpyAsync(dRecip.get(), vb.data(), sizeof(T) * k, cudaMemcpyHostToDevice, h.stream));
        dim3 g((unsigned)cdiv(k, 128), (unsigned)(k < 65535 ? k : 65535));
        scale_cols_by_kernel<T><<<g, 128, 0, h.stream>>>(Qb.get(), ldq, (int)k, dRecip.get());              // :19 vecs_b * recip
        LFB_LAUNCH_CHECK(h);
        gemm<T>(h, 0, 0, k, k, k, T(1), dA, lda, Qb.get(), ldq, T(0), Tmp.get(), ldq);                      // a vecs_b~
        gemm<T>(h, 1, 0, k, k, k, T(1), Qb.get(), ldq, Tmp.get(), ldq, T(0), dA, lda);                      // :20 a~
        LFB_CUDA(cudaStreamSynchronize(h.stream));                                                          // vb is read by the copy above
        symmetric_eig<T>(h, dA, k, lda, vals.data(), Q.get(), ldq);                                         // :21
        gemm<T>(h, 0, 0, k, k, k, T(1), Qb.get(), ldq, Q.get(), ldq, T(0), Tmp.get(), ldq);                 // :22 vecs_b~ vecs_a
        src = Tmp.get();
        LFB_CUDA(cudaStreamSynchronize(h.stream));                                                          // Qb / dRecip are released here
    } else {
        symmetric_eig<T>(h, dA, k, lda, vals.data(), Q.get(), ldq);
    }
    std::vector<int> perm((size_t)k);
    for (int64_t i = 0; i < k; ++i) perm[i] = (int)i;
    if (order != 0) {
        for (auto x : vals) if (x != x) return false;
        if (order == 1) std::stable_sort(perm.begin(), perm.end(), [&](int a, int b) { return vals[a] > vals[b]; });
        else std::stable_sort(perm.begin(), perm.end(), [&](int a, int b) { return vals[a] < vals[b]; });
    }
    const int64_t nout = order == 0 ? k : std::min(size, k);
    for (int64_t j = 0; j < nout; ++j) vals_host[j] = vals[perm[j]];
    DevBuf<int> dPerm(h, (size_t)k);
    LFB_CUDA(cudaMemcpyAsync(dPerm.get(), perm.data(), sizeof(int) * k, cudaMemcpyHostToDevice, h.stream));
    dim3 g((unsigned)cdiv(k, 128), (unsigned)(nout < 65535 ? std::max<int64_t>(nout, 1) : 65535));
    gather_sign_cols_kernel<T><<<g, 128, 0, h.stream>>>(src, ldq, (int)k, (int)nout, dPerm.get(), order != 0, dVecs, ldv);
    LFB_LAUNCH_CHECK(h);
    LFB_CUDA(cudaStreamSynchronize(h.stream));                                                              // perm (host) and the scratch are released
    return true;
}

#define INST(T)                                                                                                  \
    template bool sorted_eig_dev<T>(lfb_handle &, T *, int64_t, T *, int64_t, int64_t, int64_t, int, T *, T *, int64_t); \
    template void orthonormalize<T>(lfb_handle &, T *, int64_t, int64_t, int64_t, T *, int64_t, int64_t *);      \
    template void apply_constraints<T>(lfb_handle &, T *, int64_t, int64_t, int64_t, const T *, int64_t, int64_t, const T *, int64_t);
INST(float)
INST(double)
#undef INST

}  // namespace lfb
