// The dense blocks of LOBPCG, chained on the device.
//
// lobpcg/algorithm.rs is the biggest in-crate consumer of the Cholesky / triangular-solve path: every iteration calls
// `orthonormalize` (:81-97: Gram matrix, cholesky_into, solve_triangular_into on the transposed block) on tall-thin
// blocks and `apply_constraints` (:63-76: Gram, one triangular solve, one GEMM).  Through the per-factorisation entry
// points each of these would cross PCIe four times; here each is ONE call whose intermediates (the k x k Gram matrix,
// its factor) never leave HBM, built from kernels that are already parity-tested on their own: the tensor-core GEMM
// (split-K for the k x k x n Gram products), cholesky_lower and the blocked right-/left-hand triangular solves.
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

#define INST(T)                                                                                                  \
    template void orthonormalize<T>(lfb_handle &, T *, int64_t, int64_t, int64_t, T *, int64_t, int64_t *);      \
    template void apply_constraints<T>(lfb_handle &, T *, int64_t, int64_t, int64_t, const T *, int64_t, int64_t, const T *, int64_t);
INST(float)
INST(double)
#undef INST

}  // namespace lfb
