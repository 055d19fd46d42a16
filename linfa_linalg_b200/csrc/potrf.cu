// Blocked Cholesky (lower, column-major).
//
// Replaces src/cholesky.rs:51-83 (cholesky_inplace_dirty / cholesky_inplace).  The reference's
// row-by-row triple loop is restructured right-looking, two levels:
//   outer panels of `chol_nb` columns:  A22 -= P P^T (lower only) is ONE large-K DMMA GEMM per panel
//   inner 64-column blocks:             [potf2 + inverse] -> [rows below *= L^-T] -> [K=64 GEMM on the
//                                       rest of the panel]
// Latency notes (measured, profiles/): a 64x64 diagonal block is a chain of 64 dependent pivots;
// with one warp per scheduler every instruction costs ~5-8 cycles, and DP sqrt/div are ~50
// instructions each.  So (1) the diagonal kernel spreads each row's dot product over 4 lanes and
// takes 1/sqrt(d) from an f32 seed + 2 Newton steps, (2) the triangular solve below the diagonal
// block is NOT a substitution (a 2000-instruction dependent chain per row) but a multiplication by
// the explicitly inverted 64x64 block (recursive doubling, 6 levels), which is throughput-bound.
// Only the lower triangle is read or written (the `dirty` contract, cholesky.rs:17-19).
#include "common.cuh"

namespace lfb {
namespace {

constexpr int CB = 64;   // diagonal block
constexpr int SP = 68;   // shared pitch: 68 % 16 == 4 keeps (row, lane-part) accesses conflict free

template <typename T> __device__ __forceinline__ T fast_rsqrt(T d);
template <> __device__ __forceinline__ float fast_rsqrt<float>(float d) { return rsqrtf(d); }
template <> __device__ __forceinline__ double fast_rsqrt<double>(double d) {
    if (d > 1e-30 && d < 1e30) {
        double y = (double)rsqrtf((float)d);          // ~2^-22 relative error
#pragma unroll
        for (int it = 0; it < 2; ++it) {              // Newton: e -> 1.5 e^2
            const double r = fma(-d * y, y, 1.0);
            y = fma(0.5 * y, r, y);
        }
        return y;
    }
    return rsqrt(d);
}

// In:  A (n x n lower block at (row0,row0) of the big matrix).  Out: L in place, Linv (64 x 64,
// column-major ld 64, identity padded beyond n) = L^-1.  256 threads: thread t -> row t/4, part t%4.
template <typename T>
__global__ void __launch_bounds__(256) potf2_inv_kernel(T *A, int64_t ld, int n, int64_t row0, int64_t *info, T *Linv) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *s = reinterpret_cast<T *>(smem_raw);   // L   [CB][SP]
    T *x = s + CB * SP;                       // L^-1 [CB][SP]
    T *tmp = x + CB * SP;                     // scratch [32][SP] for the doubling products
    if (*info != 0) return;
    const int tid = threadIdx.x, r = tid >> 2, q = tid & 3;
    for (int e = tid; e < CB * CB; e += 256) {
        const int i = e % CB, k = e / CB;
        T v = (i == k) ? T(1) : T(0);
        if (i < n && k <= i) v = A[i + (int64_t)k * ld];
        s[i * SP + k] = v;
    }
    __syncthreads();
    int fail = 0;
    for (int j = 0; j < n; ++j) {
        T dp = T(0), vp = T(0);
        const T *rj = s + j * SP, *rr = s + r * SP;
        for (int k = q; k < j; k += 4) {
            const T l = rj[k];
            dp += l * l;
            vp += rr[k] * l;
        }
        dp += __shfl_xor_sync(0xffffffffu, dp, 1);
        vp += __shfl_xor_sync(0xffffffffu, vp, 1);
        dp += __shfl_xor_sync(0xffffffffu, dp, 2);
        vp += __shfl_xor_sync(0xffffffffu, vp, 2);
        const T d = rj[j] - dp;                        // identical in every thread
        if (d <= T(0)) {                               // cholesky.rs:69-71 (false for NaN)
            fail = j + 1;
            break;
        }
        const T inv = fast_rsqrt(d);
        T mine = T(0);
        if (q == 0 && r >= j) {
            if (r == j) {
                T sq = d * inv;                        // sqrt(d) with one correction step
                sq = fma(T(0.5) * inv, fma(-sq, sq, d), sq);
                mine = sq;
            } else {
                mine = (rr[j] - vp) * inv;
            }
        }
        __syncthreads();                               // everyone has read row j / column j inputs
        if (q == 0 && r >= j && r < n) s[r * SP + j] = mine;
        __syncthreads();
    }
    if (tid == 0 && fail) *info = row0 + fail;
    __syncthreads();
    for (int e = tid; e < n * n; e += 256) {
        const int i = e % n, k = e / n;
        if (k <= i) A[i + (int64_t)k * ld] = s[i * SP + k];
    }
    if (fail) return;

    // ---- X = L^-1 by recursive doubling: inv([A 0; B C]) = [A^-1 0; -C^-1 B A^-1, C^-1] ----
    for (int e = tid; e < CB * CB; e += 256) {
        const int i = e / CB, k = e % CB;
        x[i * SP + k] = (i == k) ? T(1) / s[i * SP + i] : T(0);
    }
    __syncthreads();
    for (int b = 1; b < CB; b <<= 1) {
        const int npair = CB / (2 * b), per = b * b;
        // tmp = B * A^-1   (B = L[o+b.., o..o+b), A^-1 = x[o.., o..) lower)
        for (int e = tid; e < npair * per; e += 256) {
            const int p = e / per, i = (e % per) / b, jj = e % b, o = p * 2 * b;
            T acc = T(0);
            for (int k = jj; k < b; ++k) acc += s[(o + b + i) * SP + o + k] * x[(o + k) * SP + o + jj];
            tmp[(p * b + i) * SP + jj] = acc;
        }
        __syncthreads();
        // X21 = -C^-1 * tmp   (C^-1 = x[o+b.., o+b..) lower)
        for (int e = tid; e < npair * per; e += 256) {
            const int p = e / per, i = (e % per) / b, jj = e % b, o = p * 2 * b;
            T acc = T(0);
            for (int k = 0; k <= i; ++k) acc += x[(o + b + i) * SP + o + b + k] * tmp[(p * b + k) * SP + jj];
            x[(o + b + i) * SP + o + jj] = -acc;
        }
        __syncthreads();
    }
    for (int e = tid; e < CB * CB; e += 256) {
        const int i = e % CB, k = e / CB;
        Linv[i + k * CB] = x[i * SP + k];
    }
}

// B (rows x 64 columns, column-major ldb) <- B * Linv^T in place; Linv is 64 x 64 lower (ld 64).
// One CTA per 128 rows: the tile and Linv are staged in shared memory, each thread owns a
// 4 (rows) x 8 (cols) register block.
template <typename T>
__global__ void __launch_bounds__(256) trsm_mult_kernel(T *B, int64_t ldb, int64_t rows, int nb, const T *__restrict__ Linv,
                                                        const int64_t *info) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *sB = reinterpret_cast<T *>(smem_raw);  // [64 k][132]  sB[k*132 + row]
    T *sL = sB + CB * 132;                    // [64 k][68]   sL[k*68 + j] = Linv[j][k]  (j >= k)
    if (*info != 0) return;
    const int tid = threadIdx.x;
    const int64_t r0 = (int64_t)blockIdx.x * 128;
    const int nr = (int)((rows - r0) < 128 ? (rows - r0) : 128);
#pragma unroll 4
    for (int e = tid; e < CB * CB; e += 256) {
        const int j = e % CB, k = e / CB;   // Linv[j + k*64]
        sL[k * 68 + j] = Linv[j + k * CB];
    }
#pragma unroll 8
    for (int e = tid; e < CB * 128; e += 256) {
        const int rr = e % 128, k = e / 128;
        sB[k * 132 + rr] = (rr < nr && k < nb) ? B[r0 + rr + (int64_t)k * ldb] : T(0);
    }
    __syncthreads();
    const int tr = (tid & 31) * 4, tc = (tid >> 5) * 8;   // rows tr..tr+3, cols tc..tc+7
    T acc[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = T(0);
    // out[row][j] = sum_{k <= j} B[row][k] * Linv[j][k]
    const int kmax = tc + 8;   // columns j < tc+8 need k <= j < tc+8
#pragma unroll 4
    for (int k = 0; k < kmax; ++k) {
        T bv[4], lv[8];
#pragma unroll
        for (int i = 0; i < 4; ++i) bv[i] = sB[k * 132 + tr + i];
#pragma unroll
        for (int j = 0; j < 8; ++j) lv[j] = sL[k * 68 + tc + j];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = fma(bv[i], lv[j], acc[i][j]);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (tr + i < nr && tc + j < nb) B[r0 + tr + i + (int64_t)(tc + j) * ldb] = acc[i][j];
}

}  // namespace

template <typename T>
void cholesky_lower(lfb_handle &h, T *A, int64_t n, int64_t ld, int clean, int64_t *d_info) {
    LFB_CUDA(cudaMemsetAsync(d_info, 0, sizeof(int64_t), h.stream));
    if (n <= 0) return;
    const size_t smem_p = sizeof(T) * (2 * CB * SP + 32 * SP);
    const size_t smem_t = sizeof(T) * (CB * 132 + CB * 68);
    static DeviceOnce cfg;   // function attributes are per device
    cfg.run(h.device, [&] {
        LFB_CUDA(cudaFuncSetAttribute(potf2_inv_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_p));
        LFB_CUDA(cudaFuncSetAttribute(trsm_mult_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_t));
    });
    DevBuf<T> Linv(h, CB * CB);
    const int64_t NB = std::max<int64_t>(CB, round_up(h.opt.chol_nb, CB));
    // panel [k0, k0+nb): 64-column blocks, everything on the handle's current stream
    auto factor_panel = [&](int64_t k0, int64_t nb) {
        const int64_t pend = k0 + nb;
        for (int64_t j0 = k0; j0 < pend; j0 += CB) {
            const int jb = (int)std::min<int64_t>(CB, pend - j0);
            T *Ajj = A + j0 + j0 * ld;
            potf2_inv_kernel<T><<<1, 256, smem_p, h.stream>>>(Ajj, ld, jb, j0, d_info, Linv.get());
            LFB_LAUNCH_CHECK(h);
            const int64_t below = n - (j0 + jb);
            if (below <= 0) continue;
            T *Bp = Ajj + jb;   // rows j0+jb.., columns j0..j0+jb
            trsm_mult_kernel<T><<<(unsigned)cdiv(below, 128), 256, smem_t, h.stream>>>(Bp, ld, below, jb, Linv.get(), d_info);
            LFB_LAUNCH_CHECK(h);
            const int64_t rem = pend - (j0 + jb);   // remaining columns of this panel
            if (rem > 0)   // A[j0+jb.., j0+jb..pend) -= Bp * Bp[0:rem, :]^T   (lower part only)
                gemm<T>(h, 0, 1, below, rem, jb, T(-1), Bp, ld, Bp, ld, T(1), A + (j0 + jb) + (j0 + jb) * ld, ld, /*lower_only=*/1);
        }
    };
    // Look-ahead: the trailing SYRK of panel k is split into the block column of panel k+1 (done
    // first) and the rest; panel k+1 is factored on a high-priority side stream while the rest of
    // the SYRK runs, so the latency-bound diagonal kernels leave the critical path.
    cudaStream_t sm = h.stream, sp = h.aux_stream;
    const bool la = h.opt.lookahead && sp != nullptr && !h.prof_on && n >= 4 * NB;
    factor_panel(0, std::min<int64_t>(NB, n));
    if (h.chol_panel_hook) h.chol_panel_hook(0, std::min<int64_t>(NB, n));
    for (int64_t k0 = 0; k0 < n; k0 += NB) {
        const int64_t nb = std::min<int64_t>(NB, n - k0);
        const int64_t pend = k0 + nb;
        const int64_t rows = n - pend;
        if (rows <= 0) break;
        const int64_t nbn = std::min<int64_t>(NB, rows);
        T *P = A + pend + k0 * ld;
        if (la) {
            gemm<T>(h, 0, 1, rows, nbn, nb, T(-1), P, ld, P, ld, T(1), A + pend + pend * ld, ld, /*lower_only=*/1);
            LFB_CUDA(cudaEventRecord(h.ev[0], sm));
            LFB_CUDA(cudaStreamWaitEvent(sp, h.ev[0], 0));
            h.stream = sp;
            try {
                factor_panel(pend, nbn);
                if (h.chol_panel_hook) h.chol_panel_hook(pend, nbn);     // h.stream is the side stream here
            } catch (...) {
                h.stream = sm;
                throw;
            }
            LFB_CUDA(cudaEventRecord(h.ev[1], sp));
            h.stream = sm;
            const int64_t rows2 = rows - nbn;
            if (rows2 > 0) {
                T *P2 = P + nbn;
                gemm<T>(h, 0, 1, rows2, rows2, nb, T(-1), P2, ld, P2, ld, T(1), A + (pend + nbn) + (pend + nbn) * ld, ld, /*lower_only=*/1);
            }
            LFB_CUDA(cudaStreamWaitEvent(sm, h.ev[1], 0));
        } else {
            gemm<T>(h, 0, 1, rows, rows, nb, T(-1), P, ld, P, ld, T(1), A + pend + pend * ld, ld, /*lower_only=*/1);
            factor_panel(pend, nbn);
            if (h.chol_panel_hook) h.chol_panel_hook(pend, nbn);
        }
    }
    if (clean) triangular_zero<T>(h, A, n, ld, /*keep_lower=*/1);
}

template void cholesky_lower<float>(lfb_handle &, float *, int64_t, int64_t, int, int64_t *);
template void cholesky_lower<double>(lfb_handle &, double *, int64_t, int64_t, int, int64_t *);

}  // namespace lfb
