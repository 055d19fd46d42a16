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
#include <array>
#include <memory>

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
__global__ void __launch_bounds__(256) potf2_inv_kernel(T *A, int64_t ld, int n, int64_t row0, int64_t *info, T *Linv, int do_inv = 1, int ldinv = CB) {
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
    if (fail || !do_inv) return;

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
        Linv[i + k * ldinv] = x[i * SP + k];
    }
}

// Second generation of the diagonal-block kernel: RIGHT-LOOKING with the block in registers.
// The 64 x 64 lower triangle is 136 blocks of 4 x 4; thread p < 136 owns block (bi >= bj) in 16 registers.  Step j:
// every thread reads the (unscaled) column j that its owners published in shared memory one step earlier, takes
// inv = 1/sqrt(a_jj) itself (no broadcast of the pivot: it is colbuf[j]), scales the 4 + 4 entries it needs and applies
// the rank-1 update to its columns > j -- at most 16 independent FMAs -- the owners of column j store its final values
// into the shared copy of L, and the owners of column j+1 publish it.  ONE barrier per step (two column buffers).  The j
// loop is unrolled by 4 (the column inside its block is a compile-time constant), so every register index is static and
// a step is ~60 instructions: the first version kept the block in local memory / behind 64 selects (~260 instructions per
// step, 39 us per block, profiles/r2_potf2.md).  Same outputs as potf2_inv_kernel: L in place, L^-1 in Linv,
// *info = first non-positive pivot (cholesky.rs:69-71).
template <typename T, int JJ>
__device__ __forceinline__ bool potf2_rl_step(T (&a)[4][4], int j, bool active, int bi, int bj, int r0, int c0, T *colbuf, T *s) {
    const T *cur = colbuf + (j & 1) * CB;
    T *nxt = colbuf + ((j + 1) & 1) * CB;
    const T d = cur[j];                                        // identical in every thread
    if (d <= T(0)) return false;                               // cholesky.rs:69-71 (false for NaN)
    const int cj = j >> 2;
    if (active && bj >= cj) {
        const T inv = fast_rsqrt(d);
        T lr[4], lc[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            lr[k] = cur[r0 + k] * inv;
            lc[k] = cur[c0 + k] * inv;
        }
        if (bj > cj) {                                         // the whole block lies right of column j
#pragma unroll
            for (int cc = 0; cc < 4; ++cc)
#pragma unroll
                for (int rr = 0; rr < 4; ++rr) a[rr][cc] = fma(-lr[rr], lc[cc], a[rr][cc]);
            if (JJ == 3 && bj == cj + 1) {                     // publish the updated, unscaled column j+1 (first of the next block column)
#pragma unroll
                for (int rr = 0; rr < 4; ++rr) nxt[r0 + rr] = a[rr][0];
            }
        } else {                                               // bj == cj: column j is column JJ of this block
#pragma unroll
            for (int cc = JJ + 1; cc < 4; ++cc)
#pragma unroll
                for (int rr = 0; rr < 4; ++rr) a[rr][cc] = fma(-lr[rr], lc[cc], a[rr][cc]);
            T sq = d * inv;                                    // sqrt(d) with one correction step
            sq = fma(T(0.5) * inv, fma(-sq, sq, d), sq);
#pragma unroll
            for (int rr = 0; rr < 4; ++rr) {
                const int r = r0 + rr;
                if (r >= j) s[r * SP + j] = r == j ? sq : lr[rr];
            }
            if (JJ < 3) {
#pragma unroll
                for (int rr = 0; rr < 4; ++rr) nxt[r0 + rr] = a[rr][JJ < 3 ? JJ + 1 : 0];
            }
        }
    }
    __syncthreads();
    return true;
}

template <typename T>
__global__ void __launch_bounds__(256) potf2_rl_inv_kernel(T *A, int64_t ld, int n, int64_t row0, int64_t *info, T *Linv, int do_inv = 1, int ldinv = CB) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *s = reinterpret_cast<T *>(smem_raw);   // L   [CB][SP]
    T *x = s + CB * SP;                       // L^-1 [CB][SP]
    T *tmp = x + CB * SP;                     // scratch [32][SP]; its head doubles as the two column buffers of the factor phase
    T *colbuf = tmp;                          // [2][CB]
    if (*info != 0) return;
    const int tid = threadIdx.x;
    const bool active = tid < 136;
    int bi = 0;
    while ((bi + 1) * (bi + 2) / 2 <= tid) ++bi;
    const int bj = tid - bi * (bi + 1) / 2;
    const int r0 = 4 * bi, c0 = 4 * bj;
    T a[4][4];
#pragma unroll
    for (int cc = 0; cc < 4; ++cc)
#pragma unroll
        for (int rr = 0; rr < 4; ++rr) {
            const int r = r0 + rr, c = c0 + cc;
            T v = (r == c) ? T(1) : T(0);                       // identity padding beyond n, zeros above the diagonal
            if (active && r < n && c <= r) v = A[r + (int64_t)c * ld];
            a[rr][cc] = v;
        }
    for (int e = tid; e < CB * CB; e += 256) {                  // shared L starts as the identity; finished columns overwrite it
        const int i = e / CB, k = e % CB;
        s[i * SP + k] = (i == k) ? T(1) : T(0);
    }
    if (active && bj == 0) {
#pragma unroll
        for (int rr = 0; rr < 4; ++rr) colbuf[r0 + rr] = a[rr][0];
    }
    __syncthreads();
    int fail = 0;
    for (int j = 0; j < n; j += 4) {
        if (!potf2_rl_step<T, 0>(a, j, active, bi, bj, r0, c0, colbuf, s)) { fail = j + 1; break; }
        if (j + 1 >= n) break;
        if (!potf2_rl_step<T, 1>(a, j + 1, active, bi, bj, r0, c0, colbuf, s)) { fail = j + 2; break; }
        if (j + 2 >= n) break;
        if (!potf2_rl_step<T, 2>(a, j + 2, active, bi, bj, r0, c0, colbuf, s)) { fail = j + 3; break; }
        if (j + 3 >= n) break;
        if (!potf2_rl_step<T, 3>(a, j + 3, active, bi, bj, r0, c0, colbuf, s)) { fail = j + 4; break; }
    }
    if (tid == 0 && fail) *info = row0 + fail;
    __syncthreads();
    for (int e = tid; e < n * n; e += 256) {
        const int i = e % n, k = e / n;
        if (k <= i) A[i + (int64_t)k * ld] = s[i * SP + k];
    }
    if (fail || !do_inv) return;

    // ---- X = L^-1 by recursive doubling, as in potf2_inv_kernel ----
    for (int e = tid; e < CB * CB; e += 256) {
        const int i = e / CB, k = e % CB;
        x[i * SP + k] = (i == k) ? T(1) / s[i * SP + i] : T(0);
    }
    __syncthreads();
    for (int b = 1; b < CB; b <<= 1) {
        const int npair = CB / (2 * b), per = b * b;
        for (int e = tid; e < npair * per; e += 256) {
            const int p = e / per, i = (e % per) / b, jj = e % b, o = p * 2 * b;
            T acc = T(0);
            for (int k = jj; k < b; ++k) acc += s[(o + b + i) * SP + o + k] * x[(o + k) * SP + o + jj];
            tmp[(p * b + i) * SP + jj] = acc;
        }
        __syncthreads();
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
        Linv[i + k * ldinv] = x[i * SP + k];
    }
}

// B (rows x 64 columns, column-major ldb) <- B * Linv^T in place; Linv is 64 x 64 lower (ld 64).
// One CTA per 128 rows: the tile and Linv are staged in shared memory, each thread owns a
// 4 (rows) x 8 (cols) register block.
template <typename T>
__global__ void __launch_bounds__(256) trsm_mult_kernel(T *B, int64_t ldb, int64_t rows, int nb, const T *__restrict__ Linv,
                                                        const int64_t *info, int ldinv = CB) {
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
        sL[k * 68 + j] = Linv[j + k * ldinv];
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

// ncols < n: factor only the first ncols columns (full height n) -- the trapezoid of one "arrival wave" of the host path
// (api.cu: cholesky_host), whose trailing columns are brought up to date later by one large-K product.  col_base: global index of
// column 0 (failure index, panel hook); reset_info = 0 keeps an earlier wave's failure.
template <typename T>
static void cholesky_lower_v1(lfb_handle &h, T *A, int64_t n, int64_t ld, int clean, int64_t *d_info, int64_t ncols = -1,
                              int64_t col_base = 0, int reset_info = 1) {
    if (reset_info) LFB_CUDA(cudaMemsetAsync(d_info, 0, sizeof(int64_t), h.stream));
    if (n <= 0) return;
    if (ncols < 0 || ncols > n) ncols = n;
    const size_t smem_p = sizeof(T) * (2 * CB * SP + 32 * SP);
    const size_t smem_t = sizeof(T) * (CB * 132 + CB * 68);
    static DeviceOnce cfg;   // function attributes are per device
    cfg.run(h.device, [&] {
        LFB_CUDA(cudaFuncSetAttribute(potf2_inv_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_p));
        LFB_CUDA(cudaFuncSetAttribute(potf2_rl_inv_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_p));
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
            if (h.opt.chol_potf2_rl) potf2_rl_inv_kernel<T><<<1, 256, smem_p, h.stream>>>(Ajj, ld, jb, j0 + col_base, d_info, Linv.get());
            else potf2_inv_kernel<T><<<1, 256, smem_p, h.stream>>>(Ajj, ld, jb, j0 + col_base, d_info, Linv.get());
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
    // The same panel on TWO streams (look-ahead mode, option chol_split_panel).  The event trace of the single-stream panel
    // (profiles/r2_chol_analysis.md) shows the factorisation GEMM bound for the first third and bound by the side-stream
    // panel afterwards: 0.8-1.3 ms per panel, of which the 64-pivot diagonal kernels are 0.32 -- the rest is the full-height
    // TRSM / K = 64 GEMM of every block sitting between two diagonal kernels.  Here the chain that the next diagonal block
    // really depends on -- potf2(j), the TRSM and update INSIDE the nb x nb diagonal block (<= 4 and <= 16 CTAs) -- runs on
    // the handle's stream, and the rows below the diagonal block follow on a second stream: TRSM(j) after potf2(j),
    // update(j) after the diagonal block's TRSM(j) (its B operand).  Same kernels, same arithmetic per entry.
    DevBuf<T> LinvAll(h, (size_t)CB * CB * (NB / CB));
    std::vector<cudaEvent_t> evP, evT;
    cudaEvent_t evJoin = nullptr;
    const bool split_ok = h.opt.chol_split_panel && h.aux2_stream != nullptr;
    if (split_ok) {
        evP.resize(NB / CB); evT.resize(NB / CB);
        for (auto &e : evP) LFB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        for (auto &e : evT) LFB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        LFB_CUDA(cudaEventCreateWithFlags(&evJoin, cudaEventDisableTiming));
    }
    struct EvGuard {
        std::vector<cudaEvent_t> &a, &b; cudaEvent_t &j;
        ~EvGuard() { for (auto e : a) cudaEventDestroy(e); for (auto e : b) cudaEventDestroy(e); if (j) cudaEventDestroy(j); }
    } ev_guard{evP, evT, evJoin};
    auto factor_panel_split = [&](int64_t k0, int64_t nb) {
        const int64_t pend = k0 + nb, tall = n - pend;          // rows below the diagonal block
        cudaStream_t s1 = h.stream, s2 = h.aux2_stream;
        LFB_CUDA(cudaEventRecord(evJoin, s1));
        LFB_CUDA(cudaStreamWaitEvent(s2, evJoin, 0));           // the second stream starts where the first one is
        int jb_i = 0;
        for (int64_t j0 = k0; j0 < pend; j0 += CB, ++jb_i) {
            const int jb = (int)std::min<int64_t>(CB, pend - j0);
            T *Ajj = A + j0 + j0 * ld;
            T *Lj = LinvAll.get() + (size_t)jb_i * CB * CB;
            if (h.opt.chol_potf2_rl) potf2_rl_inv_kernel<T><<<1, 256, smem_p, s1>>>(Ajj, ld, jb, j0 + col_base, d_info, Lj);
            else potf2_inv_kernel<T><<<1, 256, smem_p, s1>>>(Ajj, ld, jb, j0 + col_base, d_info, Lj);
            LFB_LAUNCH_CHECK(h);
            LFB_CUDA(cudaEventRecord(evP[jb_i], s1));
            const int64_t inner = pend - (j0 + jb);             // rows of the diagonal block below this 64-block
            T *Bp = Ajj + jb;                                   // rows j0+jb.., columns j0..j0+jb
            if (inner > 0) {
                trsm_mult_kernel<T><<<(unsigned)cdiv(inner, 128), 256, smem_t, s1>>>(Bp, ld, inner, jb, Lj, d_info);
                LFB_LAUNCH_CHECK(h);
                LFB_CUDA(cudaEventRecord(evT[jb_i], s1));
                gemm<T>(h, 0, 1, inner, inner, jb, T(-1), Bp, ld, Bp, ld, T(1), A + (j0 + jb) + (j0 + jb) * ld, ld, /*lower_only=*/1);
            }
            if (tall > 0) {
                T *Bt = A + pend + j0 * ld;                     // rows pend.., columns j0..j0+jb
                LFB_CUDA(cudaStreamWaitEvent(s2, evP[jb_i], 0));
                trsm_mult_kernel<T><<<(unsigned)cdiv(tall, 128), 256, smem_t, s2>>>(Bt, ld, tall, jb, Lj, d_info);
                LFB_LAUNCH_CHECK(h);
                if (inner > 0) {                                // A[pend.., j0+jb..pend) -= Bt * Bp^T
                    LFB_CUDA(cudaStreamWaitEvent(s2, evT[jb_i], 0));
                    h.stream = s2;
                    try {
                        gemm<T>(h, 0, 1, tall, inner, jb, T(-1), Bt, ld, Bp, ld, T(1), A + pend + (j0 + jb) * ld, ld, 0);
                    } catch (...) {
                        h.stream = s1;
                        throw;
                    }
                    h.stream = s1;
                }
            }
        }
        LFB_CUDA(cudaEventRecord(evJoin, s2));
        LFB_CUDA(cudaStreamWaitEvent(s1, evJoin, 0));           // the panel is complete on the first stream again
    };
    // Look-ahead: the trailing SYRK of panel k is split into the block column of panel k+1 (done
    // first) and the rest; panel k+1 is factored on a high-priority side stream while the rest of
    // the SYRK runs, so the latency-bound diagonal kernels leave the critical path.
    //
    // TN form of the trailing update (f64, option chol_tn): the finished panel P (rows x nb, column-major = "MN-major"
    // for both operands of P P^T, which costs the TMA GEMM 16 box loads per stage and its slower fragment path) is
    // transposed ONCE into Pt (nb x rows, K contiguous): the SYRK becomes C -= Pt^T Pt, K-major on both sides -- the
    // kernel's fastest variant (2 TMA loads per stage; 35.8 vs 31.3 TFLOP/s at K = 512 on this shape,
    // profiles/r1_gemm_bench.jsonl).  O(n nb) extra traffic per panel; two buffers because panel k+1 is transposed on the
    // side stream while the main stream still reads panel k's copy.
    cudaStream_t sm = h.stream, sp = h.aux_stream;
    const bool la = h.opt.lookahead && sp != nullptr && !h.prof_on && n >= 4 * NB;
    const bool tn = (sizeof(T) == 8 || h.opt.sgemm_tc != 0) && h.opt.chol_tn && n >= 2 * NB;   // f32: the tcgen05 kernel is TN only
    const int64_t ldpt = NB;
    std::unique_ptr<DevBuf<T>> PtBuf[2];
    if (tn)
        for (auto &b : PtBuf) b.reset(new DevBuf<T>(h, (size_t)ldpt * n));
    // C (rows x cols, lower part) -= P[0:rows, :] * P[0:cols, :]^T with P = panel rows r_off.. (and the matching slice of Pt)
    auto syrk = [&](const T *P, const T *Pt, int64_t r_off, int64_t rows_, int64_t cols_, int64_t nb, T *Cp) {
        if (tn) gemm<T>(h, 1, 0, rows_, cols_, nb, T(-1), Pt + r_off * ldpt, ldpt, Pt + r_off * ldpt, ldpt, T(1), Cp, ld, /*lower_only=*/1);
        else gemm<T>(h, 0, 1, rows_, cols_, nb, T(-1), P + r_off, ld, P + r_off, ld, T(1), Cp, ld, /*lower_only=*/1);
    };
    // Debug option chol_trace: CUDA-event time stamps of every stage on both streams (there is no nsys in the image), one
    // line per panel on stderr after the call: when the block-column SYRK, the side-stream panel and the rest of the SYRK
    // started and ended relative to the start of the factorisation.
    struct Mark { cudaEvent_t e; int panel; int what; };
    std::vector<Mark> marks;
    const bool trace = h.opt.chol_trace != 0;
    cudaEvent_t t_base = nullptr;
    auto mark = [&](cudaStream_t st, int panel, int what) {
        if (!trace) return;
        cudaEvent_t e;
        LFB_CUDA(cudaEventCreate(&e));
        LFB_CUDA(cudaEventRecord(e, st));
        marks.push_back({e, panel, what});
    };
    if (trace) {
        LFB_CUDA(cudaEventCreate(&t_base));
        LFB_CUDA(cudaEventRecord(t_base, h.stream));
    }
    // Panel width by position: NB while many rows remain, chol_nb_tail once fewer than chol_tail_rows are left -- there the
    // step is bound by the panel chain, whose granularity (not its length) is what a narrower panel improves, while the
    // SYRKs that lose efficiency at the smaller K are short anyway (n = 8192: 12.2 ms at nb 256 against 13.0 at 512).
    const int64_t NBT = std::min<int64_t>(NB, std::max<int64_t>(CB, round_up(h.opt.chol_nb_tail, CB)));
    auto width_at = [&](int64_t k0) { return std::min<int64_t>((n - k0 > h.opt.chol_tail_rows) ? NB : NBT, ncols - k0); };
    const int64_t nb0 = width_at(0);
    factor_panel(0, nb0);
    if (h.chol_panel_hook) h.chol_panel_hook(col_base, nb0);
    if (tn && ncols > nb0) transpose<T>(h, A + nb0, n - nb0, nb0, ld, PtBuf[0]->get(), ldpt);
    mark(h.stream, -1, 0);
    int cur = 0;
    int pi = -1;
    for (int64_t k0 = 0, nb_step = nb0; k0 < ncols; k0 += nb_step, cur ^= 1) {
        const int64_t nb = width_at(k0);
        nb_step = nb;
        ++pi;
        const int64_t pend = k0 + nb;
        const int64_t rows = n - pend;
        if (rows <= 0 || pend >= ncols) break;      // no further column of this call to update
        const int64_t nbn = width_at(pend);
        T *P = A + pend + k0 * ld;
        const T *Pt = tn ? PtBuf[cur]->get() : nullptr;
        T *PtNext = tn ? PtBuf[cur ^ 1]->get() : nullptr;
        const int64_t rows2 = rows - nbn;
        const int64_t cols2 = std::min(rows2, ncols - (pend + nbn));   // columns of the rest of the update (= rows2 for the whole matrix)
        if (la) {
            mark(sm, pi, 1);
            syrk(P, Pt, 0, rows, nbn, nb, A + pend + pend * ld);
            mark(sm, pi, 2);
            LFB_CUDA(cudaEventRecord(h.ev[0], sm));
            LFB_CUDA(cudaStreamWaitEvent(sp, h.ev[0], 0));
            h.stream = sp;
            try {
                mark(sp, pi, 3);
                if (split_ok) factor_panel_split(pend, nbn);
                else factor_panel(pend, nbn);
                mark(sp, pi, 4);
                if (h.chol_panel_hook) h.chol_panel_hook(col_base + pend, nbn);     // h.stream is the side stream here
                if (tn && cols2 > 0) transpose<T>(h, A + (pend + nbn) + pend * ld, rows2, nbn, ld, PtNext, ldpt);
            } catch (...) {
                h.stream = sm;
                throw;
            }
            LFB_CUDA(cudaEventRecord(h.ev[1], sp));
            h.stream = sm;
            if (cols2 > 0) syrk(P, Pt, nbn, rows2, cols2, nb, A + (pend + nbn) + (pend + nbn) * ld);
            mark(sm, pi, 5);
            LFB_CUDA(cudaStreamWaitEvent(sm, h.ev[1], 0));
        } else {
            syrk(P, Pt, 0, rows, std::min(rows, ncols - pend), nb, A + pend + pend * ld);
            factor_panel(pend, nbn);
            if (h.chol_panel_hook) h.chol_panel_hook(col_base + pend, nbn);
            if (tn && cols2 > 0) transpose<T>(h, A + (pend + nbn) + pend * ld, rows2, nbn, ld, PtNext, ldpt);
        }
    }
    if (trace) {
        LFB_CUDA(cudaStreamSynchronize(sm));
        LFB_CUDA(cudaStreamSynchronize(sp));
        std::map<int, std::array<float, 6>> rowsT;
        for (auto &mk : marks) {
            float t = 0.f;
            cudaEventElapsedTime(&t, t_base, mk.e);
            rowsT[mk.panel][mk.what] = t;
            cudaEventDestroy(mk.e);
        }
        cudaEventDestroy(t_base);
        fprintf(stderr, "chol_trace n=%lld nb=%lld (ms from start): panel | syrk_col start end | side panel start end | syrk_rest end\n",
                (long long)n, (long long)NB);
        for (auto &kv : rowsT)
            fprintf(stderr, "  %3d | %8.3f %8.3f | %8.3f %8.3f | %8.3f\n", kv.first, kv.second[1], kv.second[2], kv.second[3], kv.second[4],
                    kv.second[5]);
    }
    if (clean && ncols == n) triangular_zero<T>(h, A, n, ld, /*keep_lower=*/1);
}

template <typename T>
void cholesky_lower(lfb_handle &h, T *A, int64_t n, int64_t ld, int clean, int64_t *d_info) {
    // A "diagonal first" driver (side stream factors only the nb x nb diagonal block and builds L11^-1 by block doubling,
    // rows below solved by one GEMM with the explicit inverse) was built and measured in round 2: 64.4 ms against 58.0 ms
    // for this driver at n = 16384 (18.7 vs 13.5 ms at 8192) -- its diagonal chain is 8 potf2 + 30 tiny GEMMs of 19-39 us
    // each, ~1 ms per 512 columns, longer than what it removes.  Numbers in profiles/r2_chol_analysis.md; the code was
    // dropped rather than kept as a slower option.
    cholesky_lower_v1<T>(h, A, n, ld, clean, d_info);
}

// One arrival wave of the host path: columns [c0, c1) of the n x n matrix A, whose columns < c0 already hold L and whose columns
// [c0, c1) hold the caller's entries.  (a) catch-up: the trapezoid A[c0:, c0:c1] -= L[c0:, 0:c0] L[c0:c1, 0:c0]^T -- one product with
// K = c0 (the large-K regime of the GEMM, where a right-looking sweep would have made c0 / nb passes with K = nb); (b) the
// right-looking factorisation of the trapezoid with look-ahead.  first = 1 clears the failure index.
template <typename T>
void cholesky_lower_wave(lfb_handle &h, T *A, int64_t n, int64_t ld, int64_t c0, int64_t c1, int64_t *d_info, int first) {
    if (c1 <= c0) return;
    if (first) LFB_CUDA(cudaMemsetAsync(d_info, 0, sizeof(int64_t), h.stream));
    if (c0 > 0)
        gemm<T>(h, 0, 1, n - c0, c1 - c0, c0, T(-1), A + c0, ld, A + c0, ld, T(1), A + c0 + c0 * ld, ld, /*lower_only=*/1);
    cholesky_lower_v1<T>(h, A + c0 + c0 * ld, n - c0, ld, /*clean=*/0, d_info, c1 - c0, c0, /*reset_info=*/0);
}
template void cholesky_lower_wave<float>(lfb_handle &, float *, int64_t, int64_t, int64_t, int64_t, int64_t *, int);
template void cholesky_lower_wave<double>(lfb_handle &, double *, int64_t, int64_t, int64_t, int64_t, int64_t *, int);

// Device time (us per launch) of one diagonal-block kernel on a 64 x 64 SPD block: kind 0 = left-looking (first
// generation), 1 = right-looking register-blocked; +2 = factor only (no inverse).  The input is restored before every
// launch by a 32 KiB device copy, which is part of the measured time (< 2 us).
double microbench_potf2(lfb_handle &h, int kind, int reps) {
    using T = double;
    const size_t smem_p = sizeof(T) * (2 * CB * SP + 32 * SP);
    static DeviceOnce cfg;
    cfg.run(h.device, [&] {
        LFB_CUDA(cudaFuncSetAttribute(potf2_inv_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_p));
        LFB_CUDA(cudaFuncSetAttribute(potf2_rl_inv_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_p));
    });
    std::vector<T> host((size_t)CB * CB);
    for (int c = 0; c < CB; ++c)
        for (int r = 0; r < CB; ++r) host[r + c * CB] = r == c ? T(CB) : T(1) / T(1 + r + c);
    DevBuf<T> A0(h, CB * CB), A(h, CB * CB), Linv(h, CB * CB);
    DevBuf<int64_t> info(h, 1);
    LFB_CUDA(cudaMemcpyAsync(A0.get(), host.data(), sizeof(T) * CB * CB, cudaMemcpyHostToDevice, h.stream));
    LFB_CUDA(cudaMemsetAsync(info.get(), 0, sizeof(int64_t), h.stream));
    const int do_inv = (kind & 2) ? 0 : 1;
    auto once = [&]() {
        LFB_CUDA(cudaMemcpyAsync(A.get(), A0.get(), sizeof(T) * CB * CB, cudaMemcpyDeviceToDevice, h.stream));
        if (kind & 1) potf2_rl_inv_kernel<T><<<1, 256, smem_p, h.stream>>>(A.get(), CB, CB, 0, info.get(), Linv.get(), do_inv);
        else potf2_inv_kernel<T><<<1, 256, smem_p, h.stream>>>(A.get(), CB, CB, 0, info.get(), Linv.get(), do_inv);
        LFB_LAUNCH_CHECK(h);
    };
    cudaEvent_t e0, e1;
    LFB_CUDA(cudaEventCreate(&e0));
    LFB_CUDA(cudaEventCreate(&e1));
    for (int r = 0; r < 3; ++r) once();
    LFB_CUDA(cudaEventRecord(e0, h.stream));
    for (int r = 0; r < reps; ++r) once();
    LFB_CUDA(cudaEventRecord(e1, h.stream));
    LFB_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    LFB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return (double)ms * 1e3 / reps;
}

template void cholesky_lower<float>(lfb_handle &, float *, int64_t, int64_t, int, int64_t *);
template void cholesky_lower<double>(lfb_handle &, double *, int64_t, int64_t, int, int64_t *);

}  // namespace lfb
