// Cholesky-QR leaf of the tall-skinny paths (TSQR local stage, qr.rs:29-45 on tall-skinny input).
//
// The Householder leaf (householder.cu: tsqr_local_chunks) spends its time in latency-bound panel kernels: ~100 launches per
// 12288-row chunk, 4 TFLOP/s on 4,194,304 x 256 (0.11 of the FP64 pipe).  For a tall block with n <= 512 columns
// the same R (unique once diag(R) >= 0, qr.rs:96) is
//
//     G = A^T A   (ONE split-K tensor-core GEMM over all the rows: K = rows, lower tiles only, A is read once from HBM)
//     G = L L^T   (n x n Cholesky, potrf.cu),   R = L^T,   R^-1 = (L^-1)^T
//
// with forward error ~ cond(A)^2 eps instead of cond(A) eps.  So the leaf is GUARDED: it is taken only when the Cholesky
// succeeds and the bound sqrt(||L||_1 ||L||_inf ||L^-1||_1 ||L^-1||_inf) >= cond_2(A) is <= `tsqr_cholqr_cond` (default 16: a loss of at most 256 eps, inside every
// parity tolerance of the suite); anything else -- rank deficiency, graded columns, ill-conditioning -- goes to the
// Householder leaf unchanged (A is not modified by the attempt).  The guard costs one 32-byte D2H copy and a stream sync.
// What the callers build on top (explicit Q = A R^-1, the reconstruction of the reference's reflectors) is in
// householder.cu / tsqr_hr.cu.
#include "common.cuh"

namespace lfb {
namespace {

// One CTA: an upper bound of cond_2(L) from L (lower triangular, ld) and X = L^-1,
//   cond_2 <= sqrt(||L||_1 ||L||_inf ||X||_1 ||X||_inf)      (||M||_2^2 <= ||M||_1 ||M||_inf),
// the smallest diagonal entry and finiteness.  Warp per column, lanes down the rows (coalesced); row sums through
// shared-memory atomics.  out[0] = the bound, out[1] = min diag, out[2] = 1 if every entry read was finite.
template <typename T>
__global__ void __launch_bounds__(1024) cholqr_guard_kernel(const T *__restrict__ Lm, const T *__restrict__ X, int64_t ld, int n, double *out) {
    __shared__ double rowL[1024], rowX[1024];
    __shared__ double colL[32], colX[32], dmin[32];
    __shared__ int okf[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < n; i += 1024) rowL[i] = rowX[i] = 0.0;
    __syncthreads();
    double ml = 0.0, mx = 0.0, md = 1e300;
    int ok = 1;
    for (int c = warp; c < n; c += 32) {
        double sl = 0.0, sx = 0.0;
        for (int r = c + lane; r < n; r += 32) {
            const double l = fabs((double)Lm[r + (int64_t)c * ld]), x = fabs((double)X[r + (int64_t)c * ld]);
            sl += l;
            sx += x;
            atomicAdd(&rowL[r], l);
            atomicAdd(&rowX[r], x);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            sl += __shfl_xor_sync(0xffffffffu, sl, o);
            sx += __shfl_xor_sync(0xffffffffu, sx, o);
        }
        if (!isfinite(sl) || !isfinite(sx)) ok = 0;
        ml = fmax(ml, sl);
        mx = fmax(mx, sx);
        md = fmin(md, (double)Lm[c + (int64_t)c * ld]);
    }
    if (lane == 0) { colL[warp] = ml; colX[warp] = mx; dmin[warp] = md; okf[warp] = ok; }
    __syncthreads();
    if (warp == 0) {
        double rl = 0.0, rx = 0.0;
        for (int i = lane; i < n; i += 32) { rl = fmax(rl, rowL[i]); rx = fmax(rx, rowX[i]); }
        double cl = colL[lane], cx = colX[lane], dm = dmin[lane];
        int k = okf[lane];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            rl = fmax(rl, __shfl_xor_sync(0xffffffffu, rl, o));
            rx = fmax(rx, __shfl_xor_sync(0xffffffffu, rx, o));
            cl = fmax(cl, __shfl_xor_sync(0xffffffffu, cl, o));
            cx = fmax(cx, __shfl_xor_sync(0xffffffffu, cx, o));
            dm = fmin(dm, __shfl_xor_sync(0xffffffffu, dm, o));
            k &= __shfl_xor_sync(0xffffffffu, k, o);
        }
        if (lane == 0) {
            out[0] = sqrt(cl * rl * cx * rx);
            out[1] = dm;
            out[2] = (double)k;
        }
    }
}

// R = L^T (upper, strict lower zeroed) and, if asked, Rinv = X^T with X = L^-1.
template <typename T>
__global__ void cholqr_emit_kernel(const T *__restrict__ Lm, const T *__restrict__ X, int64_t ld, int n, T *__restrict__ R, int64_t ldr,
                                   T *__restrict__ Rinv, int64_t ldri) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;   // row of R
    if (i >= n) return;
    for (int j = blockIdx.y; j < n; j += gridDim.y) {
        R[i + (int64_t)j * ldr] = i <= j ? Lm[j + (int64_t)i * ld] : T(0);
        if (Rinv) Rinv[i + (int64_t)j * ldri] = i <= j ? X[j + (int64_t)i * ld] : T(0);
    }
}

}  // namespace

template <typename T>
bool cholqr_factor(lfb_handle &h, const T *A, int64_t rows, int64_t n, int64_t ld, T *R, int64_t ldr, T *Rinv, int64_t ldri) {
    if (n <= 0 || rows < n || n > 1024 || h.opt.tsqr_cholqr_cond <= 0 || h.in_capture) return false;   // (the guard synchronises: not capturable)
    const int64_t ldg = round_up(n, 2);
    DevBuf<T> G(h, (size_t)ldg * n), X(h, (size_t)ldg * n);
    DevBuf<int64_t> info(h, 1);
    DevBuf<double> guard(h, 4);
    gemm<T>(h, 1, 0, n, n, rows, T(1), A, ld, A, ld, T(0), G.get(), ldg, /*lower_only=*/1);      // G = A^T A
    double hg[3] = {0, 0, 0};
    if (n == 128 && h.opt.cholqr_fused && Rinv) {
        // one single-CTA launch for the n x n stage (panel_hr.cu): Cholesky, inverse, guard numbers, R and R^-1 -- the scratch
        // outputs are written before the verdict is known, the caller's matrix is not touched
        cholqr128<T>(h, G.get(), ldg, R, ldr, Rinv, ldri, guard.get());
        LFB_CUDA(cudaMemcpyAsync(hg, guard.get(), sizeof hg, cudaMemcpyDeviceToHost, h.stream));
        LFB_CUDA(cudaStreamSynchronize(h.stream));
        return hg[2] == 1.0 && hg[1] > 0.0 && std::isfinite(hg[0]) && hg[0] <= (double)h.opt.tsqr_cholqr_cond;
    }
    cholesky_lower<T>(h, G.get(), n, ldg, /*clean=*/0, info.get());                            // G = L L^T (lower)
    fill<T>(h, X.get(), n, n, ldg, T(0), T(1));
    trsm_left<T>(h, /*lower=*/1, /*trans=*/0, n, n, G.get(), ldg, (const T *)nullptr, X.get(), ldg);   // X = L^-1
    cholqr_guard_kernel<T><<<1, 1024, 0, h.stream>>>(G.get(), X.get(), ldg, (int)n, guard.get());
    LFB_LAUNCH_CHECK(h);
    int64_t hinfo = 0;
    LFB_CUDA(cudaMemcpyAsync(hg, guard.get(), sizeof hg, cudaMemcpyDeviceToHost, h.stream));
    LFB_CUDA(cudaMemcpyAsync(&hinfo, info.get(), sizeof hinfo, cudaMemcpyDeviceToHost, h.stream));
    LFB_CUDA(cudaStreamSynchronize(h.stream));
    const bool ok = hinfo == 0 && hg[2] == 1.0 && hg[1] > 0.0 && std::isfinite(hg[0]) && hg[0] <= (double)h.opt.tsqr_cholqr_cond;
    if (!ok) return false;
    dim3 grid((unsigned)cdiv(n, 128), (unsigned)(n < 65535 ? n : 65535));
    cholqr_emit_kernel<T><<<grid, 128, 0, h.stream>>>(G.get(), X.get(), ldg, (int)n, R, ldr, Rinv, ldri);
    LFB_LAUNCH_CHECK(h);
    return true;
}

template bool cholqr_factor<float>(lfb_handle &, const float *, int64_t, int64_t, int64_t, float *, int64_t, float *, int64_t);
template bool cholqr_factor<double>(lfb_handle &, const double *, int64_t, int64_t, int64_t, double *, int64_t, double *, int64_t);

}  // namespace lfb
