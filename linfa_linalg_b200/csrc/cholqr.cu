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
#include <memory>

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

// The same guard numbers from the UPPER factors R = L^T and RI = R^-1 (the bound is symmetric under transposition of both), for
// the blocked 256-column stage; g0 / g1: the guard outputs of the two diagonal-block launches, whose failure flags are folded in.
template <typename T>
__global__ void __launch_bounds__(1024) cholqr_guard_upper_kernel(const T *__restrict__ R, int64_t ldr, const T *__restrict__ RI, int64_t ldi, int n,
                                                                  const double *__restrict__ g0, const double *__restrict__ g1, double *out) {
    __shared__ double rowR[1024], rowI[1024];
    __shared__ double colR[32], colI[32], dmin[32];
    __shared__ int okf[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < n; i += 1024) rowR[i] = rowI[i] = 0.0;
    __syncthreads();
    double mr = 0.0, mi = 0.0, md = 1e300;
    int ok = 1;
    for (int c = warp; c < n; c += 32) {
        double sr = 0.0, si = 0.0;
        for (int r = lane; r <= c; r += 32) {
            const double a = fabs((double)R[r + (int64_t)c * ldr]), b = fabs((double)RI[r + (int64_t)c * ldi]);
            sr += a;
            si += b;
            atomicAdd(&rowR[r], a);
            atomicAdd(&rowI[r], b);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            sr += __shfl_xor_sync(0xffffffffu, sr, o);
            si += __shfl_xor_sync(0xffffffffu, si, o);
        }
        if (!isfinite(sr) || !isfinite(si)) ok = 0;
        mr = fmax(mr, sr);
        mi = fmax(mi, si);
        md = fmin(md, (double)R[c + (int64_t)c * ldr]);
    }
    if (lane == 0) { colR[warp] = mr; colI[warp] = mi; dmin[warp] = md; okf[warp] = ok; }
    __syncthreads();
    if (warp == 0) {
        double rr = 0.0, ri = 0.0;
        for (int i = lane; i < n; i += 32) { rr = fmax(rr, rowR[i]); ri = fmax(ri, rowI[i]); }
        double cr = colR[lane], ci = colI[lane], dm = dmin[lane];
        int k = okf[lane];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            rr = fmax(rr, __shfl_xor_sync(0xffffffffu, rr, o));
            ri = fmax(ri, __shfl_xor_sync(0xffffffffu, ri, o));
            cr = fmax(cr, __shfl_xor_sync(0xffffffffu, cr, o));
            ci = fmax(ci, __shfl_xor_sync(0xffffffffu, ci, o));
            dm = fmin(dm, __shfl_xor_sync(0xffffffffu, dm, o));
            k &= __shfl_xor_sync(0xffffffffu, k, o);
        }
        if (lane == 0) {
            const bool blocks_ok = g0[2] == 1.0 && g1[2] == 1.0;
            out[0] = blocks_ok ? sqrt(cr * rr * ci * ri) : 1e300;
            out[1] = blocks_ok ? dm : 0.0;
            out[2] = (blocks_ok && k) ? 1.0 : 0.0;
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
    if (n == 256 && h.opt.cholqr_fused) {
        // 2 x 2 blocks of 128 on the single-CTA kernel: R = [R11, L21^T; 0, R22], R^-1 = [R11^-1, -R11^-1 L21^T R22^-1; 0, R22^-1]
        // with L21 = G21 R11^-1 and R22 = chol(G22 - L21 L21^T)^T -- ten short launches instead of ~25 (cholesky_lower, the
        // triangular solve for the inverse, guard, emit), which is most of what is left of a TSQR stage at 8 GPUs
        constexpr int b = 128;
        std::unique_ptr<DevBuf<T>> ritmp;
        T *RI = Rinv;
        int64_t ldi = ldri;
        if (!RI) { ritmp.reset(new DevBuf<T>(h, (size_t)ldg * n)); RI = ritmp->get(); ldi = ldg; }
        DevBuf<T> L21(h, b * b), T1(h, b * b);
        DevBuf<double> gb(h, 8);
        T *Gm = G.get();
        fill<T>(h, R + b, b, b, ldr, T(0), T(0));
        fill<T>(h, RI + b, b, b, ldi, T(0), T(0));
        cholqr128<T>(h, Gm, ldg, R, ldr, RI, ldi, gb.get());
        gemm<T>(h, 0, 0, b, b, b, T(1), Gm + b, ldg, RI, ldi, T(0), L21.get(), b);                                     // L21 = G21 R11^-1
        gemm<T>(h, 0, 1, b, b, b, T(-1), L21.get(), b, L21.get(), b, T(1), Gm + b + (int64_t)b * ldg, ldg, /*lower_only=*/1);
        cholqr128<T>(h, Gm + b + (int64_t)b * ldg, ldg, R + b + (int64_t)b * ldr, ldr, RI + b + (int64_t)b * ldi, ldi, gb.get() + 4);
        transpose<T>(h, L21.get(), b, b, b, R + (int64_t)b * ldr, ldr);                                                // R12 = L21^T
        gemm<T>(h, 1, 0, b, b, b, T(1), L21.get(), b, RI + b + (int64_t)b * ldi, ldi, T(0), T1.get(), b);              // L21^T R22^-1
        gemm<T>(h, 0, 0, b, b, b, T(-1), RI, ldi, T1.get(), b, T(0), RI + (int64_t)b * ldi, ldi);                      // (R^-1)12
        cholqr_guard_upper_kernel<T><<<1, 1024, 0, h.stream>>>(R, ldr, RI, ldi, (int)n, gb.get(), gb.get() + 4, guard.get());
        LFB_LAUNCH_CHECK(h);
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
