// Blocked Golub-Kahan bidiagonalisation (the dgebrd / dlabrd structure, reference sign conventions).
//
// Replaces src/bidiagonal.rs:27-59 (alternating clear_column / clear_row, householder.rs:34-63).  The
// reference applies every reflector to the whole trailing matrix (three passes: GEMV, read and write of
// the rank-1 update = 48 B of HBM traffic per trailing element per column/row pair for f64).  Here the
// rank-1 updates of a panel of 32 column/row reflector pairs are deferred,
//     B_cur = B0 - V Y^T - X U^T       (V, U: unit-norm reflectors; y = 2 B^T v, x = 2 B u),
// so a pair costs TWO streaming passes over the stored trailing matrix (y_raw = B0^T v, x_raw = B0 u:
// 16 B/element, HBM bound) plus skinny panel corrections, and the trailing matrix is rewritten once per
// panel by two tensor-core GEMMs (K = 32).  Four launches per pair:
//   bd_head (column)  one thread-block cluster: finish x of the previous pair, bring column i up to
//                     date, make the reflector (householder.rs:9-28), V^T v and X^T v
//   bd_gemv<T>        y_raw = B0[i:, i+1:]^T v        (64x32 register tiles landed with cp.async)
//   bd_head (row)     finish y, bring row i up to date, reflector u, Y^T u and U^T u
//   bd_gemv<N>        x_raw = B0[i+1:, i+1:] u
//
// Signs.  clear_column / clear_row multiply the WHOLE remaining block by signum(returned pivot)
// (householder.rs:45-48), i.e. the reference's active block is sigma * B for a running scalar
// sigma = +-1.  reflection_axis_mut of sigma*x is sigma*v with pivot sigma*beta, so the engine runs the
// sign-free recurrence on B and applies sigma only to what the reference exposes: the stored
// reflectors, d and e.  A `None` reflector (householder.rs:22-27) is v = 0 and leaves sigma alone.
//
// rows < cols (bidiagonal.rs:45-51) is the same algorithm on the transpose (clear_row is literally
// clear_column on the reversed-axes view, householder.rs:57-63): the wide matrix is transposed on the
// device, factored as a tall one and transposed back.
#include <cooperative_groups.h>

#include <algorithm>

#include "common.cuh"
#include "dev_utils.cuh"

namespace cg = cooperative_groups;

namespace lfb {

template <typename T> void bidiagonal_unblocked(lfb_handle &h, T *A, int64_t rows, int64_t cols, int64_t ld, T *d, T *e);

namespace {

using namespace dev;

constexpr int TB = 32;     // reflector pairs per panel
constexpr int HNT = 512;   // threads per CTA of the head kernel
constexpr int SW = 32;     // columns per GEMV tile

template <typename T>
struct BdState {
    T sigma;            // running sign of the reference's active block
    int some;           // last reflector was Some(..)
    T tV[TB + 1], tX[TB + 1];    // head(column) -> head(row):   V^T v, X^T v
    T tY[TB + 1], tU[TB + 1];    // head(row) -> head(column):   Y^T u, U^T u
};

template <typename T>
__global__ void bd_init_kernel(BdState<T> *st) {
    if (threadIdx.x == 0) { st->sigma = T(1); st->some = 0; }
    for (int k = threadIdx.x; k <= TB; k += blockDim.x) { st->tV[k] = T(0); st->tX[k] = T(0); st->tY[k] = T(0); st->tU[k] = T(0); }
}

// ---------------------------------------------------------------------------------------------------
// y (+)= alpha * op(B0) x on the sub-matrix rows [ra, m) x cols [ca, n) of the column-major matrix A, all
// indices GLOBAL.  MODE 0: y[r] += alpha sum_c A[r,c] x[c].  MODE 1: y[c] += alpha sum_r A[r,c] x[r].
// A warp owns 64-row groups (2 rows per lane, one 16-byte cp.async per column of a 32-column tile; every
// lane reads back only what it fetched, so no barrier).  CTA = (cstrips tiles of columns, rch rows).
template <typename T, int MODE>
__global__ void __launch_bounds__(128, 3) bd_gemv_kernel(const T *__restrict__ A, int64_t ld, int64_t m, int64_t n, int64_t ra,
                                                         int64_t ca, const T *__restrict__ x, T alpha, T *y, int rch, int cstrips,
                                                         const int *some) {
    if (*some == 0) return;
    using V2 = typename Vec2<T>::type;
    extern __shared__ __align__(16) unsigned char gemv_smem[];
    __shared__ T sx[256];          // MODE 0: x over the CTA's columns (cstrips <= 8)
    __shared__ T scol[4][SW];      // MODE 1: per-warp column sums
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int64_t C0 = ca + (int64_t)blockIdx.x * SW * cstrips;
    const int64_t R0 = (ra & ~(int64_t)1) + (int64_t)blockIdx.y * rch;
    if (R0 >= m || C0 >= n) return;
    if (MODE == 0) {
        for (int t = threadIdx.x; t < SW * cstrips; t += blockDim.x) sx[t] = (C0 + t < n) ? x[C0 + t] : T(0);
        __syncthreads();
    }
    V2 *buf = reinterpret_cast<V2 *>(gemv_smem) + (size_t)warp * (SW * 32) + lane;
    T col[SW];
#pragma unroll
    for (int k = 0; k < SW; ++k) col[k] = T(0);
    for (int g = warp; g < rch / 64; g += nw) {
        const int64_t g0 = R0 + 64 * (int64_t)g;
        if (g0 >= m) break;
        const int64_t gr = g0 + 2 * lane;
        const bool v0 = gr >= ra && gr < m, v1 = gr + 1 >= ra && gr + 1 < m;
        const bool rows_full = g0 >= ra && g0 + 64 <= m;
        T x0 = T(0), x1 = T(0);
        if (MODE == 1) {
            if (v0) x0 = x[gr];
            if (v1) x1 = x[gr + 1];
        }
        T r0 = T(0), r1 = T(0);
        for (int s = 0; s < cstrips; ++s) {
            const int64_t Cs = C0 + (int64_t)s * SW;
            if (Cs >= n) break;
            const T *p = A + gr + Cs * ld;
            const T *xs_ = sx + s * SW;
            if (rows_full && Cs + SW <= n) {                                // interior tile
#pragma unroll
                for (int cb = 0; cb < SW; cb += 8) {
#pragma unroll
                    for (int q = 0; q < 8; ++q) cp_async<(int)sizeof(V2)>(buf + (cb + q) * 32, p + (int64_t)(cb + q) * ld);
                    cp_async_commit();
                }
#define LFB_GEMV_STEP(CB, PENDING)                                                              \
                cp_async_wait<PENDING>();                                                           \
                _Pragma("unroll") for (int q = 0; q < 8; ++q) {                                      \
                    const V2 a = buf[((CB) + q) * 32];                                               \
                    if (MODE == 0) {                                                                 \
                        const T xs = xs_[(CB) + q];                                                  \
                        r0 += a.x * xs;                                                              \
                        r1 += a.y * xs;                                                              \
                    } else {                                                                         \
                        col[(CB) + q] += a.x * x0 + a.y * x1;                                        \
                    }                                                                                \
                }
                LFB_GEMV_STEP(0, 3)
                LFB_GEMV_STEP(8, 2)
                LFB_GEMV_STEP(16, 1)
                LFB_GEMV_STEP(24, 0)
#undef LFB_GEMV_STEP
            } else {                                                        // edge tile: masked direct loads
#pragma unroll
                for (int cb = 0; cb < SW; cb += 8) {
                    T a0[8], a1[8];
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const bool cok = Cs + cb + q < n;
                        a0[q] = (v0 && cok) ? p[(int64_t)(cb + q) * ld] : T(0);
                        a1[q] = (v1 && cok) ? p[(int64_t)(cb + q) * ld + 1] : T(0);
                    }
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        if (MODE == 0) {
                            const T xs = xs_[cb + q];
                            r0 += a0[q] * xs;
                            r1 += a1[q] * xs;
                        } else {
                            col[cb + q] += a0[q] * x0 + a1[q] * x1;
                        }
                    }
                }
            }
        }
        if (MODE == 0) {
            if (v0) atomicAdd(y + gr, alpha * r0);
            if (v1) atomicAdd(y + gr + 1, alpha * r1);
        }
    }
    if (MODE == 1) {   // cstrips == 1: one tile column per CTA; combine the warps, 32 atomics per CTA
        warp_transpose_reduce<T>(col, lane);
        scol[warp][lane] = col[0];
        __syncthreads();
        if (warp == 0) {
            T sum = T(0);
            for (int w = 0; w < nw; ++w) sum += scol[w][lane];
            if (C0 + lane < n) atomicAdd(y + C0 + lane, alpha * sum);
        }
    }
}

// Unaligned fallback (odd leading dimension / unaligned base): thread per row or warp per column.
template <typename T, int MODE>
__global__ void __launch_bounds__(256) bd_gemv_simple_kernel(const T *__restrict__ A, int64_t ld, int64_t m, int64_t n, int64_t ra,
                                                             int64_t ca, const T *__restrict__ x, T alpha, T *y, const int *some) {
    if (*some == 0) return;
    if (MODE == 0) {
        const int64_t r = ra + blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
        if (r >= m) return;
        T acc = T(0);
        for (int64_t c = ca; c < n; ++c) acc += A[r + c * ld] * x[c];
        y[r] += alpha * acc;
    } else {
        const int lane = threadIdx.x & 31;
        const int64_t c = ca + blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
        if (c >= n) return;
        T acc = T(0);
        for (int64_t r = ra + lane; r < m; r += 32) acc += A[r + c * ld] * x[r];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) y[c] += alpha * acc;
    }
}

template <typename T, int MODE>
void launch_gemv(lfb_handle &h, const T *A, int64_t ld, int64_t m, int64_t n, int64_t ra, int64_t ca, const T *x, T alpha, T *y,
                 const int *some) {
    const int64_t nr = m - ra, ncol = n - ca;
    if (nr <= 0 || ncol <= 0) return;
    using V2 = typename Vec2<T>::type;
    const bool aligned = ((uintptr_t)A % (2 * sizeof(T)) == 0) && (ld % 2 == 0);
    if (!aligned) {
        if (MODE == 0) bd_gemv_simple_kernel<T, 0><<<(unsigned)cdiv(nr, 256), 256, 0, h.stream>>>(A, ld, m, n, ra, ca, x, alpha, y, some);
        else bd_gemv_simple_kernel<T, 1><<<(unsigned)cdiv(ncol, 8), 256, 0, h.stream>>>(A, ld, m, n, ra, ca, x, alpha, y, some);
        LFB_LAUNCH_CHECK(h);
        return;
    }
    static DeviceOnce cfg;   // function attributes are per device
    cfg.run(h.device, [&] {
        LFB_CUDA(cudaFuncSetAttribute(bd_gemv_kernel<T, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(4 * SW * 32 * sizeof(V2))));
    });
    const int64_t target = 4 * 3 * (int64_t)h.sm_count;     // >= 4 waves of 3 CTAs per SM
    int rch, cstrips;
    if (MODE == 0) {     // few atomics: 256 rows x up to 256 columns per CTA
        rch = 256; cstrips = 8;
        while (cstrips > 1 && cdiv(nr, rch) * cdiv(ncol, SW * cstrips) < target) cstrips >>= 1;
        while (rch > 64 && cdiv(nr, rch) * cdiv(ncol, SW * cstrips) < target) rch >>= 1;
    } else {             // one 32-column tile column x up to 1024 rows per CTA
        cstrips = 1; rch = 1024;
        while (rch > 64 && cdiv(nr, rch) * cdiv(ncol, SW) < target) rch >>= 1;
    }
    const int nw = std::min(4, rch / 64);
    const int64_t gbase = ra & ~(int64_t)1;
    dim3 grid((unsigned)cdiv(ncol, SW * cstrips), (unsigned)cdiv(m - gbase, rch));
    bd_gemv_kernel<T, MODE><<<grid, 32 * nw, (size_t)nw * SW * 32 * sizeof(V2), h.stream>>>(A, ld, m, n, ra, ca, x, alpha, y, rch, cstrips, some);
    LFB_LAUNCH_CHECK(h);
}

// ---------------------------------------------------------------------------------------------------
// The head kernel, one cluster launch.  "Long axis" = rows for the column step (kind 0), columns for
// the row step (kind 1); index t runs over [t0, tn).  P1 / P2 are the two panel matrices indexed by
// the long axis (column step: V, X; row step: Y, U).
template <typename T>
struct HeadArgs {
    int kind;              // 0 = column step (clear_column), 1 = row step (clear_row)
    int j;                 // position in the panel
    int final_only;        // only finish the pending companion column (panel end)
    int do_fin;            // a companion column is pending (column step: j > 0; row step: always)
    int64_t t0, tn;        // long-axis range
    T *src; int64_t sinc;  // reflector source, element t at src[t * sinc]   (A[t, i] or A[i, t])
    T *P1, *P2; int64_t ldp;
    int n1, n2;            // columns of P1 / P2 already final (used by the correction and the update)
    T *fin;                // companion column being finished (P2[:, j-1] or P1[:, j]), indexed by t
    const T *raw;          // its raw GEMV result, indexed by t
    const T *ta, *tb;      // coefficients of the correction (from the previous head)
    const T *ca, *cb; int64_t ldc;  // row i of the OTHER panel pair: ca[k * ldc], cb[k * ldc]
    int nca, ncb;          // entries of ca / cb in use (includes the one multiplying the finished column)
    int fin_in_p1;         // the finished column belongs to P1 (row step) or P2 (column step)
    T *pout;               // new reflector column (P1[:, j] for the column step, P2[:, j] for the row step)
    T *pivot;              // d[i] or e[i]
    T *t1out, *t2out; int n1t, n2t;   // P1^T w (n1t entries), P2^T w (n2t entries)
    T *zero; int64_t z0, zn;          // raw vector of the GEMV that follows: zero[z0..zn)
    BdState<T> *st;
};

template <typename T>
__global__ void __launch_bounds__(HNT) bd_head_kernel(const HeadArgs<T> a) {
    cg::cluster_group cl = cg::this_cluster();
    const int nc = (int)cl.num_blocks(), b = (int)cl.block_rank();
    __shared__ T slotB[2], resB[2];
    __shared__ T sred[HNT / 32][2];
    __shared__ T stt[2 * TB + 2];
    __shared__ T ta[TB + 1], tb[TB + 1], ca[TB + 1], cb[TB + 1];
    __shared__ T sw[HNT];
    __shared__ T inbox[16][2 * TB + 2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t stride = (int64_t)nc * HNT;
    const int64_t first = a.t0 + (int64_t)b * HNT + threadIdx.x;
    const bool single = first + stride >= a.tn;        // at most one entry per thread: it stays in a register
    T keep_x = T(0);
    const T sigma = a.st->sigma;
    if (threadIdx.x <= TB) {
        const int k = threadIdx.x;
        ta[k] = (a.do_fin && k < a.n1) ? a.ta[k] : T(0);
        tb[k] = (a.do_fin && k < a.n2) ? a.tb[k] : T(0);
        ca[k] = (!a.final_only && k < a.nca) ? a.ca[(int64_t)k * a.ldc] : T(0);
        cb[k] = (!a.final_only && k < a.ncb) ? a.cb[(int64_t)k * a.ldc] : T(0);
    }
    if (threadIdx.x < 2 * TB + 2) stt[threadIdx.x] = T(0);
    if (!a.final_only)
        for (int64_t z = a.z0 + (int64_t)b * HNT + threadIdx.x; z < a.zn; z += stride) a.zero[z] = T(0);
    __syncthreads();
    // coefficient of the finished companion column in the update of the source vector
    const T cfin = a.do_fin ? (a.fin_in_p1 ? ca[a.n1] : cb[a.n2]) : T(0);
    // ---- one pass over the panel rows: finish the companion (x = 2 (B u) or y = 2 (B^T v) on the deferred
    //      matrix), bring the source vector up to date, ||.||^2 and head element ----
    T part = T(0), headv = T(0);
    const int nk = a.n1 > a.n2 ? a.n1 : a.n2;
    for (int64_t t = first; t < a.tn; t += stride) {
        T f = a.do_fin ? a.raw[t] : T(0), S = T(0);
        for (int kb = 0; kb < nk; kb += 16) {
            T p1[16], p2[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                p1[q] = (kb + q < a.n1) ? a.P1[t + (int64_t)(kb + q) * a.ldp] : T(0);
                p2[q] = (kb + q < a.n2) ? a.P2[t + (int64_t)(kb + q) * a.ldp] : T(0);
            }
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                f -= p1[q] * ta[kb + q] + p2[q] * tb[kb + q];
                S += p1[q] * ca[kb + q] + p2[q] * cb[kb + q];
            }
        }
        if (a.do_fin) {
            f *= T(2);
            a.fin[t] = f;
            S += f * cfin;
        }
        if (!a.final_only) {
            const T xx = a.src[t * a.sinc] - S;
            if (!single) a.src[t * a.sinc] = xx;
            keep_x = xx;                                   // a thread that owns one entry keeps it in a register
            part += xx * xx;
            if (t == a.t0) headv = xx;
        }
    }
    if (a.final_only) return;
    T nsq, fh;
    cluster_sum2<T, HNT>(cl, part, headv, slotB, sred, resB, nsq, fh);
    // ---- householder.rs:9-28 on the sign-free vector; the reference sees sigma * (..) ----
    const T nrm = t_sqrt(nsq);
    const T s = t_signum(fh) * nrm;
    const T newsq = (nsq + t_abs(fh) * nrm) * T(2);
    const bool some = newsq != T(0);
    const T dd = t_sqrt(newsq);
    // ---- the reflector, then P1^T w and P2^T w: the CTA's entries of one pass are a contiguous block of HNT; w is
    //      staged in shared memory and warp k % 16 takes panel column k of both matrices (coalesced, independent loads) ----
    const int nt = a.n1t > a.n2t ? a.n1t : a.n2t;
    for (int64_t base = a.t0 + (int64_t)b * HNT; base < a.tn; base += stride) {
        const int64_t t = base + threadIdx.x;
        T w = T(0);
        if (t < a.tn) {
            const T xx = single ? keep_x : a.src[t * a.sinc];
            if (some) w = ((t == a.t0) ? xx + s : xx) / dd;
            a.src[t * a.sinc] = sigma * (some ? w : xx);     // what the reference stores (householder.rs:23 on sigma x)
            a.pout[t] = w;
        }
        sw[threadIdx.x] = w;
        __syncthreads();
        if (some) {
            for (int k = warp; k < nt; k += HNT / 32) {
                const T *p1 = a.P1 + (int64_t)k * a.ldp + base, *p2 = a.P2 + (int64_t)k * a.ldp + base;
                const bool k1 = k < a.n1t, k2 = k < a.n2t;
                T l1[HNT / 32], l2[HNT / 32];
#pragma unroll
                for (int q = 0; q < HNT / 32; ++q) {
                    const int tt = lane + 32 * q;
                    const bool ok = base + tt < a.tn;
                    l1[q] = (ok && k1) ? p1[tt] : T(0);
                    l2[q] = (ok && k2) ? p2[tt] : T(0);
                }
                T s1 = T(0), s2 = T(0);
#pragma unroll
                for (int q = 0; q < HNT / 32; ++q) {
                    s1 += l1[q] * sw[lane + 32 * q];
                    s2 += l2[q] * sw[lane + 32 * q];
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
                    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
                }
                if (lane == 0) { stt[k] += s1; stt[TB + 1 + k] += s2; }   // warp k % 16 owns entry k
            }
        }
        __syncthreads();
    }
    if (threadIdx.x < 2 * TB + 2) cl.map_shared_rank(&inbox[0][0], 0)[b * (2 * TB + 2) + threadIdx.x] = stt[threadIdx.x];
    cl.sync();
    if (b == 0) {
        if (threadIdx.x < 2 * TB + 2) {
            T sum = T(0);
            for (int rk = 0; rk < nc; ++rk) sum += inbox[rk][threadIdx.x];
            if (threadIdx.x <= TB) { if (threadIdx.x < a.n1t) a.t1out[threadIdx.x] = sum; }
            else if (threadIdx.x - (TB + 1) < a.n2t) a.t2out[threadIdx.x - (TB + 1)] = sum;
        }
        if (threadIdx.x == 0) {
            const T piv = some ? sigma * (-s) : T(0);        // householder.rs:24 on sigma x
            *a.pivot = piv;
            a.st->some = some ? 1 : 0;
            if (some) a.st->sigma = sigma * t_signum(piv);   // householder.rs:45-48: the block is scaled by signum(pivot)
        }
    }
}

template <typename T>
bool launch_head(lfb_handle &h, const HeadArgs<T> &args) {
    static DeviceOnce cfg;   // function attributes are per device
    cfg.run(h.device, [&] {
        cudaFuncSetAttribute(bd_head_kernel<T>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        cudaGetLastError();
    });
    const int64_t L = args.tn - args.t0;
    int nc = 1;
    while (nc < 16 && (int64_t)nc * HNT < L) nc <<= 1;
    const int cap = (int)std::max<int64_t>(1, std::min<int64_t>(16, h.opt.panel_cluster_max));
    while (nc > cap) nc >>= 1;
    cudaLaunchConfig_t c = {};
    c.gridDim = dim3((unsigned)nc);
    c.blockDim = dim3(HNT);
    c.dynamicSmemBytes = 0;
    c.stream = h.stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)nc;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    c.attrs = attr;
    c.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&c, bd_head_kernel<T>, args);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    h.launches++;
    return true;
}

// m >= n.  Returns false (nothing modified) if the cluster kernel cannot be launched.
template <typename T>
bool bidiagonal_tall(lfb_handle &h, T *A, int64_t m, int64_t n, int64_t ld, T *d, T *e) {
    const int64_t ldv = round_up(m, 2), ldy = round_up(n, 2);
    DevBuf<T> V(h, (size_t)ldv * TB), X(h, (size_t)ldv * TB), Y(h, (size_t)ldy * TB), U(h, (size_t)ldy * TB);
    DevBuf<T> xraw(h, ldv), yraw(h, ldy);
    DevBuf<BdState<T>> st(h, 1);
    char *sb = reinterpret_cast<char *>(st.get());
    T *tV = reinterpret_cast<T *>(sb + offsetof(BdState<T>, tV)), *tX = reinterpret_cast<T *>(sb + offsetof(BdState<T>, tX));
    T *tY = reinterpret_cast<T *>(sb + offsetof(BdState<T>, tY)), *tU = reinterpret_cast<T *>(sb + offsetof(BdState<T>, tU));
    const int *some = reinterpret_cast<const int *>(sb + offsetof(BdState<T>, some));
    bd_init_kernel<T><<<1, 128, 0, h.stream>>>(st.get());
    LFB_LAUNCH_CHECK(h);
    bool first = true;
    auto column_head = [&](int64_t i, int j, int final_only) {
        HeadArgs<T> c = {};
        c.kind = 0; c.j = j; c.final_only = final_only; c.do_fin = j > 0;
        c.t0 = i; c.tn = m;
        c.src = A + i * ld; c.sinc = 1;
        c.P1 = V.get(); c.P2 = X.get(); c.ldp = ldv;
        c.n1 = j; c.n2 = j > 0 ? j - 1 : 0;
        c.fin = X.get() + (int64_t)(j > 0 ? j - 1 : 0) * ldv; c.raw = xraw.get();
        c.ta = tY; c.tb = tU;
        c.ca = Y.get() + i; c.cb = U.get() + i; c.ldc = ldy;
        c.nca = j; c.ncb = j;
        c.fin_in_p1 = 0;
        c.pout = V.get() + (int64_t)(j < TB ? j : 0) * ldv;
        c.pivot = d + (i < n ? i : 0);
        c.t1out = tV; c.t2out = tX; c.n1t = j; c.n2t = j;
        c.zero = yraw.get(); c.z0 = i + 1; c.zn = n;
        c.st = st.get();
        return launch_head<T>(h, c);
    };
    auto row_head = [&](int64_t i, int j) {
        HeadArgs<T> c = {};
        c.kind = 1; c.j = j; c.final_only = 0; c.do_fin = 1;
        c.t0 = i + 1; c.tn = n;
        c.src = A + i; c.sinc = ld;
        c.P1 = Y.get(); c.P2 = U.get(); c.ldp = ldy;
        c.n1 = j; c.n2 = j;
        c.fin = Y.get() + (int64_t)j * ldy; c.raw = yraw.get();
        c.ta = tV; c.tb = tX;
        c.ca = V.get() + i; c.cb = X.get() + i; c.ldc = ldv;
        c.nca = j + 1; c.ncb = j;
        c.fin_in_p1 = 1;
        c.pout = U.get() + (int64_t)j * ldy;
        c.pivot = e + i;
        c.t1out = tY; c.t2out = tU; c.n1t = j + 1; c.n2t = j;
        c.zero = xraw.get(); c.z0 = i + 1; c.zn = m;
        c.st = st.get();
        return launch_head<T>(h, c);
    };
    for (int64_t i0 = 0; i0 < n; i0 += TB) {
        const int pb = (int)std::min<int64_t>(TB, n - i0);
        for (int j = 0; j < pb; ++j) {
            const int64_t i = i0 + j;
            if (!column_head(i, j, 0)) {                                                     // clear_column(i, 0), bidiagonal.rs:39
                if (first) return false;
                throw CudaError(LFB_ERR_CUDA, "bd_head_kernel launch failed");
            }
            first = false;
            if (i + 1 >= n) break;                                                           // last column: no row step (:43)
            launch_gemv<T, 1>(h, A, ld, m, n, i, i + 1, V.get() + (int64_t)j * ldv, T(1), yraw.get(), some);     // y_raw = B0[i:, i+1:]^T v
            if (!row_head(i, j)) throw CudaError(LFB_ERR_CUDA, "bd_head_kernel launch failed");                  // clear_row(i, 1), :40
            launch_gemv<T, 0>(h, A, ld, m, n, i + 1, i + 1, U.get() + (int64_t)j * ldy, T(1), xraw.get(), some);  // x_raw = B0[i+1:, i+1:] u
        }
        const int64_t R0 = i0 + pb;
        if (R0 < n) {
            if (!column_head(R0, pb, 1)) throw CudaError(LFB_ERR_CUDA, "bd_head_kernel launch failed");           // finish X[:, pb-1]
            T *C = A + R0 + R0 * ld;
            gemm<T>(h, 0, 1, m - R0, n - R0, pb, T(-1), V.get() + R0, ldv, Y.get() + R0, ldy, T(1), C, ld);
            gemm<T>(h, 0, 1, m - R0, n - R0, pb, T(-1), X.get() + R0, ldv, U.get() + R0, ldy, T(1), C, ld);
        }
    }
    return true;
}

}  // namespace

// Device time (us per launch, CUDA events, back to back) of the streaming GEMV of the bidiagonalisation on an
// m x n matrix: kind 0 = x_raw = B u (row sums), 1 = y_raw = B^T v (column sums).
double microbench_bd_gemv(lfb_handle &h, int kind, int64_t m, int64_t n, int reps) {
    using T = double;
    const int64_t ld = round_up(m, 2);
    DevBuf<T> A(h, (size_t)ld * n), x(h, std::max(ld, n) + 2), y(h, std::max(ld, n) + 2);
    DevBuf<BdState<T>> st(h, 1);
    LFB_CUDA(cudaMemsetAsync(A.get(), 0, sizeof(T) * ld * n, h.stream));
    LFB_CUDA(cudaMemsetAsync(x.get(), 0, sizeof(T) * (std::max(ld, n) + 2), h.stream));
    LFB_CUDA(cudaMemsetAsync(y.get(), 0, sizeof(T) * (std::max(ld, n) + 2), h.stream));
    BdState<T> hs;
    memset(&hs, 0, sizeof hs);
    hs.some = 1;
    LFB_CUDA(cudaMemcpyAsync(st.get(), &hs, sizeof hs, cudaMemcpyHostToDevice, h.stream));
    const int *some = reinterpret_cast<const int *>(reinterpret_cast<char *>(st.get()) + offsetof(BdState<T>, some));
    cudaEvent_t e0, e1;
    LFB_CUDA(cudaEventCreate(&e0));
    LFB_CUDA(cudaEventCreate(&e1));
    auto once = [&]() {
        if (kind == 0) launch_gemv<T, 0>(h, A.get(), ld, m, n, 1, 1, x.get(), T(1), y.get(), some);
        else launch_gemv<T, 1>(h, A.get(), ld, m, n, 0, 1, x.get(), T(1), y.get(), some);
    };
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

// bidiagonal.rs:27-59.  A: rows x cols column-major; d (min(rows, cols)) and e (min - 1) get the signed
// pivots, A the reflectors, exactly as the reference leaves them.
template <typename T>
void bidiagonal(lfb_handle &h, T *A, int64_t rows, int64_t cols, int64_t ld, T *d, T *e) {
    const int64_t md = std::min(rows, cols);
    if (md <= 0) return;
    if (!h.opt.bd_blocked) {
        bidiagonal_unblocked<T>(h, A, rows, cols, ld, d, e);
        return;
    }
    if (rows >= cols) {
        if (!bidiagonal_tall<T>(h, A, rows, cols, ld, d, e)) bidiagonal_unblocked<T>(h, A, rows, cols, ld, d, e);
        return;
    }
    // rows < cols: clear_row == clear_column on the transpose (householder.rs:57-63)
    const int64_t ldt = round_up(cols, 2);
    DevBuf<T> At(h, (size_t)ldt * rows);
    transpose<T>(h, A, rows, cols, ld, At.get(), ldt);
    if (bidiagonal_tall<T>(h, At.get(), cols, rows, ldt, d, e)) transpose<T>(h, At.get(), cols, rows, ldt, A, ld);
    else bidiagonal_unblocked<T>(h, A, rows, cols, ld, d, e);
}

template void bidiagonal<float>(lfb_handle &, float *, int64_t, int64_t, int64_t, float *, float *);
template void bidiagonal<double>(lfb_handle &, double *, int64_t, int64_t, int64_t, double *, double *);

}  // namespace lfb
