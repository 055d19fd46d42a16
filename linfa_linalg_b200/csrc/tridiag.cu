// Blocked symmetric tridiagonalisation (the dsytrd / dlatrd structure, reference sign conventions).
//
// Replaces src/tridiagonal.rs:31-66.  The reference performs, per column, p = 2 M v (GEMV) and three
// rank-1 GEMMs on the full trailing matrix M (24 bytes of HBM traffic per matrix element per column).
// Here the rank-2 updates of a panel of 32 reflectors are deferred:
//     M_j = A - V W^T - W V^T  (V, W: the panel's reflectors and their companions w = p - (v.p) v)
// so a column needs only ONE pass over the stored trailing matrix (the GEMV, 8 B/element, HBM bound)
// plus skinny row-parallel corrections, and the trailing matrix is updated once per panel by two
// tensor-core GEMMs (K = 32).  Four multi-CTA launches per column (grid-wide scalars travel through
// small atomics accumulators; a first version with single-CTA helper kernels spent 100 us per
// column streaming V and W through one SM):
//   trd_update   bring column i up to date with the panel so far (and finalise the previous
//                companion w_{j-1} on the fly), accumulate ||x||^2
//   trd_reflect  make the unit-norm reflector (householder.rs:9-28), off[i], store v, zero the GEMV
//                target, accumulate W^T v and V^T v
//   trd_gemv     p = 2 A22 v over the stored (stale) trailing matrix
//   trd_correct  p -= 2 V (W^T v) + 2 W (V^T v), accumulate v . p
// The algebra is identical to the reference's (also for non-symmetric input, where both compute the
// same "wrong" thing, tests/tridiagonal.rs:36-42); only the lower triangle + diagonal and `off` are
// observable (tridiagonal.rs:90-113).
//
// Second generation (default, options trd_fused / trd_symv): two launches per column.
//   trd_head  ONE thread-block cluster (<= 16 CTAs): correction + v.p of the previous column, its
//             companion w, the column update, the reflector, W^T v / V^T v -- the grid-wide scalars
//             are reduced through distributed shared memory and cluster barriers instead of kernel
//             boundaries (4 cluster barriers replace 3 launches + 1 launch of BLAS-1 work)
//   trd_symv  p = 2 A22 v reading ONLY the lower triangle: every 64x32 register tile feeds both
//             y[rows] += A x[cols] and y[cols] += A^T x[rows]; 4 B/element instead of 8, 16-byte
//             streaming loads, column sums reduced with a 31-shuffle transpose-reduce.  The trailing
//             update then also touches the lower triangle only (half the GEMM work).
#include <cooperative_groups.h>

#include "common.cuh"
#include "dev_utils.cuh"

namespace cg = cooperative_groups;

namespace lfb {
namespace {

using namespace dev;

constexpr int TB = 32;     // reflectors per panel
constexpr int NTH = 256;   // threads per CTA
constexpr int RPB = 512;   // rows per CTA

template <typename T>
struct TrdAcc {
    T nsq[2];
    T delta[2];
    T head[2];
    T tt[2][2 * TB];
    int some;
};

template <typename T>
__device__ __forceinline__ T block_sum(T v, T *sred) {  // blockDim.x == NTH
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) sred[warp] = v;
    __syncthreads();
    T s = T(0);
#pragma unroll
    for (int w = 0; w < NTH / 32; ++w) s += sred[w];
    return s;
}

// V, Wm: n x TB (ld = n), indexed by GLOBAL row.  P: p_true of the previous column (global rows).
template <typename T>
__global__ void __launch_bounds__(NTH) trd_update_kernel(T *A, int64_t ld, int64_t n, int64_t i, int j, const T *V, T *Wm,
                                                         const T *P, TrdAcc<T> *acc) {
    __shared__ T sred[NTH / 32];
    __shared__ T cw[TB], cv[TB];   // Wm[i, k], V[i, k]
    const int par = j & 1;
    const T dprev = j > 0 ? acc->delta[par ^ 1] : T(0);
    if (threadIdx.x < j) {
        const int k = threadIdx.x;
        cv[k] = V[i + (int64_t)k * n];
        cw[k] = (k == j - 1) ? P[i] - dprev * V[i + (int64_t)k * n] : Wm[i + (int64_t)k * n];
    }
    __syncthreads();
    T *col = A + i * ld;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        T d = col[i];                                   // diagonal element d_i
        for (int k = 0; k < j; ++k) d -= T(2) * cv[k] * cw[k];
        col[i] = d;
        acc->delta[par] = T(0);
    }
    T part = T(0);
    for (int64_t r = i + 1 + (int64_t)blockIdx.x * NTH + threadIdx.x; r < n; r += (int64_t)gridDim.x * NTH) {
        T x = col[r];
        if (j > 0) {
            const T vprev = V[r + (int64_t)(j - 1) * n];
            const T wprev = P[r] - dprev * vprev;       // finalise w_{j-1} (tridiagonal.rs:56-58 folded)
            Wm[r + (int64_t)(j - 1) * n] = wprev;
            for (int k = 0; k < j - 1; ++k) x -= V[r + (int64_t)k * n] * cw[k] + Wm[r + (int64_t)k * n] * cv[k];
            x -= vprev * cw[j - 1] + wprev * cv[j - 1];
        }
        col[r] = x;
        if (r == i + 1) acc->head[par] = x;
        part += x * x;
    }
    const T s = block_sum(part, sred);
    if (threadIdx.x == 0) atomicAdd(&acc->nsq[par], s);
}

template <typename T>
__global__ void __launch_bounds__(NTH) trd_reflect_kernel(T *A, int64_t ld, int64_t n, int64_t i, int j, T *V, const T *Wm, T *P,
                                                          T *off, TrdAcc<T> *acc) {
    __shared__ T sv[RPB];
    const int par = j & 1;
    const T nsq = acc->nsq[par], f = acc->head[par];
    const T nrm = t_sqrt(nsq);                       // householder.rs:13
    const T s = t_signum(f) * nrm;                   // :16
    const T newsq = (nsq + t_abs(f) * nrm) * T(2);   // :19-20
    const bool some = newsq != T(0);                 // :22
    const T d = t_sqrt(newsq);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        off[i] = some ? -s : T(0);                   // tridiagonal.rs:45
        acc->some = some ? 1 : 0;
    }
    T *col = A + i * ld;
    T *vout = V + (int64_t)j * n;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t r0 = i + 1 + (int64_t)blockIdx.x * RPB; r0 < n; r0 += (int64_t)gridDim.x * RPB) {
        const int nr = (int)min((int64_t)RPB, n - r0);
        for (int t = threadIdx.x; t < nr; t += NTH) {
            const int64_t r = r0 + t;
            T x = col[r];
            if (some) {
                x = ((r == i + 1) ? x + s : x) / d;  // :17,23
                col[r] = x;
            }
            const T v = some ? x : T(0);
            vout[r] = v;
            P[r] = T(0);
            sv[t] = v;
        }
        __syncthreads();
        // W_k . v and V_k . v over this row block, one warp per k
        if (some) {
            for (int k = warp; k < j; k += NTH / 32) {
                const T *wk = Wm + (int64_t)k * n + r0, *vk = V + (int64_t)k * n + r0;
                T s1 = T(0), s2 = T(0);
                for (int t = lane; t < nr; t += 32) {
                    s1 += wk[t] * sv[t];
                    s2 += vk[t] * sv[t];
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
                    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
                }
                if (lane == 0) {
                    atomicAdd(&acc->tt[par][k], s1);
                    atomicAdd(&acc->tt[par][TB + k], s2);
                }
            }
        }
        __syncthreads();
    }
}

// p[r] += alpha * sum_{c in split} M[r,c] x[c]   (M rows x cols column-major; thread per row)
template <typename T>
__global__ void __launch_bounds__(128) trd_gemv_kernel(const T *__restrict__ M, int64_t ld, int64_t rows, int64_t cols,
                                                       const T *__restrict__ x, T alpha, T *y, int64_t csplit, const TrdAcc<T> *acc) {
    if (acc->some == 0) return;
    const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t c0 = blockIdx.y * csplit, c1 = min(cols, c0 + csplit);
    if (r >= rows) return;
    T a0 = T(0), a1 = T(0), a2 = T(0), a3 = T(0), a4 = T(0), a5 = T(0), a6 = T(0), a7 = T(0);
    const T *m = M + r;
    int64_t c = c0;
    for (; c + 7 < c1; c += 8) {   // 8 independent loads in flight per thread
        a0 += m[c * ld] * x[c];
        a1 += m[(c + 1) * ld] * x[c + 1];
        a2 += m[(c + 2) * ld] * x[c + 2];
        a3 += m[(c + 3) * ld] * x[c + 3];
        a4 += m[(c + 4) * ld] * x[c + 4];
        a5 += m[(c + 5) * ld] * x[c + 5];
        a6 += m[(c + 6) * ld] * x[c + 6];
        a7 += m[(c + 7) * ld] * x[c + 7];
    }
    for (; c < c1; ++c) a0 += m[c * ld] * x[c];
    atomicAdd(y + r, alpha * (((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7))));
}

template <typename T>
__global__ void __launch_bounds__(NTH) trd_correct_kernel(int64_t n, int64_t i, int j, const T *V, const T *Wm, T *P,
                                                          TrdAcc<T> *acc) {
    __shared__ T sred[NTH / 32];
    __shared__ T t1[TB], t2[TB];
    const int par = j & 1;
    if (threadIdx.x < j) {
        t1[threadIdx.x] = acc->tt[par][threadIdx.x];
        t2[threadIdx.x] = acc->tt[par][TB + threadIdx.x];
    }
    __syncthreads();
    if (blockIdx.x == 0) {   // accumulators of the next column
        if (threadIdx.x == 0) acc->nsq[par ^ 1] = T(0);
        if (threadIdx.x < 2 * TB) acc->tt[par ^ 1][threadIdx.x] = T(0);
    }
    const T *v = V + (int64_t)j * n;
    T part = T(0);
    for (int64_t r = i + 1 + (int64_t)blockIdx.x * NTH + threadIdx.x; r < n; r += (int64_t)gridDim.x * NTH) {
        T pr = P[r];
        for (int k = 0; k < j; ++k) pr -= T(2) * (V[r + (int64_t)k * n] * t1[k] + Wm[r + (int64_t)k * n] * t2[k]);
        P[r] = pr;
        part += v[r] * pr;                            // tridiagonal.rs:50
    }
    const T s = block_sum(part, sred);
    if (threadIdx.x == 0) atomicAdd(&acc->delta[par], s);
}

// End of a panel: w of the last column, W[:, j] = P - delta v.
template <typename T>
__global__ void __launch_bounds__(NTH) trd_finalize_kernel(int64_t n, int64_t i, int j, const T *V, T *Wm, const T *P,
                                                           const TrdAcc<T> *acc) {
    const T dl = acc->delta[j & 1];
    for (int64_t r = i + 1 + (int64_t)blockIdx.x * NTH + threadIdx.x; r < n; r += (int64_t)gridDim.x * NTH)
        Wm[r + (int64_t)j * n] = P[r] - dl * V[r + (int64_t)j * n];
}


// ================================ second generation =============================================
constexpr int SW = 32;     // columns per SYMV strip (= column accumulators per lane)
constexpr int HNT = 512;   // threads per CTA of the head kernel

// y[r] += alpha * sum_c S[r,c] x[c] for the symmetric S whose lower triangle is stored in global rows/cols
// [i1, n) of A.  x and y are indexed by GLOBAL row.  CTA = (32-column strip, `chunk`-row block); a warp
// takes 64-row groups (2 rows per lane, one 16-byte load per column).
template <typename T, bool ALIGNED>
__global__ void __launch_bounds__(256, 2) trd_symv_kernel(const T *__restrict__ A, int64_t ld, int64_t n, int64_t i1,
                                                         const T *__restrict__ x, T alpha, T *y, int chunk,
                                                         const TrdAcc<T> *acc) {
    if (acc->some == 0) return;
    using V2 = typename Vec2<T>::type;
    __shared__ T sx[SW];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int64_t C0 = i1 + (int64_t)blockIdx.x * SW;
    const int64_t R0 = (i1 & ~(int64_t)1) + (int64_t)blockIdx.y * chunk;
    if (R0 >= n || R0 + chunk <= C0) return;          // block entirely above the diagonal (CTA-uniform)
    if (threadIdx.x < SW) sx[threadIdx.x] = (C0 + threadIdx.x < n) ? x[C0 + threadIdx.x] : T(0);
    __syncthreads();
    T col[SW];
#pragma unroll
    for (int k = 0; k < SW; ++k) col[k] = T(0);
    bool any = false;
    for (int g = warp; g < chunk / 64; g += nw) {
        const int64_t g0 = R0 + 64 * (int64_t)g;
        if (g0 >= n) break;
        if (g0 + 64 <= C0) continue;
        any = true;
        const int64_t gr = g0 + 2 * lane;
        T r0 = T(0), r1 = T(0);
        const bool fast = ALIGNED && g0 > C0 + (SW - 1) && g0 + 64 <= n;   // interior group: no masks, vector loads
        const bool v0 = gr >= i1 && gr < n, v1 = gr + 1 >= i1 && gr + 1 < n;
        T x0 = T(0), x1 = T(0);
        if (fast) {
            const V2 xv = *reinterpret_cast<const V2 *>(x + gr);
            x0 = xv.x; x1 = xv.y;
        } else {
            if (v0) x0 = x[gr];
            if (v1) x1 = x[gr + 1];
        }
        const T *p = A + gr + C0 * ld;
#pragma unroll
        for (int cb = 0; cb < SW; cb += 8) {
            V2 a[8];
            if (fast) {
#pragma unroll
                for (int q = 0; q < 8; ++q) a[q] = __ldcs(reinterpret_cast<const V2 *>(p + (int64_t)(cb + q) * ld));
            } else {                                                   // diagonal / edge group: masked scalar loads
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const int64_t gc = C0 + cb + q;
                    a[q].x = (v0 && gr >= gc && gc < n) ? p[(int64_t)(cb + q) * ld] : T(0);
                    a[q].y = (v1 && gr + 1 >= gc && gc < n) ? p[(int64_t)(cb + q) * ld + 1] : T(0);
                }
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int64_t gc = C0 + cb + q;
                const T xs = sx[cb + q];
                r0 += a[q].x * xs;
                r1 += a[q].y * xs;
                const T c0 = (!fast && gr == gc) ? T(0) : a[q].x;      // the diagonal feeds the row sum only
                const T c1 = (!fast && gr + 1 == gc) ? T(0) : a[q].y;
                col[cb + q] += c0 * x0 + c1 * x1;
            }
        }
        if (gr >= i1 && gr < n) atomicAdd(y + gr, alpha * r0);
        if (gr + 1 >= i1 && gr + 1 < n) atomicAdd(y + gr + 1, alpha * r1);
    }
    if (any) {
        warp_transpose_reduce<T>(col, lane);
        if (C0 + lane < n) atomicAdd(y + C0 + lane, alpha * col[0]);
    }
}

// Same tiling as trd_symv_kernel, but an interior 64x32 tile is first landed in shared memory with
// cp.async (each lane fetches and later reads back only its own 16 bytes per column, so no barrier
// is needed): a warp has its whole 16 KB tile in flight without holding it in registers, and three
// 4-warp CTAs per SM keep ~190 KB per SM outstanding -- what the HBM latency-bandwidth product asks for.
template <typename T>
__global__ void __launch_bounds__(128, 3) trd_symv_async_kernel(const T *__restrict__ A, int64_t ld, int64_t n, int64_t i1,
                                                             const T *__restrict__ x, T alpha, T *y, int chunk,
                                                             const TrdAcc<T> *acc) {
    if (acc->some == 0) return;
    using V2 = typename Vec2<T>::type;
    extern __shared__ __align__(16) unsigned char symv_smem[];
    __shared__ T sx[SW];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int64_t C0 = i1 + (int64_t)blockIdx.x * SW;
    const int64_t R0 = (i1 & ~(int64_t)1) + (int64_t)blockIdx.y * chunk;
    if (R0 >= n || R0 + chunk <= C0) return;
    if (threadIdx.x < SW) sx[threadIdx.x] = (C0 + threadIdx.x < n) ? x[C0 + threadIdx.x] : T(0);
    __syncthreads();
    V2 *buf = reinterpret_cast<V2 *>(symv_smem) + (size_t)warp * (SW * 32) + lane;
    T col[SW];
#pragma unroll
    for (int k = 0; k < SW; ++k) col[k] = T(0);
    bool any = false;
    for (int g = warp; g < chunk / 64; g += nw) {
        const int64_t g0 = R0 + 64 * (int64_t)g;
        if (g0 >= n) break;
        if (g0 + 64 <= C0) continue;
        any = true;
        const int64_t gr = g0 + 2 * lane;
        T r0 = T(0), r1 = T(0);
        const T *p = A + gr + C0 * ld;
        if (g0 > C0 + (SW - 1) && g0 + 64 <= n) {                      // interior tile
#pragma unroll
            for (int cb = 0; cb < SW; cb += 8) {
#pragma unroll
                for (int q = 0; q < 8; ++q) cp_async<(int)sizeof(V2)>(buf + (cb + q) * 32, p + (int64_t)(cb + q) * ld);
                cp_async_commit();
            }
            const V2 xv = *reinterpret_cast<const V2 *>(x + gr);
#define LFB_SYMV_STEP(CB, PENDING)                                                          \
            cp_async_wait<PENDING>();                                                           \
            _Pragma("unroll") for (int q = 0; q < 8; ++q) {                                      \
                const V2 a = buf[((CB) + q) * 32];                                               \
                const T xs = sx[(CB) + q];                                                       \
                r0 += a.x * xs;                                                                  \
                r1 += a.y * xs;                                                                  \
                col[(CB) + q] += a.x * xv.x + a.y * xv.y;                                        \
            }
            LFB_SYMV_STEP(0, 3)
            LFB_SYMV_STEP(8, 2)
            LFB_SYMV_STEP(16, 1)
            LFB_SYMV_STEP(24, 0)
#undef LFB_SYMV_STEP
        } else {                                                       // diagonal / edge tile: masked direct loads
            const bool v0 = gr >= i1 && gr < n, v1 = gr + 1 >= i1 && gr + 1 < n;
            const T x0 = v0 ? x[gr] : T(0), x1 = v1 ? x[gr + 1] : T(0);
#pragma unroll
            for (int cb = 0; cb < SW; cb += 8) {
                T a0[8], a1[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const int64_t gc = C0 + cb + q;
                    a0[q] = (v0 && gr >= gc && gc < n) ? p[(int64_t)(cb + q) * ld] : T(0);
                    a1[q] = (v1 && gr + 1 >= gc && gc < n) ? p[(int64_t)(cb + q) * ld + 1] : T(0);
                }
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const int64_t gc = C0 + cb + q;
                    const T xs = sx[cb + q];
                    r0 += a0[q] * xs;
                    r1 += a1[q] * xs;
                    col[cb + q] += (gr == gc ? T(0) : a0[q]) * x0 + (gr + 1 == gc ? T(0) : a1[q]) * x1;   // diagonal: row sum only
                }
            }
        }
        if (gr >= i1 && gr < n) atomicAdd(y + gr, alpha * r0);
        if (gr + 1 >= i1 && gr + 1 < n) atomicAdd(y + gr + 1, alpha * r1);
    }
    if (any) {
        warp_transpose_reduce<T>(col, lane);
        if (C0 + lane < n) atomicAdd(y + C0 + lane, alpha * col[0]);
    }
}

// Everything between two SYMVs, in one cluster launch (see the header).  Column i = i0 + j of the
// panel; rows r >= i are dealt to the cluster's threads (r = i + rank * HNT + tid + m * nc * HNT, the
// same owner in every phase).  V, Wm: ldv x TB, indexed by GLOBAL row.  P: SYMV result of the previous
// column (p before the panel correction), P2: scratch.  All loops over the panel's columns are
// unrolled with predicated loads so that a phase costs one or two memory latencies, not j of them.
template <typename T>
__global__ void __launch_bounds__(HNT) trd_head_kernel(T *A, int64_t ld, int64_t n, int64_t i, int j, T *V, T *Wm, int64_t ldv,
                                                       T *P, T *P2, T *off, TrdAcc<T> *acc, int final_only) {
    cg::cluster_group cl = cg::this_cluster();
    const int nc = (int)cl.num_blocks(), b = (int)cl.block_rank();
    __shared__ T slotA[2], slotB[2], resA[2], resB[2], spi;
    __shared__ T sred[HNT / 32][2];
    __shared__ T stt[2 * TB];
    __shared__ T cw[TB], cv[TB], t1[TB], t2[TB];
    __shared__ T sv[HNT];
    __shared__ T inbox[16][2 * TB];
    const int par = j & 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t stride = (int64_t)nc * HNT;
    const int64_t first = i + (int64_t)b * HNT + threadIdx.x;
    const int jp = j - 1;
    const bool single = first + stride >= n;          // this thread owns at most one row: it stays in registers across the phases
    T keep_p = T(0), keep_v = T(0), keep_c = T(0), keep_x = T(0);
    T *col = A + i * ld;
    const T *vp = V + (int64_t)(jp > 0 ? jp : 0) * ldv;
    T *wp = Wm + (int64_t)(jp > 0 ? jp : 0) * ldv;
    if (threadIdx.x < TB) { cw[threadIdx.x] = T(0); cv[threadIdx.x] = T(0); t1[threadIdx.x] = T(0); t2[threadIdx.x] = T(0); }
    if (threadIdx.x < 2 * TB) stt[threadIdx.x] = T(0);
    __syncthreads();
    T dprev = T(0);
    if (j > 0) {
        if (threadIdx.x < jp) {
            t1[threadIdx.x] = acc->tt[par ^ 1][threadIdx.x];
            t2[threadIdx.x] = acc->tt[par ^ 1][TB + threadIdx.x];
            cw[threadIdx.x] = Wm[i + (int64_t)threadIdx.x * ldv];       // row i of W (columns < j-1 are final)
        }
        if (threadIdx.x < j) cv[threadIdx.x] = V[i + (int64_t)threadIdx.x * ldv];
        __syncthreads();
        if (warp == 0) {     // corrected p_{j-1}[i]; W[i, j-1] follows once delta is known (every CTA needs it)
            T term = cv[lane] * t1[lane] + cw[lane] * t2[lane];     // zero padded beyond j-1
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) term += __shfl_xor_sync(0xffffffffu, term, o);
            if (lane == 0) spi = P[i] - T(2) * term;
        }
        // ---- one pass over rows of V, W: correction of p_{j-1} (tridiagonal.rs:49 on the deferred matrix),
        //      delta = v . p (:50), and the part of the column update that does not depend on delta ----
        T part = T(0);
        for (int64_t r = first; r < n; r += stride) {
            T pr = P[r], S = T(0);
            const T vjp_ = vp[r];                         // issued with the panel rows, used after them
            for (int kb = 0; kb < jp; kb += 16) {
                T vv[16], ww[16];
#pragma unroll
                for (int q = 0; q < 16; ++q) {
                    const bool ok = kb + q < jp;
                    vv[q] = ok ? V[r + (int64_t)(kb + q) * ldv] : T(0);
                    ww[q] = ok ? Wm[r + (int64_t)(kb + q) * ldv] : T(0);
                }
#pragma unroll
                for (int q = 0; q < 16; ++q) {
                    pr -= T(2) * (vv[q] * t1[kb + q] + ww[q] * t2[kb + q]);
                    S += vv[q] * cw[kb + q] + ww[q] * cv[kb + q];
                }
            }
            T cx = T(0);
            if (!final_only) { cx = col[r] - S; if (!single) col[r] = cx; }
            if (!single) P2[r] = pr;
            keep_p = pr; keep_v = vjp_; keep_c = cx;      // a thread that owns one row keeps it in registers
            part += vjp_ * pr;
        }
        T dummy;
        cluster_sum2<T, HNT>(cl, part, T(0), slotA, sred, resA, dprev, dummy);
        if (threadIdx.x == 0) cw[jp] = spi - dprev * vp[i];
        __syncthreads();
    }
    // ---- companion w_{j-1} = p - delta v (tridiagonal.rs:56-58 folded into one vector), rest of the column
    //      update, ||x||^2 and the head element.  Row i itself yields the diagonal entry d_i. ----
    T part = T(0), headv = T(0);
    for (int64_t r = first; r < n; r += stride) {
        T xx = T(0);
        if (j > 0) {
            const T vjp = single ? keep_v : vp[r];
            const T w = (single ? keep_p : P2[r]) - dprev * vjp;
            wp[r] = w;
            if (!final_only) {
                xx = (single ? keep_c : col[r]) - (vjp * cw[jp] + w * cv[jp]);
                if (!single || r == i) col[r] = xx;       // row i is the diagonal entry d_i; the rest is rewritten below
            }
        } else {
            xx = col[r];
        }
        keep_x = xx;
        if (r > i) {
            part += xx * xx;
            if (r == i + 1) headv = xx;
        }
    }
    if (final_only) {
        cl.sync();    // nobody leaves while its mailbox may still be read
        return;
    }
    T nsq, f;
    cluster_sum2<T, HNT>(cl, part, headv, slotB, sred, resB, nsq, f);
    // ---- householder.rs:9-28 ----
    const T nrm = t_sqrt(nsq);
    const T s = t_signum(f) * nrm;
    const T newsq = (nsq + t_abs(f) * nrm) * T(2);
    const bool some = newsq != T(0);
    const T d = t_sqrt(newsq);
    if (b == 0 && threadIdx.x == 0) {
        off[i] = some ? -s : T(0);                   // tridiagonal.rs:45
        acc->some = some ? 1 : 0;
    }
    T *vout = V + (int64_t)j * ldv;
    // ---- the reflector, and W^T v / V^T v: the CTA's rows of one pass are a contiguous block of HNT rows; v is staged
    //      in shared memory and warp k % 16 takes panel column k with 2 x 16 independent coalesced loads per lane ----
    for (int64_t base = i + (int64_t)b * HNT; base < n; base += stride) {
        const int64_t r = base + threadIdx.x;
        T v = T(0);
        if (r < n && r > i) {
            const T xx = single ? keep_x : col[r];
            if (some) v = ((r == i + 1) ? xx + s : xx) / d;
            col[r] = some ? v : xx;
            vout[r] = v;
            P[r] = T(0);                                 // target of the SYMV that follows
        }
        sv[threadIdx.x] = v;
        __syncthreads();
        if (some) {
            for (int k = warp; k < j; k += HNT / 32) {
                const T *wk = Wm + (int64_t)k * ldv + base, *vk = V + (int64_t)k * ldv + base;
                T lw[HNT / 32], lv[HNT / 32];
#pragma unroll
                for (int q = 0; q < HNT / 32; ++q) {
                    const int t = lane + 32 * q;
                    const bool ok = base + t < n;
                    lw[q] = ok ? wk[t] : T(0);
                    lv[q] = ok ? vk[t] : T(0);
                }
                T s1 = T(0), s2 = T(0);
#pragma unroll
                for (int q = 0; q < HNT / 32; ++q) {
                    s1 += lw[q] * sv[lane + 32 * q];
                    s2 += lv[q] * sv[lane + 32 * q];
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
                    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
                }
                if (lane == 0) { stt[k] += s1; stt[TB + k] += s2; }   // warp k % 16 owns entry k
            }
        }
        __syncthreads();
    }
    // push the per-CTA partials into rank 0's inbox (parallel DSMEM stores), one barrier, local sum
    if (threadIdx.x < 2 * TB) cl.map_shared_rank(&inbox[0][0], 0)[b * 2 * TB + threadIdx.x] = stt[threadIdx.x];
    cl.sync();
    if (b == 0 && threadIdx.x < 2 * TB) {
        T sum = T(0);
        for (int rk = 0; rk < nc; ++rk) sum += inbox[rk][threadIdx.x];
        acc->tt[par][threadIdx.x] = sum;
    }
}

template <typename T>
bool launch_head(lfb_handle &h, int nc, T *A, int64_t ld, int64_t n, int64_t i, int j, T *V, T *Wm, int64_t ldv, T *P, T *P2,
                 T *off, TrdAcc<T> *acc, int final_only) {
    static DeviceOnce cfg;   // function attributes are per device
    cfg.run(h.device, [&] {
        cudaFuncSetAttribute(trd_head_kernel<T>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        cudaGetLastError();
    });
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
    cudaError_t e = cudaLaunchKernelEx(&c, trd_head_kernel<T>, A, ld, n, i, j, V, Wm, ldv, P, P2, off, acc, final_only);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    h.launches++;
    return true;
}

template <typename T>
void launch_symv(lfb_handle &h, const T *A, int64_t ld, int64_t n, int64_t i1, const T *x, T alpha, T *y, const TrdAcc<T> *acc) {
    const int64_t L = n - i1;
    if (L <= 0) return;
    const int64_t strips = cdiv(L, SW);
    const int64_t gbase = i1 & ~(int64_t)1;
    const bool aligned = ((uintptr_t)A % (2 * sizeof(T)) == 0) && (ld % 2 == 0) && ((uintptr_t)x % (2 * sizeof(T)) == 0);
    if (aligned && h.opt.trd_symv_async) {
        using V2 = typename Vec2<T>::type;
        static DeviceOnce cfg;   // function attributes are per device
        cfg.run(h.device, [&] {
            LFB_CUDA(cudaFuncSetAttribute(trd_symv_async_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)(4 * SW * 32 * sizeof(V2))));
        });
        int chunk = 256;                                  // one 64-row tile per warp unless the matrix is small
        while (chunk > 64 && strips * cdiv(L, chunk) / 2 < 4 * 3 * (int64_t)h.sm_count) chunk >>= 1;
        const int nw = std::min(4, chunk / 64);
        dim3 grid((unsigned)strips, (unsigned)cdiv(n - gbase, chunk));
        trd_symv_async_kernel<T><<<grid, 32 * nw, (size_t)nw * SW * 32 * sizeof(V2), h.stream>>>(A, ld, n, i1, x, alpha, y, chunk, acc);
        LFB_LAUNCH_CHECK(h);
        return;
    }
    int chunk = 1024;                                     // largest row block that still gives >= 6 waves of CTAs
    while (chunk > 64 && strips * cdiv(L, chunk) / 2 < 6 * 2 * (int64_t)h.sm_count) chunk >>= 1;
    const int nw = std::min(8, chunk / 64);
    dim3 grid((unsigned)strips, (unsigned)cdiv(n - gbase, chunk));
    if (aligned) trd_symv_kernel<T, true><<<grid, 32 * nw, 0, h.stream>>>(A, ld, n, i1, x, alpha, y, chunk, acc);
    else trd_symv_kernel<T, false><<<grid, 32 * nw, 0, h.stream>>>(A, ld, n, i1, x, alpha, y, chunk, acc);
    LFB_LAUNCH_CHECK(h);
}

inline int head_cluster_size(const lfb_handle &h, int64_t L) {
    int nc = 1;
    while (nc < 16 && (int64_t)nc * HNT < L) nc <<= 1;
    const int cap = (int)std::max<int64_t>(1, std::min<int64_t>(16, h.opt.panel_cluster_max));
    while (nc > cap) nc >>= 1;
    return nc;
}

}  // namespace

template <typename T>
static void sym_tridiagonal_v1(lfb_handle &h, T *A, int64_t n, int64_t ld, T *off) {
    if (n <= 1) return;
    DevBuf<T> V(h, (size_t)n * TB), Wm(h, (size_t)n * TB), P(h, n);
    DevBuf<TrdAcc<T>> acc(h, 1);
    auto nblk = [&](int64_t L, int per) { return (unsigned)std::max<int64_t>(1, std::min<int64_t>(cdiv(L, per), 2 * h.sm_count)); };
    for (int64_t i0 = 0; i0 < n - 1; i0 += TB) {
        const int pb = (int)std::min<int64_t>(TB, n - 1 - i0);
        LFB_CUDA(cudaMemsetAsync(acc.get(), 0, sizeof(TrdAcc<T>), h.stream));
        for (int j = 0; j < pb; ++j) {
            const int64_t i = i0 + j, L = n - i - 1;
            trd_update_kernel<T><<<nblk(L, NTH), NTH, 0, h.stream>>>(A, ld, n, i, j, V.get(), Wm.get(), P.get(), acc.get());
            LFB_LAUNCH_CHECK(h);
            trd_reflect_kernel<T><<<nblk(L, RPB), NTH, 0, h.stream>>>(A, ld, n, i, j, V.get(), Wm.get(), P.get(), off, acc.get());
            LFB_LAUNCH_CHECK(h);
            {
                const T *M = A + (i + 1) + (i + 1) * ld;
                const int64_t rb = cdiv(L, 128);
                // ncu (profiles/r1_tridiag.md): 630 CTAs = 26 % occupancy reached only 2.6 TB/s; 16 CTAs of 128
                // threads per SM keep enough loads in flight for the HBM latency
                const int64_t splits = std::max<int64_t>(1, std::min<int64_t>(cdiv(16 * h.sm_count, rb), cdiv(L, 64)));
                const int64_t csplit = cdiv(L, splits);
                dim3 grid((unsigned)rb, (unsigned)cdiv(L, csplit));
                trd_gemv_kernel<T><<<grid, 128, 0, h.stream>>>(M, ld, L, L, V.get() + (int64_t)j * n + (i + 1), T(2), P.get() + (i + 1), csplit, acc.get());
                LFB_LAUNCH_CHECK(h);
            }
            trd_correct_kernel<T><<<nblk(L, NTH), NTH, 0, h.stream>>>(n, i, j, V.get(), Wm.get(), P.get(), acc.get());
            LFB_LAUNCH_CHECK(h);
        }
        {
            const int j = pb - 1;
            const int64_t i = i0 + j;
            trd_finalize_kernel<T><<<nblk(n - i - 1, NTH), NTH, 0, h.stream>>>(n, i, j, V.get(), Wm.get(), P.get(), acc.get());
            LFB_LAUNCH_CHECK(h);
        }
        // trailing update (both triangles: later GEMVs read full rows): rows/cols >= i0 + pb
        const int64_t R0 = i0 + pb, Lt = n - R0;
        if (Lt > 0) {
            T *C = A + R0 + R0 * ld;
            gemm<T>(h, 0, 1, Lt, Lt, pb, T(-1), V.get() + R0, n, Wm.get() + R0, n, T(1), C, ld);
            gemm<T>(h, 0, 1, Lt, Lt, pb, T(-1), Wm.get() + R0, n, V.get() + R0, n, T(1), C, ld);
        }
    }
}


// Second-generation driver: trd_head (cluster) + trd_symv per column, lower-triangle trailing update.
// Returns false (nothing modified) if the cluster kernel cannot be launched on this device.
template <typename T>
static bool sym_tridiagonal_v2(lfb_handle &h, T *A, int64_t n, int64_t ld, T *off) {
    const int64_t ldv = round_up(n, 2);
    DevBuf<T> V(h, (size_t)ldv * TB), Wm(h, (size_t)ldv * TB), P(h, ldv), P2(h, ldv);
    DevBuf<TrdAcc<T>> acc(h, 1);
    LFB_CUDA(cudaMemsetAsync(acc.get(), 0, sizeof(TrdAcc<T>), h.stream));
    // option trd_profile: CUDA events around every launch, summary on stderr (debug / profiles/ only)
    const bool prof = h.opt.trd_profile != 0;
    std::vector<cudaEvent_t> ev;
    auto mark = [&]() {
        if (!prof) return;
        cudaEvent_t e;
        LFB_CUDA(cudaEventCreate(&e));
        LFB_CUDA(cudaEventRecord(e, h.stream));
        ev.push_back(e);
    };
    bool first = true;
    for (int64_t i0 = 0; i0 < n - 1; i0 += TB) {
        const int pb = (int)std::min<int64_t>(TB, n - 1 - i0);
        for (int j = 0; j < pb; ++j) {
            const int64_t i = i0 + j;
            mark();
            const bool ok = launch_head<T>(h, head_cluster_size(h, n - i), A, ld, n, i, j, V.get(), Wm.get(), ldv, P.get(), P2.get(),
                                           off, acc.get(), 0);
            if (!ok) {
                if (first) return false;
                throw CudaError(LFB_ERR_CUDA, "trd_head_kernel launch failed");
            }
            first = false;
            mark();
            launch_symv<T>(h, A, ld, n, i + 1, V.get() + (int64_t)j * ldv, T(2), P.get(), acc.get());
            mark();
        }
        const int64_t R0 = i0 + pb, Lt = n - R0;
        if (Lt > 0) {
            if (!launch_head<T>(h, head_cluster_size(h, Lt), A, ld, n, R0, pb, V.get(), Wm.get(), ldv, P.get(), P2.get(), off,
                                acc.get(), 1))
                throw CudaError(LFB_ERR_CUDA, "trd_head_kernel launch failed");
            T *C = A + R0 + R0 * ld;     // only the lower triangle is read from here on
            gemm<T>(h, 0, 1, Lt, Lt, pb, T(-1), V.get() + R0, ldv, Wm.get() + R0, ldv, T(1), C, ld, 1);
            gemm<T>(h, 0, 1, Lt, Lt, pb, T(-1), Wm.get() + R0, ldv, V.get() + R0, ldv, T(1), C, ld, 1);
        }
    }
    if (prof) {
        mark();
        LFB_CUDA(cudaStreamSynchronize(h.stream));
        double th = 0, ts = 0, tg = 0;
        float t;
        for (size_t k = 0; k + 3 < ev.size() + 1 && k + 2 < ev.size(); k += 3) {
            cudaEventElapsedTime(&t, ev[k], ev[k + 1]); th += t;
            cudaEventElapsedTime(&t, ev[k + 1], ev[k + 2]); ts += t;
            if (k + 3 < ev.size()) { cudaEventElapsedTime(&t, ev[k + 2], ev[k + 3]); tg += t; }
        }
        cudaEventElapsedTime(&t, ev.front(), ev.back());
        fprintf(stderr, "[trd_profile] n=%lld total %.2f ms: head %.2f ms, symv %.2f ms, between columns (panel end: final head + 2 GEMMs) %.2f ms\n",
                (long long)n, t, th, ts, tg);
        for (auto e : ev) cudaEventDestroy(e);
    }
    return true;
}

template <typename T>
void sym_tridiagonal(lfb_handle &h, T *A, int64_t n, int64_t ld, T *off) {
    if (n <= 1) return;
    if (h.opt.trd_fused && sym_tridiagonal_v2<T>(h, A, n, ld, off)) return;
    sym_tridiagonal_v1<T>(h, A, n, ld, off);
}

// Device time (us per launch, CUDA events) of one kernel of the tridiagonalisation on an n x n problem
// at column i = 0 of a panel position j = 16: kind 0 = trd_symv (option-selected), 1 = trd_head.
double microbench_trd(lfb_handle &h, int kind, int64_t n, int reps) {
    using T = double;
    const int64_t ld = round_up(n, 2), ldv = ld;
    DevBuf<T> A(h, (size_t)ld * n), V(h, (size_t)ldv * TB), Wm(h, (size_t)ldv * TB), P(h, ldv), P2(h, ldv), off(h, n);
    DevBuf<TrdAcc<T>> acc(h, 1);
    LFB_CUDA(cudaMemsetAsync(A.get(), 0, sizeof(T) * ld * n, h.stream));
    LFB_CUDA(cudaMemsetAsync(V.get(), 0, sizeof(T) * ldv * TB, h.stream));
    LFB_CUDA(cudaMemsetAsync(Wm.get(), 0, sizeof(T) * ldv * TB, h.stream));
    LFB_CUDA(cudaMemsetAsync(P.get(), 0, sizeof(T) * ldv, h.stream));
    TrdAcc<T> hacc;
    memset(&hacc, 0, sizeof hacc);
    hacc.some = 1;
    LFB_CUDA(cudaMemcpyAsync(acc.get(), &hacc, sizeof hacc, cudaMemcpyHostToDevice, h.stream));
    cudaEvent_t e0, e1;
    LFB_CUDA(cudaEventCreate(&e0));
    LFB_CUDA(cudaEventCreate(&e1));
    auto once = [&]() {
        if (kind == 0) launch_symv<T>(h, A.get(), ld, n, 1, V.get() + 16 * ldv, T(2), P.get(), acc.get());
        else launch_head<T>(h, head_cluster_size(h, n), A.get(), ld, n, 0, 16, V.get(), Wm.get(), ldv, P.get(), P2.get(), off.get(), acc.get(), 0);
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

template void sym_tridiagonal<float>(lfb_handle &, float *, int64_t, int64_t, float *);
template void sym_tridiagonal<double>(lfb_handle &, double *, int64_t, int64_t, double *);

}  // namespace lfb
