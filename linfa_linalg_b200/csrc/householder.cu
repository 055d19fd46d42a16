// Householder panel kernels + blocked compact-WY drivers (QR, assemble_q, qt_mul).
//
// Replaces: src/householder.rs:9-28 (reflection_axis_mut), :34-51 (clear_column), :68-93
// (assemble_q); src/reflection.rs:26-32 (reflect_cols hot loop); src/qr.rs:38-41 (driver loop),
// :110-120 (qt_mul).
//
// Mathematics: the reference applies s_j * H_j per column (H_j = I - 2 v v^T with a
// UNIT-NORM v, s_j = signum of the returned pivot).  Because the +-1 row scalings commute with all
// later reflectors, the whole factorisation equals a standard Householder QR (no scaling) followed
// by one O(mn) sign fix-up:  P_j = sgn(beta_j) (P_j = P_{j-1} for a zero column),
//   diag_ref[j] = P_{j-1} beta_j,  R_ref[i, j>i] = P_i R[i, j],  v_ref,j = P_{j-1} v_j.
// The standard QR is blocked: BLAS-2 sub-panels of width <= 32 (one fused kernel per column: scale
// the reflector, update the sub-panel, and accumulate the NEXT column's Gram row in the same pass,
// so a column costs one launch and one pass over the sub-panel), compact-WY block reflectors
// I - V T V^T with T = (striu(V^T V) + I/2)^-1 (tau = 2 for unit-norm v), and GEMM trailing updates.
#include <cooperative_groups.h>

#include <array>
#include <map>
#include <vector>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace lfb {

// panel_cluster.cu
template <typename T>
bool factor_subpanel_cluster2(lfb_handle &h, T *A, int64_t ld, int64_t m, int64_t c0, int wmax, int *w_out, T *beta, T *V,
                              int64_t ldv, int64_t vrow0, int vcol0, T *Tout, int ldt);

namespace {

constexpr int W = 32;  // max sub-panel width

template <typename T> __device__ __forceinline__ T t_sqrt(T x);
template <> __device__ __forceinline__ double t_sqrt<double>(double x) { return sqrt(x); }
template <> __device__ __forceinline__ float t_sqrt<float>(float x) { return sqrtf(x); }
template <typename T> __device__ __forceinline__ T t_abs(T x) { return x < T(0) ? -x : x; }
// Rust signum: +1 for +0.0, -1 for -0.0
template <typename T> __device__ __forceinline__ T t_signum(T x) { return signbit(x) ? T(-1) : T(1); }

// 1/sqrt(d): f32 seed + 2 Newton steps (full f64 accuracy for d inside the f32 range)
template <typename T> __device__ __forceinline__ T fast_rsqrt(T d);
template <> __device__ __forceinline__ float fast_rsqrt<float>(float d) { return rsqrtf(d); }
template <> __device__ __forceinline__ double fast_rsqrt<double>(double d) {
    if (d > 1e-30 && d < 1e30) {
        double y = (double)rsqrtf((float)d);
#pragma unroll
        for (int it = 0; it < 2; ++it) {
            const double r = fma(-d * y, y, 1.0);
            y = fma(0.5 * y, r, y);
        }
        return y;
    }
    return rsqrt(d);
}

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-reduce `cnt` per-thread partials (part[0..cnt)) and atomically add them to out[0..cnt).
template <typename T, int NT>
__device__ __forceinline__ void block_reduce_add(T *part, int cnt, T *out, T *sred /* [NT/32][W] */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int t = 0; t < W; ++t) {
        if (t < cnt) {
            T s = warp_sum(part[t]);
            if (lane == 0) sred[warp * W + t] = s;
        }
    }
    __syncthreads();
    if (threadIdx.x < cnt) {
        T s = T(0);
#pragma unroll
        for (int w = 0; w < NT / 32; ++w) s += sred[w * W + threadIdx.x];
        atomicAdd(out + threadIdx.x, s);
    }
}

// Gram row + head row of the first column of a sub-panel:
//   gram[t] = sum_{r >= c} A[r,c] A[r,c+t],  head[t] = A[c, c+t],  t = 0..w-1
template <typename T>
__global__ void __launch_bounds__(256) hh_gram_init(const T *__restrict__ A, int64_t ld, int64_t m, int64_t c, int w,
                                                    T *gram, T *head) {
    __shared__ T sred[8 * W];
    T part[W];
#pragma unroll
    for (int t = 0; t < W; ++t) part[t] = T(0);
    const T *col = A + c * ld;
    for (int64_t r = c + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < m; r += (int64_t)gridDim.x * blockDim.x) {
        T x = col[r];
#pragma unroll
        for (int t = 0; t < W; ++t)
            if (t < w) {
                T a = col[r + t * ld];
                part[t] += x * a;
                if (r == c) head[t] = a;
            }
    }
    block_reduce_add<T, 256>(part, w, gram, sred);
}

// One Householder column step (see file header).  Column c = c0 + j of a sub-panel [c0, c0+w).
//   gram_cur[t] = x . A[:, c+t] over rows >= c (t = 0..w-j-1), head_cur[t] = A[c, c+t].
// Writes v in place, updates columns c+1 .. c0+w-1, accumulates gram_next / head_next for column
// c+1, stores beta[c] (= -signed_norm, 0 for a `None` column), zeroes gram_zero for the step after.
template <typename T>
__global__ void __launch_bounds__(256) hh_col_step(T *__restrict__ A, int64_t ld, int64_t m, int64_t c, int nrem,
                                                   const T *__restrict__ gram_cur, T *gram_next, T *gram_zero,
                                                   const T *__restrict__ head_cur, T *head_next, T *beta) {
    __shared__ T sred[8 * W];
    __shared__ T sfac[W];
    __shared__ T sscal[3];  // s, d, some
    if (threadIdx.x == 0) {
        T nsq = gram_cur[0];
        T nrm = t_sqrt(nsq);                          // householder.rs:13
        T f = head_cur[0];
        T s = t_signum(f) * nrm;                      // :16
        T newsq = (nsq + t_abs(f) * nrm) * T(2);      // :19-20
        bool some = newsq != T(0);                    // :22  (false for an all-zero / underflowed column)
        T d = t_sqrt(newsq);
        sscal[0] = s; sscal[1] = d; sscal[2] = some ? T(1) : T(0);
        if (blockIdx.x == 0) beta[c] = some ? -s : T(0);   // :24 / :26
    }
    if (blockIdx.x == 0 && threadIdx.x < W) gram_zero[threadIdx.x] = T(0);
    __syncthreads();
    const T s = sscal[0], d = sscal[1];
    const bool some = sscal[2] != T(0);
    if (threadIdx.x >= 1 && threadIdx.x <= nrem) {
        // v . A_t = (x . A_t + s * A_t[head]) / d ; reflection.rs:29 factor = -2 * (axis . col)
        int t = threadIdx.x;
        sfac[t] = some ? T(-2) * ((gram_cur[t] + s * head_cur[t]) / d) : T(0);
    }
    __syncthreads();

    T part[W];
#pragma unroll
    for (int t = 0; t < W; ++t) part[t] = T(0);
    T *col = A + c * ld;
    for (int64_t r = c + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < m; r += (int64_t)gridDim.x * blockDim.x) {
        T x = col[r];
        T v = T(0);
        if (some) {
            v = ((r == c) ? x + s : x) / d;           // householder.rs:17,23
            col[r] = v;
        }
        T a1 = T(0);
#pragma unroll
        for (int t = 1; t < W; ++t)
            if (t <= nrem) {
                T a = col[r + t * ld];
                if (some) {
                    a += sfac[t] * v;                 // reflection.rs:30 scaled_add
                    col[r + t * ld] = a;
                }
                if (t == 1) a1 = a;
                if (r > c) {
                    part[t - 1] += a1 * a;
                    if (r == c + 1) head_next[t - 1] = a;
                }
            }
    }
    if (nrem > 0) block_reduce_add<T, 256>(part, nrem, gram_next, sred);
}

// ------------------------------------------------------------------------------------------------
// Cluster panel kernel: a whole sub-panel (rows x w, w <= 32) lives in the distributed shared memory
// of ONE thread-block cluster (<= 16 CTAs, each owning a slab of rows).  Per column: one pass over
// the slab (scale reflector, update the remaining columns, accumulate the full Gram row of the NEXT
// column against all w columns), one deterministic cluster-wide reduction through DSMEM, one
// cluster barrier.  That replaces one kernel launch per column (14.7 us measured) by ~1 us.
// The same Gram rows give v_i . v_j for free, so the kernel also emits the compact-WY factor
// T = (striu(V^T V) + I/2)^-1 of the sub-panel and the staged V (zero above the diagonal, zero for
// `None` columns) that the trailing GEMMs consume.
template <typename T>
struct PanelArgs {
    T *A; int64_t ld, m, c0;
    int w, rpc;
    T *beta;
    T *V; int64_t ldv, vrow0; int vcol0;   // V workspace of the enclosing panel (rows from the panel's first row)
    T *Tout; int ldt;
};

template <typename T>
__global__ void __launch_bounds__(512, 1) hh_panel_cluster(PanelArgs<T> p) {
    cg::cluster_group cluster = cg::this_cluster();
    const int nc = (int)cluster.num_blocks();
    const int b = (int)cluster.block_rank();
    extern __shared__ __align__(16) unsigned char panel_smem[];
    const int w = p.w, rpc = p.rpc;
    T *S = reinterpret_cast<T *>(panel_smem);  // [w][rpc]
    T *exq = S + (size_t)w * rpc;              // [2][W] partial Gram rows (read remotely)
    T *exh = exq + 2 * W;                      // [2][W] head row (valid in the owner CTA)
    T *red = exh + 2 * W;                      // [16][W]
    T *G = red + 16 * W;                       // [W][W]  G[k*W + j] = v_k . v_j (k < j)
    T *fac = G + W * W;                        // [W]
    T *qv = fac + W;                           // [W]
    T *hv = qv + W;                            // [W]
    T *somef = hv + W;                         // [W]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t grow0 = p.c0 + (int64_t)b * rpc;
    const int nrows = (int)max((int64_t)0, min((int64_t)rpc, p.m - grow0));

    for (int k = 0; k < w; ++k)
        for (int lr = tid; lr < rpc; lr += 512)
            S[(size_t)k * rpc + lr] = lr < nrows ? p.A[(grow0 + lr) + (p.c0 + k) * p.ld] : T(0);
    for (int e = tid; e < W * W; e += 512) G[e] = T(0);
    __syncthreads();

    T part[W];
    // warp-transposed reduction of part[] + cross-warp sum; result for column k lands in dst[k]
    auto reduce_to = [&](T *dst) {
#pragma unroll
        for (int off = 16, n = 16; off >= 1; off >>= 1, n >>= 1) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                if (i < n) {
                    const bool up = (lane & off) != 0;
                    const T send = up ? part[i] : part[i + n];
                    const T keep = up ? part[i + n] : part[i];
                    part[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
                }
            }
        }
        red[warp * W + lane] = part[0];
        __syncthreads();
        if (tid < W) {
            T s = T(0);
#pragma unroll
            for (int wv = 0; wv < 16; ++wv) s += red[wv * W + tid];
            dst[tid] = s;
        }
    };

    // Gram row of column 0 and its head row
#pragma unroll
    for (int k = 0; k < W; ++k) part[k] = T(0);
    for (int lr = tid; lr < nrows; lr += 512) {
        const T x = S[lr];
#pragma unroll
        for (int k = 0; k < W; ++k)
            if (k < w) part[k] += S[(size_t)k * rpc + lr] * x;
    }
    reduce_to(exq);
    if (b == 0 && tid < w) exh[tid] = S[(size_t)tid * rpc];
    cluster.sync();

    for (int j = 0; j < w; ++j) {
        const int par = j & 1;
        const int64_t c = p.c0 + j;
        // ---- gather the cluster-wide Gram row (fixed order => deterministic) and the head row ----
        // one remote (DSMEM) load per thread: thread (bb, k) fetches CTA bb's partial for column k
        {
            const int bb = tid >> 5, k = tid & 31;   // 16 x 32 = 512 threads
            red[bb * W + k] = bb < nc ? cluster.map_shared_rank(exq, bb)[par * W + k] : T(0);
            if (bb == 0) hv[k] = cluster.map_shared_rank(exh, j / rpc)[par * W + k];
        }
        __syncthreads();
        if (tid < W) {
            T q = T(0);
#pragma unroll
            for (int bb = 0; bb < 16; ++bb) q += red[bb * W + tid];
            qv[tid] = q;
        }
        __syncthreads();
        const T nsq = qv[j], f = hv[j];
        // householder.rs:13-23 with the two square roots taken as x*rsqrt(x) (f32 seed + Newton, then one
        // correction step): the per-column latency chain is what bounds this kernel.
        const T rn = nsq > T(0) ? fast_rsqrt(nsq) : T(0);
        T nrm = nsq * rn;
        nrm = fma(T(0.5) * rn, fma(-nrm, nrm, nsq), nrm);
        const T s = t_signum(f) * nrm;                   // :16
        const T newsq = (nsq + t_abs(f) * nrm) * T(2);   // :19-20
        const bool some = newsq != T(0);                 // :22
        const T rd = some ? fast_rsqrt(newsq) : T(0);    // 1 / sqrt(new_norm_sq)
        if (tid < W) {
            const int k = tid;
            const T dotv = some ? (qv[k] + s * hv[k]) * rd : T(0);   // v_j . (column k)
            fac[k] = (k > j && k < w) ? T(-2) * dotv : T(0);        // reflection.rs:29
            if (k < j) G[k * W + j] = (somef[k] != T(0)) ? dotv : T(0);
            if (k == j) somef[j] = some ? T(1) : T(0);
            if (k == 0 && b == 0) p.beta[c] = some ? -s : T(0);     // householder.rs:24/26
        }
        __syncthreads();
        // ---- one pass over the slab ----
#pragma unroll
        for (int k = 0; k < W; ++k) part[k] = T(0);
        const bool has_next = j + 1 < w;
        const int lr_min = (int)max((int64_t)0, c - grow0);
        for (int lr = lr_min + tid; lr < nrows; lr += 512) {
            const int64_t gr = grow0 + lr;
            T v = S[(size_t)j * rpc + lr];
            if (some) {
                v = ((gr == c) ? v + s : v) * rd;        // householder.rs:17,23
                S[(size_t)j * rpc + lr] = v;
            }
            T a1 = T(0);
            if (has_next) {
                a1 = S[(size_t)(j + 1) * rpc + lr];
                if (some) {
                    a1 += fac[j + 1] * v;                // reflection.rs:30
                    S[(size_t)(j + 1) * rpc + lr] = a1;
                }
            }
            const bool nxt = has_next && gr > c;
#pragma unroll
            for (int k = 0; k < W; ++k) {
                if (k < w) {
                    T a;
                    if (k == j) a = v;
                    else if (k == j + 1) a = a1;
                    else {
                        a = S[(size_t)k * rpc + lr];
                        if (k > j && some) {
                            a += fac[k] * v;
                            S[(size_t)k * rpc + lr] = a;
                        }
                    }
                    if (nxt) part[k] += a * a1;
                }
            }
        }
        if (has_next) {
            reduce_to(exq + (par ^ 1) * W);
            __syncthreads();
            const int owner = (j + 1) / rpc, lrh = (j + 1) % rpc;
            if (b == owner && tid < w) exh[(par ^ 1) * W + tid] = S[(size_t)tid * rpc + lrh];
        }
        cluster.sync();
    }

    // ---- write back: A (in place), staged V, and T (CTA 0) ----
    for (int k = 0; k < w; ++k) {
        const bool sk = somef[k] != T(0);
        for (int lr = tid; lr < nrows; lr += 512) {
            const T val = S[(size_t)k * rpc + lr];
            const int64_t gr = grow0 + lr;
            p.A[gr + (p.c0 + k) * p.ld] = val;
            p.V[(p.vrow0 + (gr - p.c0)) + (int64_t)(p.vcol0 + k) * p.ldv] = (sk && gr >= p.c0 + k) ? val : T(0);
        }
    }
    if (b == 0) {
        for (int64_t e = tid; e < p.vrow0 * w; e += 512) p.V[(e % p.vrow0) + (p.vcol0 + e / p.vrow0) * p.ldv] = T(0);
        // T = (striu(G) + I/2)^-1, one thread per column (w <= 32)
        if (tid < w) {
            const int jc = tid;
            T tcol[W];
#pragma unroll
            for (int i = 0; i < W; ++i) tcol[i] = T(0);
#pragma unroll
            for (int i = W - 1; i >= 0; --i) {
                if (i == jc) tcol[i] = T(2);
                else if (i < jc) {
                    T sum = T(0);
#pragma unroll
                    for (int k = 0; k < W; ++k)
                        if (k > i && k <= jc) sum += G[i * W + k] * tcol[k];
                    tcol[i] = T(-2) * sum;
                }
            }
#pragma unroll
            for (int i = 0; i < W; ++i)
                if (i < w) p.Tout[i + jc * p.ldt] = tcol[i];
        }
    }
}

// Vout (rows x w, ldv) = V part of A[r0:, c0:c0+w): element (i, j) is zero for i < j ("upper" part
// holds R), and whole column j is zeroed when beta != nullptr and beta[c0+j] == 0 (a `None` column:
// the reference applies no reflection there).
template <typename T>
__global__ void copy_v_kernel(const T *__restrict__ A, int64_t ld, int64_t r0, int64_t c0, int64_t rows, int w,
                              const T *__restrict__ beta, T *__restrict__ V, int64_t ldv) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= rows) return;
    for (int j = blockIdx.y; j < w; j += gridDim.y) {
        T v = T(0);
        if (i >= j && !(beta && beta[c0 + j] == T(0))) v = A[(r0 + i) + (c0 + j) * ld];
        V[i + (int64_t)j * ldv] = v;
    }
}

// T = (striu(G) + I/2)^-1 for an nb x nb Gram matrix G = V^T V (single CTA, nb <= 128).
// Column j of T solves U t = e_j by back substitution; one thread per column.
// In-place recursive doubling (6-7 levels, each two small block products) instead of one
// back-substitution per column: inv([A B; 0 C]) = [A^-1, -A^-1 B C^-1; 0, C^-1].
template <typename T>
__global__ void __launch_bounds__(512) tinv_kernel(const T *__restrict__ G, int64_t ldg, int nb, T *__restrict__ Tm, int64_t ldt) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int P = 1;
    while (P < nb) P <<= 1;                   // padded size (<= 128)
    const int lds = P + 1;
    T *X = reinterpret_cast<T *>(smem_raw);   // [P][P+1]  starts as U, ends as U^-1
    T *tmp = X + P * lds;                     // [P/2][P+1]
    const int tid = threadIdx.x;
    for (int e = tid; e < P * P; e += 512) {
        const int i = e % P, k = e / P;
        T v = T(0);
        if (i == k) v = T(2);                                       // 1 / U_ii, U_ii = 1/2
        else if (i < k && k < nb) v = G[i + (int64_t)k * ldg];      // striu(V^T V)
        X[i * lds + k] = v;
    }
    __syncthreads();
    for (int b = 1; b < P; b <<= 1) {
        const int npair = P / (2 * b), per = b * b;
        // tmp = B * C^-1   (B = X[o.., o+b..) original, C^-1 = X[o+b.., o+b..) upper)
        for (int e = tid; e < npair * per; e += 512) {
            const int p = e / per, i = (e % per) / b, j = e % b, o = p * 2 * b;
            T acc = T(0);
            for (int k = 0; k <= j; ++k) acc += X[(o + i) * lds + o + b + k] * X[(o + b + k) * lds + o + b + j];
            tmp[(p * b + i) * lds + j] = acc;
        }
        __syncthreads();
        // X12 = -A^-1 * tmp   (A^-1 = X[o.., o..) upper)
        for (int e = tid; e < npair * per; e += 512) {
            const int p = e / per, i = (e % per) / b, j = e % b, o = p * 2 * b;
            T acc = T(0);
            for (int k = i; k < b; ++k) acc += X[(o + i) * lds + o + k] * tmp[(p * b + k) * lds + j];
            X[(o + i) * lds + o + b + j] = -acc;
        }
        __syncthreads();
    }
    for (int e = tid; e < nb * nb; e += 512) {
        const int i = e % nb, k = e / nb;
        Tm[i + (int64_t)k * ldt] = (i <= k) ? X[i * lds + k] : T(0);
    }
}

// psign[j] = P_j (running sign, header comment); diag[j] = P_{j-1} * beta[j].  Single thread scan.
template <typename T>
__global__ void sign_scan_kernel(const T *__restrict__ beta, int64_t n, T *__restrict__ psign, T *__restrict__ diag) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    T p = T(1);
    for (int64_t j = 0; j < n; ++j) {
        T b = beta[j];
        // a `None` column (householder.rs:26, :50: `rn.unwrap_or(A::zero())`) stores +0.0, never -0.0: the reference's
        // consumers take signum(diag[j]) (householder.rs:45, :89; qr.rs:116) and Rust's signum(-0.0) is -1
        diag[j] = (b != T(0)) ? p * b : T(0);
        if (b != T(0)) p = t_signum(b);
        psign[j] = p;
    }
}

// R part (r < c): *= P_r ; V part (r >= c): *= P_{c-1}.
template <typename T>
__global__ void sign_fix_kernel(T *A, int64_t ld, int64_t m, int64_t n, const T *__restrict__ psign) {
    int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r >= m) return;
    T pr = r < n ? psign[r] : T(1);
    for (int64_t c = blockIdx.y; c < n; c += gridDim.y) {
        T f = (r < c) ? pr : (c > 0 ? psign[c - 1] : T(1));
        if (f < T(0)) A[r + c * ld] = -A[r + c * ld];
    }
}

// The same two passes for ONE panel [k0, k0 + nb), so that a finished block column can leave for the host while the
// factorisation goes on (api.cu: qr_host): the running sign continues from *carry (P_{k0-1}; 1 before the first panel).
template <typename T>
__global__ void sign_scan_panel_kernel(const T *__restrict__ beta, int64_t k0, int nb, T *__restrict__ carry, T *__restrict__ psign,
                                       T *__restrict__ diag) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    T p = *carry;
    for (int64_t j = k0; j < k0 + nb; ++j) {
        const T b = beta[j];
        diag[j] = (b != T(0)) ? p * b : T(0);
        if (b != T(0)) p = t_signum(b);
        psign[j] = p;
    }
    *carry = p;
}

template <typename T>
__global__ void sign_fix_panel_kernel(T *A, int64_t ld, int64_t m, int64_t k0, int nb, const T *__restrict__ psign) {
    const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r >= m) return;
    for (int64_t c = k0 + blockIdx.y; c < k0 + nb; c += gridDim.y) {
        const T f = (r < c) ? psign[r] : (c > 0 ? psign[c - 1] : T(1));
        if (f < T(0)) A[r + c * ld] = -A[r + c * ld];
    }
}

// cum[j] = prod_{i <= j} signum(signs[i]) for j < cnt.
template <typename T>
__global__ void sign_cumprod_kernel(const T *__restrict__ signs, int64_t cnt, T *__restrict__ cum) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    T p = T(1);
    for (int64_t j = 0; j < cnt; ++j) {
        p *= t_signum(signs[j]);
        cum[j] = p;
    }
}

// Q[:, c] *= cum[min(c - shift, cnt-1)]  (c >= shift)
template <typename T>
__global__ void scale_cols_kernel(T *Q, int64_t ldq, int64_t rows, int64_t cols, int64_t shift, int64_t cnt,
                                  const T *__restrict__ cum) {
    int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r >= rows) return;
    for (int64_t c = blockIdx.y; c < cols; c += gridDim.y) {
        int64_t k = c - shift;
        if (k < 0 || cnt <= 0) continue;
        if (k > cnt - 1) k = cnt - 1;
        if (cum[k] < T(0)) Q[r + c * ldq] = -Q[r + c * ldq];
    }
}

// B[r, :] *= cum[min(r, cnt-1)]
template <typename T>
__global__ void scale_rows_kernel(T *B, int64_t ldb, int64_t rows, int64_t cols, int64_t cnt, const T *__restrict__ cum) {
    int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r >= rows || cnt <= 0) return;
    T f = cum[r < cnt ? r : cnt - 1];
    if (!(f < T(0))) return;
    for (int64_t c = blockIdx.y; c < cols; c += gridDim.y) B[r + c * ldb] = -B[r + c * ldb];
}

inline unsigned ycap(int64_t n) { return (unsigned)(n < 1 ? 1 : (n < 65535 ? n : 65535)); }

template <typename T>
void copy_v(lfb_handle &h, const T *A, int64_t ld, int64_t r0, int64_t c0, int64_t rows, int w, const T *beta, T *V,
            int64_t ldv) {
    if (rows <= 0 || w <= 0) return;
    dim3 grid((unsigned)cdiv(rows, 256), ycap(w));
    copy_v_kernel<T><<<grid, 256, 0, h.stream>>>(A, ld, r0, c0, rows, w, beta, V, ldv);
    LFB_LAUNCH_CHECK(h);
}

// Tm (nb x nb, ldt) from V (rows x nb, ldv); G is an nb x nb scratch.
template <typename T>
void build_t(lfb_handle &h, const T *V, int64_t ldv, int64_t rows, int nb, T *G, T *Tm, int64_t ldt) {
    gemm<T>(h, 1, 0, nb, nb, rows, T(1), V, ldv, V, ldv, T(0), G, nb);
    int P = 1;
    while (P < nb) P <<= 1;
    size_t smem = sizeof(T) * (size_t)(P * (P + 1) + (P / 2 + 1) * (P + 1));
    static DeviceOnce cfg;   // function attributes are per device
    cfg.run(h.device, [&] {
        LFB_CUDA(cudaFuncSetAttribute(tinv_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)(sizeof(T) * (128 * 129 + 65 * 129))));
    });
    if (nb <= 128) {
        tinv_kernel<T><<<1, 512, smem, h.stream>>>(G, nb, nb, Tm, ldt);
        LFB_LAUNCH_CHECK(h);
        return;
    }
    // nb in (128, 256]: T = [T1, -T1 G12 T2; 0, T2] with T1, T2 from the two diagonal blocks of G
    const int n1 = 128, n2 = nb - 128;
    smem = sizeof(T) * (size_t)(128 * 129 + 65 * 129);
    tinv_kernel<T><<<1, 512, smem, h.stream>>>(G, nb, n1, Tm, ldt);
    LFB_LAUNCH_CHECK(h);
    tinv_kernel<T><<<1, 512, smem, h.stream>>>(G + n1 + (int64_t)n1 * nb, nb, n2, Tm + n1 + (int64_t)n1 * ldt, ldt);
    LFB_LAUNCH_CHECK(h);
    fill<T>(h, Tm + n1, n2, n1, ldt, T(0), T(0));                                                   // T21 = 0
    T *X = G + n1;   // the (unused) lower-left block of G as scratch: n2 x n1 region holds n1 x n2? no: use rows n1.., cols 0..n1
    // X (n1 x n2) = T1 * G12 ; stored in a separate column block of G is not possible without clobbering G12,
    // so write it over G21 (n2 x n1) transposed-free: use ldg = nb with the n1 x n2 product placed at G[0:n1, 0:n2]
    // (G11 is no longer needed once T1 exists).
    X = G;
    gemm<T>(h, 0, 0, n1, n2, n1, T(1), Tm, ldt, G + (int64_t)n1 * nb, nb, T(0), X, nb);                // X = T1 G12
    gemm<T>(h, 0, 0, n1, n2, n2, T(-1), X, nb, Tm + n1 + (int64_t)n1 * ldt, ldt, T(0), Tm + (int64_t)n1 * ldt, ldt);   // T12 = -X T2
}

// C (rows x ncols, ldc) <- (I - V op(T) V^T) C,  op(T) = T^T if trans_t.  W1/W2: nb x ncols scratch.
// Vt (optional): V^T stored nb x rows (leading dimension ldvt), so that the rank-nb update is a K-major ("TN") product
// on both sides -- the only operand layout of the f32 tensor-core kernel (gemm_tf32.cu).
template <typename T>
void apply_block_reflector(lfb_handle &h, const T *V, int64_t ldv, int64_t rows, int nb, const T *Tm, int64_t ldt,
                           int trans_t, T *C, int64_t ldc, int64_t ncols, T *W1, T *W2, const T *Vt = nullptr, int64_t ldvt = 0,
                           const T *VT = nullptr) {
    if (ncols <= 0 || rows <= 0) return;
    if (VT && trans_t) {
        // T folded into the reflectors once per panel (VT = V T): W2 = T^T (V^T C) = (V T)^T C is ONE long-K product instead of a
        // long-K product followed by a 128 x ncols x 128 one that runs almost alone on the main stream
        gemm<T>(h, 1, 0, nb, ncols, rows, T(1), VT, ldv, C, ldc, T(0), W2, nb);
    } else {
        gemm<T>(h, 1, 0, nb, ncols, rows, T(1), V, ldv, C, ldc, T(0), W1, nb);          // W1 = V^T C
        gemm<T>(h, trans_t, 0, nb, ncols, nb, T(1), Tm, ldt, W1, nb, T(0), W2, nb);     // W2 = op(T) W1
    }
    if (Vt) gemm<T>(h, 1, 0, rows, ncols, nb, T(-1), Vt, ldvt, W2, nb, T(1), C, ldc);   // C -= (V^T)^T W2
    else gemm<T>(h, 0, 0, rows, ncols, nb, T(-1), V, ldv, W2, nb, T(1), C, ldc);    // C -= V W2
}

// Unblocked-in-sub-panel Householder on columns [c0, c0+w) of A (m x ., ld), rows >= c0.
template <typename T>
void factor_subpanel(lfb_handle &h, T *A, int64_t ld, int64_t m, int64_t c0, int w, T *beta, T *scratch /* 5*W */) {
    T *gram[3] = {scratch, scratch + W, scratch + 2 * W};
    T *head[2] = {scratch + 3 * W, scratch + 4 * W};
    LFB_CUDA(cudaMemsetAsync(scratch, 0, sizeof(T) * 5 * W, h.stream));
    int64_t rows = m - c0;
    unsigned nblk = (unsigned)std::min<int64_t>(cdiv(rows, 256), 2 * h.sm_count);
    if (nblk < 1) nblk = 1;
    hh_gram_init<T><<<nblk, 256, 0, h.stream>>>(A, ld, m, c0, w, gram[0], head[0]);
    LFB_LAUNCH_CHECK(h);
    for (int j = 0; j < w; ++j) {
        int64_t c = c0 + j;
        int64_t r = m - c;
        unsigned nb = (unsigned)std::min<int64_t>(cdiv(r, 256), 2 * h.sm_count);
        if (nb < 1) nb = 1;
        hh_col_step<T><<<nb, 256, 0, h.stream>>>(A, ld, m, c, w - j - 1, gram[j % 3], gram[(j + 1) % 3], gram[(j + 2) % 3],
                                                 head[j % 2], head[(j + 1) % 2], beta);
        LFB_LAUNCH_CHECK(h);
    }
}

// Picks (cluster size, sub-panel width, rows per CTA) so that the sub-panel fits the cluster's
// distributed shared memory; returns false if it cannot (very tall panels) or clusters are off.
template <typename T>
bool plan_cluster(lfb_handle &h, int64_t rows, int wmax, int *nc, int *w, int *rpc, size_t *smem) {
    if (!h.opt.panel_cluster || rows <= 0) return false;
    const size_t extra = sizeof(T) * (size_t)(24 * W + W * W);
    if (h.smem_optin <= extra + 4096) return false;
    const int64_t cap = (int64_t)((h.smem_optin - extra - 1024) / sizeof(T));   // elements of S per CTA
    int c = rows > 2048 ? 16 : rows > 1024 ? 8 : rows > 512 ? 4 : rows > 256 ? 2 : 1;
    if (c > h.opt.panel_cluster_max) c = (int)h.opt.panel_cluster_max;
    int64_t r = round_up(cdiv(rows, c), 32);
    int ww = wmax;
    while (ww > 8 && r * ww > cap) ww -= 8;
    if (r * ww > cap) return false;
    *nc = c; *w = ww; *rpc = (int)r;
    *smem = sizeof(T) * (size_t)(r * ww) + extra;
    return true;
}

template <typename T>
bool factor_subpanel_cluster(lfb_handle &h, T *A, int64_t ld, int64_t m, int64_t c0, int nc, int w, int rpc, size_t smem,
                             T *beta, T *V, int64_t ldv, int64_t vrow0, int vcol0, T *Tout, int ldt) {
    static DeviceOnce cfg;   // function attributes are per device
    cfg.run(h.device, [&] {
        LFB_CUDA(cudaFuncSetAttribute(hh_panel_cluster<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h.smem_optin));
        LFB_CUDA(cudaFuncSetAttribute(hh_panel_cluster<T>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    });
    PanelArgs<T> p;
    p.A = A; p.ld = ld; p.m = m; p.c0 = c0; p.w = w; p.rpc = rpc; p.beta = beta;
    p.V = V; p.ldv = ldv; p.vrow0 = vrow0; p.vcol0 = vcol0; p.Tout = Tout; p.ldt = ldt;
    cudaLaunchConfig_t cfgl = {};
    cfgl.gridDim = dim3((unsigned)nc);
    cfgl.blockDim = dim3(512);
    cfgl.dynamicSmemBytes = smem;
    cfgl.stream = h.stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)nc;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfgl.attrs = attr;
    cfgl.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfgl, hh_panel_cluster<T>, p);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return false;   // e.g. the cluster cannot be co-scheduled: caller falls back
    }
    h.launches++;
    return true;
}

}  // namespace

// Standard (unscaled) blocked Householder QR of A (m x n, m >= n); beta[n] on device.
// fin != nullptr (host path with overlapped download): after a panel is factored its block column gets the reference's signs at
// once (fin->psign / fin->diag / fin->carry) and h.qr_panel_hook is told; the caller then skips the global sign pass.
template <typename T> struct QrFinish { T *psign, *diag, *carry; };
template <typename T>
static void qr_factor_std(lfb_handle &h, T *A, int64_t m, int64_t n, int64_t ld, T *beta, const QrFinish<T> *fin = nullptr) {
    // f32 on the tensor-core GEMM: K = nb = 256 amortises the per-tile prologue / epilogue of the rank-nb update (124 vs 141 ms at 16384^2)
    const bool f32_tc = sizeof(T) == 4 && h.opt.sgemm_tc != 0 && n >= 2048;
    const int NB = (int)std::max<int64_t>(W, std::min<int64_t>(f32_tc ? h.opt.qr_nb_f32 : h.opt.qr_nb, 256));
    const int SUB = (int)std::max<int64_t>(1, std::min<int64_t>(h.opt.qr_sub, W));
    const int64_t ldv = round_up(m, 4);     // 16-byte columns for TMA in both precisions
    // f32 with the tensor-core GEMM: a transposed copy of every finished panel's V (nb x rows, K-major) for C -= V W
    // (f64, option qr_vt: the same copy turns C -= V W into the TMA kernel's K-major/K-major variant, 2 box loads per stage instead of 9)
    const bool use_vt = ((sizeof(T) == 4 && h.opt.sgemm_tc != 0) || (sizeof(T) == 8 && h.opt.qr_vt != 0)) && m >= 256 && n > NB;
    DevBuf<T> Vt0(h, use_vt ? (size_t)NB * ldv : 1), Vt1(h, use_vt ? (size_t)NB * ldv : 1);
    T *Vtbuf[2] = {use_vt ? Vt0.get() : nullptr, use_vt ? Vt1.get() : nullptr};
    const bool fold_t = (sizeof(T) == 8 ? h.opt.qr_fold_t != 0 : h.opt.qr_fold_t >= 2) && n > NB;
    DevBuf<T> VT0(h, fold_t ? (size_t)NB * ldv : 1), VT1(h, fold_t ? (size_t)NB * ldv : 1);
    T *VTbuf[2] = {fold_t ? VT0.get() : nullptr, fold_t ? VT1.get() : nullptr};
    // two generations of the panel workspaces (V, T): with look-ahead panel k+1 is factored while the
    // trailing update of panel k still reads V_k / T_k
    DevBuf<T> Vb0(h, (size_t)ldv * NB), Vb1(h, (size_t)ldv * NB);
    DevBuf<T> Tb0(h, (size_t)NB * NB), Tb1(h, (size_t)NB * NB);
    T *Vbuf[2] = {Vb0.get(), Vb1.get()}, *Tbuf[2] = {Tb0.get(), Tb1.get()};
    DevBuf<T> G(h, (size_t)NB * NB), Ts(h, (size_t)W * W);
    DevBuf<T> W1p(h, (size_t)NB * NB), W2p(h, (size_t)NB * NB);   // panel-internal applies (<= NB columns)
    DevBuf<T> W1(h, (size_t)NB * std::max<int64_t>(n, 1)), W2(h, (size_t)NB * std::max<int64_t>(n, 1));
    DevBuf<T> scratch(h, 5 * W);

    // Factors panel [k0, k0+nb) on the handle's current stream; stages V (rows from k0) and, if the
    // panel has a trailing matrix, builds its compact-WY T.
    // The panel as a guarded Cholesky-QR + Householder reconstruction (cholqr.cu, tsqr_hr.cu): Gram GEMM, nb x nb Cholesky,
    // Q_top = A_top R^-1, LU of (Q_top - S) -> the reflectors of the top block, beta and U', rows below = A_2 (R^-1 U'^-1) -- the
    // SAME reflectors, pivots and R a Householder sweep over the panel produces (the reconstruction is exact, not "up to signs"),
    // in ~30 short launches that are GEMMs where the height matters, against 0.5 ms (128 rows) - 2.9 ms (16384 rows) for the
    // cluster kernels: with those the QR of 16384^2 is bound by the panel chain from panel 60 of 128 on (qr_trace,
    // profiles/r2_qr_panel.md).  Taken only when the guard accepts the panel (cond bound <= tsqr_cholqr_cond, Cholesky succeeded);
    // rank-deficient, graded or ill-conditioned panels go to the cluster kernels, A untouched by the attempt.
    auto panel_cholqr = [&](int64_t k0, int nb, T *V, T *Tm, bool *t_built) -> bool {
        const int64_t prow = m - k0;
        T *P = A + k0 + k0 * ld;
        const int64_t ldu = round_up(nb, 2);
        DevBuf<T> R(h, (size_t)ldu * nb), Rinv(h, (size_t)ldu * nb), U(h, (size_t)ldu * nb), Qt(h, (size_t)ldu * nb);
        if (!cholqr_factor<T>(h, P, prow, nb, ld, R.get(), ldu, Rinv.get(), ldu)) return false;
        if (nb == 128 && h.opt.cholqr_fused) {
            // U receives M = R^-1 U^-1 C; the top block, beta, the staged top of V and T come out of the same launch
            hr_panel128<T>(h, P, ld, R.get(), ldu, Rinv.get(), ldu, beta + k0, U.get(), ldu, Tm, NB, V, ldv);
            if (prow > nb) {
                gemm<T>(h, 0, 0, prow - nb, nb, nb, T(1), P + nb, ld, U.get(), ldu, T(0), V + nb, ldv);
                copy2d<T>(h, V + nb, ldv, P + nb, ld, prow - nb, nb);
            }
            hr_panel128_join(h);       // T (second side stream) is complete from here on
            *t_built = true;
            return true;
        }
        gemm<T>(h, 0, 0, nb, nb, nb, T(1), P, ld, Rinv.get(), ldu, T(0), Qt.get(), ldu);                 // Q_top
        hh_reconstruct_top<T>(h, Qt.get(), nb, ldu, R.get(), ldu, U.get(), ldu, beta + k0, /*internal=*/1);
        trsm_right_upper<T>(h, nb, nb, U.get(), ldu, Rinv.get(), ldu);                                   // Rinv <- R^-1 U'^-1
        if (prow > nb) {
            gemm<T>(h, 0, 0, prow - nb, nb, nb, T(1), P + nb, ld, Rinv.get(), ldu, T(0), V + nb, ldv);   // staged V, rows below the top block
            copy2d<T>(h, V + nb, ldv, P + nb, ld, prow - nb, nb);
        }
        copy2d<T>(h, Qt.get(), ldu, P, ld, nb, nb);
        copy_v<T>(h, A, ld, k0, k0, nb, nb, (const T *)nullptr, V, ldv);                                 // top of the staged V: zero above the diagonal
        return true;
    };

    auto finish_panel = [&](int64_t k0, int nb) {        // on the stream the panel was factored on
        if (!fin) return;
        sign_scan_panel_kernel<T><<<1, 32, 0, h.stream>>>(beta, k0, nb, fin->carry, fin->psign, fin->diag);
        LFB_LAUNCH_CHECK(h);
        dim3 grid((unsigned)cdiv(m, 256), (unsigned)std::min(nb, 64));
        sign_fix_panel_kernel<T><<<grid, 256, 0, h.stream>>>(A, ld, m, k0, nb, fin->psign);
        LFB_LAUNCH_CHECK(h);
        if (h.qr_panel_hook) h.qr_panel_hook(k0, nb);
    };
    auto factor_panel_raw = [&](int64_t k0, int nb, T *V, T *Tm, T *Vt) {
        bool t_built = false;
        const bool chq = (sizeof(T) == 8 ? h.opt.qr_panel_cholqr >= 1 : h.opt.qr_panel_cholqr >= 2);
        int s_start = 0;
        if (chq && nb == 256 && h.opt.cholqr_fused && (m - k0) >= 2 * (int64_t)nb) {
            // A 256-column panel (what the f32 tensor-core update wants as K) as TWO fused 128-column Cholesky-QR panels: factor A,
            // apply it to B's 128 columns, factor B, and assemble the compact-WY factor of the pair,
            //   T = [T_A, -T_A (V_A^T V_B) T_B; 0, T_B].
            // If B is declined by the guard the Householder sub-panels take over from column 128 (A stays).
            bool tA = false, tB = false;
            const int64_t prow = m - k0;
            if (panel_cholqr(k0, 128, V, Tm, &tA) && tA) {
                apply_block_reflector<T>(h, V, ldv, prow, 128, Tm, NB, /*trans_t=*/1, A + k0 + (k0 + 128) * ld, ld, 128, W1p, W2p);
                fill<T>(h, V + (int64_t)128 * ldv, 128, 128, ldv, T(0), T(0));      // V_B has no rows above its diagonal block
                if (panel_cholqr(k0 + 128, 128, V + 128 + (int64_t)128 * ldv, Tm + 128 + (int64_t)128 * NB, &tB) && tB) {
                    if (n - (k0 + nb) > 0) {
                        gemm<T>(h, 1, 0, 128, 128, prow - 128, T(1), V + 128, ldv, V + 128 + (int64_t)128 * ldv, ldv, T(0), G, 128);   // V_A^T V_B
                        gemm<T>(h, 0, 0, 128, 128, 128, T(1), Tm, NB, G, 128, T(0), W1p, 128);
                        gemm<T>(h, 0, 0, 128, 128, 128, T(-1), W1p, 128, Tm + 128 + (int64_t)128 * NB, NB, T(0), Tm + (int64_t)128 * NB, NB);
                        fill<T>(h, Tm + 128, 128, 128, NB, T(0), T(0));
                        if (Vt) transpose<T>(h, V, prow, nb, ldv, Vt, NB);
                    }
                    return;
                }
                s_start = 128;
            }
        }
        if (s_start == 0 && chq && nb >= 64 && (m - k0) >= 2 * (int64_t)nb && panel_cholqr(k0, nb, V, Tm, &t_built)) {
            if (n - (k0 + nb) > 0) {
                if (!t_built) build_t<T>(h, V, ldv, m - k0, nb, G, Tm, NB);
                if (Vt) transpose<T>(h, V, m - k0, nb, ldv, Vt, NB);
            }
            return;
        }
        bool v_staged = true;   // the cluster kernels stage V as they go
        for (int s0 = s_start; s0 < nb;) {
            const int64_t c0 = k0 + s0;
            const int64_t rows = m - c0;
            int w = std::min(SUB, nb - s0);
            int nc = 0, wc = 0, rpc = 0;
            size_t smem = 0;
            bool done = false;
            if (h.opt.panel_cluster >= 2) {
                done = factor_subpanel_cluster2<T>(h, A, ld, m, c0, w, &wc, beta, V, ldv, c0 - k0, s0, Ts, W);
                if (done) w = wc;
                else h.opt.panel_cluster = 1;   // refused once: try the first-generation kernel
            }
            if (!done && plan_cluster<T>(h, rows, w, &nc, &wc, &rpc, &smem)) {
                done = factor_subpanel_cluster<T>(h, A, ld, m, c0, nc, wc, rpc, smem, beta, V, ldv, c0 - k0, s0, Ts, W);
                if (done) w = wc;
                else h.opt.panel_cluster = 0;   // launch refused once: stay on the per-column path
            }
            const int64_t rest = (k0 + nb) - (c0 + w);  // remaining columns of this panel
            if (!done) {
                v_staged = false;
                factor_subpanel<T>(h, A, ld, m, c0, w, beta, scratch);
                if (rest > 0) {
                    copy_v<T>(h, A, ld, c0, c0, rows, w, beta, V, ldv);
                    build_t<T>(h, V, ldv, rows, w, G, Ts, W);
                }
            }
            if (rest > 0) {
                const T *Vs = done ? V + (c0 - k0) + (int64_t)s0 * ldv : V;
                apply_block_reflector<T>(h, Vs, ldv, rows, w, Ts, W, /*trans_t=*/1, A + c0 + (c0 + w) * ld, ld, rest, W1p, W2p);
            }
            s0 += w;
        }
        if (n - (k0 + nb) > 0) {
            const int64_t rows = m - k0;
            if (!v_staged) copy_v<T>(h, A, ld, k0, k0, rows, nb, beta, V, ldv);
            build_t<T>(h, V, ldv, rows, nb, G, Tm, NB);
            if (Vt) transpose<T>(h, V, rows, nb, ldv, Vt, NB);
        }
    };
    auto factor_panel = [&](int64_t k0, int nb, T *V, T *Tm, T *Vt, T *VT) {
        factor_panel_raw(k0, nb, V, Tm, Vt);
        if (VT && n - (k0 + nb) > 0) gemm<T>(h, 0, 0, m - k0, nb, nb, T(1), V, ldv, Tm, NB, T(0), VT, ldv);   // VT = V T
        finish_panel(k0, nb);
    };

    cudaStream_t sm = h.stream, sp = h.aux_stream;
    const bool la = h.opt.lookahead && sp != nullptr && !h.prof_on && n >= 4 * NB;
    // Debug option qr_trace (like chol_trace): event time stamps of the look-ahead pipeline, one line per panel on stderr.
    struct Mark { cudaEvent_t e; int panel; int what; };
    std::vector<Mark> marks;
    const bool trace = h.opt.qr_trace != 0 && la;
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
    factor_panel(0, (int)std::min<int64_t>(NB, n), Vbuf[0], Tbuf[0], Vtbuf[0], VTbuf[0]);
    int cur = 0;
    int pi = -1;
    for (int64_t k0 = 0; k0 < n; k0 += NB, cur ^= 1) {
        ++pi;
        const int nb = (int)std::min<int64_t>(NB, n - k0);
        const int64_t trail = n - (k0 + nb);
        if (trail <= 0) break;
        const int64_t rows = m - k0;
        const int nbn = (int)std::min<int64_t>(NB, trail);
        T *C = A + k0 + (k0 + nb) * ld;
        if (la) {
            // trailing update of the NEXT panel's columns first, then factor that panel on the side
            // stream while the rest of the trailing matrix is updated here
            mark(sm, pi, 0);
            apply_block_reflector<T>(h, Vbuf[cur], ldv, rows, nb, Tbuf[cur], NB, 1, C, ld, nbn, W1, W2, Vtbuf[cur], NB, VTbuf[cur]);
            mark(sm, pi, 1);
            LFB_CUDA(cudaEventRecord(h.ev[2], sm));
            LFB_CUDA(cudaStreamWaitEvent(sp, h.ev[2], 0));
            // the rest of the update is queued BEFORE the panel: the Cholesky-QR panel ends its guard with a host
            // synchronisation of the side stream, and the main stream must already hold its work by then
            if (trail > nbn)
                apply_block_reflector<T>(h, Vbuf[cur], ldv, rows, nb, Tbuf[cur], NB, 1, C + (int64_t)nbn * ld, ld, trail - nbn, W1, W2, Vtbuf[cur], NB, VTbuf[cur]);
            mark(sm, pi, 4);
            h.stream = sp;
            try {
                mark(sp, pi, 2);
                factor_panel(k0 + nb, nbn, Vbuf[cur ^ 1], Tbuf[cur ^ 1], Vtbuf[cur ^ 1], VTbuf[cur ^ 1]);
                mark(sp, pi, 3);
            } catch (...) {
                h.stream = sm;
                throw;
            }
            LFB_CUDA(cudaEventRecord(h.ev[3], sp));
            h.stream = sm;
            LFB_CUDA(cudaStreamWaitEvent(sm, h.ev[3], 0));
        } else {
            apply_block_reflector<T>(h, Vbuf[cur], ldv, rows, nb, Tbuf[cur], NB, 1, C, ld, trail, W1, W2, Vtbuf[cur], NB, VTbuf[cur]);
            factor_panel(k0 + nb, nbn, Vbuf[cur ^ 1], Tbuf[cur ^ 1], Vtbuf[cur ^ 1], VTbuf[cur ^ 1]);
        }
    }
    if (trace) {
        LFB_CUDA(cudaStreamSynchronize(sm));
        LFB_CUDA(cudaStreamSynchronize(sp));
        std::map<int, std::array<float, 5>> rowsT;
        for (auto &mk : marks) {
            float t = 0.f;
            cudaEventElapsedTime(&t, t_base, mk.e);
            rowsT[mk.panel][mk.what] = t;
            cudaEventDestroy(mk.e);
        }
        cudaEventDestroy(t_base);
        fprintf(stderr, "qr_trace m=%lld n=%lld nb=%d (ms from start): panel | next-panel columns start end | side panel start end | rest of update end\n",
                (long long)m, (long long)n, NB);
        for (auto &kv : rowsT)
            fprintf(stderr, "  %3d | %8.3f %8.3f | %8.3f %8.3f | %8.3f\n", kv.first, kv.second[0], kv.second[1], kv.second[2], kv.second[3], kv.second[4]);
    }
}

template <typename T>
void qr_factor(lfb_handle &h, T *A, int64_t m, int64_t n, int64_t ld, T *diag) {
    if (n <= 0) return;
    DevBuf<T> beta(h, n), psign(h, n);
    if (h.qr_panel_hook) {      // per-panel sign pass: block columns are final as soon as their panel is
        DevBuf<T> carry(h, 1);
        fill<T>(h, carry.get(), 1, 1, 1, T(1), T(1));
        const QrFinish<T> fin{psign.get(), diag, carry.get()};
        qr_factor_std<T>(h, A, m, n, ld, beta, &fin);
        return;
    }
    qr_factor_std<T>(h, A, m, n, ld, beta);
    sign_scan_kernel<T><<<1, 32, 0, h.stream>>>(beta, n, psign, diag);
    LFB_LAUNCH_CHECK(h);
    dim3 grid((unsigned)cdiv(m, 256), ycap(n));
    sign_fix_kernel<T><<<grid, 256, 0, h.stream>>>(A, ld, m, n, psign);
    LFB_LAUNCH_CHECK(h);
}

// R-only local QR for TSQR: R (n x n, upper, diag >= 0).  A is overwritten.
template <typename T>
__global__ void extract_r_kernel(const T *__restrict__ A, int64_t ld, int64_t n, const T *__restrict__ diag, T *__restrict__ R, int64_t ldr) {
    int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r >= n) return;
    for (int64_t c = blockIdx.y; c < n; c += gridDim.y) {
        T v = T(0);
        if (r < c) v = A[r + c * ld];
        else if (r == c) v = diag[r] < T(0) ? -diag[r] : diag[r];   // qr.rs:96
        R[r + c * ldr] = v;
    }
}

// Local stage of TSQR.  A tall block is cut into row chunks that fit the cluster panel kernel
// (tsqr_chunk rows); the chunks are factored CONCURRENTLY on a pool of sub-handles (each chunk's
// panel kernel occupies one 16-SM cluster, so ~8 chunks fill the GPU), their R factors are stacked
// and the stack is reduced the same way (a tree inside the GPU).  R of the stack == R of the block.
template <typename T>
static void tsqr_local_chunks(lfb_handle &h, T *A, int64_t rows, int64_t cols, int64_t ld, T *R, int64_t ldr, int64_t CH) {
    const int64_t nch = rows / CH;   // the last chunk absorbs the remainder (< 2 CH rows)
    const int64_t lds = nch * cols;
    DevBuf<T> Rstack(h, (size_t)lds * cols);
    const int NS = (int)std::min<int64_t>(std::max<int64_t>(h.opt.tsqr_streams, 1), nch);
    lfb_ensure_subs(h, NS);
    LFB_CUDA(cudaEventRecord(h.ev[4], h.stream));
    for (int s = 0; s < NS; ++s) LFB_CUDA(cudaStreamWaitEvent(h.subs[s]->stream, h.ev[4], 0));
    for (int64_t i = 0; i < nch; ++i) {
        lfb_handle &sub = *h.subs[i % NS];
        const int64_t r0 = i * CH, nr = (i == nch - 1) ? rows - r0 : CH;
        DevBuf<T> diag(sub, cols);
        qr_factor<T>(sub, A + r0, nr, cols, ld, diag);
        dim3 grid((unsigned)cdiv(cols, 256), ycap(cols));
        extract_r_kernel<T><<<grid, 256, 0, sub.stream>>>(A + r0, ld, cols, diag, Rstack.get() + i * cols, lds);
        LFB_LAUNCH_CHECK(sub);
    }
    for (int s = 0; s < NS; ++s) {
        LFB_CUDA(cudaEventRecord(h.subs[s]->ev[5], h.subs[s]->stream));
        LFB_CUDA(cudaStreamWaitEvent(h.stream, h.subs[s]->ev[5], 0));
        h.launches += h.subs[s]->launches;
        h.subs[s]->launches = 0;
    }
    const bool prev = h.in_capture;   // the stacked R factors are an internal buffer: no graph cache entry for them
    h.in_capture = true;
    try {
        tsqr_local_r<T>(h, Rstack, lds, cols, lds, R, ldr);
    } catch (...) {
        h.in_capture = prev;
        throw;
    }
    h.in_capture = prev;
}

template <typename T>
void tsqr_local_r(lfb_handle &h, T *A, int64_t rows, int64_t cols, int64_t ld, T *R, int64_t ldr) {
    if (cols <= 0) return;
    // Tall and well conditioned: the Cholesky-QR leaf (cholqr.cu) -- one Gram GEMM over all the rows + an n x n Cholesky.
    // Guarded; when it declines, A is untouched and the Householder leaf below runs as before.
    if (!h.is_sub && rows >= 4 * cols && cols <= 512 && cholqr_factor<T>(h, A, rows, cols, ld, R, ldr, (T *)nullptr, 0)) return;
    const int64_t CH = std::max<int64_t>(h.opt.tsqr_chunk, 2 * cols);
    if (rows < 2 * CH || h.is_sub) {
        DevBuf<T> diag(h, cols);
        qr_factor<T>(h, A, rows, cols, ld, diag);
        dim3 grid((unsigned)cdiv(cols, 256), ycap(cols));
        extract_r_kernel<T><<<grid, 256, 0, h.stream>>>(A, ld, cols, diag, R, ldr);
        LFB_LAUNCH_CHECK(h);
        return;
    }
    cudaStream_t user = h.stream;
    if (!h.opt.tsqr_graph || h.in_capture || h.prof_on) {
        tsqr_local_chunks<T>(h, A, rows, cols, ld, R, ldr, CH);
        return;
    }
    // Option tsqr_graph: the chunked stage is thousands of short launches on several streams; the second call with
    // the same buffers and shape captures it into a CUDA graph and later calls replay it (one launch from the
    // host).  Measured on B200: 145 ms replayed vs 147 ms launched -- the stage is bound by the panel cluster
    // kernels and the skinny GEMMs on the GPU, not by the host -- so the option is off by default.  A NULL
    // (legacy) user stream cannot be captured: the graph then runs on the handle's own stream, fenced against
    // the legacy stream on both sides.
    lfb_handle::GraphEntry *ent = nullptr;
    for (auto &g : h.graphs)
        if (g.a == A && g.r == R && g.rows == rows && g.cols == cols && g.ld == ld && g.ldr == ldr && g.chunk == CH &&
            g.streams == h.opt.tsqr_streams && g.elem == sizeof(T))
            ent = &g;
    if (!ent) {
        h.graphs.push_back({A, R, rows, cols, ld, ldr, CH, h.opt.tsqr_streams, sizeof(T), nullptr, 0});
        tsqr_local_chunks<T>(h, A, rows, cols, ld, R, ldr, CH);
        return;
    }
    cudaStream_t cs = user ? user : h.own_stream;
    if (!user) {
        if (!h.ev_graph[0]) {
            LFB_CUDA(cudaEventCreateWithFlags(&h.ev_graph[0], cudaEventDisableTiming));
            LFB_CUDA(cudaEventCreateWithFlags(&h.ev_graph[1], cudaEventDisableTiming));
        }
        LFB_CUDA(cudaEventRecord(h.ev_graph[0], user));
        LFB_CUDA(cudaStreamWaitEvent(cs, h.ev_graph[0], 0));
    }
    if (!ent->exec) {
        const int64_t l0 = h.launches;
        cudaGraph_t graph = nullptr;
        h.stream = cs;
        h.in_capture = true;
        LFB_CUDA(cudaStreamBeginCapture(cs, cudaStreamCaptureModeRelaxed));
        try {
            tsqr_local_chunks<T>(h, A, rows, cols, ld, R, ldr, CH);
        } catch (...) {
            cudaStreamEndCapture(cs, &graph);
            if (graph) cudaGraphDestroy(graph);
            cudaGetLastError();
            h.stream = user;
            h.in_capture = false;
            throw;
        }
        h.stream = user;
        h.in_capture = false;
        LFB_CUDA(cudaStreamEndCapture(cs, &graph));
        cudaGraphExec_t exec = nullptr;
        cudaError_t ge = cudaGraphInstantiate(&exec, graph, 0);
        cudaGraphDestroy(graph);
        if (ge != cudaSuccess) {
            cudaGetLastError();
            throw CudaError(LFB_ERR_CUDA, std::string("cudaGraphInstantiate failed: ") + cudaGetErrorString(ge));
        }
        ent->exec = exec;
        ent->launches = h.launches - l0;
    } else {
        h.launches += ent->launches;
    }
    LFB_CUDA(cudaGraphLaunch(ent->exec, cs));
    if (!user) {
        LFB_CUDA(cudaEventRecord(h.ev_graph[1], cs));
        LFB_CUDA(cudaStreamWaitEvent(user, h.ev_graph[1], 0));
    }
}


// TSQR with the orthogonal factor kept (the first half of making tall-skinny results fit
// QRDecomp, qr.rs:68-73).  On return A holds the EXPLICIT thin Q (rows x cols) of the block and R (cols x cols,
// upper, diag >= 0, qr.rs:96) its triangular factor, A_in = Q R.  Same chunking as tsqr_local_r: every chunk is
// factored to the reference's compact form, its thin Q_i is assembled (householder.rs:68-93) into the scratch
// Wk (rows x cols, leading dimension ldw), the stacked R_i are reduced recursively to (Qs, R), and
// Q[chunk i] = Q_i Qs[i*cols .. (i+1)*cols, :] is written back over the chunk by one tensor-core GEMM per chunk.
template <typename T>
void tsqr_explicit_q(lfb_handle &h, T *A, int64_t rows, int64_t cols, int64_t ld, T *Wk, int64_t ldw, T *R, int64_t ldr) {
    if (cols <= 0 || rows <= 0) return;
    if (!h.is_sub && rows >= 4 * cols && cols <= 512) {
        // Cholesky-QR leaf: R and R^-1 from the Gram matrix, Q = A R^-1 as ONE tensor-core GEMM (K = n) through the scratch
        const int64_t ldi = round_up(cols, 2);
        DevBuf<T> Rinv(h, (size_t)ldi * cols);
        if (cholqr_factor<T>(h, A, rows, cols, ld, R, ldr, Rinv.get(), ldi)) {
            gemm<T>(h, 0, 0, rows, cols, cols, T(1), A, ld, Rinv.get(), ldi, T(0), Wk, ldw);
            copy2d<T>(h, Wk, ldw, A, ld, rows, cols);
            return;
        }
    }
    const int64_t CH = std::max<int64_t>(h.opt.tsqr_chunk, 2 * cols);
    if (rows < 2 * CH || h.is_sub) {
        DevBuf<T> diag(h, cols);
        qr_factor<T>(h, A, rows, cols, ld, diag);
        dim3 grid((unsigned)cdiv(cols, 256), ycap(cols));
        extract_r_kernel<T><<<grid, 256, 0, h.stream>>>(A, ld, cols, diag, R, ldr);
        LFB_LAUNCH_CHECK(h);
        assemble_q<T>(h, A, rows, cols, ld, 0, diag, Wk, ldw);
        copy2d<T>(h, Wk, ldw, A, ld, rows, cols);
        return;
    }
    const int64_t nch = rows / CH;   // the last chunk absorbs the remainder (< 2 CH rows)
    const int64_t srows = nch * cols, lds = round_up(srows, 2);   // even leading dimension: 16-byte aligned columns
    DevBuf<T> Rstack(h, (size_t)lds * cols), Ws(h, (size_t)lds * cols);
    const int NS = (int)std::min<int64_t>(std::max<int64_t>(h.opt.tsqr_streams, 1), nch);
    lfb_ensure_subs(h, NS);
    LFB_CUDA(cudaEventRecord(h.ev[4], h.stream));
    for (int s = 0; s < NS; ++s) LFB_CUDA(cudaStreamWaitEvent(h.subs[s]->stream, h.ev[4], 0));
    for (int64_t i = 0; i < nch; ++i) {
        lfb_handle &sub = *h.subs[i % NS];
        const int64_t r0 = i * CH, nr = (i == nch - 1) ? rows - r0 : CH;
        DevBuf<T> diag(sub, cols);
        qr_factor<T>(sub, A + r0, nr, cols, ld, diag);
        dim3 grid((unsigned)cdiv(cols, 256), ycap(cols));
        extract_r_kernel<T><<<grid, 256, 0, sub.stream>>>(A + r0, ld, cols, diag, Rstack.get() + i * cols, lds);
        LFB_LAUNCH_CHECK(sub);
        assemble_q<T>(sub, A + r0, nr, cols, ld, 0, diag, Wk + r0, ldw);
    }
    for (int s = 0; s < NS; ++s) {
        LFB_CUDA(cudaEventRecord(h.subs[s]->ev[5], h.subs[s]->stream));
        LFB_CUDA(cudaStreamWaitEvent(h.stream, h.subs[s]->ev[5], 0));
        h.launches += h.subs[s]->launches;
        h.subs[s]->launches = 0;
    }
    tsqr_explicit_q<T>(h, Rstack, srows, cols, lds, Ws, lds, R, ldr);   // Rstack <- Qs
    for (int64_t i = 0; i < nch; ++i) {
        const int64_t r0 = i * CH, nr = (i == nch - 1) ? rows - r0 : CH;
        gemm<T>(h, 0, 0, nr, cols, cols, T(1), Wk + r0, ldw, Rstack.get() + i * cols, lds, T(0), A + r0, ld);
    }
}

// Q <- Q Qs in place (Q rows x cols, Qs cols x cols): the second level of a multi-GPU TSQR, where Qs is this rank's
// block of the explicit Q of the stacked R factors.  Row chunks go through a scratch so the GEMM never aliases.
template <typename T>
void tsqr_apply_q(lfb_handle &h, T *Q, int64_t rows, int64_t cols, int64_t ld, const T *Qs, int64_t ldqs) {
    if (rows <= 0 || cols <= 0) return;
    const int64_t CH = std::min<int64_t>(rows, 65536);
    const int64_t ldt = round_up(CH, 2);
    DevBuf<T> tmp(h, (size_t)ldt * cols);
    for (int64_t r0 = 0; r0 < rows; r0 += CH) {
        const int64_t nr = std::min(CH, rows - r0);
        copy2d<T>(h, Q + r0, ld, tmp, ldt, nr, cols);
        gemm<T>(h, 0, 0, nr, cols, cols, T(1), tmp, ldt, Qs, ldqs, T(0), Q + r0, ld);
    }
}

template <typename T>
void assemble_q(lfb_handle &h, const T *M, int64_t rows, int64_t cols, int64_t ld, int64_t shift, const T *signs, T *Q,
                int64_t ldq) {
    const int64_t dim = std::min(rows, cols);
    if (rows <= 0 || dim <= 0) return;
    fill<T>(h, Q, rows, dim, ldq, T(0), T(1));
    const int64_t nref = dim - shift;
    if (nref <= 0) return;
    const int NB = 128;
    const int64_t ldv = round_up(rows, 2);
    DevBuf<T> V(h, (size_t)ldv * NB), Tm(h, (size_t)NB * NB), G(h, (size_t)NB * NB);
    DevBuf<T> W1(h, (size_t)NB * dim), W2(h, (size_t)NB * dim), cum(h, nref);
    const int64_t npanels = cdiv(nref, NB);
    for (int64_t p = npanels - 1; p >= 0; --p) {
        const int64_t i0 = p * NB;
        const int nb = (int)std::min<int64_t>(NB, nref - i0);
        const int64_t r0 = i0 + shift;
        const int64_t prow = rows - r0;
        copy_v<T>(h, M, ld, r0, i0, prow, nb, (const T *)nullptr, V, ldv);
        build_t<T>(h, V, ldv, prow, nb, G, Tm, NB);
        // householder.rs:87  res[i+shift.., i..]: only columns >= i0 are touched
        apply_block_reflector<T>(h, V, ldv, prow, nb, Tm, NB, /*trans_t=*/0, Q + r0 + i0 * ldq, ldq, dim - i0, W1, W2);
    }
    sign_cumprod_kernel<T><<<1, 32, 0, h.stream>>>(signs, nref, cum);
    LFB_LAUNCH_CHECK(h);
    dim3 grid((unsigned)cdiv(rows, 256), ycap(dim));
    scale_cols_kernel<T><<<grid, 256, 0, h.stream>>>(Q, ldq, rows, dim, shift, nref, cum);
    LFB_LAUNCH_CHECK(h);
}

template <typename T>
void qt_mul(lfb_handle &h, const T *QR, int64_t rows, int64_t cols, int64_t ld, const T *diag, T *B, int64_t bcols,
            int64_t ldb) {
    if (cols <= 0 || bcols <= 0 || rows <= 0) return;
    const int NB = 128;
    const int64_t ldv = round_up(rows, 2);
    DevBuf<T> V(h, (size_t)ldv * NB), Tm(h, (size_t)NB * NB), G(h, (size_t)NB * NB);
    DevBuf<T> W1(h, (size_t)NB * bcols), W2(h, (size_t)NB * bcols), cum(h, cols);
    for (int64_t i0 = 0; i0 < cols; i0 += NB) {
        const int nb = (int)std::min<int64_t>(NB, cols - i0);
        const int64_t prow = rows - i0;
        copy_v<T>(h, QR, ld, i0, i0, prow, nb, (const T *)nullptr, V, ldv);
        build_t<T>(h, V, ldv, prow, nb, G, Tm, NB);
        apply_block_reflector<T>(h, V, ldv, prow, nb, Tm, NB, /*trans_t=*/1, B + i0, ldb, bcols, W1, W2);
    }
    sign_cumprod_kernel<T><<<1, 32, 0, h.stream>>>(diag, cols, cum);
    LFB_LAUNCH_CHECK(h);
    dim3 grid((unsigned)cdiv(rows, 256), ycap(bcols));
    scale_rows_kernel<T><<<grid, 256, 0, h.stream>>>(B, ldb, rows, bcols, cols, cum);
    LFB_LAUNCH_CHECK(h);
}

#define INST(T)                                                                                                   \
    template void qr_factor<T>(lfb_handle &, T *, int64_t, int64_t, int64_t, T *);                                \
    template void tsqr_local_r<T>(lfb_handle &, T *, int64_t, int64_t, int64_t, T *, int64_t);                    \
    template void tsqr_explicit_q<T>(lfb_handle &, T *, int64_t, int64_t, int64_t, T *, int64_t, T *, int64_t);   \
    template void tsqr_apply_q<T>(lfb_handle &, T *, int64_t, int64_t, int64_t, const T *, int64_t);              \
    template void assemble_q<T>(lfb_handle &, const T *, int64_t, int64_t, int64_t, int64_t, const T *, T *, int64_t); \
    template void qt_mul<T>(lfb_handle &, const T *, int64_t, int64_t, int64_t, const T *, T *, int64_t, int64_t);
INST(float)
INST(double)
#undef INST

}  // namespace lfb
