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
#include "common.cuh"

namespace lfb {
namespace {

constexpr int TB = 32;     // reflectors per panel
constexpr int NTH = 256;   // threads per CTA
constexpr int RPB = 512;   // rows per CTA

template <typename T> __device__ __forceinline__ T t_sqrt(T x);
template <> __device__ __forceinline__ double t_sqrt<double>(double x) { return sqrt(x); }
template <> __device__ __forceinline__ float t_sqrt<float>(float x) { return sqrtf(x); }
template <typename T> __device__ __forceinline__ T t_abs(T x) { return x < T(0) ? -x : x; }
template <typename T> __device__ __forceinline__ T t_signum(T x) { return signbit(x) ? T(-1) : T(1); }

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

}  // namespace

template <typename T>
void sym_tridiagonal(lfb_handle &h, T *A, int64_t n, int64_t ld, T *off) {
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

template void sym_tridiagonal<float>(lfb_handle &, float *, int64_t, int64_t, float *);
template void sym_tridiagonal<double>(lfb_handle &, double *, int64_t, int64_t, double *);

}  // namespace lfb
