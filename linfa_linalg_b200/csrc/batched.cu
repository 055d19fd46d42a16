// Batched thin QR of many small row-major matrices (m, n <= 32): one warp per matrix, lane = column.
//
// Each matrix is src/qr.rs:38-41 (clear_column per column, householder.rs:34-51) executed literally,
// including the sign scaling, so the compact output is bit-for-bit in the reference's format.
// HBM-bound by design: a matrix is read once (coalesced 128-byte rows), lives in registers
// (lane j holds column j), and is written once.  The reflector is broadcast through a per-warp
// shared-memory slab (one LDS.128 per 4 elements instead of 32 shuffles).
#include "common.cuh"

namespace lfb {
namespace {

template <typename T> __device__ __forceinline__ T t_sqrt(T x);
template <> __device__ __forceinline__ double t_sqrt<double>(double x) { return sqrt(x); }
template <> __device__ __forceinline__ float t_sqrt<float>(float x) { return sqrtf(x); }
template <typename T> __device__ __forceinline__ T t_abs(T x) { return x < T(0) ? -x : x; }
template <typename T> __device__ __forceinline__ T t_signum(T x) { return signbit(x) ? T(-1) : T(1); }

constexpr int WPB = 8;  // warps (matrices) per CTA

template <typename T>
__global__ void __launch_bounds__(WPB * 32) qr_batched_kernel(T *__restrict__ A, int64_t batch, int m, int n,
                                                              T *__restrict__ diag) {
    __shared__ __align__(16) T sv[WPB][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t b = (int64_t)blockIdx.x * WPB + warp; b < batch; b += (int64_t)gridDim.x * WPB) {
        T *mat = A + b * (int64_t)m * n;
        T a[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) a[i] = (i < m && lane < n) ? mat[i * n + lane] : T(0);
        T mydiag = T(0);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            if (j < n) {
                // lane j: reflector of its own column, rows j.. (householder.rs:9-28)
                T s = T(0), d = T(1);
                int some = 0;
                if (lane == j) {
                    T nsq = T(0);
#pragma unroll
                    for (int i = j; i < 32; ++i) nsq += a[i] * a[i];
                    T nrm = t_sqrt(nsq);
                    T f = a[j];
                    s = t_signum(f) * nrm;
                    T newsq = (nsq + t_abs(f) * nrm) * T(2);
                    some = newsq != T(0);
                    d = t_sqrt(newsq);
                    if (some) {
                        a[j] = f + s;
#pragma unroll
                        for (int i = j; i < 32; ++i) a[i] = a[i] / d;
                        mydiag = -s;
                    }
#pragma unroll
                    for (int i = 0; i < 32; ++i) sv[warp][i] = (i >= j && i < m) ? a[i] : T(0);
                }
                __syncwarp();
                some = __shfl_sync(0xffffffffu, some, j);
                const T sg = t_signum(-__shfl_sync(0xffffffffu, s, j));   // signum of the returned pivot
                if (some && lane > j && lane < n) {
                    T dot = T(0);
#pragma unroll
                    for (int i = j; i < 32; ++i) dot += sv[warp][i] * a[i];
                    const T fac = T(-2) * dot;                             // reflection.rs:29
#pragma unroll
                    for (int i = j; i < 32; ++i) a[i] = sg * (a[i] + fac * sv[warp][i]);   // :30 + householder.rs:48
                }
                __syncwarp();
            }
        }
#pragma unroll
        for (int i = 0; i < 32; ++i)
            if (i < m && lane < n) mat[i * n + lane] = a[i];
        if (lane < n) diag[b * n + lane] = mydiag;
    }
}

}  // namespace

template <typename T>
void qr_batched(lfb_handle &h, T *A, int64_t batch, int64_t m, int64_t n, T *diag) {
    if (batch <= 0 || n <= 0) return;
    int64_t blocks = std::min<int64_t>(cdiv(batch, WPB), (int64_t)h.sm_count * 8);
    qr_batched_kernel<T><<<(unsigned)blocks, WPB * 32, 0, h.stream>>>(A, batch, (int)m, (int)n, diag);
    LFB_LAUNCH_CHECK(h);
}

template void qr_batched<float>(lfb_handle &, float *, int64_t, int64_t, int64_t, float *);
template void qr_batched<double>(lfb_handle &, double *, int64_t, int64_t, int64_t, double *);

}  // namespace lfb
