// Device-side helpers shared by the panel ("head") and streaming (SYMV / GEMV) kernels.
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include <cstdint>

namespace lfb {
namespace dev {

namespace cg = cooperative_groups;

template <typename T> __device__ __forceinline__ T t_sqrt(T x);
template <> __device__ __forceinline__ double t_sqrt<double>(double x) { return sqrt(x); }
template <> __device__ __forceinline__ float t_sqrt<float>(float x) { return sqrtf(x); }
template <typename T> __device__ __forceinline__ T t_abs(T x) { return x < T(0) ? -x : x; }
template <typename T> __device__ __forceinline__ T t_signum(T x) { return signbit(x) ? T(-1) : T(1); }


template <typename T> struct Vec2;
template <> struct Vec2<double> { using type = double2; };
template <> struct Vec2<float> { using type = float2; };


// c[k] holds this lane's partial sum for column k; afterwards c[0] of lane l is the warp total of
// column l (butterfly that halves the live values each step: 16+8+4+2+1 shuffles).
template <typename T>
__device__ __forceinline__ void warp_transpose_reduce(T (&c)[32], int lane) {
#pragma unroll
    for (int s = 0; s < 5; ++s) {
        const int o = 16 >> s;
        const bool up = (lane & o) != 0;
#pragma unroll
        for (int k = 0; k < o; ++k) {
            const T send = up ? c[k] : c[k + o];
            const T keep = up ? c[k + o] : c[k];
            c[k] = keep + __shfl_xor_sync(0xffffffffu, send, o);
        }
    }
}


// 16 values per lane -> lane l (and l ^ 16) holds the warp total of value l & 15 in c[0].
template <typename T>
__device__ __forceinline__ void warp_transpose_reduce16(T (&c)[16], int lane) {
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        const int o = 8 >> s;
        const bool up = (lane & o) != 0;
#pragma unroll
        for (int k = 0; k < o; ++k) {
            const T send = up ? c[k] : c[k + o];
            const T keep = up ? c[k + o] : c[k];
            c[k] = keep + __shfl_xor_sync(0xffffffffu, send, o);
        }
    }
    c[0] += __shfl_xor_sync(0xffffffffu, c[0], 16);
}


template <int BYTES>
__device__ __forceinline__ void cp_async(void *smem_dst, const void *gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    if constexpr (BYTES == 16) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
    else if constexpr (BYTES == 8) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
    else asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }


// Sum of (a, b2) over every thread of every CTA of the cluster.  `slot` is this CTA's 2-element
// mailbox (used once per kernel), read by all CTAs through distributed shared memory.
template <typename T, int NT>
__device__ __forceinline__ void cluster_sum2(cg::cluster_group &cl, T a, T b2, T *slot, T (*sred)[2], T *res, T &oa, T &ob) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b2 += __shfl_xor_sync(0xffffffffu, b2, o);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { sred[warp][0] = a; sred[warp][1] = b2; }
    __syncthreads();
    if (threadIdx.x == 0) {
        T s0 = T(0), s1 = T(0);
        for (int w = 0; w < NT / 32; ++w) { s0 += sred[w][0]; s1 += sred[w][1]; }
        slot[0] = s0; slot[1] = s1;
    }
    cl.sync();
    if (warp == 0) {
        const int nc = (int)cl.num_blocks();
        T s0 = T(0), s1 = T(0);
        if (lane < nc) {
            const T *rs = cl.map_shared_rank(slot, lane);
            s0 = rs[0]; s1 = rs[1];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            s0 += __shfl_xor_sync(0xffffffffu, s0, o);
            s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        }
        if (lane == 0) { res[0] = s0; res[1] = s1; }
    }
    __syncthreads();
    oa = res[0]; ob = res[1];
}


}  // namespace dev
}  // namespace lfb
