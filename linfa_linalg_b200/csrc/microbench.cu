// FP64 pipe ceilings measured in the same run as the benchmark (bench.py reports roofline fractions
// against these AND against the nominal figure): a register-resident DFMA chain and a DMMA.8x8x4
// (mma.sync f64) chain, both with enough independent accumulators to saturate the pipe.
#include "common.cuh"

namespace lfb {
namespace {

__global__ void __launch_bounds__(256) dfma_kernel(double *out, int iters, double a, double b) {
    double acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i];
    if (s == 123.456) out[0] = s;
}

__global__ void __launch_bounds__(256) dmma_kernel(double *out, int iters, double a, double b) {
    double acc[16][2];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i][0] = acc[i][1] = threadIdx.x * 1e-3 + i;
    double fa = a + threadIdx.x * 1e-6, fb = b;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(acc[i][0]), "+d"(acc[i][1])
                         : "d"(fa), "d"(fb));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i][0] + acc[i][1];
    if (s == 123.456) out[0] = s;
}

// kind 2: the FP32 FFMA ceiling (the denominator the f32 trailing updates are reported against)
__global__ void __launch_bounds__(256) ffma_kernel(float *out, int iters, float a, float b) {
    float acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x * 1e-3f + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = fmaf(acc[i], a, b);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i];
    if (s == 123.456f) out[0] = s;
}

}  // namespace

double microbench_fp64(lfb_handle &h, int kind) {
    DevBuf<double> out(h, 1);
    const int iters = 4096;
    const int blocks = h.sm_count * 4;
    cudaEvent_t e0, e1;
    LFB_CUDA(cudaEventCreate(&e0));
    LFB_CUDA(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        LFB_CUDA(cudaEventRecord(e0, h.stream));
        if (kind == 0) dfma_kernel<<<blocks, 256, 0, h.stream>>>(out, iters, 0.999999, 1e-9);
        else if (kind == 2) ffma_kernel<<<blocks, 256, 0, h.stream>>>(reinterpret_cast<float *>(out.get()), iters * 4, 0.999999f, 1e-9f);
        else dmma_kernel<<<blocks, 256, 0, h.stream>>>(out, iters, 0.999999, 1e-9);
        LFB_LAUNCH_CHECK(h);
        LFB_CUDA(cudaEventRecord(e1, h.stream));
        LFB_CUDA(cudaEventSynchronize(e1));
        float ms = 0;
        LFB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    double flops;
    if (kind == 0) flops = 2.0 * 16 * (double)iters * 256.0 * blocks;            // 1 FMA per lane per op
    else if (kind == 2) flops = 2.0 * 16 * (double)iters * 4 * 256.0 * blocks;
    else flops = 2.0 * 256.0 * 16 * (double)iters * 8.0 * blocks;               // 8x8x4 MACs per warp-op, 8 warps/CTA
    return flops / (best * 1e-3) / 1e9;
}

}  // namespace lfb
