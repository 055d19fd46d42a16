#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;
__global__ void k_sync(int iters, float* out) {
    cg::cluster_group cluster = cg::this_cluster();
    float acc = threadIdx.x;
    for (int i = 0; i < iters; ++i) { cluster.sync(); acc += 1.f; }
    if (acc == 12345.f) out[0] = acc;
}
__global__ void k_push(int iters, float* out) {
    cg::cluster_group cluster = cg::this_cluster();
    __shared__ double inbox[2][16][32];
    int nc = cluster.num_blocks(), b = cluster.block_rank();
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double acc = 0;
    for (int i = 0; i < iters; ++i) {
        int par = i & 1;
        for (int dst = warp; dst < nc; dst += 8) cluster.map_shared_rank(&inbox[0][0][0], dst)[(par * 16 + b) * 32 + lane] = acc + i;
        cluster.sync();
        double q = 0;
        for (int bb = 0; bb < 16; ++bb) q += inbox[par][bb][lane];
        acc += q * 1e-9;
    }
    if (acc == 12345.0) out[0] = acc;
}
__global__ void k_bsync(int iters, float* out) {
    float acc = threadIdx.x;
    for (int i = 0; i < iters; ++i) { __syncthreads(); acc += 1.f; }
    if (acc == 12345.f) out[0] = acc;
}
template <typename K> float run(K kern, int nc, int threads, int iters) {
    float* d; cudaMalloc(&d, 4);
    cudaLaunchConfig_t cfg = {}; cfg.gridDim = dim3(nc); cfg.blockDim = dim3(threads);
    cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = nc; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaLaunchKernelEx(&cfg, kern, iters, d); cudaDeviceSynchronize();
    cudaEventRecord(e0); cudaLaunchKernelEx(&cfg, kern, iters, d); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); 
    cudaError_t e = cudaGetLastError(); if (e != cudaSuccess) printf("err %s\n", cudaGetErrorString(e));
    cudaFree(d); return ms * 1e3f / iters;
}
int main() {
    for (int nc : {1, 2, 4, 8, 16}) printf("cluster.sync  nc=%2d 256thr: %.3f us   512thr: %.3f us   push+sync+sum: %.3f us\n", nc, run(k_sync, nc, 256, 2000), run(k_sync, nc, 512, 2000), run(k_push, nc, 256, 2000));
    printf("__syncthreads 256thr: %.4f us\n", run(k_bsync, 1, 256, 20000));
    return 0;
}
