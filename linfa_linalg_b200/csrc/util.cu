// Layout utilities: tiled transposes (row-major ndarray view <-> the engine's column-major HBM
// layout), fills, 2-D copies, triangular masking (src/triangular.rs:37-53).
#include "common.cuh"

namespace lfb {
namespace {

// out (cols x rows, ldout) = in^T, in is rows x cols column-major with ldin.
template <typename T>
__global__ void transpose_kernel(const T *__restrict__ in, int64_t rows, int64_t cols, int64_t ldin,
                                 T *__restrict__ out, int64_t ldout) {
    __shared__ T tile[32][33];
    int64_t r0 = (int64_t)blockIdx.x * 32, c0 = (int64_t)blockIdx.y * 32;
    int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
#pragma unroll
    for (int i = 0; i < 32; i += 8) {
        int64_t r = r0 + tx, c = c0 + ty + i;
        if (r < rows && c < cols) tile[ty + i][tx] = in[r + c * ldin];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 32; i += 8) {
        int64_t c = c0 + tx, r = r0 + ty + i;  // out element (c, r)
        if (r < rows && c < cols) out[c + r * ldout] = tile[tx][ty + i];
    }
}

// In-place transpose of a square matrix: one CTA per tile pair (bi >= bj).
template <typename T>
__global__ void transpose_sq_kernel(T *a, int64_t n, int64_t ld, int nt) {
    __shared__ T t1[32][33], t2[32][33];
    // map linear block id -> (bi, bj) with bi >= bj
    int64_t id = blockIdx.x;
    int64_t bi = (int64_t)((sqrt(8.0 * (double)id + 1.0) - 1.0) * 0.5);
    while (bi * (bi + 1) / 2 > id) --bi;
    while ((bi + 1) * (bi + 2) / 2 <= id) ++bi;
    int64_t bj = id - bi * (bi + 1) / 2;
    int tx = threadIdx.x, ty = threadIdx.y;
    int64_t r0 = bi * 32, c0 = bj * 32;
#pragma unroll
    for (int i = 0; i < 32; i += 8) {
        int64_t r = r0 + tx, c = c0 + ty + i;
        if (r < n && c < n) t1[ty + i][tx] = a[r + c * ld];
        int64_t r2 = c0 + tx, c2 = r0 + ty + i;
        if (r2 < n && c2 < n) t2[ty + i][tx] = a[r2 + c2 * ld];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 32; i += 8) {
        // block (bj, bi) <- t1^T ; block (bi, bj) <- t2^T
        int64_t r = c0 + tx, c = r0 + ty + i;
        if (r < n && c < n) a[r + c * ld] = t1[tx][ty + i];
        if (bi != bj) {
            int64_t r2 = r0 + tx, c2 = c0 + ty + i;
            if (r2 < n && c2 < n) a[r2 + c2 * ld] = t2[tx][ty + i];
        }
    }
}

template <typename T>
__global__ void fill_kernel(T *p, int64_t rows, int64_t cols, int64_t ld, T offdiag, T diag) {
    int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r >= rows) return;
    for (int64_t c = blockIdx.y; c < cols; c += gridDim.y) p[r + c * ld] = (r == c) ? diag : offdiag;
}

template <typename T>
__global__ void copy2d_kernel(const T *__restrict__ in, int64_t ldin, T *__restrict__ out, int64_t ldout,
                              int64_t rows, int64_t cols) {
    int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r >= rows) return;
    for (int64_t c = blockIdx.y; c < cols; c += gridDim.y) out[r + c * ldout] = in[r + c * ldin];
}

template <typename T>
__global__ void tri_zero_kernel(T *a, int64_t n, int64_t ld, int keep_lower) {
    int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r >= n) return;
    for (int64_t c = blockIdx.y; c < n; c += gridDim.y) {
        bool zero = keep_lower ? (c > r) : (c < r);
        if (zero) a[r + c * ld] = T(0);
    }
}

inline unsigned ycap(int64_t n) { return (unsigned)(n < 1 ? 1 : (n < 65535 ? n : 65535)); }

}  // namespace

template <typename T>
void transpose(lfb_handle &h, const T *in, int64_t rows, int64_t cols, int64_t ldin, T *out, int64_t ldout) {
    if (rows <= 0 || cols <= 0) return;
    // grid.y is limited to 65535 blocks: chunk over columns
    const int64_t maxc = 65535LL * 32;
    for (int64_t c0 = 0; c0 < cols; c0 += maxc) {
        int64_t nc = cols - c0 < maxc ? cols - c0 : maxc;
        dim3 grid((unsigned)cdiv(rows, 32), (unsigned)cdiv(nc, 32));
        transpose_kernel<T><<<grid, dim3(32, 8), 0, h.stream>>>(in + c0 * ldin, rows, nc, ldin, out + c0, ldout);
        LFB_LAUNCH_CHECK(h);
    }
}

template <typename T>
void transpose_inplace_square(lfb_handle &h, T *a, int64_t n, int64_t ld) {
    if (n <= 1) return;
    int64_t nt = cdiv(n, 32);
    int64_t blocks = nt * (nt + 1) / 2;
    transpose_sq_kernel<T><<<(unsigned)blocks, dim3(32, 8), 0, h.stream>>>(a, n, ld, (int)nt);
    LFB_LAUNCH_CHECK(h);
}

template <typename T>
void fill(lfb_handle &h, T *p, int64_t rows, int64_t cols, int64_t ld, T offdiag, T diag) {
    if (rows <= 0 || cols <= 0) return;
    dim3 grid((unsigned)cdiv(rows, 256), ycap(cols));
    fill_kernel<T><<<grid, 256, 0, h.stream>>>(p, rows, cols, ld, offdiag, diag);
    LFB_LAUNCH_CHECK(h);
}

template <typename T>
void copy2d(lfb_handle &h, const T *in, int64_t ldin, T *out, int64_t ldout, int64_t rows, int64_t cols) {
    if (rows <= 0 || cols <= 0) return;
    dim3 grid((unsigned)cdiv(rows, 256), ycap(cols));
    copy2d_kernel<T><<<grid, 256, 0, h.stream>>>(in, ldin, out, ldout, rows, cols);
    LFB_LAUNCH_CHECK(h);
}

template <typename T>
void triangular_zero(lfb_handle &h, T *a, int64_t n, int64_t ld, int keep_lower) {
    if (n <= 1) return;
    dim3 grid((unsigned)cdiv(n, 256), ycap(n));
    tri_zero_kernel<T><<<grid, 256, 0, h.stream>>>(a, n, ld, keep_lower);
    LFB_LAUNCH_CHECK(h);
}

#define INST(T)                                                                                              \
    template void transpose<T>(lfb_handle &, const T *, int64_t, int64_t, int64_t, T *, int64_t);            \
    template void transpose_inplace_square<T>(lfb_handle &, T *, int64_t, int64_t);                          \
    template void fill<T>(lfb_handle &, T *, int64_t, int64_t, int64_t, T, T);                               \
    template void copy2d<T>(lfb_handle &, const T *, int64_t, T *, int64_t, int64_t, int64_t);               \
    template void triangular_zero<T>(lfb_handle &, T *, int64_t, int64_t, int);
INST(float)
INST(double)
#undef INST

}  // namespace lfb
