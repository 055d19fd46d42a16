// Golub-Kahan bidiagonalisation, first (BLAS-2) generation: fused GPU kernels, one reflector at a
// time, HBM-bound.  (The symmetric tridiagonalisation moved to tridiag.cu, blocked.)
//
// Replaces src/tridiagonal.rs:31-66 (sym_tridiagonal hot loop :40-60) and src/bidiagonal.rs:27-59
// (alternating clear_column / clear_row, householder.rs:34-63).  The arithmetic follows the
// reference literally (full-matrix H M H with p = 2 M v; sign-scaled one-sided reflections), so the
// stored reflectors, `off`, `d`, `e` carry the reference's signs directly.
#include "common.cuh"

namespace lfb {
namespace {

template <typename T> __device__ __forceinline__ T t_sqrt(T x);
template <> __device__ __forceinline__ double t_sqrt<double>(double x) { return sqrt(x); }
template <> __device__ __forceinline__ float t_sqrt<float>(float x) { return sqrtf(x); }
template <typename T> __device__ __forceinline__ T t_abs(T x) { return x < T(0) ? -x : x; }
template <typename T> __device__ __forceinline__ T t_signum(T x) { return signbit(x) ? T(-1) : T(1); }

template <typename T>
__device__ __forceinline__ T block_sum(T v, T *sred) {  // blockDim.x multiple of 32, <= 1024
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) sred[warp] = v;
    __syncthreads();
    T s = T(0);
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += sred[w];
    return s;
}

// State shared between the kernels of one reflector step.
template <typename T>
struct StepState {
    T sign;   // signum of the returned pivot (householder.rs:45)
    T some;   // 1 if a reflection is performed, 0 for `None`
};

// householder.rs:9-28 on a strided vector x (len L, stride inc), single CTA.
// Writes v in place, a contiguous copy to vc[0..L), the returned scalar (or 0) to *out, the step
// state, and zeroes y[0..ylen) (the accumulator of the following GEMV).
template <typename T>
__global__ void __launch_bounds__(1024) reflector_kernel(T *x, int64_t L, int64_t inc, T *vc, T *out,
                                                         StepState<T> *st, T *y, int64_t ylen) {
    __shared__ T sred[32];
    T part = T(0);
    for (int64_t i = threadIdx.x; i < L; i += blockDim.x) {
        T a = x[i * inc];
        part += a * a;
    }
    const T nsq = block_sum(part, sred);
    const T nrm = t_sqrt(nsq);
    const T f = x[0];
    const T s = t_signum(f) * nrm;
    const T newsq = (nsq + t_abs(f) * nrm) * T(2);
    const bool some = newsq != T(0);
    const T d = t_sqrt(newsq);
    __syncthreads();  // everyone has read x[0]
    for (int64_t i = threadIdx.x; i < L; i += blockDim.x) {
        T a = x[i * inc];
        if (some) {
            a = ((i == 0) ? a + s : a) / d;
            x[i * inc] = a;
        }
        vc[i] = a;
    }
    for (int64_t i = threadIdx.x; i < ylen; i += blockDim.x) y[i] = T(0);
    if (threadIdx.x == 0) {
        *out = some ? -s : T(0);
        st->sign = t_signum(-s);
        st->some = some ? T(1) : T(0);
    }
}

// y[r] += alpha * sum_{c in split} M[r,c] x[c]   (M rows x cols column-major; thread per row)
template <typename T>
__global__ void __launch_bounds__(128) gemv_n_kernel(const T *__restrict__ M, int64_t ld, int64_t rows, int64_t cols,
                                                     const T *__restrict__ x, T alpha, T *y, int64_t csplit,
                                                     const StepState<T> *st) {
    if (st->some == T(0)) return;
    int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int64_t c0 = blockIdx.y * csplit, c1 = min(cols, c0 + csplit);
    if (r >= rows) return;
    T acc0 = T(0), acc1 = T(0), acc2 = T(0), acc3 = T(0);
    const T *m = M + r;
    int64_t c = c0;
    for (; c + 3 < c1; c += 4) {
        acc0 += m[c * ld] * x[c];
        acc1 += m[(c + 1) * ld] * x[c + 1];
        acc2 += m[(c + 2) * ld] * x[c + 2];
        acc3 += m[(c + 3) * ld] * x[c + 3];
    }
    for (; c < c1; ++c) acc0 += m[c * ld] * x[c];
    atomicAdd(y + r, alpha * ((acc0 + acc1) + (acc2 + acc3)));
}

// y[c] = alpha * sum_r M[r,c] x[r]  (warp per column, coalesced along rows)
template <typename T>
__global__ void __launch_bounds__(256) gemv_t_kernel(const T *__restrict__ M, int64_t ld, int64_t rows, int64_t cols,
                                                     const T *__restrict__ x, T alpha, T *y, const StepState<T> *st) {
    if (st->some == T(0)) return;
    const int lane = threadIdx.x & 31;
    int64_t c = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
    if (c >= cols) return;
    const T *m = M + c * ld;
    T acc = T(0);
    for (int64_t r = lane; r < rows; r += 32) acc += m[r] * x[r];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) y[c] = alpha * acc;
}

// M[r,c] = sign * (M[r,c] - 2 a[r] b[c])     (reflection.rs:29-30 then householder.rs:48)
template <typename T>
__global__ void __launch_bounds__(256) ger_sign_kernel(T *M, int64_t ld, int64_t rows, int64_t cols,
                                                       const T *__restrict__ a, const T *__restrict__ b,
                                                       const StepState<T> *st) {
    if (st->some == T(0)) return;
    const T sign = st->sign;
    int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const T ar = T(-2) * a[r];
    for (int64_t c = blockIdx.y; c < cols; c += gridDim.y) {
        T v = M[r + c * ld] + ar * b[c];
        M[r + c * ld] = sign * v;
    }
}

// tridiagonal.rs:50-58: dot = v.p ; w = p - dot v   (single CTA)
template <typename T>
__global__ void __launch_bounds__(1024) tri_w_kernel(const T *__restrict__ v, const T *__restrict__ p, T *w, int64_t L,
                                                     const StepState<T> *st) {
    __shared__ T sred[32];
    if (st->some == T(0)) return;
    T part = T(0);
    for (int64_t i = threadIdx.x; i < L; i += blockDim.x) part += v[i] * p[i];
    const T dot = block_sum(part, sred);
    for (int64_t i = threadIdx.x; i < L; i += blockDim.x) w[i] = p[i] - dot * v[i];
}

// M -= w v^T + v w^T   ( == M - p v^T - v p^T + 2 dot v v^T, tridiagonal.rs:56-58 )
template <typename T>
__global__ void __launch_bounds__(256) syr2_kernel(T *M, int64_t ld, int64_t L, const T *__restrict__ v,
                                                   const T *__restrict__ w, const StepState<T> *st) {
    if (st->some == T(0)) return;
    int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r >= L) return;
    const T vr = v[r], wr = w[r];
    for (int64_t c = blockIdx.y; c < L; c += gridDim.y) M[r + c * ld] -= wr * v[c] + vr * w[c];
}

inline unsigned ycap(int64_t n, int64_t cap) { return (unsigned)(n < 1 ? 1 : (n < cap ? n : cap)); }

template <typename T>
void launch_gemv_n(lfb_handle &h, const T *M, int64_t ld, int64_t rows, int64_t cols, const T *x, T alpha, T *y,
                   const StepState<T> *st) {
    if (rows <= 0 || cols <= 0) return;
    int64_t rb = cdiv(rows, 128);
    int64_t splits = std::max<int64_t>(1, std::min<int64_t>(cdiv(16 * h.sm_count, rb), cdiv(cols, 64)));
    int64_t csplit = cdiv(cols, splits);
    dim3 grid((unsigned)rb, (unsigned)cdiv(cols, csplit));
    gemv_n_kernel<T><<<grid, 128, 0, h.stream>>>(M, ld, rows, cols, x, alpha, y, csplit, st);
    LFB_LAUNCH_CHECK(h);
}

}  // namespace

// One clear_column on the column-major matrix (householder.rs:34-51).
template <typename T>
static void clear_column_dev(lfb_handle &h, T *A, int64_t rows, int64_t cols, int64_t ld, int64_t icol, int64_t shift,
                             T *out, T *v, T *y, StepState<T> *st) {
    const int64_t r0 = icol + shift, L = rows - r0, wd = cols - icol - 1;
    if (L <= 0) {  // empty axis cannot happen for valid bidiagonal calls
        return;
    }
    T *x = A + r0 + icol * ld;
    reflector_kernel<T><<<1, 1024, 0, h.stream>>>(x, L, 1, v, out, st, y, 0);
    LFB_LAUNCH_CHECK(h);
    if (wd <= 0) return;
    T *Tt = A + r0 + (icol + 1) * ld;
    gemv_t_kernel<T><<<(unsigned)cdiv(wd, 8), 256, 0, h.stream>>>(Tt, ld, L, wd, v, T(1), y, st);   // y = Tt^T v
    LFB_LAUNCH_CHECK(h);
    dim3 grid((unsigned)cdiv(L, 256), ycap(wd, 2048));
    ger_sign_kernel<T><<<grid, 256, 0, h.stream>>>(Tt, ld, L, wd, v, y, st);
    LFB_LAUNCH_CHECK(h);
}

// One clear_row (householder.rs:57-63): clear_column on the transposed view.
template <typename T>
static void clear_row_dev(lfb_handle &h, T *A, int64_t rows, int64_t cols, int64_t ld, int64_t irow, int64_t shift,
                          T *out, T *v, T *y, StepState<T> *st) {
    const int64_t c0 = irow + shift, L = cols - c0, ht = rows - irow - 1;
    if (L <= 0) return;
    T *x = A + irow + c0 * ld;  // row irow, columns c0.. (stride ld)
    reflector_kernel<T><<<1, 1024, 0, h.stream>>>(x, L, ld, v, out, st, y, ht > 0 ? ht : 0);
    LFB_LAUNCH_CHECK(h);
    if (ht <= 0) return;
    T *Tt = A + (irow + 1) + c0 * ld;  // rows irow+1.., cols c0..
    launch_gemv_n<T>(h, Tt, ld, ht, L, v, T(1), y, st);   // y = Tt v
    dim3 grid((unsigned)cdiv(ht, 256), ycap(L, 2048));
    ger_sign_kernel<T><<<grid, 256, 0, h.stream>>>(Tt, ld, ht, L, y, v, st);   // Tt = sign (Tt - 2 y v^T)
    LFB_LAUNCH_CHECK(h);
}

template <typename T>
void bidiagonal_unblocked(lfb_handle &h, T *A, int64_t rows, int64_t cols, int64_t ld, T *d, T *e) {
    const int64_t md = std::min(rows, cols);
    if (md <= 0) return;
    const int64_t mx = std::max(rows, cols);
    DevBuf<T> v(h, mx), y(h, mx);
    DevBuf<StepState<T>> st(h, 1);
    if (rows >= cols) {  // bidiagonal.rs:38-44
        for (int64_t i = 0; i + 1 < md; ++i) {
            clear_column_dev<T>(h, A, rows, cols, ld, i, 0, d + i, v, y, st);
            clear_row_dev<T>(h, A, rows, cols, ld, i, 1, e + i, v, y, st);
        }
        clear_column_dev<T>(h, A, rows, cols, ld, md - 1, 0, d + md - 1, v, y, st);
    } else {             // :45-51
        for (int64_t i = 0; i + 1 < md; ++i) {
            clear_row_dev<T>(h, A, rows, cols, ld, i, 0, d + i, v, y, st);
            clear_column_dev<T>(h, A, rows, cols, ld, i, 1, e + i, v, y, st);
        }
        clear_row_dev<T>(h, A, rows, cols, ld, md - 1, 0, d + md - 1, v, y, st);
    }
}

#define INST(T)                                                                       \
    template void bidiagonal_unblocked<T>(lfb_handle &, T *, int64_t, int64_t, int64_t, T *, T *);
INST(float)
INST(double)
#undef INST

}  // namespace lfb
