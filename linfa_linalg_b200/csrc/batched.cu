// Batched thin QR of many small row-major matrices (m, n <= 32): one warp per matrix.
//
// Each matrix is src/qr.rs:38-41 (clear_column per column, householder.rs:34-51).  HBM-bound by
// design: a matrix is read once (coalesced 128-byte rows), lives on chip, and is written once.
//
// Layout of the work (second generation; the first kept everything in 250 registers with 32 fully
// unrolled column steps and was instruction-fetch / latency bound at 7 % of the HBM roofline):
//   * lane t keeps COLUMN t in registers (static indexing only);
//   * per column step j the owner lane publishes its column through a per-warp shared slab, the
//     norm is then a 5-step shuffle reduction with lane = row, every lane computes the scalars,
//     the unit reflector v goes back through shared memory (one float per lane out, 16-byte
//     broadcast loads in) and lanes t > j do dot + axpy on their register column;
//   * rows above the current 8-row segment are skipped statically (4 code segments);
//   * the reference's per-column sign scaling (householder.rs:45-48) is applied once at the end:
//     R[i, t>i] *= P_i, v_j *= P_{j-1}, diag_j = P_{j-1} beta_j with P_j = sgn(beta_j).
#include "common.cuh"

namespace lfb {
namespace {

template <typename T> __device__ __forceinline__ T t_sqrt(T x);
template <> __device__ __forceinline__ double t_sqrt<double>(double x) { return sqrt(x); }
template <> __device__ __forceinline__ float t_sqrt<float>(float x) { return sqrtf(x); }
template <typename T> __device__ __forceinline__ T t_abs(T x) { return x < T(0) ? -x : x; }
template <typename T> __device__ __forceinline__ T t_signum(T x) { return signbit(x) ? T(-1) : T(1); }

// 1/sqrt(x): f32 = MUFU.RSQ; f64 = f32 seed + 2 Newton steps (full accuracy inside the f32 range)
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

template <typename T> constexpr int wpb() { return sizeof(T) == 8 ? 4 : 8; }   // warps (matrices in flight) per CTA
constexpr int VP = 36;    // pitch of the per-warp reflector store (16-byte aligned rows)

template <typename T, int RS>
__device__ __forceinline__ void column_steps(T (&a)[32], T *colbuf, T *vstore, int lane, int n, int jend, T &dg) {
    for (int j = RS; j < jend; ++j) {
        // 1. the owner lane publishes rows RS.. of its column
        if (lane == j) {   // 16-byte stores
            if constexpr (sizeof(T) == 4) {
#pragma unroll
                for (int r = RS; r < 32; r += 4)
                    *reinterpret_cast<float4 *>(colbuf + r) = make_float4(a[r], a[r + 1], a[r + 2], a[r + 3]);
            } else {
#pragma unroll
                for (int r = RS; r < 32; r += 2) *reinterpret_cast<double2 *>(colbuf + r) = make_double2(a[r], a[r + 1]);
            }
        }
        __syncwarp();
        const T x = colbuf[lane];                                   // lane = row
        const T xa = lane >= j ? x : T(0);
        T nsq = xa * xa;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) nsq += __shfl_xor_sync(0xffffffffu, nsq, o);
        const T f = __shfl_sync(0xffffffffu, x, j);
        // householder.rs:13-23; sqrt(x) = x * rsqrt(x) (+ one correction step) and the division by
        // sqrt(new_norm_sq) is a multiplication: the 32 column steps of a matrix are one latency chain
        const T rn = nsq > T(0) ? fast_rsqrt(nsq) : T(0);
        T nrm = nsq * rn;
        nrm = fma(T(0.5) * rn, fma(-nrm, nrm, nsq), nrm);
        const T s = t_signum(f) * nrm;                              // :16
        const T newsq = (nsq + t_abs(f) * nrm) * T(2);              // :19-20
        const bool some = newsq != T(0);                            // :22
        const T rd = some ? fast_rsqrt(newsq) : T(0);
        const T v = some ? ((lane == j ? x + s : xa) * rd) : T(0);  // :17,23 (zero above the pivot row)
        vstore[j * VP + lane] = some ? v : x;   // a `None` column is left untouched
        if (lane == j) dg = some ? -s : T(0);                       // :24/26
        __syncwarp();
        // 2./3. lanes right of the pivot column: dot + axpy on their register column (reflection.rs:29-30)
        if (some && lane > j && lane < n) {
            T vr[32 - RS];
            if constexpr (sizeof(T) == 4) {   // 16-byte broadcast loads
#pragma unroll
                for (int r = RS; r < 32; r += 4) {
                    const float4 q = *reinterpret_cast<const float4 *>(vstore + j * VP + r);
                    vr[r - RS] = q.x; vr[r + 1 - RS] = q.y; vr[r + 2 - RS] = q.z; vr[r + 3 - RS] = q.w;
                }
            } else {
#pragma unroll
                for (int r = RS; r < 32; r += 2) {
                    const double2 q = *reinterpret_cast<const double2 *>(vstore + j * VP + r);
                    vr[r - RS] = q.x; vr[r + 1 - RS] = q.y;
                }
            }
            T d0 = T(0), d1 = T(0), d2 = T(0), d3 = T(0);           // four chains: the dot is latency bound
#pragma unroll
            for (int r = RS; r < 32; r += 4) {
                d0 += vr[r - RS] * a[r];
                d1 += vr[r + 1 - RS] * a[r + 1];
                d2 += vr[r + 2 - RS] * a[r + 2];
                d3 += vr[r + 3 - RS] * a[r + 3];
            }
            const T fac = T(-2) * ((d0 + d1) + (d2 + d3));
#pragma unroll
            for (int r = RS; r < 32; ++r) a[r] += fac * vr[r - RS];
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(wpb<T>() * 32, 2) qr_batched_kernel(T *__restrict__ A, int64_t batch, int m, int n,
                                                                 T *__restrict__ diag) {
    constexpr int WPB = wpb<T>();
    __shared__ __align__(16) T s_col[WPB][32];
    __shared__ __align__(16) T s_v[WPB][32 * VP];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    T *colbuf = s_col[warp], *vstore = s_v[warp];
    for (int64_t b = (int64_t)blockIdx.x * WPB + warp; b < batch; b += (int64_t)gridDim.x * WPB) {
        T *mat = A + b * (int64_t)m * n;
        T a[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) a[i] = (i < m && lane < n) ? mat[i * n + lane] : T(0);
        T dg = T(0);
        column_steps<T, 0>(a, colbuf, vstore, lane, n, min(8, n), dg);
        if (n > 8) column_steps<T, 8>(a, colbuf, vstore, lane, n, min(16, n), dg);
        if (n > 16) column_steps<T, 16>(a, colbuf, vstore, lane, n, min(24, n), dg);
        if (n > 24) column_steps<T, 24>(a, colbuf, vstore, lane, n, min(32, n), dg);
        // sign convention of the reference, applied once: running sign P_r (every lane scans all pivots)
        T p = T(1), prevp = T(1);
#pragma unroll
        for (int r = 0; r < 32; ++r) {
            if (lane == r) prevp = p;                               // P_{lane-1}
            const T br = __shfl_sync(0xffffffffu, dg, r);
            if (br != T(0)) p = t_signum(br);
            if (r < lane) a[r] *= p;                                // R[r, lane] *= P_r
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            if (i < m && lane < n) {
                const T val = (i < lane) ? a[i] : prevp * vstore[lane * VP + i];
                mat[i * n + lane] = val;
            }
        }
        if (lane < n) diag[b * n + lane] = (dg != T(0)) ? prevp * dg : T(0);   // a `None` pivot is +0.0, never -0.0 (householder.rs:50)
        __syncwarp();
    }
}


// ---- f32, four matrices per warp ------------------------------------------------------------------
// ncu on the one-matrix-per-warp kernel: issue bound, 5.2k warp instructions per matrix of which only
// 1.3k are FMAs -- the per-column overhead (publish, reduce, scalars, broadcast) is paid by a whole
// warp for one matrix.  Here a matrix belongs to 8 lanes; lane g keeps columns g, g+8, g+16, g+24 in
// registers (interleaved so the triangular work stays balanced), so every overhead instruction serves
// four matrices and the FMA share rises to ~60 %.
template <int Q>   // column slot of the pivot column: j = 8*Q + jj, rows RS = 8*Q .. 31 are live
__device__ __forceinline__ void column_steps4(float (&a)[4][32], float *colbuf, float *vbuf, int g, int n, float (&dg)[4]) {
    constexpr int RS = 8 * Q;
    const unsigned gmask = 0xffu << (threadIdx.x & 24);   // the 8 lanes of this matrix
    for (int jj = 0; jj < 8; ++jj) {
        const int j = RS + jj;
        if (j >= n) break;   // n is matrix-uniform and warp-uniform
        if (g == jj) {
#pragma unroll
            for (int r = RS; r < 32; r += 4)
                *reinterpret_cast<float4 *>(colbuf + r) = make_float4(a[Q][r], a[Q][r + 1], a[Q][r + 2], a[Q][r + 3]);
        }
        __syncwarp();
        // norm over rows >= j: lane g sums rows g, g+8, g+16, g+24
        float nsq = 0.f, xs[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int r = g + 8 * t;
            xs[t] = (r >= RS) ? colbuf[r] : 0.f;
            if (r >= j) nsq += xs[t] * xs[t];
        }
        nsq += __shfl_xor_sync(gmask, nsq, 1);
        nsq += __shfl_xor_sync(gmask, nsq, 2);
        nsq += __shfl_xor_sync(gmask, nsq, 4);
        const float f = colbuf[j];
        const float rn = nsq > 0.f ? rsqrtf(nsq) : 0.f;
        float nrm = nsq * rn;                                        // householder.rs:13
        nrm = fmaf(0.5f * rn, fmaf(-nrm, nrm, nsq), nrm);
        const float s = (signbit(f) ? -1.f : 1.f) * nrm;             // :16
        const float newsq = (nsq + fabsf(f) * nrm) * 2.f;            // :19-20
        const bool some = newsq != 0.f;                              // :22
        const float rd = some ? rsqrtf(newsq) : 0.f;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int r = g + 8 * t;
            if (r >= RS) vbuf[r] = (some && r >= j) ? ((r == j ? xs[t] + s : xs[t]) * rd) : 0.f;   // :17,23
        }
        if (g == jj) dg[Q] = some ? -s : 0.f;                        // :24/26
        __syncwarp();
        if (some) {
            float vr[32 - RS];
#pragma unroll
            for (int r = RS; r < 32; r += 4) {
                const float4 q4 = *reinterpret_cast<const float4 *>(vbuf + r);
                vr[r - RS] = q4.x; vr[r + 1 - RS] = q4.y; vr[r + 2 - RS] = q4.z; vr[r + 3 - RS] = q4.w;
            }
#pragma unroll
            for (int qq = Q; qq < 4; ++qq) {
                const int c = 8 * qq + g;
                if (c == j) {                    // the pivot column keeps v (rows >= j), R entries above stay
#pragma unroll
                    for (int r = RS; r < 32; ++r) a[qq][r] = (r >= j) ? vr[r - RS] : a[qq][r];
                } else if (c > j && c < n) {     // reflection.rs:29-30
                    float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
#pragma unroll
                    for (int r = RS; r < 32; r += 4) {
                        d0 += vr[r - RS] * a[qq][r];
                        d1 += vr[r + 1 - RS] * a[qq][r + 1];
                        d2 += vr[r + 2 - RS] * a[qq][r + 2];
                        d3 += vr[r + 3 - RS] * a[qq][r + 3];
                    }
                    const float fac = -2.f * ((d0 + d1) + (d2 + d3));
#pragma unroll
                    for (int r = RS; r < 32; ++r) a[qq][r] += fac * vr[r - RS];
                }
            }
        }
        __syncwarp();
    }
}

// ---- f32, 32 x 32 exactly: every column step specialised at compile time ---------------------------
// ncu on qr_batched4_kernel (profiles/r1_batched.md): 2966 warp instructions per matrix, only 32 % of them FFMA;
// 37 % are ISETP / IMAD / LOP3 / FSEL / BRA / LEA -- the row and column predicates of a run-time pivot index j
// (`r >= j`, `c == j`, `c > j`).  With J a template parameter the row ranges are exact (rows >= J, not rows
// >= 8 * (J / 8)), the column blocks right of the pivot block run unpredicated, and only the pivot block keeps
// two lane predicates.
template <int J>
__device__ __forceinline__ void column_step_fixed(float (&a)[4][32], float *colbuf, float *vbuf, int g, float (&dg)[4],
                                                  unsigned gmask) {
    constexpr int Q = J / 8, JJ = J % 8, R4 = J & ~3;
    if (g == JJ) {
#pragma unroll
        for (int r = R4; r < 32; r += 4)
            *reinterpret_cast<float4 *>(colbuf + r) = make_float4(a[Q][r], a[Q][r + 1], a[Q][r + 2], a[Q][r + 3]);
    }
    __syncwarp();
    float nsq = 0.f, xs[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const int r = g + 8 * t;
        xs[t] = (8 * t + 7 >= J) ? ((r >= J) ? colbuf[r] : 0.f) : 0.f;     // rows of this lane: g, g+8, g+16, g+24
        nsq = fmaf(xs[t], xs[t], nsq);
    }
    // the four matrices of the warp run in lockstep: full-mask shuffles (xor 1, 2, 4 stay inside the 8-lane group)
    // avoid the WARPSYNC / collective bracket a partial mask costs
    nsq += __shfl_xor_sync(0xffffffffu, nsq, 1);
    nsq += __shfl_xor_sync(0xffffffffu, nsq, 2);
    nsq += __shfl_xor_sync(0xffffffffu, nsq, 4);
    const float f = colbuf[J];
    const float rn = nsq > 0.f ? rsqrtf(nsq) : 0.f;
    float nrm = nsq * rn;                                        // householder.rs:13
    nrm = fmaf(0.5f * rn, fmaf(-nrm, nrm, nsq), nrm);
    const float s = (signbit(f) ? -1.f : 1.f) * nrm;             // :16
    const float newsq = (nsq + fabsf(f) * nrm) * 2.f;            // :19-20
    const bool some = newsq != 0.f;                              // :22
    const float rd = some ? rsqrtf(newsq) : 0.f;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        if (8 * t + 7 >= R4) {
            const int r = g + 8 * t;
            if (r >= R4) vbuf[r] = (some && r >= J) ? ((r == J ? xs[t] + s : xs[t]) * rd) : 0.f;   // :17,23
        }
    }
    if (g == JJ) dg[Q] = some ? -s : 0.f;                        // :24/26
    __syncwarp();
    if (some) {
        float vr[32 - R4];
#pragma unroll
        for (int r = R4; r < 32; r += 4) {
            const float4 q4 = *reinterpret_cast<const float4 *>(vbuf + r);
            vr[r - R4] = q4.x; vr[r + 1 - R4] = q4.y; vr[r + 2 - R4] = q4.z; vr[r + 3 - R4] = q4.w;
        }
        {   // the pivot's own block of columns: lane JJ keeps v, lanes > JJ are reflected, lanes < JJ are done
            float d0 = 0.f, d1 = 0.f;
#pragma unroll
            for (int r = J; r < 32; r += 2) {
                d0 = fmaf(vr[r - R4], a[Q][r], d0);
                if (r + 1 < 32) d1 = fmaf(vr[r + 1 - R4], a[Q][r + 1], d1);
            }
            const float fac = (g > JJ) ? -2.f * (d0 + d1) : 0.f;
            const bool piv = g == JJ;
#pragma unroll
            for (int r = J; r < 32; ++r) a[Q][r] = piv ? vr[r - R4] : fmaf(fac, vr[r - R4], a[Q][r]);   // reflection.rs:29-30
        }
#pragma unroll
        for (int qq = Q + 1; qq < 4; ++qq) {   // columns right of the pivot block: no predicates at all
            float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
#pragma unroll
            for (int r = J; r < 32; r += 4) {
                d0 = fmaf(vr[r - R4], a[qq][r], d0);
                if (r + 1 < 32) d1 = fmaf(vr[r + 1 - R4], a[qq][r + 1], d1);
                if (r + 2 < 32) d2 = fmaf(vr[r + 2 - R4], a[qq][r + 2], d2);
                if (r + 3 < 32) d3 = fmaf(vr[r + 3 - R4], a[qq][r + 3], d3);
            }
            const float fac = -2.f * ((d0 + d1) + (d2 + d3));
#pragma unroll
            for (int r = J; r < 32; ++r) a[qq][r] = fmaf(fac, vr[r - R4], a[qq][r]);
        }
    }
    __syncwarp();
}

// Third generation of the column step: the pivot lane holds its whole column in registers, so it makes the
// reflector alone (the other lanes run the same instructions on their own column of the block -- free under SIMT --
// and discard the result) and v crosses shared memory ONCE.  Compared with column_step_fixed this removes one of the
// two shared-memory round trips and the three shuffles from the dependent chain of a step (ncu: short_scoreboard
// 1.37 and wait 1.05 stalls per issue with only two warps per scheduler to hide them).  MEASURED: 1.21 ms against
// 1.06 ms for column_step_fixed -- the redundant norm / scaling arithmetic on eight lanes (+14 % instructions) costs
// more than the shorter chain saves -- so it is kept only behind batched_quad = 3.
template <int J>
__device__ __forceinline__ void column_step_fixed2(float (&a)[4][32], float *vbuf, int g, int lane, float (&dg)[4]) {
    constexpr int Q = J / 8, JJ = J % 8, R4 = J & ~3;
    const bool piv = g == JJ;
    float n0 = 0.f, n1 = 0.f, n2 = 0.f, n3 = 0.f;
#pragma unroll
    for (int r = J; r < 32; r += 4) {
        n0 = fmaf(a[Q][r], a[Q][r], n0);
        if (r + 1 < 32) n1 = fmaf(a[Q][r + 1], a[Q][r + 1], n1);
        if (r + 2 < 32) n2 = fmaf(a[Q][r + 2], a[Q][r + 2], n2);
        if (r + 3 < 32) n3 = fmaf(a[Q][r + 3], a[Q][r + 3], n3);
    }
    const float nsq = (n0 + n1) + (n2 + n3);                     // householder.rs:12
    const float f = a[Q][J];
    const float rn = nsq > 0.f ? rsqrtf(nsq) : 0.f;
    float nrm = nsq * rn;                                        // :13
    nrm = fmaf(0.5f * rn, fmaf(-nrm, nrm, nsq), nrm);
    const float s = (signbit(f) ? -1.f : 1.f) * nrm;             // :16
    const float newsq = (nsq + fabsf(f) * nrm) * 2.f;            // :19-20
    const bool some = newsq != 0.f;                              // :22
    const float rd = some ? rsqrtf(newsq) : 0.f;
    float v[32 - R4];
#pragma unroll
    for (int r = R4; r < 32; ++r) v[r - R4] = (r < J) ? 0.f : (some ? ((r == J ? a[Q][r] + s : a[Q][r]) * rd) : 0.f);   // :17,23
    if (piv) {
#pragma unroll
        for (int r = R4; r < 32; r += 4)
            *reinterpret_cast<float4 *>(vbuf + r) = make_float4(v[r - R4], v[r + 1 - R4], v[r + 2 - R4], v[r + 3 - R4]);
        dg[Q] = some ? -s : 0.f;                                 // :24/26
        if (some) {
#pragma unroll
            for (int r = J; r < 32; ++r) a[Q][r] = v[r - R4];    // the pivot column keeps v
        }
    }
    const bool some_m = __shfl_sync(0xffffffffu, some ? 1 : 0, (lane & 24) + JJ) != 0;    // the pivot lane's verdict
    __syncwarp();
    if (some_m) {
        float vr[32 - R4];
#pragma unroll
        for (int r = R4; r < 32; r += 4) {
            const float4 q4 = *reinterpret_cast<const float4 *>(vbuf + r);
            vr[r - R4] = q4.x; vr[r + 1 - R4] = q4.y; vr[r + 2 - R4] = q4.z; vr[r + 3 - R4] = q4.w;
        }
        {   // the pivot's own block: lanes > JJ are reflected; fac = 0 leaves the others (the pivot lane holds v) alone
            float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
#pragma unroll
            for (int r = J; r < 32; r += 4) {
                d0 = fmaf(vr[r - R4], a[Q][r], d0);
                if (r + 1 < 32) d1 = fmaf(vr[r + 1 - R4], a[Q][r + 1], d1);
                if (r + 2 < 32) d2 = fmaf(vr[r + 2 - R4], a[Q][r + 2], d2);
                if (r + 3 < 32) d3 = fmaf(vr[r + 3 - R4], a[Q][r + 3], d3);
            }
            const float fac = (g > JJ) ? -2.f * ((d0 + d1) + (d2 + d3)) : 0.f;
#pragma unroll
            for (int r = J; r < 32; ++r) a[Q][r] = fmaf(fac, vr[r - R4], a[Q][r]);   // reflection.rs:29-30
        }
#pragma unroll
        for (int qq = Q + 1; qq < 4; ++qq) {
            float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
#pragma unroll
            for (int r = J; r < 32; r += 4) {
                d0 = fmaf(vr[r - R4], a[qq][r], d0);
                if (r + 1 < 32) d1 = fmaf(vr[r + 1 - R4], a[qq][r + 1], d1);
                if (r + 2 < 32) d2 = fmaf(vr[r + 2 - R4], a[qq][r + 2], d2);
                if (r + 3 < 32) d3 = fmaf(vr[r + 3 - R4], a[qq][r + 3], d3);
            }
            const float fac = -2.f * ((d0 + d1) + (d2 + d3));
#pragma unroll
            for (int r = J; r < 32; ++r) a[qq][r] = fmaf(fac, vr[r - R4], a[qq][r]);
        }
    }
    __syncwarp();
}

template <int J0, int GEN>
__device__ __forceinline__ void column_steps_fixed8(float (&a)[4][32], float *colbuf, float *vbuf, int g, int lane, float (&dg)[4],
                                                    unsigned gmask) {
#define LFB_STEP(J)                                                            \
    if constexpr (GEN == 2) column_step_fixed<J>(a, colbuf, vbuf, g, dg, gmask); \
    else column_step_fixed2<J>(a, vbuf, g, lane, dg);
    LFB_STEP(J0 + 0) LFB_STEP(J0 + 1) LFB_STEP(J0 + 2) LFB_STEP(J0 + 3)
    LFB_STEP(J0 + 4) LFB_STEP(J0 + 5) LFB_STEP(J0 + 6) LFB_STEP(J0 + 7)
#undef LFB_STEP
}

template <int GEN>
__global__ void __launch_bounds__(128, 2) qr_batched4_32_kernel(float *__restrict__ A, int64_t batch, float *__restrict__ diag) {
    __shared__ __align__(16) float s_col[4][4][32];
    __shared__ __align__(16) float s_v[4][4][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane & 7, grp = lane >> 3;
    float *colbuf = s_col[warp][grp], *vbuf = s_v[warp][grp];
    const unsigned gmask = 0xffu << (lane & 24);
    const int64_t nquads = (batch + 3) / 4;
    for (int64_t qd = (int64_t)blockIdx.x * 4 + warp; qd < nquads; qd += (int64_t)gridDim.x * 4) {
        const int64_t b = qd * 4 + grp;
        const bool live = b < batch;
        float *mat = A + (live ? b : 0) * 1024;
        float a[4][32];
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int i = 0; i < 32; ++i) a[q][i] = live ? mat[i * 32 + 8 * q + g] : 0.f;
        float dg[4] = {0.f, 0.f, 0.f, 0.f};
        column_steps_fixed8<0, GEN>(a, colbuf, vbuf, g, lane, dg, gmask);
        column_steps_fixed8<8, GEN>(a, colbuf, vbuf, g, lane, dg, gmask);
        column_steps_fixed8<16, GEN>(a, colbuf, vbuf, g, lane, dg, gmask);
        column_steps_fixed8<24, GEN>(a, colbuf, vbuf, g, lane, dg, gmask);
        // reference sign convention, applied once: running sign P_r over the pivots of this matrix
        float p = 1.f, prev[4] = {1.f, 1.f, 1.f, 1.f};
#pragma unroll
        for (int r = 0; r < 32; ++r) {
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (8 * q + g == r) prev[q] = p;                                 // P_{c-1}
            const float br = __shfl_sync(0xffffffffu, dg[r >> 3], (lane & 24) + (r & 7));
            if (br != 0.f) p = signbit(br) ? -1.f : 1.f;
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (8 * q + 7 > r && r < 8 * q + g) a[q][r] *= p;                // R[r, c] *= P_r
        }
        if (live) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int c = 8 * q + g;
#pragma unroll
                for (int i = 0; i < 32; ++i) mat[i * 32 + c] = (i < c) ? a[q][i] : prev[q] * a[q][i];
                diag[b * 32 + c] = (dg[q] != 0.0f) ? prev[q] * dg[q] : 0.0f;   // `None` pivot: +0.0 (householder.rs:50)
            }
        }
        __syncwarp();
    }
}

__global__ void __launch_bounds__(128, 2) qr_batched4_kernel(float *__restrict__ A, int64_t batch, int m, int n, float *__restrict__ diag) {
    __shared__ __align__(16) float s_col[4][4][32];
    __shared__ __align__(16) float s_v[4][4][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane & 7, grp = lane >> 3;
    float *colbuf = s_col[warp][grp], *vbuf = s_v[warp][grp];
    const unsigned gmask = 0xffu << (lane & 24);
    const int64_t nquads = (batch + 3) / 4;
    for (int64_t qd = (int64_t)blockIdx.x * 4 + warp; qd < nquads; qd += (int64_t)gridDim.x * 4) {
        const int64_t b = qd * 4 + grp;
        const bool live = b < batch;
        float *mat = A + (live ? b : 0) * (int64_t)m * n;
        float a[4][32];
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int i = 0; i < 32; ++i) a[q][i] = (live && i < m && 8 * q + g < n) ? mat[i * n + 8 * q + g] : 0.f;
        float dg[4] = {0.f, 0.f, 0.f, 0.f};
        column_steps4<0>(a, colbuf, vbuf, g, n, dg);
        if (n > 8) column_steps4<1>(a, colbuf, vbuf, g, n, dg);
        if (n > 16) column_steps4<2>(a, colbuf, vbuf, g, n, dg);
        if (n > 24) column_steps4<3>(a, colbuf, vbuf, g, n, dg);
        // reference sign convention, applied once: running sign P_r over the pivots of this matrix
        float p = 1.f, prev[4] = {1.f, 1.f, 1.f, 1.f};
#pragma unroll
        for (int r = 0; r < 32; ++r) {
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (8 * q + g == r) prev[q] = p;                                 // P_{c-1}
            const float br = __shfl_sync(gmask, dg[r >> 3], (lane & 24) + (r & 7));
            if (br != 0.f) p = signbit(br) ? -1.f : 1.f;
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (r < 8 * q + g) a[q][r] *= p;                                 // R[r, c] *= P_r
        }
        if (live) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int c = 8 * q + g;
                if (c < n) {
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (i < m) mat[i * n + c] = (i < c) ? a[q][i] : prev[q] * a[q][i];
                    diag[b * n + c] = (dg[q] != 0.0f) ? prev[q] * dg[q] : 0.0f;   // `None` pivot: +0.0 (householder.rs:50)
                }
            }
        }
        __syncwarp();
    }
}

}  // namespace

// ---- batched Cholesky of n x n (n <= 32) matrices, packed row-major [batch][n][n] -------------------------
// cholesky.rs:51-83 per matrix (the reference has no batched entry point: this is the loop a caller writes).  Warp per
// matrix, lane = row: lane i keeps row i of the LOWER triangle in registers (the strict upper triangle is never read
// and, for the dirty variant, never written -- cholesky.rs:17-19).  Right-looking: column j is scaled by 1/sqrt(pivot),
// then every lane subtracts l_ij * l_kj from its row, l_kj arriving by shuffle from lane k.  fail[b] = the first row
// whose pivot is not positive (cholesky.rs:69-71), -1 otherwise; a failed matrix is left partly factored, as in the
// reference.
template <typename T>
__global__ void __launch_bounds__(128) chol_batched_kernel(T *__restrict__ A, int64_t batch, int n, int clean, int *__restrict__ fail) {
    __shared__ T tile[4][32][33];                           // per-warp staging: coalesced global access, conflict-free row reads
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t b = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
    if (b >= batch) return;
    T *mat = A + b * (int64_t)n * n;
    const int nn = n * n;
    for (int e = lane; e < nn; e += 32) tile[warp][e / n][e % n] = mat[e];       // the whole matrix, 128-byte transactions
    __syncwarp();
    T row[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) row[c] = (lane < n && c <= lane) ? tile[warp][lane][c] : T(0);   // only the lower triangle is used
    int bad = -1;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        if (j < n && bad < 0) {                            // warp-uniform
            const T p = __shfl_sync(0xffffffffu, row[j], j);
            if (p <= T(0)) bad = j;                        // cholesky.rs:69 (false for NaN: the reference goes on with a NaN factor)
            if (bad < 0) {
                const T d = t_sqrt(p);
                const T lij = (lane == j) ? d : row[j] / d;
                if (lane >= j) row[j] = lij;
#pragma unroll
                for (int k = j + 1; k < 32; ++k) {
                    if (k < n) {
                        const T lkj = __shfl_sync(0xffffffffu, lij, k);
                        if (lane >= k) row[k] -= lij * lkj;
                    }
                }
            }
        }
    }
    __syncwarp();
    if (lane < n) {
#pragma unroll
        for (int c = 0; c < 32; ++c) {
            if (c <= lane) tile[warp][lane][c] = row[c];
            else if (clean && c < n) tile[warp][lane][c] = T(0);                 // cholesky.rs:78-82; dirty: the staged originals go back
        }
    }
    __syncwarp();
    for (int e = lane; e < nn; e += 32) mat[e] = tile[warp][e / n][e % n];
    if (lane == 0) fail[b] = bad;
}

// ---- f32, 32 x 32: eight lanes per matrix, the layout of qr_batched4_32_kernel --------------------------------------
// The warp-per-matrix kernel above runs 262144 matrices in 2.40 ms (0.9 TB/s, 0.14 of HBM): one shuffle per updated column
// per step and a lane that owns a ROW idles for most of the triangle.  Here lane g of an 8-lane group keeps columns g, g + 8,
// g + 16, g + 24 (4 x 32 registers, rows down the register index), every column step is specialised at compile time, the scaled
// pivot column crosses shared memory once per step (8 STS.128 by its owner, 8 LDS.128 by everybody), and the rank-1 update is
// plain FFMAs on whole register columns.  Entries above the diagonal are scratch in registers: the dirty variant never stores
// them (cholesky.rs:17-19), the clean one stores zeros (cholesky.rs:78-82).  A non-positive pivot (cholesky.rs:69-71) freezes
// that matrix: from then on its group publishes zeros, so the other three matrices of the warp go on without divergence.
template <int J>
__device__ __forceinline__ void chol_step8(float (&a)[4][32], float *vbuf, int g, int lane, int &bad) {
    constexpr int Q = J / 8, JJ = J % 8, R4 = J & ~3;
    const float p = __shfl_sync(0xffffffffu, a[Q][J], (lane & 24) + JJ);
    if (bad < 0 && p <= 0.f) bad = J;                    // false for NaN: the reference goes on with a NaN factor
    const bool ok = bad < 0;
    const float d = sqrtf(p);
    const float rinv = 1.f / d;
    if (g == JJ) {
        float l[32 - R4];
#pragma unroll
        for (int r = R4; r < 32; ++r) l[r - R4] = (r < J || !ok) ? 0.f : (r == J ? d : a[Q][r] * rinv);
        if (ok) {
#pragma unroll
            for (int r = J; r < 32; ++r) a[Q][r] = l[r - R4];
        }
#pragma unroll
        for (int r = R4; r < 32; r += 4)
            *reinterpret_cast<float4 *>(vbuf + r) = make_float4(l[r - R4], l[r + 1 - R4], l[r + 2 - R4], l[r + 3 - R4]);
    }
    __syncwarp();
    if constexpr (J < 31) {
        float l[32 - R4];
#pragma unroll
        for (int r = R4; r < 32; r += 4) {
            const float4 q4 = *reinterpret_cast<const float4 *>(vbuf + r);
            l[r - R4] = q4.x; l[r + 1 - R4] = q4.y; l[r + 2 - R4] = q4.z; l[r + 3 - R4] = q4.w;
        }
        {   // the pivot's own slot: only the columns right of it (g > JJ) change; a select, not a multiply by zero (NaN safety)
            const bool upd = g > JJ;
            const float lc = upd ? vbuf[8 * Q + g] : 0.f;
#pragma unroll
            for (int r = J + 1; r < 32; ++r) a[Q][r] = upd ? fmaf(-l[r - R4], lc, a[Q][r]) : a[Q][r];
        }
#pragma unroll
        for (int q = Q + 1; q < 4; ++q) {
            const float lc = vbuf[8 * q + g];
#pragma unroll
            for (int r = 8 * q; r < 32; ++r) a[q][r] = fmaf(-l[r - R4], lc, a[q][r]);
        }
    }
    __syncwarp();
}

template <int J0>
__device__ __forceinline__ void chol_steps8(float (&a)[4][32], float *vbuf, int g, int lane, int &bad) {
    chol_step8<J0 + 0>(a, vbuf, g, lane, bad); chol_step8<J0 + 1>(a, vbuf, g, lane, bad);
    chol_step8<J0 + 2>(a, vbuf, g, lane, bad); chol_step8<J0 + 3>(a, vbuf, g, lane, bad);
    chol_step8<J0 + 4>(a, vbuf, g, lane, bad); chol_step8<J0 + 5>(a, vbuf, g, lane, bad);
    chol_step8<J0 + 6>(a, vbuf, g, lane, bad); chol_step8<J0 + 7>(a, vbuf, g, lane, bad);
}

__global__ void __launch_bounds__(128, 3) chol_batched8_32_kernel(float *__restrict__ A, int64_t batch, int clean, int *__restrict__ fail) {
    __shared__ __align__(16) float s_v[4][4][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane & 7, grp = lane >> 3;
    float *vbuf = s_v[warp][grp];
    const int64_t nquads = (batch + 3) / 4;
    for (int64_t qd = (int64_t)blockIdx.x * 4 + warp; qd < nquads; qd += (int64_t)gridDim.x * 4) {
        const int64_t b = qd * 4 + grp;
        const bool live = b < batch;
        float *mat = A + (live ? b : 0) * 1024;
        float a[4][32];
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int i = 0; i < 32; ++i) a[q][i] = live ? mat[i * 32 + 8 * q + g] : (i == 8 * q + g ? 1.f : 0.f);
        int bad = -1;
        chol_steps8<0>(a, vbuf, g, lane, bad);
        chol_steps8<8>(a, vbuf, g, lane, bad);
        chol_steps8<16>(a, vbuf, g, lane, bad);
        chol_steps8<24>(a, vbuf, g, lane, bad);
        if (live) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int c = 8 * q + g;
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    if (i >= c) mat[i * 32 + c] = a[q][i];
                    else if (clean) mat[i * 32 + c] = 0.f;
                }
            }
            if (g == 0) fail[b] = bad;
        }
        __syncwarp();
    }
}

template <typename T>
void cholesky_batched(lfb_handle &h, T *A, int64_t batch, int64_t n, int clean, int *fail) {
    if (batch <= 0 || n <= 0) return;
    if constexpr (sizeof(T) == 4) {
        if (n == 32 && h.opt.batched_quad) {
            const int64_t blocks4 = std::min<int64_t>(cdiv(cdiv(batch, 4), 4), (int64_t)h.sm_count * 16);
            chol_batched8_32_kernel<<<(unsigned)blocks4, 128, 0, h.stream>>>(A, batch, clean, fail);
            LFB_LAUNCH_CHECK(h);
            return;
        }
    }
    chol_batched_kernel<T><<<(unsigned)cdiv(batch, 4), 128, 0, h.stream>>>(A, batch, (int)n, clean, fail);
    LFB_LAUNCH_CHECK(h);
}
template void cholesky_batched<float>(lfb_handle &, float *, int64_t, int64_t, int, int *);
template void cholesky_batched<double>(lfb_handle &, double *, int64_t, int64_t, int, int *);

template <typename T>
void qr_batched(lfb_handle &h, T *A, int64_t batch, int64_t m, int64_t n, T *diag) {
    if (batch <= 0 || n <= 0) return;
    if constexpr (sizeof(T) == 4) {
        if (h.opt.batched_quad) {
            int64_t blocks4 = std::min<int64_t>(cdiv(cdiv(batch, 4), 4), (int64_t)h.sm_count * 16);
            if (m == 32 && n == 32 && h.opt.batched_quad >= 3)
                qr_batched4_32_kernel<3><<<(unsigned)blocks4, 128, 0, h.stream>>>(A, batch, diag);
            else if (m == 32 && n == 32 && h.opt.batched_quad == 2)
                qr_batched4_32_kernel<2><<<(unsigned)blocks4, 128, 0, h.stream>>>(A, batch, diag);
            else
                qr_batched4_kernel<<<(unsigned)blocks4, 128, 0, h.stream>>>(A, batch, (int)m, (int)n, diag);
            LFB_LAUNCH_CHECK(h);
            return;
        }
    }
    constexpr int WPB = wpb<T>();
    int64_t blocks = std::min<int64_t>(cdiv(batch, WPB), (int64_t)h.sm_count * 16);
    qr_batched_kernel<T><<<(unsigned)blocks, WPB * 32, 0, h.stream>>>(A, batch, (int)m, (int)n, diag);
    LFB_LAUNCH_CHECK(h);
}

template void qr_batched<float>(lfb_handle &, float *, int64_t, int64_t, int64_t, float *);
template void qr_batched<double>(lfb_handle &, double *, int64_t, int64_t, int64_t, double *);

}  // namespace lfb
