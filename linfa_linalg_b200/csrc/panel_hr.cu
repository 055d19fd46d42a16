// The 128-column panel of the blocked QR as two single-CTA kernels (qr.rs:29-45 through householder.cu: qr_factor_std).
//
// qr_trace on QR 16384^2 (profiles/r2_qr_panel.md): from panel 60 of 128 on, the step is bound by the panel chain on the
// look-ahead stream, and that chain is ~25 dependent launches of ~24 us each (0.6 ms alone, 0.5 - 2.3 ms next to the trailing
// GEMM, which every launch has to squeeze past).  The two n x n stages between the Gram GEMM and the tall GEMM are latency,
// not work (128^3 flops each), so each becomes ONE launch with the block resident in shared memory:
//
//   cholqr128_kernel     G = A^T A  ->  R = chol(G)^T, R^-1, and the guard numbers of cholqr.cu (condition bound, smallest
//                        pivot, finiteness);
//   hr_panel128_kernel   Q_top = A_top R^-1, LU of (Q_top - S) (tsqr_hr.cu: the Householder reconstruction), then everything the
//                        driver needs from it: the top block of the factor in the driver's internal convention, beta, the
//                        matrix M = R^-1 U^-1 C that turns the rows below into reflector rows (V_2 = A_2 M, one tall GEMM), and
//                        the compact-WY factor  T = -C^-1 U S Y_1^-T C^-1  (Ballard et al.'s T for Y, rescaled to the unit-norm
//                        reflectors V = Y C) -- which replaces the second Gram GEMM + triangular inversion of build_t.
//
// Every O(n^3) stage is a "register sweep": lane tr of warp tc (NW warps) owns the 4 x (128 / NW) elements (tr + 32 x, tc + NW y) of
// the working matrix in registers; a step reads one published pivot row / column from shared memory, updates its elements (dead
// ones are multiplied by zero: no divergence, no imbalance), and the owners of the next pivot row / column publish it -- one
// barrier per step.  What bounds a step is the SM's ONE shared-memory instruction per cycle (LFB_PANEL_DBG=1 prints clock64 per
// phase): in-place loops with a shared access per FMA were 3-4 x slower than the first register version, and with 32 warps and
// 4 x 4 tiles every warp re-reads the whole pivot column (384 wavefronts, ~1200 cycles per step); NW = 8 warps with 4 x 16 tiles
// halve that.  The two dense products (Q_top = A_top R^-1 and T) are register-tiled the same way.
//
// The formulas were checked against a plain Householder sweep in NumPy before this file was written (V, beta, R, T all to 1e-14
// at 700 x 128); tests/test_gpu_parity*.py hold the result to the oracle.
#include <cstdio>
#include <cstdlib>

#include "common.cuh"

namespace lfb {
namespace {

constexpr int PN = 128, PLD = PN + 1;
constexpr int NW = 8;                      // warps per CTA
constexpr int YN = PN / NW;                // columns per thread: tc + NW * y
constexpr int NTH = NW * 32;
#define LFB_MARK(i) do { if (dbg && threadIdx.x == 0) dbg[i] = clock64(); } while (0)
__device__ __forceinline__ int pk(int i, int j) { return j * (j + 1) / 2 + i; }   // packed upper triangle, i <= j

// Column s of the register tile with a run-time (warp-uniform) s.  A dynamic index would send the tile to local memory and a
// select chain costs 4 (YN - 1) selects on the critical path of a step; the switch executes four moves.
#define LFB_COL_CASES(OP) \
    OP(0) OP(1) OP(2) OP(3) OP(4) OP(5) OP(6) OP(7) OP(8) OP(9) OP(10) OP(11) OP(12) OP(13) OP(14) OP(15)
template <typename T> __device__ __forceinline__ void get_col(const T (&a)[4][YN], int s, T (&v)[4]) {
    switch (s) {
#define LFB_OP(n) case n: if (n < YN) { v[0] = a[0][n < YN ? n : 0]; v[1] = a[1][n < YN ? n : 0]; v[2] = a[2][n < YN ? n : 0]; v[3] = a[3][n < YN ? n : 0]; } break;
        LFB_COL_CASES(LFB_OP)
#undef LFB_OP
        default: break;
    }
}
template <typename T> __device__ __forceinline__ void set_col(T (&a)[4][YN], int s, const T (&v)[4]) {
    switch (s) {
#define LFB_OP(n) case n: if (n < YN) { a[0][n < YN ? n : 0] = v[0]; a[1][n < YN ? n : 0] = v[1]; a[2][n < YN ? n : 0] = v[2]; a[3][n < YN ? n : 0] = v[3]; } break;
        LFB_COL_CASES(LFB_OP)
#undef LFB_OP
        default: break;
    }
}
template <typename T> __device__ __forceinline__ T row_of(const T (&a)[4][YN], int s, int y) {
    return s == 0 ? a[0][y] : s == 1 ? a[1][y] : s == 2 ? a[2][y] : a[3][y];
}

#define FOR_X _Pragma("unroll") for (int x = 0; x < 4; ++x)
#define FOR_Y _Pragma("unroll") for (int y = 0; y < YN; ++y)

template <typename T>
__global__ void __launch_bounds__(NTH, 1) cholqr128_kernel(const T *__restrict__ G, int64_t ldg, T *__restrict__ R, int64_t ldr,
                                                           T *__restrict__ Rinv, int64_t ldri, double *__restrict__ guard, long long *dbg) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *S = reinterpret_cast<T *>(smem_raw);      // [PN][PLD]: G, then lower = L and strict upper = (L^-1)^T
    T *rk = S + PN * PLD;                        // L_kk
    T *xd = rk + PN;                             // 1 / L_kk
    T *buf = xd + PN;                            // [2][PN] published pivot column / row
    double *red = reinterpret_cast<double *>(buf + 2 * PN);   // 4 x PN norms
    __shared__ int s_fail;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tr = lane, tc = warp;
    if (tid == 0) s_fail = 0;
    LFB_MARK(0);
    for (int j = warp; j < PN; j += NW)
        for (int i = j + lane; i < PN; i += 32) S[i * PLD + j] = G[i + (int64_t)j * ldg];
    __syncthreads();
    T a[4][YN];
    FOR_X FOR_Y {
        const int i = tr + 32 * x, j = tc + NW * y;
        a[x][y] = i >= j ? S[i * PLD + j] : S[j * PLD + i];
    }
    if (tc == 0) {
        FOR_X buf[tr + 32 * x] = a[x][0];
    }
    __syncthreads();
    LFB_MARK(1);
    // Cholesky, right-looking.  Column k keeps the unscaled Schur-complement entries a_ik (L_ik = a_ik / sqrt(a_kk)); the full
    // square is updated (the matrix stays symmetric), so one published column serves rows and columns.
    for (int k = 0; k < PN; ++k) {
        const T *cb = buf + (k & 1) * PN;
        T *cn = buf + ((k & 1) ^ 1) * PN;
        const T d = cb[k];
        if (!(d > T(0)) || !isfinite(d)) {       // uniform: every thread reads the same d
            if (tid == 0) s_fail = 1;
            break;
        }
        if (tid == 0) rk[k] = sqrt(d);
        const T dinv = T(1) / d;
        T l[4], u[YN];
        FOR_X l[x] = (tr + 32 * x > k) ? -cb[tr + 32 * x] * dinv : T(0);
        FOR_Y u[y] = (tc + NW * y > k) ? cb[tc + NW * y] : T(0);
        FOR_X FOR_Y a[x][y] = fma(l[x], u[y], a[x][y]);
        if (k + 1 < PN && tc == ((k + 1) & (NW - 1))) {
            T v[4];
            get_col(a, (k + 1) / NW, v);
            FOR_X cn[tr + 32 * x] = v[x];
        }
        __syncthreads();
    }
    __syncthreads();
    if (s_fail) {
        if (tid == 0) { guard[0] = 1e300; guard[1] = 0.0; guard[2] = 0.0; }
        return;
    }
    LFB_MARK(2);
    if (tid < PN) xd[tid] = T(1) / rk[tid];
    __syncthreads();
    FOR_X FOR_Y {
        const int i = tr + 32 * x, j = tc + NW * y;
        if (i > j) S[i * PLD + j] = a[x][y] * xd[j];
        else if (i == j) S[i * PLD + j] = rk[j];
    }
    // X = L^-1 by a forward sweep on W = I: row k of X is W[k, :] / L_kk, then W[i, :] -= L_ik X[k, :] for i > k
    FOR_X FOR_Y a[x][y] = (tr + 32 * x == tc + NW * y) ? T(1) : T(0);
    if (tr == 0) {
        FOR_Y {
            a[0][y] *= xd[0];
            buf[tc + NW * y] = a[0][y];
        }
    }
    __syncthreads();
    for (int k = 0; k < PN; ++k) {
        const T *rb = buf + (k & 1) * PN;
        T *rn = buf + ((k & 1) ^ 1) * PN;
        T l[4], u[YN];
        FOR_X l[x] = (tr + 32 * x > k) ? -S[(tr + 32 * x) * PLD + k] : T(0);
        FOR_Y u[y] = rb[tc + NW * y];
        FOR_X FOR_Y a[x][y] = fma(l[x], u[y], a[x][y]);
        if (k + 1 < PN && tr == ((k + 1) & 31)) {
            const T sc = xd[k + 1];
            const int xs = (k + 1) >> 5;
            FOR_Y {
                const T v = row_of(a, xs, y) * sc;
                rn[tc + NW * y] = v;
                FOR_X if (x == xs) a[x][y] = v;
            }
        }
        __syncthreads();
    }
    LFB_MARK(3);
    // X (lower) goes to the strict upper part of S transposed, for the guard sums and a coalesced write
    FOR_X FOR_Y {
        const int i = tr + 32 * x, j = tc + NW * y;
        if (i > j) S[j * PLD + i] = a[x][y];
    }
    __syncthreads();
    LFB_MARK(4);
    // guard: sqrt(||L||_1 ||L||_inf ||X||_1 ||X||_inf) >= cond_2(L)
    if (tid < PN) {
        const int t = tid;
        double cl = 0.0, rl = 0.0, cx = fabs((double)xd[t]), rx = fabs((double)xd[t]);
        for (int i = t; i < PN; ++i) cl += fabs((double)S[i * PLD + t]);
        for (int j = 0; j <= t; ++j) rl += fabs((double)S[t * PLD + j]);
        for (int i = t + 1; i < PN; ++i) cx += fabs((double)S[t * PLD + i]);      // X_it
        for (int j = 0; j < t; ++j) rx += fabs((double)S[j * PLD + t]);           // X_tj
        red[t] = cl; red[PN + t] = rl; red[2 * PN + t] = cx; red[3 * PN + t] = rx;
    }
    __syncthreads();
    if (warp == 0) {
        double m[4] = {0.0, 0.0, 0.0, 0.0}, dm = 1e300;
        for (int t = lane; t < PN; t += 32) {
#pragma unroll
            for (int q = 0; q < 4; ++q) m[q] = fmax(m[q], red[q * PN + t]);
            dm = fmin(dm, (double)rk[t]);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
            for (int q = 0; q < 4; ++q) m[q] = fmax(m[q], __shfl_xor_sync(0xffffffffu, m[q], o));
            dm = fmin(dm, __shfl_xor_sync(0xffffffffu, dm, o));
        }
        if (lane == 0) {
            const double b = sqrt(m[0] * m[1] * m[2] * m[3]);
            guard[0] = b;
            guard[1] = dm;
            guard[2] = isfinite(b) ? 1.0 : 0.0;
        }
    }
    LFB_MARK(5);
    for (int j = warp; j < PN; j += NW)
        for (int i = lane; i < PN; i += 32) {
            R[i + (int64_t)j * ldr] = i <= j ? S[j * PLD + i] : T(0);                            // L_ji
            Rinv[i + (int64_t)j * ldri] = i < j ? S[i * PLD + j] : (i == j ? xd[j] : T(0));     // X_ji
        }
    LFB_MARK(6);
}

// STAGE 0: everything in one launch.  STAGES 1 / 2 / 3 split it so that the part nothing downstream waits for runs beside the tall
// GEMM on a second stream: 1 = Q_top and its LU (Y_1, U and the s / pivot / c' vectors go to the scratch block Qg), 2 = M and the top
// block (what the tall GEMM V_2 = A_2 M and the panel's write-back need), 3 = Z and T (needed only when the reflector is APPLIED).
template <typename T, int STAGE>
__global__ void __launch_bounds__(NTH, 1) hr_panel128_kernel(T *__restrict__ Atop, int64_t ld, const T *__restrict__ R, int64_t ldr,
                                                            const T *__restrict__ Rinv, int64_t ldri, T *__restrict__ beta,
                                                            T *__restrict__ M, int64_t ldm, T *__restrict__ Tm, int64_t ldt,
                                                            T *__restrict__ Vtop, int64_t ldv, T *__restrict__ Qg, long long *dbg) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *Q = reinterpret_cast<T *>(smem_raw);      // [PN][PLD]: A_top, then Y_1 (strictly lower) and U
    T *P = Q + PN * PLD;                         // packed upper: R^-1, later Z
    T *sv = P + PN * (PN + 1) / 2;               // s_k
    T *pv = sv + PN;                             // U_kk
    T *pinv = pv + PN;                           // 1 / U_kk
    T *cv = pinv + PN;                           // c'_k
    T *colb = cv + PN;                           // [2][PN]
    T *rowb = colb + 2 * PN;                     // [2][PN]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tr = lane, tc = warp;
    T a[4][YN];
    LFB_MARK(0);
    if constexpr (STAGE >= 2) {                  // Y_1 / U and the vectors come from stage 1
        for (int e = tid; e < PN * PLD; e += NTH) Q[e] = Qg[e];
        for (int e = tid; e < 4 * PN; e += NTH) sv[e] = Qg[PN * PLD + e];      // sv, pv, pinv, cv are contiguous
        if constexpr (STAGE == 2) {
            for (int j = warp; j < PN; j += NW)
                for (int i = lane; i <= j; i += 32) P[pk(i, j)] = Rinv[i + (int64_t)j * ldri];
        }
        __syncthreads();
    }
    if constexpr (STAGE <= 1) {
    for (int j = warp; j < PN; j += NW)
        for (int i = lane; i < PN; i += 32) {
            Q[i * PLD + j] = Atop[i + (int64_t)j * ld];
            if (i <= j) P[pk(i, j)] = Rinv[i + (int64_t)j * ldri];
        }
    __syncthreads();
    LFB_MARK(1);
    // Q_top = A_top R^-1, register tile per thread; the result stays in registers as the input of the LU
    FOR_X FOR_Y a[x][y] = T(0);
    for (int k = 0; k <= tc + NW * (YN - 1); ++k) {   // R^-1 is upper triangular: column j needs k <= j
        T l[4], u[YN];
        FOR_X l[x] = Q[(tr + 32 * x) * PLD + k];
        FOR_Y u[y] = (k <= tc + NW * y) ? P[pk(k, tc + NW * y)] : T(0);
        FOR_X FOR_Y a[x][y] = fma(l[x], u[y], a[x][y]);
    }
    if (tc == 0) {
        FOR_X colb[tr + 32 * x] = a[x][0];
    }
    if (tr == 0) {
        FOR_Y rowb[tc + NW * y] = a[0][y];
    }
    __syncthreads();
    LFB_MARK(2);
    // LU of (Q_top - S) without pivoting (|pivot| = 1 + |q_kk| >= 1); multipliers stay unscaled (a_ik) until the write-back
    for (int k = 0; k < PN; ++k) {
        const T *cb = colb + (k & 1) * PN, *rb = rowb + (k & 1) * PN;
        T *cn = colb + ((k & 1) ^ 1) * PN, *rn = rowb + ((k & 1) ^ 1) * PN;
        const T q = cb[k];
        const T sk = q < T(0) ? T(1) : T(-1);    // s_k = -sgn(q_kk), sgn(0) = +1 (tsqr_hr.cu)
        const T p = q - sk;
        const T pin = T(1) / p;
        if (tid == 0) { sv[k] = sk; pv[k] = p; pinv[k] = pin; }
        T l[4], u[YN];
        FOR_X l[x] = (tr + 32 * x > k) ? -cb[tr + 32 * x] * pin : T(0);
        FOR_Y u[y] = (tc + NW * y > k) ? rb[tc + NW * y] : T(0);
        FOR_X FOR_Y a[x][y] = fma(l[x], u[y], a[x][y]);
        if (k + 1 < PN) {
            if (tc == ((k + 1) & (NW - 1))) {
                T v[4];
                get_col(a, (k + 1) / NW, v);
                FOR_X cn[tr + 32 * x] = v[x];
            }
            if (tr == ((k + 1) & 31)) {
                FOR_Y rn[tc + NW * y] = row_of(a, (k + 1) >> 5, y);
            }
        }
        __syncthreads();
    }
    LFB_MARK(3);
    // back to shared memory: Y_1 strictly below the diagonal (scaled now), U on and above it
    FOR_X FOR_Y {
        const int i = tr + 32 * x, j = tc + NW * y;
        Q[i * PLD + j] = i > j ? a[x][y] * pinv[j] : (i == j ? pv[j] : a[x][y]);
    }
    if (tid < PN) {
        const int k = tid;
        cv[k] = -sv[k] * sqrt(fabs(pv[k]) / T(2));
        beta[k] = sv[k] * fabs(R[k + (int64_t)k * ldr]);
    }
    if constexpr (STAGE == 1) {
        __syncthreads();
        for (int e = tid; e < PN * PLD; e += NTH) Qg[e] = Q[e];
        for (int e = tid; e < 4 * PN; e += NTH) Qg[PN * PLD + e] = sv[e];
        return;
    }
    }   // STAGE <= 1
    if constexpr (STAGE == 0 || STAGE == 2) {
    // M = R^-1 U^-1 by a column sweep on W = R^-1: column j of M is W[:, j] / U_jj, then W[:, j'] -= M[:, j] U[j, j'] for j' > j
    FOR_X FOR_Y {
        const int i = tr + 32 * x, j = tc + NW * y;
        a[x][y] = i <= j ? P[pk(i, j)] : T(0);
    }
    __syncthreads();                             // Q, pinv, cv complete; every thread has its part of R^-1
    if (tc == 0) {
        FOR_X {
            a[x][0] *= pinv[0];
            colb[tr + 32 * x] = a[x][0];
        }
    }
    __syncthreads();
    LFB_MARK(4);
    for (int j = 0; j < PN; ++j) {
        const T *cb = colb + (j & 1) * PN;
        T *cn = colb + ((j & 1) ^ 1) * PN;
        T l[4], u[YN];
        FOR_X l[x] = -cb[tr + 32 * x];
        FOR_Y u[y] = (tc + NW * y > j) ? Q[j * PLD + tc + NW * y] : T(0);
        FOR_X FOR_Y a[x][y] = fma(l[x], u[y], a[x][y]);
        if (j + 1 < PN && tc == ((j + 1) & (NW - 1))) {
            const T sc = pinv[j + 1];
            T v[4];
            get_col(a, (j + 1) / NW, v);
            FOR_X {
                v[x] *= sc;
                cn[tr + 32 * x] = v[x];
            }
            set_col(a, (j + 1) / NW, v);
        }
        __syncthreads();
    }
    FOR_X FOR_Y {
        const int i = tr + 32 * x, j = tc + NW * y;
        M[i + (int64_t)j * ldm] = i <= j ? a[x][y] * cv[j] : T(0);
    }
    }   // M
    LFB_MARK(5);
    if ((STAGE == 0 || STAGE == 3) && Tm) {
        // Z = Y_1^-T C^-1 (upper) by a row sweep from the bottom on W = C^-1: row k of Z is final when the sweep reaches it (unit
        // diagonal), then W[i, :] -= Y_1[k, i] Z[k, :] for i < k
        FOR_X FOR_Y a[x][y] = (tr + 32 * x == tc + NW * y) ? T(1) / cv[tc + NW * y] : T(0);
        if (tr == 31) {
            FOR_Y rowb[PN + tc + NW * y] = a[3][y];       // row 127 goes to buffer (127 & 1) = 1
        }
        __syncthreads();
        for (int k = PN - 1; k >= 0; --k) {
            const T *rb = rowb + (k & 1) * PN;
            T *rn = rowb + ((k & 1) ^ 1) * PN;
            T l[4], u[YN];
            FOR_X l[x] = (tr + 32 * x < k) ? -Q[k * PLD + tr + 32 * x] : T(0);
            FOR_Y u[y] = rb[tc + NW * y];
            FOR_X FOR_Y a[x][y] = fma(l[x], u[y], a[x][y]);
            if (k > 0 && tr == ((k - 1) & 31)) {
                FOR_Y rn[tc + NW * y] = row_of(a, (k - 1) >> 5, y);
            }
            __syncthreads();
        }
        LFB_MARK(6);
        // T = -C^-1 (U S) Z: Z to the packed buffer, then a register-tiled product of two upper triangles
        FOR_X FOR_Y {
            const int i = tr + 32 * x, j = tc + NW * y;
            if (i <= j) P[pk(i, j)] = a[x][y];
            a[x][y] = T(0);
        }
        __syncthreads();
        for (int k = tr; k <= tc + NW * (YN - 1); ++k) {    // U_ik needs k >= i, z_kj needs k <= j
            const T sk = sv[k];
            T l[4], u[YN];
            FOR_X l[x] = (k >= tr + 32 * x) ? Q[(tr + 32 * x) * PLD + k] * sk : T(0);
            FOR_Y u[y] = (k <= tc + NW * y) ? P[pk(k, tc + NW * y)] : T(0);
            FOR_X FOR_Y a[x][y] = fma(l[x], u[y], a[x][y]);
        }
        FOR_X FOR_Y {
            const int i = tr + 32 * x, j = tc + NW * y;
            Tm[i + (int64_t)j * ldt] = i <= j ? -a[x][y] / cv[i] : T(0);
        }
    }
    LFB_MARK(7);
    // the top block in the driver's convention: s_i R[i, j] above the diagonal, the reflector heads c'_j y_ij on and below it
    if constexpr (STAGE == 0 || STAGE == 2)
    for (int j = warp; j < PN; j += NW)
        for (int i = lane; i < PN; i += 32) {
            T v;
            if (i < j) {
                const T r = R[i + (int64_t)j * ldr];
                v = sv[i] < T(0) ? -r : r;
            } else {
                v = (i == j) ? cv[j] : Q[i * PLD + j] * cv[j];
            }
            Atop[i + (int64_t)j * ld] = v;
            Vtop[i + (int64_t)j * ldv] = i >= j ? v : T(0);
        }
    LFB_MARK(8);
}

#undef FOR_X
#undef FOR_Y

}  // namespace

template <typename T>
void cholqr128(lfb_handle &h, const T *G, int64_t ldg, T *R, int64_t ldr, T *Rinv, int64_t ldri, double *guard) {
    const size_t smem = sizeof(T) * (size_t)(PN * PLD + 4 * PN) + sizeof(double) * 4 * PN;
    static DeviceOnce cfg;
    cfg.run(h.device, [&] {
        LFB_CUDA(cudaFuncSetAttribute(cholqr128_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    });
    static const bool dbg_on = getenv("LFB_PANEL_DBG") != nullptr;     // debug: per-phase clock64 deltas of thread 0 on stderr
    long long *dbg = nullptr;
    if (dbg_on) { LFB_CUDA(cudaMalloc(&dbg, 16 * sizeof(long long))); LFB_CUDA(cudaMemset(dbg, 0, 16 * sizeof(long long))); }
    cholqr128_kernel<T><<<1, NTH, smem, h.stream>>>(G, ldg, R, ldr, Rinv, ldri, guard, dbg);
    LFB_LAUNCH_CHECK(h);
    if (dbg_on) {
        long long hd[16];
        LFB_CUDA(cudaStreamSynchronize(h.stream));
        LFB_CUDA(cudaMemcpy(hd, dbg, sizeof hd, cudaMemcpyDeviceToHost));
        cudaFree(dbg);
        fprintf(stderr, "cholqr128 cycles: load %lld | cholesky %lld | X sweep %lld | X to smem %lld | guard %lld | emit %lld\n", hd[1] - hd[0],
                hd[2] - hd[1], hd[3] - hd[2], hd[4] - hd[3], hd[5] - hd[4], hd[6] - hd[5]);
    }
}

template <typename T>
void hr_panel128(lfb_handle &h, T *Atop, int64_t ld, const T *R, int64_t ldr, const T *Rinv, int64_t ldri, T *beta, T *M, int64_t ldm,
                 T *Tm, int64_t ldt, T *Vtop, int64_t ldv) {
    const size_t smem = sizeof(T) * (size_t)(PN * PLD + PN * (PN + 1) / 2 + 8 * PN);
    static DeviceOnce cfg;
    cfg.run(h.device, [&] {
        LFB_CUDA(cudaFuncSetAttribute(hr_panel128_kernel<T, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        LFB_CUDA(cudaFuncSetAttribute(hr_panel128_kernel<T, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        LFB_CUDA(cudaFuncSetAttribute(hr_panel128_kernel<T, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        LFB_CUDA(cudaFuncSetAttribute(hr_panel128_kernel<T, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    });
    static const bool dbg_on = getenv("LFB_PANEL_DBG") != nullptr;
    // Split launch (option hr_split): LU stage, then the M / top-block stage on this stream and the Z / T stage on the handle's second
    // side stream -- T is not needed until the reflector is applied, so its two sweeps run beside the tall GEMM V_2 = A_2 M and the
    // write-back of the panel.  hr_panel128_join() makes this stream wait for it.
    const bool split = h.opt.hr_split && Tm && !dbg_on && h.aux2_stream && h.aux2_stream != h.stream && !h.in_capture;
    if (split) {
        if (!h.hr_scratch) {
            LFB_CUDA(cudaMalloc(&h.hr_scratch, sizeof(double) * (size_t)(PN * PLD + 4 * PN)));
            LFB_CUDA(cudaEventCreateWithFlags(&h.hr_ev[0], cudaEventDisableTiming));
            LFB_CUDA(cudaEventCreateWithFlags(&h.hr_ev[1], cudaEventDisableTiming));
        }
        T *Qg = reinterpret_cast<T *>(h.hr_scratch);
        cudaStream_t s1 = h.stream, s2 = h.aux2_stream;
        hr_panel128_kernel<T, 1><<<1, NTH, smem, s1>>>(Atop, ld, R, ldr, Rinv, ldri, beta, M, ldm, Tm, ldt, Vtop, ldv, Qg, nullptr);
        LFB_LAUNCH_CHECK(h);
        LFB_CUDA(cudaEventRecord(h.hr_ev[0], s1));
        LFB_CUDA(cudaStreamWaitEvent(s2, h.hr_ev[0], 0));
        hr_panel128_kernel<T, 3><<<1, NTH, smem, s2>>>(Atop, ld, R, ldr, Rinv, ldri, beta, M, ldm, Tm, ldt, Vtop, ldv, Qg, nullptr);
        LFB_LAUNCH_CHECK(h);
        LFB_CUDA(cudaEventRecord(h.hr_ev[1], s2));
        h.hr_pending = true;
        hr_panel128_kernel<T, 2><<<1, NTH, smem, s1>>>(Atop, ld, R, ldr, Rinv, ldri, beta, M, ldm, Tm, ldt, Vtop, ldv, Qg, nullptr);
        LFB_LAUNCH_CHECK(h);
        return;
    }
    long long *dbg = nullptr;
    if (dbg_on) { LFB_CUDA(cudaMalloc(&dbg, 16 * sizeof(long long))); LFB_CUDA(cudaMemset(dbg, 0, 16 * sizeof(long long))); }
    hr_panel128_kernel<T, 0><<<1, NTH, smem, h.stream>>>(Atop, ld, R, ldr, Rinv, ldri, beta, M, ldm, Tm, ldt, Vtop, ldv, (T *)nullptr, dbg);
    LFB_LAUNCH_CHECK(h);
    if (dbg_on) {
        long long hd[16];
        LFB_CUDA(cudaStreamSynchronize(h.stream));
        LFB_CUDA(cudaMemcpy(hd, dbg, sizeof hd, cudaMemcpyDeviceToHost));
        cudaFree(dbg);
        fprintf(stderr, "hr_panel128 cycles: load %lld | Q R^-1 %lld | LU %lld | write-back + M init %lld | M sweep %lld | Z sweep %lld | T %lld | top block %lld\n",
                hd[1] - hd[0], hd[2] - hd[1], hd[3] - hd[2], hd[4] - hd[3], hd[5] - hd[4], hd[6] - hd[5], hd[7] - hd[6], hd[8] - hd[7]);
    }
}

// The stream that launched a split hr_panel128 waits here for its Z / T stage (T is complete afterwards).
void hr_panel128_join(lfb_handle &h) {
    if (!h.hr_pending) return;
    LFB_CUDA(cudaStreamWaitEvent(h.stream, h.hr_ev[1], 0));
    h.hr_pending = false;
}

#define INST(T)                                                                                                      \
    template void cholqr128<T>(lfb_handle &, const T *, int64_t, T *, int64_t, T *, int64_t, double *);               \
    template void hr_panel128<T>(lfb_handle &, T *, int64_t, const T *, int64_t, const T *, int64_t, T *, T *, int64_t, \
                                 T *, int64_t, T *, int64_t);
INST(float)
INST(double)
#undef INST

}  // namespace lfb
