// Column-major GEMM core of the engine: C = alpha op(A) op(B) + beta C.
//
// f64: warp-level FP64 tensor-core MMA (mma.sync m8n8k4 -> SASS DMMA.8x8x4, the native FP64 MMA
//      shape on sm_100a; tcgen05 has no f64 kind).  CTA tile 128x128x16, 8 warps, warp tile 64x32,
//      register double-buffered global->shared staging, conflict-free padded shared layouts for all
//      four operand layouts.  Optional split-K (atomics) for skinny outputs with a long K -- the
//      shape of every W = V^T C, Gram (V^T V) and TSQR product in this library.
// f32: register-tiled FFMA kernel (exact f32 arithmetic; no TF32 rounding on this generic path).
//
// This is the reference's XXX at src/reflection.rs:23 ("Can use matrix multiplication algorithm
// instead of iterative algorithm") made concrete: every trailing update funnels here.
#include <memory>

#include "common.cuh"

namespace lfb {

namespace {

constexpr int BM = 128, BN = 128, BK = 16;

struct GemmP {
    int M, N, K;
    int64_t lda, ldb, ldc;
    const void *A, *B;
    void *C;
    double alpha, beta;
    int lower_only;  // store only elements with m >= n
    int ksplit;      // K range per blockIdx.z (multiple of BK)
    int atomic;      // accumulate alpha*acc with atomicAdd (C pre-scaled by beta)
    int vecA, vecB, vecC;
    int64_t zstride; // deterministic split-K: split z writes its partial tile to C + z * zstride (a workspace slice)
};

__device__ __forceinline__ double2 ldg2(const double *p, bool v0, bool v1, bool vec) {
    double2 r = make_double2(0.0, 0.0);
    if (vec && v1) {
        r = *reinterpret_cast<const double2 *>(p);
    } else {
        if (v0) r.x = p[0];
        if (v1) r.y = p[1];
    }
    return r;
}

__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

// AMODE 0: A is M x K column-major (m contiguous)   -> smem [BK][BM+4]
// AMODE 1: A is K x M column-major (k contiguous)   -> smem [BM][BK+4]   (op(A) = A^T)
// BMODE 0: B is K x N column-major (k contiguous)   -> smem [BN][BK+4]
// BMODE 1: B is N x K column-major (n contiguous)   -> smem [BK][BN+4]   (op(B) = B^T)
template <int AMODE, int BMODE>
__global__ void __launch_bounds__(256) dgemm_kernel(GemmP p) {
    extern __shared__ double smem[];
    constexpr int A_LD = AMODE == 0 ? BM + 4 : BK + 4;
    constexpr int A_SZ = AMODE == 0 ? BK * (BM + 4) : BM * (BK + 4);
    constexpr int B_LD = BMODE == 0 ? BK + 4 : BN + 4;
    constexpr int B_SZ = BMODE == 0 ? BN * (BK + 4) : BK * (BN + 4);
    double *sA = smem;
    double *sB = smem + 2 * A_SZ;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gid = lane >> 2, tig = lane & 3;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    if (p.lower_only && m0 + BM <= n0) return;
    const int kbeg = blockIdx.z * p.ksplit;
    const int kend = min(p.K, kbeg + p.ksplit);
    if (kbeg >= kend) return;
    const int KT = (kend - kbeg + BK - 1) / BK;

    const double *__restrict__ A = static_cast<const double *>(p.A);
    const double *__restrict__ B = static_cast<const double *>(p.B);
    const bool vecA = p.vecA, vecB = p.vecB;

    double2 ra[4], rb[4];

    auto gload = [&](int kt) {
        const int k0 = kbeg + kt * BK;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (AMODE == 0) {
                int k = (tid >> 6) + 4 * i, m = 2 * (tid & 63);
                int gm = m0 + m, gk = k0 + k;
                bool kv = gk < kend;
                ra[i] = ldg2(A + gm + (int64_t)gk * p.lda, kv && gm < p.M, kv && gm + 1 < p.M, vecA);
            } else {
                int m = (tid >> 3) + 32 * i, k = 2 * (tid & 7);
                int gm = m0 + m, gk = k0 + k;
                bool mv = gm < p.M;
                ra[i] = ldg2(A + gk + (int64_t)gm * p.lda, mv && gk < kend, mv && gk + 1 < kend, vecA);
            }
            if (BMODE == 0) {
                int n = (tid >> 3) + 32 * i, k = 2 * (tid & 7);
                int gn = n0 + n, gk = k0 + k;
                bool nv = gn < p.N;
                rb[i] = ldg2(B + gk + (int64_t)gn * p.ldb, nv && gk < kend, nv && gk + 1 < kend, vecB);
            } else {
                int k = (tid >> 6) + 4 * i, n = 2 * (tid & 63);
                int gn = n0 + n, gk = k0 + k;
                bool kv = gk < kend;
                rb[i] = ldg2(B + gn + (int64_t)gk * p.ldb, kv && gn < p.N, kv && gn + 1 < p.N, vecB);
            }
        }
    };
    auto sstore = [&](int buf) {
        double *a = sA + buf * A_SZ, *b = sB + buf * B_SZ;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (AMODE == 0) {
                int k = (tid >> 6) + 4 * i, m = 2 * (tid & 63);
                *reinterpret_cast<double2 *>(a + k * A_LD + m) = ra[i];
            } else {
                int m = (tid >> 3) + 32 * i, k = 2 * (tid & 7);
                *reinterpret_cast<double2 *>(a + m * A_LD + k) = ra[i];
            }
            if (BMODE == 0) {
                int n = (tid >> 3) + 32 * i, k = 2 * (tid & 7);
                *reinterpret_cast<double2 *>(b + n * B_LD + k) = rb[i];
            } else {
                int k = (tid >> 6) + 4 * i, n = 2 * (tid & 63);
                *reinterpret_cast<double2 *>(b + k * B_LD + n) = rb[i];
            }
        }
    };

    const int wm0 = (warp & 1) * 64, wn0 = (warp >> 1) * 32;
    double acc[4][8][2];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[j][i][0] = acc[j][i][1] = 0.0;

    gload(0);
    sstore(0);
    __syncthreads();

    for (int kt = 0; kt < KT; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < KT) gload(kt + 1);
        const double *a = sA + buf * A_SZ, *b = sB + buf * B_SZ;
#pragma unroll
        for (int ks = 0; ks < BK / 4; ++ks) {
            double fa[8], fb[4];
#pragma unroll
            for (int i = 0; i < 8; ++i)
                fa[i] = AMODE == 0 ? a[(ks * 4 + tig) * A_LD + wm0 + 8 * i + gid]
                                   : a[(wm0 + 8 * i + gid) * A_LD + ks * 4 + tig];
#pragma unroll
            for (int j = 0; j < 4; ++j)
                fb[j] = BMODE == 0 ? b[(wn0 + 8 * j + gid) * B_LD + ks * 4 + tig]
                                   : b[(ks * 4 + tig) * B_LD + wn0 + 8 * j + gid];
            // MMA roles are swapped on purpose: the 8x4 "A" fragment carries op(B) (row = n) and the
            // 4x8 "B" fragment carries op(A) (col = m), so each lane's accumulator pair is two
            // CONSECUTIVE m of one column n -> 16-byte stores into column-major C.
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int i = 0; i < 8; ++i) dmma884(acc[j][i][0], acc[j][i][1], fb[j], fa[i]);
        }
        if (kt + 1 < KT) sstore(buf ^ 1);
        __syncthreads();
    }

    double *__restrict__ C = static_cast<double *>(p.C) + (int64_t)blockIdx.z * p.zstride;
    const double alpha = p.alpha, beta = p.beta;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int n = n0 + wn0 + 8 * j + gid;
        if (n >= p.N) continue;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int m = m0 + wm0 + 8 * i + 2 * tig;
            if (m >= p.M) continue;
            double *c = C + m + (int64_t)n * p.ldc;
            bool w0 = !p.lower_only || m >= n;
            bool w1 = (m + 1 < p.M) && (!p.lower_only || m + 1 >= n);
            double v0 = alpha * acc[j][i][0], v1 = alpha * acc[j][i][1];
            if (p.atomic) {
                if (w0) atomicAdd(c, v0);
                if (w1) atomicAdd(c + 1, v1);
            } else if (p.vecC && w0 && w1) {
                double2 o = make_double2(v0, v1);
                if (beta != 0.0) {
                    double2 old = *reinterpret_cast<double2 *>(c);
                    o.x += beta * old.x;
                    o.y += beta * old.y;
                }
                *reinterpret_cast<double2 *>(c) = o;
            } else {
                if (w0) c[0] = v0 + (beta != 0.0 ? beta * c[0] : 0.0);
                if (w1) c[1] = v1 + (beta != 0.0 ? beta * c[1] : 0.0);
            }
        }
    }
}

// ---- f64 v2: 16 warps (4 per scheduler), 4-stage cp.async pipeline -----------------------------
// ncu on the 8-warp kernel above (profiles/r1_dgemm_v1.md): DMMA pipe 77 % active, top stalls "wait"
// and short scoreboard with only 2 warps per scheduler.  This variant keeps the 128x128x16 CTA tile
// but runs 16 warps (warp tile 32x32, <= 128 registers) and feeds shared memory with 16-byte
// cp.async (zero-filled at the edges), one barrier per k-tile.  Needs 16-byte aligned operands.
constexpr int V2_STAGES = 4;

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem, int src_bytes) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

template <int AMODE, int BMODE>
__global__ void __launch_bounds__(512, 1) dgemm_v2_kernel(GemmP p) {
    extern __shared__ __align__(16) double smem[];
    constexpr int A_LD = AMODE == 0 ? BM + 4 : BK + 4;
    constexpr int A_SZ = AMODE == 0 ? BK * (BM + 4) : BM * (BK + 4);
    constexpr int B_LD = BMODE == 0 ? BK + 4 : BN + 4;
    constexpr int B_SZ = BMODE == 0 ? BN * (BK + 4) : BK * (BN + 4);
    double *sA = smem;
    double *sB = smem + V2_STAGES * A_SZ;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gid = lane >> 2, tig = lane & 3;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    if (p.lower_only && m0 + BM <= n0) return;
    const int kbeg = blockIdx.z * p.ksplit;
    const int kend = min(p.K, kbeg + p.ksplit);
    if (kbeg >= kend) return;
    const int KT = (kend - kbeg + BK - 1) / BK;

    const double *__restrict__ A = static_cast<const double *>(p.A);
    const double *__restrict__ B = static_cast<const double *>(p.B);

    auto issue = [&](int kt, int slot) {
        const int k0 = kbeg + kt * BK;
        double *a = sA + slot * A_SZ, *b = sB + slot * B_SZ;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            {
                int gm, gk, nval;
                double *dst;
                const double *src;
                if (AMODE == 0) {
                    int k = (tid >> 6) + 8 * i, m = 2 * (tid & 63);
                    gm = m0 + m; gk = k0 + k;
                    nval = gk < kend ? max(0, min(2, p.M - gm)) : 0;
                    dst = a + k * A_LD + m;
                    src = A + gm + (int64_t)gk * p.lda;
                } else {
                    int m = (tid >> 3) + 64 * i, k = 2 * (tid & 7);
                    gm = m0 + m; gk = k0 + k;
                    nval = gm < p.M ? max(0, min(2, kend - gk)) : 0;
                    dst = a + m * A_LD + k;
                    src = A + gk + (int64_t)gm * p.lda;
                }
                cp_async16(dst, nval ? src : A, nval * 8);
            }
            {
                int gn, gk, nval;
                double *dst;
                const double *src;
                if (BMODE == 0) {
                    int n = (tid >> 3) + 64 * i, k = 2 * (tid & 7);
                    gn = n0 + n; gk = k0 + k;
                    nval = gn < p.N ? max(0, min(2, kend - gk)) : 0;
                    dst = b + n * B_LD + k;
                    src = B + gk + (int64_t)gn * p.ldb;
                } else {
                    int k = (tid >> 6) + 8 * i, n = 2 * (tid & 63);
                    gn = n0 + n; gk = k0 + k;
                    nval = gk < kend ? max(0, min(2, p.N - gn)) : 0;
                    dst = b + k * B_LD + n;
                    src = B + gn + (int64_t)gk * p.ldb;
                }
                cp_async16(dst, nval ? src : B, nval * 8);
            }
        }
    };

    const int wm0 = (warp & 3) * 32, wn0 = (warp >> 2) * 32;
    double acc[4][4][2];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[j][i][0] = acc[j][i][1] = 0.0;

#pragma unroll
    for (int s = 0; s < V2_STAGES - 1; ++s) {
        if (s < KT) issue(s, s);
        cp_async_commit();
    }

    for (int kt = 0; kt < KT; ++kt) {
        cp_async_wait<V2_STAGES - 2>();   // k-tile kt has landed (for this thread's copies)
        __syncthreads();                  // ... and for everyone's; slot (kt-1)%S is free again
        {
            const int nk = kt + V2_STAGES - 1;
            if (nk < KT) issue(nk, nk % V2_STAGES);
            cp_async_commit();
        }
        const int slot = kt % V2_STAGES;
        const double *a = sA + slot * A_SZ, *b = sB + slot * B_SZ;
#pragma unroll
        for (int ks = 0; ks < BK / 4; ++ks) {
            double fa[4], fb[4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
                fa[i] = AMODE == 0 ? a[(ks * 4 + tig) * A_LD + wm0 + 8 * i + gid]
                                   : a[(wm0 + 8 * i + gid) * A_LD + ks * 4 + tig];
#pragma unroll
            for (int j = 0; j < 4; ++j)
                fb[j] = BMODE == 0 ? b[(wn0 + 8 * j + gid) * B_LD + ks * 4 + tig]
                                   : b[(ks * 4 + tig) * B_LD + wn0 + 8 * j + gid];
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int i = 0; i < 4; ++i) dmma884(acc[j][i][0], acc[j][i][1], fb[j], fa[i]);
        }
    }
    cp_async_wait<0>();

    double *__restrict__ C = static_cast<double *>(p.C) + (int64_t)blockIdx.z * p.zstride;
    const double alpha = p.alpha, beta = p.beta;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int n = n0 + wn0 + 8 * j + gid;
        if (n >= p.N) continue;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int m = m0 + wm0 + 8 * i + 2 * tig;
            if (m >= p.M) continue;
            double *c = C + m + (int64_t)n * p.ldc;
            bool w0 = !p.lower_only || m >= n;
            bool w1 = (m + 1 < p.M) && (!p.lower_only || m + 1 >= n);
            double v0 = alpha * acc[j][i][0], v1 = alpha * acc[j][i][1];
            if (p.atomic) {
                if (w0) atomicAdd(c, v0);
                if (w1) atomicAdd(c + 1, v1);
            } else if (p.vecC && w0 && w1) {
                double2 o = make_double2(v0, v1);
                if (beta != 0.0) {
                    double2 old = *reinterpret_cast<double2 *>(c);
                    o.x += beta * old.x;
                    o.y += beta * old.y;
                }
                *reinterpret_cast<double2 *>(c) = o;
            } else {
                if (w0) c[0] = v0 + (beta != 0.0 ? beta * c[0] : 0.0);
                if (w1) c[1] = v1 + (beta != 0.0 ? beta * c[1] : 0.0);
            }
        }
    }
}

// ---- f32: 64x64x16 tile, 256 threads, 4x4 micro-tile, exact FFMA ------------------------------
struct SgemmP {
    int M, N, K;
    int64_t a_sm, a_sk, b_sk, b_sn, ldc;  // element strides of op(A)[m,k], op(B)[k,n]
    const float *A, *B;
    float *C;
    float alpha, beta;
    int lower_only, ksplit, atomic;
    int64_t zstride;   // as GemmP::zstride
};

__global__ void __launch_bounds__(256) sgemm_kernel(SgemmP p) {
    constexpr int TM = 64, TN = 64, TK = 16;
    __shared__ __align__(16) float sA[TK][TM + 4];
    __shared__ __align__(16) float sB[TK][TN + 4];
    const int tid = threadIdx.x;
    const int m0 = blockIdx.x * TM, n0 = blockIdx.y * TN;
    if (p.lower_only && m0 + TM <= n0) return;
    const int kbeg = blockIdx.z * p.ksplit, kend = min(p.K, kbeg + p.ksplit);
    if (kbeg >= kend) return;
    const int tx = tid & 15, ty = tid >> 4;
    float acc[4][4] = {};
    for (int k0 = kbeg; k0 < kend; k0 += TK) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int e = tid + 256 * i;  // 1024 elements per operand tile
            int m, k;
            if (p.a_sm == 1) { m = e & 63; k = e >> 6; } else { k = e & 15; m = e >> 4; }
            int gm = m0 + m, gk = k0 + k;
            sA[k][m] = (gm < p.M && gk < kend) ? p.A[gm * p.a_sm + gk * p.a_sk] : 0.f;
            int n;
            if (p.b_sn == 1) { n = e & 63; k = e >> 6; } else { k = e & 15; n = e >> 4; }
            int gn = n0 + n;
            gk = k0 + k;
            sB[k][n] = (gn < p.N && gk < kend) ? p.B[gk * p.b_sk + gn * p.b_sn] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < TK; ++k) {
            float4 a = *reinterpret_cast<const float4 *>(&sA[k][tx * 4]);
            float4 b = *reinterpret_cast<const float4 *>(&sB[k][ty * 4]);
            float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[j][i] = fmaf(av[i], bv[j], acc[j][i]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        int n = n0 + ty * 4 + j;
        if (n >= p.N) continue;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int m = m0 + tx * 4 + i;
            if (m >= p.M) continue;
            if (p.lower_only && m < n) continue;
            float *c = p.C + (int64_t)blockIdx.z * p.zstride + m + (int64_t)n * p.ldc;
            float v = p.alpha * acc[j][i];
            if (p.atomic) atomicAdd(c, v);
            else *c = v + (p.beta != 0.f ? p.beta * *c : 0.f);
        }
    }
}

template <typename T>
__global__ void scale_kernel(T *C, int64_t M, int64_t N, int64_t ldc, T beta, int lower_only) {
    int64_t m = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (m >= M) return;
    for (int64_t n = blockIdx.y; n < N; n += gridDim.y) {
        if (lower_only && m < n) continue;
        T *c = C + m + n * ldc;
        *c = beta == T(0) ? T(0) : beta * *c;
    }
}

template <int AM, int BMo>
void launch_dgemm(lfb_handle &h, const GemmP &p, dim3 grid) {
    constexpr int A_SZ = AM == 0 ? BK * (BM + 4) : BM * (BK + 4);
    constexpr int B_SZ = BMo == 0 ? BN * (BK + 4) : BK * (BN + 4);
    constexpr size_t smem = sizeof(double) * 2 * (A_SZ + B_SZ);
    constexpr size_t smem2 = sizeof(double) * V2_STAGES * (A_SZ + B_SZ);
    static DeviceOnce configured;   // function attributes are per device
    configured.run(h.device, [&] {
        LFB_CUDA(cudaFuncSetAttribute(dgemm_kernel<AM, BMo>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        LFB_CUDA(cudaFuncSetAttribute(dgemm_v2_kernel<AM, BMo>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
    });
    if (p.vecA && p.vecB && h.opt.gemm_v2) {
        dgemm_v2_kernel<AM, BMo><<<grid, 512, smem2, h.stream>>>(p);
    } else {
        dgemm_kernel<AM, BMo><<<grid, 256, smem, h.stream>>>(p);
    }
    LFB_LAUNCH_CHECK(h);
}

// C = beta C + sum_z W_z in the fixed order z = 0, 1, ...: second stage of the deterministic split-K of the fallback kernels
template <typename T>
__global__ void splitk_reduce_any_kernel(const T *__restrict__ W, int64_t ldw, int64_t zstride, int splits, T *__restrict__ C, int64_t M,
                                         int64_t N, int64_t ldc, T beta, int lower_only) {
    const int64_t m = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (m >= M) return;
    for (int64_t n = blockIdx.y; n < N; n += gridDim.y) {
        if (lower_only && m < n) continue;
        const T *w = W + m + n * ldw;
        T acc = T(0);
        for (int z = 0; z < splits; ++z) acc += w[(int64_t)z * zstride];
        T *c = C + m + n * ldc;
        *c = beta == T(0) ? acc : beta * *c + acc;
    }
}

inline int choose_splits(lfb_handle &h, int64_t tiles, int64_t K, int bk) {
    if (!h.opt.gemm_splitk) return 1;
    if (tiles >= h.sm_count) return 1;
    int64_t want = (2 * h.sm_count + tiles - 1) / tiles;
    int64_t maxs = K / (8 * bk);  // at least 8 k-tiles per split
    if (maxs < 1) maxs = 1;
    return (int)(want < maxs ? want : maxs);
}

}  // namespace

// Defined in gemm_tma.cu: returns true if it handled the call.
bool dgemm_tma_try(lfb_handle &h, int ta, int tb, int64_t M, int64_t N, int64_t K, double alpha, const double *A,
                   int64_t lda, const double *B, int64_t ldb, double beta, double *C, int64_t ldc, int lower_only);

static void dgemm_impl(lfb_handle &h, int ta, int tb, int64_t M, int64_t N, int64_t K, double alpha, const double *A,
                       int64_t lda, const double *B, int64_t ldb, double beta, double *C, int64_t ldc, int lower_only);

template <>
void gemm<double>(lfb_handle &h, int ta, int tb, int64_t M, int64_t N, int64_t K, double alpha, const double *A,
                  int64_t lda, const double *B, int64_t ldb, double beta, double *C, int64_t ldc, int lower_only) {
    if (!h.prof_on) {
        dgemm_impl(h, ta, tb, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, lower_only);
        return;
    }
    cudaEvent_t e0 = h.prof_event(), e1 = h.prof_event();
    LFB_CUDA(cudaEventRecord(e0, h.stream));
    dgemm_impl(h, ta, tb, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, lower_only);
    LFB_CUDA(cudaEventRecord(e1, h.stream));
    h.prof_flops += (lower_only ? 1.0 : 2.0) * (double)M * (double)N * (double)K;
}

static void dgemm_impl(lfb_handle &h, int ta, int tb, int64_t M, int64_t N, int64_t K, double alpha, const double *A,
                       int64_t lda, const double *B, int64_t ldb, double beta, double *C, int64_t ldc, int lower_only) {
    if (M <= 0 || N <= 0) return;
    if (K <= 0 || alpha == 0.0) {
        if (beta != 1.0) {
            dim3 g((unsigned)cdiv(M, 256), (unsigned)(N < 65535 ? N : 65535));
            scale_kernel<double><<<g, 256, 0, h.stream>>>(C, M, N, ldc, beta, lower_only);
            LFB_LAUNCH_CHECK(h);
        }
        return;
    }
    if (h.opt.gemm_tma && dgemm_tma_try(h, ta, tb, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, lower_only)) return;

    GemmP p;
    p.M = (int)M; p.N = (int)N; p.K = (int)K;
    p.lda = lda; p.ldb = ldb; p.ldc = ldc;
    p.A = A; p.B = B; p.C = C;
    p.alpha = alpha; p.beta = beta;
    p.lower_only = lower_only;
    p.vecA = ((uintptr_t)A % 16 == 0) && (lda % 2 == 0);
    p.vecB = ((uintptr_t)B % 16 == 0) && (ldb % 2 == 0);
    p.vecC = ((uintptr_t)C % 16 == 0) && (ldc % 2 == 0);
    int64_t tm = cdiv(M, BM), tn = cdiv(N, BN);
    int splits = choose_splits(h, tm * tn, K, BK);
    p.ksplit = (int)round_up(cdiv(K, splits), BK);
    splits = (int)cdiv(K, p.ksplit);
    p.zstride = 0;
    std::unique_ptr<DevBuf<double>> work;
    const int64_t ldw = round_up(M, 2);
    const bool det = splits > 1 && h.opt.gemm_deterministic;     // slices + ordered reduce instead of atomicAdd (bit-reproducible)
    if (det) {
        work.reset(new DevBuf<double>(h, (size_t)ldw * N * splits));
        p.C = work->get(); p.ldc = ldw; p.zstride = ldw * N;
        p.beta = 0.0; p.atomic = 0; p.vecC = 1;
    } else {
        p.atomic = splits > 1;
        if (p.atomic && beta != 1.0) {
            dim3 g((unsigned)cdiv(M, 256), (unsigned)(N < 65535 ? N : 65535));
            scale_kernel<double><<<g, 256, 0, h.stream>>>(C, M, N, ldc, beta, lower_only);
            LFB_LAUNCH_CHECK(h);
        }
    }
    dim3 grid((unsigned)tm, (unsigned)tn, (unsigned)splits);
    if (ta == 0 && tb == 0) launch_dgemm<0, 0>(h, p, grid);
    else if (ta == 1 && tb == 0) launch_dgemm<1, 0>(h, p, grid);
    else if (ta == 0 && tb == 1) launch_dgemm<0, 1>(h, p, grid);
    else launch_dgemm<1, 1>(h, p, grid);
    if (det) {
        dim3 g((unsigned)cdiv(M, 128), (unsigned)(N < 65535 ? N : 65535));
        splitk_reduce_any_kernel<double><<<g, 128, 0, h.stream>>>(work->get(), ldw, p.zstride, splits, C, M, N, ldc, beta, lower_only);
        LFB_LAUNCH_CHECK(h);
    }
}

// Defined in gemm_tf32.cu (tcgen05 / TMEM, 3xTF32): returns true if it handled the call.
bool sgemm_tc_try(lfb_handle &h, int ta, int tb, int64_t M, int64_t N, int64_t K, float alpha, const float *A, int64_t lda,
                  const float *B, int64_t ldb, float beta, float *C, int64_t ldc, int lower_only);

template <>
void gemm<float>(lfb_handle &h, int ta, int tb, int64_t M, int64_t N, int64_t K, float alpha, const float *A,
                 int64_t lda, const float *B, int64_t ldb, float beta, float *C, int64_t ldc, int lower_only) {
    if (M <= 0 || N <= 0) return;
    if (K <= 0 || alpha == 0.f) {
        if (beta != 1.f) {
            dim3 g((unsigned)cdiv(M, 256), (unsigned)(N < 65535 ? N : 65535));
            scale_kernel<float><<<g, 256, 0, h.stream>>>(C, M, N, ldc, beta, lower_only);
            LFB_LAUNCH_CHECK(h);
        }
        return;
    }
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (h.prof_on) {
        e0 = h.prof_event(); e1 = h.prof_event();
        LFB_CUDA(cudaEventRecord(e0, h.stream));
        h.prof_flops += (lower_only ? 1.0 : 2.0) * (double)M * (double)N * (double)K;
    }
    struct ProfEnd {   // closes the profiler's event pair on every exit path
        lfb_handle &h; cudaEvent_t e;
        ~ProfEnd() { if (e) cudaEventRecord(e, h.stream); }
    } prof_end{h, e1};
    if (sgemm_tc_try(h, ta, tb, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, lower_only)) return;
    SgemmP p;
    p.M = (int)M; p.N = (int)N; p.K = (int)K;
    p.a_sm = ta ? lda : 1; p.a_sk = ta ? 1 : lda;
    p.b_sk = tb ? ldb : 1; p.b_sn = tb ? 1 : ldb;
    p.ldc = ldc; p.A = A; p.B = B; p.C = C; p.alpha = alpha; p.beta = beta; p.lower_only = lower_only;
    int64_t tm = cdiv(M, 64), tn = cdiv(N, 64);
    int splits = choose_splits(h, tm * tn, K, 16);
    p.ksplit = (int)round_up(cdiv(K, splits), 16);
    splits = (int)cdiv(K, p.ksplit);
    p.zstride = 0;
    std::unique_ptr<DevBuf<float>> work;
    const int64_t ldw = round_up(M, 4);
    const bool det = splits > 1 && h.opt.gemm_deterministic;
    if (det) {
        work.reset(new DevBuf<float>(h, (size_t)ldw * N * splits));
        p.C = work->get(); p.ldc = ldw; p.zstride = ldw * N;
        p.beta = 0.f; p.atomic = 0;
    } else {
        p.atomic = splits > 1;
        if (p.atomic && beta != 1.f) {
            dim3 g((unsigned)cdiv(M, 256), (unsigned)(N < 65535 ? N : 65535));
            scale_kernel<float><<<g, 256, 0, h.stream>>>(C, M, N, ldc, beta, lower_only);
            LFB_LAUNCH_CHECK(h);
        }
    }
    dim3 grid((unsigned)tm, (unsigned)tn, (unsigned)splits);
    sgemm_kernel<<<grid, 256, 0, h.stream>>>(p);
    LFB_LAUNCH_CHECK(h);
    if (det) {
        dim3 g((unsigned)cdiv(M, 128), (unsigned)(N < 65535 ? N : 65535));
        splitk_reduce_any_kernel<float><<<g, 128, 0, h.stream>>>(work->get(), ldw, p.zstride, splits, C, M, N, ldc, beta, lower_only);
        LFB_LAUNCH_CHECK(h);
    }
}

}  // namespace lfb
