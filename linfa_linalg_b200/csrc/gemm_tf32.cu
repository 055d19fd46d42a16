// f32 trailing updates on the 5th-generation tensor cores: tcgen05.mma.kind::tf32 with the accumulator in TMEM, operands
// staged by TMA (SASS: UTMALDG / UTCMMA / LDTM), 3xTF32 error compensation so that the result stays at f32 accuracy.
//
//   C (M x N, column-major) = alpha * A^T B + beta * C,   A stored K x M, B stored K x N  (both K-contiguous: "TN")
//
// This is the shape every large f32 GEMM of the factorisations is brought to: W = V^T C directly, C -= V W and the Cholesky
// SYRK through a transposed copy of the (skinny) panel -- so ONE operand layout (K-major, 128-byte rows, SWIZZLE_128B, the
// canonical UMMA K-major layout that TMA produces by itself) covers the path.
//
// 3xTF32: the tensor core multiplies TF32 (10 explicit mantissa bits).  Every f32 operand x is split in shared memory into
// hi = x with the low 13 mantissa bits cleared and lo = x - hi (exact), itself cut to TF32; the product is accumulated in
// f32 as  a_lo b_hi + a_hi b_lo + a_hi b_hi  (the a_lo b_lo term, <= 2^-22 relative, is dropped).
//
// Accumulation.  Measured on B200 (tests/test_gpu_f32_tc.py, profiles/r2_f32_tc.md): the tensor core adds into the TMEM
// accumulator with TRUNCATION, so a sum of K positive products accumulated entirely in TMEM is biased by ~(3K/8) 2^-24
// relative (4.4e-5 at K = 2048: 150x worse than FFMA, however well the operands are split).  The accumulator is therefore
// PROMOTED: every CHUNK = 2 k-tiles (64 values of K, 24 MMAs) go into a fresh TMEM buffer (two buffers, 2 x 128 columns),
// which the four epilogue warps drain with tcgen05.ld and add -- round to nearest, CUDA cores -- into 128 f32 registers per
// thread while the tensor core fills the other buffer.  The truncation bias then no longer grows with K: <= 24 2^-24 of one
// chunk's partial sum.  Stated tolerance: |C - exact| <= (2^-21 + 16 2^-24) sum_k |a_k b_k| + f32 summation of K/64 chunk
// sums; the parity tests hold the f32 factorisations that run on this kernel to the same c n eps_32 bounds as the FFMA path.
// Option "sgemm_tc": 1 = 3xTF32 (default for large aligned TN products), 0 = FFMA kernel only, 2 = single-pass TF32
// (benchmark yard-stick, ~1e-3 relative: never a default).
//
// CTA = 128 x 128 output tile, 6 warps, one role each (no CTA-wide barrier in the main loop):
//   warp 0      TMA producer: per k-tile (32 f32 = one 128-byte swizzle row) the raw A and B tiles -> shared, full_raw[s]
//   warps 2-5   splitters: raw tile -> hi (in place) and lo (second tile), fence.proxy.async, arrive on full_split[s];
//               afterwards the same four warps are the epilogue (tcgen05.ld 32 lanes x 32 columns each, alpha/beta, coalesced
//               column-major stores: lane = row)
//   warp 1      TMEM allocation; one lane issues 3 tcgen05.mma per 8-wide K slice; tcgen05.commit frees the stage
//               (empty[s]) and finally publishes the accumulator (tmem_full)
#include <cuda.h>

#include <memory>

#include "common.cuh"

namespace lfb {
namespace {

constexpr int BM = 128, BN = 128, BKF = 32;         // BKF f32 = 128 bytes = one swizzle row
constexpr int TILE_BYTES = BM * BKF * 4;            // 16 KiB (BM == BN)
constexpr int STAGE_BYTES = 4 * TILE_BYTES;         // A hi | A lo | B hi | B lo
constexpr int STAGES = 3;
constexpr int NTHREADS = 192;
constexpr int CHUNK = 2;                            // k-tiles accumulated in TMEM before promotion to registers (3xTF32 mode)

struct TcP {
    int M, N, K;
    int64_t ldc;
    float *C;
    float alpha, beta;
    int lower_only, mode;   // mode: 1 = single TF32 pass, 3 = 3xTF32
    int kt_per_split;       // split-K (skinny outputs): CTA z covers k-tiles [z * kt_per_split, ...) and writes its partial tile
    int64_t zstride;        //   to C + z * zstride (a workspace slice, beta = 0); a reduce kernel sums the slices in order
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done = 0;
    long long t0 = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (!done) {   // never hang the GPU on a protocol bug: trap after ~2 s
            if (t0 == 0) t0 = clock64();
            else if (clock64() - t0 > 4000000000LL) __trap();
        }
    }
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(dst)),
        "l"((uint64_t)map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
        : "memory");
}

// Shared-memory matrix descriptor of a K-major, 128-byte-swizzled tile (rows of 128 bytes, 8-row groups 1024 bytes apart):
// start address >> 4 | LBO (ignored for swizzled K-major; 1) | SBO = 1024 >> 4 | version 1 (sm_100) | SWIZZLE_128B.
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t saddr) {
    uint64_t d = (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// Instruction descriptor: D = F32, A = B = TF32, both K-major, N = BN, M = 128, dense, no negate.
__device__ __forceinline__ uint32_t umma_idesc_tf32() {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// hi = x with the 13 low mantissa bits cleared (exactly a TF32 value whatever rounding the tensor core applies to its
// inputs), lo = x - hi (exact in f32), cut to TF32 the same way.
__device__ __forceinline__ void split_tf32(float x, float &hi, float &lo) {
    hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
    lo = __uint_as_float(__float_as_uint(x - hi) & 0xFFFFE000u);
}

__global__ void __launch_bounds__(NTHREADS, 1)
sgemm_tf32_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, TcP p) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // swizzle atoms are 1024-byte aligned
    uint64_t *full_raw = reinterpret_cast<uint64_t *>(smem + STAGES * STAGE_BYTES);
    uint64_t *full_split = full_raw + STAGES;
    uint64_t *empty = full_split + STAGES;
    uint64_t *tmem_full = empty + STAGES;               // [2]
    uint64_t *tmem_empty = tmem_full + 2;               // [2]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_empty + 2);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    if (p.lower_only && m0 + BM <= n0) return;                    // tile strictly above the diagonal
    const int KT_all = (p.K + BKF - 1) / BKF;
    const int kt0 = blockIdx.z * p.kt_per_split;                  // first k-tile of this split
    const int KT = min(KT_all - kt0, p.kt_per_split);
    const bool x3 = p.mode == 3;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_raw[s], 1);
            mbar_init(&full_split[s], 4);                         // one arrive per splitter warp
            mbar_init(&empty[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&tmem_full[b], 1);
            mbar_init(&tmem_empty[b], 4);                         // one arrive per epilogue warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {                                              // TMEM: two accumulator buffers of BN f32 columns x 128 lanes
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(2 * BN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            for (int kt = 0; kt < KT; ++kt) {
                const int slot = kt % STAGES;
                const uint32_t phase = (kt / STAGES) & 1;
                mbar_wait(&empty[slot], phase ^ 1);
                mbar_expect_tx(&full_raw[slot], 2 * TILE_BYTES);
                unsigned char *st = smem + slot * STAGE_BYTES;
                tma_load_2d(st, &mapA, (kt0 + kt) * BKF, m0, &full_raw[slot]);                    // A hi (raw)
                tma_load_2d(st + 2 * TILE_BYTES, &mapB, (kt0 + kt) * BKF, n0, &full_raw[slot]);   // B hi (raw)
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: one thread on behalf of the CTA =====
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_tf32();
            for (int kt = 0; kt < KT; ++kt) {
                const int slot = kt % STAGES;
                const uint32_t phase = (kt / STAGES) & 1;
                // 3xTF32: chunk c of CHUNK k-tiles accumulates from zero into TMEM buffer c & 1 (drained by the epilogue warps)
                const int c = x3 ? kt / CHUNK : 0, b = c & 1;
                const bool chunk_first = x3 ? (kt % CHUNK == 0) : (kt == 0);
                const bool chunk_last = x3 ? (kt % CHUNK == CHUNK - 1 || kt == KT - 1) : (kt == KT - 1);
                if (x3 && chunk_first) {
                    mbar_wait(&tmem_empty[b], ((c >> 1) & 1) ^ 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                }
                mbar_wait(x3 ? &full_split[slot] : &full_raw[slot], phase);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t tmem_d = tmem_base + (uint32_t)(b * BN);
                const uint32_t sa = smem_u32(smem + slot * STAGE_BYTES);
                const uint64_t a_hi = umma_desc_k_sw128(sa), a_lo = umma_desc_k_sw128(sa + TILE_BYTES);
                const uint64_t b_hi = umma_desc_k_sw128(sa + 2 * TILE_BYTES), b_lo = umma_desc_k_sw128(sa + 3 * TILE_BYTES);
#pragma unroll
                for (int ks = 0; ks < BKF / 8; ++ks) {            // UMMA_K = 8 TF32 = 32 bytes: +2 in the (addr >> 4) field
                    const uint64_t o = (uint64_t)(2 * ks);
                    const uint32_t acc = (chunk_first && ks == 0) ? 0u : 1u;
                    if (x3) {
                        umma_tf32(tmem_d, a_lo + o, b_hi + o, idesc, acc);
                        umma_tf32(tmem_d, a_hi + o, b_lo + o, idesc, 1u);
                        umma_tf32(tmem_d, a_hi + o, b_hi + o, idesc, 1u);
                    } else {
                        umma_tf32(tmem_d, a_hi + o, b_hi + o, idesc, acc);
                    }
                }
                umma_commit(&empty[slot]);                        // arrives when the MMAs above have read the stage
                if (chunk_last) umma_commit(&tmem_full[b]);       // ... and when this chunk's accumulator is complete
            }
        }
    } else {
        // ===== splitters (main loop), then epilogue =====
        const int st_id = tid - 64;                               // 0..127
        const int q = warp & 3;                                   // TMEM lane quarter this warp may read
        const uint32_t lane_base = tmem_base + ((uint32_t)(32 * q) << 16);
        float acc[BN];                                            // 3xTF32: running f32 sums of this thread's row (promoted chunks)
#pragma unroll
        for (int j = 0; j < BN; ++j) acc[j] = 0.f;
        // drain chunk c: TMEM buffer c & 1 -> registers, round-to-nearest adds; then hand the buffer back to the MMA warp
        auto promote = [&](int c) {
            const int b = c & 1;
            mbar_wait(&tmem_full[b], (c >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int g = 0; g < BN / 32; ++g) {
                uint32_t v[32];
                tmem_ld_32x32(lane_base + (uint32_t)(b * BN + 32 * g), v);
#pragma unroll
                for (int j = 0; j < 32; ++j) acc[32 * g + j] += __uint_as_float(v[j]);
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[b]);
        };
        if (x3) {
            for (int kt = 0; kt < KT; ++kt) {
                const int slot = kt % STAGES;
                const uint32_t phase = (kt / STAGES) & 1;
                mbar_wait(&full_raw[slot], phase);
                unsigned char *st = smem + slot * STAGE_BYTES;
#pragma unroll
                for (int op = 0; op < 2; ++op) {                  // A then B: raw tile -> hi in place, lo next to it
                    float4 *hi = reinterpret_cast<float4 *>(st + op * 2 * TILE_BYTES);
                    float4 *lo = reinterpret_cast<float4 *>(st + op * 2 * TILE_BYTES + TILE_BYTES);
#pragma unroll
                    for (int i = 0; i < TILE_BYTES / 16 / 128; ++i) {
                        const int idx = st_id + i * 128;
                        const float4 x = hi[idx];
                        float4 h, l;
                        split_tf32(x.x, h.x, l.x);
                        split_tf32(x.y, h.y, l.y);
                        split_tf32(x.z, h.z, l.z);
                        split_tf32(x.w, h.w, l.w);
                        hi[idx] = h;
                        lo[idx] = l;
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core
                __syncwarp();
                if (lane == 0) mbar_arrive(&full_split[slot]);
                // one chunk behind the splitting: chunk c - 1's MMAs only need tiles that are already split, so this
                // wait cannot dead-lock, and the tensor core works on chunk c meanwhile
                if (kt % CHUNK == CHUNK - 1 && kt / CHUNK >= 1) promote(kt / CHUNK - 1);
            }
            const int nchunks = (KT + CHUNK - 1) / CHUNK;
            // chunks not drained inside the loop: the last one, and the one before it when the last chunk is partial
            if (nchunks >= 2 && KT % CHUNK != 0) promote(nchunks - 2);
            promote(nchunks - 1);
        } else {
            mbar_wait(&tmem_full[0], 0);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int g = 0; g < BN / 32; ++g) {
                uint32_t v[32];
                tmem_ld_32x32(lane_base + (uint32_t)(32 * g), v);
#pragma unroll
                for (int j = 0; j < 32; ++j) acc[32 * g + j] = __uint_as_float(v[j]);
            }
        }
        const int m = m0 + 32 * q + lane;
        const float alpha = p.alpha, beta = p.zstride ? 0.f : p.beta;
        float *C = p.C + (int64_t)blockIdx.z * p.zstride;
        if (m < p.M) {
            // lanes = consecutive rows: every access below is a coalesced 128-byte line per warp.  The old values of C are
            // fetched 32 columns at a time BEFORE any store of the group: a load-store-load chain (the compiler cannot
            // prove that the columns do not alias) made the epilogue 128 dependent DRAM round trips per thread, 90 us per
            // tile with beta = 1 (profiles/r2_f32_tc.md).
#pragma unroll
            for (int g = 0; g < BN / 32; ++g) {
                float old[32];
                if (beta != 0.f) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int n = n0 + 32 * g + j;
                        old[j] = (n < p.N && (!p.lower_only || m >= n)) ? C[m + (int64_t)n * p.ldc] : 0.f;
                    }
                }
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int n = n0 + 32 * g + j;
                    if (n < p.N && (!p.lower_only || m >= n)) {
                        const float r = alpha * acc[32 * g + j];
                        C[m + (int64_t)n * p.ldc] = beta == 0.f ? r : fmaf(beta, old[j], r);
                    }
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * BN) : "memory");
    }
}

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                             const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeFn get_encode_f32() {
    static const EncodeFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            return (EncodeFn)p;
        cudaGetLastError();
        return (EncodeFn) nullptr;
    }();
    return fn;
}

// K-major f32 operand stored K x R column-major (leading dimension ld): dim0 = K (contiguous), dim1 = R; box 32 x 128.
bool make_map_f32(CUtensorMap *map, const float *ptr, int64_t K, int64_t R, int64_t ld) {
    EncodeFn enc = get_encode_f32();
    if (!enc) return false;
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)R};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
    cuuint32_t box[2] = {(cuuint32_t)BKF, (cuuint32_t)BM};
    cuuint32_t es[2] = {1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)ptr, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// C = beta C + sum_z W_z in the fixed order z = 0, 1, ...: the deterministic second stage of split-K.
__global__ void splitk_reduce_f32_kernel(const float *__restrict__ W, int64_t ldw, int64_t zstride, int splits, float *__restrict__ C,
                                         int64_t M, int64_t N, int64_t ldc, float beta, int lower_only) {
    const int64_t m = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (m >= M) return;
    for (int64_t n = blockIdx.y; n < N; n += gridDim.y) {
        if (lower_only && m < n) continue;
        const float *w = W + m + n * ldw;
        float acc = 0.f;
        for (int z = 0; z < splits; ++z) acc += w[(int64_t)z * zstride];
        float *c = C + m + n * ldc;
        *c = beta == 0.f ? acc : fmaf(beta, *c, acc);
    }
}

}  // namespace

// Returns true if the product was taken by the tensor-core kernel (large, aligned, TN); false -> the caller's FFMA kernel.
bool sgemm_tc_try(lfb_handle &h, int ta, int tb, int64_t M, int64_t N, int64_t K, float alpha, const float *A, int64_t lda,
                  const float *B, int64_t ldb, float beta, float *C, int64_t ldc, int lower_only) {
    if (h.opt.sgemm_tc == 0 || ta != 1 || tb != 0) return false;
    if (M < 128 || N < 128 || K < 32) return false;
    if (((uintptr_t)A & 15) || ((uintptr_t)B & 15) || (lda & 3) || (ldb & 3)) return false;   // TMA: 16-byte bases and strides
    if (M >= (1LL << 31) || N >= (1LL << 31) || K >= (1LL << 31)) return false;
    alignas(64) CUtensorMap ma, mb;
    if (!make_map_f32(&ma, A, K, M, lda) || !make_map_f32(&mb, B, K, N, ldb)) return false;
    TcP p;
    p.M = (int)M; p.N = (int)N; p.K = (int)K;
    p.ldc = ldc; p.C = C; p.alpha = alpha; p.beta = beta; p.lower_only = lower_only;
    p.mode = h.opt.sgemm_tc == 2 ? 1 : 3;
    constexpr size_t smem = STAGES * STAGE_BYTES + (3 * STAGES + 4) * sizeof(uint64_t) + 16 + 1024;
    static DeviceOnce cfg;   // function attributes are per device
    cfg.run(h.device, [&] {
        LFB_CUDA(cudaFuncSetAttribute(sgemm_tf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    });
    // Split-K for skinny outputs (W = V^T C: one row of tiles, K = the matrix height): enough (tile, split) items for ~4 per
    // SM, at least 16 k-tiles each.  Besides the wave balance this keeps CTAs short, so that the cluster launches of the
    // look-ahead panel kernel find a drained GPC quickly.  Deterministic: slices + ordered reduce, as in the f64 kernel.
    const int64_t tiles = cdiv(M, BM) * cdiv(N, BN), KT = cdiv(K, BKF);
    int64_t splits = 1;
    if (h.opt.gemm_splitk && tiles < 2 * h.sm_count) splits = std::max<int64_t>(1, std::min(cdiv(4 * h.sm_count, tiles), KT / 16));
    p.kt_per_split = (int)cdiv(KT, splits);
    splits = cdiv(KT, p.kt_per_split);
    p.zstride = 0;
    const int64_t ldw = round_up(M, 4);
    std::unique_ptr<DevBuf<float>> work;
    if (splits > 1) {
        work.reset(new DevBuf<float>(h, (size_t)ldw * N * splits));
        p.zstride = ldw * N;
        p.C = work->get();
        p.ldc = ldw;
    }
    dim3 grid((unsigned)cdiv(M, BM), (unsigned)cdiv(N, BN), (unsigned)splits);
    sgemm_tf32_kernel<<<grid, NTHREADS, smem, h.stream>>>(ma, mb, p);
    LFB_LAUNCH_CHECK(h);
    if (splits > 1) {
        dim3 g((unsigned)cdiv(M, 128), (unsigned)(N < 65535 ? N : 65535));
        splitk_reduce_f32_kernel<<<g, 128, 0, h.stream>>>(work->get(), ldw, p.zstride, (int)splits, C, M, N, ldc, beta, lower_only);
        LFB_LAUNCH_CHECK(h);
    }
    return true;
}

}  // namespace lfb
