// TMA-fed, warp-specialised FP64 GEMM (the aligned fast path of lfb::gemm<double>).
//
//   * one producer warp: cp.async.bulk.tensor (TMA, SASS UTMALDG) into a 5-stage 128B-swizzled
//     shared-memory ring, completion on `full` mbarriers (expect_tx);
//   * sixteen consumer warps (4 per scheduler, warp tile 32x32): DMMA.8x8x4 straight from the
//     swizzled tiles, release a stage with one `empty` mbarrier arrive per warp.
// There is NO CTA-wide barrier in the main loop: the ncu profile of the barrier-per-k-tile kernels
// (profiles/) shows the restart bubble after each __syncthreads as their main loss; here warps
// drift apart and the FP64 tensor pipe always has a ready warp.
//
// Bank-conflict-free fragment loads from the TMA layout (no padding is possible with TMA): the MMA
// does not care which physical row a fragment lane holds as long as the epilogue agrees, so rows
// are permuted inside each 16-row group:
//   K-major tile  (box 16k x 128 rows, 128B rows):  row = 16*grp + 2*gid + parity
//   MN-major tile (8 boxes of 16 rows(mn) x 16 k):  row = 16*grp + perm(gid) + 4*parity,
//                                                   perm(g) = g0 | g2<<1 | g1<<3
// which makes the 16 lanes of each 64-bit shared-load phase hit 16 distinct 8-byte slots under the
// 128B swizzle (chunk index ^= row%8).
#include <cuda.h>

#include <memory>

#include "common.cuh"

namespace lfb {
namespace {

constexpr int BM = 128, BN = 128, BK = 16;
constexpr int STAGES = 5;
constexpr int TILE_BYTES = BM * BK * 8;           // 16 KB per operand per stage
constexpr int STAGE_BYTES = 2 * TILE_BYTES;       // 32 KB
constexpr int NCONS = 16;                         // consumer warps
constexpr int NTHREADS = (NCONS + 1) * 32;

struct TmaP {
    int M, N, K;
    int64_t ldc;
    double *C;
    double alpha, beta;
    int lower_only, ksplit, atomic, vecC;
    int ptn;            // dgemm_tma2_kernel with tri: number of 64-column tile columns
    int tri;            // lower_only on a square tile grid: blockIdx.x enumerates the LIVE tiles (i >= j) row by row -- no dead CTAs
    int partial;        // deterministic split-K: split z writes alpha * (its partial product) to C + z * zstride (ld = ldc), no atomics
    int64_t zstride;
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
__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

__device__ __forceinline__ int perm8(int g) { return (g & 1) | (((g >> 2) & 1) << 1) | (((g >> 1) & 1) << 3); }

// physical row (0..31 inside the warp tile) held by lane-group g of fragment f
template <int KMAJOR>
__device__ __forceinline__ int frag_row(int f, int g) {
    return KMAJOR ? 16 * (f >> 1) + 2 * g + (f & 1) : 16 * (f >> 1) + perm8(g) + 4 * (f & 1);
}

// AK / BKm: 1 if that operand's tile is K-major in shared memory.
//   A: AMODE 1 (A stored K x M, k contiguous) -> K-major ; AMODE 0 (M x K) -> MN-major
//   B: BMODE 0 (B stored K x N, k contiguous) -> K-major ; BMODE 1 (N x K) -> MN-major
template <int AK, int BKm>
__global__ void __launch_bounds__(NTHREADS, 1)
dgemm_tma_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, TmaP p) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    // the 128B swizzle pattern is a function of the shared address: align the ring to 1024 bytes
    unsigned char *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + STAGES * STAGE_BYTES);
    uint64_t *empty = full + STAGES;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int ti = blockIdx.x, tj = blockIdx.y;
    if (p.tri) {                                                   // t = i (i + 1) / 2 + j,  j <= i
        const int t = blockIdx.x;
        ti = (int)((sqrtf(8.f * (float)t + 1.f) - 1.f) * 0.5f);
        while (ti * (ti + 1) / 2 > t) --ti;
        while ((ti + 1) * (ti + 2) / 2 <= t) ++ti;
        tj = t - ti * (ti + 1) / 2;
    }
    const int m0 = ti * BM, n0 = tj * BN;
    if (p.lower_only && m0 + BM <= n0) return;
    const int kbeg = blockIdx.z * p.ksplit;
    const int kend = min(p.K, kbeg + p.ksplit);
    if (kbeg >= kend) return;
    const int KT = (kend - kbeg + BK - 1) / BK;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], NCONS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    if (warp == NCONS) {
        // ===== TMA producer =====
        if (lane == 0) {
            for (int kt = 0; kt < KT; ++kt) {
                const int slot = kt % STAGES;
                const uint32_t phase = (kt / STAGES) & 1;
                mbar_wait(&empty[slot], phase ^ 1);
                mbar_expect_tx(&full[slot], STAGE_BYTES);
                unsigned char *sa = smem + slot * STAGE_BYTES, *sb = sa + TILE_BYTES;
                const int k0 = kbeg + kt * BK;
                if (AK) {
                    tma_load_2d(sa, &mapA, k0, m0, &full[slot]);
                } else {
#pragma unroll
                    for (int b = 0; b < 8; ++b) tma_load_2d(sa + b * 2048, &mapA, m0 + 16 * b, k0, &full[slot]);
                }
                if (BKm) {
                    tma_load_2d(sb, &mapB, k0, n0, &full[slot]);
                } else {
#pragma unroll
                    for (int b = 0; b < 8; ++b) tma_load_2d(sb + b * 2048, &mapB, n0 + 16 * b, k0, &full[slot]);
                }
            }
        }
        return;
    }

    // ===== consumers =====
    const int gid = lane >> 2, tig = lane & 3;
    const int wm0 = (warp & 3) * 32, wn0 = (warp >> 2) * 32;
    // per-fragment byte offsets: addr(ks) = base + ks*KS_STEP + (pre ^ ks-dependent constant)
    uint32_t abase[4], apre[4], bbase[4], bpre[4];
#pragma unroll
    for (int f = 0; f < 4; ++f) {
        {
            const int r = wm0 + frag_row<AK>(f, gid);
            if (AK) {  // row*128 + ((2ks ^ pre) << 4) + (tig&1)*8,  pre = (r&7) ^ (tig>>1)
                abase[f] = r * 128 + (tig & 1) * 8;
                apre[f] = (uint32_t)(((r & 7) ^ (tig >> 1)) << 4);
            } else {   // (r>>4)*2048 + (ks*4+tig)*128 + ((((r&15)>>1) ^ ((ks&1)<<2 | tig)) << 4) + (r&1)*8
                abase[f] = (r >> 4) * 2048 + tig * 128 + (r & 1) * 8;
                apre[f] = (uint32_t)(((((r & 15) >> 1) ^ tig)) << 4);
            }
        }
        {
            const int r = wn0 + frag_row<BKm>(f, gid);
            if (BKm) {
                bbase[f] = r * 128 + (tig & 1) * 8;
                bpre[f] = (uint32_t)(((r & 7) ^ (tig >> 1)) << 4);
            } else {
                bbase[f] = (r >> 4) * 2048 + tig * 128 + (r & 1) * 8;
                bpre[f] = (uint32_t)(((((r & 15) >> 1) ^ tig)) << 4);
            }
        }
    }

    double acc[4][4][2];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[j][i][0] = acc[j][i][1] = 0.0;

    for (int kt = 0; kt < KT; ++kt) {
        const int slot = kt % STAGES;
        const uint32_t phase = (kt / STAGES) & 1;
        mbar_wait(&full[slot], phase);
        const unsigned char *sa = smem + slot * STAGE_BYTES, *sb = sa + TILE_BYTES;
#pragma unroll
        for (int ks = 0; ks < BK / 4; ++ks) {
            double fa[4], fb[4];
#pragma unroll
            for (int f = 0; f < 4; ++f) {
                const uint32_t oa = AK ? abase[f] + (apre[f] ^ (uint32_t)(ks << 5))
                                       : abase[f] + ks * 512 + (apre[f] ^ (uint32_t)((ks & 1) << 6));
                fa[f] = *reinterpret_cast<const double *>(sa + oa);
                const uint32_t ob = BKm ? bbase[f] + (bpre[f] ^ (uint32_t)(ks << 5))
                                        : bbase[f] + ks * 512 + (bpre[f] ^ (uint32_t)((ks & 1) << 6));
                fb[f] = *reinterpret_cast<const double *>(sb + ob);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int i = 0; i < 4; ++i) dmma884(acc[j][i][0], acc[j][i][1], fb[j], fa[i]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[slot]);
    }

    // epilogue: acc[j][i][e] is C[m = A-row(i, 2*tig+e)][n = B-row(j, gid)]
    // Stage the tile through the (now idle) pipeline ring so that C is read and written with
    // coalesced 16-byte accesses and all loads of a thread are in flight together: with K = nb = 128
    // (the QR trailing update) the epilogue is a third of the tile's time.
    constexpr int LDT = BM + 2;
    double *tile = reinterpret_cast<double *>(smem);   // [BN][LDT], element (m, n) at n*LDT + m
    asm volatile("bar.sync 1, %0;" ::"n"(NCONS * 32) : "memory");   // every consumer is done with the ring
    const double alpha = p.alpha, beta = p.partial ? 0.0 : p.beta;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int nl = wn0 + frag_row<BKm>(j, gid);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int e = 0; e < 2; ++e) tile[nl * LDT + wm0 + frag_row<AK>(i, 2 * tig + e)] = alpha * acc[j][i][e];
    }
    asm volatile("bar.sync 1, %0;" ::"n"(NCONS * 32) : "memory");
    double *__restrict__ C = p.C + (p.partial ? (int64_t)blockIdx.z * p.zstride : 0);
    if (p.vecC && !p.atomic) {
        constexpr int PER = 8;   // double2 chunks in flight per thread (2 rounds of 8 cover the tile)
        for (int round = 0; round < (BM / 2) * BN / (NCONS * 32 * PER); ++round) {
        double2 old[PER];
        bool ok[PER];
#pragma unroll
        for (int it = 0; it < PER; ++it) {
            const int idx = tid + (round * PER + it) * NCONS * 32;
            const int nl = idx / (BM / 2), ml = (idx % (BM / 2)) * 2;
            const int m = m0 + ml, n = n0 + nl;
            ok[it] = m < p.M && n < p.N && (!p.lower_only || m + 1 >= n);
            old[it] = make_double2(0.0, 0.0);
            if (ok[it] && beta != 0.0) {
                if (m + 1 < p.M) old[it] = *reinterpret_cast<const double2 *>(C + m + (int64_t)n * p.ldc);
                else old[it].x = C[m + (int64_t)n * p.ldc];
            }
        }
#pragma unroll
        for (int it = 0; it < PER; ++it) {
            if (!ok[it]) continue;
            const int idx = tid + (round * PER + it) * NCONS * 32;
            const int nl = idx / (BM / 2), ml = (idx % (BM / 2)) * 2;
            const int m = m0 + ml, n = n0 + nl;
            const double2 t = *reinterpret_cast<const double2 *>(tile + nl * LDT + ml);
            double2 o = make_double2(t.x + beta * old[it].x, t.y + beta * old[it].y);
            double *c = C + m + (int64_t)n * p.ldc;
            const bool w0 = !p.lower_only || m >= n, w1 = m + 1 < p.M;
            if (w0 && w1) *reinterpret_cast<double2 *>(c) = o;
            else {
                if (w0) c[0] = o.x;
                if (w1) c[1] = o.y;
            }
        }
        }  // round
    } else {
        for (int idx = tid; idx < BM * BN; idx += NCONS * 32) {
            const int nl = idx / BM, ml = idx % BM;
            const int m = m0 + ml, n = n0 + nl;
            if (m >= p.M || n >= p.N || (p.lower_only && m < n)) continue;
            double *c = C + m + (int64_t)n * p.ldc;
            const double v = tile[nl * LDT + ml];
            if (p.atomic) atomicAdd(c, v);
            else *c = v + (beta != 0.0 ? beta * *c : 0.0);
        }
    }
}

// ---- two CTAs per SM -------------------------------------------------------------------------------------------------
// With K = nb = 128 (QR trailing update) a 128 x 128 tile's main loop is 8 k-steps, and all sixteen consumer warps of the ONE
// resident CTA leave it together: the FP64 tensor pipe idles for the whole epilogue (C read-modify-write), 24 us per tile against
// 17 us of DMMA time.  A persistent tile loop does not change that (profiles/r2_dgemm_tn.md).  Here the tile is 128 x 64 with eight
// warps (same 32 x 32 warp tile, same fragment code), a 4-stage 24 KB ring and <= 128 registers, so TWO CTAs share an SM and one's
// epilogue overlaps the other's main loop.  No producer warp (a ninth warp would round the register allocation up to twelve):
// thread 0 refills the slot the CTA finished one iteration earlier.
constexpr int BN2 = 64, NC2 = 8, ST2 = 4;
constexpr int TILE_B2 = BN2 * BK * 8;             // 8 KB
constexpr int STAGE2 = TILE_BYTES + TILE_B2;      // 24 KB

template <int AK, int BKm>
__global__ void __launch_bounds__(NC2 * 32, 2)
dgemm_tma2_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, TmaP p) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + ST2 * STAGE2);
    uint64_t *empty = full + ST2;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int ti = blockIdx.x, tj = blockIdx.y;
    if (p.tri) {   // lower_only: row i of 128-row tiles has min(2 (i + 1), ptn) live 64-column tiles; blockIdx.x enumerates them row by row
        const int t = blockIdx.x;
        const int full_rows = p.ptn / 2;                      // rows i < full_rows hold 2 (i + 1) tiles: t < full_rows (full_rows + 1)
        if (t < full_rows * (full_rows + 1)) {
            ti = (int)((sqrtf(4.f * (float)t + 1.f) - 1.f) * 0.5f);
            while (ti * (ti + 1) > t) --ti;
            while ((ti + 1) * (ti + 2) <= t) ++ti;
            tj = t - ti * (ti + 1);
        } else {
            const int r = t - full_rows * (full_rows + 1);
            ti = full_rows + r / p.ptn;
            tj = r % p.ptn;
        }
    }
    const int m0 = ti * BM, n0 = tj * BN2;
    if (p.lower_only && m0 + BM <= n0) return;
    const int kbeg = blockIdx.z * p.ksplit;
    const int kend = min(p.K, kbeg + p.ksplit);
    if (kbeg >= kend) return;
    const int KT = (kend - kbeg + BK - 1) / BK;

    auto load_stage = [&](int kt) {
        const int slot = kt % ST2;
        mbar_wait(&empty[slot], ((kt / ST2) & 1) ^ 1);      // the previous use of the slot has been consumed by all eight warps
        mbar_expect_tx(&full[slot], STAGE2);
        unsigned char *sa = smem + slot * STAGE2, *sb = sa + TILE_BYTES;
        const int k0 = kbeg + kt * BK;
        if (AK) {
            tma_load_2d(sa, &mapA, k0, m0, &full[slot]);
        } else {
#pragma unroll
            for (int b = 0; b < 8; ++b) tma_load_2d(sa + b * 2048, &mapA, m0 + 16 * b, k0, &full[slot]);
        }
        if (BKm) {
            tma_load_2d(sb, &mapB, k0, n0, &full[slot]);
        } else {
#pragma unroll
            for (int b = 0; b < 4; ++b) tma_load_2d(sb + b * 2048, &mapB, n0 + 16 * b, k0, &full[slot]);
        }
    };

    if (tid == 0) {
        for (int s = 0; s < ST2; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], NC2);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        for (int kt = 0; kt < ST2 - 1 && kt < KT; ++kt) load_stage(kt);
    }
    __syncthreads();

    const int gid = lane >> 2, tig = lane & 3;
    const int wm0 = (warp & 3) * 32, wn0 = (warp >> 2) * 32;
    uint32_t abase[4], apre[4], bbase[4], bpre[4];
#pragma unroll
    for (int f = 0; f < 4; ++f) {
        {
            const int r = wm0 + frag_row<AK>(f, gid);
            if (AK) {
                abase[f] = r * 128 + (tig & 1) * 8;
                apre[f] = (uint32_t)(((r & 7) ^ (tig >> 1)) << 4);
            } else {
                abase[f] = (r >> 4) * 2048 + tig * 128 + (r & 1) * 8;
                apre[f] = (uint32_t)(((((r & 15) >> 1) ^ tig)) << 4);
            }
        }
        {
            const int r = wn0 + frag_row<BKm>(f, gid);
            if (BKm) {
                bbase[f] = r * 128 + (tig & 1) * 8;
                bpre[f] = (uint32_t)(((r & 7) ^ (tig >> 1)) << 4);
            } else {
                bbase[f] = (r >> 4) * 2048 + tig * 128 + (r & 1) * 8;
                bpre[f] = (uint32_t)(((((r & 15) >> 1) ^ tig)) << 4);
            }
        }
    }

    double acc[4][4][2];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[j][i][0] = acc[j][i][1] = 0.0;

    for (int kt = 0; kt < KT; ++kt) {
        const int slot = kt % ST2;
        mbar_wait(&full[slot], (kt / ST2) & 1);
        const unsigned char *sa = smem + slot * STAGE2, *sb = sa + TILE_BYTES;
#pragma unroll
        for (int ks = 0; ks < BK / 4; ++ks) {
            double fa[4], fb[4];
#pragma unroll
            for (int f = 0; f < 4; ++f) {
                const uint32_t oa = AK ? abase[f] + (apre[f] ^ (uint32_t)(ks << 5))
                                       : abase[f] + ks * 512 + (apre[f] ^ (uint32_t)((ks & 1) << 6));
                fa[f] = *reinterpret_cast<const double *>(sa + oa);
                const uint32_t ob = BKm ? bbase[f] + (bpre[f] ^ (uint32_t)(ks << 5))
                                        : bbase[f] + ks * 512 + (bpre[f] ^ (uint32_t)((ks & 1) << 6));
                fb[f] = *reinterpret_cast<const double *>(sb + ob);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int i = 0; i < 4; ++i) dmma884(acc[j][i][0], acc[j][i][1], fb[j], fa[i]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[slot]);
        if (tid == 0 && kt + ST2 - 1 < KT) load_stage(kt + ST2 - 1);   // slot of iteration kt - 1
        __syncwarp();
    }

    // epilogue through the (now idle) ring, as in dgemm_tma_kernel
    constexpr int LDT = BM + 2;
    double *tile = reinterpret_cast<double *>(smem);   // [BN2][LDT]
    __syncthreads();
    const double alpha = p.alpha, beta = p.partial ? 0.0 : p.beta;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int nl = wn0 + frag_row<BKm>(j, gid);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int e = 0; e < 2; ++e) tile[nl * LDT + wm0 + frag_row<AK>(i, 2 * tig + e)] = alpha * acc[j][i][e];
    }
    __syncthreads();
    double *__restrict__ C = p.C + (p.partial ? (int64_t)blockIdx.z * p.zstride : 0);
    constexpr int PER = 8;
#pragma unroll 1
    for (int round = 0; round < (BM / 2) * BN2 / (NC2 * 32 * PER); ++round) {
        double2 old[PER];
        bool ok[PER];
#pragma unroll
        for (int it = 0; it < PER; ++it) {
            const int idx = tid + (round * PER + it) * NC2 * 32;
            const int nl = idx / (BM / 2), ml = (idx % (BM / 2)) * 2;
            const int m = m0 + ml, n = n0 + nl;
            ok[it] = m < p.M && n < p.N && (!p.lower_only || m + 1 >= n);
            old[it] = make_double2(0.0, 0.0);
            if (ok[it] && beta != 0.0) {
                if (m + 1 < p.M) old[it] = *reinterpret_cast<const double2 *>(C + m + (int64_t)n * p.ldc);
                else old[it].x = C[m + (int64_t)n * p.ldc];
            }
        }
#pragma unroll
        for (int it = 0; it < PER; ++it) {
            if (!ok[it]) continue;
            const int idx = tid + (round * PER + it) * NC2 * 32;
            const int nl = idx / (BM / 2), ml = (idx % (BM / 2)) * 2;
            const int m = m0 + ml, n = n0 + nl;
            const double2 t = *reinterpret_cast<const double2 *>(tile + nl * LDT + ml);
            double2 o = make_double2(t.x + beta * old[it].x, t.y + beta * old[it].y);
            double *c = C + m + (int64_t)n * p.ldc;
            const bool w0 = !p.lower_only || m >= n, w1 = m + 1 < p.M;
            if (w0 && w1) *reinterpret_cast<double2 *>(c) = o;
            else {
                if (w0) c[0] = o.x;
                if (w1) c[1] = o.y;
            }
        }
    }
}


typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                             const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeFn get_encode() {
    // resolved once (thread-safe: C++11 static initialisation); several host threads drive several devices in multi.cu
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

// 2-D f64 tensor map over a column-major buffer: dim0 = contiguous extent d0, dim1 = d1 (stride ld).
bool make_map(CUtensorMap *map, const double *ptr, int64_t d0, int64_t d1, int64_t ld, int box0, int box1) {
    EncodeFn enc = get_encode();
    if (!enc) return false;
    cuuint64_t dims[2] = {(cuuint64_t)d0, (cuuint64_t)d1};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 8};
    cuuint32_t box[2] = {(cuuint32_t)box0, (cuuint32_t)box1};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void *)ptr, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

template <typename K>
__global__ void scale_kernel_d(K *C, int64_t M, int64_t N, int64_t ldc, K beta, int lower_only) {
    int64_t m = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (m >= M) return;
    for (int64_t n = blockIdx.y; n < N; n += gridDim.y) {
        if (lower_only && m < n) continue;
        K *c = C + m + n * ldc;
        *c = beta == K(0) ? K(0) : beta * *c;
    }
}

// C = beta C + sum_z W_z in the fixed order z = 0, 1, ...: the deterministic second stage of split-K.
__global__ void splitk_reduce_kernel(const double *__restrict__ W, int64_t ldw, int64_t zstride, int splits, double *__restrict__ C,
                                     int64_t M, int64_t N, int64_t ldc, double beta, int lower_only) {
    const int64_t m = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (m >= M) return;
    for (int64_t n = blockIdx.y; n < N; n += gridDim.y) {
        if (lower_only && m < n) continue;
        const double *w = W + m + n * ldw;
        double acc = 0.0;
        for (int z = 0; z < splits; ++z) acc += w[(int64_t)z * zstride];
        double *c = C + m + n * ldc;
        *c = beta == 0.0 ? acc : fma(beta, *c, acc);
    }
}

template <int AK, int BKm>
void launch(lfb_handle &h, const CUtensorMap &ma, const CUtensorMap &mb, const TmaP &p, dim3 grid) {
    constexpr size_t smem = STAGES * STAGE_BYTES + 2 * STAGES * sizeof(uint64_t) + 1024;
    static DeviceOnce cfg;   // function attributes are per device
    cfg.run(h.device, [&] {
        LFB_CUDA(cudaFuncSetAttribute(dgemm_tma_kernel<AK, BKm>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    });
    dgemm_tma_kernel<AK, BKm><<<grid, NTHREADS, smem, h.stream>>>(ma, mb, p);
    LFB_LAUNCH_CHECK(h);
}

template <int AK, int BKm>
void launch2(lfb_handle &h, const CUtensorMap &ma, const CUtensorMap &mb, const TmaP &p, dim3 grid) {
    constexpr size_t smem = ST2 * STAGE2 + 2 * ST2 * sizeof(uint64_t) + 1024;
    static DeviceOnce cfg;
    cfg.run(h.device, [&] {
        LFB_CUDA(cudaFuncSetAttribute(dgemm_tma2_kernel<AK, BKm>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    });
    dgemm_tma2_kernel<AK, BKm><<<grid, NC2 * 32, smem, h.stream>>>(ma, mb, p);
    LFB_LAUNCH_CHECK(h);
}

}  // namespace

bool dgemm_tma_try(lfb_handle &h, int ta, int tb, int64_t M, int64_t N, int64_t K, double alpha, const double *A,
                   int64_t lda, const double *B, int64_t ldb, double beta, double *C, int64_t ldc, int lower_only) {
    // TMA needs 16-byte aligned bases and strides; tiny problems stay on the cp.async kernel.
    if (((uintptr_t)A & 15) || ((uintptr_t)B & 15) || (lda & 1) || (ldb & 1)) return false;
    if (M < 64 || N < 32 || K < 64) return false;
    if (M >= (1LL << 31) || N >= (1LL << 31) || K >= (1LL << 31)) return false;
    alignas(64) CUtensorMap ma, mb;
    const int AK = ta ? 1 : 0;   // A stored K x M -> K-major tile
    const int BKm = tb ? 0 : 1;  // B stored K x N -> K-major tile
    bool ok = AK ? make_map(&ma, A, K, M, lda, BK, BM) : make_map(&ma, A, M, K, lda, 16, BK);
    ok = ok && (BKm ? make_map(&mb, B, K, N, ldb, BK, BN) : make_map(&mb, B, N, K, ldb, 16, BK));
    if (!ok) return false;

    TmaP p;
    p.M = (int)M; p.N = (int)N; p.K = (int)K;
    p.ldc = ldc; p.C = C; p.alpha = alpha; p.beta = beta; p.lower_only = lower_only;
    p.vecC = (((uintptr_t)C & 15) == 0) && ((ldc & 1) == 0);
    const int64_t tm = cdiv(M, BM), tn = cdiv(N, BN);
    int splits = 1;
    if (h.opt.gemm_splitk && tm * tn < h.sm_count) {
        // enough (tile, split) work items for `gemm_split_waves` waves: with one row of tiles (W = V^T C) two
        // waves of 128 CTAs on 148 SMs waste 14 % to wave quantisation
        int64_t want = cdiv(h.opt.gemm_split_waves * h.sm_count, tm * tn), maxs = std::max<int64_t>(1, K / (8 * BK));
        splits = (int)std::min(want, maxs);
    }
    p.ksplit = (int)round_up(cdiv(K, splits), BK);
    splits = (int)cdiv(K, p.ksplit);
    p.partial = 0;
    p.zstride = 0;
    // Split-K, second stage.  Default (gemm_deterministic): every split writes its partial tile to a workspace slice and a
    // reduce kernel sums the slices in a fixed order, so results are bit-reproducible run to run like the reference's
    // sequential loops; the slices cost 2 * splits * M * N * 8 B of extra traffic on outputs that are skinny by construction
    // (< 148 tiles).  gemm_deterministic = 0 restores the atomicAdd epilogue.
    std::unique_ptr<DevBuf<double>> work;
    const int64_t ldw = round_up(M, 2);
    if (splits > 1 && h.opt.gemm_deterministic) {
        work.reset(new DevBuf<double>(h, (size_t)ldw * N * splits));
        p.partial = 1;
        p.zstride = ldw * N;
        p.C = work->get();
        p.ldc = ldw;
        p.vecC = 1;
        p.atomic = 0;
    } else {
        p.atomic = splits > 1;
        if (p.atomic && beta != 1.0) {
            dim3 g((unsigned)cdiv(M, 256), (unsigned)(N < 65535 ? N : 65535));
            scale_kernel_d<double><<<g, 256, 0, h.stream>>>(C, M, N, ldc, beta, lower_only);
            LFB_LAUNCH_CHECK(h);
        }
    }
    // short main loops, full waves, no split-K, C addressable by 16-byte pairs: 128 x 64 tiles, two CTAs per SM
    const bool tma2 = h.opt.gemm_tma2 && p.vecC && !p.atomic &&
                      ((splits == 1 && K <= h.opt.gemm_tma2_maxk && tm * tn >= h.sm_count) || (p.partial && (h.opt.gemm_tma2 >= 2 || tm == 1)));
    if (tma2) {
        alignas(64) CUtensorMap mb2;
        if (BKm ? make_map(&mb2, B, K, N, ldb, BK, BN2) : true) {
            const CUtensorMap &mbu = BKm ? mb2 : mb;
            const int64_t tn2 = cdiv(N, BN2);
            dim3 grid2((unsigned)tm, (unsigned)tn2, (unsigned)splits);
            p.tri = 0;
            if (lower_only && tm > 1 && tm < 20000) {
                // live tiles row by row: rows i < tn2 / 2 hold 2 (i + 1), the rest all tn2 (kernel: the inverse map)
                const int64_t fr = std::min<int64_t>(tn2 / 2, tm);
                const int64_t live = fr * (fr + 1) + (tm - fr) * tn2;
                if (live < (1LL << 31)) {
                    p.tri = 1;
                    p.ptn = (int)tn2;
                    grid2 = dim3((unsigned)live, 1, (unsigned)splits);
                }
            }
            if (AK && BKm) launch2<1, 1>(h, ma, mbu, p, grid2);
            else if (AK && !BKm) launch2<1, 0>(h, ma, mbu, p, grid2);
            else if (!AK && BKm) launch2<0, 1>(h, ma, mbu, p, grid2);
            else launch2<0, 0>(h, ma, mbu, p, grid2);
            if (p.partial) {
                dim3 g((unsigned)cdiv(M, 128), (unsigned)(N < 65535 ? N : 65535));
                splitk_reduce_kernel<<<g, 128, 0, h.stream>>>(work->get(), ldw, p.zstride, splits, C, M, N, ldc, beta, lower_only);
                LFB_LAUNCH_CHECK(h);
            }
            return true;
        }
    }
    dim3 grid((unsigned)tm, (unsigned)tn, (unsigned)splits);
    p.tri = 0;
    if (lower_only && tm == tn && tm > 1 && tm < 40000) {          // the SYRK of the Cholesky: only the tm (tm + 1) / 2 live tiles are launched
        p.tri = 1;
        grid = dim3((unsigned)(tm * (tm + 1) / 2), 1, (unsigned)splits);
    }
    if (AK && BKm) launch<1, 1>(h, ma, mb, p, grid);
    else if (AK && !BKm) launch<1, 0>(h, ma, mb, p, grid);
    else if (!AK && BKm) launch<0, 1>(h, ma, mb, p, grid);
    else launch<0, 0>(h, ma, mb, p, grid);
    if (p.partial) {
        dim3 g((unsigned)cdiv(M, 128), (unsigned)(N < 65535 ? N : 65535));
        splitk_reduce_kernel<<<g, 128, 0, h.stream>>>(work->get(), ldw, p.zstride, splits, C, M, N, ldc, beta, lower_only);
        LFB_LAUNCH_CHECK(h);
    }
    return true;
}

}  // namespace lfb
