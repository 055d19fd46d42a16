// Cluster panel kernel, third generation (see householder.cu for the first one and the maths).
//
// A whole sub-panel (rows x w, w in {8,16,24,32}) lives in the distributed shared memory of ONE
// thread-block cluster (<= 16 CTAs, each owning a slab of rows).  Per column: one pass over the slab
// (scale the reflector, update the remaining columns, accumulate the full Gram row of the NEXT column
// against all w columns) and one cluster-wide all-reduce of that Gram row.
//
// ncu on the earlier generations (profiles/r1_panel_cluster.md): latency bound -- ~10 cycles per issued
// instruction with 2 warps per scheduler, 36 % of the stalls at block barriers around single-warp
// sections, 14 % on DSMEM *loads*.  This generation therefore
//   * PUSHES partial Gram rows (remote shared-memory stores are fire-and-forget) into an inbox in every
//     CTA instead of pulling them, so after the cluster barrier every warp reduces the 16 partials from
//     its OWN shared memory, redundantly, in a fixed order (deterministic, no block barrier);
//   * keeps the per-column scalars and the update factors warp-private (shuffles + a per-warp slab);
//   * needs one __syncthreads and one cluster barrier per column;
//   * takes the sub-panel width as a template parameter and handles two consecutive rows per lane with
//     16-byte shared-memory accesses.
// The kernel also emits the compact-WY factor T = (striu(V^T V) + I/2)^-1 of the sub-panel (the Gram
// rows contain v_i . v_j) and the staged V (zero above the diagonal, zero for `None` columns).
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace lfb {
namespace {

constexpr int WMAX = 32;
constexpr int NT = 256;
constexpr int NW = NT / 32;
constexpr int MAXC = 16;

template <typename T> struct Vec2;
template <> struct Vec2<double> { using type = double2; };
template <> struct Vec2<float> { using type = float2; };

template <typename T> __device__ __forceinline__ T t_abs(T x) { return x < T(0) ? -x : x; }
template <typename T> __device__ __forceinline__ T t_signum(T x) { return signbit(x) ? T(-1) : T(1); }
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

template <typename V>
__device__ __forceinline__ V make_zero() {
    V v;
    v.x = 0;
    v.y = 0;
    return v;
}

template <typename T>
struct PanelArgs2 {
    T *A; int64_t ld, m, c0;
    int rpc;
    T *beta;
    T *V; int64_t ldv, vrow0; int vcol0;
    T *Tout; int ldt;
    long long *dbg;   // optional: per-phase cycle counters (CTA 0, thread 0), see tools/panel_phases.py
};

// part[0..32) per lane -> lane l ends with the warp-wide sum of part[l]
template <typename T>
__device__ __forceinline__ T warp_transposed_reduce(T (&part)[WMAX], int lane) {
#pragma unroll
    for (int off = 16, n = 16; off >= 1; off >>= 1, n >>= 1) {
        const bool up = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (i < n) {
                const T send = up ? part[i] : part[i + n];
                const T keep = up ? part[i + n] : part[i];
                part[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
            }
        }
    }
    return part[0];
}

template <typename T, int WD>
__global__ void __launch_bounds__(NT, 1) hh_panel_cluster3(PanelArgs2<T> p) {
    using V2 = typename Vec2<T>::type;
    cg::cluster_group cluster = cg::this_cluster();
    const int nc = (int)cluster.num_blocks();
    const int b = (int)cluster.block_rank();
    extern __shared__ __align__(16) unsigned char panel_smem[];
    const int rpc = p.rpc;
    T *S = reinterpret_cast<T *>(panel_smem);   // [WD][rpc]
    T *inbox = S + (size_t)WD * rpc;            // [2][MAXC][32]  partial Gram rows pushed by every CTA
    T *inhead = inbox + 2 * MAXC * WMAX;        // [2][32]        head row pushed by its owner
    T *red = inhead + 2 * WMAX;                 // [NW][32]
    T *facw = red + NW * WMAX;                  // [NW][32]       warp-private update factors
    T *G = facw + NW * WMAX;                    // [32][32]       v_k . v_j (k < j), kept by warp 0
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t grow0 = p.c0 + (int64_t)b * rpc;
    const int nrows = (int)max((int64_t)0, min((int64_t)rpc, p.m - grow0));

#pragma unroll 4
    for (int k = 0; k < WD; ++k) {
        const T *src = p.A + grow0 + (p.c0 + k) * p.ld;
        for (int lr = tid; lr < rpc; lr += NT) S[(size_t)k * rpc + lr] = lr < nrows ? src[lr] : T(0);
    }
    for (int e = tid; e < 2 * MAXC * WMAX + 2 * WMAX; e += NT) inbox[e] = T(0);   // inbox + inhead
    for (int e = tid; e < WMAX * WMAX; e += NT) G[e] = T(0);
    __syncthreads();
    cluster.sync();   // every inbox is zeroed before anyone pushes into it

    T part[WMAX];
    // CTA-level reduction of part[] and push of the result (plus the head row of column jn, if this CTA
    // owns it) into slot `slot` of every CTA's inbox.  Ends with the cluster barrier.
    auto reduce_and_push = [&](int slot, int jn) {
        const T r = warp_transposed_reduce(part, lane);
        red[warp * WMAX + lane] = r;
        __syncthreads();   // also orders this column's slab updates before the head-row read below
        T sb[NW];
#pragma unroll
        for (int wv = 0; wv < NW; ++wv) sb[wv] = red[wv * WMAX + lane];
#pragma unroll
        for (int st = 1; st < NW; st <<= 1)
#pragma unroll
            for (int wv = 0; wv < NW; wv += 2 * st) sb[wv] += sb[wv + st];
        const T s = sb[0];
        const int owner = jn / rpc, lrh = jn % rpc;
        const T hval = (b == owner && lane < WD) ? S[(size_t)lane * rpc + lrh] : T(0);
        for (int dst = warp; dst < nc; dst += NW) {   // warp w serves CTAs w, w+8
            T *rin = cluster.map_shared_rank(inbox, dst);
            rin[(slot * MAXC + b) * WMAX + lane] = s;
            if (b == owner) cluster.map_shared_rank(inhead, dst)[slot * WMAX + lane] = hval;
        }
        cluster.sync();
    };

#pragma unroll
    for (int k = 0; k < WMAX; ++k) part[k] = T(0);
    for (int lr = tid; lr < nrows; lr += NT) {
        const T x = S[lr];
#pragma unroll
        for (int k = 0; k < WD; ++k) part[k] += S[(size_t)k * rpc + lr] * x;
    }
    reduce_and_push(0, 0);

    unsigned somemask = 0;   // bit k: column k produced a reflection (identical in every thread)
    long long tph[5] = {0, 0, 0, 0, 0};
    const bool dbg = p.dbg != nullptr && b == 0 && tid == 0;
    for (int j = 0; j < WD; ++j) {
        const int par = j & 1;
        const int64_t c = p.c0 + j;
        long long tq0 = dbg ? clock64() : 0;
        // every warp reduces the 16 pushed partials from its own shared memory, in a fixed order
        T qb[MAXC];
#pragma unroll
        for (int bb = 0; bb < MAXC; ++bb) qb[bb] = inbox[(par * MAXC + bb) * WMAX + lane];
#pragma unroll
        for (int st = 1; st < MAXC; st <<= 1)   // fixed pairwise tree: 4 dependent adds instead of 16
#pragma unroll
            for (int bb = 0; bb < MAXC; bb += 2 * st) qb[bb] += qb[bb + st];
        const T q = qb[0];
        const T hd = inhead[par * WMAX + lane];
        if (dbg) { long long t = clock64(); tph[0] += t - tq0; tq0 = t; }
        const T nsq = __shfl_sync(0xffffffffu, q, j), f = __shfl_sync(0xffffffffu, hd, j);
        const T rn = nsq > T(0) ? fast_rsqrt(nsq) : T(0);
        T nrm = nsq * rn;                                   // householder.rs:13
        nrm = fma(T(0.5) * rn, fma(-nrm, nrm, nsq), nrm);
        const T s = t_signum(f) * nrm;                      // :16
        const T newsq = (nsq + t_abs(f) * nrm) * T(2);      // :19-20
        const bool some = newsq != T(0);                    // :22
        const T rd = some ? fast_rsqrt(newsq) : T(0);
        const T dotv = some ? (q + s * hd) * rd : T(0);     // v_j . (column `lane`)
        facw[warp * WMAX + lane] = (lane > j && lane < WD) ? T(-2) * dotv : T(0);   // reflection.rs:29
        if (warp == 0) {
            if (lane < j) G[lane * WMAX + j] = ((somemask >> lane) & 1u) ? dotv : T(0);
            if (lane == 0 && b == 0) p.beta[c] = some ? -s : T(0);                  // householder.rs:24/26
        }
        if (some) somemask |= 1u << j;
        __syncwarp();
        if (dbg) { long long t = clock64(); tph[1] += t - tq0; tq0 = t; }
        const T *fac = facw + warp * WMAX;
        // ---- one pass over the slab, two consecutive rows per lane ----
#pragma unroll
        for (int k = 0; k < WMAX; ++k) part[k] = T(0);
        const bool has_next = j + 1 < WD;
        const int lr_min = (int)max((int64_t)0, c - grow0) & ~1;
        for (int lr = lr_min + 2 * tid; lr < nrows; lr += 2 * NT) {
            const int64_t gr = grow0 + lr;
            const bool a0 = gr >= c, a1 = gr + 1 >= c && lr + 1 < nrows;
            V2 v = *reinterpret_cast<const V2 *>(S + (size_t)j * rpc + lr);
            if (some) {
                if (a0) v.x = ((gr == c) ? v.x + s : v.x) * rd;              // householder.rs:17,23
                if (a1) v.y = ((gr + 1 == c) ? v.y + s : v.y) * rd;
                *reinterpret_cast<V2 *>(S + (size_t)j * rpc + lr) = v;
            }
            const T vx = a0 ? v.x : T(0), vy = a1 ? v.y : T(0);
            V2 n1 = make_zero<V2>();
            if (has_next) {
                n1 = *reinterpret_cast<const V2 *>(S + (size_t)(j + 1) * rpc + lr);
                if (some) {
                    n1.x += fac[j + 1] * vx;                                 // reflection.rs:30
                    n1.y += fac[j + 1] * vy;
                    *reinterpret_cast<V2 *>(S + (size_t)(j + 1) * rpc + lr) = n1;
                }
            }
            const T mx = (has_next && gr > c) ? n1.x : T(0);
            const T my = (has_next && gr + 1 > c && lr + 1 < nrows) ? n1.y : T(0);
            // Chunks of 8 columns: ALL loads of a chunk are issued before any store.  (Interleaving
            // "load column k, update, store column k" serialises on the store->load ordering of the
            // same shared array: measured 3800 cycles per pass instead of ~500.)
#pragma unroll
            for (int k0 = 0; k0 < WD; k0 += 8) {
                V2 a[8];
                T fk[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const int k = k0 + q;
                    a[q] = *reinterpret_cast<const V2 *>(S + (size_t)k * rpc + lr);
                    fk[q] = fac[k];
                }
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const int k = k0 + q;
                    if (k == j) { a[q].x = v.x; a[q].y = v.y; }
                    else if (k == j + 1) a[q] = n1;
                    else if (k > j && some) {
                        a[q].x += fk[q] * vx;
                        a[q].y += fk[q] * vy;
                    }
                    part[k] += a[q].x * mx + a[q].y * my;
                }
                if (some) {
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const int k = k0 + q;
                        if (k > j + 1) *reinterpret_cast<V2 *>(S + (size_t)k * rpc + lr) = a[q];
                    }
                }
            }
        }
        if (dbg) { long long t = clock64(); tph[2] += t - tq0; tq0 = t; }
        if (has_next) reduce_and_push(par ^ 1, j + 1);
        if (dbg) { long long t = clock64(); tph[3] += t - tq0; tq0 = t; }
    }
    if (dbg) for (int q = 0; q < 4; ++q) p.dbg[q] = tph[q];
    __syncthreads();

    // ---- write back: A (in place), staged V, and T (CTA 0) ----
#pragma unroll 4
    for (int k = 0; k < WD; ++k) {
        const bool sk = (somemask >> k) & 1u;
        T *dstA = p.A + grow0 + (p.c0 + k) * p.ld;
        T *dstV = p.V + (p.vrow0 + (grow0 - p.c0)) + (int64_t)(p.vcol0 + k) * p.ldv;
        for (int lr = tid; lr < nrows; lr += NT) {
            const T val = S[(size_t)k * rpc + lr];
            dstA[lr] = val;
            dstV[lr] = (sk && grow0 + lr >= p.c0 + k) ? val : T(0);
        }
    }
    if (b == 0) {
        for (int64_t e = tid; e < p.vrow0 * WD; e += NT) p.V[(e % p.vrow0) + (p.vcol0 + e / p.vrow0) * p.ldv] = T(0);
        // T = (striu(G) + I/2)^-1, one thread per column (G was written by warp 0 only)
        if (tid < WD) {
            const int jc = tid;
            T tcol[WD];
#pragma unroll
            for (int i = 0; i < WD; ++i) tcol[i] = T(0);
#pragma unroll
            for (int i = WD - 1; i >= 0; --i) {
                if (i == jc) tcol[i] = T(2);
                else if (i < jc) {
                    T sum = T(0);
#pragma unroll
                    for (int k = 0; k < WD; ++k)
                        if (k > i && k <= jc) sum += G[i * WMAX + k] * tcol[k];
                    tcol[i] = T(-2) * sum;
                }
            }
#pragma unroll
            for (int i = 0; i < WD; ++i) p.Tout[i + jc * p.ldt] = tcol[i];
        }
    }
}

constexpr size_t panel_extra_elems() { return 2 * MAXC * WMAX + 2 * WMAX + 2 * NW * WMAX + WMAX * WMAX; }

template <typename T, int WD>
bool launch_panel(lfb_handle &h, const PanelArgs2<T> &p, int nc, size_t smem) {
    static DeviceOnce cfg;   // function attributes are per device
    cfg.run(h.device, [&] {
        LFB_CUDA(cudaFuncSetAttribute(hh_panel_cluster3<T, WD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h.smem_optin));
        LFB_CUDA(cudaFuncSetAttribute(hh_panel_cluster3<T, WD>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    });
    cudaLaunchConfig_t cfgl = {};
    cfgl.gridDim = dim3((unsigned)nc);
    cfgl.blockDim = dim3(NT);
    cfgl.dynamicSmemBytes = smem;
    cfgl.stream = h.stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)nc;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfgl.attrs = attr;
    cfgl.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfgl, hh_panel_cluster3<T, WD>, p);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    h.launches++;
    return true;
}

}  // namespace

// Factors columns [c0, c0 + w) of A with one cluster launch; w (8/16/24/32, <= wmax) is chosen so
// that the sub-panel fits the cluster's shared memory.  Returns false if it cannot run (caller
// falls back to the per-column path).
template <typename T>
bool factor_subpanel_cluster2(lfb_handle &h, T *A, int64_t ld, int64_t m, int64_t c0, int wmax, int *w_out, T *beta, T *V,
                              int64_t ldv, int64_t vrow0, int vcol0, T *Tout, int ldt) {
    const int64_t rows = m - c0;
    if (rows <= 0 || wmax < 8) return false;
    const size_t extra = sizeof(T) * panel_extra_elems();
    if (h.smem_optin <= extra + 4096) return false;
    const int64_t cap = (int64_t)((h.smem_optin - extra - 1024) / sizeof(T));
    int nc = rows > 2048 ? 16 : rows > 1024 ? 8 : rows > 512 ? 4 : rows > 256 ? 2 : 1;
    if (nc > h.opt.panel_cluster_max) nc = (int)h.opt.panel_cluster_max;
    const int64_t rpc = round_up(cdiv(rows, nc), 32);
    int w = (wmax / 8) * 8;
    while (w > 8 && rpc * w > cap) w -= 8;
    if (rpc * w > cap) return false;
    PanelArgs2<T> p;
    p.A = A; p.ld = ld; p.m = m; p.c0 = c0; p.rpc = (int)rpc; p.beta = beta;
    p.V = V; p.ldv = ldv; p.vrow0 = vrow0; p.vcol0 = vcol0; p.Tout = Tout; p.ldt = ldt;
    p.dbg = (long long *)h.panel_dbg;
    const size_t smem = sizeof(T) * (size_t)(rpc * w) + extra;
    bool ok = false;
    switch (w) {
        case 8: ok = launch_panel<T, 8>(h, p, nc, smem); break;
        case 16: ok = launch_panel<T, 16>(h, p, nc, smem); break;
        case 24: ok = launch_panel<T, 24>(h, p, nc, smem); break;
        case 32: ok = launch_panel<T, 32>(h, p, nc, smem); break;
        default: return false;
    }
    if (ok) *w_out = w;
    return ok;
}

template bool factor_subpanel_cluster2<float>(lfb_handle &, float *, int64_t, int64_t, int64_t, int, int *, float *, float *,
                                              int64_t, int64_t, int, float *, int);
template bool factor_subpanel_cluster2<double>(lfb_handle &, double *, int64_t, int64_t, int64_t, int, int *, double *, double *,
                                               int64_t, int64_t, int, double *, int);

}  // namespace lfb
