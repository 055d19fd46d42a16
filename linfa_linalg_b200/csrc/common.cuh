// Shared host-side plumbing of liblinfa_b200: handle, workspace pool, error handling.
#pragma once
#include <cuda_runtime.h>

#include <cmath>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <functional>
#include <map>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "../../include/linfa_b200.h"

namespace lfb {

struct CudaError : std::runtime_error {
    int code;
    CudaError(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};

#define LFB_CUDA(expr)                                                                          \
    do {                                                                                        \
        cudaError_t e__ = (expr);                                                               \
        if (e__ != cudaSuccess) {                                                               \
            char buf__[512];                                                                    \
            snprintf(buf__, sizeof buf__, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), \
                     __FILE__, __LINE__);                                                       \
            throw ::lfb::CudaError(e__ == cudaErrorMemoryAllocation ? LFB_ERR_ALLOC : LFB_ERR_CUDA, buf__); \
        }                                                                                       \
    } while (0)

#define LFB_LAUNCH_CHECK(h)                 \
    do {                                    \
        (h).launches++;                     \
        LFB_CUDA(cudaGetLastError());       \
    } while (0)

struct Options {
    int64_t qr_nb = 128;     // outer panel width of blocked compact-WY QR
    int64_t qr_nb_f32 = 256; // same for f32 when the trailing updates run on the tcgen05 kernel (n >= 2048)
    int64_t qr_vt = 1;       // f64 QR: rank-nb update through a transposed copy of V (K-major tiles on both sides: 2 TMA box loads per stage instead of 9,
                             // which matters since the two-CTA kernel issues them from a compute warp): 231.1 vs 234.0 ms, same bits
    int64_t qr_panel_cholqr = 2; // blocked QR: panel = guarded Cholesky-QR + Householder reconstruction (tsqr_hr.cu) when its condition bound passes, else the cluster panel kernels; 1 = f64 only, 2 = f32 too (256-column panels as two fused 128-column ones: 125 -> 106 ms at 16384^2), 0 = off
    int64_t cholqr_fused = 1;    // 128-column Cholesky-QR stages as single-CTA kernels (panel_hr.cu): Cholesky + inverse + guard, and reconstruction + M + T
    int64_t qr_overlap_d2h = 1;  // host QR (pinned memory, cols >= 2048): finished block columns go back to the host during the factorisation
    int64_t qr_fold_t = 1;       // f64 QR: VT = V T once per panel on the look-ahead stream, so the trailing update needs (V T)^T C instead of T^T (V^T C)
    int64_t hr_split = 1;        // fused QR panel: the Z / T stage of the reconstruction kernel on the second side stream, beside the tall GEMM
    int64_t qr_trace = 0;    // debug: event time stamps of every stage of the look-ahead pipeline on stderr
    int64_t qr_sub = 32;     // inner BLAS-2 sub-panel width (<= 32)
    int64_t chol_base = 64;  // recursion base of Cholesky / TRSM (<= 64)
    int64_t chol_nb = 512;   // right-looking panel width of Cholesky (K of the trailing SYRK)
    int64_t chol_tn = 1;     // f64: trailing SYRK in TN form on a transposed copy of the panel (K-major TMA tiles on both sides)
    int64_t chol_nb_tail = 256;    // panel width once fewer than chol_tail_rows rows remain
    int64_t chol_tail_rows = 8192;
    int64_t chol_split_panel = 1; // look-ahead panel on TWO side streams: diagonal-block chain | rows below it (profiles/r2_chol_analysis.md)
    int64_t chol_trace = 0;    // debug: event time stamps of every stage of the look-ahead pipeline on stderr
    int64_t chol_potf2_rl = 1; // diagonal 64 x 64 blocks: right-looking register-blocked kernel (0 = first-generation left-looking)
    int64_t gemm_tma = 1;    // use the TMA-fed DGEMM when operands are 16-byte aligned
    int64_t gemm_tma2 = 1;   // f64 TMA GEMM with 128 x 64 tiles and two CTAs per SM (one's epilogue under the other's main loop): K <= gemm_tma2_maxk with at
                             // least one wave of tiles, and split-K products with one row of tiles (W = V^T C); 2 = every split-K product; 0 = off
    int64_t gemm_tma2_maxk = 1024;
    int64_t gemm_splitk = 1; // allow split-K for skinny-output GEMMs
    int64_t gemm_deterministic = 1; // split-K partial tiles summed in a fixed order by a reduce kernel (0 = atomicAdd epilogue)
    int64_t gemm_split_waves = 6; // target waves of (tile, split-K) work items for skinny outputs
    int64_t sgemm_tc = 1;    // f32 TN products >= 128 x 128 x 32: 1 = tcgen05 3xTF32 (TMEM accumulator), 2 = single-pass TF32 (yard-stick only), 0 = FFMA kernel
    int64_t gemm_v2 = 1;     // 16-warp cp.async DGEMM when operands are 16-byte aligned
    int64_t panel_cluster = 2; // cluster/DSMEM panel kernel: 2 = second generation, 1 = first, 0 = per-column launches
    int64_t lookahead = 1;     // factor the next panel on a side stream while the trailing update runs
    int64_t panel_cluster_max = 16; // largest cluster size tried (16 is non-portable but supported on B200)
    int64_t batched_quad = 2;       // f32 batched QR: 1 = four matrices per warp; 32 x 32: 2 = compile-time column steps (1.06 ms),
                                    // 3 = reflector made by the pivot lane alone, one shared-memory trip per step (measured slower: 1.21 ms)
    int64_t tsqr_chunk = 12288;     // rows per concurrently factored chunk (12288 x 32 f64 is the widest sub-panel a 16-CTA cluster holds)
    int64_t trd_fused = 1;          // tridiagonalisation: cluster head kernel + lower-triangle SYMV (0 = first generation)
    int64_t trd_symv_async = 1;     // SYMV tiles staged through shared memory with cp.async (0 = direct register loads)
    int64_t trd_profile = 0;        // debug: events around every tridiagonalisation launch, summary on stderr
    int64_t bd_blocked = 1;         // bidiagonalisation: blocked (deferred rank-1 updates); 0 = one reflector at a time
    int64_t rot_staged = 1;         // Givens wavefront: coefficients staged through shared memory with cp.async
    int64_t rot_serial = 0;         // debug: one chain per pass
    int64_t eigh_stable_2x2 = 1;    // eigh.rs:111 basis without cancellation (0 = the reference's formula verbatim)
    int64_t fast_hypot = 1;         // host recurrence: sqrt(x^2 + y^2) instead of hypot when far from underflow
    int64_t chol_waves = 3;         // host Cholesky (n >= 8192): factor in this many arrival waves of block columns while the rest is still crossing PCIe (1 = upload first)
    int64_t chol_overlap_d2h = 1;   // host Cholesky (dirty, n >= 2048): finished block columns go back to the host during the factorisation
    int64_t host_staging = 1;       // host Cholesky on PAGEABLE memory (n >= 2048): gather / scatter through the pinned buffer with host threads
    int64_t tsqr_streams = 8;       // chunks in flight (each panel kernel occupies one 16-SM cluster)
    int64_t qr_tsqr_auto = 0;       // 1: lfb_qr_* takes the TSQR + Householder-reconstruction route for tall-skinny inputs (rows >= 2 chunks, cols <= 512)
    int64_t tsqr_cholqr_cond = 16;  // tall-skinny leaf: Cholesky-QR (Gram GEMM + n x n Cholesky) when its cond_2 bound <= this; 0 = always Householder
    int64_t hr_lu_blocked = 1;      // Householder reconstruction: shared-memory panel LU of the top block (0 = one pivot at a time in global memory)
    int64_t tsqr_graph = 0;         // 1: replay the local TSQR stage of a (buffer, shape) seen before as one CUDA graph
                                    // (measured: 145 vs 147 ms -- the stage is GPU bound, not launch bound -- so off by default)
};

// A few persistent host threads for the staging copies between pageable caller memory and the handle's pinned buffer
// (one core's memcpy is ~8 GB/s, PCIe 5 x16 ~50 GB/s).  run(n, f) executes f(0..n-1) across the pool and returns when all
// are done; used by one call at a time, like the handle.
struct HostPool {
    std::vector<std::thread> threads;
    std::mutex mu;
    std::condition_variable cv_start, cv_done;
    std::function<void(int)> job;
    uint64_t generation = 0;
    int next = 0, total = 0, pending = 0;
    bool stop = false;
    explicit HostPool(int n) {
        for (int i = 0; i < n; ++i) threads.emplace_back([this] { loop(); });
    }
    ~HostPool() {
        {
            std::lock_guard<std::mutex> lk(mu);
            stop = true;
        }
        cv_start.notify_all();
        for (auto &t : threads) t.join();
    }
    void loop() {
        uint64_t seen = 0;
        for (;;) {
            std::unique_lock<std::mutex> lk(mu);
            cv_start.wait(lk, [&] { return stop || (generation != seen && next < total); });
            if (stop) return;
            while (next < total) {
                const int i = next++;
                lk.unlock();
                job(i);
                lk.lock();
                if (--pending == 0) cv_done.notify_all();
            }
            seen = generation;
        }
    }
    void run(int n, const std::function<void(int)> &f) {
        if (n <= 0) return;
        if (threads.empty() || n == 1) {
            for (int i = 0; i < n; ++i) f(i);
            return;
        }
        std::unique_lock<std::mutex> lk(mu);
        job = f;
        next = 0; total = n; pending = n;
        ++generation;
        cv_start.notify_all();
        cv_done.wait(lk, [&] { return pending == 0; });
    }
};

}  // namespace lfb

// The opaque C handle.
struct lfb_handle {
    int device = 0;
    cudaStream_t stream = nullptr;      // stream in use
    cudaStream_t own_stream = nullptr;  // created by lfb_create
    cudaStream_t aux_stream = nullptr;  // high-priority side stream for look-ahead panel factorisation
    cudaStream_t aux2_stream = nullptr; // second high-priority side stream (Cholesky: the below-diagonal half of a panel)
    cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    void *panel_dbg = nullptr;          // device buffer for the panel kernel's phase counters (debug)
    bool is_sub = false;                // a worker handle owned by another handle (TSQR chunk pool)
    std::vector<lfb_handle *> subs;     // created on demand by lfb_ensure_subs
    int sm_count = 148;
    size_t smem_optin = 0;
    std::string err;
    int64_t launches = 0;
    lfb::Options opt;

    // ---- caching device allocator (the engine reuses workspaces across calls) ----
    // A freed block remembers the stream that was current when it was released: work still queued there may be reading it, so
    // only a request made on the SAME stream (ordered behind that work) may take it.  The look-ahead drivers run two or three
    // streams of one handle concurrently and both sides allocate split-K workspaces; without the tag a side-stream GEMM could be
    // handed the workspace a main-stream GEMM is still reducing.  switch_stream() synchronises and clears the tags.
    struct Block { void *p; size_t bytes; bool used; cudaStream_t st; bool tagged; };
    void untag_blocks() { for (auto &b : blocks) b.tagged = false; }
    std::vector<Block> blocks;
    void *pinned = nullptr; size_t pinned_bytes = 0;
    lfb::HostPool *host_pool = nullptr;          // created on first use (pageable staging copies)
    lfb::HostPool &pool() {
        if (!host_pool) {
            unsigned hc = std::thread::hardware_concurrency();
            int n = (int)std::min<unsigned>(8, std::max<unsigned>(2, hc / 2));
            host_pool = new lfb::HostPool(n);
        }
        return *host_pool;
    }

    // ---- CUDA graphs of launch-bound multi-stream stages (TSQR local stage), keyed by buffer and shape ----
    struct GraphEntry {
        const void *a; const void *r; int64_t rows, cols, ld, ldr, chunk, streams; size_t elem;
        cudaGraphExec_t exec; int64_t launches;
    };
    std::vector<GraphEntry> graphs;
    bool in_capture = false;
    cudaEvent_t ev_graph[2] = {nullptr, nullptr};

    // ---- Cholesky host path: called (at enqueue time) right after panel [k0, k0 + nb) has been factored on `stream`,
    //      so that the finished block column can start its way back to the host while the trailing update runs ----
    std::function<void(int64_t k0, int64_t nb)> chol_panel_hook;
    std::function<void(int64_t k0, int64_t nb)> qr_panel_hook;     // host QR: block column [k0, k0 + nb) is final (reference signs applied) on h.stream
    cudaStream_t copy_stream = nullptr;
    cudaStream_t upload_stream = nullptr;
    void *hr_scratch = nullptr;             // panel_hr.cu: Y_1 / U and the pivot vectors handed from the LU stage to the M and T stages
    cudaEvent_t hr_ev[2] = {nullptr, nullptr};
    bool hr_pending = false;   // host Cholesky in arrival waves: H2D pieces (and their transposes) while earlier columns are factored
    void drop_graphs() {
        for (auto &g : graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
        graphs.clear();
    }

    // ---- optional GEMM profiler (bench.py roofline): CUDA events around every GEMM launch ----
    bool prof_on = false;
    std::vector<cudaEvent_t> prof_ev;
    size_t prof_used = 0;
    double prof_flops = 0.0;
    cudaEvent_t prof_event() {
        if (prof_used == prof_ev.size()) {
            cudaEvent_t e;
            LFB_CUDA(cudaEventCreate(&e));
            prof_ev.push_back(e);
        }
        return prof_ev[prof_used++];
    }

    void *dalloc(size_t bytes) {
        if (bytes == 0) bytes = 256;
        bytes = (bytes + 255) & ~size_t(255);
        int best = -1;
        for (int i = 0; i < (int)blocks.size(); ++i)
            if (!blocks[i].used && (!blocks[i].tagged || blocks[i].st == stream) && blocks[i].bytes >= bytes && blocks[i].bytes <= 2 * bytes + (1 << 20))
                if (best < 0 || blocks[i].bytes < blocks[best].bytes) best = i;
        if (best >= 0) { blocks[best].used = true; return blocks[best].p; }
        void *p = nullptr;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e != cudaSuccess) {
            // free cached blocks and retry once
            cudaGetLastError();
            trim();
            e = cudaMalloc(&p, bytes);
            if (e != cudaSuccess) {
                cudaGetLastError();
                throw lfb::CudaError(LFB_ERR_ALLOC, "device allocation of " + std::to_string(bytes) + " bytes failed");
            }
        }
        blocks.push_back({p, bytes, true, nullptr, false});
        return p;
    }
    void dfree(void *p) {
        for (auto &b : blocks) if (b.p == p) { b.used = false; b.st = stream; b.tagged = true; return; }
    }
    void trim() {
        drop_graphs();   // captured graphs hold pointers into the pool
        for (auto *sub : subs) sub->trim();
        std::vector<Block> keep;
        for (auto &b : blocks) { if (b.used) keep.push_back(b); else cudaFree(b.p); }
        blocks.swap(keep);
    }
    void *pinned_buf(size_t bytes) {
        if (bytes > pinned_bytes) {
            if (pinned) cudaFreeHost(pinned);
            pinned = nullptr; pinned_bytes = 0;
            LFB_CUDA(cudaMallocHost(&pinned, bytes));
            pinned_bytes = bytes;
        }
        return pinned;
    }
};

namespace lfb {

// "Done once per device" latch for cudaFuncSetAttribute calls (function attributes belong to a device's context, so a
// process that drives several GPUs has to set them on each).  The latch is taken under a mutex and only set after
// the body has run, so two handles used from two threads cannot launch before the attribute is in place.
struct DeviceOnce {
    std::mutex mu;
    bool done[64] = {};
    template <typename F>
    void run(int dev, F &&body) {
        std::lock_guard<std::mutex> g(mu);
        const bool tracked = dev >= 0 && dev < 64;
        if (tracked && done[dev]) return;
        body();                        // may throw (LFB_CUDA): the latch then stays open
        if (tracked) done[dev] = true;
    }
};

// RAII device buffer from the handle's pool.
template <typename T>
struct DevBuf {
    lfb_handle *h; T *p; size_t n;
    DevBuf(lfb_handle &hh, size_t count) : h(&hh), p((T *)hh.dalloc(count * sizeof(T))), n(count) {}
    ~DevBuf() { if (p) h->dfree(p); }
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    T *get() const { return p; }
    operator T *() const { return p; }
};

// Makes sure h.subs holds at least n worker handles (own streams, events, workspace pools).
void lfb_ensure_subs(lfb_handle &h, int n);

inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline int64_t round_up(int64_t a, int64_t b) { return cdiv(a, b) * b; }

// ---- internal engine API (all column-major device pointers, async on h.stream) ----
template <typename T>
void gemm(lfb_handle &h, int ta, int tb, int64_t M, int64_t N, int64_t K, T alpha, const T *A, int64_t lda,
          const T *B, int64_t ldb, T beta, T *C, int64_t ldc, int lower_only = 0);

template <typename T> void transpose(lfb_handle &h, const T *in, int64_t rows, int64_t cols, int64_t ldin, T *out, int64_t ldout);
template <typename T> void transpose_inplace_square(lfb_handle &h, T *a, int64_t n, int64_t ld);
template <typename T> void fill(lfb_handle &h, T *p, int64_t rows, int64_t cols, int64_t ld, T offdiag, T diag);
template <typename T> void copy2d(lfb_handle &h, const T *in, int64_t ldin, T *out, int64_t ldout, int64_t rows, int64_t cols);

template <typename T> void qr_factor(lfb_handle &h, T *A, int64_t m, int64_t n, int64_t ld, T *diag);
template <typename T> void assemble_q(lfb_handle &h, const T *M, int64_t rows, int64_t cols, int64_t ld, int64_t shift,
                                      const T *signs, T *Q, int64_t ldq);
template <typename T> void qt_mul(lfb_handle &h, const T *QR, int64_t rows, int64_t cols, int64_t ld, const T *diag,
                                  T *B, int64_t bcols, int64_t ldb);
template <typename T> void cholesky_lower(lfb_handle &h, T *A, int64_t n, int64_t ld, int clean, int64_t *d_info);
template <typename T> void cholesky_lower_wave(lfb_handle &h, T *A, int64_t n, int64_t ld, int64_t c0, int64_t c1, int64_t *d_info, int first);
// op(A) X = B, A n x n column-major, lower != 0 -> A is lower triangular; trans != 0 -> solve A^T X = B.
template <typename T> void trsm_left(lfb_handle &h, int lower, int trans, int64_t n, int64_t nrhs, const T *A, int64_t lda,
                                     const T *ext_diag, T *B, int64_t ldb);
template <typename T> void sym_tridiagonal(lfb_handle &h, T *A, int64_t n, int64_t ld, T *off);
template <typename T> void bidiagonal(lfb_handle &h, T *A, int64_t rows, int64_t cols, int64_t ld, T *d, T *e);
// eigh.rs:10-129: dA consumed; vals is a HOST array (reference order); dQ device n x n or nullptr.
template <typename T> void symmetric_eig(lfb_handle &h, T *dA, int64_t n, int64_t ld, T *vals, T *dQ, int64_t ldq);
// svd.rs:17-221: dA consumed; sv is a HOST array (reference order); dU rows x dim or nullptr; dV = Vt^T, cols x dim, or nullptr.
template <typename T> void svd_dev(lfb_handle &h, T *dA, int64_t rows, int64_t cols, int64_t ld, T *sv, T *dU, int64_t ldu, T *dV, int64_t ldv);
template <typename T> void qr_batched(lfb_handle &h, T *A, int64_t batch, int64_t m, int64_t n, T *diag);
// n <= 32, packed row-major [batch][n][n]; fail[b] = first row with a non-positive pivot or -1.
template <typename T> void cholesky_batched(lfb_handle &h, T *A, int64_t batch, int64_t n, int clean, int *fail);
template <typename T> void tsqr_local_r(lfb_handle &h, T *A, int64_t rows, int64_t cols, int64_t ld, T *R, int64_t ldr);
// TSQR keeping the orthogonal factor: A <- explicit thin Q, R <- triangular factor (diag >= 0); Wk: rows x cols scratch.
template <typename T> void tsqr_explicit_q(lfb_handle &h, T *A, int64_t rows, int64_t cols, int64_t ld, T *Wk, int64_t ldw, T *R, int64_t ldr);
template <typename T> void tsqr_apply_q(lfb_handle &h, T *Q, int64_t rows, int64_t cols, int64_t ld, const T *Qs, int64_t ldqs);
// X U = B in place on B (rows x n), U upper triangular n x n.
template <typename T> void trsm_right_upper(lfb_handle &h, int64_t rows, int64_t n, const T *U, int64_t ldu, T *B, int64_t ldb);
// X op(Tri) = B in place on B (rows x n): trans_lower = 0 -> Tri is upper, 1 -> Tri is lower and the system is X Tri^T = B.
template <typename T> void trsm_right(lfb_handle &h, int64_t rows, int64_t n, const T *Tri, int64_t ldt, int trans_lower, T *B,
                                      int64_t ldb, const int64_t *info);
// LOBPCG's dense blocks kept on the device (lobpcg/algorithm.rs:63-97), lobpcg_blocks.cu.
template <typename T> void orthonormalize(lfb_handle &h, T *V, int64_t rows, int64_t cols, int64_t ld, T *Lm, int64_t ldl, int64_t *d_info);
template <typename T> void apply_constraints(lfb_handle &h, T *V, int64_t n, int64_t k, int64_t ldv, const T *Lyy, int64_t m, int64_t ldl,
                                             const T *Y, int64_t ldy);
// lobpcg/algorithm.rs:16-44 on device-resident k x k operands (lobpcg_blocks.cu); false if an eigenvalue is NaN.
template <typename T> bool sorted_eig_dev(lfb_handle &h, T *dA, int64_t lda, T *dB, int64_t ldb, int64_t k, int64_t size, int order, T *vals_host,
                                          T *dVecs, int64_t ldv);
// Householder reconstruction (tsqr_hr.cu): top n x n block of an explicit Q -> reference compact form; U' for the rows below.
template <typename T> void hh_reconstruct_top(lfb_handle &h, T *Qtop, int64_t n, int64_t ld, const T *R, int64_t ldr, T *U, int64_t ldu, T *diag, int internal = 0);
// Tall-skinny thin QR in the reference's compact form (identical contract to qr_factor) via TSQR + reconstruction.
template <typename T> void qr_tsqr(lfb_handle &h, T *A, int64_t rows, int64_t cols, int64_t ld, T *diag);
// Cholesky-QR leaf (cholqr.cu): R (n x n upper, diag >= 0) and optionally R^-1 of a tall block WITHOUT touching A; returns
// false (nothing written) when the Gram matrix is not safely positive definite -- the caller then takes the Householder route.
template <typename T> void cholqr128(lfb_handle &h, const T *G, int64_t ldg, T *R, int64_t ldr, T *Rinv, int64_t ldri, double *guard);
template <typename T> void hr_panel128(lfb_handle &h, T *Atop, int64_t ld, const T *R, int64_t ldr, const T *Rinv, int64_t ldri, T *beta, T *M, int64_t ldm,
                                       T *Tm, int64_t ldt, T *Vtop, int64_t ldv);
void hr_panel128_join(lfb_handle &h);
template <typename T> bool cholqr_factor(lfb_handle &h, const T *A, int64_t rows, int64_t n, int64_t ld, T *R, int64_t ldr, T *Rinv, int64_t ldri);
template <typename T> void triangular_zero(lfb_handle &h, T *A, int64_t n, int64_t ld, int keep_lower);
double microbench_fp64(lfb_handle &h, int kind);
double microbench_trd(lfb_handle &h, int kind, int64_t n, int reps);
double microbench_potf2(lfb_handle &h, int kind, int reps);
double microbench_bd_gemv(lfb_handle &h, int kind, int64_t m, int64_t n, int reps);

}  // namespace lfb
