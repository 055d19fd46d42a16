// Multi-GPU entry points behind the C ABI (include/linfa_b200.h, "one box, several devices"): ONE process, one
// lfb_handle + one host worker thread per device, NCCL (ncclCommInitAll, resolved with dlopen so that the library has
// no link-time dependency and shares the copy a host process may already have loaded) for the only exchange the hot
// path has -- the n x n R factors of a row-sharded tall-skinny QR -- and no collective at all for the batch-sharded
// small-matrix factorisations.  This is what a Rust caller of `QRInto::qr_into` (qr.rs:29-45) reaches through the
// shim: it cannot spawn one process per GPU the way bench.py's torchrun launch does (linfa_linalg_b200/dist.py keeps
// that launch mode and mirrors this file's sharding).
//
// Row-sharded tall-skinny qr_into (device i owns rows [b_i, e_i), b/e as dist.py: shard_range):
//   P1  every device: upload its rows; explicit-Q TSQR of the block -> Q_i (in place), R_i          [parallel]
//   C1  ncclAllGather of the R_i (n*n values per rank)                                              [NVLink]
//   P2  every device: stack the R factors, explicit-Q QR of the (G n) x n stack -> Qs, R (replicated);
//       Q_i <- Q_i Qs[i]; device 0: Householder reconstruction of the top n x n block -> U', diag   [parallel]
//   C2  ncclBroadcast of U' and diag (one buffer) from device 0
//   P3  every device: rows <- rows U'^-1 (one right-hand TRSM); download                            [parallel]
// R-only TSQR stops after C1 + the QR of the stack.  Results are identical in contract to lfb_qr_* / lfb_qr_tsqr_*.
#include <dlfcn.h>
#include <nccl.h>

#include <condition_variable>
#include <memory>
#include <thread>

#include "common.cuh"
#include "host_io.cuh"

using namespace lfb;

namespace {

struct NcclApi {
    void *lib = nullptr;
    decltype(&ncclCommInitAll) CommInitAll = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclAllGather) AllGather = nullptr;
    decltype(&ncclBroadcast) Broadcast = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    decltype(&ncclCommCount) CommCount = nullptr;
    decltype(&ncclGetVersion) GetVersion = nullptr;
    std::string why;
    bool load() {
        const char *override_path = getenv("LFB_NCCL_LIB");
        if (override_path && *override_path) lib = dlopen(override_path, RTLD_NOW | RTLD_LOCAL);
        // a copy already mapped by the host process (e.g. the one PyTorch bundles) wins: two NCCLs in one process would
        // export the same symbols from two versions
        if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
        if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
        if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
        if (!lib) { why = std::string("libnccl.so.2 not found: ") + (dlerror() ? dlerror() : "?"); return false; }
#define LFB_SYM(field, name)                                                        \
        field = reinterpret_cast<decltype(field)>(dlsym(lib, name));                \
        if (!field) { why = std::string("NCCL symbol missing: ") + name; return false; }
        LFB_SYM(CommInitAll, "ncclCommInitAll")
        LFB_SYM(CommDestroy, "ncclCommDestroy")
        LFB_SYM(AllGather, "ncclAllGather")
        LFB_SYM(Broadcast, "ncclBroadcast")
        LFB_SYM(GroupStart, "ncclGroupStart")
        LFB_SYM(GroupEnd, "ncclGroupEnd")
        LFB_SYM(GetErrorString, "ncclGetErrorString")
        LFB_SYM(CommCount, "ncclCommCount")
        LFB_SYM(GetVersion, "ncclGetVersion")
#undef LFB_SYM
        return true;
    }
};

template <typename T> ncclDataType_t nccl_type();
template <> ncclDataType_t nccl_type<float>() { return ncclFloat32; }
template <> ncclDataType_t nccl_type<double>() { return ncclFloat64; }

}  // namespace

struct lfb_multi {
    std::vector<lfb_handle *> hs;
    std::vector<int> devices;
    NcclApi nccl;
    std::vector<ncclComm_t> comms;
    int nccl_ranks = 0;
    int nccl_version = 0;
    std::string err;
    std::vector<cudaEvent_t> tev[2];   // timing events, one pair per device

    // ---- one persistent host thread per device: the local stages are thousands of launches on several streams, so the
    //      enqueue itself has to run in parallel across devices ----
    std::mutex mu;
    std::condition_variable cv_start, cv_done;
    uint64_t generation = 0;
    int pending = 0;
    bool stop = false;
    std::function<void(int)> job;
    std::vector<std::thread> threads;
    std::vector<int> wcode;
    std::vector<std::string> werr;

    void worker(int i) {
        uint64_t seen = 0;
        for (;;) {
            std::function<void(int)> f;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv_start.wait(lk, [&] { return stop || generation != seen; });
                if (stop) return;
                seen = generation;
                f = job;
            }
            int code = LFB_OK;
            std::string msg;
            try {
                LFB_CUDA(cudaSetDevice(devices[i]));
                f(i);
            } catch (const lfb::CudaError &e) {
                code = e.code; msg = e.what(); cudaGetLastError();
            } catch (const std::exception &e) {
                code = LFB_ERR_CUDA; msg = e.what();
            }
            {
                std::lock_guard<std::mutex> lk(mu);
                wcode[i] = code; werr[i] = msg;
                if (--pending == 0) cv_done.notify_all();
            }
        }
    }
    // Runs f(i) for every device i on its own thread; returns the first failure (device order).
    int run_all(const std::function<void(int)> &f) {
        {
            std::lock_guard<std::mutex> lk(mu);
            job = f;
            pending = (int)hs.size();
            ++generation;
        }
        cv_start.notify_all();
        std::unique_lock<std::mutex> lk(mu);
        cv_done.wait(lk, [&] { return pending == 0; });
        for (size_t i = 0; i < hs.size(); ++i)
            if (wcode[i] != LFB_OK) {
                err = "device " + std::to_string(devices[i]) + ": " + werr[i];
                return wcode[i];
            }
        return LFB_OK;
    }
    int nccl_fail(ncclResult_t r, const char *what) {
        err = std::string(what) + ": " + (nccl.GetErrorString ? nccl.GetErrorString(r) : "NCCL error");
        return LFB_ERR_NCCL;
    }
};

namespace {

inline void shard_range(int64_t total, int world, int rank, int64_t *b, int64_t *e) {   // == dist.py: shard_range
    const int64_t base = total / world, rem = total % world;
    *b = rank * base + std::min<int64_t>(rank, rem);
    *e = *b + base + (rank < rem ? 1 : 0);
}

int mfail(lfb_multi *m, int code, const char *msg) {
    if (m) m->err = msg;
    return code;
}

// All-gather of one n x n (contiguous) factor per device into every device's `all` buffer, on the handles' streams.
template <typename T>
int allgather_r(lfb_multi *m, const std::vector<T *> &send, const std::vector<T *> &all, size_t count) {
    const int G = (int)m->hs.size();
    ncclResult_t r = m->nccl.GroupStart();
    if (r != ncclSuccess) return m->nccl_fail(r, "ncclGroupStart");
    for (int i = 0; i < G; ++i) {
        cudaSetDevice(m->devices[i]);
        r = m->nccl.AllGather(send[i], all[i], count, nccl_type<T>(), m->comms[i], m->hs[i]->stream);
        if (r != ncclSuccess) { m->nccl.GroupEnd(); return m->nccl_fail(r, "ncclAllGather"); }
    }
    r = m->nccl.GroupEnd();
    if (r != ncclSuccess) return m->nccl_fail(r, "ncclGroupEnd");
    for (auto *h : m->hs) h->launches++;   // one collective kernel per device
    return LFB_OK;
}

template <typename T>
int broadcast0(lfb_multi *m, const std::vector<T *> &buf, size_t count) {
    const int G = (int)m->hs.size();
    ncclResult_t r = m->nccl.GroupStart();
    if (r != ncclSuccess) return m->nccl_fail(r, "ncclGroupStart");
    for (int i = 0; i < G; ++i) {
        cudaSetDevice(m->devices[i]);
        r = m->nccl.Broadcast(buf[i], buf[i], count, nccl_type<T>(), 0, m->comms[i], m->hs[i]->stream);
        if (r != ncclSuccess) { m->nccl.GroupEnd(); return m->nccl_fail(r, "ncclBroadcast"); }
    }
    r = m->nccl.GroupEnd();
    if (r != ncclSuccess) return m->nccl_fail(r, "ncclGroupEnd");
    for (auto *h : m->hs) h->launches++;
    return LFB_OK;
}

// Per-device state of one tall-skinny call (buffers live from P1 to P3).
template <typename T>
struct TsqrDev {
    T *A = nullptr; int64_t rows = 0, ld = 0;          // this device's row block (column-major)
    std::unique_ptr<DevBuf<T>> own, Wk, R, Rall, stack, Ws, UD;
};

// Core of the row-sharded factorisation on device-resident blocks.  want_q: deliver the reference's compact factor
// (qr_into contract) in the blocks + diag; otherwise R only.  On return (asynchronous on every handle's stream):
//   st[i].R  : n x n column-major (ld n) final R (diag >= 0), on every device
//   st[i].UD : (want_q) U' (ldu x n) followed by diag (n values), on every device
//
// want_q takes the FOLDED route when the Cholesky-QR leaf (cholqr.cu) accepts every device's block: with
// Q_i = A_i R_i^-1 Qs_i (Qs = the Q of the stacked R factors) the reflector rows are Y_i = Q_i U'^-1 = A_i W_i,
// W_i = R_i^-1 Qs_i U'^-1 (n x n), so each device touches its rows with ONE Gram GEMM and ONE GEMM by W_i; only device 0
// forms the n x n top block Q_top = A_top M_0 for the reconstruction.  If any block is declined (ill-conditioned / rank
// deficient) every device takes the Householder route: explicit-Q TSQR, Q_i <- Q_i Qs_i, right-hand TRSM.
template <typename T>
int tsqr_core(lfb_multi *m, std::vector<TsqrDev<T>> &st, int64_t n, bool want_q) {
    const int G = (int)m->hs.size();
    const int64_t ldu = round_up(n, 2);
    const int64_t srows = (int64_t)G * n, lds = round_up(srows, 2);
    std::vector<int> leaf_ok(G, 0);
    std::vector<std::unique_ptr<DevBuf<T>>> Rinv(G), Qt(G);
    int rc = m->run_all([&](int i) {
        lfb_handle &h = *m->hs[i];
        TsqrDev<T> &d = st[i];
        d.R.reset(new DevBuf<T>(h, (size_t)n * n));
        if (G > 1) d.Rall.reset(new DevBuf<T>(h, (size_t)G * n * n));
        if (want_q) {
            Rinv[i].reset(new DevBuf<T>(h, (size_t)ldu * n));
            leaf_ok[i] = (d.rows >= 4 * n && n <= 512 && cholqr_factor<T>(h, d.A, d.rows, n, d.ld, d.R->get(), n, Rinv[i]->get(), ldu)) ? 1 : 0;
        } else {
            tsqr_local_r<T>(h, d.A, d.rows, n, d.ld, d.R->get(), n);
        }
    });
    if (rc != LFB_OK) return rc;
    bool folded = want_q;
    for (int i = 0; i < G; ++i) folded = folded && leaf_ok[i];
    if (want_q && !folded) {
        // Householder route for everybody; the leaf attempt is switched off meanwhile so that it is not repeated per block
        rc = m->run_all([&](int i) {
            lfb_handle &h = *m->hs[i];
            TsqrDev<T> &d = st[i];
            const int64_t keep = h.opt.tsqr_cholqr_cond;
            h.opt.tsqr_cholqr_cond = 0;
            try {
                const int64_t ldw = round_up(d.rows, 2);
                d.Wk.reset(new DevBuf<T>(h, (size_t)ldw * n));
                tsqr_explicit_q<T>(h, d.A, d.rows, n, d.ld, d.Wk->get(), ldw, d.R->get(), n);
                d.Wk.reset();      // stream-ordered reuse: every later user of this handle's pool runs on the same stream
            } catch (...) {
                h.opt.tsqr_cholqr_cond = keep;
                throw;
            }
            h.opt.tsqr_cholqr_cond = keep;
        });
        if (rc != LFB_OK) return rc;
    }
    if (G > 1) {
        std::vector<T *> send(G), all(G);
        for (int i = 0; i < G; ++i) { send[i] = st[i].R->get(); all[i] = st[i].Rall->get(); }
        rc = allgather_r<T>(m, send, all, (size_t)n * n);
        if (rc != LFB_OK) return rc;
    }
    rc = m->run_all([&](int i) {
        lfb_handle &h = *m->hs[i];
        TsqrDev<T> &d = st[i];
        if (G > 1) {
            d.stack.reset(new DevBuf<T>(h, (size_t)lds * n));
            for (int g = 0; g < G; ++g) copy2d<T>(h, d.Rall->get() + (size_t)g * n * n, n, d.stack->get() + (size_t)g * n, lds, n, n);
            if (want_q) {
                d.Ws.reset(new DevBuf<T>(h, (size_t)lds * n));
                tsqr_explicit_q<T>(h, d.stack->get(), srows, n, lds, d.Ws->get(), lds, d.R->get(), n);   // stack <- Qs
                if (folded) {   // M_i = R_i^-1 Qs_i, kept in Ws (its head is free again)
                    gemm<T>(h, 0, 0, n, n, n, T(1), Rinv[i]->get(), ldu, d.stack->get() + (size_t)i * n, lds, T(0), d.Ws->get(), ldu);
                    LFB_CUDA(cudaMemcpyAsync(Rinv[i]->get(), d.Ws->get(), sizeof(T) * ldu * n, cudaMemcpyDeviceToDevice, h.stream));
                } else {
                    tsqr_apply_q<T>(h, d.A, d.rows, n, d.ld, d.stack->get() + (size_t)i * n, lds);
                }
            } else {
                tsqr_local_r<T>(h, d.stack->get(), srows, n, lds, d.R->get(), n);
            }
        }
        if (want_q) {
            d.UD.reset(new DevBuf<T>(h, (size_t)ldu * n + n));
            if (i == 0) {
                T *top = d.A;
                int64_t ldtop = d.ld;
                if (folded) {   // Q_top = A_top M_0, out of place: A's rows stay intact until the final GEMM
                    Qt[0].reset(new DevBuf<T>(h, (size_t)ldu * n));
                    gemm<T>(h, 0, 0, n, n, n, T(1), d.A, d.ld, Rinv[0]->get(), ldu, T(0), Qt[0]->get(), ldu);
                    top = Qt[0]->get();
                    ldtop = ldu;
                }
                hh_reconstruct_top<T>(h, top, n, ldtop, d.R->get(), n, d.UD->get(), ldu, d.UD->get() + (size_t)ldu * n);
            }
        }
    });
    if (rc != LFB_OK || !want_q) return rc;
    if (G > 1) {
        std::vector<T *> ud(G);
        for (int i = 0; i < G; ++i) ud[i] = st[i].UD->get();
        rc = broadcast0<T>(m, ud, (size_t)ldu * n + n);
        if (rc != LFB_OK) return rc;
    }
    return m->run_all([&](int i) {
        lfb_handle &h = *m->hs[i];
        TsqrDev<T> &d = st[i];
        const int64_t off = i == 0 ? n : 0;
        if (folded) {
            trsm_right_upper<T>(h, n, n, d.UD->get(), ldu, Rinv[i]->get(), ldu);                      // W_i = M_i U'^-1
            tsqr_apply_q<T>(h, d.A + off, d.rows - off, n, d.ld, Rinv[i]->get(), ldu);                // rows <- rows W_i
            if (i == 0) copy2d<T>(h, Qt[0]->get(), ldu, d.A, d.ld, n, n);
        } else {
            trsm_right_upper<T>(h, d.rows - off, n, d.UD->get(), ldu, d.A + off, d.ld);
        }
    });
}

template <typename T>
int tsqr_host(lfb_multi *m, T *a, int64_t rows, int64_t cols, int64_t rs, int64_t cs, T *diag, T *r_out, int64_t r_rs, int64_t r_cs) {
    if (!m) return LFB_INVALID_ARGUMENT;
    m->err.clear();
    if (rows < 0 || cols < 0) return mfail(m, LFB_INVALID_ARGUMENT, "negative dimension");
    if (rows < cols) return mfail(m, LFB_NOT_THIN, "Expected matrix rows >= cols");                 // qr.rs:34-36
    if (cols == 0) return LFB_OK;                                                                  // qr.rs:383-388
    const int G = (int)m->hs.size();
    const bool want_q = diag != nullptr;
    // every shard has to be thin itself (rows_i >= cols); a matrix too short for that is not worth sharding
    if (rows / G < cols) {
        lfb_handle *h0 = m->hs[0];
        int st;
        if (want_q) st = sizeof(T) == 8 ? lfb_qr_tsqr_f64(h0, (double *)a, rows, cols, rs, cs, (double *)diag)
                                        : lfb_qr_tsqr_f32(h0, (float *)a, rows, cols, rs, cs, (float *)diag);
        else {
            std::vector<T> dtmp((size_t)cols);
            std::vector<T> copy((size_t)rows * cols);
            for (int64_t i = 0; i < rows; ++i) for (int64_t j = 0; j < cols; ++j) copy[i * cols + j] = a[i * rs + j * cs];
            st = sizeof(T) == 8 ? lfb_qr_f64(h0, (double *)copy.data(), rows, cols, cols, 1, (double *)dtmp.data())
                                : lfb_qr_f32(h0, (float *)copy.data(), rows, cols, cols, 1, (float *)dtmp.data());
            if (st == LFB_OK)
                for (int64_t i = 0; i < cols; ++i) for (int64_t j = 0; j < cols; ++j)
                    r_out[i * r_rs + j * r_cs] = i < j ? copy[i * cols + j] : (i == j ? std::fabs(dtmp[i]) : T(0));   // qr.rs:91-98
        }
        if (st != LFB_OK) m->err = lfb_last_error(h0);
        return st;
    }
    std::vector<TsqrDev<T>> st(G);
    std::vector<int64_t> b(G), e(G);
    for (int i = 0; i < G; ++i) shard_range(rows, G, i, &b[i], &e[i]);
    int rc = m->run_all([&](int i) {
        lfb_handle &h = *m->hs[i];
        TsqrDev<T> &d = st[i];
        d.rows = e[i] - b[i];
        d.ld = round_up(d.rows, 2);
        d.own.reset(new DevBuf<T>(h, (size_t)d.ld * cols));
        d.A = d.own->get();
        upload<T>(h, a + b[i] * rs, d.rows, cols, rs, cs, d.A, d.ld);
    });
    if (rc != LFB_OK) return rc;
    rc = tsqr_core<T>(m, st, cols, want_q);
    if (rc != LFB_OK) return rc;
    return m->run_all([&](int i) {
        lfb_handle &h = *m->hs[i];
        TsqrDev<T> &d = st[i];
        if (want_q) {
            download<T>(h, d.A, d.ld, a + b[i] * rs, d.rows, cols, rs, cs);
            if (i == 0) download_vec<T>(h, d.UD->get() + (size_t)round_up(cols, 2) * cols, cols, diag);
        } else if (i == 0) {
            download<T>(h, d.R->get(), cols, r_out, cols, cols, r_rs, r_cs);
        }
        LFB_CUDA(cudaStreamSynchronize(h.stream));
    });
}

template <typename T>
int tsqr_dev(lfb_multi *m, T *const *blocks, const int64_t *rows, int64_t cols, const int64_t *ld, T *const *d_diag, T *const *d_r,
             bool want_q) {
    if (!m) return LFB_INVALID_ARGUMENT;
    m->err.clear();
    if (!blocks || !rows || !ld || cols < 0) return mfail(m, LFB_INVALID_ARGUMENT, "bad arguments");
    if (cols == 0) return LFB_OK;
    const int G = (int)m->hs.size();
    for (int i = 0; i < G; ++i)
        if (rows[i] < cols) return mfail(m, LFB_NOT_THIN, "every device's row block must have rows >= cols (qr.rs:34-36 per block)");
    std::vector<TsqrDev<T>> st(G);
    for (int i = 0; i < G; ++i) { st[i].A = blocks[i]; st[i].rows = rows[i]; st[i].ld = ld[i]; }
    int rc = tsqr_core<T>(m, st, cols, want_q);
    if (rc != LFB_OK) return rc;
    return m->run_all([&](int i) {
        lfb_handle &h = *m->hs[i];
        TsqrDev<T> &d = st[i];
        if (d_r && d_r[i]) LFB_CUDA(cudaMemcpyAsync(d_r[i], d.R->get(), sizeof(T) * cols * cols, cudaMemcpyDeviceToDevice, h.stream));
        if (want_q && d_diag && d_diag[i])
            LFB_CUDA(cudaMemcpyAsync(d_diag[i], d.UD->get() + (size_t)round_up(cols, 2) * cols, sizeof(T) * cols, cudaMemcpyDeviceToDevice, h.stream));
    });
}

// ---- batch-sharded small factorisations: no collective ----
template <typename T>
int qr_batched_host_multi(lfb_multi *m, T *a, int64_t batch, int64_t mm, int64_t n, T *diag) {
    if (!m) return LFB_INVALID_ARGUMENT;
    m->err.clear();
    if (mm < n) return mfail(m, LFB_NOT_THIN, "Expected matrix rows >= cols");
    if (mm > 32 || n > 32) return mfail(m, LFB_UNSUPPORTED, "batched QR supports m, n <= 32");
    if (batch <= 0 || n == 0) return LFB_OK;
    const int G = (int)m->hs.size();
    return m->run_all([&](int i) {
        int64_t b, e;
        shard_range(batch, G, i, &b, &e);
        const int64_t nb = e - b;
        if (nb <= 0) return;
        lfb_handle &h = *m->hs[i];
        DevBuf<T> dA(h, (size_t)nb * mm * n), dD(h, (size_t)nb * n);
        LFB_CUDA(cudaMemcpyAsync(dA.get(), a + b * mm * n, sizeof(T) * nb * mm * n, cudaMemcpyHostToDevice, h.stream));
        qr_batched<T>(h, dA, nb, mm, n, dD);
        LFB_CUDA(cudaMemcpyAsync(a + b * mm * n, dA.get(), sizeof(T) * nb * mm * n, cudaMemcpyDeviceToHost, h.stream));
        LFB_CUDA(cudaMemcpyAsync(diag + b * n, dD.get(), sizeof(T) * nb * n, cudaMemcpyDeviceToHost, h.stream));
        LFB_CUDA(cudaStreamSynchronize(h.stream));
    });
}

template <typename T>
int cholesky_batched_host_multi(lfb_multi *m, T *a, int64_t batch, int64_t n, int clean, int64_t *fail_matrix, int64_t *fail_index) {
    if (!m) return LFB_INVALID_ARGUMENT;
    m->err.clear();
    if (fail_matrix) *fail_matrix = -1;
    if (fail_index) *fail_index = -1;
    if (n > 32) return mfail(m, LFB_UNSUPPORTED, "batched Cholesky supports n <= 32");
    if (batch <= 0 || n <= 0) return LFB_OK;
    const int G = (int)m->hs.size();
    std::vector<int> hf((size_t)batch);
    int rc = m->run_all([&](int i) {
        int64_t b, e;
        shard_range(batch, G, i, &b, &e);
        const int64_t nb = e - b;
        if (nb <= 0) return;
        lfb_handle &h = *m->hs[i];
        DevBuf<T> dA(h, (size_t)nb * n * n);
        DevBuf<int> dF(h, (size_t)nb);
        LFB_CUDA(cudaMemcpyAsync(dA.get(), a + b * n * n, sizeof(T) * nb * n * n, cudaMemcpyHostToDevice, h.stream));
        cholesky_batched<T>(h, dA, nb, n, clean, dF);
        LFB_CUDA(cudaMemcpyAsync(a + b * n * n, dA.get(), sizeof(T) * nb * n * n, cudaMemcpyDeviceToHost, h.stream));
        LFB_CUDA(cudaMemcpyAsync(hf.data() + b, dF.get(), sizeof(int) * nb, cudaMemcpyDeviceToHost, h.stream));
        LFB_CUDA(cudaStreamSynchronize(h.stream));
    });
    if (rc != LFB_OK) return rc;
    for (int64_t k = 0; k < batch; ++k)      // the first failure in BATCH order, whichever device met it (cholesky.rs:69-71 per matrix)
        if (hf[k] >= 0) {
            if (fail_matrix) *fail_matrix = k;
            if (fail_index) *fail_index = hf[k];
            return mfail(m, LFB_NOT_POSITIVE_DEFINITE, "Matrix is not positive definite");
        }
    return LFB_OK;
}

}  // namespace

extern "C" {

int lfb_create_multi(lfb_multi **out, const int *devices, int n_devices) {
    if (!out) return LFB_INVALID_ARGUMENT;
    *out = nullptr;
    if (n_devices < 1 || n_devices > 64) return LFB_INVALID_ARGUMENT;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) { cudaGetLastError(); return LFB_ERR_CUDA; }   // no CPU fallback
    std::unique_ptr<lfb_multi> m(new lfb_multi());
    for (int i = 0; i < n_devices; ++i) {
        const int dev = devices ? devices[i] : i;
        if (dev < 0 || dev >= count) return LFB_INVALID_ARGUMENT;
        for (int j = 0; j < i; ++j) if (m->devices[j] == dev) return LFB_INVALID_ARGUMENT;   // one handle per device
        m->devices.push_back(dev);
    }
    auto cleanup = [&] { for (auto *h : m->hs) lfb_destroy(h); m->hs.clear(); };
    for (int dev : m->devices) {
        lfb_handle *h = nullptr;
        if (lfb_create(&h, dev) != LFB_OK) { cleanup(); return LFB_ERR_CUDA; }
        m->hs.push_back(h);
    }
    if (n_devices > 1) {
        if (!m->nccl.load()) { cleanup(); fprintf(stderr, "liblinfa_b200: %s\n", m->nccl.why.c_str()); return LFB_ERR_NCCL; }
        m->comms.resize(n_devices);
        ncclResult_t r = m->nccl.CommInitAll(m->comms.data(), n_devices, m->devices.data());
        if (r != ncclSuccess) {
            fprintf(stderr, "liblinfa_b200: ncclCommInitAll failed: %s\n", m->nccl.GetErrorString(r));
            m->comms.clear(); cleanup();
            return LFB_ERR_NCCL;
        }
        m->nccl.CommCount(m->comms[0], &m->nccl_ranks);
        m->nccl.GetVersion(&m->nccl_version);
    }
    for (int k = 0; k < 2; ++k) {
        m->tev[k].resize(n_devices);
        for (int i = 0; i < n_devices; ++i) {
            cudaSetDevice(m->devices[i]);
            if (cudaEventCreate(&m->tev[k][i]) != cudaSuccess) { cudaGetLastError(); m->tev[k][i] = nullptr; }
        }
    }
    m->wcode.assign(n_devices, LFB_OK);
    m->werr.assign(n_devices, "");
    lfb_multi *raw = m.release();
    for (int i = 0; i < n_devices; ++i) raw->threads.emplace_back([raw, i] { raw->worker(i); });
    *out = raw;
    return LFB_OK;
}

int lfb_destroy_multi(lfb_multi *m) {
    if (!m) return LFB_OK;
    {
        std::lock_guard<std::mutex> lk(m->mu);
        m->stop = true;
    }
    m->cv_start.notify_all();
    for (auto &t : m->threads) t.join();
    for (size_t i = 0; i < m->hs.size(); ++i) {
        cudaSetDevice(m->devices[i]);
        cudaStreamSynchronize(m->hs[i]->stream);
    }
    for (auto c : m->comms) if (c) m->nccl.CommDestroy(c);
    for (int k = 0; k < 2; ++k)
        for (size_t i = 0; i < m->tev[k].size(); ++i) if (m->tev[k][i]) { cudaSetDevice(m->devices[i]); cudaEventDestroy(m->tev[k][i]); }
    for (auto *h : m->hs) lfb_destroy(h);
    delete m;
    return LFB_OK;
}

const char *lfb_multi_last_error(lfb_multi *m) { return m ? m->err.c_str() : "null handle"; }
int lfb_multi_device_count(lfb_multi *m) { return m ? (int)m->hs.size() : 0; }
int lfb_multi_nccl_ranks(lfb_multi *m) { return m ? m->nccl_ranks : 0; }
int lfb_multi_nccl_version(lfb_multi *m) { return m ? m->nccl_version : 0; }
lfb_handle *lfb_multi_handle(lfb_multi *m, int i) { return (m && i >= 0 && i < (int)m->hs.size()) ? m->hs[i] : nullptr; }

int lfb_multi_set_option(lfb_multi *m, const char *key, int64_t value) {
    if (!m) return LFB_INVALID_ARGUMENT;
    for (auto *h : m->hs) {
        const int st = lfb_set_option(h, key, value);
        if (st != LFB_OK) return st;
    }
    return LFB_OK;
}

int64_t lfb_multi_launch_count(lfb_multi *m) {
    int64_t s = 0;
    if (m) for (auto *h : m->hs) s += h->launches;
    return s;
}

int lfb_multi_synchronize(lfb_multi *m) {
    if (!m) return LFB_INVALID_ARGUMENT;
    for (size_t i = 0; i < m->hs.size(); ++i) {
        if (cudaSetDevice(m->devices[i]) != cudaSuccess || cudaStreamSynchronize(m->hs[i]->stream) != cudaSuccess) {
            m->err = cudaGetErrorString(cudaGetLastError());
            return LFB_ERR_CUDA;
        }
    }
    return LFB_OK;
}

// Device-side timing across devices: `begin` synchronises every stream and records a start event on each; `end` records a
// stop event on each stream, synchronises, and returns the MAX over devices of the elapsed time (the multi-GPU timing rule).
int lfb_multi_time_begin(lfb_multi *m) {
    if (!m) return LFB_INVALID_ARGUMENT;
    int st = lfb_multi_synchronize(m);
    if (st != LFB_OK) return st;
    for (size_t i = 0; i < m->hs.size(); ++i) {
        cudaSetDevice(m->devices[i]);
        if (cudaEventRecord(m->tev[0][i], m->hs[i]->stream) != cudaSuccess) return mfail(m, LFB_ERR_CUDA, "cudaEventRecord failed");
    }
    return LFB_OK;
}
int lfb_multi_time_end(lfb_multi *m, double *max_ms) {
    if (!m || !max_ms) return LFB_INVALID_ARGUMENT;
    for (size_t i = 0; i < m->hs.size(); ++i) {
        cudaSetDevice(m->devices[i]);
        if (cudaEventRecord(m->tev[1][i], m->hs[i]->stream) != cudaSuccess) return mfail(m, LFB_ERR_CUDA, "cudaEventRecord failed");
    }
    double mx = 0.0;
    for (size_t i = 0; i < m->hs.size(); ++i) {
        cudaSetDevice(m->devices[i]);
        float t = 0.f;
        if (cudaEventSynchronize(m->tev[1][i]) != cudaSuccess || cudaEventElapsedTime(&t, m->tev[0][i], m->tev[1][i]) != cudaSuccess)
            return mfail(m, LFB_ERR_CUDA, "event timing failed");
        mx = std::max(mx, (double)t);
    }
    *max_ms = mx;
    return LFB_OK;
}

// ---- qr.rs:29-45 on a row-sharded tall-skinny matrix ----
int lfb_qr_tsqr_multi_f64(lfb_multi *m, double *a, int64_t rows, int64_t cols, int64_t rs, int64_t cs, double *diag) {
    if (!diag && cols > 0) return mfail(m, LFB_INVALID_ARGUMENT, "diag is null");
    return tsqr_host<double>(m, a, rows, cols, rs, cs, diag, nullptr, 0, 0);
}
int lfb_qr_tsqr_multi_f32(lfb_multi *m, float *a, int64_t rows, int64_t cols, int64_t rs, int64_t cs, float *diag) {
    if (!diag && cols > 0) return mfail(m, LFB_INVALID_ARGUMENT, "diag is null");
    return tsqr_host<float>(m, a, rows, cols, rs, cs, diag, nullptr, 0, 0);
}
int lfb_tsqr_r_multi_f64(lfb_multi *m, const double *a, int64_t rows, int64_t cols, int64_t rs, int64_t cs, double *r, int64_t r_rs, int64_t r_cs) {
    if (!r && cols > 0) return mfail(m, LFB_INVALID_ARGUMENT, "r is null");
    return tsqr_host<double>(m, const_cast<double *>(a), rows, cols, rs, cs, nullptr, r, r_rs, r_cs);
}
int lfb_tsqr_r_multi_f32(lfb_multi *m, const float *a, int64_t rows, int64_t cols, int64_t rs, int64_t cs, float *r, int64_t r_rs, int64_t r_cs) {
    if (!r && cols > 0) return mfail(m, LFB_INVALID_ARGUMENT, "r is null");
    return tsqr_host<float>(m, const_cast<float *>(a), rows, cols, rs, cs, nullptr, r, r_rs, r_cs);
}
int lfb_qr_tsqr_multi_dev_f64(lfb_multi *m, double *const *d_blocks, const int64_t *rows, int64_t cols, const int64_t *ld,
                              double *const *d_diag, double *const *d_r) {
    return tsqr_dev<double>(m, d_blocks, rows, cols, ld, d_diag, d_r, true);
}
int lfb_tsqr_r_multi_dev_f64(lfb_multi *m, double *const *d_blocks, const int64_t *rows, int64_t cols, const int64_t *ld, double *const *d_r) {
    return tsqr_dev<double>(m, d_blocks, rows, cols, ld, nullptr, d_r, false);
}

// ---- batch-sharded ----
int lfb_qr_batched_multi_f32(lfb_multi *m, float *a, int64_t batch, int64_t mm, int64_t n, float *diag) { return qr_batched_host_multi<float>(m, a, batch, mm, n, diag); }
int lfb_qr_batched_multi_f64(lfb_multi *m, double *a, int64_t batch, int64_t mm, int64_t n, double *diag) { return qr_batched_host_multi<double>(m, a, batch, mm, n, diag); }
int lfb_cholesky_batched_multi_f32(lfb_multi *m, float *a, int64_t batch, int64_t n, int clean, int64_t *fm, int64_t *fi) {
    return cholesky_batched_host_multi<float>(m, a, batch, n, clean, fm, fi);
}
int lfb_cholesky_batched_multi_f64(lfb_multi *m, double *a, int64_t batch, int64_t n, int clean, int64_t *fm, int64_t *fi) {
    return cholesky_batched_host_multi<double>(m, a, batch, n, clean, fm, fi);
}
int lfb_qr_batched_multi_dev_f32(lfb_multi *m, float *const *d_a, const int64_t *batch, int64_t mm, int64_t n, float *const *d_diag) {
    if (!m) return LFB_INVALID_ARGUMENT;
    m->err.clear();
    if (!d_a || !batch || !d_diag) return mfail(m, LFB_INVALID_ARGUMENT, "bad arguments");
    if (mm < n) return mfail(m, LFB_NOT_THIN, "Expected matrix rows >= cols");
    if (mm > 32 || n > 32) return mfail(m, LFB_UNSUPPORTED, "batched QR supports m, n <= 32");
    // one kernel launch per device: enqueued from the calling thread, asynchronous on every handle's stream
    for (size_t i = 0; i < m->hs.size(); ++i) {
        if (batch[i] <= 0) continue;
        try {
            LFB_CUDA(cudaSetDevice(m->devices[i]));
            qr_batched<float>(*m->hs[i], d_a[i], batch[i], mm, n, d_diag[i]);
        } catch (const lfb::CudaError &e) {
            m->err = e.what();
            cudaGetLastError();
            return e.code;
        }
    }
    return LFB_OK;
}

}  // extern "C"
