// extern "C" entry points of liblinfa_b200.so (include/linfa_b200.h): argument checks that mirror
// the reference's LinalgError behaviour, packing of arbitrary-stride host views into the engine's
// column-major HBM layout, and the device-resident variants.
#include <memory>

#include <chrono>

#include "common.cuh"
#include "host_io.cuh"

using namespace lfb;

namespace {

int fail(lfb_handle *h, int code, const char *msg) {
    if (h) h->err = msg;
    return code;
}

#define LFB_API_BEGIN(h)                                   \
    if (!(h)) return LFB_INVALID_ARGUMENT;                 \
    (h)->err.clear();                                      \
    try {                                                  \
        LFB_CUDA(cudaSetDevice((h)->device));
#define LFB_API_END(h)                                     \
    }                                                      \
    catch (const lfb::CudaError &e) {                      \
        (h)->err = e.what();                               \
        cudaGetLastError();                                \
        return e.code;                                     \
    }                                                      \
    catch (const std::exception &e) {                      \
        (h)->err = e.what();                               \
        return LFB_ERR_CUDA;                               \
    }                                                      \
    return LFB_OK;

// ---------------------------------------------------------------------------------------------
// Tall-skinny route of qr_into: TSQR + Householder reconstruction (tsqr_hr.cu) delivers the same compact factor.
// Taken when the caller asks for it (lfb_qr_tsqr_*), or, with option qr_tsqr_auto, whenever the matrix is tall enough
// to be cut into at least two chunks and skinny enough for the single-CTA LU of the reconstruction.
inline bool tsqr_route(const lfb_handle &h, int64_t rows, int64_t cols, bool force) {
    if (cols < 1) return false;
    if (force) return true;
    const int64_t CH = std::max<int64_t>(h.opt.tsqr_chunk, 2 * cols);
    return h.opt.qr_tsqr_auto != 0 && cols <= 512 && rows >= 2 * CH;
}

template <typename T>
int qr_host(lfb_handle *h, T *a, int64_t rows, int64_t cols, int64_t rs, int64_t cs, T *diag, bool force_tsqr = false) {
    if (rows < 0 || cols < 0) return fail(h, LFB_INVALID_ARGUMENT, "negative dimension");
    if (rows < cols) return fail(h, LFB_NOT_THIN, "Expected matrix rows >= cols");           // qr.rs:34-36
    if (cols == 0) return LFB_OK;                                                            // 0x0 legal, qr.rs:383-388
    LFB_API_BEGIN(h)
    const int64_t ld = round_up(rows, 4);     // 16-byte columns for TMA in both precisions
    DevBuf<T> dA(*h, (size_t)ld * cols), dD(*h, cols);
    upload<T>(*h, a, rows, cols, rs, cs, dA, ld);
    if (tsqr_route(*h, rows, cols, force_tsqr)) {
        qr_tsqr<T>(*h, dA, rows, cols, ld, dD);
        download<T>(*h, dA, ld, a, rows, cols, rs, cs);
        download_vec<T>(*h, dD, cols, diag);
    } else {
        // Large matrix in page-locked host memory: every block column of the factor starts its way back as soon as its panel is
        // factored and signed (householder.cu: finish_panel), on a copy stream behind an event -- the D2H traffic (2 GiB at
        // 16384^2 f64, ~43 ms) hides behind the trailing updates, as in cholesky_host.
        int64_t hld = 0;
        const Layout lay = classify(rows, cols, rs, cs, &hld);
        bool pinned_host = false;
        {
            cudaPointerAttributes pa;
            if (cudaPointerGetAttributes(&pa, a) == cudaSuccess) pinned_host = pa.type == cudaMemoryTypeHost || pa.type == cudaMemoryTypeManaged;
            else cudaGetLastError();
        }
        const bool overlap = h->opt.qr_overlap_d2h && pinned_host && lay != L_GEN && cols >= 2048;
        std::vector<cudaEvent_t> pev;
        std::unique_ptr<DevBuf<T>> stage;
        if (overlap) {
            if (!h->copy_stream) LFB_CUDA(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
            if (lay == L_ROW) stage.reset(new DevBuf<T>(*h, (size_t)rows * cols));
            lfb_handle *hh = h;
            T *dAp = dA.get();
            T *stg = stage ? stage->get() : nullptr;
            h->qr_panel_hook = [hh, dAp, stg, a, rows, ld, hld, lay, &pev](int64_t k0, int64_t nb) {
                cudaEvent_t e;
                LFB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
                pev.push_back(e);
                LFB_CUDA(cudaEventRecord(e, hh->stream));
                LFB_CUDA(cudaStreamWaitEvent(hh->copy_stream, e, 0));
                if (lay == L_COL) {
                    LFB_CUDA(cudaMemcpy2DAsync(a + k0 * hld, hld * sizeof(T), dAp + k0 * ld, ld * sizeof(T), rows * sizeof(T), nb,
                                               cudaMemcpyDeviceToHost, hh->copy_stream));
                } else {                 // row-major host: the block column as rows of nb entries
                    T *t = stg + (size_t)k0 * rows;
                    cudaStream_t keep = hh->stream;
                    hh->stream = hh->copy_stream;
                    try {
                        transpose<T>(*hh, dAp + k0 * ld, rows, nb, ld, t, nb);
                    } catch (...) {
                        hh->stream = keep;
                        throw;
                    }
                    hh->stream = keep;
                    LFB_CUDA(cudaMemcpy2DAsync(a + k0, hld * sizeof(T), t, nb * sizeof(T), nb * sizeof(T), rows, cudaMemcpyDeviceToHost,
                                               hh->copy_stream));
                }
            };
        }
        try {
            qr_factor<T>(*h, dA, rows, cols, ld, dD);
        } catch (...) {
            h->qr_panel_hook = nullptr;
            if (h->copy_stream) cudaStreamSynchronize(h->copy_stream);
            cudaStreamSynchronize(h->stream);
            for (auto e : pev) cudaEventDestroy(e);
            throw;
        }
        h->qr_panel_hook = nullptr;
        if (overlap) {
            download_vec<T>(*h, dD, cols, diag);
            LFB_CUDA(cudaStreamSynchronize(h->copy_stream));
            LFB_CUDA(cudaStreamSynchronize(h->stream));
            for (auto e : pev) cudaEventDestroy(e);
        } else {
            download<T>(*h, dA, ld, a, rows, cols, rs, cs);
            download_vec<T>(*h, dD, cols, diag);
        }
    }
    LFB_API_END(h)
}

template <typename T>
int assemble_q_host(lfb_handle *h, const T *m, int64_t rows, int64_t cols, int64_t rs, int64_t cs, int64_t shift,
                    const T *signs, T *q, int64_t q_rs, int64_t q_cs) {
    if (rows < 0 || cols < 0 || shift < 0) return fail(h, LFB_INVALID_ARGUMENT, "negative dimension");
    const int64_t dim = std::min(rows, cols);
    if (shift > dim) return fail(h, LFB_INVALID_ARGUMENT, "shift exceeds matrix dimension");   // householder.rs:67 (panics there)
    if (rows == 0 || dim == 0) return LFB_OK;
    LFB_API_BEGIN(h)
    const int64_t ld = round_up(rows, 2);
    const int64_t nref = dim - shift;
    DevBuf<T> dM(*h, (size_t)ld * cols), dQ(*h, (size_t)ld * dim), dS(*h, std::max<int64_t>(nref, 1));
    upload<T>(*h, m, rows, cols, rs, cs, dM, ld);
    upload_vec<T>(*h, signs, nref, dS);
    assemble_q<T>(*h, dM, rows, cols, ld, shift, dS, dQ, ld);
    download<T>(*h, dQ, ld, q, rows, dim, q_rs, q_cs);
    LFB_API_END(h)
}

template <typename T>
int qt_mul_host(lfb_handle *h, const T *qr, int64_t rows, int64_t cols, int64_t rs, int64_t cs, const T *diag, T *b,
                int64_t bcols, int64_t b_rs, int64_t b_cs) {
    if (rows < cols) return fail(h, LFB_NOT_THIN, "Expected matrix rows >= cols");
    if (rows == 0 || cols == 0 || bcols == 0) return LFB_OK;
    LFB_API_BEGIN(h)
    const int64_t ld = round_up(rows, 2);
    DevBuf<T> dM(*h, (size_t)ld * cols), dB(*h, (size_t)ld * bcols), dD(*h, cols);
    upload<T>(*h, qr, rows, cols, rs, cs, dM, ld);
    upload<T>(*h, b, rows, bcols, b_rs, b_cs, dB, ld);
    upload_vec<T>(*h, diag, cols, dD);
    qt_mul<T>(*h, dM, rows, cols, ld, dD, dB, bcols, ld);
    download<T>(*h, dB, ld, b, rows, bcols, b_rs, b_cs);
    LFB_API_END(h)
}

// Strided 2-D block of PAGEABLE caller memory (height rows of `width` elements, `pitch` elements apart) <-> a compact pinned
// buffer (pitch = width), copied by the handle's host threads in row groups of ~1 MiB.
template <typename T>
void host_gather(lfb_handle &h, const T *src, int64_t pitch, int64_t width, int64_t height, T *dst) {
    const int64_t per = std::max<int64_t>(1, (int64_t)(1 << 20) / std::max<int64_t>(1, width * (int64_t)sizeof(T)));
    const int jobs = (int)cdiv(height, per);
    h.pool().run(jobs, [&](int j) {
        const int64_t r0 = j * per, r1 = std::min(height, r0 + per);
        for (int64_t r = r0; r < r1; ++r) memcpy(dst + r * width, src + r * pitch, sizeof(T) * width);
    });
}
template <typename T>
void host_scatter(lfb_handle &h, const T *src, int64_t width, int64_t height, T *dst, int64_t pitch) {
    const int64_t per = std::max<int64_t>(1, (int64_t)(1 << 20) / std::max<int64_t>(1, width * (int64_t)sizeof(T)));
    const int jobs = (int)cdiv(height, per);
    h.pool().run(jobs, [&](int j) {
        const int64_t r0 = j * per, r1 = std::min(height, r0 + per);
        for (int64_t r = r0; r < r1; ++r) memcpy(dst + r * pitch, src + r * width, sizeof(T) * width);
    });
}

template <typename T>
int cholesky_host(lfb_handle *h, T *a, int64_t rows, int64_t cols, int64_t rs, int64_t cs, int clean, int64_t *fail_index) {
    if (rows != cols) return fail(h, LFB_NOT_SQUARE, "Matrix is not square");                 // lib.rs:64-71
    if (fail_index) *fail_index = -1;
    const int64_t n = rows;
    if (n == 0) return LFB_OK;                                                               // cholesky.rs:270-272
    int64_t info = 0;
    const auto t_entry = std::chrono::steady_clock::now();
    auto since_entry = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_entry).count(); };
    double t_enq0 = 0, t_enq1 = 0;
    LFB_API_BEGIN(h)
    const int64_t ld = round_up(n, 4);      // 16-byte columns for TMA in both precisions
    DevBuf<T> dA(*h, (size_t)ld * n);
    DevBuf<int64_t> dInfo(*h, 1);
    // The reference reads (and, for the dirty variant, writes) ONLY the lower triangle
    // (cholesky.rs:51-76), so only lower-triangular trapezoids cross PCIe: ~53 % of the bytes.
    int64_t hld = 0;
    const Layout lay = classify(n, n, rs, cs, &hld);
    const bool tri = lay != L_GEN && n >= 2048;
    // Bands of the trapezoid upload.  A diagonal block of the factorisation (chol_nb wide, a multiple of 64) must lie
    // inside ONE band: the overlapped D2H hook below copies whole block columns including the upper part of their
    // diagonal block, which is only defined (= the caller's own data) if that block was uploaded as part of a band.
    const int64_t NBc = std::max<int64_t>(64, round_up(h->opt.chol_nb, 64));
    const int64_t BAND = round_up(std::max<int64_t>(1024, NBc), NBc);
    // Pageable caller memory (what an ndarray owns): cudaMemcpy from it is staged by the driver at ~6-10 GB/s and blocks the
    // calling thread.  Instead the handle's host threads gather every band into the handle's pinned buffer while the DMA of
    // the previous band runs, and on the way back scatter each finished block column while the factorisation continues.
    bool pinned_host = false;
    {
        cudaPointerAttributes pa;
        if (cudaPointerGetAttributes(&pa, a) == cudaSuccess) pinned_host = pa.type == cudaMemoryTypeHost || pa.type == cudaMemoryTypeManaged;
        else cudaGetLastError();
    }
    const bool staged = tri && !pinned_host && h->opt.host_staging;
    T *stagebuf = nullptr;
    if (staged) {
        size_t tri_elems = 0;
        for (int64_t r0 = 0; r0 < n; r0 += BAND) tri_elems += (size_t)(std::min(n, r0 + BAND) - r0) * (size_t)std::max(std::min(n, r0 + BAND), n - r0);
        stagebuf = (T *)h->pinned_buf(sizeof(T) * tri_elems);
    }
    // Dirty factorisation of a large matrix: every finished block column of L starts its way back to the host as
    // soon as its panel is factored (a copy stream waits on an event recorded behind the panel), so the D2H traffic
    // hides behind the trailing updates instead of following the factorisation.
    // Only for page-locked host memory: a D2H copy into pageable memory is staged synchronously and would stall
    // the thread that is still enqueueing the factorisation.
    const bool overlap = tri && !clean && h->opt.chol_overlap_d2h && (pinned_host || staged);
    std::vector<cudaEvent_t> pev;
    struct Piece { cudaEvent_t done; int64_t k0, nb; T *st; };   // staged: block columns waiting for their host scatter
    std::vector<Piece> pieces;
    size_t stage_off = 0;
    std::unique_ptr<DevBuf<T>> stage;
    if (overlap) {
        if (!h->copy_stream) LFB_CUDA(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
        if (lay != L_COL) stage.reset(new DevBuf<T>(*h, (size_t)n * n));
        lfb_handle *hh = h;
        T *dAp = dA.get();
        T *stg = stage ? stage->get() : nullptr;
        h->chol_panel_hook = [hh, dAp, stg, a, n, ld, hld, lay, &pev, staged, stagebuf, &pieces, &stage_off](int64_t k0, int64_t nb) {
            cudaEvent_t e;
            LFB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            pev.push_back(e);
            LFB_CUDA(cudaEventRecord(e, hh->stream));
            LFB_CUDA(cudaStreamWaitEvent(hh->copy_stream, e, 0));
            const int64_t below = n - k0;
            if (staged) {            // pageable caller memory: device -> pinned staging now, host threads scatter it later
                T *st = stagebuf + stage_off;
                stage_off += (size_t)below * nb;
                if (lay == L_COL) {
                    LFB_CUDA(cudaMemcpy2DAsync(st, below * sizeof(T), dAp + k0 + k0 * ld, ld * sizeof(T), below * sizeof(T), nb,
                                               cudaMemcpyDeviceToHost, hh->copy_stream));
                } else {
                    T *t = stg + (size_t)k0 * n;
                    cudaStream_t keep = hh->stream;
                    hh->stream = hh->copy_stream;
                    try {
                        transpose<T>(*hh, dAp + k0 + k0 * ld, below, nb, ld, t, nb);
                    } catch (...) {
                        hh->stream = keep;
                        throw;
                    }
                    hh->stream = keep;
                    LFB_CUDA(cudaMemcpyAsync(st, t, sizeof(T) * below * nb, cudaMemcpyDeviceToHost, hh->copy_stream));
                }
                cudaEvent_t d;
                LFB_CUDA(cudaEventCreateWithFlags(&d, cudaEventDisableTiming));
                pev.push_back(d);
                LFB_CUDA(cudaEventRecord(d, hh->copy_stream));
                pieces.push_back({d, k0, nb, st});
                return;
            }
            if (lay == L_COL) {      // host column-major: the block column is a 2-D copy as it is
                LFB_CUDA(cudaMemcpy2DAsync(a + k0 + k0 * hld, hld * sizeof(T), dAp + k0 + k0 * ld, ld * sizeof(T), below * sizeof(T), nb,
                                           cudaMemcpyDeviceToHost, hh->copy_stream));
            } else {                 // host row-major: transpose the block column into nb-wide rows first
                T *t = stg + (size_t)k0 * n;
                cudaStream_t keep = hh->stream;
                hh->stream = hh->copy_stream;
                try {
                    transpose<T>(*hh, dAp + k0 + k0 * ld, below, nb, ld, t, nb);
                } catch (...) {
                    hh->stream = keep;
                    throw;
                }
                hh->stream = keep;
                LFB_CUDA(cudaMemcpy2DAsync(a + k0 * hld + k0, hld * sizeof(T), t, nb * sizeof(T), nb * sizeof(T), below, cudaMemcpyDeviceToHost,
                                           hh->copy_stream));
            }
        };
    }
    // Arrival waves (n >= 8192): with the upload first, the 20 ms the lower trapezoid of a 16384^2 f64 matrix spends on PCIe are
    // fully exposed (a right-looking step 0 touches every column).  Instead the block columns cross in order on their own stream
    // and the factorisation runs in waves over the columns that have arrived: wave w = one large-K catch-up product for its
    // trapezoid + a right-looking sweep restricted to its columns (potrf.cu: cholesky_lower_wave).  Boundaries from the model
    // time(w) = max(arrival, previous wave) + flops(w) / rate: upload-bound up to ~0.3 n, compute-bound afterwards.
    std::vector<int64_t> wend;
    if (tri && n >= 8192 && h->opt.chol_waves > 1) {
        static const double frac[3][3] = {{0.28, 0, 0}, {0.16, 0.31, 0}, {0.09, 0.19, 0.34}};
        const int nwv = (int)std::min<int64_t>(h->opt.chol_waves, 4);
        for (int w = 0; w + 1 < nwv; ++w) {
            int64_t e = std::min<int64_t>(n, std::max<int64_t>(BAND, (int64_t)(frac[nwv - 2][w] * (double)n / (double)BAND + 0.5) * BAND));
            if (wend.empty() || e > wend.back()) wend.push_back(e);
        }
        if (wend.empty() || wend.back() < n) wend.push_back(n);
        if (wend.size() < 2) wend.clear();
    }
    const bool waves = !wend.empty();
    std::unique_ptr<DevBuf<T>> rowtmp;
    if (waves) {
        // uploads are enqueued wave by wave below, next to the compute that consumes them
    } else if (!tri) {
        upload<T>(*h, a, n, n, rs, cs, dA, ld);
    } else if (staged) {
        T *sp = stagebuf;
        if (lay != L_COL) rowtmp.reset(new DevBuf<T>(*h, (size_t)hld * n));
        for (int64_t b0 = 0; b0 < n; b0 += BAND) {
            const int64_t b1 = std::min(n, b0 + BAND);
            if (lay == L_COL) {      // columns b0..b1, rows b0..n-1 of each: width n - b0, height b1 - b0, pitch hld
                host_gather<T>(*h, a + b0 + b0 * hld, hld, n - b0, b1 - b0, sp);
                LFB_CUDA(cudaMemcpy2DAsync(dA.get() + b0 + b0 * ld, ld * sizeof(T), sp, (n - b0) * sizeof(T), (n - b0) * sizeof(T), b1 - b0,
                                           cudaMemcpyHostToDevice, h->stream));
                sp += (size_t)(n - b0) * (b1 - b0);
            } else {                 // rows b0..b1, columns 0..b1-1 of each: width b1, height b1 - b0
                host_gather<T>(*h, a + b0 * hld, hld, b1, b1 - b0, sp);
                LFB_CUDA(cudaMemcpy2DAsync(rowtmp->get() + b0 * hld, hld * sizeof(T), sp, b1 * sizeof(T), b1 * sizeof(T), b1 - b0,
                                           cudaMemcpyHostToDevice, h->stream));
                sp += (size_t)b1 * (b1 - b0);
            }
        }
        if (lay != L_COL) transpose<T>(*h, rowtmp->get(), n, n, hld, dA, ld);
    } else if (lay == L_COL) {   // column j holds rows j..n-1: bands of columns
        for (int64_t c0 = 0; c0 < n; c0 += BAND) {
            const int64_t c1 = std::min(n, c0 + BAND);
            LFB_CUDA(cudaMemcpy2DAsync(dA.get() + c0 + c0 * ld, ld * sizeof(T), a + c0 + c0 * hld, hld * sizeof(T),
                                       (n - c0) * sizeof(T), c1 - c0, cudaMemcpyHostToDevice, h->stream));
        }
    } else {                     // row-major: row i holds columns 0..i: bands of rows, then transpose on the device
        DevBuf<T> tmp(*h, (size_t)hld * n);
        for (int64_t r0 = 0; r0 < n; r0 += BAND) {
            const int64_t r1 = std::min(n, r0 + BAND);
            LFB_CUDA(cudaMemcpy2DAsync(tmp.get() + r0 * hld, hld * sizeof(T), a + r0 * hld, hld * sizeof(T), r1 * sizeof(T), r1 - r0,
                                       cudaMemcpyHostToDevice, h->stream));
        }
        transpose<T>(*h, tmp.get(), n, n, hld, dA, ld);   // the unread upper triangle carries whatever tmp held
    }
    std::vector<cudaEvent_t> wev;
    try {
        if (!waves) {
            cholesky_lower<T>(*h, dA, n, ld, clean, dInfo);
        } else {
            if (!h->upload_stream) LFB_CUDA(cudaStreamCreateWithFlags(&h->upload_stream, cudaStreamNonBlocking));
            cudaStream_t us = h->upload_stream, ms = h->stream;
            {   // the upload stream starts behind whatever the caller's stream holds (dA may be a recycled block)
                cudaEvent_t e0;
                LFB_CUDA(cudaEventCreateWithFlags(&e0, cudaEventDisableTiming));
                wev.push_back(e0);
                LFB_CUDA(cudaEventRecord(e0, ms));
                LFB_CUDA(cudaStreamWaitEvent(us, e0, 0));
            }
            static const bool wdbg = getenv("LFB_WAVE_DBG") != nullptr && atoi(getenv("LFB_WAVE_DBG")) == 1;      // debug: event time stamps of every wave on stderr (2: host time stamps only)
            std::vector<cudaEvent_t> tev;
            auto stamp = [&](cudaStream_t st) {
                if (!wdbg) return;
                cudaEvent_t e;
                LFB_CUDA(cudaEventCreate(&e));
                LFB_CUDA(cudaEventRecord(e, st));
                tev.push_back(e);
            };
            stamp(ms);
            t_enq0 = since_entry();
            size_t tmp_elems = 0;
            if (lay != L_COL)
                for (int64_t c0 = 0; c0 < n; c0 += BAND) tmp_elems += (size_t)(n - c0) * (size_t)(std::min(n, c0 + BAND) - c0);
            if (lay != L_COL) rowtmp.reset(new DevBuf<T>(*h, tmp_elems));
            T *sp = stagebuf;
            T *tp = rowtmp ? rowtmp->get() : nullptr;
            int64_t c_begin = 0;
            for (size_t w = 0; w < wend.size(); ++w) {
                const int64_t c_end = wend[w];
                for (int64_t c0 = c_begin; c0 < c_end; c0 += BAND) {       // block columns c0..c1, rows c0..n-1
                    const int64_t c1 = std::min(c_end, c0 + BAND), wd = c1 - c0, ht = n - c0;
                    if (lay == L_COL) {
                        if (staged) {
                            host_gather<T>(*h, a + c0 + c0 * hld, hld, ht, wd, sp);
                            LFB_CUDA(cudaMemcpy2DAsync(dA.get() + c0 + c0 * ld, ld * sizeof(T), sp, ht * sizeof(T), ht * sizeof(T), wd,
                                                       cudaMemcpyHostToDevice, us));
                            sp += (size_t)ht * wd;
                        } else {
                            LFB_CUDA(cudaMemcpy2DAsync(dA.get() + c0 + c0 * ld, ld * sizeof(T), a + c0 + c0 * hld, hld * sizeof(T), ht * sizeof(T), wd,
                                                       cudaMemcpyHostToDevice, us));
                        }
                    } else {             // row-major host: row r holds this block column as wd contiguous entries
                        if (staged) {
                            host_gather<T>(*h, a + c0 * hld + c0, hld, wd, ht, sp);
                            LFB_CUDA(cudaMemcpyAsync(tp, sp, sizeof(T) * (size_t)ht * wd, cudaMemcpyHostToDevice, us));
                            sp += (size_t)ht * wd;
                        } else {
                            LFB_CUDA(cudaMemcpy2DAsync(tp, wd * sizeof(T), a + c0 * hld + c0, hld * sizeof(T), wd * sizeof(T), ht,
                                                       cudaMemcpyHostToDevice, us));
                        }
                        h->stream = us;
                        try {
                            transpose<T>(*h, tp, wd, ht, wd, dA.get() + c0 + c0 * ld, ld);
                        } catch (...) {
                            h->stream = ms;
                            throw;
                        }
                        h->stream = ms;
                        tp += (size_t)ht * wd;
                    }
                }
                cudaEvent_t e;
                LFB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
                wev.push_back(e);
                LFB_CUDA(cudaEventRecord(e, us));
                stamp(us);
                stamp(ms);
                LFB_CUDA(cudaStreamWaitEvent(ms, e, 0));
                cholesky_lower_wave<T>(*h, dA, n, ld, c_begin, c_end, dInfo, w == 0);
                stamp(ms);
                c_begin = c_end;
            }
            t_enq1 = since_entry();
            if (wdbg) {
                fprintf(stderr, "chol waves host: first enqueue at %.3f ms after entry, all waves enqueued at %.3f ms\n", t_enq0, t_enq1);
                LFB_CUDA(cudaStreamSynchronize(ms));
                LFB_CUDA(cudaStreamSynchronize(us));
                fprintf(stderr, "chol waves n=%lld (ms from start): wave end-col | upload done | main stream free | wave done\n", (long long)n);
                for (size_t w = 0; w < wend.size(); ++w) {
                    float t[3];
                    for (int q = 0; q < 3; ++q) cudaEventElapsedTime(&t[q], tev[0], tev[1 + 3 * w + q]);
                    fprintf(stderr, "  %6lld | %8.3f | %8.3f | %8.3f\n", (long long)wend[w], t[0], t[1], t[2]);
                }
                for (auto e : tev) cudaEventDestroy(e);
            }
            if (clean) triangular_zero<T>(*h, dA.get(), n, ld, /*keep_lower=*/1);
        }
    } catch (...) {
        if (h->upload_stream) cudaStreamSynchronize(h->upload_stream);
        for (auto e : wev) cudaEventDestroy(e);
        wev.clear();
        h->chol_panel_hook = nullptr;
        // copies queued by the hook may still be reading dA / the staging buffer: drain before the DevBufs are released
        if (h->copy_stream) cudaStreamSynchronize(h->copy_stream);
        cudaStreamSynchronize(h->stream);
        for (auto e : pev) cudaEventDestroy(e);
        throw;
    }
    h->chol_panel_hook = nullptr;
    for (auto e : wev) cudaEventDestroy(e);      // (destroying a recorded event is legal: its waits stay in place)
    LFB_CUDA(cudaMemcpyAsync(&info, dInfo.get(), sizeof(int64_t), cudaMemcpyDeviceToHost, h->stream));
    if (overlap) {
        // staged: scatter every block column into the caller's storage as soon as its D2H has landed -- the factorisation
        // is fully enqueued by now and keeps running while the host threads copy
        for (auto &pc : pieces) {
            LFB_CUDA(cudaEventSynchronize(pc.done));
            const int64_t below = n - pc.k0;
            if (lay == L_COL) host_scatter<T>(*h, pc.st, below, pc.nb, a + pc.k0 + pc.k0 * hld, hld);   // nb columns of `below` rows
            else host_scatter<T>(*h, pc.st, pc.nb, below, a + pc.k0 * hld + pc.k0, hld);               // `below` rows of nb entries
        }
        LFB_CUDA(cudaStreamSynchronize(h->copy_stream));
        LFB_CUDA(cudaStreamSynchronize(h->stream));
        for (auto e : pev) cudaEventDestroy(e);
    } else if (staged && !clean) {
        // no overlap requested: same staging, after the factorisation
        T *sp = stagebuf;
        std::unique_ptr<DevBuf<T>> tmp;
        if (lay != L_COL) {
            tmp.reset(new DevBuf<T>(*h, (size_t)hld * n));
            transpose<T>(*h, dA, n, n, ld, tmp->get(), hld);
        }
        for (int64_t b0 = 0; b0 < n; b0 += BAND) {
            const int64_t b1 = std::min(n, b0 + BAND);
            if (lay == L_COL) {
                LFB_CUDA(cudaMemcpy2DAsync(sp, (n - b0) * sizeof(T), dA.get() + b0 + b0 * ld, ld * sizeof(T), (n - b0) * sizeof(T), b1 - b0,
                                           cudaMemcpyDeviceToHost, h->stream));
                LFB_CUDA(cudaStreamSynchronize(h->stream));
                host_scatter<T>(*h, sp, n - b0, b1 - b0, a + b0 + b0 * hld, hld);
                sp += (size_t)(n - b0) * (b1 - b0);
            } else {
                LFB_CUDA(cudaMemcpy2DAsync(sp, b1 * sizeof(T), tmp->get() + b0 * hld, hld * sizeof(T), b1 * sizeof(T), b1 - b0,
                                           cudaMemcpyDeviceToHost, h->stream));
                LFB_CUDA(cudaStreamSynchronize(h->stream));
                host_scatter<T>(*h, sp, b1, b1 - b0, a + b0 * hld, hld);
                sp += (size_t)b1 * (b1 - b0);
            }
        }
    } else if (!tri || clean) {
        download<T>(*h, dA, ld, a, n, n, rs, cs);
    } else if (lay == L_COL) {
        for (int64_t c0 = 0; c0 < n; c0 += BAND) {
            const int64_t c1 = std::min(n, c0 + BAND);
            LFB_CUDA(cudaMemcpy2DAsync(a + c0 + c0 * hld, hld * sizeof(T), dA.get() + c0 + c0 * ld, ld * sizeof(T),
                                       (n - c0) * sizeof(T), c1 - c0, cudaMemcpyDeviceToHost, h->stream));
        }
        LFB_CUDA(cudaStreamSynchronize(h->stream));
    } else {
        DevBuf<T> tmp(*h, (size_t)hld * n);
        transpose<T>(*h, dA, n, n, ld, tmp.get(), hld);
        for (int64_t r0 = 0; r0 < n; r0 += BAND) {
            const int64_t r1 = std::min(n, r0 + BAND);
            LFB_CUDA(cudaMemcpy2DAsync(a + r0 * hld, hld * sizeof(T), tmp.get() + r0 * hld, hld * sizeof(T), r1 * sizeof(T), r1 - r0,
                                       cudaMemcpyDeviceToHost, h->stream));
        }
        LFB_CUDA(cudaStreamSynchronize(h->stream));
    }
    if (getenv("LFB_WAVE_DBG")) fprintf(stderr, "chol host: returning %.3f ms after entry\n", since_entry());
    if (info != 0) {
        if (fail_index) *fail_index = info - 1;
        h->err = "Matrix is not positive definite";
        return LFB_NOT_POSITIVE_DEFINITE;
    }
    LFB_API_END(h)
}

template <typename T>
int solve_triangular_host(lfb_handle *h, const T *a, int64_t a_rows, int64_t a_cols, int64_t a_rs, int64_t a_cs, T *b,
                          int64_t b_rows, int64_t b_cols, int64_t b_rs, int64_t b_cs, int uplo, const T *ext_diag) {
    if (a_rows != a_cols) return fail(h, LFB_NOT_SQUARE, "Matrix is not square");            // triangular.rs:102
    if (b_rows != a_rows) return fail(h, LFB_WRONG_ROWS, "Matrix has the wrong number of rows");  // :103-108
    if (uplo != LFB_UPPER && uplo != LFB_LOWER) return fail(h, LFB_INVALID_ARGUMENT, "bad uplo");
    const int64_t n = a_rows;
    if (n == 0 || b_cols == 0) return LFB_OK;
    LFB_API_BEGIN(h)
    const int64_t ld = round_up(n, 2);
    DevBuf<T> dA(*h, (size_t)ld * n), dB(*h, (size_t)ld * b_cols), dD(*h, n);
    upload<T>(*h, a, n, n, a_rs, a_cs, dA, ld);
    upload<T>(*h, b, n, b_cols, b_rs, b_cs, dB, ld);
    if (ext_diag) upload_vec<T>(*h, ext_diag, n, dD);
    trsm_left<T>(*h, uplo == LFB_LOWER, 0, n, b_cols, dA, ld, ext_diag ? dD.get() : (const T *)nullptr, dB, ld);
    download<T>(*h, dB, ld, b, n, b_cols, b_rs, b_cs);
    LFB_API_END(h)
}

template <typename T>
int triangular_inplace_host(lfb_handle *h, T *a, int64_t rows, int64_t cols, int64_t rs, int64_t cs, int uplo) {
    if (rows != cols) return fail(h, LFB_NOT_SQUARE, "Matrix is not square");
    if (rows == 0) return LFB_OK;
    LFB_API_BEGIN(h)
    const int64_t n = rows, ld = round_up(n, 2);
    DevBuf<T> dA(*h, (size_t)ld * n);
    upload<T>(*h, a, n, n, rs, cs, dA, ld);
    triangular_zero<T>(*h, dA, n, ld, uplo == LFB_LOWER);
    download<T>(*h, dA, ld, a, n, n, rs, cs);
    LFB_API_END(h)
}

// ---- fused solve drivers: one upload, every stage on the device, one download ----------
// True if any of the n device values is exactly zero (QRDecomp::is_invertible, qr.rs:194-197).
template <typename T>
bool any_zero_dev(lfb_handle &h, const T *d, int64_t n) {
    std::vector<T> v((size_t)std::max<int64_t>(n, 1));
    download_vec<T>(h, d, n, v.data());
    for (int64_t i = 0; i < n; ++i)
        if (v[i] == T(0)) return true;
    return false;
}

// d <- |d| (qr.rs:149,176 pass `|diag|` as the diagonal of R)
template <typename T>
__global__ void abs_kernel(T *d, int64_t n) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) d[i] = d[i] < T(0) ? -d[i] : d[i];
}
template <typename T>
void abs_dev(lfb_handle &h, T *d, int64_t n) {
    if (n <= 0) return;
    abs_kernel<T><<<(unsigned)cdiv(n, 256), 256, 0, h.stream>>>(d, n);
    LFB_LAUNCH_CHECK(h);
}

// qr.rs:207-229 LeastSquaresQrInto::least_squares_into: thin (rows >= cols): qr_into + solve_into (:124-152);
// wide: qr_into of the transpose + solve_tr_into (:156-181).  x: cols x bcols.
template <typename T>
int least_squares_host(lfb_handle *h, const T *a, int64_t rows, int64_t cols, int64_t rs, int64_t cs, const T *b, int64_t b_rows,
                       int64_t bcols, int64_t b_rs, int64_t b_cs, T *x, int64_t x_rs, int64_t x_cs) {
    if (b_rows != rows) return fail(h, LFB_WRONG_ROWS, "Matrix has the wrong number of rows");      // qr.rs:128-133 / :160-165
    if (rows == 0 || cols == 0) return LFB_OK;
    LFB_API_BEGIN(h)
    if (rows >= cols) {
        const int64_t ld = round_up(rows, 2);
        DevBuf<T> dA(*h, (size_t)ld * cols), dB(*h, (size_t)ld * std::max<int64_t>(bcols, 1)), dD(*h, cols);
        upload<T>(*h, a, rows, cols, rs, cs, dA, ld);
        upload<T>(*h, b, rows, bcols, b_rs, b_cs, dB, ld);
        qr_factor<T>(*h, dA, rows, cols, ld, dD);                                                    // :32-44
        if (any_zero_dev<T>(*h, dD, cols)) return fail(h, LFB_NON_INVERTIBLE, "Matrix is not invertible");   // :134-136
        if (bcols > 0) {
            qt_mul<T>(*h, dA, rows, cols, ld, dD, dB, bcols, ld);                                    // :139
            abs_dev<T>(*h, dD, cols);
            trsm_left<T>(*h, 0, 0, cols, bcols, dA, ld, dD, dB, ld);                                 // :145-150: R x = Q^T b, |diag| as the diagonal
            download<T>(*h, dB, ld, x, cols, bcols, x_rs, x_cs);
        }
    } else {
        // A^T = Q R (cols x rows, thin); solve R^T m = b, x = Q m
        const int64_t ldt = round_up(cols, 2), ldb = round_up(rows, 2);
        DevBuf<T> dA(*h, (size_t)ldb * cols), dT(*h, (size_t)ldt * rows), dB(*h, (size_t)ldb * std::max<int64_t>(bcols, 1)), dD(*h, rows);
        upload<T>(*h, a, rows, cols, rs, cs, dA, ldb);
        transpose<T>(*h, dA, rows, cols, ldb, dT, ldt);
        upload<T>(*h, b, rows, bcols, b_rs, b_cs, dB, ldb);
        qr_factor<T>(*h, dT, cols, rows, ldt, dD);
        if (any_zero_dev<T>(*h, dD, rows)) return fail(h, LFB_NON_INVERTIBLE, "Matrix is not invertible");   // :166-168
        if (bcols > 0) {
            DevBuf<T> dQ(*h, (size_t)ldt * rows), dX(*h, (size_t)ldt * bcols), dAbs(*h, rows);
            assemble_q<T>(*h, dT, cols, rows, ldt, 0, dD, dQ, ldt);                                  // :180 generate_q (needs the signed diag)
            LFB_CUDA(cudaMemcpyAsync(dAbs.get(), dD.get(), sizeof(T) * rows, cudaMemcpyDeviceToDevice, h->stream));
            abs_dev<T>(*h, dAbs, rows);
            trsm_left<T>(*h, 0, 1, rows, bcols, dT, ldt, dAbs, dB, ldb);                             // :172-177: R^T m = b
            gemm<T>(*h, 0, 0, cols, bcols, rows, T(1), dQ, ldt, dB, ldb, T(0), dX, ldt);             // :180: Q m
            download<T>(*h, dX, ldt, x, cols, bcols, x_rs, x_cs);
        }
    }
    LFB_API_END(h)
}

// QRDecomp::solve_into (qr.rs:124-152) on an existing decomposition: x (cols x bcols) = R^-1 (Q^T b)[..cols].
template <typename T>
int qr_solve_host(lfb_handle *h, const T *qr, int64_t rows, int64_t cols, int64_t rs, int64_t cs, const T *diag, const T *b,
                  int64_t b_rows, int64_t bcols, int64_t b_rs, int64_t b_cs, T *x, int64_t x_rs, int64_t x_cs) {
    if (rows < cols) return fail(h, LFB_NOT_THIN, "Expected matrix rows >= cols");
    if (b_rows != rows) return fail(h, LFB_WRONG_ROWS, "Matrix has the wrong number of rows");      // :128-133
    for (int64_t i = 0; i < cols; ++i)
        if (diag[i] == T(0)) return fail(h, LFB_NON_INVERTIBLE, "Matrix is not invertible");         // :134-136
    if (rows == 0 || cols == 0 || bcols == 0) return LFB_OK;
    LFB_API_BEGIN(h)
    const int64_t ld = round_up(rows, 2);
    DevBuf<T> dM(*h, (size_t)ld * cols), dB(*h, (size_t)ld * bcols), dD(*h, cols), dAbs(*h, cols);
    std::vector<T> ad((size_t)cols);
    for (int64_t i = 0; i < cols; ++i) ad[i] = std::fabs(diag[i]);
    upload<T>(*h, qr, rows, cols, rs, cs, dM, ld);
    upload<T>(*h, b, rows, bcols, b_rs, b_cs, dB, ld);
    upload_vec<T>(*h, diag, cols, dD);
    upload_vec<T>(*h, ad.data(), cols, dAbs);
    qt_mul<T>(*h, dM, rows, cols, ld, dD, dB, bcols, ld);
    trsm_left<T>(*h, 0, 0, cols, bcols, dM, ld, dAbs, dB, ld);
    download<T>(*h, dB, ld, x, cols, bcols, x_rs, x_cs);
    LFB_API_END(h)
}

// QRDecomp::solve_tr_into (qr.rs:156-181): x (rows x bcols) = Q m with R^T m = b.
template <typename T>
int qr_solve_tr_host(lfb_handle *h, const T *qr, int64_t rows, int64_t cols, int64_t rs, int64_t cs, const T *diag, const T *b,
                     int64_t b_rows, int64_t bcols, int64_t b_rs, int64_t b_cs, T *x, int64_t x_rs, int64_t x_cs) {
    if (rows < cols) return fail(h, LFB_NOT_THIN, "Expected matrix rows >= cols");
    if (b_rows != cols) return fail(h, LFB_WRONG_ROWS, "Matrix has the wrong number of rows");      // :160-165
    for (int64_t i = 0; i < cols; ++i)
        if (diag[i] == T(0)) return fail(h, LFB_NON_INVERTIBLE, "Matrix is not invertible");         // :166-168
    if (rows == 0 || cols == 0 || bcols == 0) return LFB_OK;
    LFB_API_BEGIN(h)
    const int64_t ld = round_up(rows, 4), ldb = round_up(cols, 2);
    DevBuf<T> dM(*h, (size_t)ld * cols), dQ(*h, (size_t)ld * cols), dB(*h, (size_t)ldb * bcols), dX(*h, (size_t)ld * bcols);
    DevBuf<T> dD(*h, cols), dAbs(*h, cols);
    std::vector<T> ad((size_t)cols);
    for (int64_t i = 0; i < cols; ++i) ad[i] = std::fabs(diag[i]);
    upload<T>(*h, qr, rows, cols, rs, cs, dM, ld);
    upload<T>(*h, b, cols, bcols, b_rs, b_cs, dB, ldb);
    upload_vec<T>(*h, diag, cols, dD);
    upload_vec<T>(*h, ad.data(), cols, dAbs);
    trsm_left<T>(*h, 0, 1, cols, bcols, dM, ld, dAbs, dB, ldb);                                      // :172-177  R^T m = b
    assemble_q<T>(*h, dM, rows, cols, ld, 0, dD, dQ, ld);                                            // :180 generate_q
    gemm<T>(*h, 0, 0, rows, bcols, cols, T(1), dQ, ld, dB, ldb, T(0), dX, ld);                       // :180 Q m
    download<T>(*h, dX, ld, x, rows, bcols, x_rs, x_cs);
    LFB_API_END(h)
}

// cholesky.rs:136-144 SolveCInplace::solvec_inplace (b in place; a receives its Cholesky factor in the lower
// triangle when write_factor != 0, as `cholesky_inplace_dirty` leaves it) and, with b == nullptr, cholesky.rs:178-182
// InverseCInplace::invc_inplace (the identity right-hand side is generated on the device; result in x).
template <typename T>
int solvec_host(lfb_handle *h, T *a, int64_t rows, int64_t cols, int64_t rs, int64_t cs, int write_factor, T *b, int64_t b_rows,
                int64_t bcols, int64_t b_rs, int64_t b_cs, int64_t *fail_index) {
    if (rows != cols) return fail(h, LFB_NOT_SQUARE, "Matrix is not square");
    if (b_rows != rows) return fail(h, LFB_WRONG_ROWS, "Matrix has the wrong number of rows");      // triangular.rs:103-108
    if (fail_index) *fail_index = -1;
    const int64_t n = rows;
    if (n == 0) return LFB_OK;
    int64_t info = 0;
    LFB_API_BEGIN(h)
    const int64_t ld = round_up(n, 2);
    DevBuf<T> dA(*h, (size_t)ld * n), dB(*h, (size_t)ld * std::max<int64_t>(bcols, 1));
    DevBuf<int64_t> dInfo(*h, 1);
    upload<T>(*h, a, n, n, rs, cs, dA, ld);
    if (b) upload<T>(*h, b, n, bcols, b_rs, b_cs, dB, ld);
    cholesky_lower<T>(*h, dA, n, ld, 0, dInfo);                                                      // :139 (dirty)
    LFB_CUDA(cudaMemcpyAsync(&info, dInfo.get(), sizeof(int64_t), cudaMemcpyDeviceToHost, h->stream));
    LFB_CUDA(cudaStreamSynchronize(h->stream));
    if (info != 0) {
        if (fail_index) *fail_index = info - 1;
        h->err = "Matrix is not positive definite";
        return LFB_NOT_POSITIVE_DEFINITE;
    }
    if (bcols > 0 && b) {
        trsm_left<T>(*h, 1, 0, n, bcols, dA, ld, (const T *)nullptr, dB, ld);                        // :140  L y = b
        trsm_left<T>(*h, 1, 1, n, bcols, dA, ld, (const T *)nullptr, dB, ld);                        // :141  L^T x = y
        download<T>(*h, dB, ld, b, n, bcols, b_rs, b_cs);
    }
    if (write_factor) download<T>(*h, dA, ld, a, n, n, rs, cs);
    LFB_API_END(h)
}

// lobpcg/algorithm.rs:81-97 orthonormalize: v (rows x cols) <- v L^-T, l (cols x cols) <- L = chol(v^T v), one round trip.
template <typename T>
int orthonormalize_host(lfb_handle *h, T *v, int64_t rows, int64_t cols, int64_t rs, int64_t cs, T *l, int64_t l_rs, int64_t l_cs,
                        int64_t *fail_index) {
    if (rows < 0 || cols < 0) return fail(h, LFB_INVALID_ARGUMENT, "negative dimension");
    if (fail_index) *fail_index = -1;
    if (cols == 0) return LFB_OK;
    int64_t info = 0;
    LFB_API_BEGIN(h)
    const int64_t ld = round_up(std::max<int64_t>(rows, 1), 2), ldl = round_up(cols, 2);
    DevBuf<T> dV(*h, (size_t)ld * cols), dL(*h, (size_t)ldl * cols);
    DevBuf<int64_t> dInfo(*h, 1);
    if (rows > 0) upload<T>(*h, v, rows, cols, rs, cs, dV, ld);
    orthonormalize<T>(*h, dV, rows, cols, ld, dL, ldl, dInfo);
    LFB_CUDA(cudaMemcpyAsync(&info, dInfo.get(), sizeof(int64_t), cudaMemcpyDeviceToHost, h->stream));
    LFB_CUDA(cudaStreamSynchronize(h->stream));
    if (info != 0) {
        if (fail_index) *fail_index = info - 1;
        h->err = "Matrix is not positive definite";
        return LFB_NOT_POSITIVE_DEFINITE;
    }
    if (rows > 0) download<T>(*h, dV, ld, v, rows, cols, rs, cs);
    if (l) download<T>(*h, dL, ldl, l, cols, cols, l_rs, l_cs);
    LFB_API_END(h)
}

// lobpcg/algorithm.rs:63-76 apply_constraints: v (n x k) -= y (L_yy^-1 (y^T v)), one round trip.
template <typename T>
int apply_constraints_host(lfb_handle *h, T *v, int64_t n, int64_t k, int64_t rs, int64_t cs, const T *lyy, int64_t m, int64_t l_rs,
                           int64_t l_cs, const T *y, int64_t y_rows, int64_t y_cols, int64_t y_rs, int64_t y_cs) {
    if (n < 0 || k < 0 || m < 0) return fail(h, LFB_INVALID_ARGUMENT, "negative dimension");
    if (y_rows != n || y_cols != m) return fail(h, LFB_INVALID_ARGUMENT, "y must be (rows of v) x (order of cholesky_yy)");
    if (n == 0 || k == 0 || m == 0) return LFB_OK;
    LFB_API_BEGIN(h)
    const int64_t ld = round_up(n, 2), ldl = round_up(m, 2);
    DevBuf<T> dV(*h, (size_t)ld * k), dY(*h, (size_t)ld * m), dL(*h, (size_t)ldl * m);
    upload<T>(*h, v, n, k, rs, cs, dV, ld);
    upload<T>(*h, y, n, m, y_rs, y_cs, dY, ld);
    upload<T>(*h, lyy, m, m, l_rs, l_cs, dL, ldl);
    apply_constraints<T>(*h, dV, n, k, ld, dL, m, ldl, dY, ld);
    download<T>(*h, dV, ld, v, n, k, rs, cs);
    LFB_API_END(h)
}

// lobpcg/algorithm.rs:16-44 on host views: one upload, both eigendecompositions and the products between them on the device.
template <typename T>
int sorted_eig_host(lfb_handle *h, const T *a, int64_t k, int64_t a_rs, int64_t a_cs, const T *b, int64_t b_rs, int64_t b_cs, int64_t size,
                    int order, T *vals, T *vecs, int64_t v_rs, int64_t v_cs) {
    if (k < 0 || size < 0 || order < 0 || order > 2) return fail(h, LFB_INVALID_ARGUMENT, "bad arguments");
    if (k == 0) return LFB_OK;
    if (!vals || !vecs) return fail(h, LFB_INVALID_ARGUMENT, "null output");
    LFB_API_BEGIN(h)
    const int64_t ld = round_up(k, 2);
    const int64_t nout = order == 0 ? k : std::min(size, k);
    DevBuf<T> dA(*h, (size_t)ld * k), dB(*h, b ? (size_t)ld * k : 1), dV(*h, (size_t)ld * std::max<int64_t>(nout, 1));
    upload<T>(*h, a, k, k, a_rs, a_cs, dA, ld);
    if (b) upload<T>(*h, b, k, k, b_rs, b_cs, dB, ld);
    if (!sorted_eig_dev<T>(*h, dA, ld, b ? dB.get() : nullptr, ld, k, size, order, vals, dV, ld))
        return fail(h, LFB_INVALID_ARGUMENT, "NaN values in array");
    if (nout > 0) download<T>(*h, dV, ld, vecs, k, nout, v_rs, v_cs);
    LFB_API_END(h)
}

template <typename T>
int invc_host(lfb_handle *h, const T *a, int64_t rows, int64_t cols, int64_t rs, int64_t cs, T *inv, int64_t i_rs, int64_t i_cs,
              int64_t *fail_index) {
    if (rows != cols) return fail(h, LFB_NOT_SQUARE, "Matrix is not square");
    if (fail_index) *fail_index = -1;
    const int64_t n = rows;
    if (n == 0) return LFB_OK;
    int64_t info = 0;
    LFB_API_BEGIN(h)
    const int64_t ld = round_up(n, 2);
    DevBuf<T> dA(*h, (size_t)ld * n), dB(*h, (size_t)ld * n);
    DevBuf<int64_t> dInfo(*h, 1);
    upload<T>(*h, a, n, n, rs, cs, dA, ld);
    fill<T>(*h, dB, n, n, ld, T(0), T(1));                                                           // :180 Array2::eye
    cholesky_lower<T>(*h, dA, n, ld, 0, dInfo);
    LFB_CUDA(cudaMemcpyAsync(&info, dInfo.get(), sizeof(int64_t), cudaMemcpyDeviceToHost, h->stream));
    LFB_CUDA(cudaStreamSynchronize(h->stream));
    if (info != 0) {
        if (fail_index) *fail_index = info - 1;
        h->err = "Matrix is not positive definite";
        return LFB_NOT_POSITIVE_DEFINITE;
    }
    trsm_left<T>(*h, 1, 0, n, n, dA, ld, (const T *)nullptr, dB, ld);
    trsm_left<T>(*h, 1, 1, n, n, dA, ld, (const T *)nullptr, dB, ld);
    download<T>(*h, dB, ld, inv, n, n, i_rs, i_cs);
    LFB_API_END(h)
}

template <typename T>
int sym_tridiagonal_host(lfb_handle *h, T *a, int64_t rows, int64_t cols, int64_t rs, int64_t cs, T *off) {
    if (rows != cols) return fail(h, LFB_NOT_SQUARE, "Matrix is not square");                 // tridiagonal.rs:32
    if (rows < 1) return fail(h, LFB_EMPTY_MATRIX, "Matrix is empty");                        // :33-35
    const int64_t n = rows;
    if (n == 1) return LFB_OK;
    LFB_API_BEGIN(h)
    const int64_t ld = round_up(n, 2);
    DevBuf<T> dA(*h, (size_t)ld * n), dOff(*h, n);
    upload<T>(*h, a, n, n, rs, cs, dA, ld);
    sym_tridiagonal<T>(*h, dA, n, ld, dOff);
    download<T>(*h, dA, ld, a, n, n, rs, cs);
    download_vec<T>(*h, dOff, n - 1, off);
    LFB_API_END(h)
}

template <typename T>
int eigh_host(lfb_handle *h, const T *a, int64_t rows, int64_t cols, int64_t rs, int64_t cs, T *vals, T *vecs, int64_t vrs,
              int64_t vcs) {
    if (rows != cols) return fail(h, LFB_NOT_SQUARE, "Matrix is not square");                 // eigh.rs:15 (check_square)
    if (rows > 0 && !vals) return fail(h, LFB_INVALID_ARGUMENT, "vals is null");
    const int64_t n = rows;
    if (n == 0) return LFB_OK;                                                                // :16-25
    LFB_API_BEGIN(h)
    const int64_t ld = round_up(n, 2);
    DevBuf<T> dA(*h, (size_t)ld * n), dQ(*h, vecs ? (size_t)ld * n : 1);
    upload<T>(*h, a, n, n, rs, cs, dA, ld);
    symmetric_eig<T>(*h, dA, n, ld, vals, vecs ? dQ.get() : nullptr, ld);
    if (vecs) download<T>(*h, dQ, ld, vecs, n, n, vrs, vcs);
    LFB_API_END(h)
}

template <typename T>
int svd_host(lfb_handle *h, const T *a, int64_t rows, int64_t cols, int64_t rs, int64_t cs, T *sv, T *u, int64_t urs, int64_t ucs,
             T *vt, int64_t vrs, int64_t vcs) {
    if (rows <= 0 || cols <= 0) return fail(h, LFB_EMPTY_MATRIX, "Matrix is empty");           // svd.rs:23-25
    if (!sv) return fail(h, LFB_INVALID_ARGUMENT, "sigma is null");
    LFB_API_BEGIN(h)
    const int64_t dim = std::min(rows, cols);
    const int64_t ld = round_up(rows, 2), ldv = round_up(cols, 2);
    DevBuf<T> dA(*h, (size_t)ld * cols), dU(*h, u ? (size_t)ld * dim : 1), dV(*h, vt ? (size_t)ldv * dim : 1);
    upload<T>(*h, a, rows, cols, rs, cs, dA, ld);
    svd_dev<T>(*h, dA, rows, cols, ld, sv, u ? dU.get() : nullptr, ld, vt ? dV.get() : nullptr, ldv);
    if (u) download<T>(*h, dU, ld, u, rows, dim, urs, ucs);
    if (vt) download<T>(*h, dV, ldv, vt, cols, dim, vcs, vrs);      // the device holds V = Vt^T: swap the strides
    LFB_API_END(h)
}

template <typename T>
int bidiagonal_host(lfb_handle *h, T *a, int64_t rows, int64_t cols, int64_t rs, int64_t cs, T *d, T *e) {
    const int64_t md = std::min(rows, cols);
    if (md <= 0) return fail(h, LFB_EMPTY_MATRIX, "Matrix is empty");                         // bidiagonal.rs:30-32
    LFB_API_BEGIN(h)
    const int64_t ld = round_up(rows, 2);
    DevBuf<T> dA(*h, (size_t)ld * cols), dD(*h, md), dE(*h, md);
    upload<T>(*h, a, rows, cols, rs, cs, dA, ld);
    bidiagonal<T>(*h, dA, rows, cols, ld, dD, dE);
    download<T>(*h, dA, ld, a, rows, cols, rs, cs);
    download_vec<T>(*h, dD, md, d);
    if (md > 1) download_vec<T>(*h, dE, md - 1, e);
    LFB_API_END(h)
}

template <typename T>
int qr_batched_host(lfb_handle *h, T *a, int64_t batch, int64_t m, int64_t n, T *diag) {
    if (m < n) return fail(h, LFB_NOT_THIN, "Expected matrix rows >= cols");
    if (m > 32 || n > 32) return fail(h, LFB_UNSUPPORTED, "batched QR supports m, n <= 32");
    if (batch <= 0 || n == 0) return LFB_OK;
    LFB_API_BEGIN(h)
    DevBuf<T> dA(*h, (size_t)batch * m * n), dD(*h, (size_t)batch * n);
    LFB_CUDA(cudaMemcpyAsync(dA.get(), a, sizeof(T) * batch * m * n, cudaMemcpyHostToDevice, h->stream));
    qr_batched<T>(*h, dA, batch, m, n, dD);
    LFB_CUDA(cudaMemcpyAsync(a, dA.get(), sizeof(T) * batch * m * n, cudaMemcpyDeviceToHost, h->stream));
    LFB_CUDA(cudaMemcpyAsync(diag, dD.get(), sizeof(T) * batch * n, cudaMemcpyDeviceToHost, h->stream));
    LFB_CUDA(cudaStreamSynchronize(h->stream));
    LFB_API_END(h)
}

// cholesky.rs:51-83 over a packed batch.  Returns LFB_NOT_POSITIVE_DEFINITE if any matrix fails; *fail_matrix /
// *fail_index then name the first failing matrix (batch order) and its failing row.
template <typename T>
int cholesky_batched_host(lfb_handle *h, T *a, int64_t batch, int64_t n, int clean, int64_t *fail_matrix, int64_t *fail_index) {
    if (fail_matrix) *fail_matrix = -1;
    if (fail_index) *fail_index = -1;
    if (n > 32) return fail(h, LFB_UNSUPPORTED, "batched Cholesky supports n <= 32");
    if (batch <= 0 || n <= 0) return LFB_OK;
    LFB_API_BEGIN(h)
    DevBuf<T> dA(*h, (size_t)batch * n * n);
    DevBuf<int> dF(*h, (size_t)batch);
    std::vector<int> hf((size_t)batch);
    LFB_CUDA(cudaMemcpyAsync(dA.get(), a, sizeof(T) * batch * n * n, cudaMemcpyHostToDevice, h->stream));
    cholesky_batched<T>(*h, dA, batch, n, clean, dF);
    LFB_CUDA(cudaMemcpyAsync(a, dA.get(), sizeof(T) * batch * n * n, cudaMemcpyDeviceToHost, h->stream));
    LFB_CUDA(cudaMemcpyAsync(hf.data(), dF.get(), sizeof(int) * batch, cudaMemcpyDeviceToHost, h->stream));
    LFB_CUDA(cudaStreamSynchronize(h->stream));
    for (int64_t b = 0; b < batch; ++b)
        if (hf[b] >= 0) {
            if (fail_matrix) *fail_matrix = b;
            if (fail_index) *fail_index = hf[b];
            h->err = "Matrix is not positive definite";
            return LFB_NOT_POSITIVE_DEFINITE;
        }
    LFB_API_END(h)
}

}  // namespace

extern "C" {

const char *lfb_version(void) { return "linfa_b200 0.1 (sm_100a; DMMA FP64 GEMM core; compact-WY QR; recursive Cholesky)"; }

int lfb_create(lfb_handle **out, int device) {
    if (!out) return LFB_INVALID_ARGUMENT;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0 || device < 0 || device >= count) {
        cudaGetLastError();
        return LFB_ERR_CUDA;  // no CPU fallback by design
    }
    lfb_handle *h = new lfb_handle();
    h->device = device;
    try {
        LFB_CUDA(cudaSetDevice(device));
        LFB_CUDA(cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking));
        h->stream = h->own_stream;
        int lo = 0, hi = 0;
        LFB_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        LFB_CUDA(cudaStreamCreateWithPriority(&h->aux_stream, cudaStreamNonBlocking, hi));
        LFB_CUDA(cudaStreamCreateWithPriority(&h->aux2_stream, cudaStreamNonBlocking, hi));
        for (auto &e : h->ev) LFB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        cudaDeviceProp prop;
        LFB_CUDA(cudaGetDeviceProperties(&prop, device));
        h->sm_count = prop.multiProcessorCount;
        h->smem_optin = prop.sharedMemPerBlockOptin;
    } catch (const std::exception &) {
        cudaGetLastError();
        lfb_destroy(h);   // releases whatever was created before the failure
        return LFB_ERR_CUDA;
    }
    *out = h;
    return LFB_OK;
}

int lfb_destroy(lfb_handle *h) {
    if (!h) return LFB_OK;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    h->drop_graphs();
    for (auto &e : h->ev_graph) if (e) cudaEventDestroy(e);
    for (auto *s : h->subs) lfb_destroy(s);
    h->subs.clear();
    for (auto &b : h->blocks) cudaFree(b.p);
    for (auto &e : h->prof_ev) if (e) cudaEventDestroy(e);
    if (h->panel_dbg) cudaFree(h->panel_dbg);
    if (h->pinned) cudaFreeHost(h->pinned);
    delete h->host_pool;
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    if (h->upload_stream) cudaStreamDestroy(h->upload_stream);
    if (h->hr_scratch) { cudaFree(h->hr_scratch); cudaEventDestroy(h->hr_ev[0]); cudaEventDestroy(h->hr_ev[1]); }
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    if (h->aux_stream) cudaStreamDestroy(h->aux_stream);
    if (h->aux2_stream) cudaStreamDestroy(h->aux2_stream);
    for (auto &e : h->ev) if (e) cudaEventDestroy(e);
    delete h;
    return LFB_OK;
}

}  // extern "C"

void lfb::lfb_ensure_subs(lfb_handle &h, int n) {
    while ((int)h.subs.size() < n) {
        lfb_handle *s = nullptr;
        if (lfb_create(&s, h.device) != LFB_OK || !s) throw lfb::CudaError(LFB_ERR_CUDA, "could not create a worker handle");
        s->is_sub = true;
        s->opt = h.opt;
        h.subs.push_back(s);
    }
    for (auto *s : h.subs) s->opt = h.opt;
}

extern "C" {

const char *lfb_last_error(lfb_handle *h) { return h ? h->err.c_str() : "null handle"; }

// The workspace pool is not stream-aware (a DevBuf released at the end of an async `_dev` call may still be in use by
// kernels queued on the stream), which is safe as long as every later user is queued on the SAME stream.  Switching
// streams therefore drains the outgoing one first.
static int switch_stream(lfb_handle *h, cudaStream_t s) {
    if (!h) return LFB_INVALID_ARGUMENT;
    if (s == h->stream) return LFB_OK;
    LFB_API_BEGIN(h)
    LFB_CUDA(cudaStreamSynchronize(h->stream));
    h->untag_blocks();      // every call joins its side streams into h->stream before it returns, and that stream is idle now
    h->stream = s;
    LFB_API_END(h)
}

int lfb_set_stream(lfb_handle *h, void *s) { return switch_stream(h, (cudaStream_t)s); }   // verbatim: NULL is the legacy default stream

int lfb_use_own_stream(lfb_handle *h) { return h ? switch_stream(h, h->own_stream) : LFB_INVALID_ARGUMENT; }

int lfb_synchronize(lfb_handle *h) {
    LFB_API_BEGIN(h)
    LFB_CUDA(cudaStreamSynchronize(h->stream));
    h->untag_blocks();
    LFB_API_END(h)
}

int64_t lfb_launch_count(lfb_handle *h) { return h ? h->launches : 0; }

int lfb_set_option(lfb_handle *h, const char *key, int64_t value) {
    if (!h || !key) return LFB_INVALID_ARGUMENT;
    lfb::Options &o = h->opt;
    // name, field, smallest and largest accepted value (anything else: LFB_INVALID_ARGUMENT, option unchanged)
    struct Opt { const char *name; int64_t *field; int64_t lo, hi; };
    const int64_t BIG = int64_t(1) << 40;
    const Opt table[] = {
        {"qr_nb", &o.qr_nb, 32, 1024}, {"qr_nb_f32", &o.qr_nb_f32, 32, 1024}, {"qr_vt", &o.qr_vt, 0, 1}, {"qr_sub", &o.qr_sub, 1, 32}, {"chol_base", &o.chol_base, 1, 64},
        {"chol_nb", &o.chol_nb, 64, 8192}, {"chol_tn", &o.chol_tn, 0, 1}, {"chol_nb_tail", &o.chol_nb_tail, 64, 8192}, {"chol_tail_rows", &o.chol_tail_rows, 0, BIG}, {"chol_split_panel", &o.chol_split_panel, 0, 1}, {"chol_trace", &o.chol_trace, 0, 1}, {"chol_potf2_rl", &o.chol_potf2_rl, 0, 1}, {"gemm_tma", &o.gemm_tma, 0, 1}, {"gemm_splitk", &o.gemm_splitk, 0, 1}, {"gemm_deterministic", &o.gemm_deterministic, 0, 1}, {"tsqr_cholqr_cond", &o.tsqr_cholqr_cond, 0, 1 << 20},
        {"gemm_v2", &o.gemm_v2, 0, 1}, {"sgemm_tc", &o.sgemm_tc, 0, 2}, {"gemm_split_waves", &o.gemm_split_waves, 1, 64}, {"panel_cluster", &o.panel_cluster, 0, 2},
        {"panel_cluster_max", &o.panel_cluster_max, 1, 16}, {"lookahead", &o.lookahead, 0, 1}, {"tsqr_chunk", &o.tsqr_chunk, 64, BIG},
        {"batched_quad", &o.batched_quad, 0, 4}, {"gemm_tma2", &o.gemm_tma2, 0, 2}, {"qr_trace", &o.qr_trace, 0, 1}, {"hr_split", &o.hr_split, 0, 1}, {"qr_fold_t", &o.qr_fold_t, 0, 2}, {"qr_overlap_d2h", &o.qr_overlap_d2h, 0, 1}, {"chol_waves", &o.chol_waves, 1, 4}, {"qr_panel_cholqr", &o.qr_panel_cholqr, 0, 2}, {"cholqr_fused", &o.cholqr_fused, 0, 1}, {"gemm_tma2_maxk", &o.gemm_tma2_maxk, 16, 1 << 30}, {"tsqr_streams", &o.tsqr_streams, 1, 64}, {"tsqr_graph", &o.tsqr_graph, 0, 1}, {"hr_lu_blocked", &o.hr_lu_blocked, 0, 1},
        {"qr_tsqr_auto", &o.qr_tsqr_auto, 0, 1}, {"trd_fused", &o.trd_fused, 0, 1}, {"chol_overlap_d2h", &o.chol_overlap_d2h, 0, 1}, {"host_staging", &o.host_staging, 0, 1},
        {"rot_staged", &o.rot_staged, 0, 1}, {"eigh_stable_2x2", &o.eigh_stable_2x2, 0, 1}, {"rot_serial", &o.rot_serial, 0, 1},
        {"fast_hypot", &o.fast_hypot, 0, 1}, {"bd_blocked", &o.bd_blocked, 0, 1}, {"trd_profile", &o.trd_profile, 0, 1},
        {"trd_symv_async", &o.trd_symv_async, 0, 1},
    };
    for (const Opt &e : table)
        if (std::strcmp(e.name, key) == 0) {
            if (value < e.lo || value > e.hi) return fail(h, LFB_INVALID_ARGUMENT, "option value out of range");
            *e.field = value;
            return LFB_OK;
        }
    return fail(h, LFB_INVALID_ARGUMENT, "unknown option");
}

#define DEF2(name, T, sfx, args, call) \
    int name##sfx args { return call; }

int lfb_qr_f64(lfb_handle *h, double *a, int64_t r, int64_t c, int64_t rs, int64_t cs, double *diag) { return qr_host<double>(h, a, r, c, rs, cs, diag); }
int lfb_qr_f32(lfb_handle *h, float *a, int64_t r, int64_t c, int64_t rs, int64_t cs, float *diag) { return qr_host<float>(h, a, r, c, rs, cs, diag); }

int lfb_qr_tsqr_f64(lfb_handle *h, double *a, int64_t r, int64_t c, int64_t rs, int64_t cs, double *diag) { return qr_host<double>(h, a, r, c, rs, cs, diag, true); }
int lfb_qr_tsqr_f32(lfb_handle *h, float *a, int64_t r, int64_t c, int64_t rs, int64_t cs, float *diag) { return qr_host<float>(h, a, r, c, rs, cs, diag, true); }
int lfb_assemble_q_f64(lfb_handle *h, const double *m, int64_t r, int64_t c, int64_t rs, int64_t cs, int64_t shift, const double *signs, double *q, int64_t qrs, int64_t qcs) {
    return assemble_q_host<double>(h, m, r, c, rs, cs, shift, signs, q, qrs, qcs);
}
int lfb_assemble_q_f32(lfb_handle *h, const float *m, int64_t r, int64_t c, int64_t rs, int64_t cs, int64_t shift, const float *signs, float *q, int64_t qrs, int64_t qcs) {
    return assemble_q_host<float>(h, m, r, c, rs, cs, shift, signs, q, qrs, qcs);
}
int lfb_qt_mul_f64(lfb_handle *h, const double *qr, int64_t r, int64_t c, int64_t rs, int64_t cs, const double *diag, double *b, int64_t bc, int64_t brs, int64_t bcs) {
    return qt_mul_host<double>(h, qr, r, c, rs, cs, diag, b, bc, brs, bcs);
}
int lfb_qt_mul_f32(lfb_handle *h, const float *qr, int64_t r, int64_t c, int64_t rs, int64_t cs, const float *diag, float *b, int64_t bc, int64_t brs, int64_t bcs) {
    return qt_mul_host<float>(h, qr, r, c, rs, cs, diag, b, bc, brs, bcs);
}
int lfb_cholesky_f64(lfb_handle *h, double *a, int64_t r, int64_t c, int64_t rs, int64_t cs, int clean, int64_t *fi) { return cholesky_host<double>(h, a, r, c, rs, cs, clean, fi); }
int lfb_cholesky_f32(lfb_handle *h, float *a, int64_t r, int64_t c, int64_t rs, int64_t cs, int clean, int64_t *fi) { return cholesky_host<float>(h, a, r, c, rs, cs, clean, fi); }
int lfb_solve_triangular_f64(lfb_handle *h, const double *a, int64_t ar, int64_t ac, int64_t ars, int64_t acs, double *b, int64_t br, int64_t bc, int64_t brs, int64_t bcs, int uplo, const double *ed) {
    return solve_triangular_host<double>(h, a, ar, ac, ars, acs, b, br, bc, brs, bcs, uplo, ed);
}
int lfb_solve_triangular_f32(lfb_handle *h, const float *a, int64_t ar, int64_t ac, int64_t ars, int64_t acs, float *b, int64_t br, int64_t bc, int64_t brs, int64_t bcs, int uplo, const float *ed) {
    return solve_triangular_host<float>(h, a, ar, ac, ars, acs, b, br, bc, brs, bcs, uplo, ed);
}
int lfb_triangular_inplace_f64(lfb_handle *h, double *a, int64_t r, int64_t c, int64_t rs, int64_t cs, int uplo) { return triangular_inplace_host<double>(h, a, r, c, rs, cs, uplo); }
int lfb_triangular_inplace_f32(lfb_handle *h, float *a, int64_t r, int64_t c, int64_t rs, int64_t cs, int uplo) { return triangular_inplace_host<float>(h, a, r, c, rs, cs, uplo); }
int lfb_sym_tridiagonal_f64(lfb_handle *h, double *a, int64_t r, int64_t c, int64_t rs, int64_t cs, double *off) { return sym_tridiagonal_host<double>(h, a, r, c, rs, cs, off); }
int lfb_sym_tridiagonal_f32(lfb_handle *h, float *a, int64_t r, int64_t c, int64_t rs, int64_t cs, float *off) { return sym_tridiagonal_host<float>(h, a, r, c, rs, cs, off); }
int lfb_eigh_f64(lfb_handle *h, const double *a, int64_t r, int64_t c, int64_t rs, int64_t cs, double *vals, double *vecs, int64_t vrs, int64_t vcs) { return eigh_host<double>(h, a, r, c, rs, cs, vals, vecs, vrs, vcs); }
int lfb_eigh_f32(lfb_handle *h, const float *a, int64_t r, int64_t c, int64_t rs, int64_t cs, float *vals, float *vecs, int64_t vrs, int64_t vcs) { return eigh_host<float>(h, a, r, c, rs, cs, vals, vecs, vrs, vcs); }
#define LFB_SOLVE_ENTRIES(SFX, T)                                                                                          \
    int lfb_least_squares_##SFX(lfb_handle *h, const T *a, int64_t r, int64_t c, int64_t rs, int64_t cs, const T *b, int64_t br, \
                                int64_t bc, int64_t brs, int64_t bcs, T *x, int64_t xrs, int64_t xcs) {                    \
        return least_squares_host<T>(h, a, r, c, rs, cs, b, br, bc, brs, bcs, x, xrs, xcs);                                \
    }                                                                                                                      \
    int lfb_qr_solve_##SFX(lfb_handle *h, const T *qr, int64_t r, int64_t c, int64_t rs, int64_t cs, const T *diag, const T *b, \
                           int64_t br, int64_t bc, int64_t brs, int64_t bcs, T *x, int64_t xrs, int64_t xcs) {             \
        return qr_solve_host<T>(h, qr, r, c, rs, cs, diag, b, br, bc, brs, bcs, x, xrs, xcs);                              \
    }                                                                                                                      \
    int lfb_qr_solve_tr_##SFX(lfb_handle *h, const T *qr, int64_t r, int64_t c, int64_t rs, int64_t cs, const T *diag, const T *b, \
                              int64_t br, int64_t bc, int64_t brs, int64_t bcs, T *x, int64_t xrs, int64_t xcs) {          \
        return qr_solve_tr_host<T>(h, qr, r, c, rs, cs, diag, b, br, bc, brs, bcs, x, xrs, xcs);                           \
    }                                                                                                                      \
    int lfb_solvec_##SFX(lfb_handle *h, T *a, int64_t r, int64_t c, int64_t rs, int64_t cs, int write_factor, T *b, int64_t br, \
                         int64_t bc, int64_t brs, int64_t bcs, int64_t *fail_index) {                                      \
        return solvec_host<T>(h, a, r, c, rs, cs, write_factor, b, br, bc, brs, bcs, fail_index);                          \
    }                                                                                                                      \
    int lfb_invc_##SFX(lfb_handle *h, const T *a, int64_t r, int64_t c, int64_t rs, int64_t cs, T *inv, int64_t irs,       \
                       int64_t ics, int64_t *fail_index) {                                                                 \
        return invc_host<T>(h, a, r, c, rs, cs, inv, irs, ics, fail_index);                                                \
    }
LFB_SOLVE_ENTRIES(f64, double)
LFB_SOLVE_ENTRIES(f32, float)
#undef LFB_SOLVE_ENTRIES
#define LFB_LOBPCG_ENTRIES(SFX, T)                                                                                         \
    int lfb_orthonormalize_##SFX(lfb_handle *h, T *v, int64_t r, int64_t c, int64_t rs, int64_t cs, T *l, int64_t lrs,     \
                                 int64_t lcs, int64_t *fail_index) {                                                       \
        return orthonormalize_host<T>(h, v, r, c, rs, cs, l, lrs, lcs, fail_index);                                        \
    }                                                                                                                      \
    int lfb_apply_constraints_##SFX(lfb_handle *h, T *v, int64_t n, int64_t k, int64_t rs, int64_t cs, const T *lyy,       \
                                    int64_t m, int64_t lrs, int64_t lcs, const T *y, int64_t yr, int64_t yc, int64_t yrs,  \
                                    int64_t ycs) {                                                                         \
        return apply_constraints_host<T>(h, v, n, k, rs, cs, lyy, m, lrs, lcs, y, yr, yc, yrs, ycs);                       \
    }
LFB_LOBPCG_ENTRIES(f64, double)
LFB_LOBPCG_ENTRIES(f32, float)
#undef LFB_LOBPCG_ENTRIES
int lfb_sorted_eig_f64(lfb_handle *h, const double *a, int64_t k, int64_t ars, int64_t acs, const double *b, int64_t brs, int64_t bcs, int64_t size,
                       int order, double *vals, double *vecs, int64_t vrs, int64_t vcs) {
    return sorted_eig_host<double>(h, a, k, ars, acs, b, brs, bcs, size, order, vals, vecs, vrs, vcs);
}
int lfb_sorted_eig_f32(lfb_handle *h, const float *a, int64_t k, int64_t ars, int64_t acs, const float *b, int64_t brs, int64_t bcs, int64_t size,
                       int order, float *vals, float *vecs, int64_t vrs, int64_t vcs) {
    return sorted_eig_host<float>(h, a, k, ars, acs, b, brs, bcs, size, order, vals, vecs, vrs, vcs);
}
int lfb_sorted_eig_dev_f64(lfb_handle *h, double *d_a, int64_t lda, double *d_b, int64_t ldb, int64_t k, int64_t size, int order,
                           double *vals_host, double *d_vecs, int64_t ldv) {
    if (k < 0 || size < 0 || order < 0 || order > 2 || (k > 0 && (!vals_host || !d_vecs || !d_a))) return fail(h, LFB_INVALID_ARGUMENT, "bad arguments");
    LFB_API_BEGIN(h)
    if (!sorted_eig_dev<double>(*h, d_a, lda, d_b, ldb, k, size, order, vals_host, d_vecs, ldv)) return fail(h, LFB_INVALID_ARGUMENT, "NaN values in array");
    LFB_API_END(h)
}
int lfb_orthonormalize_dev_f64(lfb_handle *h, double *d_v, int64_t rows, int64_t cols, int64_t ld, double *d_l, int64_t ldl,
                               int64_t *d_info) {
    if (rows < 0 || cols < 0 || !d_info) return fail(h, LFB_INVALID_ARGUMENT, "bad arguments");
    LFB_API_BEGIN(h)
    orthonormalize<double>(*h, d_v, rows, cols, ld, d_l, ldl, d_info);
    LFB_API_END(h)
}
int lfb_apply_constraints_dev_f64(lfb_handle *h, double *d_v, int64_t n, int64_t k, int64_t ldv, const double *d_lyy, int64_t m,
                                  int64_t ldl, const double *d_y, int64_t ldy) {
    if (n < 0 || k < 0 || m < 0) return fail(h, LFB_INVALID_ARGUMENT, "negative dimension");
    LFB_API_BEGIN(h)
    apply_constraints<double>(*h, d_v, n, k, ldv, d_lyy, m, ldl, d_y, ldy);
    LFB_API_END(h)
}
int lfb_svd_f64(lfb_handle *h, const double *a, int64_t r, int64_t c, int64_t rs, int64_t cs, double *sv, double *u, int64_t urs, int64_t ucs, double *vt, int64_t vrs, int64_t vcs) { return svd_host<double>(h, a, r, c, rs, cs, sv, u, urs, ucs, vt, vrs, vcs); }
int lfb_svd_f32(lfb_handle *h, const float *a, int64_t r, int64_t c, int64_t rs, int64_t cs, float *sv, float *u, int64_t urs, int64_t ucs, float *vt, int64_t vrs, int64_t vcs) { return svd_host<float>(h, a, r, c, rs, cs, sv, u, urs, ucs, vt, vrs, vcs); }
int lfb_bidiagonal_f64(lfb_handle *h, double *a, int64_t r, int64_t c, int64_t rs, int64_t cs, double *d, double *e) { return bidiagonal_host<double>(h, a, r, c, rs, cs, d, e); }
int lfb_bidiagonal_f32(lfb_handle *h, float *a, int64_t r, int64_t c, int64_t rs, int64_t cs, float *d, float *e) { return bidiagonal_host<float>(h, a, r, c, rs, cs, d, e); }
int lfb_qr_batched_f32(lfb_handle *h, float *a, int64_t batch, int64_t m, int64_t n, float *diag) { return qr_batched_host<float>(h, a, batch, m, n, diag); }
int lfb_qr_batched_f64(lfb_handle *h, double *a, int64_t batch, int64_t m, int64_t n, double *diag) { return qr_batched_host<double>(h, a, batch, m, n, diag); }
int lfb_cholesky_batched_f32(lfb_handle *h, float *a, int64_t batch, int64_t n, int clean, int64_t *fm, int64_t *fi) { return cholesky_batched_host<float>(h, a, batch, n, clean, fm, fi); }
int lfb_cholesky_batched_f64(lfb_handle *h, double *a, int64_t batch, int64_t n, int clean, int64_t *fm, int64_t *fi) { return cholesky_batched_host<double>(h, a, batch, n, clean, fm, fi); }
int lfb_cholesky_batched_dev_f32(lfb_handle *h, float *d_a, int64_t batch, int64_t n, int clean, int *d_fail) {
    if (n > 32) return fail(h, LFB_UNSUPPORTED, "batched Cholesky supports n <= 32");
    LFB_API_BEGIN(h)
    cholesky_batched<float>(*h, d_a, batch, n, clean, d_fail);
    LFB_API_END(h)
}

// ---- device-resident variants (async on the handle's stream) ----
int lfb_qr_dev_f64(lfb_handle *h, double *d_a, int64_t rows, int64_t cols, int64_t ld, double *d_diag) {
    if (rows < cols) return fail(h, LFB_NOT_THIN, "Expected matrix rows >= cols");
    LFB_API_BEGIN(h)
    qr_factor<double>(*h, d_a, rows, cols, ld, d_diag);
    LFB_API_END(h)
}
int lfb_qr_dev_f32(lfb_handle *h, float *d_a, int64_t rows, int64_t cols, int64_t ld, float *d_diag) {
    if (rows < cols) return fail(h, LFB_NOT_THIN, "Expected matrix rows >= cols");
    LFB_API_BEGIN(h)
    qr_factor<float>(*h, d_a, rows, cols, ld, d_diag);
    LFB_API_END(h)
}
int lfb_cholesky_dev_f64(lfb_handle *h, double *d_a, int64_t n, int64_t ld, int clean, int64_t *d_info) {
    LFB_API_BEGIN(h)
    cholesky_lower<double>(*h, d_a, n, ld, clean, d_info);
    LFB_API_END(h)
}
int lfb_cholesky_dev_f32(lfb_handle *h, float *d_a, int64_t n, int64_t ld, int clean, int64_t *d_info) {
    LFB_API_BEGIN(h)
    cholesky_lower<float>(*h, d_a, n, ld, clean, d_info);
    LFB_API_END(h)
}
int lfb_assemble_q_dev_f64(lfb_handle *h, const double *d_m, int64_t rows, int64_t cols, int64_t ld, int64_t shift,
                           const double *d_signs, double *d_q, int64_t ldq) {
    if (shift > std::min(rows, cols)) return fail(h, LFB_INVALID_ARGUMENT, "shift exceeds matrix dimension");
    LFB_API_BEGIN(h)
    assemble_q<double>(*h, d_m, rows, cols, ld, shift, d_signs, d_q, ldq);
    LFB_API_END(h)
}
int lfb_sym_tridiagonal_dev_f64(lfb_handle *h, double *d_a, int64_t n, int64_t ld, double *d_off) {
    if (n < 1) return fail(h, LFB_EMPTY_MATRIX, "Matrix is empty");
    LFB_API_BEGIN(h)
    sym_tridiagonal<double>(*h, d_a, n, ld, d_off);
    LFB_API_END(h)
}
int lfb_eigh_dev_f64(lfb_handle *h, double *d_a, int64_t n, int64_t ld, double *vals_host, double *d_q, int64_t ldq) {
    if (n < 0 || (n > 0 && !vals_host)) return fail(h, LFB_INVALID_ARGUMENT, "bad arguments");
    LFB_API_BEGIN(h)
    symmetric_eig<double>(*h, d_a, n, ld, vals_host, d_q, ldq);
    LFB_API_END(h)
}
int lfb_svd_dev_f64(lfb_handle *h, double *d_a, int64_t rows, int64_t cols, int64_t ld, double *sv_host, double *d_u, int64_t ldu,
                    double *d_v, int64_t ldv) {
    if (std::min(rows, cols) < 1) return fail(h, LFB_EMPTY_MATRIX, "Matrix is empty");
    if (!sv_host) return fail(h, LFB_INVALID_ARGUMENT, "sigma is null");
    LFB_API_BEGIN(h)
    svd_dev<double>(*h, d_a, rows, cols, ld, sv_host, d_u, ldu, d_v, ldv);
    LFB_API_END(h)
}
int lfb_bidiagonal_dev_f64(lfb_handle *h, double *d_a, int64_t rows, int64_t cols, int64_t ld, double *d_d, double *d_e) {
    if (std::min(rows, cols) < 1) return fail(h, LFB_EMPTY_MATRIX, "Matrix is empty");
    LFB_API_BEGIN(h)
    bidiagonal<double>(*h, d_a, rows, cols, ld, d_d, d_e);
    LFB_API_END(h)
}
int lfb_qr_batched_dev_f32(lfb_handle *h, float *d_a, int64_t batch, int64_t m, int64_t n, float *d_diag) {
    if (m < n) return fail(h, LFB_NOT_THIN, "Expected matrix rows >= cols");
    if (m > 32 || n > 32) return fail(h, LFB_UNSUPPORTED, "batched QR supports m, n <= 32");
    LFB_API_BEGIN(h)
    qr_batched<float>(*h, d_a, batch, m, n, d_diag);
    LFB_API_END(h)
}
int lfb_tsqr_local_r_dev_f64(lfb_handle *h, double *d_a, int64_t rows, int64_t cols, int64_t ld, double *d_r, int64_t ldr) {
    if (rows < cols) return fail(h, LFB_NOT_THIN, "Expected matrix rows >= cols");
    LFB_API_BEGIN(h)
    tsqr_local_r<double>(*h, d_a, rows, cols, ld, d_r, ldr);
    LFB_API_END(h)
}
int lfb_qr_tsqr_dev_f64(lfb_handle *h, double *d_a, int64_t rows, int64_t cols, int64_t ld, double *d_diag) {
    if (rows < cols) return fail(h, LFB_NOT_THIN, "Expected matrix rows >= cols");
    LFB_API_BEGIN(h)
    qr_tsqr<double>(*h, d_a, rows, cols, ld, d_diag);
    LFB_API_END(h)
}
int lfb_qr_tsqr_dev_f32(lfb_handle *h, float *d_a, int64_t rows, int64_t cols, int64_t ld, float *d_diag) {
    if (rows < cols) return fail(h, LFB_NOT_THIN, "Expected matrix rows >= cols");
    LFB_API_BEGIN(h)
    qr_tsqr<float>(*h, d_a, rows, cols, ld, d_diag);
    LFB_API_END(h)
}
int lfb_tsqr_explicit_q_dev_f64(lfb_handle *h, double *d_a, int64_t rows, int64_t cols, int64_t ld, double *d_r, int64_t ldr) {
    if (rows < cols) return fail(h, LFB_NOT_THIN, "Expected matrix rows >= cols");
    if (cols < 1) return LFB_OK;
    LFB_API_BEGIN(h)
    const int64_t ldw = round_up(rows, 2);
    DevBuf<double> Wk(*h, (size_t)ldw * cols);
    tsqr_explicit_q<double>(*h, d_a, rows, cols, ld, Wk, ldw, d_r, ldr);
    LFB_API_END(h)
}
int lfb_tsqr_leaf_dev_f64(lfb_handle *h, const double *d_a, int64_t rows, int64_t cols, int64_t ld, double *d_r, int64_t ldr,
                          double *d_rinv, int64_t ldri, int *ok) {
    if (!ok) return fail(h, LFB_INVALID_ARGUMENT, "ok is null");
    *ok = 0;
    if (rows < cols) return fail(h, LFB_NOT_THIN, "Expected matrix rows >= cols");
    LFB_API_BEGIN(h)
    *ok = (rows >= 4 * cols && cols <= 512 && cholqr_factor<double>(*h, d_a, rows, cols, ld, d_r, ldr, d_rinv, ldri)) ? 1 : 0;
    LFB_API_END(h)
}
int lfb_tsqr_apply_q_dev_f64(lfb_handle *h, double *d_q, int64_t rows, int64_t cols, int64_t ld, const double *d_qs, int64_t ldqs) {
    if (rows < 0 || cols < 0) return fail(h, LFB_INVALID_ARGUMENT, "negative dimension");
    LFB_API_BEGIN(h)
    tsqr_apply_q<double>(*h, d_q, rows, cols, ld, d_qs, ldqs);
    LFB_API_END(h)
}
int lfb_hh_reconstruct_top_dev_f64(lfb_handle *h, double *d_qtop, int64_t n, int64_t ld, const double *d_r, int64_t ldr,
                                   double *d_u, int64_t ldu, double *d_diag) {
    if (n < 0) return fail(h, LFB_INVALID_ARGUMENT, "negative dimension");
    LFB_API_BEGIN(h)
    hh_reconstruct_top<double>(*h, d_qtop, n, ld, d_r, ldr, d_u, ldu, d_diag);
    LFB_API_END(h)
}
int lfb_hh_reconstruct_rows_dev_f64(lfb_handle *h, double *d_q, int64_t rows, int64_t n, int64_t ld, const double *d_u, int64_t ldu) {
    if (rows < 0 || n < 0) return fail(h, LFB_INVALID_ARGUMENT, "negative dimension");
    LFB_API_BEGIN(h)
    trsm_right_upper<double>(*h, rows, n, d_u, ldu, d_q, ld);
    LFB_API_END(h)
}
int lfb_gemm_dev_f64(lfb_handle *h, int ta, int tb, int64_t m, int64_t n, int64_t k, double alpha, const double *d_a,
                     int64_t lda, const double *d_b, int64_t ldb, double beta, double *d_c, int64_t ldc) {
    LFB_API_BEGIN(h)
    gemm<double>(*h, ta, tb, m, n, k, alpha, d_a, lda, d_b, ldb, beta, d_c, ldc);
    LFB_API_END(h)
}
int lfb_gemm_dev_f32(lfb_handle *h, int ta, int tb, int64_t m, int64_t n, int64_t k, float alpha, const float *d_a,
                     int64_t lda, const float *d_b, int64_t ldb, float beta, float *d_c, int64_t ldc) {
    LFB_API_BEGIN(h)
    gemm<float>(*h, ta, tb, m, n, k, alpha, d_a, lda, d_b, ldb, beta, d_c, ldc);
    LFB_API_END(h)
}
int lfb_profile_begin(lfb_handle *h) {
    if (!h) return LFB_INVALID_ARGUMENT;
    h->prof_on = true;
    h->prof_used = 0;
    h->prof_flops = 0.0;
    return LFB_OK;
}
int lfb_profile_end(lfb_handle *h, double *gemm_ms, double *gemm_flops, int64_t *gemm_calls) {
    LFB_API_BEGIN(h)
    h->prof_on = false;
    LFB_CUDA(cudaStreamSynchronize(h->stream));
    double ms = 0.0;
    for (size_t i = 0; i + 1 < h->prof_used; i += 2) {
        float t = 0.f;
        LFB_CUDA(cudaEventElapsedTime(&t, h->prof_ev[i], h->prof_ev[i + 1]));
        ms += t;
    }
    if (gemm_ms) *gemm_ms = ms;
    if (gemm_flops) *gemm_flops = h->prof_flops;
    if (gemm_calls) *gemm_calls = (int64_t)(h->prof_used / 2);
    LFB_API_END(h)
}
int lfb_debug_panel_phases(lfb_handle *h, long long *out4) {
    LFB_API_BEGIN(h)
    if (!h->panel_dbg) {
        LFB_CUDA(cudaMalloc(&h->panel_dbg, 8 * sizeof(long long)));
        LFB_CUDA(cudaMemset(h->panel_dbg, 0, 8 * sizeof(long long)));
    }
    LFB_CUDA(cudaStreamSynchronize(h->stream));
    LFB_CUDA(cudaDeviceSynchronize());
    if (out4) LFB_CUDA(cudaMemcpy(out4, h->panel_dbg, 4 * sizeof(long long), cudaMemcpyDeviceToHost));
    LFB_API_END(h)
}
int lfb_microbench_fp64(lfb_handle *h, int kind, double *gflops) {
    if (!gflops) return LFB_INVALID_ARGUMENT;
    LFB_API_BEGIN(h)
    *gflops = microbench_fp64(*h, kind);
    LFB_API_END(h)
}

int lfb_microbench_kernel(lfb_handle *h, const char *name, int64_t n, int reps, double *us_per_launch) {
    if (!us_per_launch || !name || n < 0 || reps < 1) return LFB_INVALID_ARGUMENT;
    LFB_API_BEGIN(h)
    const std::string k(name);
    if (k == "trd_symv") *us_per_launch = microbench_trd(*h, 0, n, reps);
    else if (k == "trd_head") *us_per_launch = microbench_trd(*h, 1, n, reps);
    else if (k == "bd_gemv_n") *us_per_launch = microbench_bd_gemv(*h, 0, 4 * n, n, reps);   // the 4:1 aspect of configs[4]
    else if (k == "bd_gemv_t") *us_per_launch = microbench_bd_gemv(*h, 1, 4 * n, n, reps);
    else if (k == "potf2") *us_per_launch = microbench_potf2(*h, (int)(n & 3), reps);      // n: 0 old, 1 new, +2 = factor only
    else return fail(h, LFB_INVALID_ARGUMENT, "unknown kernel name");
    LFB_API_END(h)
}

}  // extern "C"
