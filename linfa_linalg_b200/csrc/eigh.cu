// Symmetric eigendecomposition: src/eigh.rs:10-129 (symmetric_eig) end to end.
//
// Phase 1 (scale by max|a|, tridiagonalise, generate Q) runs on the kernels of tridiag.cu / householder.cu.
// Phase 2, the implicit symmetric QR iteration with Wilkinson shifts (eigh.rs:51-128), is O(n^2) SCALAR work
// on the tridiagonal (d, e) -- a strictly sequential bulge chase -- plus O(n^3) rotation work on Q.  Here the
// scalar recurrence runs on the host, exactly as the reference orders it, and emits "chains": runs of Givens
// rotations acting on consecutive column pairs (i, i+1), i = p .. p+cnt-1 (one chain per QR sweep, a chain of
// one for a deflated 2x2 block).  Every row of Q sees the same sequence of rotations, so the device applies
// them row-parallel.  K chains are applied in ONE pass over Q as a wavefront: chain s runs two columns
// behind chain s-1, a thread keeps a sliding window of 2K row entries in registers, each element of Q is
// read and written once per K sweeps (the reference touches it once per sweep), and the host is already
// chasing the next K sweeps while the device applies the previous ones.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <limits>
#include <memory>
#include <vector>

#include "common.cuh"
#include "dev_utils.cuh"

namespace lfb {

namespace {

constexpr int KCH = 16;         // chains per pass
constexpr int WIN = 2 * KCH;    // register window (columns)
constexpr int PAD = 2 * (KCH - 1);

template <typename T> struct UIntOf;
template <> struct UIntOf<double> { using type = unsigned long long; };
template <> struct UIntOf<float> { using type = unsigned int; };

// max |a_ij| (eigh.rs:27-30; f::max ignores NaN).  |x| >= 0, so the IEEE bit pattern orders like an unsigned.
template <typename T>
__global__ void __launch_bounds__(256) absmax_kernel(const T *__restrict__ A, int64_t ld, int64_t rows, int64_t cols, T *out) {
    using U = typename UIntOf<T>::type;
    T m = T(0);
    for (int64_t c = blockIdx.y; c < cols; c += gridDim.y)
        for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x) {
            const T v = fabs(A[r + c * ld]);
            if (v > m) m = v;      // false for NaN
        }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const T w = __shfl_xor_sync(0xffffffffu, m, o);
        if (w > m) m = w;
    }
    if ((threadIdx.x & 31) == 0 && m > T(0)) {
        U bits;
        memcpy(&bits, &m, sizeof(T));
        atomicMax(reinterpret_cast<U *>(out), bits);
    }
}

// matrix /= amax when amax != 0 (eigh.rs:32-34)
template <typename T>
__global__ void __launch_bounds__(256) scale_div_kernel(T *A, int64_t ld, int64_t rows, int64_t cols, const T *amax) {
    const T d = *amax;
    if (d == T(0)) return;
    for (int64_t c = blockIdx.y; c < cols; c += gridDim.y)
        for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x)
            A[r + c * ld] /= d;
}

template <typename T>
__global__ void extract_diag_kernel(const T *__restrict__ A, int64_t ld, int64_t n, T *diag) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) diag[i] = A[i + i * ld];
}

// Applies KCH chains of rotations to the columns [lo, lo + span] of Q (nrows x ., column-major), thread per row.
// Cp / Sp: [KCH][len] with len = span + 2 PAD + 1; entry x + PAD of chain s rotates columns (lo+x, lo+x+1):
//     a' = a c + s b ;  b' = -s a + b c          (givens.rs:97-100 rotate_rows), identity elsewhere.
// At time tau chain s works on x = tau - 2 s, so it never overtakes chain s - 1.
//
// Everything that comes from global memory is landed in shared memory with cp.async, DEPTH steps ahead: the
// row's next columns (a ring of RING slots per thread) and, once per TS steps, the coefficients of the next
// TS steps.  Two earlier versions kept the prefetched values in registers; ncu (profiles/r1_eigh_rot.md)
// showed every DFMA stalled on `long_scoreboard`: a warp has only six scoreboards, so 32 loads in flight alias
// with the shared-memory loads of the coefficients and the arithmetic ends up waiting for DRAM anyway.
constexpr int TS = 32;          // time steps per coefficient stage (multiple of WIN)
constexpr int RNT = 64;         // threads (rows) per CTA
constexpr int DEPTH = 32;       // columns prefetched ahead
constexpr int RING = 48;        // slots of the per-thread column ring (> DEPTH)
static_assert(TS % WIN == 0 && TS <= DEPTH && RING > DEPTH, "staging constants");

template <typename T, bool STAGED>
__global__ void __launch_bounds__(RNT) rot_wave_kernel(T *Q, int64_t ldq, int64_t nrows, int64_t lo, int64_t span,
                                                       const T *__restrict__ Cp, const T *__restrict__ Sp, int64_t len) {
    __shared__ T sc[2][KCH][TS], ss[2][KCH][TS];
    __shared__ T sq[RING][RNT];
    const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const bool active = r < nrows;
    T *q = Q + (active ? r : 0) + lo * ldq;
    const int64_t ntau = span + PAD + 1;      // the last step only stores column `span`
    auto stage = [&](int buf, int64_t tb) {   // coefficients of steps [tb, tb + TS) -> sc/ss[buf]  (no commit)
        for (int e = threadIdx.x; e < KCH * TS; e += RNT) {
            const int s = e / TS, k = e % TS;
            const int64_t idx = tb + k - 2 * s + PAD;
            if (idx < len) {
                dev::cp_async<(int)sizeof(T)>(&sc[buf][s][k], Cp + (int64_t)s * len + idx);
                dev::cp_async<(int)sizeof(T)>(&ss[buf][s][k], Sp + (int64_t)s * len + idx);
            }
        }
    };
    auto prefetch = [&](int64_t x) {          // column x of this thread's row -> ring slot x % RING  (no commit)
        if (active && x <= span) dev::cp_async<(int)sizeof(T)>(&sq[x % RING][threadIdx.x], q + x * ldq);
    };
    T w[WIN];
#pragma unroll
    for (int u = 0; u < WIN; ++u) w[u] = T(0);
    if (active) w[0] = q[0];
    if (STAGED) stage(0, 0);
    for (int j = 0; j < DEPTH; ++j) {         // group j carries column j + 1 (the virtual steps -DEPTH .. -1)
        prefetch(j + 1);
        dev::cp_async_commit();
    }
    int it = 0;
    for (int64_t tb = 0; tb < ntau; tb += TS, ++it) {
        const int buf = it & 1;
        // the coefficient copies of this stage travelled with the first step of the previous stage (or with the
        // prologue): they are older than the newest DEPTH - 1 groups
        dev::cp_async_wait<DEPTH - 1>();
        __syncthreads();
        for (int g = 0; g < TS; g += WIN) {
#pragma unroll
            for (int u = 0; u < WIN; ++u) {
                const int64_t tau = tb + g + u;
                if (tau < ntau) {
                    if (STAGED && g + u == 0 && tb + TS < ntau) stage(buf ^ 1, tb + TS);
                    prefetch(tau + 1 + DEPTH);
                    dev::cp_async_commit();
                    dev::cp_async_wait<DEPTH>();                                 // the group of step tau - DEPTH: column tau + 1
                    w[(u + 1) % WIN] = (active && tau + 1 <= span) ? sq[(tau + 1) % RING][threadIdx.x] : T(0);
#pragma unroll
                    for (int s = 0; s < KCH; ++s) {
                        T c, sn;
                        if (STAGED) {
                            c = sc[buf][s][g + u]; sn = ss[buf][s][g + u];
                        } else {
                            const int64_t idx = tau - 2 * s + PAD;
                            c = Cp[(int64_t)s * len + idx]; sn = Sp[(int64_t)s * len + idx];
                        }
                        const int i0 = (u - 2 * s + 4 * WIN) % WIN, i1 = (u - 2 * s + 1 + 4 * WIN) % WIN;
                        const T a = w[i0], b = w[i1];
                        w[i0] = a * c + sn * b;
                        w[i1] = -sn * a + b * c;
                    }
                    const int64_t xf = tau - PAD;                               // column xf has seen its last rotation
                    if (active && xf >= 0) q[xf * ldq] = w[(u - PAD + 4 * WIN) % WIN];
                }
            }
        }
        __syncthreads();      // everyone is done with `buf` before the next stage's first step overwrites buf ^ 1 ... buf
    }
    dev::cp_async_wait<0>();
}

// ---- host side: the scalar recurrences of eigh.rs on (diag, off), emitting chains ------------------------
template <typename T>
struct Chain {
    int64_t p;               // first column
    std::vector<T> c, s;     // rotation t acts on columns (p + t, p + t + 1)
};

template <typename T> inline T h_signum(T x) { return std::signbit(x) ? T(-1) : T(1); }

// x.hypot(y) (givens.rs:19).  The matrix was scaled to max|a| = 1, so x^2 + y^2 cannot overflow; the plain
// square root is used unless the squares get close to underflow.
static bool FAST_HYPOT = true;
template <typename T>
inline T h_hypot(T x, T y) {
    const T r2 = x * x + y * y;
    return (FAST_HYPOT && r2 > std::numeric_limits<T>::min() * T(1e16)) ? std::sqrt(r2) : std::hypot(x, y);
}

// eigh.rs:178-186
template <typename T>
inline T wilkinson_shift(T tmm, T tnn, T tmn) {
    if (tmn != T(0)) {
        const T tmn_sq = tmn * tmn;
        const T d = (tmm - tnn) * T(0.5);
        return tnn - tmn_sq / (d + h_signum(d) * std::sqrt(d * d + tmn_sq));
    }
    return tnn;
}

// eigh.rs:131-170
template <typename T>
inline void delimit_subproblem(const std::vector<T> &diag, std::vector<T> &off, int64_t end, T eps, int64_t &start, int64_t &nend) {
    int64_t n = end;
    while (n > 0) {
        const int64_t m = n - 1;
        if (std::fabs(off[m]) > eps * (std::fabs(diag[n]) + std::fabs(diag[m]))) break;
        n -= 1;
    }
    if (n == 0) { start = 0; nend = 0; return; }
    int64_t ns = n - 1;
    while (ns > 0) {
        const int64_t m = ns - 1;
        if (off[m] == T(0) || std::fabs(off[m]) <= eps * (std::fabs(diag[ns]) + std::fabs(diag[m]))) {
            off[m] = T(0);
            break;
        }
        ns -= 1;
    }
    start = ns; nend = n;
}

// Streams batches of KCH chains to the device and launches the wavefront kernel.
template <typename T>
struct RotationApplier {
    lfb_handle &h;
    T *Q; int64_t ldq, nrows;
    std::vector<Chain<T>> pending;
    T *pin[2] = {nullptr, nullptr};
    size_t pin_elems = 0;
    cudaEvent_t ev[2] = {nullptr, nullptr};
    bool ev_used[2] = {false, false};
    int cur = 0;
    T *dbuf = nullptr; size_t dbuf_elems = 0;
    double t_wait = 0, t_pack = 0; int64_t batches = 0;     // host-side profile (option trd_profile)

    RotationApplier(lfb_handle &hh, T *q, int64_t ld, int64_t n) : h(hh), Q(q), ldq(ld), nrows(n) {
        pin_elems = (size_t)2 * KCH * (size_t)(n + 2 * PAD + 2);
        for (int b = 0; b < 2; ++b) {
            LFB_CUDA(cudaMallocHost(&pin[b], pin_elems * sizeof(T)));
            LFB_CUDA(cudaEventCreateWithFlags(&ev[b], cudaEventDisableTiming));
        }
        dbuf_elems = pin_elems;
        dbuf = (T *)h.dalloc(dbuf_elems * sizeof(T));
    }
    ~RotationApplier() {
        for (int b = 0; b < 2; ++b) {
            if (ev[b]) cudaEventDestroy(ev[b]);
            if (pin[b]) cudaFreeHost(pin[b]);
        }
        if (dbuf) h.dfree(dbuf);
    }
    void push(Chain<T> &&ch) {
        if (ch.c.empty()) return;
        pending.push_back(std::move(ch));
        if ((int)pending.size() >= (h.opt.rot_serial ? 1 : KCH)) flush();
    }
    void flush() {
        if (pending.empty()) return;
        int64_t lo = pending[0].p, hi = pending[0].p + (int64_t)pending[0].c.size();
        for (auto &ch : pending) {
            lo = std::min(lo, ch.p);
            hi = std::max(hi, ch.p + (int64_t)ch.c.size());
        }
        const int64_t span = hi - lo, len = span + 2 * PAD + 1;
        const auto c0 = std::chrono::steady_clock::now();
        if (ev_used[cur]) LFB_CUDA(cudaEventSynchronize(ev[cur]));     // the copy out of this staging buffer is done
        const auto c1 = std::chrono::steady_clock::now();
        T *C = pin[cur], *S = pin[cur] + (size_t)KCH * len;
        std::fill(C, C + (size_t)KCH * len, T(1));
        std::fill(S, S + (size_t)KCH * len, T(0));
        for (size_t s = 0; s < pending.size(); ++s) {
            const auto &ch = pending[s];
            const int64_t off = ch.p - lo + PAD;
            std::copy(ch.c.begin(), ch.c.end(), C + s * len + off);
            std::copy(ch.s.begin(), ch.s.end(), S + s * len + off);
        }
        LFB_CUDA(cudaMemcpyAsync(dbuf, pin[cur], sizeof(T) * 2 * KCH * len, cudaMemcpyHostToDevice, h.stream));
        LFB_CUDA(cudaEventRecord(ev[cur], h.stream));
        ev_used[cur] = true;
        if (h.opt.rot_staged)
            rot_wave_kernel<T, true><<<(unsigned)cdiv(nrows, RNT), RNT, 0, h.stream>>>(Q, ldq, nrows, lo, span, dbuf, dbuf + (size_t)KCH * len, len);
        else
            rot_wave_kernel<T, false><<<(unsigned)cdiv(nrows, RNT), RNT, 0, h.stream>>>(Q, ldq, nrows, lo, span, dbuf, dbuf + (size_t)KCH * len, len);
        LFB_LAUNCH_CHECK(h);
        cur ^= 1;
        pending.clear();
        const auto c2 = std::chrono::steady_clock::now();
        t_wait += std::chrono::duration<double>(c1 - c0).count();
        t_pack += std::chrono::duration<double>(c2 - c1).count();
        ++batches;
    }
};

}  // namespace

// eigh.rs:10-129.  dA (n x n column-major, ld) is consumed.  vals: HOST array of n entries (reference order,
// unsorted).  dQ: device n x n (ldq) for the eigenvectors as columns, or nullptr (eigvalsh).
template <typename T>
void symmetric_eig(lfb_handle &h, T *dA, int64_t n, int64_t ld, T *vals, T *dQ, int64_t ldq) {
    if (n < 1) return;                                                          // :16-25
    FAST_HYPOT = h.opt.fast_hypot != 0;
    const T eps = std::numeric_limits<T>::epsilon();                            // :214 (A::epsilon())
    DevBuf<T> scal(h, 2), dOff(h, n), dDiag(h, n);
    LFB_CUDA(cudaMemsetAsync(scal.get(), 0, 2 * sizeof(T), h.stream));
    {
        dim3 grid((unsigned)std::min<int64_t>(cdiv(n, 256), 64), (unsigned)std::min<int64_t>(n, 1024));
        absmax_kernel<T><<<grid, 256, 0, h.stream>>>(dA, ld, n, n, scal.get());            // :27-30
        LFB_LAUNCH_CHECK(h);
        scale_div_kernel<T><<<grid, 256, 0, h.stream>>>(dA, ld, n, n, scal.get());          // :32-34
        LFB_LAUNCH_CHECK(h);
    }
    sym_tridiagonal<T>(h, dA, n, ld, dOff.get());                                             // :36
    if (dQ) assemble_q<T>(h, dA, n, n, ld, 1, dOff.get(), dQ, ldq);                           // :37-41 (tridiagonal.rs:90-92)
    extract_diag_kernel<T><<<(unsigned)cdiv(n, 256), 256, 0, h.stream>>>(dA, ld, n, dDiag.get());
    LFB_LAUNCH_CHECK(h);
    std::vector<T> diag(n), off(std::max<int64_t>(n - 1, 1));
    T amax = T(0);
    LFB_CUDA(cudaMemcpyAsync(diag.data(), dDiag.get(), sizeof(T) * n, cudaMemcpyDeviceToHost, h.stream));
    if (n > 1) LFB_CUDA(cudaMemcpyAsync(off.data(), dOff.get(), sizeof(T) * (n - 1), cudaMemcpyDeviceToHost, h.stream));
    LFB_CUDA(cudaMemcpyAsync(&amax, scal.get(), sizeof(T), cudaMemcpyDeviceToHost, h.stream));
    LFB_CUDA(cudaStreamSynchronize(h.stream));
    for (int64_t i = 0; i + 1 < n; ++i) off[i] = std::fabs(off[i]);                          // tridiagonal.rs:96-101
    if (n == 1) {                                                                            // :44-47
        vals[0] = diag[0] * amax;
        return;
    }
    std::unique_ptr<RotationApplier<T>> app;
    if (dQ) app.reset(new RotationApplier<T>(h, dQ, ldq, n));

    const auto t_loop0 = std::chrono::steady_clock::now();
    int64_t start, end;
    delimit_subproblem<T>(diag, off, n - 1, eps, start, end);                                // :49
    while (end != start) {                                                                   // :51
        const int64_t subdim = end - start + 1;
        if (subdim > 2) {                                                                    // :55
            const int64_t m = end - 1, nn = end;
            T x = diag[start] - wilkinson_shift<T>(diag[m], diag[nn], off[m]);               // :59
            T y = off[start];
            Chain<T> ch;
            ch.p = start;
            if (dQ) { ch.c.reserve(nn - start); ch.s.reserve(nn - start); }
            for (int64_t i = start; i < nn; ++i) {                                           // :62
                const int64_t j = i + 1;
                if (y == T(0)) break;                                                        // givens.rs:18 -> :96-98
                const T r = h_hypot(x, y), c = x / r, s = -y / r;                            // givens.rs:19-21
                if (i > start) off[i - 1] = r;                                               // :66-68
                const T cc = c * c, ss = s * s, cs = c * s;
                const T mii = diag[i], mjj = diag[j], mij = off[i];
                const T b = cs * mij * T(2);
                diag[i] = cc * mii + ss * mjj - b;                                           // :78-80
                diag[j] = ss * mii + cc * mjj + b;
                off[i] = cs * (mii - mjj) + mij * (cc - ss);
                if (i != nn - 1) {                                                           // :82-86
                    x = off[i];
                    y = -s * off[i + 1];
                    off[i + 1] *= c;
                }
                if (dQ) { ch.c.push_back(c); ch.s.push_back(-s); }                           // :89-94: the inverse rotation
            }
            if (dQ) app->push(std::move(ch));
            if (std::fabs(off[m]) <= eps * (std::fabs(diag[m]) + std::fabs(diag[nn]))) end -= 1;   // :100-102
        } else if (subdim == 2) {                                                            // :103
            const T h00 = diag[start], h10 = off[start], h11 = diag[start + 1];
            const T val = (h00 - h11) * T(0.5);                                              // :188-198 compute_2x2_eigvals
            const T discr = h10 * h10 + val * val;                                           // >= 0 for a symmetric block
            const T sq = std::sqrt(discr), half = (h00 + h11) * T(0.5);
            const T e0 = half + sq, e1 = half - sq;
            T b0 = e0 - diag[start + 1], b1 = off[start];                                    // :111
            if (h.opt.eigh_stable_2x2 && h00 < h11) {
                // (e0 - h11, h10) and (h10, e0 - h00) span the same line ((e0 - h11)(e0 - h00) = h10^2).  For
                // h00 < h11 the reference's e0 - h11 = h10^2 / (e0 - h00) cancels catastrophically when h10 is
                // barely above the deflation threshold: its rounding noise is then as large as h10 itself and
                // the "rotation" mixes two well-separated eigenvectors (seen with the reference's own algorithm
                // on the CPU oracle: residual 0.1 .. 1.7 instead of 3e-10 at n = 600 for ~20 % of 1e-15-sized
                // perturbations of the input).  Same direction and sign, computed without cancellation:
                b0 = std::fabs(h10);
                b1 = h_signum(h10) * (e0 - h00);
            }
            diag[start] = e0;
            diag[start + 1] = e1;
            if (dQ) {                                                                        // :116-121, givens.rs:48-57
                const T norm = std::hypot(b0, b1);
                if (norm > eps) {
                    Chain<T> ch;
                    ch.p = start;
                    ch.c.push_back(b0 / norm);
                    ch.s.push_back(b1 / norm);
                    app->push(std::move(ch));
                }
            }
            end -= 1;
        }
        delimit_subproblem<T>(diag, off, end, eps, start, end);                              // :125-127
    }
    if (dQ) {
        app->flush();
        const auto c0 = std::chrono::steady_clock::now();
        LFB_CUDA(cudaStreamSynchronize(h.stream));
        if (h.opt.trd_profile)
            fprintf(stderr, "[eigh_profile] n=%lld: host recurrence + packing %.3f s (waiting for a staging buffer %.3f s, packing + "
                            "launch %.3f s, %lld passes), final drain of the device queue %.3f s\n", (long long)n,
                    std::chrono::duration<double>(c0 - t_loop0).count(), app->t_wait, app->t_pack, (long long)app->batches,
                    std::chrono::duration<double>(std::chrono::steady_clock::now() - c0).count());
    }
    for (int64_t i = 0; i < n; ++i) vals[i] = diag[i] * amax;                                // :130
}

template void symmetric_eig<float>(lfb_handle &, float *, int64_t, int64_t, float *, float *, int64_t);
template void symmetric_eig<double>(lfb_handle &, double *, int64_t, int64_t, double *, double *, int64_t);

}  // namespace lfb
