// Shared machinery of the Givens phases of eigh (eigh.cu) and svd (svd.cu): max|a| scaling kernels, the
// rotation wavefront kernel and the host-side applier that streams chains of rotations to the device.
#pragma once
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

// One rotation on two arbitrary columns (svd.rs:312,348 rotate a strided column pair): a' = a c + s b; b' = -s a + b c.
template <typename T>
__global__ void __launch_bounds__(256) rot_pair_kernel(T *Q, int64_t ldq, int64_t nrows, int64_t ca, int64_t cb, T c, T s) {
    const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r >= nrows) return;
    const T a = Q[r + ca * ldq], b = Q[r + cb * ldq];
    Q[r + ca * ldq] = a * c + s * b;
    Q[r + cb * ldq] = -s * a + b * c;
}

template <typename T>
__global__ void __launch_bounds__(256) scale_col_kernel(T *Q, int64_t ldq, int64_t nrows, int64_t col, T f) {
    const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r < nrows) Q[r + col * ldq] *= f;
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
// thread_local: set from the handle's option at the start of every eigh / svd call; two handles on two threads do not share it
static thread_local bool FAST_HYPOT = true;
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
        while (!ch.c.empty() && ch.c.back() == T(1) && ch.s.back() == T(0)) { ch.c.pop_back(); ch.s.pop_back(); }   // trailing identities
        if (ch.c.empty()) return;
        pending.push_back(std::move(ch));
        if ((int)pending.size() >= (h.opt.rot_serial ? 1 : KCH)) flush();
    }
    // A single rotation on columns (ca, cb), in order with the chains pushed so far.
    void pair(int64_t ca, int64_t cb, T c, T s) {
        flush();
        rot_pair_kernel<T><<<(unsigned)cdiv(nrows, 256), 256, 0, h.stream>>>(Q, ldq, nrows, ca, cb, c, s);
        LFB_LAUNCH_CHECK(h);
    }
    void scale_col(int64_t col, T f) {
        flush();
        scale_col_kernel<T><<<(unsigned)cdiv(nrows, 256), 256, 0, h.stream>>>(Q, ldq, nrows, col, f);
        LFB_LAUNCH_CHECK(h);
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

}  // namespace lfb
