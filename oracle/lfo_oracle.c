/*
 * ORACLE -- test infrastructure only.  Never linked into, imported by, or
 * called from the product (linfa_linalg_b200/).  Only tests/, 
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library, and only as the checker / reported CPU baseline.
 *
 * CPU restatement (plain C, single thread) of the Householder / Cholesky /
 * triangular hot path of rust-ml/linfa-linalg v0.2.1:
 *   src/householder.rs:9-93, src/reflection.rs:26-37, src/qr.rs:32-44,91-120,
 *   src/cholesky.rs:51-83, src/triangular.rs:37-53,95-144,
 *   src/tridiagonal.rs:31-66, src/bidiagonal.rs:27-59, and (for the end-to-end eigh parity of the
 *   Givens phase) src/eigh.rs:10-199 with src/givens.rs:12-106.
 * The Rust crate itself cannot be built in this image (no rustc/cargo), so
 * parity is PINNED against the reference's own known-answer tests instead
 * (tests/test_oracle_kat.py reproduces every KAT listed in SURVEY.md 8c) and
 * cross-checked against LAPACK after sign normalisation.
 *
 * Build: see oracle/Makefile  (gcc -O3 -march=x86-64-v3 (portable to the GPU box host), no fast-math: the
 * reference does no reassociation beyond ndarray's unrolled dot).
 */
#include <math.h>
#include <stdint.h>
#include <stddef.h>

#define T double
#define SFX _f64
#define SQRT sqrt
#define FABS fabs
#define HYPOT hypot
#include "lfo_impl.inc"
#undef T
#undef SFX
#undef SQRT
#undef FABS
#undef HYPOT

#define T float
#define SFX _f32
#define SQRT sqrtf
#define FABS fabsf
#define HYPOT hypotf
#include "lfo_impl.inc"
#undef T
#undef SFX
#undef SQRT
#undef FABS
#undef HYPOT

/* Batched thin QR over `batch` packed row-major m x n matrices (the reference has no batched
 * entry point: this is qr.rs:32-44 in a loop, as a caller would write it). */
void lfo_qr_batched_f32(float *a, int64_t batch, int64_t m, int64_t n, float *diag) {
    for (int64_t b = 0; b < batch; ++b) lfo_qr_f32(a + b * m * n, m, n, n, 1, diag + b * n);
}
void lfo_qr_batched_f64(double *a, int64_t batch, int64_t m, int64_t n, double *diag) {
    for (int64_t b = 0; b < batch; ++b) lfo_qr_f64(a + b * m * n, m, n, n, 1, diag + b * n);
}
