/*
 * linfa_b200.h -- C ABI of liblinfa_b200.so, the B200 (sm_100a) dense-factorisation engine behind
 * linfa-linalg's Householder / Cholesky / triangular hot path.
 *
 * The reference (rust-ml/linfa-linalg v0.2.1) has no FFI of its own: its "interface" is the set of
 * blanket trait impls on ndarray ArrayBase<S, Ix2>.  Each entry point below replaces the BODY of one
 * of those trait methods (cited per function as file:line under /root/reference) and is what the
 * Rust shim in INTEGRATION.md binds with `extern "C"`.  One call per whole factorisation -- never
 * per reflector -- so the blocked compact-WY structure lives behind the boundary.
 *
 * Conventions
 *  - Plain pointers and sizes only.  `*_f32` / `*_f64` symbol families (A: NdFloat is f32|f64).
 *  - HOST entry points take ndarray-style strided views: (ptr, rows, cols, row_stride, col_stride),
 *    strides in ELEMENTS and SIGNED (negative and non-unit strides are legal, tests/common.rs:12-43).
 *    Results are written back in place into the caller's storage, exactly like the `*_into` /
 *    `*_inplace` trait methods.  The library never keeps a host pointer after it returns.
 *  - DEVICE entry points (`*_dev_*`) take device pointers to COLUMN-MAJOR buffers with a leading
 *    dimension (the engine's HBM layout) and run asynchronously on the handle's stream.
 *  - Every function returns an lfb_status.  Data-dependent failures (NotPositiveDefinite) are
 *    status codes, shape errors mirror LinalgError (src/lib.rs:33-60), CUDA failures are >= 100.
 *  - There is NO CPU fallback: without a CUDA device lfb_create fails with LFB_ERR_CUDA.
 *  - Blocking: host entry points synchronise the stream before returning.  A handle may be used
 *    from any thread, one call at a time.
 */
#ifndef LINFA_B200_H
#define LINFA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct lfb_handle lfb_handle;
typedef struct lfb_multi lfb_multi;   /* several devices of one box behind one handle (see the end of this file) */

typedef enum lfb_status {
    LFB_OK = 0,
    LFB_NOT_POSITIVE_DEFINITE = 1, /* LinalgError::NotPositiveDefinite (cholesky.rs:69-71) */
    LFB_NOT_THIN = 2,              /* LinalgError::NotThin            (qr.rs:34-36)        */
    LFB_NOT_SQUARE = 3,            /* LinalgError::NotSquare          (lib.rs:64-71)       */
    LFB_EMPTY_MATRIX = 4,          /* LinalgError::EmptyMatrix        (tridiagonal.rs:33-35, bidiagonal.rs:30-32) */
    LFB_WRONG_ROWS = 5,            /* LinalgError::WrongRows          (triangular.rs:103-108, qr.rs:128-133) */
    LFB_NON_INVERTIBLE = 6,        /* LinalgError::NonInvertible      (qr.rs:134-136)      */
    LFB_INVALID_ARGUMENT = 7,
    LFB_UNSUPPORTED = 8,
    LFB_ERR_CUDA = 100,
    LFB_ERR_ALLOC = 101,
    LFB_ERR_NCCL = 102             /* libnccl.so.2 missing, or a communicator / collective failed */
} lfb_status;

/* triangular.rs:10-13  enum UPLO */
#define LFB_UPPER 0
#define LFB_LOWER 1

/* ---- lifetime ------------------------------------------------------------------------------ */
/* Creates a handle on CUDA device `device` (own non-blocking stream, workspace pool). */
int lfb_create(lfb_handle **out, int device);
int lfb_destroy(lfb_handle *h);
/* Message of the last failure on this handle ("" if none).  Valid until the next call. */
const char *lfb_last_error(lfb_handle *h);
/* Run subsequent calls on a caller-owned cudaStream_t, taken verbatim (NULL = the legacy default
 * stream); lfb_use_own_stream restores the handle's own non-blocking stream. */
int lfb_set_stream(lfb_handle *h, void *cuda_stream);
int lfb_use_own_stream(lfb_handle *h);
int lfb_synchronize(lfb_handle *h);
/* Version / build string, and the number of kernels this handle has launched so far. */
const char *lfb_version(void);
int64_t lfb_launch_count(lfb_handle *h);
/* Tunables: "qr_nb", "qr_sub", "chol_base", "gemm_tma" (0/1), ...; returns LFB_INVALID_ARGUMENT if unknown or out of range.
 * Routes that depend on the data and can be pinned for reproducibility studies (results agree to rounding either way):
 *   "qr_panel_cholqr" (default 1): 128-column panels of lfb_qr_f64 as guarded Cholesky-QR + Householder reconstruction when
 *       the panel's condition bound is <= "tsqr_cholqr_cond" (16), else Householder panel kernels; 0 = always Householder.
 *   "chol_waves" (default 3): lfb_cholesky_* on a host view with n >= 8192 factors in that many arrival waves of block columns
 *       while the rest of the matrix is still crossing PCIe; 1 = upload first (the summation order of the updates differs).
 *   "gemm_deterministic" (default 1): split-K partial sums reduced in a fixed order (bit-reproducible run to run). */
int lfb_set_option(lfb_handle *h, const char *key, int64_t value);

/* ---- QR: qr.rs:29-45 QRInto::qr_into (driver loop :38-41 over householder.rs:34-51) --------- */
/* In place: on return a[i.., i] holds the unit-norm reflector v_i, a[i, i+1..] row i of R (sign
 * scaled as the reference does), diag[i] the signed pivot (QRDecomp{qr,diag}, qr.rs:68-73).
 * rows < cols -> LFB_NOT_THIN.  0x0 is legal. */
int lfb_qr_f64(lfb_handle *h, double *a, int64_t rows, int64_t cols, int64_t rs, int64_t cs, double *diag);
int lfb_qr_f32(lfb_handle *h, float *a, int64_t rows, int64_t cols, int64_t rs, int64_t cs, float *diag);

/* qr.rs:29-45 again, for TALL-SKINNY matrices: identical contract and results (same compact factor, same diag, to
 * rounding), computed as TSQR over row chunks + Householder reconstruction (LU of Q - S) instead of one panel sweep
 * over all the rows.  lfb_qr_* takes this route by itself when option "qr_tsqr_auto" is 1 and the matrix has at least
 * two chunks of rows ("tsqr_chunk") and at most 512 columns.  On rank-deficient input both routes return a valid
 * factorisation (Q R = A, Q orthonormal) but not the same one -- R is not unique there. */
int lfb_qr_tsqr_f64(lfb_handle *h, double *a, int64_t rows, int64_t cols, int64_t rs, int64_t cs, double *diag);
int lfb_qr_tsqr_f32(lfb_handle *h, float *a, int64_t rows, int64_t cols, int64_t rs, int64_t cs, float *diag);

/* householder.rs:68-93 assemble_q (used by qr.rs:86-88 generate_q, tridiagonal.rs:90-92,
 * bidiagonal.rs:90-109).  m is the compact factor (rows x cols, read only), signs[i] = diag_fn(i)
 * (min(rows,cols) - shift entries are read), q is rows x min(rows,cols) written with (q_rs,q_cs). */
int lfb_assemble_q_f64(lfb_handle *h, const double *m, int64_t rows, int64_t cols, int64_t rs, int64_t cs,
                       int64_t shift, const double *signs, double *q, int64_t q_rs, int64_t q_cs);
int lfb_assemble_q_f32(lfb_handle *h, const float *m, int64_t rows, int64_t cols, int64_t rs, int64_t cs,
                       int64_t shift, const float *signs, float *q, int64_t q_rs, int64_t q_cs);

/* qr.rs:110-120 QRDecomp::qt_mul: b <- Q^T b in place, b is rows x bcols. */
int lfb_qt_mul_f64(lfb_handle *h, const double *qr, int64_t rows, int64_t cols, int64_t rs, int64_t cs,
                   const double *diag, double *b, int64_t bcols, int64_t b_rs, int64_t b_cs);
int lfb_qt_mul_f32(lfb_handle *h, const float *qr, int64_t rows, int64_t cols, int64_t rs, int64_t cs,
                   const float *diag, float *b, int64_t bcols, int64_t b_rs, int64_t b_cs);

/* ---- Cholesky: cholesky.rs:51-83 cholesky_inplace_dirty / cholesky_inplace ------------------- */
/* Reads only the lower triangle.  clean != 0 zeroes the strict upper triangle (:78-82), clean == 0
 * leaves it untouched.  On LFB_NOT_POSITIVE_DEFINITE *fail_index (may be NULL) is the first row j
 * whose pivot is <= 0 (:69-71); the matrix is then partially overwritten, as in the reference. */
int lfb_cholesky_f64(lfb_handle *h, double *a, int64_t rows, int64_t cols, int64_t rs, int64_t cs,
                     int clean, int64_t *fail_index);
int lfb_cholesky_f32(lfb_handle *h, float *a, int64_t rows, int64_t cols, int64_t rs, int64_t cs,
                     int clean, int64_t *fail_index);

/* ---- triangular.rs:95-144 solve_triangular_system ------------------------------------------- */
/* Solves a x = b in place on b (n x bcols); reads only the `uplo` triangle of a.  ext_diag == NULL
 * takes the diagonal from a (SolveTriangularInplace, :161-168); otherwise diag_fn(i) = ext_diag[i]
 * (qr.rs:149,176 pass |diag|).  A zero diagonal yields inf/NaN, never an error (:274-276). */
int lfb_solve_triangular_f64(lfb_handle *h, const double *a, int64_t a_rows, int64_t a_cols, int64_t a_rs, int64_t a_cs,
                             double *b, int64_t b_rows, int64_t b_cols, int64_t b_rs, int64_t b_cs,
                             int uplo, const double *ext_diag);
int lfb_solve_triangular_f32(lfb_handle *h, const float *a, int64_t a_rows, int64_t a_cols, int64_t a_rs, int64_t a_cs,
                             float *b, int64_t b_rows, int64_t b_cols, int64_t b_rs, int64_t b_cs,
                             int uplo, const float *ext_diag);

/* triangular.rs:37-53 triangular_inplace (zero the other strict triangle). */
int lfb_triangular_inplace_f64(lfb_handle *h, double *a, int64_t rows, int64_t cols, int64_t rs, int64_t cs, int uplo);
int lfb_triangular_inplace_f32(lfb_handle *h, float *a, int64_t rows, int64_t cols, int64_t rs, int64_t cs, int uplo);

/* ---- fused solve drivers: one upload, every stage on the device, one download -------------------------
 * qr.rs:207-248 LeastSquaresQrInto / LeastSquaresQr: a (rows x cols), b (rows x bcols) -> x (cols x bcols).
 * rows >= cols: qr_into + QRDecomp::solve_into (qr.rs:124-152); rows < cols: QR of the transpose +
 * solve_tr_into (qr.rs:156-181).  LFB_WRONG_ROWS, LFB_NON_INVERTIBLE (a zero on diag(R), qr.rs:194-197). */
int lfb_least_squares_f64(lfb_handle *h, const double *a, int64_t rows, int64_t cols, int64_t rs, int64_t cs,
                          const double *b, int64_t b_rows, int64_t bcols, int64_t b_rs, int64_t b_cs,
                          double *x, int64_t x_rs, int64_t x_cs);
int lfb_least_squares_f32(lfb_handle *h, const float *a, int64_t rows, int64_t cols, int64_t rs, int64_t cs,
                          const float *b, int64_t b_rows, int64_t bcols, int64_t b_rs, int64_t b_cs,
                          float *x, int64_t x_rs, int64_t x_cs);
/* qr.rs:124-152 QRDecomp::solve_into (and :200-203 inverse with b = I) on an existing compact factor + diag. */
int lfb_qr_solve_f64(lfb_handle *h, const double *qr, int64_t rows, int64_t cols, int64_t rs, int64_t cs, const double *diag,
                     const double *b, int64_t b_rows, int64_t bcols, int64_t b_rs, int64_t b_cs,
                     double *x, int64_t x_rs, int64_t x_cs);
int lfb_qr_solve_f32(lfb_handle *h, const float *qr, int64_t rows, int64_t cols, int64_t rs, int64_t cs, const float *diag,
                     const float *b, int64_t b_rows, int64_t bcols, int64_t b_rs, int64_t b_cs,
                     float *x, int64_t x_rs, int64_t x_cs);
/* qr.rs:156-181 QRDecomp::solve_tr_into on an existing compact factor + diag: x (rows x bcols) = Q (R^-T b), b is cols x bcols.
 * The triangular solve (R^T m = b with |diag| as the diagonal, :172-177), generate_q (:180) and the product Q m run on the
 * device in one round trip.  LFB_WRONG_ROWS if b does not have `cols` rows (:160-165), LFB_NON_INVERTIBLE (:166-168). */
int lfb_qr_solve_tr_f64(lfb_handle *h, const double *qr, int64_t rows, int64_t cols, int64_t rs, int64_t cs, const double *diag,
                        const double *b, int64_t b_rows, int64_t bcols, int64_t b_rs, int64_t b_cs,
                        double *x, int64_t x_rs, int64_t x_cs);
int lfb_qr_solve_tr_f32(lfb_handle *h, const float *qr, int64_t rows, int64_t cols, int64_t rs, int64_t cs, const float *diag,
                        const float *b, int64_t b_rows, int64_t bcols, int64_t b_rs, int64_t b_cs,
                        float *x, int64_t x_rs, int64_t x_cs);
/* cholesky.rs:118-163 SolveCInplace / SolveC: b (n x bcols) is overwritten with the solution of A x = b;
 * with write_factor != 0, a receives its Cholesky factor in the lower triangle (solvec_inplace, :136-144).
 * LFB_NOT_POSITIVE_DEFINITE reports the failing pivot in *fail_index. */
int lfb_solvec_f64(lfb_handle *h, double *a, int64_t rows, int64_t cols, int64_t rs, int64_t cs, int write_factor,
                   double *b, int64_t b_rows, int64_t bcols, int64_t b_rs, int64_t b_cs, int64_t *fail_index);
int lfb_solvec_f32(lfb_handle *h, float *a, int64_t rows, int64_t cols, int64_t rs, int64_t cs, int write_factor,
                   float *b, int64_t b_rows, int64_t bcols, int64_t b_rs, int64_t b_cs, int64_t *fail_index);
/* cholesky.rs:166-199 InverseCInplace / InverseC: inv (n x n) = A^-1; the identity is generated on the device. */
int lfb_invc_f64(lfb_handle *h, const double *a, int64_t rows, int64_t cols, int64_t rs, int64_t cs,
                 double *inv, int64_t i_rs, int64_t i_cs, int64_t *fail_index);
int lfb_invc_f32(lfb_handle *h, const float *a, int64_t rows, int64_t cols, int64_t rs, int64_t cs,
                 float *inv, int64_t i_rs, int64_t i_cs, int64_t *fail_index);

/* ---- the dense blocks of LOBPCG, one call each (lobpcg/algorithm.rs) -----------------
 * lobpcg/algorithm.rs:81-97 orthonormalize(v) -> (u, gram_vv_fac): v (rows x cols) is overwritten with
 * u = v L^-T where L = cholesky_into(v^T v) (:82-83; lower, strict upper zeroed) is written to l (cols x cols; may be
 * NULL).  The Gram matrix and its factor never leave the device.  LFB_NOT_POSITIVE_DEFINITE (+ *fail_index, the
 * failing pivot row) when v^T v is not numerically SPD, exactly where cholesky_into fails (:83); v, l untouched. */
int lfb_orthonormalize_f64(lfb_handle *h, double *v, int64_t rows, int64_t cols, int64_t rs, int64_t cs,
                           double *l, int64_t l_rs, int64_t l_cs, int64_t *fail_index);
int lfb_orthonormalize_f32(lfb_handle *h, float *v, int64_t rows, int64_t cols, int64_t rs, int64_t cs,
                           float *l, int64_t l_rs, int64_t l_cs, int64_t *fail_index);
/* lobpcg/algorithm.rs:63-76 apply_constraints(v, cholesky_yy, y): v (n x k) -= y u with u the solution of
 * cholesky_yy u = y^T v (:68-75; cholesky_yy is m x m, its lower triangle is read; y is n x m).  A y that is not
 * n x m is LFB_INVALID_ARGUMENT (ndarray panics on the shape mismatch). */
int lfb_apply_constraints_f64(lfb_handle *h, double *v, int64_t n, int64_t k, int64_t rs, int64_t cs,
                              const double *cholesky_yy, int64_t m, int64_t l_rs, int64_t l_cs,
                              const double *y, int64_t y_rows, int64_t y_cols, int64_t y_rs, int64_t y_cs);
int lfb_apply_constraints_f32(lfb_handle *h, float *v, int64_t n, int64_t k, int64_t rs, int64_t cs,
                              const float *cholesky_yy, int64_t m, int64_t l_rs, int64_t l_cs,
                              const float *y, int64_t y_rows, int64_t y_cols, int64_t y_rs, int64_t y_cs);

/* lobpcg/algorithm.rs:16-44 generalized_eig / sorted_eig: the small dense eigenproblems of every LOBPCG iteration, as ONE call.
 * a: k x k view; b: k x k view or NULL (then the plain problem).  order: 0 = the pair as `eigh_into` / `generalized_eig` return
 * it (vals: k entries, vecs: k x k); 1 = Largest, 2 = Smallest: `sort_eig(order)` (eigh.rs:275-325), columns multiplied by the
 * signum of their first entry (sign BIT, :40-41), truncated to `size` (vals: size entries, vecs: k x size).  Both eigen-
 * decompositions, the k x k products between them, the sort's gather and the sign fix run on the device in one upload /
 * download.  LFB_INVALID_ARGUMENT if an eigenvalue is NaN (the reference's sort panics). */
int lfb_sorted_eig_f64(lfb_handle *h, const double *a, int64_t k, int64_t a_rs, int64_t a_cs, const double *b, int64_t b_rs, int64_t b_cs,
                       int64_t size, int order, double *vals, double *vecs, int64_t v_rs, int64_t v_cs);
int lfb_sorted_eig_f32(lfb_handle *h, const float *a, int64_t k, int64_t a_rs, int64_t a_cs, const float *b, int64_t b_rs, int64_t b_cs,
                       int64_t size, int order, float *vals, float *vecs, int64_t v_rs, int64_t v_cs);
/* The same on device-resident column-major operands (d_a, d_b consumed; d_b may be NULL); vals_host is HOST memory, d_vecs is
 * k x size (k x k for order 0) on the device: with lfb_orthonormalize_dev_f64, lfb_apply_constraints_dev_f64 and lfb_gemm_dev_f64
 * every n x k block of a LOBPCG iteration (algorithm.rs:193-417) stays in HBM -- tests/test_gpu_lobpcg_resident.py runs the loop. */
int lfb_sorted_eig_dev_f64(lfb_handle *h, double *d_a, int64_t lda, double *d_b, int64_t ldb, int64_t k, int64_t size, int order,
                           double *vals_host, double *d_vecs, int64_t ldv);

/* ---- eigh.rs:202-268 EighInto / Eigh / EigValshInto / EigValsh (symmetric_eig, eigh.rs:10-129) ---------
 * a: n x n view (only read; the reference consumes `self`, nothing of it is observable afterwards).
 * vals: n contiguous entries, in the reference's own (unsorted) order -- EigSort stays host-side.
 * vecs: n x n view that receives the eigenvectors as columns, or NULL for eigvalsh (no Q is formed).
 * Scale by max|a|, tridiagonalise and generate Q on the device; the scalar implicit-QR recurrence of
 * eigh.rs:51-128 runs on the host and its Givens rotations are applied to Q on the device, eight sweeps per
 * pass.  Returns LFB_NOT_SQUARE like check_square (lib.rs:64-71); n = 0 is legal. */
int lfb_eigh_f64(lfb_handle *h, const double *a, int64_t rows, int64_t cols, int64_t rs, int64_t cs,
                 double *vals, double *vecs, int64_t vrs, int64_t vcs);
int lfb_eigh_f32(lfb_handle *h, const float *a, int64_t rows, int64_t cols, int64_t rs, int64_t cs,
                 float *vals, float *vecs, int64_t vrs, int64_t vcs);

/* ---- svd.rs:415-479 SVDInto / SVD (svd, svd.rs:17-221) ------------------------------------------------
 * a: rows x cols view (only read; the reference consumes `self`).  sigma: min(rows, cols) contiguous entries in the
 * reference's own (unsorted) order, all >= 0 -- SvdSort stays host-side.  u: rows x min view or NULL (calc_u = false);
 * vt: min x cols view or NULL (calc_vt = false).  Scale by max|a|, bidiagonalise, generate U and V on the device; the
 * scalar implicit-shift QR recurrence of svd.rs:55-208 (eps = 5 * machine epsilon, svd.rs:441) runs on the host and
 * its Givens rotations are applied to U and V on the device.  LFB_EMPTY_MATRIX for an empty input (svd.rs:23-25). */
int lfb_svd_f64(lfb_handle *h, const double *a, int64_t rows, int64_t cols, int64_t rs, int64_t cs, double *sigma,
                double *u, int64_t urs, int64_t ucs, double *vt, int64_t vrs, int64_t vcs);
int lfb_svd_f32(lfb_handle *h, const float *a, int64_t rows, int64_t cols, int64_t rs, int64_t cs, float *sigma,
                float *u, int64_t urs, int64_t ucs, float *vt, int64_t vrs, int64_t vcs);

/* ---- tridiagonal.rs:31-66 sym_tridiagonal ---------------------------------------------------- */
/* In place: diag(a) becomes the tridiagonal's diagonal, a[i+1.., i] the reflectors; off (n-1) gets
 * the SIGNED off-diagonal (TridiagonalDecomp, :71-77).  n == 0 -> LFB_EMPTY_MATRIX. */
int lfb_sym_tridiagonal_f64(lfb_handle *h, double *a, int64_t rows, int64_t cols, int64_t rs, int64_t cs, double *off);
int lfb_sym_tridiagonal_f32(lfb_handle *h, float *a, int64_t rows, int64_t cols, int64_t rs, int64_t cs, float *off);

/* ---- bidiagonal.rs:27-59 bidiagonal ---------------------------------------------------------- */
/* In place: signed diagonal d (min(r,c)) and off-diagonal e (min(r,c)-1); reflectors stay in a
 * (BidiagonalDecomp, :64-69).  Empty -> LFB_EMPTY_MATRIX. */
int lfb_bidiagonal_f64(lfb_handle *h, double *a, int64_t rows, int64_t cols, int64_t rs, int64_t cs, double *d, double *e);
int lfb_bidiagonal_f32(lfb_handle *h, float *a, int64_t rows, int64_t cols, int64_t rs, int64_t cs, float *d, float *e);

/* ---- batched thin QR of `batch` packed row-major m x n matrices (qr.rs:32-44 per matrix) ------ */
/* a: [batch][m][n] contiguous, diag: [batch][n].  Batch-sharded across GPUs by the caller. */
int lfb_qr_batched_f32(lfb_handle *h, float *a, int64_t batch, int64_t m, int64_t n, float *diag);
int lfb_qr_batched_f64(lfb_handle *h, double *a, int64_t batch, int64_t m, int64_t n, double *diag);
/* cholesky.rs:51-83 over `batch` packed row-major n x n matrices (n <= 32), in place (clean != 0 zeroes the strict upper
 * triangles, cholesky.rs:78-82).  LFB_NOT_POSITIVE_DEFINITE if any matrix fails: *fail_matrix / *fail_index name the
 * first one in batch order and its failing row.  Batch-sharded across GPUs like lfb_qr_batched_*. */
int lfb_cholesky_batched_f32(lfb_handle *h, float *a, int64_t batch, int64_t n, int clean, int64_t *fail_matrix, int64_t *fail_index);
int lfb_cholesky_batched_f64(lfb_handle *h, double *a, int64_t batch, int64_t n, int clean, int64_t *fail_matrix, int64_t *fail_index);

/* ---- device-resident entry points (column-major, leading dimension, async on the stream) ----- */
int lfb_qr_dev_f64(lfb_handle *h, double *d_a, int64_t rows, int64_t cols, int64_t ld, double *d_diag);
int lfb_qr_dev_f32(lfb_handle *h, float *d_a, int64_t rows, int64_t cols, int64_t ld, float *d_diag);
/* lower Cholesky of a column-major matrix (== upper^T of the row-major view); d_info[0] = 0 or fail row + 1 */
int lfb_cholesky_dev_f64(lfb_handle *h, double *d_a, int64_t n, int64_t ld, int clean, int64_t *d_info);
int lfb_cholesky_dev_f32(lfb_handle *h, float *d_a, int64_t n, int64_t ld, int clean, int64_t *d_info);
/* q (rows x min(rows,cols), column-major ldq) from a compact factor, as lfb_assemble_q */
int lfb_assemble_q_dev_f64(lfb_handle *h, const double *d_m, int64_t rows, int64_t cols, int64_t ld,
                           int64_t shift, const double *d_signs, double *d_q, int64_t ldq);
int lfb_sym_tridiagonal_dev_f64(lfb_handle *h, double *d_a, int64_t n, int64_t ld, double *d_off);
/* d_a (n x n column-major, consumed) -> vals_host (n, HOST memory), d_q (n x n device, or NULL). */
int lfb_eigh_dev_f64(lfb_handle *h, double *d_a, int64_t n, int64_t ld, double *vals_host, double *d_q, int64_t ldq);
/* d_a (rows x cols column-major, consumed) -> sigma_host (min, HOST memory), d_u (rows x min, or NULL),
 * d_v (cols x min: V = Vt^T column-major, i.e. the row-major buffer of Vt; or NULL). */
int lfb_svd_dev_f64(lfb_handle *h, double *d_a, int64_t rows, int64_t cols, int64_t ld, double *sigma_host,
                    double *d_u, int64_t ldu, double *d_v, int64_t ldv);
int lfb_bidiagonal_dev_f64(lfb_handle *h, double *d_a, int64_t rows, int64_t cols, int64_t ld, double *d_d, double *d_e);
/* batched: d_a is [batch][m][n] row-major packed (the ndarray layout), in place */
int lfb_qr_batched_dev_f32(lfb_handle *h, float *d_a, int64_t batch, int64_t m, int64_t n, float *d_diag);
/* d_fail: device int[batch], first failing row per matrix or -1. */
int lfb_cholesky_batched_dev_f32(lfb_handle *h, float *d_a, int64_t batch, int64_t n, int clean, int *d_fail);
/* Tall-skinny local stage of TSQR: R (cols x cols, column-major ldr, diag >= 0, strict lower zeroed)
 * of a rows x cols column-major block.  d_a is overwritten (compact local factor). */
int lfb_tsqr_local_r_dev_f64(lfb_handle *h, double *d_a, int64_t rows, int64_t cols, int64_t ld, double *d_r, int64_t ldr);
/* Tall-skinny qr_into on the device: d_a (rows x cols) becomes the reference's compact factor, d_diag the signed
 * pivots -- TSQR + Householder reconstruction, same contract as lfb_qr_dev_*. */
int lfb_qr_tsqr_dev_f64(lfb_handle *h, double *d_a, int64_t rows, int64_t cols, int64_t ld, double *d_diag);
int lfb_qr_tsqr_dev_f32(lfb_handle *h, float *d_a, int64_t rows, int64_t cols, int64_t ld, float *d_diag);
/* Its building blocks, for a row-sharded matrix (linfa_linalg_b200/dist.py: tsqr_qr):
 *   explicit_q:       d_a <- explicit thin Q of the block, d_r <- its R (cols x cols, diag >= 0)
 *   apply_q:          d_q <- d_q * d_qs   (d_qs: this rank's cols x cols block of the Q of the stacked R factors)
 *   reconstruct_top:  first n rows of the global Q (on the rank that owns them) -> top block of the compact factor,
 *                     d_u <- U' (n x n) for everybody else's rows, d_diag <- signed pivots
 *   reconstruct_rows: d_q (rows x n, any rows below the top block) <- d_q * U'^-1  == the reflector rows */
int lfb_tsqr_explicit_q_dev_f64(lfb_handle *h, double *d_a, int64_t rows, int64_t cols, int64_t ld, double *d_r, int64_t ldr);
/* The Cholesky-QR leaf on its own (csrc/cholqr.cu): R (cols x cols upper, diag >= 0) and R^-1 of a tall block from its Gram
 * matrix, d_a NOT modified.  *ok = 1 if the leaf accepted the block (Cholesky succeeded and the cond_2(R) bound <= option
 * "tsqr_cholqr_cond"), else 0 and nothing is written: the caller then takes lfb_tsqr_explicit_q_dev_f64.  Synchronises the
 * stream (the decision is made on the host).  With it a row-sharded caller folds everything after the leaf into n x n
 * products: rows <- rows * (R_i^-1 Qs_i U'^-1)  (linfa_linalg_b200/dist.py: tsqr_qr, csrc/multi.cu). */
int lfb_tsqr_leaf_dev_f64(lfb_handle *h, const double *d_a, int64_t rows, int64_t cols, int64_t ld, double *d_r, int64_t ldr,
                          double *d_rinv, int64_t ldri, int *ok);
int lfb_tsqr_apply_q_dev_f64(lfb_handle *h, double *d_q, int64_t rows, int64_t cols, int64_t ld, const double *d_qs, int64_t ldqs);
int lfb_hh_reconstruct_top_dev_f64(lfb_handle *h, double *d_qtop, int64_t n, int64_t ld, const double *d_r, int64_t ldr,
                                   double *d_u, int64_t ldu, double *d_diag);
int lfb_hh_reconstruct_rows_dev_f64(lfb_handle *h, double *d_q, int64_t rows, int64_t n, int64_t ld, const double *d_u, int64_t ldu);
/* LOBPCG blocks on device-resident column-major operands, async on the stream, so the iteration's blocks can stay in
 * HBM across calls: d_info is a device int64 (0, or failing pivot row + 1 -- d_v is then unspecified). */
int lfb_orthonormalize_dev_f64(lfb_handle *h, double *d_v, int64_t rows, int64_t cols, int64_t ld, double *d_l, int64_t ldl,
                               int64_t *d_info);
int lfb_apply_constraints_dev_f64(lfb_handle *h, double *d_v, int64_t n, int64_t k, int64_t ldv, const double *d_cholesky_yy,
                                  int64_t m, int64_t ldl, const double *d_y, int64_t ldy);
/* General column-major GEMM on the engine's own kernels (used by tests / yard-stick benches):
 * C = alpha op(A) op(B) + beta C,  ta/tb: 0 = N, 1 = T. */
int lfb_gemm_dev_f64(lfb_handle *h, int ta, int tb, int64_t m, int64_t n, int64_t k, double alpha,
                     const double *d_a, int64_t lda, const double *d_b, int64_t ldb, double beta, double *d_c, int64_t ldc);
int lfb_gemm_dev_f32(lfb_handle *h, int ta, int tb, int64_t m, int64_t n, int64_t k, float alpha,
                     const float *d_a, int64_t lda, const float *d_b, int64_t ldb, float beta, float *d_c, int64_t ldc);

/* ---- GEMM profiler for the roofline line of bench.py: CUDA events around every FP64 GEMM launch
 * between begin and end; end synchronises and returns summed device time, algorithmic flops, calls. */
int lfb_profile_begin(lfb_handle *h);
int lfb_profile_end(lfb_handle *h, double *gemm_ms, double *gemm_flops, int64_t *gemm_calls);

/* Debug: first call arms per-phase cycle counters in the cluster panel kernel, later calls read the
 * counters of the most recent launch (gather, scalars, row pass, reduce+push+barrier). */
int lfb_debug_panel_phases(lfb_handle *h, long long *out4);

/* ---- micro-benchmarks used by bench.py to measure the FP64 pipe ceiling in the same run ------- */
/* kind: 0 = DFMA register chain, 1 = DMMA.8x8x4 (mma.sync f64), 2 = FP32 FFMA register chain.  Returns achieved GFLOP/s. */
int lfb_microbench_fp64(lfb_handle *h, int kind, double *gflops);
/* Device time (us per launch, CUDA events, back-to-back launches) of one internal kernel on an n x n f64
 * problem: "trd_symv" (lower-triangle SYMV of the tridiagonalisation), "trd_head" (its cluster kernel),
 * "bd_gemv_n" / "bd_gemv_t" (streaming GEMVs of the bidiagonalisation on a 4n x n matrix); "potf2" (the 64 x 64 diagonal
 * block kernel of Cholesky; here n selects the variant: 0 left-looking, 1 right-looking, +2 without the inverse). */
int lfb_microbench_kernel(lfb_handle *h, const char *name, int64_t n, int reps, double *us_per_launch);

/* ==== one box, several GPUs: the two paths that shard naturally behind the same boundary ===================
 * ONE process: lfb_create_multi makes one engine handle and one host worker thread per listed device and, for more than one
 * device, an NCCL communicator over them (ncclCommInitAll; libnccl.so.2 is resolved at run time -- the copy the host process
 * already has mapped, if any -- so the library has no link-time NCCL dependency; LFB_ERR_NCCL if it cannot be had).
 * These are the entry points the Rust shim binds for `QRInto::qr_into` (qr.rs:29-45) on a tall-skinny matrix and for loops
 * over many small matrices; a Rust caller has one process and cannot use bench.py's one-process-per-GPU launch.
 * devices == NULL means devices 0 .. n_devices-1.  Handles are used by one call at a time, like lfb_handle. */
int lfb_create_multi(lfb_multi **out, const int *devices, int n_devices);
int lfb_destroy_multi(lfb_multi *m);
const char *lfb_multi_last_error(lfb_multi *m);
int lfb_multi_device_count(lfb_multi *m);
int lfb_multi_nccl_ranks(lfb_multi *m);     /* size of the NCCL communicator (0 with a single device: no collective exists) */
int lfb_multi_nccl_version(lfb_multi *m);   /* ncclGetVersion code of the library in use, 0 if none */
lfb_handle *lfb_multi_handle(lfb_multi *m, int i);            /* device i's engine handle (options, launch counts) */
int lfb_multi_set_option(lfb_multi *m, const char *key, int64_t value);   /* lfb_set_option on every device's handle */
int64_t lfb_multi_launch_count(lfb_multi *m);                 /* kernels + collectives launched on all devices so far */
int lfb_multi_synchronize(lfb_multi *m);
/* Device-side timing of whatever is enqueued between the two calls: start/stop events on every device's stream, result =
 * MAX over devices of the elapsed time (ms). */
int lfb_multi_time_begin(lfb_multi *m);
int lfb_multi_time_end(lfb_multi *m, double *max_ms);

/* qr.rs:29-45 QRInto::qr_into on a tall-skinny host view, ROWS SHARDED over the devices (device i owns the contiguous row
 * block [i*rows/G ...), first rows%G blocks one row longer): same contract and results as lfb_qr_* / lfb_qr_tsqr_* -- the
 * reference's compact factor written back in place, diag = signed pivots.  Per device: explicit-Q TSQR of its rows; ONE
 * ncclAllGather of the n x n R factors; QR of the stacked R (replicated); Householder reconstruction of the top block on
 * device 0; ONE ncclBroadcast of U' and diag; one right-hand TRSM per device.  No row data ever crosses NVLink.
 * A matrix with fewer than G*cols rows is factored on device 0 alone. */
int lfb_qr_tsqr_multi_f64(lfb_multi *m, double *a, int64_t rows, int64_t cols, int64_t rs, int64_t cs, double *diag);
int lfb_qr_tsqr_multi_f32(lfb_multi *m, float *a, int64_t rows, int64_t cols, int64_t rs, int64_t cs, float *diag);
/* R only (qr.rs:91-98 into_r of the same factorisation; a is read, not written): r is cols x cols, upper, diag >= 0. */
int lfb_tsqr_r_multi_f64(lfb_multi *m, const double *a, int64_t rows, int64_t cols, int64_t rs, int64_t cs, double *r, int64_t r_rs, int64_t r_cs);
int lfb_tsqr_r_multi_f32(lfb_multi *m, const float *a, int64_t rows, int64_t cols, int64_t rs, int64_t cs, float *r, int64_t r_rs, int64_t r_cs);
/* Device-resident variants: d_blocks[i] is device i's rows[i] x cols column-major block (leading dimension ld[i], rows[i] >=
 * cols), overwritten with its rows of the compact factor (resp. destroyed for the R-only call); d_diag[i] (cols values) and
 * d_r[i] (cols x cols column-major, ld = cols) receive the replicated results on every device whose pointer is not NULL
 * (either array may itself be NULL).  Asynchronous on the handles' streams: finish with lfb_multi_synchronize. */
int lfb_qr_tsqr_multi_dev_f64(lfb_multi *m, double *const *d_blocks, const int64_t *rows, int64_t cols, const int64_t *ld,
                              double *const *d_diag, double *const *d_r);
int lfb_tsqr_r_multi_dev_f64(lfb_multi *m, double *const *d_blocks, const int64_t *rows, int64_t cols, const int64_t *ld, double *const *d_r);
/* Batched small factorisations, BATCH SHARDED (same split rule), no collective: contracts of lfb_qr_batched_* and
 * lfb_cholesky_batched_* (the first failing matrix is reported in batch order, whichever device met it). */
int lfb_qr_batched_multi_f32(lfb_multi *m, float *a, int64_t batch, int64_t mm, int64_t n, float *diag);
int lfb_qr_batched_multi_f64(lfb_multi *m, double *a, int64_t batch, int64_t mm, int64_t n, double *diag);
int lfb_cholesky_batched_multi_f32(lfb_multi *m, float *a, int64_t batch, int64_t n, int clean, int64_t *fail_matrix, int64_t *fail_index);
int lfb_cholesky_batched_multi_f64(lfb_multi *m, double *a, int64_t batch, int64_t n, int clean, int64_t *fail_matrix, int64_t *fail_index);
/* d_a[i]: device i's [batch[i]][mm][n] packed shard, d_diag[i]: [batch[i]][n]; asynchronous on the handles' streams. */
int lfb_qr_batched_multi_dev_f32(lfb_multi *m, float *const *d_a, const int64_t *batch, int64_t mm, int64_t n, float *const *d_diag);

#ifdef __cplusplus
}
#endif
#endif /* LINFA_B200_H */
