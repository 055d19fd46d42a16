//! `extern "C"` declarations of include/linfa_b200.h (every entry point the replaced trait bodies call).
//! Strides are in ELEMENTS and signed, exactly what `ArrayBase::strides()` returns; pointers are `as_mut_ptr()` of the
//! view (element [0, 0]).
#![allow(non_camel_case_types, dead_code)]
use std::os::raw::{c_char, c_int};

#[repr(C)]
pub struct lfb_handle {
    _p: [u8; 0],
}
#[repr(C)]
pub struct lfb_multi {
    _p: [u8; 0],
}

pub const LFB_OK: c_int = 0;
pub const LFB_NOT_POSITIVE_DEFINITE: c_int = 1;
pub const LFB_NOT_THIN: c_int = 2;
pub const LFB_NOT_SQUARE: c_int = 3;
pub const LFB_EMPTY_MATRIX: c_int = 4;
pub const LFB_WRONG_ROWS: c_int = 5;
pub const LFB_NON_INVERTIBLE: c_int = 6;
pub const LFB_UPPER: c_int = 0;
pub const LFB_LOWER: c_int = 1;

macro_rules! typed_entries {
    ($t:ty, $qr:ident, $qr_tsqr:ident, $assemble_q:ident, $qt_mul:ident, $cholesky:ident, $solve_tri:ident, $tri_inplace:ident,
     $tridiag:ident, $bidiag:ident, $eigh:ident, $svd:ident, $lsq:ident, $qr_solve:ident, $solvec:ident, $invc:ident,
     $ortho:ident, $constraints:ident, $qr_batched:ident, $chol_batched:ident, $qr_tsqr_multi:ident, $tsqr_r_multi:ident,
     $qr_batched_multi:ident, $chol_batched_multi:ident) => {
        extern "C" {
            pub fn $qr(h: *mut lfb_handle, a: *mut $t, rows: i64, cols: i64, rs: i64, cs: i64, diag: *mut $t) -> c_int;
            pub fn $qr_tsqr(h: *mut lfb_handle, a: *mut $t, rows: i64, cols: i64, rs: i64, cs: i64, diag: *mut $t) -> c_int;
            pub fn $assemble_q(h: *mut lfb_handle, m: *const $t, rows: i64, cols: i64, rs: i64, cs: i64, shift: i64,
                               signs: *const $t, q: *mut $t, q_rs: i64, q_cs: i64) -> c_int;
            pub fn $qt_mul(h: *mut lfb_handle, qr: *const $t, rows: i64, cols: i64, rs: i64, cs: i64, diag: *const $t,
                           b: *mut $t, bcols: i64, b_rs: i64, b_cs: i64) -> c_int;
            pub fn $cholesky(h: *mut lfb_handle, a: *mut $t, rows: i64, cols: i64, rs: i64, cs: i64, clean: c_int,
                             fail_index: *mut i64) -> c_int;
            pub fn $solve_tri(h: *mut lfb_handle, a: *const $t, ar: i64, ac: i64, ars: i64, acs: i64, b: *mut $t, br: i64,
                              bc: i64, brs: i64, bcs: i64, uplo: c_int, ext_diag: *const $t) -> c_int;
            pub fn $tri_inplace(h: *mut lfb_handle, a: *mut $t, rows: i64, cols: i64, rs: i64, cs: i64, uplo: c_int) -> c_int;
            pub fn $tridiag(h: *mut lfb_handle, a: *mut $t, rows: i64, cols: i64, rs: i64, cs: i64, off: *mut $t) -> c_int;
            pub fn $bidiag(h: *mut lfb_handle, a: *mut $t, rows: i64, cols: i64, rs: i64, cs: i64, d: *mut $t, e: *mut $t) -> c_int;
            pub fn $eigh(h: *mut lfb_handle, a: *const $t, rows: i64, cols: i64, rs: i64, cs: i64, vals: *mut $t,
                         vecs: *mut $t, vrs: i64, vcs: i64) -> c_int;
            pub fn $svd(h: *mut lfb_handle, a: *const $t, rows: i64, cols: i64, rs: i64, cs: i64, sigma: *mut $t, u: *mut $t,
                        urs: i64, ucs: i64, vt: *mut $t, vrs: i64, vcs: i64) -> c_int;
            pub fn $lsq(h: *mut lfb_handle, a: *const $t, rows: i64, cols: i64, rs: i64, cs: i64, b: *const $t, br: i64,
                        bc: i64, brs: i64, bcs: i64, x: *mut $t, xrs: i64, xcs: i64) -> c_int;
            pub fn $qr_solve(h: *mut lfb_handle, qr: *const $t, rows: i64, cols: i64, rs: i64, cs: i64, diag: *const $t,
                             b: *const $t, br: i64, bc: i64, brs: i64, bcs: i64, x: *mut $t, xrs: i64, xcs: i64) -> c_int;
            pub fn $solvec(h: *mut lfb_handle, a: *mut $t, rows: i64, cols: i64, rs: i64, cs: i64, write_factor: c_int,
                           b: *mut $t, br: i64, bc: i64, brs: i64, bcs: i64, fail_index: *mut i64) -> c_int;
            pub fn $invc(h: *mut lfb_handle, a: *const $t, rows: i64, cols: i64, rs: i64, cs: i64, inv: *mut $t, irs: i64,
                         ics: i64, fail_index: *mut i64) -> c_int;
            pub fn $ortho(h: *mut lfb_handle, v: *mut $t, rows: i64, cols: i64, rs: i64, cs: i64, l: *mut $t, lrs: i64,
                          lcs: i64, fail_index: *mut i64) -> c_int;
            pub fn $constraints(h: *mut lfb_handle, v: *mut $t, n: i64, k: i64, rs: i64, cs: i64, cholesky_yy: *const $t,
                                m: i64, lrs: i64, lcs: i64, y: *const $t, yr: i64, yc: i64, yrs: i64, ycs: i64) -> c_int;
            pub fn $qr_batched(h: *mut lfb_handle, a: *mut $t, batch: i64, m: i64, n: i64, diag: *mut $t) -> c_int;
            pub fn $chol_batched(h: *mut lfb_handle, a: *mut $t, batch: i64, n: i64, clean: c_int, fail_matrix: *mut i64,
                                 fail_index: *mut i64) -> c_int;
            // one box, several GPUs (one process): rows / batch sharded inside the library, NCCL for the R factors
            pub fn $qr_tsqr_multi(m: *mut lfb_multi, a: *mut $t, rows: i64, cols: i64, rs: i64, cs: i64, diag: *mut $t) -> c_int;
            pub fn $tsqr_r_multi(m: *mut lfb_multi, a: *const $t, rows: i64, cols: i64, rs: i64, cs: i64, r: *mut $t,
                                 r_rs: i64, r_cs: i64) -> c_int;
            pub fn $qr_batched_multi(m: *mut lfb_multi, a: *mut $t, batch: i64, mm: i64, n: i64, diag: *mut $t) -> c_int;
            pub fn $chol_batched_multi(m: *mut lfb_multi, a: *mut $t, batch: i64, n: i64, clean: c_int,
                                       fail_matrix: *mut i64, fail_index: *mut i64) -> c_int;
        }
    };
}

typed_entries!(f64, lfb_qr_f64, lfb_qr_tsqr_f64, lfb_assemble_q_f64, lfb_qt_mul_f64, lfb_cholesky_f64, lfb_solve_triangular_f64,
               lfb_triangular_inplace_f64, lfb_sym_tridiagonal_f64, lfb_bidiagonal_f64, lfb_eigh_f64, lfb_svd_f64,
               lfb_least_squares_f64, lfb_qr_solve_f64, lfb_solvec_f64, lfb_invc_f64, lfb_orthonormalize_f64,
               lfb_apply_constraints_f64, lfb_qr_batched_f64, lfb_cholesky_batched_f64, lfb_qr_tsqr_multi_f64,
               lfb_tsqr_r_multi_f64, lfb_qr_batched_multi_f64, lfb_cholesky_batched_multi_f64);
typed_entries!(f32, lfb_qr_f32, lfb_qr_tsqr_f32, lfb_assemble_q_f32, lfb_qt_mul_f32, lfb_cholesky_f32, lfb_solve_triangular_f32,
               lfb_triangular_inplace_f32, lfb_sym_tridiagonal_f32, lfb_bidiagonal_f32, lfb_eigh_f32, lfb_svd_f32,
               lfb_least_squares_f32, lfb_qr_solve_f32, lfb_solvec_f32, lfb_invc_f32, lfb_orthonormalize_f32,
               lfb_apply_constraints_f32, lfb_qr_batched_f32, lfb_cholesky_batched_f32, lfb_qr_tsqr_multi_f32,
               lfb_tsqr_r_multi_f32, lfb_qr_batched_multi_f32, lfb_cholesky_batched_multi_f32);

extern "C" {
    pub fn lfb_create(out: *mut *mut lfb_handle, device: c_int) -> c_int;
    pub fn lfb_destroy(h: *mut lfb_handle) -> c_int;
    pub fn lfb_last_error(h: *mut lfb_handle) -> *const c_char;
    pub fn lfb_set_option(h: *mut lfb_handle, key: *const c_char, value: i64) -> c_int;
    pub fn lfb_synchronize(h: *mut lfb_handle) -> c_int;
    pub fn lfb_version() -> *const c_char;

    pub fn lfb_create_multi(out: *mut *mut lfb_multi, devices: *const c_int, n_devices: c_int) -> c_int;
    pub fn lfb_destroy_multi(m: *mut lfb_multi) -> c_int;
    pub fn lfb_multi_last_error(m: *mut lfb_multi) -> *const c_char;
    pub fn lfb_multi_device_count(m: *mut lfb_multi) -> c_int;
    pub fn lfb_multi_nccl_ranks(m: *mut lfb_multi) -> c_int;
    pub fn lfb_multi_set_option(m: *mut lfb_multi, key: *const c_char, value: i64) -> c_int;
    pub fn lfb_multi_synchronize(m: *mut lfb_multi) -> c_int;
}
