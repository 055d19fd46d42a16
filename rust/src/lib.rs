//! Replaced bodies of linfa-linalg's hot-path trait methods over liblinfa_b200.so.
//!
//! Every public trait, struct and signature of the reference stays as it is (`QRInto`, `QRDecomp`, `CholeskyInplace`,
//! `SolveTriangularInplace`, `SymmetricTridiagonal`, `Bidiagonal`, `EighInto`, `SVDInto`, ...); only the bodies named in
//! INTEGRATION.md section 2 change, and they all have the shape shown here for `qr_into`, `cholesky_inplace_dirty` and
//! `solve_triangular_system`: shape checks in Rust (before any device work, as in the reference), one FFI call per
//! factorisation, status code -> `LinalgError`.  There is no CPU fallback: without a CUDA device `handle()` panics.
//!
//! This crate is NOT compiled in the repository's image (no Rust toolchain there): it is the source a maintainer adds.
mod ffi;

use ndarray::{ArrayBase, DataMut, Ix2, NdFloat};
use std::any::TypeId;
use std::os::raw::c_int;

#[derive(Debug, thiserror::Error)]
#[non_exhaustive]
pub enum LinalgError {
    // the reference's variants, src/lib.rs:33-60
    #[error("Matrix of ({rows}, {cols}) is not square")]
    NotSquare { rows: usize, cols: usize },
    #[error("Expected matrix rows({rows}) >= cols({cols})")]
    NotThin { rows: usize, cols: usize },
    #[error("Matrix is not positive definite")]
    NotPositiveDefinite,
    #[error("Matrix is non-invertible")]
    NonInvertible,
    #[error("Matrix is empty")]
    EmptyMatrix,
    #[error("Matrix must have {expected} rows, not {actual}")]
    WrongRows { expected: usize, actual: usize },
    // new: CUDA / allocation / NCCL failures of the engine (the enum is #[non_exhaustive], src/lib.rs:34)
    #[error("linfa_b200 device error {code}: {message}")]
    Device { code: i32, message: String },
}
pub type Result<T> = std::result::Result<T, LinalgError>;

thread_local! {
    /// One engine handle (device 0, own stream, workspace pool) per thread: a handle serves one call at a time.
    static HANDLE: *mut ffi::lfb_handle = unsafe {
        let mut h = std::ptr::null_mut();
        assert_eq!(ffi::lfb_create(&mut h, 0), ffi::LFB_OK, "no CUDA device: linfa_b200 has no CPU fallback");
        h
    };
}
fn handle() -> *mut ffi::lfb_handle {
    HANDLE.with(|h| *h)
}

/// `A: NdFloat` is exactly f32 | f64 and `'static`, so the scalar family is chosen by TypeId and every public bound of
/// the reference (`A: NdFloat`) stays as it is.
fn is_f64<A: 'static>() -> bool {
    TypeId::of::<A>() == TypeId::of::<f64>()
}

/// Status code of the C ABI -> the reference's error variants (shape arguments are the ones the Rust side checked).
fn status(code: c_int, rows: usize, cols: usize, other_rows: usize) -> Result<()> {
    match code {
        ffi::LFB_OK => Ok(()),
        ffi::LFB_NOT_POSITIVE_DEFINITE => Err(LinalgError::NotPositiveDefinite),
        ffi::LFB_NOT_THIN => Err(LinalgError::NotThin { rows, cols }),
        ffi::LFB_NOT_SQUARE => Err(LinalgError::NotSquare { rows, cols }),
        ffi::LFB_EMPTY_MATRIX => Err(LinalgError::EmptyMatrix),
        ffi::LFB_WRONG_ROWS => Err(LinalgError::WrongRows { expected: rows, actual: other_rows }),
        ffi::LFB_NON_INVERTIBLE => Err(LinalgError::NonInvertible),
        c => {
            let msg = unsafe { std::ffi::CStr::from_ptr(ffi::lfb_last_error(handle())) }.to_string_lossy().into_owned();
            Err(LinalgError::Device { code: c, message: msg })
        }
    }
}

/// Body of `QRInto::qr_into` (src/qr.rs:29-45): returns `diag`; the caller builds `QRDecomp { qr: self, diag }` (:43).
pub fn qr_into_body<A: NdFloat, S: DataMut<Elem = A>>(a: &mut ArrayBase<S, Ix2>) -> Result<ndarray::Array1<A>> {
    let (rows, cols) = a.dim();
    if rows < cols {
        return Err(LinalgError::NotThin { rows, cols }); // src/qr.rs:34-36, before any device work
    }
    let mut diag = ndarray::Array1::<A>::zeros(cols);
    let (rs, cs) = (a.strides()[0] as i64, a.strides()[1] as i64);
    let code = unsafe {
        if is_f64::<A>() {
            ffi::lfb_qr_f64(handle(), a.as_mut_ptr() as *mut f64, rows as i64, cols as i64, rs, cs, diag.as_mut_ptr() as *mut f64)
        } else {
            ffi::lfb_qr_f32(handle(), a.as_mut_ptr() as *mut f32, rows as i64, cols as i64, rs, cs, diag.as_mut_ptr() as *mut f32)
        }
    };
    status(code, rows, cols, 0)?;
    Ok(diag)
}

/// Body of `CholeskyInplace::cholesky_inplace_dirty` / `cholesky_inplace` (src/cholesky.rs:51-83).
pub fn cholesky_inplace_body<A: NdFloat, S: DataMut<Elem = A>>(a: &mut ArrayBase<S, Ix2>, clean: bool) -> Result<()> {
    let (rows, cols) = a.dim();
    if rows != cols {
        return Err(LinalgError::NotSquare { rows, cols }); // check_square, src/lib.rs:64-71
    }
    let (rs, cs) = (a.strides()[0] as i64, a.strides()[1] as i64);
    let mut fail: i64 = -1;
    let code = unsafe {
        if is_f64::<A>() {
            ffi::lfb_cholesky_f64(handle(), a.as_mut_ptr() as *mut f64, rows as i64, cols as i64, rs, cs, clean as c_int, &mut fail)
        } else {
            ffi::lfb_cholesky_f32(handle(), a.as_mut_ptr() as *mut f32, rows as i64, cols as i64, rs, cs, clean as c_int, &mut fail)
        }
    };
    status(code, rows, cols, 0)
}

/// Body of `triangular::solve_triangular_system` (src/triangular.rs:95-144): `ext_diag` is `Some(|diag|)` at the two QR
/// call sites (src/qr.rs:149,176) and `None` for `SolveTriangularInplace` (the diagonal of `a` itself).
pub fn solve_triangular_body<A: NdFloat, Sa: ndarray::Data<Elem = A>, Sb: DataMut<Elem = A>>(
    a: &ArrayBase<Sa, Ix2>, b: &mut ArrayBase<Sb, Ix2>, lower: bool, ext_diag: Option<&[A]>,
) -> Result<()> {
    let (rows, cols) = a.dim();
    if rows != cols {
        return Err(LinalgError::NotSquare { rows, cols });
    }
    if b.nrows() != rows {
        return Err(LinalgError::WrongRows { expected: rows, actual: b.nrows() }); // src/triangular.rs:103-108
    }
    let uplo = if lower { ffi::LFB_LOWER } else { ffi::LFB_UPPER };
    let dp = ext_diag.map(|d| d.as_ptr()).unwrap_or(std::ptr::null());
    let code = unsafe {
        if is_f64::<A>() {
            ffi::lfb_solve_triangular_f64(handle(), a.as_ptr() as *const f64, rows as i64, cols as i64, a.strides()[0] as i64,
                a.strides()[1] as i64, b.as_mut_ptr() as *mut f64, b.nrows() as i64, b.ncols() as i64, b.strides()[0] as i64,
                b.strides()[1] as i64, uplo, dp as *const f64)
        } else {
            ffi::lfb_solve_triangular_f32(handle(), a.as_ptr() as *const f32, rows as i64, cols as i64, a.strides()[0] as i64,
                a.strides()[1] as i64, b.as_mut_ptr() as *mut f32, b.nrows() as i64, b.ncols() as i64, b.strides()[0] as i64,
                b.strides()[1] as i64, uplo, dp as *const f32)
        }
    };
    status(code, rows, cols, b.nrows())
}

/// Several GPUs of one box behind the same call: `qr_into` of a tall-skinny matrix, rows sharded over `devices`
/// (include/linfa_b200.h: lfb_qr_tsqr_multi_*).  One `MultiEngine` per process; NCCL is set up inside the library.
pub struct MultiEngine(*mut ffi::lfb_multi);
impl MultiEngine {
    pub fn new(devices: &[i32]) -> Result<Self> {
        let mut m = std::ptr::null_mut();
        let code = unsafe { ffi::lfb_create_multi(&mut m, devices.as_ptr(), devices.len() as c_int) };
        if code != ffi::LFB_OK {
            return Err(LinalgError::Device { code, message: "lfb_create_multi failed (100 = CUDA, 102 = NCCL)".into() });
        }
        Ok(MultiEngine(m))
    }
    pub fn qr_into_body(&self, a: &mut ndarray::ArrayViewMut2<f64>) -> Result<ndarray::Array1<f64>> {
        let (rows, cols) = a.dim();
        if rows < cols {
            return Err(LinalgError::NotThin { rows, cols });
        }
        let mut diag = ndarray::Array1::<f64>::zeros(cols);
        let code = unsafe {
            ffi::lfb_qr_tsqr_multi_f64(self.0, a.as_mut_ptr(), rows as i64, cols as i64, a.strides()[0] as i64,
                                       a.strides()[1] as i64, diag.as_mut_ptr())
        };
        if code != ffi::LFB_OK {
            let msg = unsafe { std::ffi::CStr::from_ptr(ffi::lfb_multi_last_error(self.0)) }.to_string_lossy().into_owned();
            return Err(LinalgError::Device { code, message: msg });
        }
        Ok(diag)
    }
}
impl Drop for MultiEngine {
    fn drop(&mut self) {
        unsafe { ffi::lfb_destroy_multi(self.0) };
    }
}
