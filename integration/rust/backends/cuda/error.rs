use std::ffi::CStr;

use thiserror::Error;

use super::ffi;

/// `Backend::Error` (backends/common/backend.rs:11). Status codes are `uzu_status` of include/uzu_b200.h; the message is the
/// library's thread-local `uzu_last_error()`.
#[derive(Debug, Error)]
pub enum CudaError {
    #[error("CUDA backend: {0}")]
    Library(String),
    #[error("not supported by the CUDA backend: {0}")]
    NotSupported(&'static str),
}

pub(crate) fn check(status: i32) -> Result<(), CudaError> {
    if status == 0 {
        return Ok(());
    }
    let msg = unsafe { CStr::from_ptr(ffi::uzu_last_error()) }.to_string_lossy().into_owned();
    Err(CudaError::Library(format!("status {status}: {msg}")))
}
