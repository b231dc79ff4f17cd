//! CUDA (B200, sm_100a) backend: thin Rust wrappers over the C ABI of libuzu_b200.so (include/uzu_b200.h).
mod backend;
mod buffer;
mod command_buffer;
mod context;
mod error;
pub mod ffi;
pub mod ffi_generated;
pub mod kernel;

pub use backend::Cuda;
pub use buffer::{CudaBuffer, CudaSparseBuffer};
pub use command_buffer::{CudaCommandBuffer, CudaCompleted, CudaEncoding, CudaExecutable, CudaInitial, CudaPending};
pub use context::CudaContext;
pub use error::CudaError;
