//! The `extern "C"` layer over include/uzu_b200.h is GENERATED (ffi_generated.rs, tools/gen_rust_ffi.py); this module re-exports it under the
//! name the rest of the backend uses and keeps the start-up ABI check.
pub use super::ffi_generated::*;

/// Call once at start-up (CudaContext::new does): a binding compiled against a different header revision must not run. Compares
/// `size_of` of every generated `#[repr(C)]` struct with what the loaded library reports (uzu_abi_struct_size, runtime.cu).
pub fn abi_self_check() {
    for (name, size) in super::ffi_generated::STRUCT_SIZES {
        let c = std::ffi::CString::new(*name).unwrap();
        let lib = unsafe { uzu_abi_struct_size(c.as_ptr()) };
        assert_eq!(lib, *size, "ABI mismatch: {name}");
    }
}
