use std::{ffi::c_void, ops::Range, ptr::NonNull};

use crate::backends::common::{Buffer, DenseBuffer, SparseBuffer};

use super::{error::check, ffi, Cuda, CudaContext, CudaError};

/// buffer/mod.rs:11-17 + buffer/dense.rs:5-7
#[derive(Debug)]
pub struct CudaBuffer {
    raw: NonNull<ffi::uzu_buffer>,
}
unsafe impl Send for CudaBuffer {}
unsafe impl Sync for CudaBuffer {}

impl CudaBuffer {
    pub(crate) fn new(ctx: &CudaContext, size: usize, kind: i32) -> Result<Self, CudaError> {
        let mut raw = std::ptr::null_mut();
        check(unsafe { ffi::uzu_buffer_create(ctx.raw.as_ptr(), size, kind, &mut raw) })?;
        Ok(Self { raw: NonNull::new(raw).unwrap() })
    }

    /// After the loader has written the weights through `cpu_ptr`: prefetch the managed pages to the device once.
    pub fn make_resident(&self, ctx: &CudaContext) -> Result<(), CudaError> {
        check(unsafe { ffi::uzu_buffer_make_resident(ctx.raw.as_ptr(), self.raw.as_ptr()) })
    }
}

impl Buffer for CudaBuffer {
    type Backend = Cuda;
    fn gpu_ptr(&self) -> usize {
        unsafe { ffi::uzu_buffer_gpu_ptr(self.raw.as_ptr()) as usize }
    }
    fn size(&self) -> usize {
        unsafe { ffi::uzu_buffer_size(self.raw.as_ptr()) }
    }
}

impl DenseBuffer for CudaBuffer {
    fn cpu_ptr(&self) -> NonNull<c_void> {
        NonNull::new(unsafe { ffi::uzu_buffer_cpu_ptr(self.raw.as_ptr()) }).expect("managed / pinned buffers are CPU addressable")
    }
}

impl Drop for CudaBuffer {
    fn drop(&mut self) {
        unsafe { ffi::uzu_buffer_destroy(self.raw.as_ptr()) }
    }
}

/// buffer/sparse.rs:5-19 — CUDA virtual memory management: one reserved VA range, physical pages mapped on demand
/// (the KV cache, mixer/attention/state.rs:113-172).
#[derive(Debug)]
pub struct CudaSparseBuffer {
    raw: NonNull<ffi::uzu_sparse_buffer>,
}
unsafe impl Send for CudaSparseBuffer {}
unsafe impl Sync for CudaSparseBuffer {}

impl CudaSparseBuffer {
    pub(crate) fn new(ctx: &CudaContext, capacity: usize) -> Result<Self, CudaError> {
        let mut raw = std::ptr::null_mut();
        check(unsafe { ffi::uzu_sparse_buffer_create(ctx.raw.as_ptr(), capacity, &mut raw) })?;
        Ok(Self { raw: NonNull::new(raw).unwrap() })
    }
    fn page_list(pages: &Range<usize>) -> Vec<u32> {
        pages.clone().map(|p| p as u32).collect()
    }
}

impl Buffer for CudaSparseBuffer {
    type Backend = Cuda;
    fn gpu_ptr(&self) -> usize {
        unsafe { ffi::uzu_sparse_buffer_gpu_ptr(self.raw.as_ptr()) as usize }
    }
    fn size(&self) -> usize {
        unsafe { ffi::uzu_sparse_buffer_size(self.raw.as_ptr()) }
    }
}

impl SparseBuffer for CudaSparseBuffer {
    fn map(&mut self, _context: &CudaContext, pages: &Range<usize>) -> Result<(), CudaError> {
        let list = Self::page_list(pages);
        check(unsafe { ffi::uzu_sparse_buffer_map(self.raw.as_ptr(), list.as_ptr(), list.len()) })
    }
    fn unmap(&mut self, _context: &CudaContext, pages: &Range<usize>) -> Result<(), CudaError> {
        let list = Self::page_list(pages);
        check(unsafe { ffi::uzu_sparse_buffer_unmap(self.raw.as_ptr(), list.as_ptr(), list.len()) })
    }
    fn page_size_bytes(&self) -> usize {
        unsafe { ffi::uzu_sparse_buffer_page_size_bytes(self.raw.as_ptr()) }
    }
}

impl Drop for CudaSparseBuffer {
    fn drop(&mut self) {
        unsafe { ffi::uzu_sparse_buffer_destroy(self.raw.as_ptr()) }
    }
}
