use std::{ffi::CString, path::Path, ptr::NonNull, sync::Arc};

use crate::backends::common::{Allocation, AllocationPool, AllocationType, Allocator, Context, DeviceCapabilities};

use super::{error::check, ffi, Cuda, CudaBuffer, CudaError, CudaInitial, CudaSparseBuffer};

/// backends/common/context.rs:5-48. One CUDA stream per context: submission order = execution order, like the CPU backend's
/// single worker thread (cpu/context.rs:21-27).
pub struct CudaContext {
    pub(crate) raw: NonNull<ffi::uzu_context>,
    allocator: Arc<Allocator<Cuda>>,
}

// The library serialises on the context's stream; handles are plain pointers.
unsafe impl Send for CudaContext {}
unsafe impl Sync for CudaContext {}

impl Context for CudaContext {
    type Backend = Cuda;

    fn new() -> Result<Arc<Self>, CudaError> {
        ffi::abi_self_check();      // a glue layer compiled against another header revision must not run
        let mut raw = std::ptr::null_mut();
        // -1: CUDA current device or $UZU_DEVICE (one process per GPU for tensor parallel runs)
        check(unsafe { ffi::uzu_context_create(-1, &mut raw) })?;
        let raw = NonNull::new(raw).expect("uzu_context_create returned OK with a null context");
        Ok(Arc::new_cyclic(|weak| CudaContext { raw, allocator: Allocator::new(weak.clone()) }))
    }

    fn create_command_buffer(&self, name: Option<&str>) -> Result<CudaInitial, CudaError> {
        let cname = CString::new(name.unwrap_or("")).unwrap();
        let mut raw = std::ptr::null_mut();
        check(unsafe { ffi::uzu_command_buffer_create(self.raw.as_ptr(), cname.as_ptr(), &mut raw) })?;
        Ok(CudaInitial::from_raw(raw))
    }

    /// DenseBuffer::cpu_ptr must work (the loader preads weights through it, parameters/loader.rs:162-179): managed memory with
    /// preferred location = device; `CudaBuffer::make_resident` migrates the pages to HBM once after the load.
    fn create_buffer(&self, size: usize) -> Result<CudaBuffer, CudaError> {
        CudaBuffer::new(self, size, ffi::UZU_BUFFER_MANAGED as i32)
    }

    fn create_allocation(&self, size: usize, allocation_type: AllocationType<Cuda>) -> Result<Allocation<Cuda>, CudaError> {
        self.allocator.allocate(size, allocation_type)
    }

    fn create_allocation_pool(&self, reusable: bool) -> AllocationPool<Cuda> {
        self.allocator.create_pool(reusable)
    }

    fn create_sparse_buffer(&self, capacity: usize) -> Result<CudaSparseBuffer, CudaError> {
        CudaSparseBuffer::new(self, capacity)
    }

    fn peak_memory_usage(&self) -> Option<usize> {
        let mut bytes = 0usize;
        (unsafe { ffi::uzu_context_peak_memory_usage(self.raw.as_ptr(), &mut bytes) } == 0).then_some(bytes)
    }

    fn enable_capture() {}

    fn start_capture(&self, trace_path: &Path) -> Result<(), CudaError> {
        let p = CString::new(trace_path.to_string_lossy().as_bytes()).unwrap();
        check(unsafe { ffi::uzu_context_start_capture(self.raw.as_ptr(), p.as_ptr()) })   // cudaProfilerStart (ncu --profile-from-start off)
    }

    fn stop_capture(&self) -> Result<(), CudaError> {
        check(unsafe { ffi::uzu_context_stop_capture(self.raw.as_ptr()) })
    }

    fn device_capabilities(&self) -> DeviceCapabilities {
        let bits = unsafe { ffi::uzu_context_device_capabilities(self.raw.as_ptr()) };
        DeviceCapabilities::from_bits_truncate(bits)        // bit 0 = SPARSE_BUFFERS (CUDA VMM)
    }
}

impl Drop for CudaContext {
    fn drop(&mut self) {
        unsafe { ffi::uzu_context_destroy(self.raw.as_ptr()) }
    }
}
