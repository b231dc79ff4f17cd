use std::{ffi::CString, time::Duration};

use crate::backends::common::{
    AccessFlags, Buffer, BufferGpuAddressRangeExt, BufferRangeMut, BufferRangeRef, CommandBuffer, CommandBufferCompleted, CommandBufferEncoding,
    CommandBufferExecutable, CommandBufferInitial, CommandBufferPending,
};

use super::{error::check, ffi, Cuda, CudaError};

/// backends/common/command_buffer.rs:15-125 as a typestate over ONE `uzu_command_buffer`: Encoding = kernels are enqueued on the
/// context's stream as they are encoded (encode never blocks); submit records the end event; wait_until_completed is the only blocking
/// call and the only place an error can surface (sticky error recorded by an `encode`).
pub struct CudaCommandBuffer;

struct Raw(*mut ffi::uzu_command_buffer);
unsafe impl Send for Raw {}
impl Drop for Raw {
    fn drop(&mut self) {
        unsafe { ffi::uzu_command_buffer_destroy(self.0) }
    }
}

pub struct CudaInitial(Raw);
pub struct CudaEncoding(Raw);
pub struct CudaExecutable(Raw);
pub struct CudaPending(Raw);
pub struct CudaCompleted(Raw);

impl CudaInitial {
    pub(crate) fn from_raw(raw: *mut ffi::uzu_command_buffer) -> Self {
        Self(Raw(raw))
    }
}
impl CudaEncoding {
    /// what every generated `<Name>Kernel::encode` passes to `uzu_<name>_encode`
    pub fn raw(&mut self) -> *mut ffi::uzu_command_buffer {
        (self.0).0
    }
}

impl CommandBuffer for CudaCommandBuffer {
    type Backend = Cuda;
    type Initial = CudaInitial;
    type Encoding = CudaEncoding;
    type Executable = CudaExecutable;
    type Pending = CudaPending;
    type Completed = CudaCompleted;
}

impl CommandBufferInitial for CudaInitial {
    type CommandBuffer = CudaCommandBuffer;
    fn start_encoding(self) -> CudaEncoding {
        // a failure here (wrong state) is recorded as the buffer's sticky error and reported by wait_until_completed
        unsafe { ffi::uzu_command_buffer_start_encoding((self.0).0) };
        CudaEncoding(self.0)
    }
}

impl CommandBufferEncoding for CudaEncoding {
    type CommandBuffer = CudaCommandBuffer;

    fn encode_copy<Src: Buffer<Backend = Cuda>, Dst: Buffer<Backend = Cuda>>(&mut self, src: BufferRangeRef<Src>, dst: BufferRangeMut<Dst>) {
        let (s, d) = (src.gpu_address_range(), dst.gpu_address_range());
        unsafe { ffi::uzu_command_buffer_encode_copy((self.0).0, s.start as u64, d.start as u64, s.end - s.start) }
    }

    fn encode_fill<Dst: Buffer<Backend = Cuda>>(&mut self, dst: BufferRangeMut<Dst>, value: u8) {
        let d = dst.gpu_address_range();
        unsafe { ffi::uzu_command_buffer_encode_fill((self.0).0, d.start as u64, d.end - d.start, value) }
    }

    /// One in-order stream: every barrier of the reference's hazard tracker is already implied.
    fn encode_barrier(&mut self, _after: AccessFlags, _before: AccessFlags) {}

    fn push_debug_group(&mut self, name: &str) {
        let n = CString::new(name).unwrap();
        unsafe { ffi::uzu_command_buffer_push_debug_group((self.0).0, n.as_ptr()) }   // NVTX range
    }
    fn pop_debug_group(&mut self) {
        unsafe { ffi::uzu_command_buffer_pop_debug_group((self.0).0) }
    }

    fn end_encoding(self) -> CudaExecutable {
        unsafe { ffi::uzu_command_buffer_end_encoding((self.0).0) };
        CudaExecutable(self.0)
    }
}

impl CommandBufferExecutable for CudaExecutable {
    type CommandBuffer = CudaCommandBuffer;
    fn submit(self) -> CudaPending {
        unsafe { ffi::uzu_command_buffer_submit((self.0).0) };
        CudaPending(self.0)
    }
}

impl CommandBufferPending for CudaPending {
    type CommandBuffer = CudaCommandBuffer;
    fn wait_until_completed(self) -> Result<CudaCompleted, CudaError> {
        check(unsafe { ffi::uzu_command_buffer_wait_until_completed((self.0).0) })?;
        Ok(CudaCompleted(self.0))
    }
}

impl CommandBufferCompleted for CudaCompleted {
    type CommandBuffer = CudaCommandBuffer;
    fn gpu_execution_time(&self) -> Duration {
        let mut seconds = 0.0f64;
        unsafe { ffi::uzu_command_buffer_gpu_execution_time((self.0).0, &mut seconds) };
        Duration::from_secs_f64(seconds.max(0.0))
    }
}
