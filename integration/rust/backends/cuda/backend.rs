use crate::backends::common::Backend;

use super::{kernel::CudaKernels, CudaBuffer, CudaCommandBuffer, CudaContext, CudaError, CudaSparseBuffer};

/// backends/common/backend.rs:5-18. Alignment 256: TMA / 128-bit vector loads; the allocator clamps
/// `size.next_power_of_two()` into [MIN, MAX] (allocator.rs:128-129).
#[derive(Debug, Clone)]
pub struct Cuda;

impl Backend for Cuda {
    type Context = CudaContext;
    type CommandBuffer = CudaCommandBuffer;
    type DenseBuffer = CudaBuffer;
    type SparseBuffer = CudaSparseBuffer;
    type Kernels = CudaKernels;
    type Error = CudaError;

    const NAME: &'static str = "cuda-b200";
    const MIN_ALLOCATION_ALIGNMENT: usize = 256;
    const MAX_ALLOCATION_ALIGNMENT: usize = 256;
    const ALLOCATION_GRANULARITY: usize = 8 << 20;
    const MAX_INLINE_BYTES: usize = 4096;
}
