//! `Backend::Kernels` for CUDA (backends/common/kernel/mod.rs:16-25): the hand-written matmul slot + the 72 generated traits
//! (generated.rs, emitted by tools/gen_rust_kernels.py -- the `build/cuda/compiler.rs` arm of build/cpu/compiler.rs:176-658).
pub(crate) mod generated;
mod matmul;

pub use matmul::CudaMatmul;

use std::convert::Infallible;

use crate::backends::common::Kernels;

use super::Cuda;

pub struct CudaKernels;

impl Kernels for CudaKernels {
    type Backend = Cuda;

    generated::autogen_cuda_kernels!();
    type MatmulKernel = CudaMatmul;
    // Optional cores this backend does not provide are `Infallible`, as in cpu/kernel/mod.rs:35-37: attention always goes through the
    // single-pass / two-pass kernels (the library folds both into one split-KV kernel and picks its tensor-core prefill kernel itself),
    // DeltaNet prefill through the flat decode branch, no DeltaNet tree verification (speculation_supported() == false for hybrids).
    type AttentionGemmCore = Infallible;
    type DeltaNetChunkedPrefill = Infallible;
    type DeltaNetTreeVerify = Infallible;
    type RadixTopKSmall = Infallible;
}

/// `gpu_ptr + byte offset` of a kernel argument (BufferArg::into_parts, buffer/arg.rs:4-59)
pub(crate) fn addr<'a, A: crate::backends::common::BufferArg<'a, Cuda>>(arg: A) -> u64 {
    let (buffer, offset, _len) = arg.into_parts();
    (buffer.gpu_ptr() + offset) as u64
}
pub(crate) fn addr_mut<'a, A: crate::backends::common::BufferArgMut<'a, Cuda>>(arg: A) -> u64 {
    let (buffer, offset, _len) = arg.into_parts();
    (buffer.gpu_ptr() + offset) as u64
}
pub(crate) fn dt(d: crate::data_type::DataType) -> u32 {
    match d {
        crate::data_type::DataType::BF16 => 0,
        crate::data_type::DataType::F32 => 1,
        crate::data_type::DataType::F16 => 2,
        other => panic!("data type {other:?} is not used on the CUDA path"),
    }
}
