//! Example of a GENERATED kernel trait impl (what build/cuda/compiler.rs emits for backends/cpu/kernel/normalization/normalization.rs:7-49):
//! `new` stores the #[specialize] block, `encode` fills `uzu_normalization_args` in the declaration order of the CPU #[kernel] fn.
use crate::{
    backends::common::{kernel::NormalizationKernel, BufferArg, BufferArgMut, Encoder},
    data_type::DataType,
};

use super::{addr, addr_mut};
use crate::backends::cuda::{ffi, Cuda, CudaContext, CudaError};

pub struct CudaNormalizationKernel {
    in_place: bool, subtract_mean: bool, full_layer: bool, copy_to_shortcut: bool, residual_add: bool, use_hadamard: bool,
    scale_residual_sum: bool, scale_output: bool, has_biases: bool, has_scales: bool,
}

impl NormalizationKernel for CudaNormalizationKernel {
    type Backend = Cuda;

    #[allow(clippy::too_many_arguments)]
    fn new(
        _context: &CudaContext, input_t: DataType, _scale_t: DataType, output_t: DataType, _accum_t: DataType,
        in_place: bool, subtract_mean: bool, full_layer: bool, copy_to_shortcut: bool, residual_add: bool, use_hadamard: bool,
        scale_residual_sum: bool, scale_output: bool, has_biases: bool, has_scales: bool,
    ) -> Result<Self, CudaError> {
        if input_t != DataType::BF16 || output_t != DataType::BF16 {
            return Err(CudaError::NotSupported("normalization over non-bf16 activations"));
        }
        if use_hadamard {
            return Err(CudaError::NotSupported("in-norm Hadamard (Mirai RHT)"));
        }
        Ok(Self { in_place, subtract_mean, full_layer, copy_to_shortcut, residual_add, use_hadamard, scale_residual_sum, scale_output, has_biases, has_scales })
    }

    #[allow(clippy::too_many_arguments)]
    fn encode<'input, 'scales, 'biases, 'output, 'shortcut, 'hadamard_factors>(
        &self,
        input: Option<impl BufferArg<'input, Cuda>>,
        scales: Option<impl BufferArg<'scales, Cuda>>,
        biases: Option<impl BufferArg<'biases, Cuda>>,
        output: impl BufferArgMut<'output, Cuda>,
        shortcut: Option<impl BufferArgMut<'shortcut, Cuda>>,
        _hadamard_factors: Option<impl BufferArg<'hadamard_factors, Cuda>>,
        batch_size: u32, element_count: u32, epsilon: f32, scale_offset: f32, post_layer_scalar: f32,
        encoder: &mut Encoder<Cuda>,
    ) {
        let args = ffi::uzu_normalization_args {
            input: input.map(addr).unwrap_or(0), scales: scales.map(addr).unwrap_or(0), biases: biases.map(addr).unwrap_or(0),
            output: addr_mut(output), shortcut: shortcut.map(addr_mut).unwrap_or(0), hadamard_factors: 0,
            batch_size, element_count, epsilon, scale_offset, post_layer_scalar,
            in_place: self.in_place as u32, subtract_mean: self.subtract_mean as u32, full_layer: self.full_layer as u32,
            copy_to_shortcut: self.copy_to_shortcut as u32, residual_add: self.residual_add as u32, use_hadamard: self.use_hadamard as u32,
            scale_residual_sum: self.scale_residual_sum as u32, scale_output: self.scale_output as u32,
            has_biases: self.has_biases as u32, has_scales: self.has_scales as u32,
        };
        // infallible like every generated encode: invalid arguments become the command buffer's sticky error
        unsafe { ffi::uzu_normalization_encode(encoder.as_command_buffer_mut().raw(), &args) };
    }
}
