use crate::{
    backends::common::{
        kernel::matmul::{arguments::MatmulArguments, matmul_a::MatmulA, matmul_b::MatmulB, MatmulKernel},
        BufferArg, Encoder,
    },
    data_type::DataType,
};

use super::{addr, dt};
use crate::backends::cuda::{error::check, ffi, Cuda, CudaContext, CudaError};

/// backends/common/kernel/matmul/kernel.rs:12-43. One entry point for every shape: the library picks the fused dequant + GEMV
/// (m <= 16), the tcgen05 prefill GEMM with the in-kernel dequant stage (m >= 64, quantised B, bf16 A) or the generic kernel.
/// `a8_activation_plan` / `select_activation_format` keep their defaults (None / Bf16): bf16 activations only.
pub struct CudaMatmul {
    weights: DataType,
    input: DataType,
    output: DataType,
}

impl MatmulKernel for CudaMatmul {
    type Backend = Cuda;

    fn new(_context: &CudaContext, weights: DataType, input: DataType, output: DataType) -> Result<Self, CudaError> {
        Ok(Self { weights, input, output })
    }

    fn encode<'a, 'b, 'd, TB: BufferArg<'b, Cuda>>(
        &mut self,
        a: MatmulArguments<'a, 'b, 'd, Cuda, TB>,
        encoder: &mut Encoder<Cuda>,
    ) -> Result<(), CudaError> {
        let mut args = ffi::uzu_matmul_args::default();
        match a.a {
            MatmulA::FullPrecision { values, offset } => args.a = addr(values) + (offset * self.input.size_in_bytes()) as u64,
            MatmulA::Int8Symmetric { .. } => return Err(CudaError::NotSupported("int8 (A8) activations")),
        }
        let quant = |args: &mut ffi::uzu_matmul_args, kind: u32, mode, group_size: u32, signed: bool| {
            args.b_prologue = kind;
            args.b_mode = mode as u32;
            args.b_group_size = group_size;
            args.b_signed_codes = signed as u32;
        };
        match a.b {
            MatmulB::FullPrecision { b } => {
                args.b_prologue = 0;
                args.b = addr(b);
            }
            MatmulB::ScaleBiasDequant { b, scales, biases, mode, group_size, signed_codes } => {
                quant(&mut args, 1, mode, group_size, signed_codes);
                (args.b, args.b_scales, args.b_biases) = (addr(b), addr(scales), addr(biases));
            }
            MatmulB::ScaleZeroPointDequant { b, scales, zero_points, mode, group_size, signed_codes } => {
                quant(&mut args, 2, mode, group_size, signed_codes);
                (args.b, args.b_scales, args.b_zero_points) = (addr(b), addr(scales), addr(zero_points));
            }
            MatmulB::ScaleSymmetricDequant { b, scales, mode, group_size, signed_codes } => {
                quant(&mut args, 3, mode, group_size, signed_codes);
                (args.b, args.b_scales) = (addr(b), addr(scales));
            }
        }
        // MatmulDOps::rht_factors: the library runs OutputRht over D after the epilogue and adds the bias after it (kernel.rs:297-303)
        args.rht_factors = a.d_transform.rht_factors.map(|f| addr(f)).unwrap_or(0);
        args.d_transform = a.d_transform.mask().bits();
        args.ab_scale = a.d_transform.ab_scale;
        args.soft_cap = a.d_transform.soft_cap.unwrap_or(0.0);
        args.bias = a.d_transform.bias.map(|b| addr(b)).unwrap_or(0);
        args.gather_indices = a.gather_indices.map(|g| addr(g)).unwrap_or(0);
        args.d = addr(&*a.d);
        args.b_transpose = a.b_transpose as u32;
        args.b_leading_dimension = a.b_leading_dimension.unwrap_or(0);
        (args.m, args.n, args.k) = (a.m, a.n, a.k);
        (args.weights_dt, args.input_dt, args.output_dt) = (dt(self.weights), dt(self.input), dt(self.output));
        check(unsafe { ffi::uzu_matmul_validate(&args) })?;      // the only fallible part; encode itself records sticky errors
        unsafe { ffi::uzu_matmul_encode(encoder.as_command_buffer_mut().raw(), &args) };
        Ok(())
    }
}
