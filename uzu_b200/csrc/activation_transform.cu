// ActivationTransformKernel (Mirai RHT): crates/backend-uzu/src/backends/cpu/kernel/activation_transform/
// {activation_transform.rs:44-136, mod.rs:9-44}; callers encodable_block/linear/rht_wrapper.rs:214-297 (input RHT of a HybridSpec linear)
// and the matmul's output-RHT epilogue (cpu/kernel/matmul/kernel.rs:297-303, matmul.cu: UZU_D_RHT).
//
// A 32-point Walsh-Hadamard transform per 32-wide stripe is exactly one warp: lane l holds element l, the five butterfly stages
// (stride 1, 2, 4, 8, 16: lower = a + b, upper = a - b) are five xor-shuffles, all in f32 in the reference's order, so the result is
// bit-identical to the CPU loop. The work is elementwise and HBM-trivial ([m, K] activations); what matters is that it is one launch
// with coalesced 64-byte (bf16) / 128-byte (f32) row segments per warp.
//
// Quantize ops (symmetric int8 per activation group on the input-transformed row, optional integer group sums): one warp owns a span of
// max(activation group, sum group) columns (<= 8 stripes), keeps the transformed values in registers, reduces max|t| with shuffles and
// writes codes, the group divisor and the code sums.
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <string>

#include "common.cuh"
#include "uzu_b200.h"

namespace uzu {

__device__ __forceinline__ float hadamard32_lane(float v, uint32_t lane) {
#pragma unroll
    for (uint32_t stride = 1; stride < 32; stride <<= 1) {
        const float other = __shfl_xor_sync(0xffffffffu, v, stride);
        v = (lane & stride) ? __fsub_rn(other, v) : __fadd_rn(v, other);   // upper lane: a - b with a = the lower lane's value
    }
    return __fmul_rn(v, __fdiv_rn(1.0f, __fsqrt_rn(32.0f)));              // 1 / sqrt(32) evaluated in f32 like the reference
}

template <typename T>
__device__ __forceinline__ float at_load(const T* p, size_t i);
template <>
__device__ __forceinline__ float at_load<float>(const float* p, size_t i) { return p[i]; }
template <>
__device__ __forceinline__ float at_load<__nv_bfloat16>(const __nv_bfloat16* p, size_t i) { return __bfloat162float(p[i]); }
__device__ __forceinline__ void at_store(float* p, size_t i, float v) { p[i] = v; }
__device__ __forceinline__ void at_store(__nv_bfloat16* p, size_t i, float v) { p[i] = __float2bfloat16_rn(v); }

// InputRht: out = H (s o x);  OutputRht: out = s o (H x). One warp per stripe; `input` may alias `out` (each warp reads its stripe, then writes it).
template <typename T, bool INPUT_ORDER>
__global__ void __launch_bounds__(256) activation_rht_kernel(const T* input, T* out, const int32_t* factors, uint32_t rows, uint32_t cols) {
    const uint32_t lane = threadIdx.x & 31;
    const size_t stripes_per_row = cols / 32;
    const size_t stripe = (size_t)blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
    if (stripe >= (size_t)rows * stripes_per_row) return;
    const uint32_t col = (uint32_t)(stripe % stripes_per_row) * 32 + lane;
    const size_t idx = (stripe / stripes_per_row) * cols + col;
    const float f = (float)factors[col];
    float v = at_load<T>(input, idx);
    if (INPUT_ORDER) v = __fmul_rn(v, f);
    v = hadamard32_lane(v, lane);
    if (!INPUT_ORDER) v = __fmul_rn(v, f);
    at_store(out, idx, v);
}

// Quantize / QuantizeWithGroupSums. SPAN = columns one warp owns = max(activation group, sum group); both are multiples of 32.
template <typename T>
__global__ void __launch_bounds__(256) activation_quantize_kernel(const T* input, int8_t* q_out, float* scales_out, int32_t* group_sums_out,
                                                                  const int32_t* factors, uint32_t rows, uint32_t cols, uint32_t group,
                                                                  uint32_t sum_group, uint32_t span) {
    const uint32_t lane = threadIdx.x & 31;
    const size_t spans_per_row = cols / span;
    const size_t s = (size_t)blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
    if (s >= (size_t)rows * spans_per_row) return;
    const size_t row = s / spans_per_row;
    const uint32_t col0 = (uint32_t)(s % spans_per_row) * span;
    const uint32_t stripes = span / 32;                      // <= 8
    float t[8];
#pragma unroll
    for (uint32_t i = 0; i < 8; ++i) {
        if (i < stripes) {
            const uint32_t col = col0 + i * 32 + lane;
            t[i] = hadamard32_lane(__fmul_rn(at_load<T>(input, row * cols + col), (float)factors[col]), lane);
        } else {
            t[i] = 0.0f;
        }
    }
    const uint32_t stripes_per_group = group / 32;
    const uint32_t sum_stripes = sum_group ? sum_group / 32 : 0;
    int run_sum = 0;
    for (uint32_t g0 = 0; g0 < stripes; g0 += stripes_per_group) {
        // min_max_symmetric_divisor (mod.rs:9-19): magnitude = max(|min|, |max|) = max |t| over the group
        // (f32::min / f32::max skip NaNs, and so does fmaxf)
        float mag = 0.0f;
#pragma unroll
        for (uint32_t i = 0; i < 8; ++i)
            if (i >= g0 && i < g0 + stripes_per_group) mag = fmaxf(mag, fabsf(t[i]));
#pragma unroll
        for (uint32_t o = 16; o > 0; o >>= 1) mag = fmaxf(mag, __shfl_xor_sync(0xffffffffu, mag, o));
        const float divisor = (isfinite(mag) && mag > 0.0f) ? __fdiv_rn(mag, 127.0f) : 1.0f;
        if (lane == 0) scales_out[row * (cols / group) + (col0 + g0 * 32) / group] = divisor;
#pragma unroll
        for (uint32_t i = 0; i < 8; ++i)
            if (i >= g0 && i < g0 + stripes_per_group) {
                // quantize_symmetric_i8 (mod.rs:21-29): f32::round = half away from zero, clamp, saturating cast
                const float q = roundf(__fdiv_rn(t[i], divisor));
                const int code = isnan(q) ? 0 : (int)fminf(fmaxf(q, -127.0f), 127.0f);   // clamp keeps NaN, `NaN as i8` is 0
                q_out[row * cols + col0 + i * 32 + lane] = (int8_t)code;
                if (sum_stripes) {
                    int ssum = code;
#pragma unroll
                    for (uint32_t o = 16; o > 0; o >>= 1) ssum += __shfl_xor_sync(0xffffffffu, ssum, o);
                    run_sum += ssum;
                    if ((i + 1) % sum_stripes == 0) {
                        if (lane == 0) group_sums_out[row * (cols / sum_group) + (col0 + i * 32) / sum_group] = run_sum;
                        run_sum = 0;
                    }
                }
            }
    }
}

}  // namespace uzu

using namespace uzu;

extern "C" {

uzu_status uzu_activation_transform_validate(const uzu_activation_transform_args* a) {
    auto bad = [](const char* m) { return fail(UZU_ERROR_INVALID_ARGUMENT, std::string("activation_transform: ") + m); };
    if (!a) return bad("null arguments");
    if (a->data_type != UZU_DT_BF16 && a->data_type != UZU_DT_F32) return bad("data type must be bf16 or f32");
    if (a->ops > UZU_ACTIVATION_TRANSFORM_QUANTIZE_WITH_GROUP_SUMS) return bad("bad op");
    if (a->element_count == 0 || a->element_count % 32 != 0) return bad("element_count must be a positive multiple of HADAMARD_TRANSFORM_BLOCK_SIZE (32)");
    if (!a->rht_factors) return bad("null rht_factors");
    const bool quant = a->ops >= UZU_ACTIVATION_TRANSFORM_QUANTIZE;
    if (quant) {
        if (a->in_place || !a->input) return bad("quantize ops read `input` (not in place)");
        if (!a->q_out || !a->scales_out) return bad("quantize ops need q_out and scales_out");
        const uint32_t g = a->activation_scale_group_size, s = a->ops == UZU_ACTIVATION_TRANSFORM_QUANTIZE_WITH_GROUP_SUMS ? a->sum_group_size : 0;
        if (g == 0 || g % 32 != 0 || a->element_count % g != 0) return bad("activation group must be a multiple of 32 dividing element_count");
        if (a->ops == UZU_ACTIVATION_TRANSFORM_QUANTIZE_WITH_GROUP_SUMS) {
            if (!a->group_sums_out) return bad("QuantizeWithGroupSums needs group_sums_out");
            if (s == 0 || s % 32 != 0 || a->element_count % s != 0) return bad("sum group must be a multiple of 32 dividing element_count");
        }
        const uint32_t span = s > g ? s : g;
        if (span > 256 || span % g != 0 || (s && span % s != 0)) return bad("group sizes above 256 (or not nested) are not supported");
    } else {
        if (!a->fp_out) return bad("RHT ops need fp_out");
        if (!a->in_place && !a->input) return bad("out-of-place transform needs input");
    }
    return UZU_OK;
}

void uzu_activation_transform_encode(uzu_command_buffer* cmd, const uzu_activation_transform_args* a) {
    if (!encodable(cmd, "activation_transform")) return;
    if (uzu_activation_transform_validate(a) != UZU_OK) {
        cmd->record_error(UZU_ERROR_INVALID_ARGUMENT, uzu_last_error());
        return;
    }
    if (a->batch_size == 0) return;
    cudaStream_t s = cmd->ctx->stream;
    const uint64_t in = a->in_place ? a->fp_out : a->input;
    const bool f32 = a->data_type == UZU_DT_F32;
    const int32_t* fac = (const int32_t*)a->rht_factors;
    if (a->ops <= UZU_ACTIVATION_TRANSFORM_OUTPUT_RHT) {
        const size_t stripes = (size_t)a->batch_size * (a->element_count / 32);
        const uint32_t grid = (uint32_t)((stripes + 7) / 8);
        const bool input_order = a->ops == UZU_ACTIVATION_TRANSFORM_INPUT_RHT;
        if (f32) {
            if (input_order) activation_rht_kernel<float, true><<<grid, 256, 0, s>>>((const float*)in, (float*)a->fp_out, fac, a->batch_size, a->element_count);
            else activation_rht_kernel<float, false><<<grid, 256, 0, s>>>((const float*)in, (float*)a->fp_out, fac, a->batch_size, a->element_count);
        } else {
            if (input_order) activation_rht_kernel<__nv_bfloat16, true><<<grid, 256, 0, s>>>((const __nv_bfloat16*)in, (__nv_bfloat16*)a->fp_out, fac, a->batch_size, a->element_count);
            else activation_rht_kernel<__nv_bfloat16, false><<<grid, 256, 0, s>>>((const __nv_bfloat16*)in, (__nv_bfloat16*)a->fp_out, fac, a->batch_size, a->element_count);
        }
        after_launch(cmd, "activation_rht_kernel");
        return;
    }
    const uint32_t g = a->activation_scale_group_size;
    const uint32_t sg = a->ops == UZU_ACTIVATION_TRANSFORM_QUANTIZE_WITH_GROUP_SUMS ? a->sum_group_size : 0;
    const uint32_t span = sg > g ? sg : g;
    const size_t spans = (size_t)a->batch_size * (a->element_count / span);
    const uint32_t grid = (uint32_t)((spans + 7) / 8);
    if (f32) activation_quantize_kernel<float><<<grid, 256, 0, s>>>((const float*)in, (int8_t*)a->q_out, (float*)a->scales_out, (int32_t*)a->group_sums_out, fac, a->batch_size, a->element_count, g, sg, span);
    else activation_quantize_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>((const __nv_bfloat16*)in, (int8_t*)a->q_out, (float*)a->scales_out, (int32_t*)a->group_sums_out, fac, a->batch_size, a->element_count, g, sg, span);
    after_launch(cmd, "activation_quantize_kernel");
}

}  // extern "C"
