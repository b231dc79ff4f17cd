// Elementwise / gather kernels: GatedActMul, embedding lookups, LogitTransform, TensorAddScale / Copy /
// AddBias / AddSwap. Specs: backends/cpu/kernel/{gated_act_mul,embedding,logit_transform,tensor_*}/ *.rs.
// All are bandwidth-trivial (a few KB per token); they exist for drop-in completeness and are fused away
// on the engine's fused decode path. Compiled with -fmad=false (two roundings like the reference).
#include "common.cuh"

namespace uzu {

// out[t,j] = bf16( value * bf16(act(gate)) ): two bf16 roundings (gated_act_mul/mod.rs:5-12)
__global__ void __launch_bounds__(256) gated_act_mul_kernel(const uzu_gated_act_mul_args a) {
    pdl_launch_dependents();
    pdl_wait();
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
    if (j >= a.gated_dim) return;
    const __nv_bfloat16* act_operand = reinterpret_cast<const __nv_bfloat16*>(a.act_operand);
    __nv_bfloat16 value, gate;
    if (a.interleaved) {
        const size_t base = (size_t)b * 2 * a.gated_dim;
        value = act_operand[base + j];
        gate = act_operand[base + a.gated_dim + j];
    } else {
        value = reinterpret_cast<const __nv_bfloat16*>(a.value_operand)[(size_t)b * a.value_row_stride + a.value_offset + j];
        gate = act_operand[(size_t)b * a.gated_dim + j];
    }
    __nv_bfloat16 activated;
    const float gx = bf2f(gate);
    if (a.act_type == UZU_ACT_IDENTITY || (a.act_type == UZU_ACT_SOFTPLUS && gx > 20.0f)) activated = gate;
    else activated = f2bf(act_f32(a.act_type, gx));
    reinterpret_cast<__nv_bfloat16*>(a.fp_out)[(size_t)b * a.gated_dim + j] = f2bf(bf2f(value) * bf2f(activated));
}

__global__ void __launch_bounds__(256) quant_embedding_lookup_kernel(const uzu_quantized_embedding_lookup_args a) {
    pdl_launch_dependents();
    pdl_wait();
    const uint32_t d = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
    if (d >= a.model_dim) return;
    const uint32_t tok = reinterpret_cast<const uint32_t*>(a.token_ids)[b];
    __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(a.output) + (size_t)b * a.model_dim + d;
    if (tok >= a.vocab_size) {
        *out = __float2bfloat16_rn(0.0f);
        return;
    }
    const uint32_t packing = a.quantization_mode == UZU_QMODE_U4 ? 2 : 1;
    const size_t wstride = a.model_dim / packing;
    const uint32_t num_groups = (a.model_dim + a.group_size - 1) / a.group_size;
    const size_t zstride = a.quantization_mode == UZU_QMODE_U4 ? (num_groups + 1) / 2 : num_groups;
    const uint32_t g = d / a.group_size;
    const float scale = bf2f(reinterpret_cast<const __nv_bfloat16*>(a.scales)[(size_t)tok * num_groups + g]);
    const uint8_t* weights = reinterpret_cast<const uint8_t*>(a.weights);
    int32_t qv;
    if (a.quantization_mode == UZU_QMODE_U4) {
        const uint8_t packed = weights[(size_t)tok * wstride + d / 2];
        qv = (d & 1) ? (packed >> 4) : (packed & 15);
    } else if (a.quantization_mode == UZU_QMODE_I8) {
        qv = reinterpret_cast<const int8_t*>(weights)[(size_t)tok * wstride + d];
    } else {
        qv = weights[(size_t)tok * wstride + d];
    }
    float bias;
    if (a.quantization_method == UZU_QMETHOD_SCALE_BIAS) {
        bias = bf2f(reinterpret_cast<const __nv_bfloat16*>(a.biases)[(size_t)tok * num_groups + g]);
    } else if (a.quantization_method == UZU_QMETHOD_SCALE_ZERO_POINT) {
        const uint8_t* zps = reinterpret_cast<const uint8_t*>(a.zero_points);
        uint8_t zp;
        if (a.quantization_mode == UZU_QMODE_U4) {
            const uint8_t packed = zps[(size_t)tok * zstride + g / 2];
            zp = (g & 1) ? (packed >> 4) : (packed & 15);
        } else {
            zp = zps[(size_t)tok * zstride + g];
        }
        bias = -scale * (float)zp;
    } else {
        bias = -scale * (float)(a.quantization_mode == UZU_QMODE_U4 ? 8 : 128);
    }
    float of = scale * (float)qv + bias;
    of = of * a.input_scale;
    *out = f2bf(of);
}

__global__ void __launch_bounds__(256) fp_embedding_lookup_kernel(const uint32_t* token_ids, const __nv_bfloat16* weights,
                                                                  __nv_bfloat16* output, uint32_t vocab_size, uint32_t model_dim,
                                                                  float input_scale) {
    const uint32_t d = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
    if (d >= model_dim) return;
    const uint32_t tok = token_ids[b];
    __nv_bfloat16 r = __float2bfloat16_rn(0.0f);
    if (tok < vocab_size) r = f2bf(bf2f(weights[(size_t)tok * model_dim + d]) * bf2f(f2bf(input_scale)));
    output[(size_t)b * model_dim + d] = r;
}

__global__ void __launch_bounds__(256) logit_transform_kernel(__nv_bfloat16* logits, uint32_t length, float scale, float soft_cap,
                                                              uint32_t has_soft_cap) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= length) return;
    float v = bf2f(logits[i]) * scale;
    if (has_soft_cap) v = tanhf(v / soft_cap) * soft_cap;
    logits[i] = f2bf(v);
}

__global__ void __launch_bounds__(256) tensor_add_scale_kernel(const __nv_bfloat16* input, const __nv_bfloat16* bias, __nv_bfloat16* output,
                                                               uint32_t num_cols, uint32_t length, float scale) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= length) return;
    const float iv = bf2f(input ? input[i] : output[i]);
    output[i] = f2bf((iv + bf2f(bias[i % num_cols])) * scale);
}

__global__ void __launch_bounds__(256) tensor_copy_kernel(const __nv_bfloat16* src, __nv_bfloat16* dst, uint32_t length) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < length) dst[i] = src[i];
}

__global__ void __launch_bounds__(256) tensor_add_bias_kernel(const __nv_bfloat16* input, const __nv_bfloat16* bias, __nv_bfloat16* output,
                                                              uint32_t num_cols, uint32_t length) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= length) return;
    const float v = bf2f(input ? input[i] : output[i]);
    output[i] = f2bf(v + bf2f(bias[i % num_cols]));
}

__global__ void __launch_bounds__(256) tensor_add_swap_kernel(__nv_bfloat16* skip, __nv_bfloat16* main_, uint32_t length) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= length) return;
    const __nv_bfloat16 r = f2bf(bf2f(skip[i]) + bf2f(main_[i]));
    skip[i] = r;
    main_[i] = r;
}

static inline uint32_t blocks_for(uint32_t n) { return (n + 255) / 256; }

}  // namespace uzu

using namespace uzu;

extern "C" {

void uzu_gated_act_mul_encode(uzu_command_buffer* cmd, const uzu_gated_act_mul_args* a) {
    if (!encodable(cmd, "gated_act_mul")) return;
    if (!a->act_operand || !a->fp_out || (!a->interleaved && !a->value_operand)) {
        cmd->record_error(UZU_ERROR_INVALID_ARGUMENT, "gated_act_mul: null operand");
        return;
    }
    if (a->gated_dim == 0 || a->batch_dim == 0) return;
    dim3 grid(blocks_for(a->gated_dim), a->batch_dim);
    launch(cmd, "gated_act_mul_kernel", gated_act_mul_kernel, grid, dim3(256), 0, *a);
}

void uzu_quantized_embedding_lookup_encode(uzu_command_buffer* cmd, const uzu_quantized_embedding_lookup_args* a) {
    if (!encodable(cmd, "quantized_embedding_lookup")) return;
    if (!a->token_ids || !a->weights || !a->scales || !a->output || a->group_size == 0 ||
        (a->quantization_method == UZU_QMETHOD_SCALE_ZERO_POINT && !a->zero_points) ||
        (a->quantization_method == UZU_QMETHOD_SCALE_BIAS && !a->biases)) {
        cmd->record_error(UZU_ERROR_INVALID_ARGUMENT, "quantized_embedding_lookup: inconsistent arguments");
        return;
    }
    if (a->batch_size == 0 || a->model_dim == 0) return;
    dim3 grid(blocks_for(a->model_dim), a->batch_size);
    launch(cmd, "quant_embedding_lookup_kernel", quant_embedding_lookup_kernel, grid, dim3(256), 0, *a);
}

void uzu_full_precision_embedding_lookup_encode(uzu_command_buffer* cmd, uint64_t token_ids, uint64_t weights, uint64_t output,
                                                uint32_t batch_size, uint32_t vocab_size, uint32_t model_dim, float input_scale) {
    if (!encodable(cmd, "full_precision_embedding_lookup")) return;
    if (!token_ids || !weights || !output) {
        cmd->record_error(UZU_ERROR_INVALID_ARGUMENT, "full_precision_embedding_lookup: null operand");
        return;
    }
    if (batch_size == 0 || model_dim == 0) return;
    dim3 grid(blocks_for(model_dim), batch_size);
    fp_embedding_lookup_kernel<<<grid, 256, 0, cmd->ctx->stream>>>((const uint32_t*)token_ids, (const __nv_bfloat16*)weights,
                                                                   (__nv_bfloat16*)output, vocab_size, model_dim, input_scale);
    after_launch(cmd, "fp_embedding_lookup_kernel");
}

void uzu_logit_transform_encode(uzu_command_buffer* cmd, uint64_t logits, uint32_t length, float scale, float soft_cap,
                                uint32_t has_soft_cap) {
    if (!encodable(cmd, "logit_transform") || length == 0) return;
    logit_transform_kernel<<<blocks_for(length), 256, 0, cmd->ctx->stream>>>((__nv_bfloat16*)logits, length, scale, soft_cap, has_soft_cap);
    after_launch(cmd, "logit_transform_kernel");
}

void uzu_tensor_add_scale_encode(uzu_command_buffer* cmd, uint64_t input, uint64_t bias, uint64_t output, uint32_t num_cols,
                                 uint32_t length, float scale) {
    if (!encodable(cmd, "tensor_add_scale") || length == 0) return;
    tensor_add_scale_kernel<<<blocks_for(length), 256, 0, cmd->ctx->stream>>>((const __nv_bfloat16*)input, (const __nv_bfloat16*)bias,
                                                                              (__nv_bfloat16*)output, num_cols, length, scale);
    after_launch(cmd, "tensor_add_scale_kernel");
}

void uzu_tensor_copy_encode(uzu_command_buffer* cmd, uint64_t src, uint64_t dst, uint32_t length) {
    if (!encodable(cmd, "tensor_copy") || length == 0) return;
    tensor_copy_kernel<<<blocks_for(length), 256, 0, cmd->ctx->stream>>>((const __nv_bfloat16*)src, (__nv_bfloat16*)dst, length);
    after_launch(cmd, "tensor_copy_kernel");
}

void uzu_tensor_add_bias_encode(uzu_command_buffer* cmd, uint64_t input, uint64_t bias, uint64_t output, uint32_t num_cols,
                                uint32_t length) {
    if (!encodable(cmd, "tensor_add_bias") || length == 0) return;
    tensor_add_bias_kernel<<<blocks_for(length), 256, 0, cmd->ctx->stream>>>((const __nv_bfloat16*)input, (const __nv_bfloat16*)bias,
                                                                             (__nv_bfloat16*)output, num_cols, length);
    after_launch(cmd, "tensor_add_bias_kernel");
}

void uzu_tensor_add_swap_encode(uzu_command_buffer* cmd, uint64_t skip_buffer, uint64_t main_buffer, uint32_t length) {
    if (!encodable(cmd, "tensor_add_swap") || length == 0) return;
    tensor_add_swap_kernel<<<blocks_for(length), 256, 0, cmd->ctx->stream>>>((__nv_bfloat16*)skip_buffer, (__nv_bfloat16*)main_buffer, length);
    after_launch(cmd, "tensor_add_swap_kernel");
}

}  // extern "C"
