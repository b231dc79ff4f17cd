// NormalizationKernel and QKVNormKernel.
// Specs: backends/cpu/kernel/normalization/normalization.rs:50-125, attention/qkv_norm.rs:36-76
// (InputT = OutputT = bf16, AffineT = AccumT = f32, the only instantiation the engine creates:
// encodable_block/normalization.rs:84-100, mixer/attention/qkv_norm.rs:97-111).
// Rounding points are kept exactly: residual add rounded to bf16 and written back to the shortcut,
// OnlyNormalization multiplies bf16(norm) * bf16(scale+offset) in bf16, FullLayer stays in f32.
// Compiled with -fmad=false so `a*b+c` rounds twice like the reference.
#include "common.cuh"

namespace uzu {

// One CTA per row. Fast path (element_count % 8 == 0, <= 8 vectors per thread): every thread issues all its 128-bit loads first,
// keeps the (residual-added) bf16 values in registers across the block reduction, and writes shortcut / output once. The
// generic path below handles every other shape. Arithmetic and rounding points are identical in both.
template <int MAXV>
__device__ __forceinline__ void normalization_fast(const uzu_normalization_args& a, float* red) {
    const uint32_t row = blockIdx.x, n = a.element_count;
    const size_t off = (size_t)row * n;
    const __nv_bfloat16* input = (a.in_place ? reinterpret_cast<const __nv_bfloat16*>(a.output) : reinterpret_cast<const __nv_bfloat16*>(a.input)) + off;
    __nv_bfloat16* shortcut = reinterpret_cast<__nv_bfloat16*>(a.shortcut) + off;
    __nv_bfloat16* output = reinterpret_cast<__nv_bfloat16*>(a.output) + off;
    const float* scales = reinterpret_cast<const float*>(a.scales);
    const float* biases = reinterpret_cast<const float*>(a.biases);
    const uint32_t stride = blockDim.x * 8u;
    uint4 v[MAXV], sv[MAXV];
#pragma unroll
    for (int u = 0; u < MAXV; ++u) {
        const uint32_t i = threadIdx.x * 8u + u * stride;
        v[u] = make_uint4(0, 0, 0, 0); sv[u] = make_uint4(0, 0, 0, 0);
        if (i < n) {
            v[u] = *reinterpret_cast<const uint4*>(input + i);
            if (a.copy_to_shortcut && a.residual_add) sv[u] = *reinterpret_cast<const uint4*>(shortcut + i);
        }
    }
    float sum = 0.0f, sum_sq = 0.0f;
#pragma unroll
    for (int u = 0; u < MAXV; ++u) {
        const uint32_t i = threadIdx.x * 8u + u * stride;
        if (i >= n) break;
        __nv_bfloat16* e = reinterpret_cast<__nv_bfloat16*>(&v[u]);
        const __nv_bfloat16* se = reinterpret_cast<const __nv_bfloat16*>(&sv[u]);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            __nv_bfloat16 val = e[j];
            if (a.copy_to_shortcut && a.residual_add) {
                val = f2bf(bf2f(val) + bf2f(se[j]));
                if (a.scale_residual_sum) val = f2bf(bf2f(val) * a.post_layer_scalar);
            }
            e[j] = val;
            const float av = bf2f(val);
            if (a.subtract_mean) sum += av;
            sum_sq += av * av;
        }
        if (a.copy_to_shortcut) *reinterpret_cast<uint4*>(shortcut + i) = v[u];
    }
    const float nf = (float)n;
    float mean = 0.0f;
    if (a.subtract_mean) mean = block_sum(sum, red) / nf;
    sum_sq = block_sum(sum_sq, red);
    const float variance = sum_sq / nf - mean * mean;
    const float rms_inv = 1.0f / sqrtf(variance + a.epsilon);
#pragma unroll
    for (int u = 0; u < MAXV; ++u) {
        const uint32_t i = threadIdx.x * 8u + u * stride;
        if (i >= n) break;
        float sc[8], bs[8];
        if (a.has_scales) {
            const float4 s0 = *reinterpret_cast<const float4*>(scales + i), s1 = *reinterpret_cast<const float4*>(scales + i + 4);
            sc[0] = s0.x; sc[1] = s0.y; sc[2] = s0.z; sc[3] = s0.w; sc[4] = s1.x; sc[5] = s1.y; sc[6] = s1.z; sc[7] = s1.w;
        }
        if (a.has_biases) {
            const float4 b0 = *reinterpret_cast<const float4*>(biases + i), b1 = *reinterpret_cast<const float4*>(biases + i + 4);
            bs[0] = b0.x; bs[1] = b0.y; bs[2] = b0.z; bs[3] = b0.w; bs[4] = b1.x; bs[5] = b1.y; bs[6] = b1.z; bs[7] = b1.w;
        }
        __nv_bfloat16* e = reinterpret_cast<__nv_bfloat16*>(&v[u]);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float normalized = (bf2f(e[j]) - mean) * rms_inv;
            __nv_bfloat16 result;
            if (a.has_scales) {
                if (a.full_layer) result = f2bf(normalized * (sc[j] + a.scale_offset));
                else result = f2bf(bf2f(f2bf(normalized)) * bf2f(f2bf(sc[j] + a.scale_offset)));
            } else {
                result = f2bf(normalized);
            }
            if (a.has_biases) result = f2bf(bf2f(result) + bs[j]);
            if (a.scale_output) result = f2bf(bf2f(result) * bf2f(f2bf(a.post_layer_scalar)));
            e[j] = result;
        }
        *reinterpret_cast<uint4*>(output + i) = v[u];
    }
}

__global__ void __launch_bounds__(1024) normalization_kernel(const uzu_normalization_args a) {
    __shared__ float red[32];
    pdl_launch_dependents();
    pdl_wait();
    {
        const bool aligned = (a.element_count % 8 == 0) && (((a.in_place ? a.output : a.input) | a.output | a.shortcut | a.scales | a.biases) % 16 == 0);
        if (aligned && a.element_count <= blockDim.x * 8u * 4u) {
            normalization_fast<4>(a, red);
            return;
        }
    }
    const uint32_t row = blockIdx.x;
    const uint32_t n = a.element_count;
    const size_t off = (size_t)row * n;
    const __nv_bfloat16* input = a.in_place ? reinterpret_cast<const __nv_bfloat16*>(a.output) : reinterpret_cast<const __nv_bfloat16*>(a.input);
    __nv_bfloat16* shortcut = reinterpret_cast<__nv_bfloat16*>(a.shortcut);
    __nv_bfloat16* output = reinterpret_cast<__nv_bfloat16*>(a.output);
    const float* scales = reinterpret_cast<const float*>(a.scales);
    const float* biases = reinterpret_cast<const float*>(a.biases);

    float sum = 0.0f, sum_sq = 0.0f;
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
        __nv_bfloat16 val = input[off + i];
        if (a.copy_to_shortcut) {
            if (a.residual_add) {
                val = f2bf(bf2f(val) + bf2f(shortcut[off + i]));
                if (a.scale_residual_sum) val = f2bf(bf2f(val) * a.post_layer_scalar);
            }
            shortcut[off + i] = val;
        }
        const float av = bf2f(val);
        if (a.subtract_mean) sum += av;
        sum_sq += av * av;
    }
    const float nf = (float)n;
    float mean = 0.0f;
    if (a.subtract_mean) mean = block_sum(sum, red) / nf;
    sum_sq = block_sum(sum_sq, red);
    const float variance = sum_sq / nf - mean * mean;
    const float rms_inv = 1.0f / sqrtf(variance + a.epsilon);
    __syncthreads();  // shortcut writes of this block are visible to its own re-reads below

    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
        const float iv = a.residual_add ? bf2f(shortcut[off + i]) : bf2f(input[off + i]);
        const float normalized = (iv - mean) * rms_inv;
        __nv_bfloat16 result;
        if (a.has_scales) {
            const float sv = scales[i];
            if (a.full_layer) result = f2bf(normalized * (sv + a.scale_offset));
            else result = f2bf(bf2f(f2bf(normalized)) * bf2f(f2bf(sv + a.scale_offset)));
        } else {
            result = f2bf(normalized);
        }
        if (a.has_biases) result = f2bf(bf2f(result) + biases[i]);
        if (a.scale_output) result = f2bf(bf2f(result) * bf2f(f2bf(a.post_layer_scalar)));
        output[off + i] = result;
    }
}

// one warp per (row, head)
__global__ void __launch_bounds__(128) qkv_norm_kernel(const uzu_qkv_norm_args a) {
    const int lane = threadIdx.x & 31;
    pdl_launch_dependents();
    pdl_wait();
    const uint32_t unit = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (unit >= a.batch_size * a.head_count) return;
    const uint32_t b = unit / a.head_count, h = unit % a.head_count;
    const size_t off = (size_t)b * a.total_heads * a.head_dim + (size_t)(a.head_offset + h) * a.head_dim;
    const __nv_bfloat16* in = a.in_place ? reinterpret_cast<const __nv_bfloat16*>(a.qkv_output) : reinterpret_cast<const __nv_bfloat16*>(a.qkv_input);
    __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(a.qkv_output);
    const float* scales = reinterpret_cast<const float*>(a.scales);
    float total = 0.0f;
    for (uint32_t i = lane; i < a.head_dim; i += 32) {
        const float v = bf2f(in[off + i]);
        total += v * v;
    }
    total = warp_sum(total);
    const float mean_square = total / (float)a.head_dim;
    const float rms = 1.0f / sqrtf(mean_square + a.epsilon);
    for (uint32_t i = lane; i < a.head_dim; i += 32) {
        const float normalized = bf2f(in[off + i]) * rms;
        __nv_bfloat16 r;
        if (!a.has_scales) r = f2bf(normalized);
        else if (a.full_layer) r = f2bf(normalized * (scales[i] + a.scale_offset));
        else r = f2bf(bf2f(f2bf(normalized)) * bf2f(f2bf(scales[i] + a.scale_offset)));
        out[off + i] = r;
    }
}

}  // namespace uzu

extern "C" {

void uzu_normalization_encode(uzu_command_buffer* cmd, const uzu_normalization_args* a) {
    if (!uzu::encodable(cmd, "normalization")) return;
    if (a->use_hadamard || a->hadamard_factors) {
        // the reference CPU kernel is `unimplemented!` here too (normalization.rs:46-48)
        cmd->record_error(UZU_ERROR_UNSUPPORTED, "normalization: in-norm Hadamard (Mirai RHT) is not supported");
        return;
    }
    if ((!a->in_place && !a->input) || !a->output || (a->copy_to_shortcut && !a->shortcut) || (a->residual_add && !a->copy_to_shortcut) ||
        (a->has_scales && !a->scales) || (a->has_biases && !a->biases)) {
        cmd->record_error(UZU_ERROR_INVALID_ARGUMENT, "normalization: inconsistent optional arguments");
        return;
    }
    if (a->batch_size == 0 || a->element_count == 0) return;
    // 128-bit loads: 8 elements per thread per vector, up to 4 vectors per thread on the fast path
    uint32_t threads = a->element_count >= 8192 ? 512 : (a->element_count >= 2048 ? 256 : 128);
    if (a->element_count % 8 != 0) threads = a->element_count >= 4096 ? 1024 : (a->element_count >= 1024 ? 512 : 256);
    uzu::launch(cmd, "normalization_kernel", uzu::normalization_kernel, dim3(a->batch_size), dim3(threads), 0, *a);
}

void uzu_qkv_norm_encode(uzu_command_buffer* cmd, const uzu_qkv_norm_args* a) {
    if (!uzu::encodable(cmd, "qkv_norm")) return;
    if ((!a->in_place && !a->qkv_input) || !a->qkv_output || (a->has_scales && !a->scales)) {
        cmd->record_error(UZU_ERROR_INVALID_ARGUMENT, "qkv_norm: inconsistent optional arguments");
        return;
    }
    uint32_t units = a->batch_size * a->head_count;
    if (units == 0) return;
    uzu::launch(cmd, "qkv_norm_kernel", uzu::qkv_norm_kernel, dim3((units + 3) / 4), dim3(128), 0, *a);
}

}  // extern "C"
