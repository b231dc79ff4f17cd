// DeltaNet prefill: the decode recurrence (DeltaNetConvUpdate + DeltaNetUpdate, backends/cpu/kernel/gdn/{conv_update.rs:8-55,
// update.rs:13-144}) applied to m tokens in ONE launch instead of one launch per token.
//
//   OPT-IN (UZU_DELTA_PREFILL_KERNEL=1 / uzu_debug_set_delta_prefill(1)). Written after round 1's GPU budget was spent: it has NOT run on
//   hardware. The default hybrid prefill keeps the parity-tested decode kernel launched once per token inside the batched pass
//   (engine.cu: encode_delta_net); that per-token launch chain is what bounds Qwen3.5-0.8B prefill (18 layers x m launches).
//
// One CTA per v-head (heads are independent): the head's recurrent state S[Dv][128] (f32, 64 KB for Dv = 128) and the rolling conv state
// of its q / k / v channels stay on chip (shared memory / registers) for all m tokens; per token the CTA
//   1. runs the causal conv + SiLU (bf16-rounded, like the conv kernel) on its 128 q, 128 k and Dv v channels of the raw projection row,
//   2. L2-normalises q and k, scales q by Dk^-0.5, computes beta / decay from the row's scalars,
//   3. updates the state row by row (one warp per row, 4 state elements per lane): retrieved = decay*(S k), delta = beta*(v - retrieved),
//      o = decay*(S q) + delta*(k.q), S = decay*S + k*delta,
//   4. RMS-normalises o over the head, multiplies by norm_weight and SiLU(z), stores bf16.
// Same formulas and rounding points as the single-token kernels (this file is built with -fmad=false like deltanet.cu); reductions are
// tree-shaped, so parity with the oracle's sequential sums is at tolerance level, as for the decode kernel.
// Requires num_v_heads == num_k_heads (a k head's conv state has exactly one owner), head_k_dim = 128, head_v_dim <= 128, <= 7 conv taps.
#include <cstdlib>

#include "common.cuh"

namespace uzu {

constexpr int DP_THREADS = 256;
constexpr int DP_DK = 128;
constexpr int DP_MAX_TAPS = 7;

struct DeltaPrefillParams {
    uzu_delta_net_fused_update_args f;   // per-token arguments; update.in_proj / update.out / conv.in_out point at row 0
    uint32_t rows;                       // tokens
    uint32_t in_stride, out_stride;      // elements between consecutive rows of the projection / the output
};

__device__ __forceinline__ float dp_block_sum(float v, float* red, int nwarps) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.0f;
    for (int w = 0; w < nwarps; ++w) t += red[w];     // fixed order: every thread gets the same bits
    return t;
}

__global__ void __launch_bounds__(DP_THREADS) delta_net_prefill_kernel(const DeltaPrefillParams p) {
    extern __shared__ __align__(16) float dp_smem[];
    const uzu_delta_net_update_args& a = p.f.update;
    const uzu_delta_net_conv_update_args& c = p.f.conv;
    const uint32_t DV = a.head_v_dim;
    float* S = dp_smem;                        // [DV][128]
    float* sq = S + (size_t)DV * DP_DK;        // [128]
    float* sk = sq + DP_DK;                    // [128]
    float* sv = sk + DP_DK;                    // [128]
    float* so = sv + DP_DK;                    // [128]
    float* red = so + DP_DK;                   // [8]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t hv = blockIdx.x, hk = hv;   // num_v_heads == num_k_heads
    const uint32_t conv_dim = 2 * a.key_dim + a.value_dim;
    const uint32_t taps = c.kernel_size - 1;
    float* gstate = reinterpret_cast<float*>(a.state) + (size_t)hv * DV * DP_DK;

    // recurrent state -> shared memory
    for (uint32_t i = tid; i < DV * DP_DK / 4; i += DP_THREADS) reinterpret_cast<float4*>(S)[i] = reinterpret_cast<const float4*>(gstate)[i];

    // conv channels of this thread: threads 0..127 own q_j and k_j (j = tid), threads 128..255 own v_i (i = tid - 128)
    const bool qk_thread = tid < DP_DK, v_thread = tid >= DP_DK && (uint32_t)(tid - DP_DK) < DV;
    const uint32_t ch0 = qk_thread ? hk * DP_DK + tid : 2 * a.key_dim + hv * DV + (tid - DP_DK);   // q channel or v channel
    const uint32_t ch1 = a.key_dim + hk * DP_DK + tid;                                               // k channel (q/k threads)
    float w0[DP_MAX_TAPS + 1], w1[DP_MAX_TAPS + 1], st0[DP_MAX_TAPS], st1[DP_MAX_TAPS];
    float b0 = 0.0f, b1 = 0.0f;
    const float* cw = reinterpret_cast<const float*>(c.conv_weight);
    float* cs = reinterpret_cast<float*>(c.state);
#pragma unroll
    for (int t = 0; t <= DP_MAX_TAPS; ++t) {
        w0[t] = w1[t] = 0.0f;
        if (t < DP_MAX_TAPS) st0[t] = st1[t] = 0.0f;
    }
    if (qk_thread || v_thread) {
#pragma unroll
        for (int t = 0; t <= DP_MAX_TAPS; ++t)
            if ((uint32_t)t <= taps) w0[t] = cw[(size_t)ch0 * c.kernel_size + t];
#pragma unroll
        for (int t = 0; t < DP_MAX_TAPS; ++t)
            if ((uint32_t)t < taps) st0[t] = cs[(size_t)ch0 * c.state_stride + t];
        if (c.has_bias) b0 = reinterpret_cast<const float*>(c.bias)[ch0];
    }
    if (qk_thread) {
#pragma unroll
        for (int t = 0; t <= DP_MAX_TAPS; ++t)
            if ((uint32_t)t <= taps) w1[t] = cw[(size_t)ch1 * c.kernel_size + t];
#pragma unroll
        for (int t = 0; t < DP_MAX_TAPS; ++t)
            if ((uint32_t)t < taps) st1[t] = cs[(size_t)ch1 * c.state_stride + t];
        if (c.has_bias) b1 = reinterpret_cast<const float*>(c.bias)[ch1];
    }
    // one conv step (conv_update.rs:25-54): acc = bias + sum_t state[t]*w[t] + x*w[taps]; out = bf16(SiLU(acc)); state shifts in x
    auto conv_step = [&](float x, const float (&w)[DP_MAX_TAPS + 1], float (&st)[DP_MAX_TAPS], float bias) {
        float acc = bias;
#pragma unroll
        for (int t = 0; t < DP_MAX_TAPS; ++t)
            if ((uint32_t)t < taps) acc += st[t] * w[t];
        float wl = 0.0f;
#pragma unroll
        for (int t = 0; t <= DP_MAX_TAPS; ++t)
            if ((uint32_t)t == taps) wl = w[t];
        acc += x * wl;
#pragma unroll
        for (int t = 0; t < DP_MAX_TAPS - 1; ++t)
            if ((uint32_t)(t + 1) < taps) st[t] = st[t + 1];
#pragma unroll
        for (int t = 0; t < DP_MAX_TAPS; ++t)
            if ((uint32_t)t == taps - 1) st[t] = x;
        return bf2f(f2bf(act_f32(UZU_ACT_SILU, acc)));
    };

    const float a_log = reinterpret_cast<const float*>(a.a_log)[hv], dt_bias = reinterpret_cast<const float*>(a.dt_bias)[hv];
    const float nw = tid < (int)DV ? reinterpret_cast<const float*>(a.norm_weight)[tid] : 0.0f;
    const float qscale = 1.0f / sqrtf((float)DP_DK);
    __syncthreads();

    for (uint32_t t = 0; t < p.rows; ++t) {
        const __nv_bfloat16* row = reinterpret_cast<const __nv_bfloat16*>(a.in_proj) + (size_t)t * p.in_stride;
        // ---- 1. conv + SiLU ----
        float qv = 0.0f, kv = 0.0f;
        if (qk_thread) {
            qv = conv_step(bf2f(row[ch0]), w0, st0, b0);
            kv = conv_step(bf2f(row[ch1]), w1, st1, b1);
        } else if (v_thread) {
            sv[tid - DP_DK] = conv_step(bf2f(row[ch0]), w0, st0, b0);
        }
        // ---- 2. q / k normalisation, k.q, gate scalars (update.rs:60-95) ----
        const float qn = dp_block_sum(qv * qv, red, DP_THREADS / 32);
        const float kn = dp_block_sum(kv * kv, red, DP_THREADS / 32);
        const float qi = 1.0f / sqrtf(qn + 1e-6f), ki = 1.0f / sqrtf(kn + 1e-6f);
        float qq = 0.0f, kk = 0.0f;
        if (qk_thread) {
            qq = qv * qi;
            qq = qq * qscale;
            kk = kv * ki;
            sq[tid] = qq;
            sk[tid] = kk;
        }
        const float kq = dp_block_sum(kk * qq, red, DP_THREADS / 32);     // also orders the sq / sk / sv writes before the reads below
        const float beta_raw = bf2f(row[conv_dim + a.value_dim + hv]);
        const float beta = 1.0f / (1.0f + expf(-beta_raw));
        const float a_raw = bf2f(row[conv_dim + a.value_dim + a.num_v_heads + hv]);
        const float sp_in = a_raw + dt_bias;
        const float sp = sp_in > 20.0f ? sp_in : logf(1.0f + expf(sp_in));
        const float decay = expf(-expf(a_log) * sp);
        // ---- 3. state update, one warp per row ----
        const float4 q4 = *reinterpret_cast<const float4*>(&sq[lane * 4]);
        const float4 k4 = *reinterpret_cast<const float4*>(&sk[lane * 4]);
        for (uint32_t i = warp; i < DV; i += DP_THREADS / 32) {
            float4* srow = reinterpret_cast<float4*>(S + (size_t)i * DP_DK) + lane;
            const float4 s = *srow;
            float sqa = s.x * q4.x + s.y * q4.y + s.z * q4.z + s.w * q4.w;
            float ska = s.x * k4.x + s.y * k4.y + s.z * k4.z + s.w * k4.w;
            sqa = warp_sum(sqa);
            ska = warp_sum(ska);
            const float retrieved = decay * ska;
            const float delta = beta * (sv[i] - retrieved);
            if (lane == 0) so[i] = decay * sqa + delta * kq;
            float4 ns;
            ns.x = decay * s.x + k4.x * delta;
            ns.y = decay * s.y + k4.y * delta;
            ns.z = decay * s.z + k4.z * delta;
            ns.w = decay * s.w + k4.w * delta;
            *srow = ns;
        }
        __syncthreads();
        // ---- 4. RMS norm * norm_weight * SiLU(z) (update.rs:120-143) ----
        const float ov = tid < (int)DV ? so[tid] : 0.0f;
        const float ss = dp_block_sum(ov * ov, red, DP_THREADS / 32);
        const float inv_rms = 1.0f / sqrtf(ss / (float)DV + a.norm_epsilon);
        if (tid < (int)DV) {
            const float z = bf2f(row[conv_dim + hv * DV + tid]);
            const float zs = act_f32(UZU_ACT_SILU, z);
            reinterpret_cast<__nv_bfloat16*>(a.out)[(size_t)t * p.out_stride + hv * DV + tid] = f2bf(ov * inv_rms * nw * zs);
        }
        __syncthreads();                                        // so / sv / sq / sk are rewritten by the next token
    }

    // state and rolling conv state back to global memory
    for (uint32_t i = tid; i < DV * DP_DK / 4; i += DP_THREADS) reinterpret_cast<float4*>(gstate)[i] = reinterpret_cast<const float4*>(S)[i];
    if (qk_thread || v_thread) {
#pragma unroll
        for (int t = 0; t < DP_MAX_TAPS; ++t)
            if ((uint32_t)t < taps) cs[(size_t)ch0 * c.state_stride + t] = st0[t];
    }
    if (qk_thread) {
#pragma unroll
        for (int t = 0; t < DP_MAX_TAPS; ++t)
            if ((uint32_t)t < taps) cs[(size_t)ch1 * c.state_stride + t] = st1[t];
    }
}

static int g_delta_prefill = -1;   // -1: follow UZU_DELTA_PREFILL_KERNEL, 0 / 1: forced

// true = handled: the whole m-token recurrence of one DeltaNet layer in one launch
bool encode_delta_net_prefill(uzu_command_buffer* cmd, const uzu_delta_net_fused_update_args& f, uint32_t rows, uint32_t in_stride, uint32_t out_stride) {
    static const bool env_enabled = [] { const char* e = getenv("UZU_DELTA_PREFILL_KERNEL"); return !e || atoi(e) != 0; }();   // default on (validated on B200, round 2)
    if (!(g_delta_prefill < 0 ? env_enabled : g_delta_prefill != 0)) return false;
    const uzu_delta_net_update_args& a = f.update;
    const uzu_delta_net_conv_update_args& c = f.conv;
    if (rows < 2 || a.head_k_dim != DP_DK || a.head_v_dim == 0 || a.head_v_dim > 128 || (a.head_v_dim & 3u)) return false;
    if (a.num_v_heads != a.num_k_heads || a.key_dim != a.num_k_heads * DP_DK || a.value_dim != a.num_v_heads * a.head_v_dim) return false;
    if (c.kernel_size < 2 || c.kernel_size - 1 > (uint32_t)DP_MAX_TAPS || c.conv_dim != 2 * a.key_dim + a.value_dim || c.in_out != a.in_proj) return false;
    if (!a.in_proj || !a.a_log || !a.dt_bias || !a.norm_weight || !a.state || !a.out || !c.conv_weight || !c.state || (c.has_bias && !c.bias)) return false;
    if ((a.state & 15u)) return false;
    const size_t smem = ((size_t)a.head_v_dim * DP_DK + 4 * DP_DK + 8) * sizeof(float);
    static std::atomic<uint64_t> smem_opt_in{0};
    opt_in_dynamic_smem(cmd, delta_net_prefill_kernel, (int)((int)(((size_t)128 * DP_DK + 4 * DP_DK + 8) * sizeof(float))), smem_opt_in);
    DeltaPrefillParams p{f, rows, in_stride, out_stride};
    delta_net_prefill_kernel<<<a.num_v_heads, DP_THREADS, smem, cmd->ctx->stream>>>(p);
    after_launch(cmd, "delta_net_prefill_kernel");
    return true;
}

}  // namespace uzu

extern "C" void uzu_debug_set_delta_prefill(int mode) { uzu::g_delta_prefill = mode; }
