// DeltaNet decode kernels for Qwen3.5 hybrid layers:
//   DeltaNetConvUpdate  backends/cpu/kernel/gdn/conv_update.rs:8-55   (1-token causal conv + SiLU, f32 rolling state)
//   DeltaNetUpdate      backends/cpu/kernel/gdn/update.rs:13-144      (gated delta rule over S[Hv, Dv, Dk] + RMSNorm * SiLU(z))
// Parameter / state dtypes are f32 (what the engine allocates and Metal declares; the reference CPU
// kernel's `*const T` typing of these buffers is the bug documented in SURVEY.md row a11).
// State traffic: Hv*Dv*Dk*4 B read + written per layer per token (2 MB for Qwen3.5-0.8B).
#include "common.cuh"

namespace uzu {

__global__ void __launch_bounds__(256) delta_net_conv_update_kernel(const uzu_delta_net_conv_update_args a) {
    pdl_launch_dependents();
    pdl_wait();
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= a.conv_dim) return;
    const float* w = reinterpret_cast<const float*>(a.conv_weight) + (size_t)c * a.kernel_size;
    float* st = reinterpret_cast<float*>(a.state) + (size_t)c * a.state_stride;
    __nv_bfloat16* io = reinterpret_cast<__nv_bfloat16*>(a.in_out);
    const uint32_t taps = a.kernel_size - 1;
    const float x = bf2f(io[c]);
    float acc = a.has_bias ? reinterpret_cast<const float*>(a.bias)[c] : 0.0f;
    for (uint32_t t = 0; t < taps; ++t) acc += st[t] * w[t];
    acc += x * w[taps];
    io[c] = f2bf(act_f32(UZU_ACT_SILU, acc));
    for (uint32_t t = 1; t < taps; ++t) st[t - 1] = st[t];
    st[taps - 1] = x;
}

// A cluster of DN_CLUSTER CTAs per v-head; HEAD_K_DIM = 128 (the only variant the reference instantiates). Each CTA owns
// head_v_dim / DN_CLUSTER state rows (one warp per row, two rows in flight per warp); the RMS statistic of the head's
// output is exchanged through distributed shared memory and summed in rank order by every CTA (deterministic).
constexpr int DN_WARPS = 8;
constexpr int DN_CLUSTER = 8;
constexpr int DN_MAX_TAPS = 7;

// DeltaNetConvUpdate for one channel (conv_update.rs:8-55): returns the bf16-rounded SiLU output as f32, leaves the advanced rolling
// state in `ns` (the caller stores it once every reader of the old state is done)
__device__ __forceinline__ float conv_channel(const uzu_delta_net_conv_update_args& c, uint32_t ch, float x, float (&ns)[DN_MAX_TAPS]) {
    const float* w = reinterpret_cast<const float*>(c.conv_weight) + (size_t)ch * c.kernel_size;
    const float* st = reinterpret_cast<const float*>(c.state) + (size_t)ch * c.state_stride;
    const uint32_t taps = c.kernel_size - 1;
    float acc = c.has_bias ? reinterpret_cast<const float*>(c.bias)[ch] : 0.0f;
#pragma unroll
    for (uint32_t t = 0; t < DN_MAX_TAPS; ++t) {
        if (t < taps) {
            const float sv = st[t];
            acc += sv * w[t];
            if (t >= 1) ns[t - 1] = sv;
        }
    }
    acc += x * w[taps];
    ns[taps - 1] = x;
    return bf2f(f2bf(act_f32(UZU_ACT_SILU, acc)));
}
__device__ __forceinline__ void conv_store_state(const uzu_delta_net_conv_update_args& c, uint32_t ch, const float (&ns)[DN_MAX_TAPS]) {
    float* st = reinterpret_cast<float*>(c.state) + (size_t)ch * c.state_stride;
#pragma unroll
    for (uint32_t t = 0; t < DN_MAX_TAPS; ++t)
        if (t < c.kernel_size - 1) st[t] = ns[t];
}

// A cluster of DN_CLUSTER CTAs per v-head; HEAD_K_DIM = 128 (the only variant the reference instantiates). Each CTA owns
// head_v_dim / DN_CLUSTER state rows (one warp per row, four rows in flight per warp); the RMS statistic of the head's
// output is exchanged through distributed shared memory and summed in rank order by every CTA (deterministic).
// FUSED_CONV: q / k / v come from the raw projection through the causal conv (see uzu_delta_net_fused_update_args).
template <bool FUSED_CONV>
__global__ void __cluster_dims__(DN_CLUSTER, 1, 1) __launch_bounds__(DN_WARPS * 32) delta_net_update_kernel(const uzu_delta_net_fused_update_args fa) {
    constexpr int DK = 128;
    const uzu_delta_net_update_args& a = fa.update;
    __shared__ float sq[DK], sk[DK];
    __shared__ float so[256];
    __shared__ float sv[256];
    __shared__ float red[32];
    __shared__ float ss_parts[DN_CLUSTER];
    pdl_launch_dependents();
    if (!FUSED_CONV) asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");   // "this CTA is running": awaited before any DSMEM store
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t hv = blockIdx.x / DN_CLUSTER;
    uint32_t crank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
    const uint32_t hk = hv / (a.num_v_heads / a.num_k_heads);
    const uint32_t conv_dim = 2 * a.key_dim + a.value_dim;
    const __nv_bfloat16* in_proj = reinterpret_cast<const __nv_bfloat16*>(a.in_proj);
    const uint32_t rows_per_cta = (a.head_v_dim + DN_CLUSTER - 1) / DN_CLUSTER;
    const uint32_t row0 = crank * rows_per_cta;
    const uint32_t row1 = min(a.head_v_dim, row0 + rows_per_cta);
    const uint32_t nrows = row1 > row0 ? row1 - row0 : 0;
    float* state = reinterpret_cast<float*>(a.state) + (size_t)hv * a.head_v_dim * DK;

    // the recurrent state does not depend on the previous kernel: fetch this warp's rows before the grid dependency resolves
    constexpr int RPW = 4;                                                       // rows in flight per warp
    float4 srow[RPW];
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
        const uint32_t i = row0 + warp + r * DN_WARPS;
        srow[r] = i < row1 ? *(reinterpret_cast<const float4*>(state + (size_t)i * DK) + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    pdl_wait();

    // q, k of this k head and the v rows of this CTA (bf16 values as f32)
    float qv = 0.0f, kv = 0.0f;
    float nsq[DN_MAX_TAPS], nsk[DN_MAX_TAPS], nsv[DN_MAX_TAPS];
    const uint32_t cq = hk * DK + threadIdx.x, ck = a.key_dim + hk * DK + threadIdx.x;
    const uint32_t cvch = 2 * a.key_dim + hv * a.head_v_dim + row0 + (threadIdx.x - DK);     // threads DK .. DK + nrows - 1
    const bool v_thread = threadIdx.x >= DK && threadIdx.x - DK < nrows;
    if (FUSED_CONV) {
        if (threadIdx.x < DK) {
            qv = conv_channel(fa.conv, cq, bf2f(in_proj[cq]), nsq);
            kv = conv_channel(fa.conv, ck, bf2f(in_proj[ck]), nsk);
        } else if (v_thread) {
            sv[threadIdx.x - DK] = conv_channel(fa.conv, cvch, bf2f(in_proj[cvch]), nsv);
        }
        // every CTA of the cluster has now read the old conv state of this head's q / k channels
        asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    } else {
        if (threadIdx.x < DK) {
            qv = bf2f(in_proj[cq]);
            kv = bf2f(in_proj[ck]);
        } else if (v_thread) {
            sv[threadIdx.x - DK] = bf2f(in_proj[cvch]);
        }
    }

    // L2-normalise q and k, scale q by Dk^-0.5 (update.rs:60-80)
    const float qn = block_sum(qv * qv, red);
    const float kn = block_sum(kv * kv, red);
    const float qi = 1.0f / sqrtf(qn + 1e-6f), ki = 1.0f / sqrtf(kn + 1e-6f);
    const float qscale = 1.0f / sqrtf((float)DK);
    if (threadIdx.x < DK) {
        float qq = qv * qi;
        qq = qq * qscale;
        sq[threadIdx.x] = qq;
        sk[threadIdx.x] = kv * ki;
    }
    __syncthreads();
    const float kq_part = threadIdx.x < DK ? sk[threadIdx.x] * sq[threadIdx.x] : 0.0f;
    const float kq = block_sum(kq_part, red);

    const float beta_raw = bf2f(in_proj[conv_dim + a.value_dim + hv]);
    const float beta = 1.0f / (1.0f + expf(-beta_raw));
    const float a_raw = bf2f(in_proj[conv_dim + a.value_dim + a.num_v_heads + hv]);
    const float sp_in = a_raw + reinterpret_cast<const float*>(a.dt_bias)[hv];
    const float sp = sp_in > 20.0f ? sp_in : logf(1.0f + expf(sp_in));
    const float gdec = -expf(reinterpret_cast<const float*>(a.a_log)[hv]) * sp;
    const float decay = expf(gdec);

    const float4 q4 = *reinterpret_cast<const float4*>(&sq[lane * 4]);
    const float4 k4 = *reinterpret_cast<const float4*>(&sk[lane * 4]);
    for (uint32_t base = row0 + warp; base < row1; base += RPW * DN_WARPS) {
        if (base != row0 + warp) {
#pragma unroll
            for (int r = 0; r < RPW; ++r) {
                const uint32_t i = base + r * DN_WARPS;
                if (i < row1) srow[r] = *(reinterpret_cast<const float4*>(state + (size_t)i * DK) + lane);
            }
        }
#pragma unroll
        for (int r = 0; r < RPW; ++r) {
            const uint32_t i = base + r * DN_WARPS;
            if (i >= row1) break;
            const float4 s = srow[r];
            float sqa = s.x * q4.x + s.y * q4.y + s.z * q4.z + s.w * q4.w;
            float ska = s.x * k4.x + s.y * k4.y + s.z * k4.z + s.w * k4.w;
            sqa = warp_sum(sqa);
            ska = warp_sum(ska);
            const float v_i = sv[i - row0];
            const float retrieved = decay * ska;
            const float delta = beta * (v_i - retrieved);
            if (lane == 0) so[i - row0] = decay * sqa + delta * kq;
            float4 ns;
            ns.x = decay * s.x + k4.x * delta;
            ns.y = decay * s.y + k4.y * delta;
            ns.z = decay * s.z + k4.z * delta;
            ns.w = decay * s.w + k4.w * delta;
            *(reinterpret_cast<float4*>(state + (size_t)i * DK) + lane) = ns;
        }
    }
    __syncthreads();
    const float ov = threadIdx.x < nrows ? so[threadIdx.x] : 0.0f;
    const float ss_local = block_sum(ov * ov, red);
    // publish this CTA's partial into every CTA of the cluster (DSMEM), then sum the DN_CLUSTER partials in rank order
    if (FUSED_CONV) asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    else asm volatile("barrier.cluster.wait.aligned;" ::: "memory");
    if (FUSED_CONV) {
        // advance the rolling conv state: v channels by their owner, the head's q / k channels by cluster rank 0
        if (threadIdx.x < DK && crank == 0) {
            conv_store_state(fa.conv, cq, nsq);
            conv_store_state(fa.conv, ck, nsk);
        } else if (v_thread) {
            conv_store_state(fa.conv, cvch, nsv);
        }
    }
    if (threadIdx.x < DN_CLUSTER) {
        const uint32_t local = (uint32_t)__cvta_generic_to_shared(&ss_parts[crank]);
        uint32_t remote;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"((uint32_t)threadIdx.x));
        asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(remote), "f"(ss_local) : "memory");
    }
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    float ss = 0.0f;
#pragma unroll
    for (int c = 0; c < DN_CLUSTER; ++c) ss += ss_parts[c];
    const float inv_rms = 1.0f / sqrtf(ss / (float)a.head_v_dim + a.norm_epsilon);
    if (threadIdx.x < nrows) {
        const uint32_t i = row0 + threadIdx.x;
        const float nw = reinterpret_cast<const float*>(a.norm_weight)[i];
        const float z = bf2f(in_proj[conv_dim + hv * a.head_v_dim + i]);
        const float zs = act_f32(UZU_ACT_SILU, z);
        reinterpret_cast<__nv_bfloat16*>(a.out)[hv * a.head_v_dim + i] = f2bf(ov * inv_rms * nw * zs);
    }
}

}  // namespace uzu

extern "C" {

void uzu_delta_net_conv_update_encode(uzu_command_buffer* cmd, const uzu_delta_net_conv_update_args* a) {
    if (!uzu::encodable(cmd, "delta_net_conv_update")) return;
    if (!a->conv_weight || !a->in_out || !a->state || a->kernel_size < 2 || (a->has_bias && !a->bias)) {
        cmd->record_error(UZU_ERROR_INVALID_ARGUMENT, "delta_net_conv_update: inconsistent arguments");
        return;
    }
    if (a->conv_dim == 0) return;
    uzu::launch(cmd, "delta_net_conv_update_kernel", uzu::delta_net_conv_update_kernel, dim3((a->conv_dim + 255) / 256), dim3(256), 0, *a);
}

static const char* delta_net_update_check(const uzu_delta_net_update_args* a) {
    if (!a->in_proj || !a->a_log || !a->dt_bias || !a->norm_weight || !a->state || !a->out) return "null operand";
    if (a->head_k_dim != 128 || a->head_v_dim == 0 || a->head_v_dim > 256 || a->num_k_heads == 0 || a->num_v_heads % a->num_k_heads != 0)
        return "HEAD_K_DIM must be 128 and head_v_dim <= 256";
    if ((a->head_v_dim + uzu::DN_CLUSTER - 1) / uzu::DN_CLUSTER > uzu::DN_WARPS * 32 - 128) return "head_v_dim too large for the v staging threads";
    return nullptr;
}

void uzu_delta_net_update_encode(uzu_command_buffer* cmd, const uzu_delta_net_update_args* a) {
    if (!uzu::encodable(cmd, "delta_net_update")) return;
    if (const char* err = delta_net_update_check(a)) {
        cmd->record_error(UZU_ERROR_INVALID_ARGUMENT, std::string("delta_net_update: ") + err);
        return;
    }
    uzu_delta_net_fused_update_args fa{};
    fa.update = *a;
    uzu::launch(cmd, "delta_net_update_kernel", uzu::delta_net_update_kernel<false>, dim3(a->num_v_heads * uzu::DN_CLUSTER), dim3(uzu::DN_WARPS * 32), 0, fa);
}

int uzu_delta_net_fused_update_supported(const uzu_delta_net_fused_update_args* f) {
    if (!f || delta_net_update_check(&f->update)) return 0;
    const uzu_delta_net_conv_update_args& c = f->conv;
    if (!c.conv_weight || !c.state || c.kernel_size < 2 || c.kernel_size - 1 > uzu::DN_MAX_TAPS || (c.has_bias && !c.bias)) return 0;
    if (c.in_out != f->update.in_proj || c.conv_dim != 2 * f->update.key_dim + f->update.value_dim) return 0;
    if (f->update.num_v_heads != f->update.num_k_heads) return 0;       // a k head's conv state must have exactly one reader cluster
    if (f->update.key_dim != f->update.num_k_heads * 128 || f->update.value_dim != f->update.num_v_heads * f->update.head_v_dim) return 0;
    return 1;
}

void uzu_delta_net_fused_update_encode(uzu_command_buffer* cmd, const uzu_delta_net_fused_update_args* f) {
    if (!uzu::encodable(cmd, "delta_net_fused_update")) return;
    if (!uzu_delta_net_fused_update_supported(f)) {
        cmd->record_error(UZU_ERROR_UNSUPPORTED, "delta_net_fused_update: geometry not covered (encode DeltaNetConvUpdate + DeltaNetUpdate separately)");
        return;
    }
    uzu::launch(cmd, "delta_net_update_kernel", uzu::delta_net_update_kernel<true>, dim3(f->update.num_v_heads * uzu::DN_CLUSTER), dim3(uzu::DN_WARPS * 32), 0, *f);
}

}  // extern "C"
