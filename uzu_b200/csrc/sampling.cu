// UnifiedSamplingKernel: bitmask -> temperature -> top-k / top-p / min-p -> Gumbel-max -> argmax.
// Spec: backends/cpu/kernel/sampling/unified_sampling.rs:22-98; RNG stream layout
// encodable_block/sampling/gumbel.rs:1-81 (Philox4x32-10, key = 64-bit seed, counter = [offset,0,0,0],
// logit i -> (offset, word) = revidx(i, vocab)). Counter-based, so any thread layout reproduces the
// reference's per-logit noise; ties resolve to the lowest index (unified_sampling.rs:90-95).
//
// Two code paths:
//   * no filters (greedy, or temperature + Gumbel): multi-CTA argmax over a packed (value, ~index)
//     64-bit key, last CTA (ticket) merges -> one launch, ~V*2 bytes read at HBM/L2 speed;
//   * with top-k / top-p / min-p: the kept set is a prefix of the (value desc, index asc) order and each
//     condition is monotone along that order, so "kept" is a per-element predicate on (rank, mass before,
//     value). One CTA per row iterates candidates: take the best post-noise element of the current set,
//     compute its rank / preceding mass in one pass, accept if it passes the filters, otherwise shrink the
//     set to the elements ranked before it (same idea as the reference's Metal kernel,
//     metal/kernel/sampling/unified_sampling.metal:154-238).
#include "common.cuh"

namespace uzu {

__device__ __forceinline__ void philox_round(uint32_t (&c)[4], const uint32_t (&k)[2]) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k[0], n1 = lo1, n2 = hi0 ^ c[3] ^ k[1], n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}

__device__ __forceinline__ float gumbel_noise(uint64_t seed, uint32_t i, uint32_t vocab) {
    // revidx (gumbel.rs:69-81)
    const uint32_t thread_idx = i % 1024u;
    const uint32_t thread_offset = ((vocab + 4095u) / 4096u) * thread_idx;
    const uint32_t block_idx = i / 1024u;
    const uint32_t offset = thread_offset + block_idx / 4u, word = block_idx % 4u;
    uint32_t c[4] = {offset, 0u, 0u, 0u};
    uint32_t k[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    philox_round(c, k);
#pragma unroll
    for (int r = 0; r < 9; ++r) {
        k[0] += 0x9E3779B9u;
        k[1] += 0xBB67AE85u;
        philox_round(c, k);
    }
    const uint32_t w = c[word] >> 8;
    const float u = (float)(w < 1u ? 1u : w) * (1.0f / 16777216.0f);  // unit_interval (gumbel.rs:55-57)
    return -logf(-logf(u));
}

__device__ __forceinline__ float filtered_logit(const uzu_unified_sampling_args& a, const __nv_bfloat16* logits, const uint32_t* bitmask,
                                                float recip_t, uint32_t i) {
    float l = bf2f(logits[i]);
    if (a.has_bitmask && ((bitmask[i >> 5] >> (i & 31)) & 1u) == 0u) l = -INFINITY;
    if (a.has_temperature) l *= recip_t;
    return l;
}

// total order: larger value first, then lower index
__device__ __forceinline__ unsigned long long pack_key(float v, uint32_t i) {
    uint32_t u = __float_as_uint(v);
    if (v != v) u = 0u;  // NaN never wins (the reference's partial_cmp -> Equal keeps the earlier element)
    else u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
    return ((unsigned long long)u << 32) | (unsigned long long)(0xFFFFFFFFu - i);
}
__device__ __forceinline__ unsigned long long warp_max_u64(unsigned long long k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, k, o);
        k = other > k ? other : k;
    }
    return k;
}

constexpr uint32_t SAMPLING_WS_SLOTS = 7168;   // u64 partial keys
constexpr uint32_t SAMPLING_MAX_BLOCKS = 28;
constexpr uint32_t SAMPLING_ROWS_PER_LAUNCH = 256;

__global__ void __launch_bounds__(1024) sampling_argmax_kernel(const uzu_unified_sampling_args a, uint32_t row0, unsigned long long* ws,
                                                               unsigned int* tickets) {
    __shared__ unsigned long long red[32];
    __shared__ unsigned int sm_ticket;
    pdl_launch_dependents();
    pdl_wait();
    const uint32_t row = row0 + blockIdx.y, V = a.vocab_size;
    const __nv_bfloat16* logits = reinterpret_cast<const __nv_bfloat16*>(a.logits) + (size_t)row * V;
    const uint32_t* bitmask = a.has_bitmask ? reinterpret_cast<const uint32_t*>(a.bitmask) + (size_t)row * ((V + 31) / 32) : nullptr;
    const float recip_t = a.has_temperature ? 1.0f / a.temperature : 1.0f;
    const uint64_t seed = a.is_stochastic ? reinterpret_cast<const unsigned long long*>(a.seeds)[row] : 0ull;
    unsigned long long best = 0ull;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < V; i += gridDim.x * blockDim.x) {
        float v = filtered_logit(a, logits, bitmask, recip_t, i);
        if (a.is_stochastic) v += gumbel_noise(seed, i, V);
        const unsigned long long k = pack_key(v, i);
        best = k > best ? k : best;
    }
    best = warp_max_u64(best);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) red[warp] = best;
    __syncthreads();
    if (warp == 0) {
        best = red[lane];
        best = warp_max_u64(best);
        if (lane == 0) {
            if (gridDim.x == 1) {
                reinterpret_cast<uint32_t*>(a.output)[row] = V ? 0xFFFFFFFFu - (uint32_t)(best & 0xFFFFFFFFull) : 0u;
            } else {
                ws[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = best;
                __threadfence();
                sm_ticket = atomicAdd(&tickets[blockIdx.y], 1u);
            }
        }
    }
    if (gridDim.x == 1) return;
    __syncthreads();
    if (sm_ticket != gridDim.x - 1 || warp != 0) return;
    __threadfence();
    unsigned long long k = lane < gridDim.x ? __ldcg(&ws[(size_t)blockIdx.y * gridDim.x + lane]) : 0ull;
    k = warp_max_u64(k);
    if (lane == 0) {
        reinterpret_cast<uint32_t*>(a.output)[row] = 0xFFFFFFFFu - (uint32_t)(k & 0xFFFFFFFFull);
        tickets[blockIdx.y] = 0;
    }
}

// "ranked before c": strictly better in (value desc, index asc)
__device__ __forceinline__ bool ranked_before(float lj, uint32_t j, float lc, uint32_t c) { return lj > lc || (lj == lc && j < c); }

__global__ void __launch_bounds__(1024) sampling_filtered_kernel(const uzu_unified_sampling_args a) {
    __shared__ float redf[32];
    __shared__ unsigned long long redk[32];
    __shared__ unsigned int redc[32];
    __shared__ unsigned long long s_key;
    __shared__ float s_f;
    __shared__ unsigned int s_c;
    const uint32_t row = blockIdx.x, V = a.vocab_size;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const __nv_bfloat16* logits = reinterpret_cast<const __nv_bfloat16*>(a.logits) + (size_t)row * V;
    const uint32_t* bitmask = a.has_bitmask ? reinterpret_cast<const uint32_t*>(a.bitmask) + (size_t)row * ((V + 31) / 32) : nullptr;
    const float recip_t = a.has_temperature ? 1.0f / a.temperature : 1.0f;
    const uint64_t seed = a.is_stochastic ? reinterpret_cast<const unsigned long long*>(a.seeds)[row] : 0ull;

    auto block_max_key = [&](unsigned long long k) {
        k = warp_max_u64(k);
        __syncthreads();
        if (lane == 0) redk[warp] = k;
        __syncthreads();
        if (warp == 0) {
            unsigned long long t = lane < nw ? redk[lane] : 0ull;
            t = warp_max_u64(t);
            if (lane == 0) s_key = t;
        }
        __syncthreads();
        return s_key;
    };
    auto block_sum_f = [&](float v) {
        v = warp_sum(v);
        __syncthreads();
        if (lane == 0) redf[warp] = v;
        __syncthreads();
        if (warp == 0) {
            float t = lane < nw ? redf[lane] : 0.0f;
            t = warp_sum(t);
            if (lane == 0) s_f = t;
        }
        __syncthreads();
        return s_f;
    };
    auto block_sum_u = [&](unsigned int v) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        __syncthreads();
        if (lane == 0) redc[warp] = v;
        __syncthreads();
        if (warp == 0) {
            unsigned int t = lane < nw ? redc[lane] : 0u;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
            if (lane == 0) s_c = t;
        }
        __syncthreads();
        return s_c;
    };

    // pass 0: max of the filtered logits, softmax normaliser, best post-noise candidate over everything
    unsigned long long kmax = 0ull, kcand = 0ull;
    for (uint32_t i = threadIdx.x; i < V; i += blockDim.x) {
        const float l = filtered_logit(a, logits, bitmask, recip_t, i);
        const unsigned long long k = pack_key(l, i);
        kmax = k > kmax ? k : kmax;
        const float v = a.is_stochastic ? l + gumbel_noise(seed, i, V) : l;
        const unsigned long long kc = pack_key(v, i);
        kcand = kc > kcand ? kc : kcand;
    }
    kmax = block_max_key(kmax);
    kcand = block_max_key(kcand);
    const uint32_t imax = 0xFFFFFFFFu - (uint32_t)(kmax & 0xFFFFFFFFull);
    const float lmax = filtered_logit(a, logits, bitmask, recip_t, imax);
    float part = 0.0f;
    for (uint32_t i = threadIdx.x; i < V; i += blockDim.x) part += expf(filtered_logit(a, logits, bitmask, recip_t, i) - lmax);
    const float norm = block_sum_f(part);
    const float min_p_thr = a.has_min_p ? lmax + logf(a.min_p) : -INFINITY;

    uint32_t cand = 0xFFFFFFFFu - (uint32_t)(kcand & 0xFFFFFFFFull);
    uint32_t result = 0;
    for (uint32_t iter = 0; iter <= V; ++iter) {
        const float lc = filtered_logit(a, logits, bitmask, recip_t, cand);
        // one pass: rank and preceding mass of `cand`, plus the best post-noise element ranked before it
        unsigned int cnt = 0;
        float mass = 0.0f;
        unsigned long long knext = 0ull;
        for (uint32_t j = threadIdx.x; j < V; j += blockDim.x) {
            const float lj = filtered_logit(a, logits, bitmask, recip_t, j);
            if (ranked_before(lj, j, lc, cand)) {
                cnt++;
                mass += expf(lj - lmax) / norm;
                const float v = a.is_stochastic ? lj + gumbel_noise(seed, j, V) : lj;
                const unsigned long long k = pack_key(v, j);
                knext = k > knext ? k : knext;
            }
        }
        const unsigned int rank = block_sum_u(cnt);
        const float mass_before = block_sum_f(mass);
        knext = block_max_key(knext);
        const bool kept = !((a.has_top_k && rank >= a.top_k) || (a.has_top_p && mass_before >= a.top_p) || (a.has_min_p && lc < min_p_thr));
        if (kept) { result = cand; break; }
        if (rank == 0) { result = 0; break; }   // nothing survives the filters: all -inf, argmax = index 0
        cand = 0xFFFFFFFFu - (uint32_t)(knext & 0xFFFFFFFFull);
    }
    if (threadIdx.x == 0) reinterpret_cast<uint32_t*>(a.output)[row] = result;
}

}  // namespace uzu

extern "C" {

void uzu_unified_sampling_encode(uzu_command_buffer* cmd, const uzu_unified_sampling_args* a) {
    if (!uzu::encodable(cmd, "unified_sampling")) return;
    if (!a->logits || !a->output || (a->is_stochastic && !a->seeds) || (a->has_bitmask && !a->bitmask) || a->vocab_size == 0 ||
        (a->has_temperature && !(a->temperature > 0.0f))) {
        cmd->record_error(UZU_ERROR_INVALID_ARGUMENT, "unified_sampling: inconsistent arguments");
        return;
    }
    if (a->batch_size == 0) return;
    uzu_context* ctx = cmd->ctx;
    if (a->has_top_k || a->has_top_p || a->has_min_p) {
        uzu::sampling_filtered_kernel<<<a->batch_size, 1024, 0, ctx->stream>>>(*a);
        uzu::after_launch(cmd, "sampling_filtered_kernel");
        return;
    }
    uint32_t nblocks = std::min(uzu::SAMPLING_MAX_BLOCKS, (a->vocab_size + 8191u) / 8192u);
    nblocks = std::max(nblocks, 1u);
    unsigned int* tickets = reinterpret_cast<unsigned int*>(ctx->sampling_ws + uzu::SAMPLING_WS_SLOTS);
    for (uint32_t row0 = 0; row0 < a->batch_size; row0 += uzu::SAMPLING_ROWS_PER_LAUNCH) {
        const uint32_t rows = std::min(uzu::SAMPLING_ROWS_PER_LAUNCH, a->batch_size - row0);
        dim3 grid(nblocks, rows);
        uzu::launch(cmd, "sampling_argmax_kernel", uzu::sampling_argmax_kernel, grid, dim3(1024), 0, *a, row0, ctx->sampling_ws, tickets);
    }
}

}  // extern "C"
