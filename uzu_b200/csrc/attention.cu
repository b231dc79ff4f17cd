// Attention kernels: AttentionPrepare (RoPE + Q transpose + KV append), split-KV flash-decode attention
// behind the AttentionSinglePass / AttentionTwoPass1 / AttentionTwoPass2 entry points, KVCacheUpdate,
// SigmoidGate.
//
// Specs: backends/cpu/kernel/attention/{attention_prepare.rs:7-126, attention_single_pass.rs:49-126,
// attention_two_pass.rs:55-189, mask.rs:3-62, kv_cache_update.rs:7-28, sigmoid_gate.rs:7-22}.
//
// Decode attention design (HBM-bound: 2*ctx*Hkv*D*2 bytes per layer per token):
//   * one CTA per (kv head, query token, kv split); all G = Hq/Hkv query heads of the group are processed
//     against each K/V row while it is in registers, so the cache is read from HBM once, not G times;
//   * a K/V row (D bf16) is read by D/EPL adjacent lanes with 128-bit loads (whole 32 B sectors),
//     several keys per warp per step; f32 scores, f32 online softmax (expf, not exp2, to stay on the
//     reference's arithmetic), f32 output accumulators;
//   * the KV range is split across up to 32 CTAs so a batch-1 decode still fills 148 SMs; partial
//     (max, sum, o) triples merge either in the caller-visible two-pass buffers (TwoPass1 + TwoPass2
//     API) or, for the single-pass entry point, through a stream-ordered workspace where the last
//     CTA of a (kv head, token) merges (one launch).
#include <cfloat>

#include "common.cuh"

namespace uzu {

// ---------------------------------------------------------------------------------------------------
// AttentionPrepare
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) attention_prepare_kernel(const uzu_attention_prepare_args a) {
    pdl_launch_dependents();
    pdl_wait();
    const uint32_t total_heads = a.has_kv ? a.num_q_heads + 2 * a.num_kv_heads : a.num_q_heads;
    const uint32_t b = blockIdx.y, h = blockIdx.x;
    const __nv_bfloat16* qkv = reinterpret_cast<const __nv_bfloat16*>(a.qkv);
    const __nv_bfloat16* head = qkv + ((size_t)b * total_heads + h) * a.head_dim;
    const bool is_query = !a.has_kv || h < a.num_q_heads;
    const bool is_key = a.has_kv && h >= a.num_q_heads && h < a.num_q_heads + a.num_kv_heads;
    const float* cosines = reinterpret_cast<const float*>(a.cosines);
    const float* sines = reinterpret_cast<const float*>(a.sines);
    uint32_t kv_token_offset = a.kv_token_offset;
    if (a.dynamic_position) {
        kv_token_offset = *reinterpret_cast<const uint32_t*>(a.dynamic_position);
        if (a.has_rope) {
            cosines += (size_t)kv_token_offset * a.rope_dim;
            sines += (size_t)kv_token_offset * a.rope_dim;
        }
    }
    for (uint32_t d = threadIdx.x; d < a.head_dim; d += blockDim.x) {
        __nv_bfloat16 e = head[d];
        if (a.has_rope && d < a.rope_dim && (is_query || is_key)) {
            const uint32_t half = a.rope_dim / 2;
            const uint32_t paired = d < half ? d + half : d - half;
            const float input = bf2f(e);
            const float pv = bf2f(head[paired]);
            const float signed_p = d < half ? -pv : pv;
            const float c = cosines[(size_t)b * a.rope_dim + d];
            const float s = sines[(size_t)b * a.rope_dim + d];
            e = f2bf(input * c + signed_p * s);
        }
        if (is_query) {
            reinterpret_cast<__nv_bfloat16*>(a.queries)[((size_t)h * a.batch_dim + b) * a.head_dim + d] = e;
        } else if (is_key) {
            reinterpret_cast<__nv_bfloat16*>(a.keys)[((size_t)(kv_token_offset + b) * a.num_kv_heads + (h - a.num_q_heads)) * a.head_dim + d] = e;
        } else {
            reinterpret_cast<__nv_bfloat16*>(a.values)[((size_t)(kv_token_offset + b) * a.num_kv_heads + (h - a.num_q_heads - a.num_kv_heads)) * a.head_dim + d] = e;
        }
    }
}

// QKVNorm(q) + QKVNorm(k) + AttentionPrepare in one launch: one CTA per (head, token). The RMS statistic is summed by warp 0 in the
// standalone kernel's order (lane-strided, then the xor tree), the normalised row is rounded to bf16 as the in-place QKVNorm
// stores it, and RoPE reads that bf16 row from shared memory: bit-identical to the three-kernel sequence.
__global__ void __launch_bounds__(256) attention_prepare_norm_kernel(const uzu_attention_prepare_norm_args n) {
    __shared__ __nv_bfloat16 row[256];
    __shared__ float rms_s;
    pdl_launch_dependents();
    pdl_wait();
    const uzu_attention_prepare_args& a = n.prepare;
    const uint32_t total_heads = a.has_kv ? a.num_q_heads + 2 * a.num_kv_heads : a.num_q_heads;
    const uint32_t b = blockIdx.y, h = blockIdx.x;
    const __nv_bfloat16* head = reinterpret_cast<const __nv_bfloat16*>(a.qkv) + ((size_t)b * total_heads + h) * a.head_dim;
    const bool is_query = !a.has_kv || h < a.num_q_heads;
    const bool is_key = a.has_kv && h >= a.num_q_heads && h < a.num_q_heads + a.num_kv_heads;
    const uzu_qk_norm_config& nc = is_query ? n.q_norm : n.k_norm;
    const bool normed = (is_query || is_key) && nc.present;
    const float* cosines = reinterpret_cast<const float*>(a.cosines);
    const float* sines = reinterpret_cast<const float*>(a.sines);
    uint32_t kv_token_offset = a.kv_token_offset;
    if (a.dynamic_position) {
        kv_token_offset = *reinterpret_cast<const uint32_t*>(a.dynamic_position);
        if (a.has_rope) {
            cosines += (size_t)kv_token_offset * a.rope_dim;
            sines += (size_t)kv_token_offset * a.rope_dim;
        }
    }
    if (normed && threadIdx.x < 32) {
        float total = 0.0f;
        for (uint32_t i = threadIdx.x; i < a.head_dim; i += 32) {
            const float v = bf2f(head[i]);
            total += v * v;
        }
        total = warp_sum(total);
        const float mean_square = total / (float)a.head_dim;
        if (threadIdx.x == 0) rms_s = 1.0f / sqrtf(mean_square + nc.epsilon);
    }
    __syncthreads();
    for (uint32_t d = threadIdx.x; d < a.head_dim; d += blockDim.x) {
        __nv_bfloat16 r = head[d];
        if (normed) {
            const float normalized = bf2f(r) * rms_s;
            const float* scales = reinterpret_cast<const float*>(nc.scales);
            if (!nc.has_scales) r = f2bf(normalized);
            else if (nc.full_layer) r = f2bf(normalized * (scales[d] + nc.scale_offset));
            else r = f2bf(bf2f(f2bf(normalized)) * bf2f(f2bf(scales[d] + nc.scale_offset)));
        }
        row[d] = r;
    }
    __syncthreads();
    for (uint32_t d = threadIdx.x; d < a.head_dim; d += blockDim.x) {
        __nv_bfloat16 e = row[d];
        if (a.has_rope && d < a.rope_dim && (is_query || is_key)) {
            const uint32_t half = a.rope_dim / 2;
            const uint32_t paired = d < half ? d + half : d - half;
            const float input = bf2f(e);
            const float pv = bf2f(row[paired]);
            const float signed_p = d < half ? -pv : pv;
            const float c = cosines[(size_t)b * a.rope_dim + d];
            const float s = sines[(size_t)b * a.rope_dim + d];
            e = f2bf(input * c + signed_p * s);
        }
        if (is_query) {
            reinterpret_cast<__nv_bfloat16*>(a.queries)[((size_t)h * a.batch_dim + b) * a.head_dim + d] = e;
        } else if (is_key) {
            reinterpret_cast<__nv_bfloat16*>(a.keys)[((size_t)(kv_token_offset + b) * a.num_kv_heads + (h - a.num_q_heads)) * a.head_dim + d] = e;
        } else {
            reinterpret_cast<__nv_bfloat16*>(a.values)[((size_t)(kv_token_offset + b) * a.num_kv_heads + (h - a.num_q_heads - a.num_kv_heads)) * a.head_dim + d] = e;
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// mask (mask.rs:3-62)
// ---------------------------------------------------------------------------------------------------
struct AttnParams {
    uzu_attention_args a;
    float* part_o;      // [suffix, heads, nb, D]
    float* part_sum;    // [suffix, heads, nb]
    float* part_max;
    __nv_bfloat16* final_out;  // [suffix, heads, D] when merging in-kernel
    unsigned int* counters;    // one per (token, head group) when nsplits > 1 and merging in-kernel
    uint32_t nsplits, keys_per_split, nb, fuse_merge, heads_per_cta_group;
};

__device__ __forceinline__ bool should_use_key(const uzu_attention_args& a, uint32_t q_seq_idx, uint32_t prefix_length,
                                               uint32_t suffix_position, uint32_t query_position, uint32_t i) {
    bool use_key = true;
    uint32_t key_position;
    if (i >= prefix_length) {
        const uint32_t kis = i - prefix_length;
        if (a.is_trie) {
            const uzu_trie_node node = reinterpret_cast<const uzu_trie_node*>(a.trie)[kis];
            key_position = suffix_position + node.height;
            if (a.is_causal) use_key &= (q_seq_idx >= node.trie_start && q_seq_idx <= node.trie_end);
        } else {
            key_position = suffix_position + kis;
            if (a.is_causal) use_key &= (kis <= q_seq_idx);
        }
    } else {
        if (a.is_kv_cache_ring) {
            key_position = (prefix_length + i - a.ring_params.ring_offset) % prefix_length;
            use_key &= key_position < a.ring_params.ring_length;
        } else {
            key_position = i;
        }
    }
    if (a.is_sliding_window) {
        const uint32_t w = a.sliding_window_size;
        if (a.is_causal) use_key &= (key_position <= query_position && (query_position - key_position) < w);
        else if (key_position <= query_position) use_key &= ((query_position - key_position) <= w / 2);
        else use_key &= ((key_position - query_position) <= w / 2);
    }
    return use_key;
}

__device__ __forceinline__ float safe_exp_diff(float m, float gm) { return (m == -INFINITY) ? 0.0f : expf(m - gm); }

constexpr int ATTN_WARPS = 4;
constexpr int ATTN_CHUNK = 512;     // keys whose scores live in shared memory at once
constexpr int ATTN_UNROLL = 2;      // key rows per lane group per round (one more round is always in flight)

// Split-KV attention for short suffixes (decode) and, with nsplits = 1, for any suffix. One CTA = (G query heads sharing a KV
// head, one query token, one key range). Per chunk of <= ATTN_CHUNK keys: (1) scores q.k for all keys of the chunk, K rows
// loaded ATTN_UNROLL at a time with no dependency between them; (2) chunk max / exp / sum per head; (3) o += p.V with the V
// rows streamed the same way. The running (m, l, o) is rescaled only once per chunk, so the key loop has no serial
// softmax chain (the first version of this kernel updated (m, l, o) per key and spent most of its time waiting on loads).
template <int D, int G>
__global__ void __launch_bounds__(ATTN_WARPS * 32) attn_split_kernel(const AttnParams p) {
    constexpr int EPL = (D >= 128 && G < 8) ? 16 : 8;    // elements per lane (q and o live in registers: 2 * G * EPL floats)
    constexpr int CH = (G >= 8) ? ATTN_CHUNK / 2 : ATTN_CHUNK;   // scores [G][CH] + partial outputs stay under the 48 KB static limit
    constexpr int LPK = D / EPL;              // lanes per key row
    constexpr int KPW = 32 / LPK;             // key rows per warp step
    constexpr int STEP = ATTN_WARPS * KPW;    // key rows per CTA step
    constexpr int U = ATTN_UNROLL;
    static_assert(LPK >= 1 && LPK <= 32 && (EPL % 8) == 0, "bad attention tiling");
    __shared__ float sc[G][CH];
    __shared__ float sm_o[ATTN_WARPS * KPW][G][D];     // one partial output per lane group (<= 32 KB for every instantiation)
    __shared__ float sm_m[G], sm_lc[G], sm_mfin[G], sm_lfin[G];
    __shared__ unsigned int sm_ticket;
    pdl_launch_dependents();
    pdl_wait();
    uzu_attention_args a = p.a;
    uint32_t keys_per_split = p.keys_per_split;
    if (a.dynamic_position) {
        a.sequence_length = *reinterpret_cast<const uint32_t*>(a.dynamic_position) + a.suffix_length;
        keys_per_split = ((a.sequence_length + p.nsplits - 1) / p.nsplits + 15u) & ~15u;   // whole CTA steps per split
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane / LPK, li = lane % LPK, d0 = li * EPL;
    const uint32_t head0 = blockIdx.x * G, qs = blockIdx.y, split = blockIdx.z;
    const uint32_t kvh = head0 / a.gqa_factor;

    const uint32_t prefix_length = a.sequence_length - a.suffix_length;
    const uint32_t suffix_position = a.is_kv_cache_ring ? a.ring_params.ring_length : prefix_length;
    const uint32_t query_position = a.is_trie ? suffix_position + reinterpret_cast<const uzu_trie_node*>(a.trie)[qs].height
                                              : suffix_position + qs;

    const __nv_bfloat16* keys = reinterpret_cast<const __nv_bfloat16*>(a.keys) + (size_t)kvh * a.k_head_stride + d0;
    const __nv_bfloat16* values = reinterpret_cast<const __nv_bfloat16*>(a.values) + (size_t)kvh * a.v_head_stride + d0;

    float q[G][EPL];
#pragma unroll
    for (int h = 0; h < G; ++h) {
        const __nv_bfloat16* qp = reinterpret_cast<const __nv_bfloat16*>(a.queries) + ((size_t)(head0 + h) * a.suffix_length + qs) * D + d0;
#pragma unroll
        for (int v = 0; v < EPL / 8; ++v) {
            const uint4 raw = *reinterpret_cast<const uint4*>(qp + v * 8);
            const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                q[h][v * 8 + 2 * e] = a.scale * __low2float(h2[e]);
                q[h][v * 8 + 2 * e + 1] = a.scale * __high2float(h2[e]);
            }
        }
    }
    // running softmax state: (m, l) identical in every thread, o partial per lane group (summed at the end)
    float m_run[G], l_run[G], o[G][EPL];
#pragma unroll
    for (int h = 0; h < G; ++h) {
        m_run[h] = -INFINITY;
        l_run[h] = 0.0f;
#pragma unroll
        for (int e = 0; e < EPL; ++e) o[h][e] = 0.0f;
    }
    if (a.has_sinks && split == 0) {
#pragma unroll
        for (int h = 0; h < G; ++h) {
            m_run[h] = bf2f(reinterpret_cast<const __nv_bfloat16*>(a.sinks)[head0 + h]);
            l_run[h] = 1.0f;
        }
    }

    // no ring / trie / window: every key below sequence_length is visible except the causal part of the suffix itself
    const bool plain_mask = !a.is_kv_cache_ring && !a.is_trie && !a.is_sliding_window;
    auto key_visible = [&](uint32_t i) {
        if (plain_mask) return !(a.is_causal && i >= prefix_length && (i - prefix_length) > qs);
        return should_use_key(a, qs, prefix_length, suffix_position, query_position, i);
    };
    // one round = U key rows per lane group; the rows of round r + 1 are requested before round r is consumed
    auto load_round = [&](const __nv_bfloat16* base, uint32_t seq_stride, uint32_t cbeg, uint32_t n, uint32_t kb, bool masked,
                          uint4 (&dst)[U][EPL / 8], bool (&ok)[U]) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t kl = kb + u * STEP + sub, i = cbeg + kl;
            ok[u] = kl < n && (!masked || key_visible(i));
#pragma unroll
            for (int v = 0; v < EPL / 8; ++v)
                dst[u][v] = ok[u] ? *reinterpret_cast<const uint4*>(base + (size_t)i * seq_stride + v * 8) : make_uint4(0, 0, 0, 0);
        }
    };

    const uint32_t begin = min(a.sequence_length, split * keys_per_split);
    const uint32_t end = min(a.sequence_length, begin + keys_per_split);
    for (uint32_t cbeg = begin; cbeg < end; cbeg += CH) {
        const uint32_t n = min((uint32_t)CH, end - cbeg);
        uint4 cur[U][EPL / 8], nxt[U][EPL / 8];
        bool okc[U], okn[U];
        // ---- (1) scores ------------------------------------------------------------------------------------
        load_round(keys, a.k_seq_stride, cbeg, n, warp * KPW, true, cur, okc);
        for (uint32_t kb = warp * KPW; kb < n; kb += STEP * U) {
            const bool more = kb + STEP * U < n;
            if (more) load_round(keys, a.k_seq_stride, cbeg, n, kb + STEP * U, true, nxt, okn);
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const uint32_t kl = kb + u * STEP + sub;
                float kf[EPL];
#pragma unroll
                for (int v = 0; v < EPL / 8; ++v) {
                    const __nv_bfloat162* k2 = reinterpret_cast<const __nv_bfloat162*>(&cur[u][v]);
#pragma unroll
                    for (int e = 0; e < 4; ++e) { kf[v * 8 + 2 * e] = __low2float(k2[e]); kf[v * 8 + 2 * e + 1] = __high2float(k2[e]); }
                }
#pragma unroll
                for (int h = 0; h < G; ++h) {
                    float sdot = 0.0f;
#pragma unroll
                    for (int e = 0; e < EPL; ++e) sdot = fmaf(q[h][e], kf[e], sdot);   // this file is built with -fmad=false: ask for the FMA
#pragma unroll
                    for (int off = LPK / 2; off > 0; off >>= 1) sdot += __shfl_xor_sync(0xffffffffu, sdot, off);
                    if (li == 0 && kl < n) sc[h][kl] = okc[u] ? sdot : -INFINITY;
                }
            }
            if (more) {
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    okc[u] = okn[u];
#pragma unroll
                    for (int v = 0; v < EPL / 8; ++v) cur[u][v] = nxt[u][v];
                }
            }
        }
        // the first V rows travel while the chunk statistics are computed
        load_round(values, a.v_seq_stride, cbeg, n, warp * KPW, false, cur, okc);
        __syncthreads();
        // ---- (2) chunk max, probabilities relative to the new running max, chunk sum (one warp per head) -----
#pragma unroll
        for (int h = 0; h < G; ++h) {
            if ((h % ATTN_WARPS) != warp) continue;      // static register indexing of m_run
            float mc = -INFINITY;
            for (uint32_t kl = lane; kl < n; kl += 32) mc = fmaxf(mc, sc[h][kl]);
            mc = warp_max(mc);
            const float m_new = fmaxf(m_run[h], mc);
            float lc = 0.0f;
            for (uint32_t kl = lane; kl < n; kl += 32) {
                const float sv = sc[h][kl];
                const float pv = (sv == -INFINITY) ? 0.0f : expf(sv - m_new);
                sc[h][kl] = pv;
                lc += pv;
            }
            lc = warp_sum(lc);
            if (lane == 0) { sm_m[h] = m_new; sm_lc[h] = lc; }
        }
        __syncthreads();
#pragma unroll
        for (int h = 0; h < G; ++h) {
            const float m_new = sm_m[h];
            const float factor = safe_exp_diff(m_run[h], m_new);
            l_run[h] = l_run[h] * factor + sm_lc[h];
            m_run[h] = m_new;
#pragma unroll
            for (int e = 0; e < EPL; ++e) o[h][e] *= factor;
        }
        // ---- (3) o += p . V ---------------------------------------------------------------------------------
        for (uint32_t kb = warp * KPW; kb < n; kb += STEP * U) {
            const bool more = kb + STEP * U < n;
            if (more) load_round(values, a.v_seq_stride, cbeg, n, kb + STEP * U, false, nxt, okn);
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const uint32_t kl = kb + u * STEP + sub;
                if (!okc[u]) continue;
                float vf[EPL];
#pragma unroll
                for (int v = 0; v < EPL / 8; ++v) {
                    const __nv_bfloat162* v2 = reinterpret_cast<const __nv_bfloat162*>(&cur[u][v]);
#pragma unroll
                    for (int e = 0; e < 4; ++e) { vf[v * 8 + 2 * e] = __low2float(v2[e]); vf[v * 8 + 2 * e + 1] = __high2float(v2[e]); }
                }
#pragma unroll
                for (int h = 0; h < G; ++h) {
                    const float pv = sc[h][kl];
#pragma unroll
                    for (int e = 0; e < EPL; ++e) o[h][e] = fmaf(pv, vf[e], o[h][e]);
                }
            }
            if (more) {
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    okc[u] = okn[u];
#pragma unroll
                    for (int v = 0; v < EPL / 8; ++v) cur[u][v] = nxt[u][v];
                }
            }
        }
        __syncthreads();     // sc is rewritten by the next chunk
    }

    // ---- every lane group parks its partial output in shared memory; summed below, ATTN_WARPS * KPW terms per output -------
#pragma unroll
    for (int h = 0; h < G; ++h)
#pragma unroll
        for (int v = 0; v < EPL / 4; ++v)
            *reinterpret_cast<float4*>(&sm_o[warp * KPW + sub][h][d0 + v * 4]) = make_float4(o[h][v * 4], o[h][v * 4 + 1], o[h][v * 4 + 2], o[h][v * 4 + 3]);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int h = 0; h < G; ++h) { sm_mfin[h] = m_run[h]; sm_lfin[h] = l_run[h]; }
    }
    __syncthreads();

    const uint32_t H = a.num_heads;
    const bool direct = (p.nsplits == 1) && p.fuse_merge;
    for (int idx = threadIdx.x; idx < G * D; idx += blockDim.x) {
        const int h = idx / D, d = idx % D;
        float ov = 0.0f;
#pragma unroll
        for (int w = 0; w < ATTN_WARPS * KPW; ++w) ov += sm_o[w][h][d];
        const float lv = sm_lfin[h], gm = sm_mfin[h];
        const size_t o_off = (size_t)qs * H + head0 + h;
        if (direct) {
            p.final_out[o_off * D + d] = f2bf(ov / lv);
        } else {
            p.part_o[(o_off * p.nb + split) * D + d] = ov;
            if (d == 0) {
                p.part_sum[o_off * p.nb + split] = lv;
                p.part_max[o_off * p.nb + split] = (gm == -INFINITY) ? -1e9f : gm;  // two-pass init value (attention_two_pass.rs:95)
            }
        }
    }
    if (direct) return;
    if (!p.fuse_merge) {
        // caller-visible two-pass buffers have 32 blocks; blocks this launch did not produce are empty
        if (split == 0) {
            for (uint32_t blk = p.nsplits + warp; blk < p.nb; blk += ATTN_WARPS)
                for (int h = 0; h < G; ++h) {
                    const size_t o_off = (size_t)qs * H + head0 + h;
                    for (int d = lane; d < D; d += 32) p.part_o[(o_off * p.nb + blk) * D + d] = 0.0f;
                    if (lane == 0) { p.part_sum[o_off * p.nb + blk] = 0.0f; p.part_max[o_off * p.nb + blk] = -1e9f; }
                }
        }
        return;
    }
    // ---- in-kernel merge: the last CTA of this (token, head group) combines all splits (nsplits <= 32) -----------------
    __threadfence();
    __syncthreads();
    const uint32_t cidx = qs * gridDim.x + blockIdx.x;
    if (threadIdx.x == 0) sm_ticket = atomicAdd(&p.counters[cidx], 1u);
    __syncthreads();
    if (sm_ticket != p.nsplits - 1) return;
    __threadfence();
    for (int h = warp; h < G; h += ATTN_WARPS) {        // one warp per head: split factors and the global sum
        const size_t o_off = (size_t)qs * H + head0 + h;
        const bool has = (uint32_t)lane < p.nsplits;
        const float mymax = has ? __ldcg(&p.part_max[o_off * p.nb + lane]) : -INFINITY;
        const float gmax = warp_max(mymax);
        const float f = has ? expf(mymax - gmax) : 0.0f;
        const float gsum = warp_sum(has ? __ldcg(&p.part_sum[o_off * p.nb + lane]) * f : 0.0f);
        sc[h][lane] = f;
        if (lane == 0) sm_lfin[h] = gsum;
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < G * D; idx += blockDim.x) {
        const int h = idx / D, d = idx % D;
        const size_t o_off = (size_t)qs * H + head0 + h;
        const float* po = p.part_o + (o_off * p.nb) * D + d;
        float val = 0.0f;
        float pv[32];                                   // all split partials of this output in flight at once
#pragma unroll
        for (uint32_t sp = 0; sp < 32; ++sp) pv[sp] = sp < p.nsplits ? __ldcg(po + (size_t)sp * D) : 0.0f;
#pragma unroll
        for (uint32_t sp = 0; sp < 32; ++sp) val += sp < p.nsplits ? pv[sp] * sc[h][sp] : 0.0f;
        p.final_out[o_off * D + d] = f2bf(val / sm_lfin[h]);
    }
    if (threadIdx.x == 0) p.counters[cidx] = 0;
}

// AttentionTwoPass2 (attention_two_pass.rs:139-189): merge 32 blocks. One warp per (head, token).
__global__ void __launch_bounds__(128) attn_two_pass2_kernel(const uzu_attention_two_pass2_args a) {
    pdl_launch_dependents();
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const uint32_t unit = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (unit >= a.num_heads * a.suffix_length) return;
    const float* partials = reinterpret_cast<const float*>(a.partials);
    const float* sums = reinterpret_cast<const float*>(a.sums);
    const float* maxs = reinterpret_cast<const float*>(a.maxs);
    const size_t o_off = unit;  // = q_seq_idx * num_heads + head_idx
    const float mymax = maxs[o_off * 32 + lane];
    const float gmax = warp_max(mymax);
    const float f = expf(mymax - gmax);
    // global_sum accumulates block 0..31 in order in the reference; a tree sum differs by f32 rounding only
    const float gsum = warp_sum(sums[o_off * 32 + lane] * f);
    const uint32_t D = a.head_dim;
    for (uint32_t j = lane; j < D; j += 32) {
        float val = 0.0f;
        for (int b = 0; b < 32; ++b) {
            const float fb = __shfl_sync(0xffffffffu, f, b);
            val += partials[(o_off * 32 + b) * D + j] * fb;
        }
        reinterpret_cast<__nv_bfloat16*>(a.out)[o_off * D + j] = f2bf(val / gsum);
    }
}

// ---------------------------------------------------------------------------------------------------
// KVCacheUpdate / SigmoidGate
// ---------------------------------------------------------------------------------------------------
struct KvCopies {
    uzu_kv_copy c[512];
};
__global__ void __launch_bounds__(256) kv_cache_update_kernel(__nv_bfloat16* keys, __nv_bfloat16* values, const __grid_constant__ KvCopies copies,
                                                             uint32_t copy_count, uint32_t element_dim) {
    // The reference applies copies sequentially per element (copy i may read a row written by copy j < i);
    // one thread owns one element and walks the copies in order, which preserves that.
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= element_dim) return;
    for (uint32_t i = 0; i < copy_count; ++i) {
        const size_t s = (size_t)copies.c[i].source * element_dim + e, d = (size_t)copies.c[i].destination * element_dim + e;
        keys[d] = keys[s];
        values[d] = values[s];
    }
}

__global__ void __launch_bounds__(256) sigmoid_gate_kernel(const __nv_bfloat16* gate, __nv_bfloat16* output, uint32_t total) {
    pdl_launch_dependents();
    pdl_wait();
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const float g = bf2f(gate[idx]);
    const float sg = 1.0f / (1.0f + expf(-g));
    output[idx] = f2bf(bf2f(output[idx]) * sg);
}

// ---------------------------------------------------------------------------------------------------
// host dispatch
// ---------------------------------------------------------------------------------------------------
template <int D, int G>
static void launch_attn(uzu_command_buffer* cmd, const AttnParams& p, uint32_t head_groups) {
    dim3 grid(head_groups, p.a.suffix_length, p.nsplits);
    launch(cmd, "attn_split_kernel", attn_split_kernel<D, G>, grid, dim3(ATTN_WARPS * 32), 0, p);
}

template <int D>
static void launch_attn_g(uzu_command_buffer* cmd, const AttnParams& p, int g) {
    const uint32_t H = p.a.num_heads;
    switch (g) {
        case 8: launch_attn<D, 8>(cmd, p, H / 8); break;
        case 4: launch_attn<D, 4>(cmd, p, H / 4); break;
        case 2: launch_attn<D, 2>(cmd, p, H / 2); break;
        default: launch_attn<D, 1>(cmd, p, H); break;
    }
}

static bool attn_common_checks(uzu_command_buffer* cmd, const uzu_attention_args* a, const char* what) {
    if (!a->queries || !a->keys || !a->values || !a->out) {
        cmd->record_error(UZU_ERROR_INVALID_ARGUMENT, std::string(what) + ": null operand");
        return false;
    }
    if (a->head_dim != 64 && a->head_dim != 128 && a->head_dim != 256) {
        cmd->record_error(UZU_ERROR_UNSUPPORTED, std::string(what) + ": head_dim must be 64, 128 or 256");
        return false;
    }
    if (a->gqa_factor == 0 || a->num_heads % a->gqa_factor != 0 || a->suffix_length == 0 || a->sequence_length < a->suffix_length) {
        cmd->record_error(UZU_ERROR_INVALID_ARGUMENT, std::string(what) + ": bad head / sequence geometry");
        return false;
    }
    if ((a->k_head_stride | a->k_seq_stride | a->v_head_stride | a->v_seq_stride) % 8 != 0 || ((a->keys | a->values | a->queries) & 15)) {
        cmd->record_error(UZU_ERROR_UNSUPPORTED, std::string(what) + ": K/V strides must be multiples of 8 elements and 16-byte aligned");
        return false;
    }
    if ((a->is_trie && !a->trie) || (a->has_sinks && !a->sinks)) {
        cmd->record_error(UZU_ERROR_INVALID_ARGUMENT, std::string(what) + ": missing optional operand");
        return false;
    }
    return true;
}

static int pick_group(uint32_t gqa) {
    if (gqa % 8 == 0) return 8;
    if (gqa % 4 == 0) return 4;
    if (gqa % 2 == 0) return 2;
    return 1;
}

static void dispatch_attn(uzu_command_buffer* cmd, const AttnParams& p, int g) {
    switch (p.a.head_dim) {
        case 64: launch_attn_g<64>(cmd, p, g); break;
        case 128: launch_attn_g<128>(cmd, p, g); break;
        default: launch_attn_g<256>(cmd, p, g); break;
    }
}

}  // namespace uzu

extern "C" {

void uzu_attention_prepare_encode(uzu_command_buffer* cmd, const uzu_attention_prepare_args* a) {
    if (!uzu::encodable(cmd, "attention_prepare")) return;
    if (!a->qkv || !a->queries || a->head_dim == 0 || (a->has_kv && (!a->keys || !a->values || a->num_kv_heads == 0)) ||
        (a->has_rope && (!a->cosines || !a->sines || a->rope_dim == 0 || a->rope_dim > a->head_dim || (a->rope_dim & 1))) ||
        (a->num_q_heads == 0 && !a->has_kv)) {
        cmd->record_error(UZU_ERROR_INVALID_ARGUMENT, "attention_prepare: inconsistent arguments");
        return;
    }
    const uint32_t total_heads = a->has_kv ? a->num_q_heads + 2 * a->num_kv_heads : a->num_q_heads;
    if (a->batch_dim == 0 || total_heads == 0) return;
    dim3 grid(total_heads, a->batch_dim);
    uint32_t threads = a->head_dim >= 256 ? 256 : (a->head_dim >= 128 ? 128 : 64);
    uzu::launch(cmd, "attention_prepare_kernel", uzu::attention_prepare_kernel, grid, dim3(threads), 0, *a);
}

void uzu_attention_prepare_norm_encode(uzu_command_buffer* cmd, const uzu_attention_prepare_norm_args* n) {
    if (!uzu::encodable(cmd, "attention_prepare_norm")) return;
    const uzu_attention_prepare_args* a = &n->prepare;
    if (!a->qkv || !a->queries || a->head_dim == 0 || (a->has_kv && (!a->keys || !a->values || a->num_kv_heads == 0)) ||
        (a->has_rope && (!a->cosines || !a->sines || a->rope_dim == 0 || a->rope_dim > a->head_dim || (a->rope_dim & 1))) ||
        (a->num_q_heads == 0 && !a->has_kv) || (n->q_norm.present && n->q_norm.has_scales && !n->q_norm.scales) ||
        (n->k_norm.present && n->k_norm.has_scales && !n->k_norm.scales)) {
        cmd->record_error(UZU_ERROR_INVALID_ARGUMENT, "attention_prepare_norm: inconsistent arguments");
        return;
    }
    if (a->head_dim > 256) {
        cmd->record_error(UZU_ERROR_UNSUPPORTED, "attention_prepare_norm: head_dim must be <= 256 (encode QKVNorm + AttentionPrepare separately)");
        return;
    }
    const uint32_t total_heads = a->has_kv ? a->num_q_heads + 2 * a->num_kv_heads : a->num_q_heads;
    if (a->batch_dim == 0 || total_heads == 0) return;
    const uint32_t threads = a->head_dim >= 256 ? 256 : (a->head_dim >= 128 ? 128 : 64);
    uzu::launch(cmd, "attention_prepare_norm_kernel", uzu::attention_prepare_norm_kernel, dim3(total_heads, a->batch_dim), dim3(threads), 0, *n);
}

void uzu_attention_single_pass_encode(uzu_command_buffer* cmd, const uzu_attention_args* a) {
    if (!uzu::encodable(cmd, "attention_single_pass")) return;
    if (!uzu::attn_common_checks(cmd, a, "attention_single_pass")) return;
    if (uzu::encode_attention_prefill(cmd, *a)) return;   // opt-in tensor-core path for suffix >= 16 (attention_prefill.cu)
    uzu_context* ctx = cmd->ctx;
    const int g = uzu::pick_group(a->gqa_factor);
    const uint32_t head_groups = a->num_heads / g;
    // enough CTAs for ~2 waves, >= 64 keys per split, <= 32 splits, and only for short suffixes (decode)
    uint32_t nsplits = 1;
    if (a->suffix_length <= 16) {
        const uint32_t ctas = head_groups * a->suffix_length;
        const uint32_t want = (2u * (uint32_t)ctx->sm_count + ctas - 1) / ctas;
        const uint32_t cap = (a->sequence_length + 63) / 64;
        nsplits = std::max(1u, std::min(std::min(want, cap), 32u));
    }
    uzu::AttnParams p{};
    p.a = *a;
    p.final_out = reinterpret_cast<__nv_bfloat16*>(a->out);
    p.fuse_merge = 1;
    p.nsplits = nsplits;
    p.nb = nsplits;
    p.keys_per_split = nsplits > 1 ? (((a->sequence_length + nsplits - 1) / nsplits + 15u) & ~15u) : a->sequence_length;
    if (nsplits > 1) {
        const size_t units = (size_t)a->suffix_length * a->num_heads * nsplits;
        const size_t need = units * (a->head_dim + 2) * sizeof(float);
        if (need > ctx->attn_ws_bytes) {  // grows outside graph capture only (decode needs < 2 MiB)
            cudaStreamSynchronize(ctx->stream);
            cudaFree(ctx->attn_ws);
            if (cudaMalloc(&ctx->attn_ws, need) != cudaSuccess) {
                ctx->attn_ws = nullptr;
                ctx->attn_ws_bytes = 0;
                cmd->record_error(UZU_ERROR_OUT_OF_MEMORY, "attention_single_pass: workspace allocation failed");
                return;
            }
            ctx->attn_ws_bytes = need;
        }
        if ((size_t)a->suffix_length * head_groups > 65536) {
            cmd->record_error(UZU_ERROR_UNSUPPORTED, "attention_single_pass: too many (token, head group) pairs for the split path");
            return;
        }
        p.part_o = ctx->attn_ws;
        p.part_sum = ctx->attn_ws + units * a->head_dim;
        p.part_max = p.part_sum + units;
        p.counters = ctx->attn_counters;
    }
    uzu::dispatch_attn(cmd, p, g);
}

void uzu_attention_two_pass1_encode(uzu_command_buffer* cmd, const uzu_attention_args* a) {
    if (!uzu::encodable(cmd, "attention_two_pass1")) return;
    if (!uzu::attn_common_checks(cmd, a, "attention_two_pass1")) return;
    if (!a->sums || !a->maxs) {
        cmd->record_error(UZU_ERROR_INVALID_ARGUMENT, "attention_two_pass1: missing sums / maxs");
        return;
    }
    uzu_context* ctx = cmd->ctx;
    const int g = uzu::pick_group(a->gqa_factor);
    const uint32_t head_groups = a->num_heads / g;
    const uint32_t ctas = head_groups * a->suffix_length;
    const uint32_t want = (2u * (uint32_t)ctx->sm_count + ctas - 1) / ctas;
    const uint32_t cap = (a->sequence_length + 63) / 64;
    uzu::AttnParams p{};
    p.a = *a;
    p.part_o = reinterpret_cast<float*>(a->out);
    p.part_sum = reinterpret_cast<float*>(a->sums);
    p.part_max = reinterpret_cast<float*>(a->maxs);
    p.fuse_merge = 0;
    p.nsplits = std::max(1u, std::min(std::min(want, cap), 32u));
    p.nb = 32;  // TOTAL_BLOCKS_COUNT (attention_two_pass.rs:10)
    p.keys_per_split = p.nsplits > 1 ? (((a->sequence_length + p.nsplits - 1) / p.nsplits + 15u) & ~15u) : a->sequence_length;
    uzu::dispatch_attn(cmd, p, g);
}

void uzu_attention_two_pass2_encode(uzu_command_buffer* cmd, const uzu_attention_two_pass2_args* a) {
    if (!uzu::encodable(cmd, "attention_two_pass2")) return;
    if (!a->partials || !a->sums || !a->maxs || !a->out || a->head_dim == 0) {
        cmd->record_error(UZU_ERROR_INVALID_ARGUMENT, "attention_two_pass2: null operand");
        return;
    }
    const uint32_t units = a->num_heads * a->suffix_length;
    if (units == 0) return;
    uzu::launch(cmd, "attn_two_pass2_kernel", uzu::attn_two_pass2_kernel, dim3((units + 3) / 4), dim3(128), 0, *a);
}

void uzu_kv_cache_update_encode(uzu_command_buffer* cmd, const uzu_kv_cache_update_args* a) {
    if (!uzu::encodable(cmd, "kv_cache_update")) return;
    if (a->copy_count == 0 || a->element_dim == 0) return;  // flat decode path: nothing to move (state.rs:185-189)
    if (!a->in_place_keys || !a->in_place_values || !a->copies) {
        cmd->record_error(UZU_ERROR_INVALID_ARGUMENT, "kv_cache_update: null operand");
        return;
    }
    for (uint32_t done = 0; done < a->copy_count; done += 512) {
        uzu::KvCopies c;
        const uint32_t n = std::min(512u, a->copy_count - done);
        memcpy(c.c, a->copies + done, n * sizeof(uzu_kv_copy));
        uzu::kv_cache_update_kernel<<<(a->element_dim + 255) / 256, 256, 0, cmd->ctx->stream>>>(
            reinterpret_cast<__nv_bfloat16*>(a->in_place_keys), reinterpret_cast<__nv_bfloat16*>(a->in_place_values), c, n, a->element_dim);
        uzu::after_launch(cmd, "kv_cache_update_kernel");
    }
}

void uzu_sigmoid_gate_encode(uzu_command_buffer* cmd, uint64_t gate, uint64_t output, uint32_t total_elements) {
    if (!uzu::encodable(cmd, "sigmoid_gate")) return;
    if (total_elements == 0) return;
    if (!gate || !output) {
        cmd->record_error(UZU_ERROR_INVALID_ARGUMENT, "sigmoid_gate: null operand");
        return;
    }
    uzu::launch(cmd, "sigmoid_gate_kernel", uzu::sigmoid_gate_kernel, dim3((total_elements + 255) / 256), dim3(256), 0,
                reinterpret_cast<const __nv_bfloat16*>(gate), reinterpret_cast<__nv_bfloat16*>(output), total_elements);
}

}  // extern "C"
