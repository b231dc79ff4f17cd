// Prefill attention on tensor cores (suffix >= 16 query tokens): flash-attention-style tiling over the token-major KV cache.
//
//   OPT-IN (UZU_PREFILL_ATTN=1). Written after round 1's GPU budget was spent; its only hardware run so far is tools/prefill_attn_probe.py
//   (3 shapes within one bf16 step of a float64 softmax, profiles/r1_prefill_attention_probe.txt). Until tests/test_prefill_attention_gpu.py
//   (oracle ulp comparison, enabled with UZU_TEST_PREFILL_ATTN=1) and the engine tests have run through it, the default prefill path stays
//   the split-KV decode kernel applied per query token (attention.cu).
//
// Semantic spec: backends/cpu/kernel/attention/attention_single_pass.rs:49-126 (online softmax per query over the visible keys, q scaled by
// `scale`, GQA kv_head = h / gqa, output [suffix, heads, D] cast after the division by the sum), mask.rs:3-62 restricted to the plain causal /
// non-causal case (no ring, trie, sliding window, sinks: those keep the decode kernel).
//
// One CTA = 16 query tokens x one KV head; its G = heads-per-KV-head warps each own one query head and share the K / V tiles:
//   * S = Q K^T with mma.sync.m16n8k16 (bf16 x bf16 products are exact, f32 accumulate), Q fragments live in registers, the K tile
//     (64 keys x D, rows padded to D + 8 so the 32-bit fragment loads are bank-conflict free) arrives by cp.async, double buffered;
//   * online softmax in f32 registers (expf, like the reference), the row statistics reduced over the 4 lanes of a quad;
//   * O += P V with P split into two bf16 planes (hi = bf16(p), lo = bf16(p - hi)): a single rounded plane would put 2^-9 on every
//     probability, the same budget argument as the prefill GEMM's weight planes; V fragments come from ldmatrix.trans.
// Causal blocks past the CTA's last query are never loaded. Work per CTA grows with the query index; CTAs are launched in reverse
// (longest first) so the tail is short.
#include <algorithm>
#include <cmath>
#include <cstdlib>

#include "common.cuh"

namespace uzu {

constexpr int PA_BM = 16;   // query rows per CTA (one m16 tile per warp)
constexpr int PA_BN = 64;   // keys per K / V tile

__device__ __forceinline__ void pa_mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void pa_ldmatrix_x2_trans(uint32_t& r0, uint32_t& r1, const void* smem_ptr) {
    const uint32_t addr = (uint32_t)__cvta_generic_to_shared(smem_ptr);
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
__device__ __forceinline__ void pa_cp16(void* dst_smem, const void* src, bool valid) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst_smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(valid ? 16u : 0u) : "memory");
}
__device__ __forceinline__ uint32_t pa_pack(float lo_elem, float hi_elem) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo_elem, hi_elem);
    return *reinterpret_cast<uint32_t*>(&v);
}

template <int D, int G>
__global__ void __launch_bounds__(G * 32) attn_prefill_kernel(const uzu_attention_args a) {
    constexpr int LD = D + 8;                       // padded tile row (elements): 16-byte aligned rows, conflict-free fragment loads
    constexpr int TILE = PA_BN * LD;
    constexpr int CH = D / 8;                       // 16-byte chunks per row
    extern __shared__ __align__(16) uint8_t pa_smem[];
    __nv_bfloat16* sk = reinterpret_cast<__nv_bfloat16*>(pa_smem);   // [2][TILE]
    __nv_bfloat16* sv = sk + 2 * TILE;                                // [2][TILE]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, quad = lane >> 2, tq = lane & 3;
    const uint32_t m = a.suffix_length, seq = a.sequence_length, prefix = seq - m;
    const uint32_t qblk = gridDim.x - 1u - blockIdx.x;              // longest (latest queries) first
    const uint32_t q0 = qblk * PA_BM, kvh = blockIdx.y, h = kvh * G + (uint32_t)warp;
    const uint32_t q_hi = min(q0 + (uint32_t)PA_BM, m);
    const uint32_t kend = a.is_causal ? prefix + q_hi : seq;        // keys [0, kend) can be visible to some row of this CTA
    const uint32_t nblocks = (kend + PA_BN - 1) / PA_BN;
    const __nv_bfloat16* kbase = reinterpret_cast<const __nv_bfloat16*>(a.keys) + (size_t)kvh * a.k_head_stride;
    const __nv_bfloat16* vbase = reinterpret_cast<const __nv_bfloat16*>(a.values) + (size_t)kvh * a.v_head_stride;

    // ---- Q fragments (A operand, row-major 16 x 16 per k step): a0 (row g, k 2t..), a1 (row g+8), a2 (row g, k 2t+8..), a3 (row g+8) ----
    uint32_t qf[D / 16][4];
    {
        const __nv_bfloat16* qh = reinterpret_cast<const __nv_bfloat16*>(a.queries) + (size_t)h * m * D;
        const uint32_t r0 = q0 + quad, r1 = r0 + 8;
#pragma unroll
        for (int ks = 0; ks < D / 16; ++ks) {
            const int c0 = ks * 16 + 2 * tq;
            qf[ks][0] = r0 < m ? *reinterpret_cast<const uint32_t*>(qh + (size_t)r0 * D + c0) : 0u;
            qf[ks][1] = r1 < m ? *reinterpret_cast<const uint32_t*>(qh + (size_t)r1 * D + c0) : 0u;
            qf[ks][2] = r0 < m ? *reinterpret_cast<const uint32_t*>(qh + (size_t)r0 * D + c0 + 8) : 0u;
            qf[ks][3] = r1 < m ? *reinterpret_cast<const uint32_t*>(qh + (size_t)r1 * D + c0 + 8) : 0u;
        }
    }

    auto load_tiles = [&](uint32_t kb, int stage) {
        __nv_bfloat16* dk = sk + stage * TILE;
        __nv_bfloat16* dv = sv + stage * TILE;
        for (int c = threadIdx.x; c < PA_BN * CH; c += G * 32) {
            const int key = c / CH, ch = c % CH;
            const uint32_t kg = kb * PA_BN + (uint32_t)key;
            const bool ok = kg < kend;                               // rows past the visible range may be unmapped cache pages: never read
            const size_t row = ok ? kg : 0u;
            pa_cp16(dk + key * LD + ch * 8, kbase + row * a.k_seq_stride + ch * 8, ok);
            pa_cp16(dv + key * LD + ch * 8, vbase + row * a.v_seq_stride + ch * 8, ok);
        }
    };

    float o[D / 8][4];
#pragma unroll
    for (int nd = 0; nd < D / 8; ++nd) o[nd][0] = o[nd][1] = o[nd][2] = o[nd][3] = 0.0f;
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.0f, l1 = 0.0f;     // running max / sum of rows (quad) and (quad + 8)
    const uint32_t qr0 = q0 + quad, qr1 = qr0 + 8;
    // exclusive key bound per row: causal -> keys up to and including the query's own position
    const uint32_t lim0 = qr0 < m ? (a.is_causal ? prefix + qr0 + 1 : seq) : 0u;
    const uint32_t lim1 = qr1 < m ? (a.is_causal ? prefix + qr1 + 1 : seq) : 0u;

    if (nblocks > 0) load_tiles(0, 0);
    asm volatile("cp.async.commit_group;" ::: "memory");
    for (uint32_t kb = 0; kb < nblocks; ++kb) {
        const int stage = kb & 1;
        if (kb + 1 < nblocks) load_tiles(kb + 1, stage ^ 1);
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 1;" ::: "memory");        // tile kb has landed (tile kb + 1 may be in flight)
        __syncthreads();
        const __nv_bfloat16* kt = sk + stage * TILE;
        const __nv_bfloat16* vt = sv + stage * TILE;

        // ---- S = Q K^T: B fragment b0 = K[key = nt*8 + g][d = ks*16 + 2t .. +1], b1 = same key, d + 8 ----
        float s[PA_BN / 8][4];
#pragma unroll
        for (int nt = 0; nt < PA_BN / 8; ++nt) s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.0f;
#pragma unroll
        for (int ks = 0; ks < D / 16; ++ks) {
#pragma unroll
            for (int nt = 0; nt < PA_BN / 8; ++nt) {
                const __nv_bfloat16* kp = kt + (nt * 8 + quad) * LD + ks * 16 + 2 * tq;
                pa_mma(s[nt], qf[ks], *reinterpret_cast<const uint32_t*>(kp), *reinterpret_cast<const uint32_t*>(kp + 8));
            }
        }
        // ---- scale, mask, online softmax (c0,c1: row g, cols 2t, 2t+1; c2,c3: row g+8) ----
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int nt = 0; nt < PA_BN / 8; ++nt) {
            const uint32_t kg = kb * PA_BN + nt * 8 + 2 * tq;
            s[nt][0] = kg < lim0 ? s[nt][0] * a.scale : -INFINITY;
            s[nt][1] = kg + 1 < lim0 ? s[nt][1] * a.scale : -INFINITY;
            s[nt][2] = kg < lim1 ? s[nt][2] * a.scale : -INFINITY;
            s[nt][3] = kg + 1 < lim1 ? s[nt][3] * a.scale : -INFINITY;
            mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
            mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
        }
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
        const float n0 = fmaxf(m0, mx0), n1 = fmaxf(m1, mx1);
        const float c0 = n0 == -INFINITY ? 1.0f : expf(m0 - n0), c1 = n1 == -INFINITY ? 1.0f : expf(m1 - n1);   // expf(-inf) = 0
        float rs0 = 0.0f, rs1 = 0.0f;
#pragma unroll
        for (int nt = 0; nt < PA_BN / 8; ++nt) {
            s[nt][0] = n0 == -INFINITY ? 0.0f : expf(s[nt][0] - n0);
            s[nt][1] = n0 == -INFINITY ? 0.0f : expf(s[nt][1] - n0);
            s[nt][2] = n1 == -INFINITY ? 0.0f : expf(s[nt][2] - n1);
            s[nt][3] = n1 == -INFINITY ? 0.0f : expf(s[nt][3] - n1);
            rs0 += s[nt][0] + s[nt][1];
            rs1 += s[nt][2] + s[nt][3];
        }
        rs0 += __shfl_xor_sync(0xffffffffu, rs0, 1); rs0 += __shfl_xor_sync(0xffffffffu, rs0, 2);
        rs1 += __shfl_xor_sync(0xffffffffu, rs1, 1); rs1 += __shfl_xor_sync(0xffffffffu, rs1, 2);
        l0 = l0 * c0 + rs0;
        l1 = l1 * c1 + rs1;
        m0 = n0;
        m1 = n1;
#pragma unroll
        for (int nd = 0; nd < D / 8; ++nd) {
            o[nd][0] *= c0; o[nd][1] *= c0;
            o[nd][2] *= c1; o[nd][3] *= c1;
        }
        // ---- O += P V: the S accumulator layout IS the A fragment layout of the next MMA (k = key) ----
#pragma unroll
        for (int kt2 = 0; kt2 < PA_BN / 16; ++kt2) {
            uint32_t ph[4], pl[4];
            const float* e0 = s[2 * kt2];
            const float* e1 = s[2 * kt2 + 1];
            const float pv[8] = {e0[0], e0[1], e0[2], e0[3], e1[0], e1[1], e1[2], e1[3]};
#pragma unroll
            for (int r = 0; r < 4; ++r) {   // a0 = (row g, k 2t..): e0[0..1]; a1 = (row g+8): e0[2..3]; a2 = (row g, k+8): e1[0..1]; a3: e1[2..3]
                const float x = pv[2 * r], y = pv[2 * r + 1];
                ph[r] = pa_pack(x, y);
                pl[r] = pa_pack(x - __uint_as_float(ph[r] << 16), y - __uint_as_float(ph[r] & 0xffff0000u));
            }
#pragma unroll
            for (int nd = 0; nd < D / 8; ++nd) {
                uint32_t b0, b1;   // B[k = key][n = d]: V rows are keys -> transposed 8x8 loads; lanes 0-15 address rows kt2*16 + lane
                pa_ldmatrix_x2_trans(b0, b1, vt + (kt2 * 16 + (lane & 15)) * LD + nd * 8);
                pa_mma(o[nd], ph, b0, b1);
                pa_mma(o[nd], pl, b0, b1);
            }
        }
        __syncthreads();                                            // everyone is done with this stage before it is refilled
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");

    // ---- out[token, head, d] = O / l, cast after the division ----
    __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(a.out);
    const float i0 = l0 > 0.0f ? 1.0f / l0 : 0.0f, i1 = l1 > 0.0f ? 1.0f / l1 : 0.0f;
#pragma unroll
    for (int nd = 0; nd < D / 8; ++nd) {
        const int d = nd * 8 + 2 * tq;
        if (qr0 < m) *reinterpret_cast<uint32_t*>(out + ((size_t)qr0 * a.num_heads + h) * D + d) = pa_pack(o[nd][0] * i0, o[nd][1] * i0);
        if (qr1 < m) *reinterpret_cast<uint32_t*>(out + ((size_t)qr1 * a.num_heads + h) * D + d) = pa_pack(o[nd][2] * i1, o[nd][3] * i1);
    }
}

template <int D, int G>
static void launch_attn_prefill(uzu_command_buffer* cmd, const uzu_attention_args& a) {
    constexpr size_t smem = (size_t)4 * PA_BN * (D + 8) * sizeof(__nv_bfloat16);
    static std::atomic<uint64_t> smem_opt_in{0};
    opt_in_dynamic_smem(cmd, attn_prefill_kernel<D, G>, (int)((int)smem), smem_opt_in);
    const dim3 grid((a.suffix_length + PA_BM - 1) / PA_BM, a.num_heads / G);
    attn_prefill_kernel<D, G><<<grid, G * 32, smem, cmd->ctx->stream>>>(a);
    after_launch(cmd, "attn_prefill_kernel");
}

// true = handled. Plain causal / non-causal attention over a dense prefix only; everything else keeps the decode kernel.
static int g_prefill_attn = -1;   // -1: follow UZU_PREFILL_ATTN, 0 / 1: forced by uzu_debug_set_prefill_attention

bool encode_attention_prefill(uzu_command_buffer* cmd, const uzu_attention_args& a) {
    static const bool env_enabled = [] { const char* e = getenv("UZU_PREFILL_ATTN"); return !e || atoi(e) != 0; }();   // default on (validated on B200, round 2); =0 keeps the split-KV kernel
    if (!(g_prefill_attn < 0 ? env_enabled : g_prefill_attn != 0)) return false;
    if (a.suffix_length < 16 || a.dynamic_position || a.has_sinks || a.is_kv_cache_ring || a.is_trie || a.is_sliding_window) return false;
    if (a.head_dim != 64 && a.head_dim != 128) return false;
    const uint32_t g = a.gqa_factor;
    if ((g != 1 && g != 2 && g != 4 && g != 8) || a.num_heads % g) return false;
    if ((a.k_head_stride % 8) || (a.v_head_stride % 8) || (a.k_seq_stride % 8) || (a.v_seq_stride % 8)) return false;   // 16-byte cp.async rows
    if ((a.queries | a.keys | a.values | a.out) & 15u) return false;
    if (a.sequence_length < a.suffix_length) return false;
#define UZU_PA(DD, GG) launch_attn_prefill<DD, GG>(cmd, a); return true
    if (a.head_dim == 64) {
        switch (g) { case 1: UZU_PA(64, 1); case 2: UZU_PA(64, 2); case 4: UZU_PA(64, 4); default: UZU_PA(64, 8); }
    }
    switch (g) { case 1: UZU_PA(128, 1); case 2: UZU_PA(128, 2); case 4: UZU_PA(128, 4); default: UZU_PA(128, 8); }
#undef UZU_PA
}

}  // namespace uzu

extern "C" void uzu_debug_set_prefill_attention(int mode) { uzu::g_prefill_attn = mode; }
