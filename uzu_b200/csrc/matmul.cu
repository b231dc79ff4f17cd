// MatmulKernel for sm_100a: D[m,n] = epilogue(A[m,k] * dequant(B[n,k])^T).
//
// Semantic spec: backends/cpu/kernel/matmul/kernel.rs:164-295 (reference CPU backend).
//
// Decode path (m <= 16): `qmv_kernel`, a fused dequant + GEMV. Why it looks the way it does:
//   * the op is pure weight streaming (4.05 GB/token for Llama-3-8B int4), so the HBM roofline is
//     the target: at 6.5 TB/s each SM must retire ~52 int4 weights per clock, while an SM issues at
//     most 128 lane-instructions per clock. A SIMT dequant (extract, int->float, FMA with f32
//     accumulation as the reference requires) needs >= 3 instructions per weight and is issue-bound
//     below the HBM roofline. So the per-weight work is reduced to <1 instruction: nibbles are turned
//     into exact bf16 integers (128 + code) with one shift + one LOP3 per *pair* and the
//     multiply-accumulate runs on the tensor pipe via mma.sync.m16n8k16 (bf16 x bf16 -> f32; products
//     are exact, accumulation is f32 like the reference). The affine part of the dequantisation is
//     hoisted out per quantisation group:
//         sum_k x_k (s*q_k + c) = s * (sum_k x_k (128+q_k) - 128 * Sx) + c * Sx,   Sx = sum_k x_k
//     (the reference's own Metal GEMV hoists the same way, metal/kernel/matmul/common/qdot.h:88-89).
//   * every lane streams 2 x 16 B of packed weights per step with ld.global.nc.L1::no_allocate
//     (128-bit, fully used sectors); a warp covers 16 output rows x 128 k per step and the k range of
//     a row tile is interleaved across the 8 warps of a CTA, so a CTA reads 512 contiguous bytes of
//     each of its 16 rows per step.
//   * int8 weights reuse the same inner loop: a byte is two nibbles (lo, hi) and the activation
//     for the hi nibble is pre-multiplied by 16 (exact in bf16).
//   * the spare columns of the 16x8 MMA tile separate quantisation groups (group 64 -> 2 columns per
//     activation row), so one MMA serves both groups a warp touches in a step.
//   * split-K (for few-row / long-k shapes) reduces through a stream-ordered f32 workspace in a fixed
//     order (deterministic), the last CTA of a tile applies the epilogue.
//
// Everything the fast path does not cover (full-precision B, gather, odd shapes, f32 activations)
// runs on `generic_kernel`, a warp-per-output restatement of the reference loop.
#include <algorithm>

#include "common.cuh"

namespace uzu {

struct QmvParams {
    const uint8_t* w;          // packed codes, row stride = row_bytes
    const __nv_bfloat16* scales;
    const uint8_t* zero_points;
    const __nv_bfloat16* biases;
    const __nv_bfloat16* x;    // [m, k]
    void* d;                   // [m, n]
    const __nv_bfloat16* bias; // epilogue bias [n] or null
    float* ws;
    unsigned int* counters;
    uint32_t m, n, k;
    uint32_t np;               // nibbles per row
    uint32_t row_bytes;
    uint32_t groups_per_row, zp_stride;
    uint32_t group_size;       // in k elements
    uint32_t chunks_total, chunks_per_slice, kslices;
    uint32_t warps_per_tile;   // decode kernel: warps of a CTA sharing one 16-row tile (1, 2 or 4)
    // fused activation prologue (decode kernel only): the activation row is produced inside the GEMV instead of by a
    // separate Normalization / GatedActMul / SigmoidGate launch. 0 = plain (x), 1 = RMS norm, 2 = gated act, 3 = sigmoid gate
    uint32_t prologue;
    const __nv_bfloat16* pro_a;      // norm: input row; gated: fused up row [2F]; sigmoid: attention output row
    const __nv_bfloat16* pro_b;      // norm: shortcut in (or null); sigmoid: gate row
    __nv_bfloat16* pro_shortcut_out; // norm: updated residual, written by CTA 0 only (null = do not write)
    const float* pro_scales;         // norm scales f32 [k]
    float pro_eps, pro_scale_offset;
    uint32_t pro_residual_add, pro_full_layer, pro_act;
    // fused epilogue (decode kernel): rows [0, F) are `up`, rows [F, 2F) are `gate`; a CTA walks tile i then tile i + F/16 and
    // writes hidden[j] = bf16(bf16(up_j) * bf16(act(bf16(gate_j)))) (GatedActMul, gated_act_mul/mod.rs:5-12) instead of the 2F row
    uint32_t epi_gated, epi_act, pair_tiles;
    uint32_t stages;                 // cp.async ring depth of the decode kernel (host-side dispatch only)
    uint32_t method;           // uzu_quantization_method
    uint32_t bits;
    uint32_t xor_mask;         // signed_codes
    uint32_t d_is_f32;
    uint32_t accumulate, has_soft_cap;
    float ab_scale, soft_cap;
    // decode-stream copy of the matrix (decode_mega.cu mega_repack: unit-major 4608-byte units = 16 rows x 512 nibbles + coefficients), or null.
    // With it the decode kernel (METHOD == QMV_TMA) feeds its rings with ONE TMA bulk copy per stage instead of 12 cp.async per lane.
    const uint8_t* stream;
    uint32_t stream_C, pad_stream;
};

__device__ __forceinline__ void mma_16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                          uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// (code_i, code_{i+4}) of a packed word -> bf16x2 holding (128+code_i, 128+code_{i+4}); exact.
__device__ __forceinline__ uint32_t nib_pair(uint32_t w, int shift) { return ((w >> shift) & 0x000f000fu) | 0x43004300u; }

// Same, tuned for the decode kernel where the integer ALU pipe is the limiter: the shift is a multiply-high
// (IMAD.HI runs on the FMA pipe, which is otherwise idle) and (x & mask) | magic is forced into ONE LOP3 by keeping
// the magic constant in a register (ptxas emits two LOP3 when both constants are immediates).
#ifndef UZU_QA_SHIFT_MODE
#define UZU_QA_SHIFT_MODE 1     // 0: every shift on the FMA pipe (IMAD.HI); 1: plain shifts (ALU pipe); 2: shift by 8 on the ALU pipe, 4 / 12 on FMA.
                                // Measured (Llama-3-8B linears, B200): mode 1 is 13-19% faster than mode 0 once the kernel is small
                                // enough not to stall on instruction fetch (IMAD.HI showed up as dispatch stalls in ncu).
#endif
__device__ __forceinline__ uint32_t nib_pair_fast(uint32_t w, int shift, uint32_t magic_reg) {
    uint32_t sh;
    if (shift == 0) sh = w;
    else if (UZU_QA_SHIFT_MODE == 1 || (UZU_QA_SHIFT_MODE == 2 && shift == 8)) sh = w >> shift;
    else sh = __umulhi(w, 1u << (32 - shift));
    uint32_t d;
    asm("lop3.b32 %0, %1, 0x000f000f, %2, 0xEA;" : "=r"(d) : "r"(sh), "r"(magic_reg));
    return d;
}

constexpr int QMV_WARPS = 8;

// NPG: nibbles per quantisation group (32, 64, 128, 256); MT: MMA column tiles (activation rows = MT * 8 / CPM)
template <int NPG, int MT>
__global__ void __launch_bounds__(QMV_WARPS * 32) qmv_kernel(const QmvParams p) {
    constexpr int CPM = NPG >= 128 ? 1 : 128 / NPG;  // MMA columns per activation row
    constexpr int MPM = 8 / CPM;                     // activation rows per MMA
    constexpr int MROWS = MT * MPM;
    extern __shared__ __align__(16) uint8_t smem_raw[];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const uint32_t tile = blockIdx.x / p.kslices, slice = blockIdx.x % p.kslices;
    const uint32_t cb = slice * p.chunks_per_slice;
    const uint32_t ce = min(p.chunks_total, cb + p.chunks_per_slice);
    const uint32_t nc = ce - cb;
    // groups touched by this slice (in units of NPG nibbles)
    const uint32_t grp_begin = (cb * 128u) / NPG;
    const uint32_t grp_end = min(p.groups_per_row, (ce * 128u + NPG - 1) / NPG);
    const uint32_t ngl = grp_end - grp_begin;

    const uint32_t coef_stride = ((ngl + 15u) & ~15u) + 2u;                            // float2 units; rows land on distinct banks
    const uint32_t row_items = (nc * 16u + 31u) & ~31u;                                // staged words per activation row (warp aligned)
    uint4* xs = reinterpret_cast<uint4*>(smem_raw);                                   // [m][row_items] = [m][nc][4 t][4 w]
    float2* coef = reinterpret_cast<float2*>(smem_raw + (size_t)p.m * row_items * 16); // [16 rows][coef_stride] (scale, Sx coefficient)
    float* sx = reinterpret_cast<float*>(coef + 16 * coef_stride);                    // [m][ngl]
    float* red = sx + (size_t)p.m * ngl;                                              // [QMV_WARPS][MT*4][32]

    const uint32_t row_a = min(tile * 16u + (uint32_t)g, p.n - 1), row_b = min(tile * 16u + (uint32_t)g + 8u, p.n - 1);
    const uint8_t* wa_base = p.w + (size_t)row_a * p.row_bytes;
    const uint8_t* wb_base = p.w + (size_t)row_b * p.row_bytes;

    // first weight chunk is requested before the prologue so its HBM latency overlaps the staging work
    uint4 wa = make_uint4(0, 0, 0, 0), wb = make_uint4(0, 0, 0, 0);
    uint32_t c = cb + warp;
    if (c < ce) {
        const uint32_t pos = c * 128u + (uint32_t)t * 32u;
        if (pos < p.np) {
            wa = ldg_stream_u4(wa_base + pos / 2);
            wb = ldg_stream_u4(wb_base + pos / 2);
        }
    }

    // ---- prologue 1: permuted bf16 activation fragments + per-group activation sums -----------------------
    {
        constexpr uint32_t IPG = NPG / 8;   // staged 8-nibble words per quantisation group (4..32), lanes of one warp
        const uint32_t items_padded = p.m * row_items;
        for (uint32_t it = tid; it < items_padded; it += blockDim.x) {
            uint4 out = make_uint4(0, 0, 0, 0);
            float part = 0.0f;
            const uint32_t r = it / row_items, li = it % row_items;
            uint32_t gl = 0xffffffffu;
            const bool in_row = li < nc * 16u;
            if (in_row) {
                const uint32_t w_ = li & 3, t_ = (li >> 2) & 3, c_ = li >> 4;
                const uint32_t pos = (cb + c_) * 128u + t_ * 32u + w_ * 8u;  // nibble position of the word
                gl = pos / NPG - grp_begin;
                if (pos < p.np) {
                    if (p.bits == 4) {
                        const uint4 v = *reinterpret_cast<const uint4*>(p.x + (size_t)r * p.k + pos);
                        // v = (x0,x1),(x2,x3),(x4,x5),(x6,x7) -> (x0,x4),(x1,x5),(x2,x6),(x3,x7)
                        out.x = __byte_perm(v.x, v.z, 0x5410);
                        out.y = __byte_perm(v.x, v.z, 0x7632);
                        out.z = __byte_perm(v.y, v.w, 0x5410);
                        out.w = __byte_perm(v.y, v.w, 0x7632);
                        const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
                        for (int e = 0; e < 4; ++e) { part += __low2float(h2[e]); part += __high2float(h2[e]); }
                    } else {
                        // 8-bit: nibble 2j = lo(code_j) -> x_j, nibble 2j+1 = hi(code_j) -> 16*x_j
                        const uint2 v = *reinterpret_cast<const uint2*>(p.x + (size_t)r * p.k + pos / 2);
                        __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&v.x);  // x0,x1
                        __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&v.y);  // x2,x3
                        const float x0 = __low2float(a), x1 = __high2float(a), x2 = __low2float(b), x3 = __high2float(b);
                        // positions p0..p7 = x0,16x0,x1,16x1,x2,16x2,x3,16x3 ; want (p0,p4),(p1,p5),(p2,p6),(p3,p7)
                        __nv_bfloat162 o0 = __floats2bfloat162_rn(x0, x2);
                        __nv_bfloat162 o1 = __floats2bfloat162_rn(16.0f * x0, 16.0f * x2);
                        __nv_bfloat162 o2 = __floats2bfloat162_rn(x1, x3);
                        __nv_bfloat162 o3 = __floats2bfloat162_rn(16.0f * x1, 16.0f * x3);
                        out.x = *reinterpret_cast<uint32_t*>(&o0);
                        out.y = *reinterpret_cast<uint32_t*>(&o1);
                        out.z = *reinterpret_cast<uint32_t*>(&o2);
                        out.w = *reinterpret_cast<uint32_t*>(&o3);
                        part = ((x0 + x1) + x2) + x3;
                    }
                }
                xs[it] = out;
            }
            // fixed-shape butterfly over the IPG consecutive lanes that hold one group (deterministic)
#pragma unroll
            for (uint32_t o = 1; o < IPG; o <<= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
            if (in_row && (li & (IPG - 1)) == 0 && gl < ngl) sx[(size_t)r * ngl + gl] = part;
        }
    }
    // ---- prologue 2: per (row, group) dequantisation coefficients: value = s * dot(128 + code) + z * Sx ----
    {
        const float mult128 = p.bits == 4 ? 128.0f : 128.0f * 17.0f;
        const float sym_mid = p.bits == 4 ? 8.0f : 128.0f;
        for (uint32_t it = tid; it < 16u * ngl; it += blockDim.x) {
            const uint32_t gl = it % ngl, rr = it / ngl;
            const uint32_t row = min(tile * 16u + rr, p.n - 1), gi = grp_begin + gl;
            const float sc = __bfloat162float(p.scales[(size_t)row * p.groups_per_row + gi]);
            float z;
            if (p.method == UZU_QMETHOD_SCALE_ZERO_POINT) {
                float zp;
                if (p.bits == 4) {
                    const uint8_t byte = p.zero_points[(size_t)row * p.zp_stride + (gi >> 1)];
                    zp = (float)((gi & 1) ? (byte >> 4) : (byte & 15));
                } else {
                    zp = (float)p.zero_points[(size_t)row * p.zp_stride + gi];
                }
                z = -sc * (zp + mult128);
            } else if (p.method == UZU_QMETHOD_SCALE_BIAS) {
                z = __bfloat162float(p.biases[(size_t)row * p.groups_per_row + gi]) - sc * mult128;
            } else {
                z = -sc * (sym_mid + mult128);
            }
            coef[(size_t)rr * coef_stride + gl] = make_float2(sc, z);
        }
    }
    __syncthreads();

    // ---- main loop ---------------------------------------------------------------------------------
    float acc[MT][4];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[mt][i] = 0.0f;

    // B-fragment ownership: lane (n = g, t) feeds MMA column n with the k-slots of quad lane t.
    const int col_sub = g % CPM, col_mrow = g / CPM;
    const int lane_sub = CPM == 1 ? 0 : (CPM == 2 ? (t >> 1) : t);
    const float2* coef_a = coef + (size_t)g * coef_stride;
    const float2* coef_b = coef + (size_t)(g + 8) * coef_stride;

    for (; c < ce; c += QMV_WARPS) {
        // software pipeline: request the next chunk of this warp before working on the current one
        uint4 wa_n = make_uint4(0, 0, 0, 0), wb_n = make_uint4(0, 0, 0, 0);
        {
            const uint32_t cn = c + QMV_WARPS;
            const uint32_t posn = cn * 128u + (uint32_t)t * 32u;
            if (cn < ce && posn < p.np) {
                wa_n = ldg_stream_u4(wa_base + posn / 2);
                wb_n = ldg_stream_u4(wb_base + posn / 2);
            }
        }
        float d[MT][4];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
            for (int i = 0; i < 4; ++i) d[mt][i] = 0.0f;

        const uint32_t wav[4] = {wa.x ^ p.xor_mask, wa.y ^ p.xor_mask, wa.z ^ p.xor_mask, wa.w ^ p.xor_mask};
        const uint32_t wbv[4] = {wb.x ^ p.xor_mask, wb.y ^ p.xor_mask, wb.z ^ p.xor_mask, wb.w ^ p.xor_mask};
#pragma unroll
        for (int w_ = 0; w_ < 4; ++w_) {
            const uint32_t a0 = nib_pair(wav[w_], 0), a1 = nib_pair(wav[w_], 4), a2 = nib_pair(wav[w_], 8), a3 = nib_pair(wav[w_], 12);
            const uint32_t b0 = nib_pair(wbv[w_], 0), b1 = nib_pair(wbv[w_], 4), b2 = nib_pair(wbv[w_], 8), b3 = nib_pair(wbv[w_], 12);
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
                uint4 xb = make_uint4(0, 0, 0, 0);
                const uint32_t mr = mt * MPM + col_mrow;
                if (col_sub == lane_sub && mr < p.m) xb = xs[(size_t)mr * row_items + ((c - cb) * 4 + t) * 4 + w_];
                // mma A regs: (row g, k 2t..), (row g+8, k 2t..), (row g, k 2t+8..), (row g+8, k 2t+8..)
                mma_16816(d[mt], a0, b0, a1, b1, xb.x, xb.y);
                mma_16816(d[mt], a2, b2, a3, b3, xb.z, xb.w);
            }
        }

        // ---- per-group affine part: D columns of this thread are 2t and 2t+1 ---------------------------------
        const int c0 = 2 * t, c1 = 2 * t + 1;
        uint32_t gl0, gl1;
        if (CPM == 1) gl0 = gl1 = (c * 128u) / NPG - grp_begin;
        else { gl0 = c * CPM + (c0 % CPM) - grp_begin; gl1 = gl0 + 1; }
        const bool v0 = gl0 < ngl, v1 = gl1 < ngl;
        float2 ca0 = make_float2(0.f, 0.f), ca1 = ca0, cb0 = ca0, cb1 = ca0;
        if (CPM == 1) {
            if (v0) { ca0 = ca1 = coef_a[gl0]; cb0 = cb1 = coef_b[gl0]; }
        } else if (v1) {   // both groups valid: one 128-bit shared load per row (gl0 is even)
            const float4 fa = *reinterpret_cast<const float4*>(coef_a + gl0);
            const float4 fb = *reinterpret_cast<const float4*>(coef_b + gl0);
            ca0 = make_float2(fa.x, fa.y); ca1 = make_float2(fa.z, fa.w);
            cb0 = make_float2(fb.x, fb.y); cb1 = make_float2(fb.z, fb.w);
        } else if (v0) {
            ca0 = coef_a[gl0]; cb0 = coef_b[gl0];
        }
        // NPG > 128: a group spans several chunks; its Sx term is added once, on the group's first chunk
        const bool first_chunk_of_group = (NPG <= 128) || ((c * 128u) % NPG == 0);
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
            const uint32_t mr0 = mt * MPM + c0 / CPM, mr1 = mt * MPM + c1 / CPM;
            float sx0 = 0.0f, sx1 = 0.0f;
            if (first_chunk_of_group) {
                if (v0 && mr0 < p.m) sx0 = sx[(size_t)mr0 * ngl + gl0];
                if (v1 && mr1 < p.m) sx1 = sx[(size_t)mr1 * ngl + gl1];
            }
            acc[mt][0] += ca0.x * d[mt][0] + ca0.y * sx0;
            acc[mt][1] += ca1.x * d[mt][1] + ca1.y * sx1;
            acc[mt][2] += cb0.x * d[mt][2] + cb0.y * sx0;
            acc[mt][3] += cb1.x * d[mt][3] + cb1.y * sx1;
        }
        wa = wa_n;
        wb = wb_n;
    }

    // ---- reduce across the CTA's warps (fixed order) ---------------------------------------------------
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int i = 0; i < 4; ++i) red[((size_t)warp * MT * 4 + mt * 4 + i) * 32 + lane] = acc[mt][i];
    __syncthreads();
    if (warp != 0) return;
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float s = 0.0f;
#pragma unroll
            for (int w_ = 0; w_ < QMV_WARPS; ++w_) s += red[((size_t)w_ * MT * 4 + mt * 4 + i) * 32 + lane];
            acc[mt][i] = s;
        }
    // combine the per-group columns of one activation row
    float outv[MT][2][2];  // [mt][row half: g / g+8][slot]
    int out_mrow[MT][2];
    int nslots;
    if (CPM == 1) {
        nslots = 2;
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
            outv[mt][0][0] = acc[mt][0]; outv[mt][0][1] = acc[mt][1];
            outv[mt][1][0] = acc[mt][2]; outv[mt][1][1] = acc[mt][3];
            out_mrow[mt][0] = mt * MPM + 2 * t; out_mrow[mt][1] = mt * MPM + 2 * t + 1;
        }
    } else if (CPM == 2) {
        nslots = 1;
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
            outv[mt][0][0] = acc[mt][0] + acc[mt][1];
            outv[mt][1][0] = acc[mt][2] + acc[mt][3];
            outv[mt][0][1] = outv[mt][1][1] = 0.0f;
            out_mrow[mt][0] = mt * MPM + t; out_mrow[mt][1] = -1;
        }
    } else {
        nslots = 1;
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
            float a = acc[mt][0] + acc[mt][1], b = acc[mt][2] + acc[mt][3];
            float a2 = __shfl_xor_sync(0xffffffffu, a, 1), b2 = __shfl_xor_sync(0xffffffffu, b, 1);
            // lanes t and t^1 hold subs {0,1} and {2,3}; add in sub order so both lanes agree bit-for-bit
            outv[mt][0][0] = (t & 1) ? (a2 + a) : (a + a2);
            outv[mt][1][0] = (t & 1) ? (b2 + b) : (b + b2);
            outv[mt][0][1] = outv[mt][1][1] = 0.0f;
            out_mrow[mt][0] = (t & 1) ? -1 : mt * MPM + t / 2; out_mrow[mt][1] = -1;
        }
    }

    auto epilogue_store = [&](uint32_t row, uint32_t mrow, float v) {
        if (row >= p.n || mrow >= p.m) return;
        const size_t oi = (size_t)mrow * p.n + row;
        float value = p.ab_scale * v;
        if (p.accumulate) value += p.d_is_f32 ? reinterpret_cast<float*>(p.d)[oi] : __bfloat162float(reinterpret_cast<__nv_bfloat16*>(p.d)[oi]);
        if (p.bias) value += __bfloat162float(p.bias[row]);
        if (p.has_soft_cap) value = p.soft_cap * tanhf(value / p.soft_cap);
        if (p.d_is_f32) reinterpret_cast<float*>(p.d)[oi] = value;
        else reinterpret_cast<__nv_bfloat16*>(p.d)[oi] = __float2bfloat16_rn(value);
    };

    if (p.kslices == 1) {
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
            for (int s = 0; s < nslots; ++s) {
                if (out_mrow[mt][s] < 0) continue;
                epilogue_store(tile * 16u + g, (uint32_t)out_mrow[mt][s], outv[mt][0][s]);
                epilogue_store(tile * 16u + g + 8, (uint32_t)out_mrow[mt][s], outv[mt][1][s]);
            }
        return;
    }
    // split-K: partials -> workspace [tile][slice][mrow][16 rows]; last CTA of the tile reduces in slice order
    float* wst = p.ws + ((size_t)tile * p.kslices + slice) * (16 * MROWS);
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
        for (int s = 0; s < nslots; ++s) {
            if (out_mrow[mt][s] < 0) continue;
            wst[out_mrow[mt][s] * 16 + g] = outv[mt][0][s];
            wst[out_mrow[mt][s] * 16 + g + 8] = outv[mt][1][s];
        }
    __threadfence();
    __syncwarp();
    unsigned int ticket = 0;
    if (lane == 0) ticket = atomicAdd(&p.counters[tile], 1u);
    ticket = __shfl_sync(0xffffffffu, ticket, 0);
    if (ticket != p.kslices - 1) return;
    __threadfence();
    const float* wt = p.ws + (size_t)tile * p.kslices * (16 * MROWS);
    for (int e = lane; e < 16 * MROWS; e += 32) {
        float s = 0.0f;
        for (uint32_t sl = 0; sl < p.kslices; ++sl) s += __ldcg(wt + (size_t)sl * (16 * MROWS) + e);
        epilogue_store(tile * 16u + (e & 15), (uint32_t)(e >> 4), s);
    }
    if (lane == 0) p.counters[tile] = 0;  // ready for the next launch (stream-ordered)
}

// -------------------------------------------------------------------------------------------------
// Streaming variant (the decode fast path): persistent warps, no CTA-level synchronisation after the
// one-time activation staging. A work item is (16-row tile, k-slice); items are dealt round-robin to
// all resident warps (item id interleaved across CTAs so every SM gets the same share). A warp streams
// its item in "super-chunks" of 4 x 128 nibbles per row: per lane 8 x 16 B of packed weights plus the
// matching scale / zero-point vectors, double-buffered in registers, so ~8 KB per warp are in flight
// while the previous super-chunk is being multiplied. Dequantisation coefficients are derived in
// registers from the vector-loaded scales (no shared-memory round trip, no per-item prologue).
// -------------------------------------------------------------------------------------------------
constexpr int QS_WARPS = 4;
constexpr int QS_SC = 4;   // chunks per super-chunk

template <int NPG>
struct SuperChunk {
    static constexpr int GPS = (QS_SC * 128) / NPG;          // groups per super-chunk: 16, 8, 4, 2
    static constexpr int SW = GPS >= 2 ? GPS / 2 : 1;        // 32-bit words of bf16 scales (and MLX biases)
    uint4 wa[QS_SC], wb[QS_SC];
    uint32_t sa[SW], sb[SW];
    uint32_t ca[SW], cb[SW];                                 // zero points (packed) or biases
};

template <int WORDS>
__device__ __forceinline__ void ldg_words(const void* ptr, uint32_t (&out)[WORDS]) {
    if constexpr (WORDS == 8) {
        const uint4 a = __ldg(reinterpret_cast<const uint4*>(ptr)), b = __ldg(reinterpret_cast<const uint4*>(ptr) + 1);
        out[0] = a.x; out[1] = a.y; out[2] = a.z; out[3] = a.w; out[4] = b.x; out[5] = b.y; out[6] = b.z; out[7] = b.w;
    } else if constexpr (WORDS == 4) {
        const uint4 a = __ldg(reinterpret_cast<const uint4*>(ptr));
        out[0] = a.x; out[1] = a.y; out[2] = a.z; out[3] = a.w;
    } else if constexpr (WORDS == 2) {
        const uint2 a = __ldg(reinterpret_cast<const uint2*>(ptr));
        out[0] = a.x; out[1] = a.y;
    } else {
        out[0] = __ldg(reinterpret_cast<const uint32_t*>(ptr));
    }
}

template <int NPG, int MT>
__global__ void __launch_bounds__(QS_WARPS * 32) qmv_stream_kernel(const QmvParams p) {
    constexpr int CPM = NPG >= 128 ? 1 : 128 / NPG;
    constexpr int MPM = 8 / CPM;
    constexpr int MROWS = MT * MPM;
    using SCk = SuperChunk<NPG>;
    constexpr int GPS = SCk::GPS, SW = SCk::SW;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    pdl_launch_dependents();
    pdl_wait();

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const uint32_t nc_all = p.chunks_total;
    const uint32_t row_items = (nc_all * 16u + 31u) & ~31u;
    const uint32_t ngroups = p.groups_per_row;
    uint4* xs = reinterpret_cast<uint4*>(smem_raw);                                     // [m][row_items]
    float* sx = reinterpret_cast<float*>(smem_raw + (size_t)p.m * row_items * 16);     // [m][ngroups]

    // ---- one-time staging of the activations (whole k range) -------------------------------------------------
    {
        constexpr uint32_t IPG = NPG / 8;
        const uint32_t items_padded = p.m * row_items;
        for (uint32_t it = tid; it < items_padded; it += blockDim.x) {
            uint4 out = make_uint4(0, 0, 0, 0);
            float part = 0.0f;
            const uint32_t r = it / row_items, li = it % row_items;
            const bool in_row = li < nc_all * 16u;
            uint32_t gl = 0xffffffffu;
            if (in_row) {
                const uint32_t pos = li * 8u;   // li = (chunk*4 + t)*4 + w  ->  nibble position = li * 8
                gl = pos / NPG;
                if (pos < p.np) {
                    if (p.bits == 4) {
                        const uint4 v = *reinterpret_cast<const uint4*>(p.x + (size_t)r * p.k + pos);
                        out.x = __byte_perm(v.x, v.z, 0x5410);
                        out.y = __byte_perm(v.x, v.z, 0x7632);
                        out.z = __byte_perm(v.y, v.w, 0x5410);
                        out.w = __byte_perm(v.y, v.w, 0x7632);
                        const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
                        for (int e = 0; e < 4; ++e) { part += __low2float(h2[e]); part += __high2float(h2[e]); }
                    } else {
                        const uint2 v = *reinterpret_cast<const uint2*>(p.x + (size_t)r * p.k + pos / 2);
                        __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&v.x);
                        __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&v.y);
                        const float x0 = __low2float(a), x1 = __high2float(a), x2 = __low2float(b), x3 = __high2float(b);
                        __nv_bfloat162 o0 = __floats2bfloat162_rn(x0, x2);
                        __nv_bfloat162 o1 = __floats2bfloat162_rn(16.0f * x0, 16.0f * x2);
                        __nv_bfloat162 o2 = __floats2bfloat162_rn(x1, x3);
                        __nv_bfloat162 o3 = __floats2bfloat162_rn(16.0f * x1, 16.0f * x3);
                        out.x = *reinterpret_cast<uint32_t*>(&o0);
                        out.y = *reinterpret_cast<uint32_t*>(&o1);
                        out.z = *reinterpret_cast<uint32_t*>(&o2);
                        out.w = *reinterpret_cast<uint32_t*>(&o3);
                        part = ((x0 + x1) + x2) + x3;
                    }
                }
                xs[it] = out;
            }
#pragma unroll
            for (uint32_t o = 1; o < IPG; o <<= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
            if (in_row && (li & (IPG - 1)) == 0 && gl < ngroups) sx[(size_t)r * ngroups + gl] = part;
        }
    }
    __syncthreads();

    const float mult128 = p.bits == 4 ? 128.0f : 128.0f * 17.0f;
    const float sym_mid = p.bits == 4 ? 8.0f : 128.0f;
    const int col_sub = g % CPM, col_mrow = g / CPM;
    const int lane_sub = CPM == 1 ? 0 : (CPM == 2 ? (t >> 1) : t);
    const uint32_t total_warps = gridDim.x * QS_WARPS;
    const uint32_t items_total = ((p.n + 15u) / 16u) * p.kslices;

    for (uint32_t item = blockIdx.x + gridDim.x * warp; item < items_total; item += total_warps) {
        const uint32_t tile = item / p.kslices, slice = item % p.kslices;
        const uint32_t cb = slice * p.chunks_per_slice;
        const uint32_t ce = min(p.chunks_total, cb + p.chunks_per_slice);
        const uint32_t row_a = min(tile * 16u + (uint32_t)g, p.n - 1), row_b = min(tile * 16u + (uint32_t)g + 8u, p.n - 1);
        const uint8_t* wa_base = p.w + (size_t)row_a * p.row_bytes + (size_t)t * 16;
        const uint8_t* wb_base = p.w + (size_t)row_b * p.row_bytes + (size_t)t * 16;

        auto load_super = [&](SCk& sc, uint32_t c0) {
#pragma unroll
            for (int j = 0; j < QS_SC; ++j) {
                const uint32_t c = c0 + j;
                if (c < ce) {
                    sc.wa[j] = ldg_stream_u4(wa_base + (size_t)c * 64);
                    sc.wb[j] = ldg_stream_u4(wb_base + (size_t)c * 64);
                } else {
                    sc.wa[j] = make_uint4(0, 0, 0, 0);
                    sc.wb[j] = make_uint4(0, 0, 0, 0);
                }
            }
            const uint32_t gi = (c0 * 128u) / NPG;   // first group of the super-chunk (multiple of GPS)
            if constexpr (GPS >= 2) {
                ldg_words<SW>(p.scales + (size_t)row_a * ngroups + gi, sc.sa);
                ldg_words<SW>(p.scales + (size_t)row_b * ngroups + gi, sc.sb);
            } else {
                sc.sa[0] = *reinterpret_cast<const uint16_t*>(p.scales + (size_t)row_a * ngroups + gi);
                sc.sb[0] = *reinterpret_cast<const uint16_t*>(p.scales + (size_t)row_b * ngroups + gi);
            }
            if (p.method == UZU_QMETHOD_SCALE_BIAS) {
                if constexpr (GPS >= 2) {
                    ldg_words<SW>(p.biases + (size_t)row_a * ngroups + gi, sc.ca);
                    ldg_words<SW>(p.biases + (size_t)row_b * ngroups + gi, sc.cb);
                } else {
                    sc.ca[0] = *reinterpret_cast<const uint16_t*>(p.biases + (size_t)row_a * ngroups + gi);
                    sc.cb[0] = *reinterpret_cast<const uint16_t*>(p.biases + (size_t)row_b * ngroups + gi);
                }
            } else if (p.method == UZU_QMETHOD_SCALE_ZERO_POINT) {
                // packed zero points of GPS groups: 4-bit -> GPS/2 bytes, 8-bit -> GPS bytes (<= 16 B)
                const uint32_t zbytes = p.bits == 4 ? (GPS + 1) / 2 : GPS;
                const uint8_t* za = p.zero_points + (size_t)row_a * p.zp_stride + (p.bits == 4 ? gi / 2 : gi);
                const uint8_t* zb = p.zero_points + (size_t)row_b * p.zp_stride + (p.bits == 4 ? gi / 2 : gi);
#pragma unroll
                for (int w_ = 0; w_ < SW; ++w_) {
                    if ((uint32_t)w_ * 4u < zbytes) {
                        if (zbytes >= 4) {
                            sc.ca[w_] = __ldg(reinterpret_cast<const uint32_t*>(za) + w_);
                            sc.cb[w_] = __ldg(reinterpret_cast<const uint32_t*>(zb) + w_);
                        } else if (zbytes == 2) {
                            sc.ca[w_] = __ldg(reinterpret_cast<const uint16_t*>(za));
                            sc.cb[w_] = __ldg(reinterpret_cast<const uint16_t*>(zb));
                        } else {
                            sc.ca[w_] = __ldg(za);
                            sc.cb[w_] = __ldg(zb);
                        }
                    } else {
                        sc.ca[w_] = 0; sc.cb[w_] = 0;
                    }
                }
            }
        };

        float acc[MT][4];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[mt][i] = 0.0f;

        // (scale, Sx coefficient) of local group `lg` (compile-time constant) for rows a and b
        auto coef_of = [&](const SCk& sc, int lg, float& s_a, float& z_a, float& s_b, float& z_b) {
            const uint32_t wa_ = sc.sa[lg >> 1], wb_ = sc.sb[lg >> 1];
            s_a = __uint_as_float((lg & 1) ? (wa_ & 0xffff0000u) : (wa_ << 16));
            s_b = __uint_as_float((lg & 1) ? (wb_ & 0xffff0000u) : (wb_ << 16));
            if (p.method == UZU_QMETHOD_SCALE_ZERO_POINT) {
                float zpa, zpb;
                if (p.bits == 4) {
                    zpa = (float)((sc.ca[lg >> 3] >> ((lg & 7) * 4)) & 15u);
                    zpb = (float)((sc.cb[lg >> 3] >> ((lg & 7) * 4)) & 15u);
                } else {
                    zpa = (float)((sc.ca[lg >> 2] >> ((lg & 3) * 8)) & 255u);
                    zpb = (float)((sc.cb[lg >> 2] >> ((lg & 3) * 8)) & 255u);
                }
                z_a = -s_a * (zpa + mult128);
                z_b = -s_b * (zpb + mult128);
            } else if (p.method == UZU_QMETHOD_SCALE_BIAS) {
                const uint32_t ba_ = sc.ca[lg >> 1], bb_ = sc.cb[lg >> 1];
                z_a = __uint_as_float((lg & 1) ? (ba_ & 0xffff0000u) : (ba_ << 16)) - s_a * mult128;
                z_b = __uint_as_float((lg & 1) ? (bb_ & 0xffff0000u) : (bb_ << 16)) - s_b * mult128;
            } else {
                z_a = -s_a * (sym_mid + mult128);
                z_b = -s_b * (sym_mid + mult128);
            }
        };

        auto compute_super = [&](const SCk& sc, uint32_t c0) {
#pragma unroll
            for (int j = 0; j < QS_SC; ++j) {
                const uint32_t c = c0 + j;
                if (c >= ce) break;
                float d[MT][4];
#pragma unroll
                for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                    for (int i = 0; i < 4; ++i) d[mt][i] = 0.0f;
                const uint32_t wav[4] = {sc.wa[j].x ^ p.xor_mask, sc.wa[j].y ^ p.xor_mask, sc.wa[j].z ^ p.xor_mask, sc.wa[j].w ^ p.xor_mask};
                const uint32_t wbv[4] = {sc.wb[j].x ^ p.xor_mask, sc.wb[j].y ^ p.xor_mask, sc.wb[j].z ^ p.xor_mask, sc.wb[j].w ^ p.xor_mask};
#pragma unroll
                for (int w_ = 0; w_ < 4; ++w_) {
                    const uint32_t a0 = nib_pair(wav[w_], 0), a1 = nib_pair(wav[w_], 4), a2 = nib_pair(wav[w_], 8), a3 = nib_pair(wav[w_], 12);
                    const uint32_t b0 = nib_pair(wbv[w_], 0), b1 = nib_pair(wbv[w_], 4), b2 = nib_pair(wbv[w_], 8), b3 = nib_pair(wbv[w_], 12);
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt) {
                        uint4 xb = make_uint4(0, 0, 0, 0);
                        const uint32_t mr = mt * MPM + col_mrow;
                        if (col_sub == lane_sub && mr < p.m) xb = xs[(size_t)mr * row_items + (c * 4 + t) * 4 + w_];
                        mma_16816(d[mt], a0, b0, a1, b1, xb.x, xb.y);
                        mma_16816(d[mt], a2, b2, a3, b3, xb.z, xb.w);
                    }
                }
                // affine part. Local group indices of this thread's two D columns (2t, 2t+1):
                float sa0, za0, sb0, zb0, sa1, za1, sb1, zb1;
                uint32_t gi0, gi1;
                if constexpr (CPM == 1) {
                    constexpr int dummy = 0; (void)dummy;
                    const int lg = (j * 128) / NPG;
                    coef_of(sc, lg, sa0, za0, sb0, zb0);
                    sa1 = sa0; za1 = za0; sb1 = sb0; zb1 = zb0;
                    gi0 = gi1 = (c * 128u) / NPG;
                } else if constexpr (CPM == 2) {
                    coef_of(sc, 2 * j, sa0, za0, sb0, zb0);
                    coef_of(sc, 2 * j + 1, sa1, za1, sb1, zb1);
                    gi0 = c * 2u; gi1 = gi0 + 1u;
                } else {
                    float e0, f0, g0_, h0, e1, f1, g1_, h1;
                    coef_of(sc, 4 * j, sa0, za0, sb0, zb0);
                    coef_of(sc, 4 * j + 1, sa1, za1, sb1, zb1);
                    coef_of(sc, 4 * j + 2, e0, f0, g0_, h0);
                    coef_of(sc, 4 * j + 3, e1, f1, g1_, h1);
                    if (t & 1) { sa0 = e0; za0 = f0; sb0 = g0_; zb0 = h0; sa1 = e1; za1 = f1; sb1 = g1_; zb1 = h1; }
                    gi0 = c * 4u + 2u * (t & 1); gi1 = gi0 + 1u;
                }
                const bool first_chunk_of_group = (NPG <= 128) || ((j & 1) == 0);
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
                    const uint32_t mr0 = mt * MPM + (2 * t) / CPM, mr1 = mt * MPM + (2 * t + 1) / CPM;
                    float sx0 = 0.0f, sx1 = 0.0f;
                    if (first_chunk_of_group) {
                        if (mr0 < p.m) sx0 = sx[(size_t)mr0 * ngroups + gi0];
                        if (mr1 < p.m) sx1 = sx[(size_t)mr1 * ngroups + gi1];
                    }
                    acc[mt][0] += sa0 * d[mt][0] + za0 * sx0;
                    acc[mt][1] += sa1 * d[mt][1] + za1 * sx1;
                    acc[mt][2] += sb0 * d[mt][2] + zb0 * sx0;
                    acc[mt][3] += sb1 * d[mt][3] + zb1 * sx1;
                }
            }
        };

        SCk cur, nxt;
        load_super(cur, cb);
        for (uint32_t c0 = cb; c0 < ce; c0 += QS_SC) {
            const bool more = c0 + QS_SC < ce;
            if (more) load_super(nxt, c0 + QS_SC);
            compute_super(cur, c0);
            if (more) cur = nxt;
        }

        // ---- epilogue of the item (warp-local) ---------------------------------------------------------------------
        float outv[MT][2][2];
        int out_mrow[MT][2];
        int nslots;
        if (CPM == 1) {
            nslots = 2;
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
                outv[mt][0][0] = acc[mt][0]; outv[mt][0][1] = acc[mt][1];
                outv[mt][1][0] = acc[mt][2]; outv[mt][1][1] = acc[mt][3];
                out_mrow[mt][0] = mt * MPM + 2 * t; out_mrow[mt][1] = mt * MPM + 2 * t + 1;
            }
        } else if (CPM == 2) {
            nslots = 1;
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
                outv[mt][0][0] = acc[mt][0] + acc[mt][1];
                outv[mt][1][0] = acc[mt][2] + acc[mt][3];
                outv[mt][0][1] = outv[mt][1][1] = 0.0f;
                out_mrow[mt][0] = mt * MPM + t; out_mrow[mt][1] = -1;
            }
        } else {
            nslots = 1;
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
                float a = acc[mt][0] + acc[mt][1], b = acc[mt][2] + acc[mt][3];
                float a2 = __shfl_xor_sync(0xffffffffu, a, 1), b2 = __shfl_xor_sync(0xffffffffu, b, 1);
                outv[mt][0][0] = (t & 1) ? (a2 + a) : (a + a2);
                outv[mt][1][0] = (t & 1) ? (b2 + b) : (b + b2);
                outv[mt][0][1] = outv[mt][1][1] = 0.0f;
                out_mrow[mt][0] = (t & 1) ? -1 : mt * MPM + t / 2; out_mrow[mt][1] = -1;
            }
        }
        auto epilogue_store = [&](uint32_t row, uint32_t mrow, float v) {
            if (row >= p.n || mrow >= p.m) return;
            const size_t oi = (size_t)mrow * p.n + row;
            float value = p.ab_scale * v;
            if (p.accumulate) value += p.d_is_f32 ? reinterpret_cast<float*>(p.d)[oi] : __bfloat162float(reinterpret_cast<__nv_bfloat16*>(p.d)[oi]);
            if (p.bias) value += __bfloat162float(p.bias[row]);
            if (p.has_soft_cap) value = p.soft_cap * tanhf(value / p.soft_cap);
            if (p.d_is_f32) reinterpret_cast<float*>(p.d)[oi] = value;
            else reinterpret_cast<__nv_bfloat16*>(p.d)[oi] = __float2bfloat16_rn(value);
        };
        if (p.kslices == 1) {
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
                for (int s_ = 0; s_ < nslots; ++s_) {
                    if (out_mrow[mt][s_] < 0) continue;
                    epilogue_store(tile * 16u + g, (uint32_t)out_mrow[mt][s_], outv[mt][0][s_]);
                    epilogue_store(tile * 16u + g + 8, (uint32_t)out_mrow[mt][s_], outv[mt][1][s_]);
                }
            continue;
        }
        float* wst = p.ws + ((size_t)tile * p.kslices + slice) * (16 * MROWS);
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
            for (int s_ = 0; s_ < nslots; ++s_) {
                if (out_mrow[mt][s_] < 0) continue;
                wst[out_mrow[mt][s_] * 16 + g] = outv[mt][0][s_];
                wst[out_mrow[mt][s_] * 16 + g + 8] = outv[mt][1][s_];
            }
        __threadfence();
        __syncwarp();
        unsigned int ticket = 0;
        if (lane == 0) ticket = atomicAdd(&p.counters[tile], 1u);
        ticket = __shfl_sync(0xffffffffu, ticket, 0);
        if (ticket != p.kslices - 1) continue;
        __threadfence();
        const float* wt = p.ws + (size_t)tile * p.kslices * (16 * MROWS);
        for (int e = lane; e < 16 * MROWS; e += 32) {
            float sum = 0.0f;
            for (uint32_t sl = 0; sl < p.kslices; ++sl) sum += __ldcg(wt + (size_t)sl * (16 * MROWS) + e);
            epilogue_store(tile * 16u + (e & 15), (uint32_t)(e >> 4), sum);
        }
        __syncwarp();
        if (lane == 0) p.counters[tile] = 0;
    }
}

// -------------------------------------------------------------------------------------------------
// Decode kernel: the m == 1 specialisation of the streaming variant (the per-token hot path).
// Profiling the generic streaming kernel showed the integer ALU pipe (shift + LOP3 of the nibble
// extraction, coefficient unpacking, register moves), not HBM, as the limiter. For one activation row
// the 8 columns of the MMA tile are free, so the 4 chunks of a super-chunk are steered into *different*
// column pairs (chunk j -> columns 2j, 2j+1 for group size 64; column j for 128): all 32 MMAs of a
// super-chunk accumulate into one 4-register fragment, lane t of a quad ends up owning chunk t, and each
// lane unpacks only the two (scale, zero-point) pairs of its own chunk: the affine work drops 4x, the
// accumulator is zeroed once per 4 KB, and the two register buffers ping-pong without copies.
// -------------------------------------------------------------------------------------------------
// (The register-pipelined instance of this specialisation, qmv_decode_kernel, was removed in round 2: it was reachable only through
// UZU_QMV_REGS after the cp.async variant below replaced it; rows too long for the rings take the streaming kernel above.)

// -------------------------------------------------------------------------------------------------
// cp.async variant of the decode kernel: the same arithmetic, but the weight stream lands in a per-warp
// shared-memory ring (QA_STAGES super-chunks deep) instead of registers. Measured on B200: with the
// GEMV access pattern the HBM pipe needs >100 KB in flight per SM to reach ~6 TB/s (tools/bw_probe.cu);
// registers can hold ~8 KB per warp, the ring holds QA_STAGES-1 x 4.5 KB per warp with no register cost,
// each lane consuming exactly the bytes it copied (no cross-thread synchronisation on the ring).
// -------------------------------------------------------------------------------------------------
#ifndef UZU_QA_ACCS
#define UZU_QA_ACCS 2
#endif
constexpr int QMV_TMA = 3;             // pseudo quantisation method: zero-point form read from the decode-stream layout through TMA bulk copies
constexpr int QA_STAGES = 2;         // base ring depth used for shared-memory sizing; the launch may deepen it (pick_stages)
constexpr uint32_t QA_STAGE_BYTES = 4096 + 512;

// TMA bulk copy + mbarrier helpers of the decode-stream variant (METHOD == QMV_TMA)
__device__ __forceinline__ void qa_mbar_init(uint32_t addr) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(addr) : "memory"); }
__device__ __forceinline__ void qa_mbar_expect_tx(uint32_t addr, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(addr), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool qa_mbar_try_wait(uint32_t addr, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
    return ok != 0;
}
// (weights are read once per token: L2 evict-first keeps the small hot rows -- activations, residual, norm scales -- resident)
__device__ __forceinline__ void qa_bulk_copy(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t mbar, uint64_t policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(mbar), "l"(policy) : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t smem_addr, const void* gptr) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(smem_addr), "l"(gptr) : "memory");
}
__device__ __forceinline__ void cp_async8(uint32_t smem_addr, const void* gptr) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(smem_addr), "l"(gptr) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t smem_addr, const void* gptr) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(smem_addr), "l"(gptr) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }

// METHOD / BITS / FUSED are compile-time so that the hot loop carries one dequantisation variant and the plain GEMV carries no
// prologue / epilogue code: the first single-instance version was ~10.5 K SASS instructions and stalled mostly on instruction
// fetch (ncu: no_instruction was the top stall at 1-2 warps per scheduler).
// PRO: 0 none, 1 RMSNorm (+ residual add), 3 sigmoid gate (prologue 2, gated act, moved to the producer's epilogue); EPI: GatedActMul
template <int NPG, int STAGES, int METHOD, int BITS, int PRO, bool EPI>
__global__ void __launch_bounds__(QS_WARPS * 32) qmv_decode_async_kernel(const QmvParams p) {
    constexpr bool TMA = METHOD == QMV_TMA;
    constexpr uint32_t method = TMA ? (uint32_t)UZU_QMETHOD_SCALE_ZERO_POINT : (uint32_t)METHOD, bits = BITS;
    static_assert(NPG == 64 || NPG == 128, "decode kernel covers int4 gs64 / gs128 and int8 gs64");
    __shared__ uint64_t tma_bars[QS_WARPS * 4];      // TMA mode: one mbarrier per (warp, ring stage)
    constexpr int CPM = NPG >= 128 ? 1 : 2;
    constexpr int GPS = 512 / NPG;    // groups per super-chunk: 8 or 4
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const uint32_t nc_all = p.chunks_total;
    const uint32_t row_items = (nc_all * 16u + 31u) & ~31u;
    const uint32_t ngroups = p.groups_per_row;
    uint4* xs = reinterpret_cast<uint4*>(smem_raw);                            // [row_items] + one zero block of 4 x uint4
    float* sx = reinterpret_cast<float*>(smem_raw + (size_t)(row_items + 4) * 16);   // [ngroups]
    float* red = sx + ((ngroups + 3u) & ~3u);      // [2][QS_WARPS][16]
    // per-warp cp.async ring: STAGES x ([8 vectors][32 lanes] uint4 weights + [4 words][32 lanes] u32 scale / zero-point words)
    uint8_t* ring_base = reinterpret_cast<uint8_t*>(red + 2 * QS_WARPS * 16) + (size_t)warp * STAGES * QA_STAGE_BYTES;
    const uint32_t ring_s = (uint32_t)__cvta_generic_to_shared(ring_base);
    // fused prologue: the produced activation row (bf16 [k]) lives after the rings
    __nv_bfloat16* xrow = reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<uint8_t*>(red + 2 * QS_WARPS * 16) + (size_t)QS_WARPS * STAGES * QA_STAGE_BYTES);
    uint32_t magic;
    asm volatile("mov.b32 %0, 0x43004300;" : "=r"(magic));
    const uint32_t bars_s = (uint32_t)__cvta_generic_to_shared(tma_bars) + (uint32_t)warp * STAGES * 8u;
    uint64_t l2_policy = 0;
    if (TMA) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(l2_policy));
    if (TMA) {
        if (tid < QS_WARPS * STAGES) qa_mbar_init((uint32_t)__cvta_generic_to_shared(tma_bars) + (uint32_t)tid * 8u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        __syncthreads();
    }

    struct Item {
        uint32_t tile, slice, cb, ce;
        const uint8_t *wa_base, *wb_base;
        const __nv_bfloat16 *sa_base, *sb_base;
        const uint8_t *za_base, *zb_base;      // zero points (bytes) or MLX biases
        uint32_t za_off, zb_off;               // byte offset of this lane's first zero-point byte at super-chunk 0
    };
    const float mult128 = bits == 4 ? 128.0f : 128.0f * 17.0f;
    const float sym_mid = bits == 4 ? 8.0f : 128.0f;
    const bool lane_has_groups = 2 * t < GPS;       // NPG = 128: only lanes t = 0, 1 own columns
    // which chunk of a super-chunk this lane feeds into the MMA B operand (column n = g)
    const int b_chunk = CPM == 2 ? (g >> 1) : g;
    const bool b_lane = CPM == 2 ? ((g & 1) == (t >> 1)) : true;

    auto setup = [&](Item& it, uint32_t item) {
        it.tile = item / p.kslices;
        it.slice = item % p.kslices;
        it.cb = it.slice * p.chunks_per_slice;
        it.ce = min(p.chunks_total, it.cb + p.chunks_per_slice);
        if (TMA) return;                 // the stream is addressed by (tile, super-chunk) alone
        const uint32_t row_a = min(it.tile * 16u + (uint32_t)g, p.n - 1), row_b = min(it.tile * 16u + (uint32_t)g + 8u, p.n - 1);
        it.wa_base = p.w + (size_t)row_a * p.row_bytes + (size_t)t * 16;
        it.wb_base = p.w + (size_t)row_b * p.row_bytes + (size_t)t * 16;
        it.sa_base = p.scales + (size_t)row_a * ngroups + 2 * t;
        it.sb_base = p.scales + (size_t)row_b * ngroups + 2 * t;
        if (method == UZU_QMETHOD_SCALE_BIAS) {
            it.za_base = reinterpret_cast<const uint8_t*>(p.biases + (size_t)row_a * ngroups + 2 * t);
            it.zb_base = reinterpret_cast<const uint8_t*>(p.biases + (size_t)row_b * ngroups + 2 * t);
        } else if (bits == 4) {
            it.za_base = p.zero_points + (size_t)row_a * p.zp_stride + t;
            it.zb_base = p.zero_points + (size_t)row_b * p.zp_stride + t;
        } else {
            it.za_base = p.zero_points + (size_t)row_a * p.zp_stride + 2 * t;
            it.zb_base = p.zero_points + (size_t)row_b * p.zp_stride + 2 * t;
        }
        it.za_off = row_a * p.zp_stride + (bits == 4 ? t : 2 * t);
        it.zb_off = row_b * p.zp_stride + (bits == 4 ? t : 2 * t);
    };
    // one stage = one super-chunk of this warp: 8 x 16 B of packed weights per lane + its scale / zero-point words,
    // copied global -> shared with cp.async (no registers held while in flight)
    auto issue_stage = [&](const Item& it, uint32_t c0, uint32_t slot) {
        if (TMA) {
            // one 4608-byte unit of the decode stream = exactly this stage (codes in lane order + scale / zero-point words)
            if (lane == 0) {
                qa_mbar_expect_tx(bars_s + slot * 8u, QA_STAGE_BYTES);
                qa_bulk_copy(ring_s + slot * QA_STAGE_BYTES, p.stream + ((size_t)it.tile * p.stream_C + c0 / 4u) * QA_STAGE_BYTES, QA_STAGE_BYTES, bars_s + slot * 8u, l2_policy);
            }
            return;
        }
        const uint32_t sbase = ring_s + slot * QA_STAGE_BYTES + (uint32_t)lane * 16u;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            cp_async16(sbase + (uint32_t)j * 512u, it.wa_base + (size_t)(c0 + j) * 64);
            cp_async16(sbase + (uint32_t)(4 + j) * 512u, it.wb_base + (size_t)(c0 + j) * 64);
        }
        if (lane_has_groups) {
            const uint32_t wbase = ring_s + slot * QA_STAGE_BYTES + 4096u + (uint32_t)lane * 4u;
            const uint32_t gi = (c0 * 128u) / NPG;
            cp_async4(wbase, it.sa_base + gi);
            cp_async4(wbase + 128u, it.sb_base + gi);
            if (method == UZU_QMETHOD_SCALE_ZERO_POINT) {
                // copy the aligned 32-bit word that holds this lane's zero-point byte(s); the consumer shifts them out
                const uint32_t goff = bits == 4 ? gi / 2 : gi;
                cp_async4(wbase + 256u, p.zero_points + ((it.za_off + goff) & ~3u));
                cp_async4(wbase + 384u, p.zero_points + ((it.zb_off + goff) & ~3u));
            } else if (method == UZU_QMETHOD_SCALE_BIAS) {
                cp_async4(wbase + 256u, it.za_base + (size_t)gi * 2);
                cp_async4(wbase + 384u, it.zb_base + (size_t)gi * 2);
            }
        }
    };

    // ---- work distribution ------------------------------------------------------------------------------------
    // A CTA owns an item = (group of TPC row tiles, global k-slice). Inside the item the k range of a tile is interleaved
    // over WPT warps by super-chunk (so the CTA reads WPT x 256 contiguous bytes of every row per step) and the WPT
    // partial sums meet in shared memory: no global atomics unless the host asked for a global k split (few-row shapes).
    const uint32_t WPT = p.warps_per_tile, TPC = QS_WARPS / WPT;
    const uint32_t tile_local = warp / WPT, kpart = warp % WPT;
    const uint32_t tiles = (p.n + 15u) / 16u;
    // paired mode (GatedActMul epilogue; host guarantees kslices == 1 and WPT <= 2): the first TPC/2 tiles of a CTA item are
    // `up` tiles q, the other TPC/2 the matching `gate` tiles q + F/16, so both halves of an output row finish in the same CTA
    constexpr bool paired = EPI;
    const uint32_t hp = TPC / 2u;
    const uint32_t tile_groups = paired ? (p.pair_tiles + hp - 1) / hp : (tiles + TPC - 1) / TPC;
    const uint32_t items_cta = tile_groups * p.kslices;
    const uint32_t item_first = blockIdx.x;
    auto next_item = [&](uint32_t it_) { return it_ + gridDim.x; };

    auto setup_cta = [&](Item& it, uint32_t item) {
        if (paired) {
            const uint32_t base = item * hp + tile_local % hp;
            setup(it, min(base, p.pair_tiles - 1u) + (tile_local / hp) * p.pair_tiles);
            if (base >= p.pair_tiles) it.tile = 0xffffffffu;
            return;
        }
        const uint32_t tg = item / p.kslices;
        setup(it, (tg * TPC + tile_local) * p.kslices + item % p.kslices);
    };
    auto first_sc = [&](const Item& it) { return it.cb + kpart * 4u; };

    // Two cursors walk this warp's stages (item ascending, super-chunk ascending): `ic` issues cp.async stages STAGES-1
    // ahead of `cur`, the one being consumed, across item boundaries. The weights do not depend on the previous kernel, so
    // the first stages are requested before waiting on the producer of the activations (programmatic dependent launch).
    struct Cursor { uint32_t item, c0; Item it; };
    auto cursor_begin = [&](Cursor& c, uint32_t item) {
        c.item = item;
        if (item < items_cta) { setup_cta(c.it, item); c.c0 = first_sc(c.it); }
    };
    auto cursor_has = [&](const Cursor& c) { return c.item < items_cta && c.it.tile < tiles && c.c0 < c.it.ce; };
    Cursor ic;
    cursor_begin(ic, item_first);
    uint32_t issued = 0;
    auto issue_next = [&]() {
        while (ic.item < items_cta && !cursor_has(ic)) cursor_begin(ic, next_item(ic.item));
        if (ic.item < items_cta) {
            issue_stage(ic.it, ic.c0, issued % STAGES);
            ic.c0 += 4u * WPT;
        }
        if (!TMA) cp_async_commit();       // always commit (possibly empty) so wait_group<STAGES-1> means "oldest stage landed"
        ++issued;
    };
#pragma unroll
    for (int s_ = 0; s_ < STAGES - 1; ++s_) issue_next();
    Item cur;
    uint32_t item = item_first;
    uint32_t consumed = 0;
    if (PRO == 1) {
        // the norm scales are static weights: pull them towards L2 while the producer of the activations is still running
        for (uint32_t line = blockIdx.x * blockDim.x + tid; line < p.k / 32u; line += gridDim.x * blockDim.x)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(p.pro_scales + (size_t)line * 32));
    }
    pdl_launch_dependents();
    pdl_wait();
    if (tid < 4) xs[row_items + tid] = make_uint4(0, 0, 0, 0);

    const __nv_bfloat16* xsrc = p.x;
    if (PRO != 0) {
        // ---- fused prologue: every CTA recomputes the (tiny) activation row; same arithmetic and rounding points as the
        // standalone kernels (normalization.rs:50-125, gated_act_mul/mod.rs:5-12, sigmoid_gate.rs:7-22) -------------------------
        const uint32_t K = p.k;
        constexpr int PB = 2;   // items (8 elements each) per thread per batch: all global loads of a batch are issued before use
        const uint32_t stride = blockDim.x * 8u;
        if (PRO == 1) {
            float ssq = 0.0f;
            for (uint32_t base = tid * 8u; base < K; base += stride * PB) {
                uint4 va[PB], vb[PB];
#pragma unroll
                for (int u = 0; u < PB; ++u) {
                    const uint32_t i = base + u * stride;
                    va[u] = make_uint4(0, 0, 0, 0); vb[u] = make_uint4(0, 0, 0, 0);
                    if (i < K) {
                        va[u] = *reinterpret_cast<const uint4*>(p.pro_a + i);
                        if (p.pro_residual_add) vb[u] = *reinterpret_cast<const uint4*>(p.pro_b + i);
                    }
                }
#pragma unroll
                for (int u = 0; u < PB; ++u) {
                    const uint32_t i = base + u * stride;
                    if (i >= K) break;
                    __nv_bfloat162* a2 = reinterpret_cast<__nv_bfloat162*>(&va[u]);
                    if (p.pro_residual_add) {
                        const __nv_bfloat162* b2 = reinterpret_cast<const __nv_bfloat162*>(&vb[u]);
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            a2[e] = __floats2bfloat162_rn(__fadd_rn(__low2float(a2[e]), __low2float(b2[e])), __fadd_rn(__high2float(a2[e]), __high2float(b2[e])));
                    }
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float lo = __low2float(a2[e]), hi = __high2float(a2[e]);
                        ssq = __fadd_rn(ssq, __fmul_rn(lo, lo));
                        ssq = __fadd_rn(ssq, __fmul_rn(hi, hi));
                    }
                    *reinterpret_cast<uint4*>(xrow + i) = va[u];
                    if (p.pro_shortcut_out && blockIdx.x == 0) *reinterpret_cast<uint4*>(p.pro_shortcut_out + i) = va[u];
                }
            }
            ssq = warp_sum(ssq);
            if (lane == 0) red[warp] = ssq;
            __syncthreads();
            float tot = 0.0f;
#pragma unroll
            for (int w_ = 0; w_ < QS_WARPS; ++w_) tot = __fadd_rn(tot, red[w_]);
            const float rms_inv = __frcp_rn(__fsqrt_rn(__fadd_rn(__fdiv_rn(tot, (float)K), p.pro_eps)));
            for (uint32_t base = tid * 8u; base < K; base += stride * PB) {
                float4 s0[PB], s1[PB];
#pragma unroll
                for (int u = 0; u < PB; ++u) {
                    const uint32_t i = base + u * stride;
                    if (i < K) {
                        s0[u] = __ldg(reinterpret_cast<const float4*>(p.pro_scales + i));
                        s1[u] = __ldg(reinterpret_cast<const float4*>(p.pro_scales + i + 4));
                    }
                }
#pragma unroll
                for (int u = 0; u < PB; ++u) {
                    const uint32_t i = base + u * stride;
                    if (i >= K) break;
                    uint4 va = *reinterpret_cast<const uint4*>(xrow + i);
                    __nv_bfloat162* a2 = reinterpret_cast<__nv_bfloat162*>(&va);
                    const float sc[8] = {s0[u].x, s0[u].y, s0[u].z, s0[u].w, s1[u].x, s1[u].y, s1[u].z, s1[u].w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        float v[2] = {__low2float(a2[e]), __high2float(a2[e])};
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const float normalized = __fmul_rn(v[h], rms_inv);
                            const float so = __fadd_rn(sc[2 * e + h], p.pro_scale_offset);
                            if (p.pro_full_layer) v[h] = __fmul_rn(normalized, so);
                            else v[h] = __fmul_rn(__bfloat162float(__float2bfloat16_rn(normalized)), __bfloat162float(__float2bfloat16_rn(so)));
                        }
                        a2[e] = __floats2bfloat162_rn(v[0], v[1]);
                    }
                    *reinterpret_cast<uint4*>(xrow + i) = va;
                }
            }
        } else {
            // prologue 3: x = attn * sigmoid(gate)
            const __nv_bfloat16* pv = p.pro_a;
            const __nv_bfloat16* pg = p.pro_b;
            for (uint32_t base = tid * 8u; base < K; base += stride * PB) {
                uint4 vv[PB], vg[PB];
#pragma unroll
                for (int u = 0; u < PB; ++u) {
                    const uint32_t i = base + u * stride;
                    vv[u] = make_uint4(0, 0, 0, 0); vg[u] = make_uint4(0, 0, 0, 0);
                    if (i < K) {
                        vv[u] = *reinterpret_cast<const uint4*>(pv + i);
                        vg[u] = *reinterpret_cast<const uint4*>(pg + i);
                    }
                }
#pragma unroll
                for (int u = 0; u < PB; ++u) {
                    const uint32_t i = base + u * stride;
                    if (i >= K) break;
                    __nv_bfloat162* v2 = reinterpret_cast<__nv_bfloat162*>(&vv[u]);
                    const __nv_bfloat162* g2 = reinterpret_cast<const __nv_bfloat162*>(&vg[u]);
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float m0 = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-__low2float(g2[e]))));
                        const float m1 = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-__high2float(g2[e]))));
                        v2[e] = __floats2bfloat162_rn(__fmul_rn(__low2float(v2[e]), m0), __fmul_rn(__high2float(v2[e]), m1));
                    }
                    *reinterpret_cast<uint4*>(xrow + i) = vv[u];
                }
            }
        }
        __syncthreads();
        xsrc = xrow;
    }

    {   // one-time staging of the activation row: all global loads of a batch are issued before any is consumed
        constexpr uint32_t IPG = NPG / 8;
        constexpr int BATCH = 4;
        for (uint32_t base = tid; base < row_items; base += blockDim.x * BATCH) {
            uint4 v[BATCH];
#pragma unroll
            for (int u = 0; u < BATCH; ++u) {
                const uint32_t it = base + u * blockDim.x;
                const uint32_t pos = it * 8u;
                v[u] = make_uint4(0, 0, 0, 0);
                if (it < nc_all * 16u && pos < p.np) {
                    if (bits == 4) v[u] = *reinterpret_cast<const uint4*>(xsrc + pos);
                    else {
                        const uint2 h = *reinterpret_cast<const uint2*>(xsrc + pos / 2);
                        v[u].x = h.x; v[u].y = h.y;
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < BATCH; ++u) {
                const uint32_t it = base + u * blockDim.x;
                if (it >= row_items) break;      // warp-uniform: row_items is a multiple of 32
                const uint32_t pos = it * 8u;
                uint4 out;
                float part = 0.0f;
                if (bits == 4) {
                    out.x = __byte_perm(v[u].x, v[u].z, 0x5410);
                    out.y = __byte_perm(v[u].x, v[u].z, 0x7632);
                    out.z = __byte_perm(v[u].y, v[u].w, 0x5410);
                    out.w = __byte_perm(v[u].y, v[u].w, 0x7632);
                    const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&v[u]);
#pragma unroll
                    for (int e = 0; e < 4; ++e) { part += __low2float(h2[e]); part += __high2float(h2[e]); }
                } else {
                    __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&v[u].x);
                    __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&v[u].y);
                    const float x0 = __low2float(a), x1 = __high2float(a), x2 = __low2float(b), x3 = __high2float(b);
                    __nv_bfloat162 o0 = __floats2bfloat162_rn(x0, x2);
                    __nv_bfloat162 o1 = __floats2bfloat162_rn(16.0f * x0, 16.0f * x2);
                    __nv_bfloat162 o2 = __floats2bfloat162_rn(x1, x3);
                    __nv_bfloat162 o3 = __floats2bfloat162_rn(16.0f * x1, 16.0f * x3);
                    out.x = *reinterpret_cast<uint32_t*>(&o0);
                    out.y = *reinterpret_cast<uint32_t*>(&o1);
                    out.z = *reinterpret_cast<uint32_t*>(&o2);
                    out.w = *reinterpret_cast<uint32_t*>(&o3);
                    part = ((x0 + x1) + x2) + x3;
                }
                if (it < nc_all * 16u) xs[it] = out;
#pragma unroll
                for (uint32_t o = 1; o < IPG; o <<= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
                const uint32_t gl = pos / NPG;
                if ((it & (IPG - 1)) == 0 && gl < ngroups) sx[gl] = part;
            }
        }
    }
    __syncthreads();

    uint32_t parity = 0;
    for (; item < items_cta; item = next_item(item), parity ^= 1u) {
        setup_cta(cur, item);
        const uint32_t tile = cur.tile, slice = cur.slice, ce = cur.ce;
        const uint32_t step = 4u * WPT;

        float acc0 = 0.0f, acc1 = 0.0f, acc2 = 0.0f, acc3 = 0.0f;

        auto compute = [&](uint32_t slot, uint32_t c0) {
            const uint4* sw = reinterpret_cast<const uint4*>(ring_base + (size_t)slot * QA_STAGE_BYTES) + lane;
            const uint32_t* swd = reinterpret_cast<const uint32_t*>(ring_base + (size_t)slot * QA_STAGE_BYTES + 4096) + lane;
            // two independent accumulator fragments halve the dependent HMMA chain (32 MMAs per super-chunk)
            float d[4] = {0.0f, 0.0f, 0.0f, 0.0f}, d2[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#if UZU_QA_ACCS == 4
            float d3[4] = {0.0f, 0.0f, 0.0f, 0.0f}, d4[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#endif
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                uint4 va = sw[j * 32], vb = sw[(4 + j) * 32];
                if (p.xor_mask) {
                    va.x ^= p.xor_mask; va.y ^= p.xor_mask; va.z ^= p.xor_mask; va.w ^= p.xor_mask;
                    vb.x ^= p.xor_mask; vb.y ^= p.xor_mask; vb.z ^= p.xor_mask; vb.w ^= p.xor_mask;
                }
                const uint32_t wav[4] = {va.x, va.y, va.z, va.w};
                const uint32_t wbv[4] = {vb.x, vb.y, vb.z, vb.w};
                const bool feeds = b_lane && b_chunk == j;
                // lanes that do not feed this chunk's column read a shared zero block (no predication, no register zeroing)
                const uint4* xrow = feeds ? xs + ((size_t)(c0 + j) * 4 + t) * 4 : xs + row_items;
#pragma unroll
                for (int w_ = 0; w_ < 4; ++w_) {
                    const uint4 xb = xrow[w_];
                    const uint32_t a0 = nib_pair_fast(wav[w_], 0, magic), a1 = nib_pair_fast(wav[w_], 4, magic), a2 = nib_pair_fast(wav[w_], 8, magic), a3 = nib_pair_fast(wav[w_], 12, magic);
                    const uint32_t b0 = nib_pair_fast(wbv[w_], 0, magic), b1 = nib_pair_fast(wbv[w_], 4, magic), b2 = nib_pair_fast(wbv[w_], 8, magic), b3 = nib_pair_fast(wbv[w_], 12, magic);
#if UZU_QA_ACCS == 4
                    if (w_ & 1) {
                        mma_16816(d3, a0, b0, a1, b1, xb.x, xb.y);
                        mma_16816(d4, a2, b2, a3, b3, xb.z, xb.w);
                    } else
#endif
                    {
                        mma_16816(d, a0, b0, a1, b1, xb.x, xb.y);
                        mma_16816(d2, a2, b2, a3, b3, xb.z, xb.w);
                    }
                }
            }
#if UZU_QA_ACCS == 4
#pragma unroll
            for (int i = 0; i < 4; ++i) { d[i] += d3[i]; d2[i] += d4[i]; }
#endif
#pragma unroll
            for (int i = 0; i < 4; ++i) d[i] += d2[i];
            // this lane's two D columns (2t, 2t+1) hold groups gi + 2t and gi + 2t + 1 of rows g (d0, d1) and g+8 (d2, d3)
            if (lane_has_groups) {
                const uint32_t gi = (c0 * 128u) / NPG + 2 * t;
                const float2 sxv = *reinterpret_cast<const float2*>(sx + gi);
                const uint32_t bsa = swd[0], bsb = swd[32];
                uint32_t bca = swd[64], bcb = swd[96];
                if (method == UZU_QMETHOD_SCALE_ZERO_POINT && !TMA) {
                    const uint32_t goff = bits == 4 ? ((c0 * 128u) / NPG) / 2 : (c0 * 128u) / NPG;
                    bca >>= ((cur.za_off + goff) & 3u) * 8u;
                    bcb >>= ((cur.zb_off + goff) & 3u) * 8u;
                }
                const float sa0 = __uint_as_float(bsa << 16), sa1 = __uint_as_float(bsa & 0xffff0000u);
                const float sb0 = __uint_as_float(bsb << 16), sb1 = __uint_as_float(bsb & 0xffff0000u);
                if (method == UZU_QMETHOD_SCALE_ZERO_POINT) {
                    float ka0, ka1, kb0, kb1;   // value = s * (d + k * Sx)
                    if (TMA) {              // the stream carries the zero points as exact bf16 pairs
                        ka0 = -(__uint_as_float(bca << 16) + mult128); ka1 = -(__uint_as_float(bca & 0xffff0000u) + mult128);
                        kb0 = -(__uint_as_float(bcb << 16) + mult128); kb1 = -(__uint_as_float(bcb & 0xffff0000u) + mult128);
                    } else if (bits == 4) {
                        ka0 = -((float)(bca & 15u) + mult128); ka1 = -((float)((bca >> 4) & 15u) + mult128);
                        kb0 = -((float)(bcb & 15u) + mult128); kb1 = -((float)((bcb >> 4) & 15u) + mult128);
                    } else {
                        ka0 = -((float)(bca & 255u) + mult128); ka1 = -((float)((bca >> 8) & 255u) + mult128);
                        kb0 = -((float)(bcb & 255u) + mult128); kb1 = -((float)((bcb >> 8) & 255u) + mult128);
                    }
                    acc0 += sa0 * (d[0] + ka0 * sxv.x);
                    acc1 += sa1 * (d[1] + ka1 * sxv.y);
                    acc2 += sb0 * (d[2] + kb0 * sxv.x);
                    acc3 += sb1 * (d[3] + kb1 * sxv.y);
                } else if (method == UZU_QMETHOD_SCALE_BIAS) {
                    const float ba0 = __uint_as_float(bca << 16), ba1 = __uint_as_float(bca & 0xffff0000u);
                    const float bb0 = __uint_as_float(bcb << 16), bb1 = __uint_as_float(bcb & 0xffff0000u);
                    acc0 += sa0 * (d[0] - mult128 * sxv.x) + ba0 * sxv.x;
                    acc1 += sa1 * (d[1] - mult128 * sxv.y) + ba1 * sxv.y;
                    acc2 += sb0 * (d[2] - mult128 * sxv.x) + bb0 * sxv.x;
                    acc3 += sb1 * (d[3] - mult128 * sxv.y) + bb1 * sxv.y;
                } else {
                    const float k = -(sym_mid + mult128);
                    acc0 += sa0 * (d[0] + k * sxv.x);
                    acc1 += sa1 * (d[1] + k * sxv.y);
                    acc2 += sb0 * (d[2] + k * sxv.x);
                    acc3 += sb1 * (d[3] + k * sxv.y);
                }
            }
        };

        if (tile < tiles) {
            for (uint32_t c0 = first_sc(cur); c0 < ce; c0 += step) {
                // refill the slot consumed last iteration first (each lane only ever reads back its own cp.async bytes, so
                // there is no cross-lane hazard), then wait until the oldest of the STAGES outstanding stages has landed
                if (TMA) {
                    // every lane is done with the slot about to be refilled (generic-proxy reads before the async-proxy write)
                    __syncwarp();
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    issue_next();
                    const uint32_t bar = bars_s + (consumed % STAGES) * 8u, parity = (consumed / STAGES) & 1u;
                    uint32_t spins = 0;
                    while (!qa_mbar_try_wait(bar, parity))
                        if (++spins > (1u << 24)) __trap();          // a protocol bug must fail the launch, not hang the GPU
                } else {
                    issue_next();
                    cp_async_wait<STAGES - 1>();
                }
                compute(consumed % STAGES, c0);
                ++consumed;
            }
        }

        // ---- quad reduction (fixed order), CTA reduction through shared memory, epilogue --------------------------------
        float ra = acc0 + acc1, rb = acc2 + acc3;
        ra += __shfl_xor_sync(0xffffffffu, ra, 1);
        rb += __shfl_xor_sync(0xffffffffu, rb, 1);
        ra += __shfl_xor_sync(0xffffffffu, ra, 2);
        rb += __shfl_xor_sync(0xffffffffu, rb, 2);
        float* redp = red + parity * (QS_WARPS * 16);
        if (WPT > 1 || paired) {
            if (t == 0) {
                redp[warp * 16 + g] = ra;
                redp[warp * 16 + g + 8] = rb;
            }
            __syncthreads();
        }
        auto epilogue_store = [&](uint32_t row, float v) {
            if (row >= p.n) return;
            float value = p.ab_scale * v;
            if (p.accumulate) value += p.d_is_f32 ? reinterpret_cast<float*>(p.d)[row] : __bfloat162float(reinterpret_cast<__nv_bfloat16*>(p.d)[row]);
            if (p.bias) value += __bfloat162float(p.bias[row]);
            if (p.has_soft_cap) value = p.soft_cap * tanhf(value / p.soft_cap);
            if (p.d_is_f32) reinterpret_cast<float*>(p.d)[row] = value;
            else reinterpret_cast<__nv_bfloat16*>(p.d)[row] = __float2bfloat16_rn(value);
        };
        if (paired) {
            // warp f < TPC/2 finishes pair f: matmul outputs rounded to bf16 (+ bias) exactly as the unfused GEMV stores them,
            // then GatedActMul's two roundings (gated_act_mul/mod.rs:5-12)
            const uint32_t my_pair = item * hp + warp;
            if (warp < hp && my_pair < p.pair_tiles && lane < 16) {
                float up = 0.0f, gate = 0.0f;
                for (uint32_t kp = 0; kp < WPT; ++kp) {
                    up += redp[(warp * WPT + kp) * 16 + lane];
                    gate += redp[((warp + hp) * WPT + kp) * 16 + lane];
                }
                const uint32_t row = my_pair * 16u + lane;
                up = p.ab_scale * up; gate = p.ab_scale * gate;
                if (p.bias) { up += __bfloat162float(p.bias[row]); gate += __bfloat162float(p.bias[row + p.pair_tiles * 16u]); }
                up = __bfloat162float(__float2bfloat16_rn(up));
                gate = __bfloat162float(__float2bfloat16_rn(gate));
                const float m_ = __bfloat162float(__float2bfloat16_rn(act_f32_nofma(p.epi_act, gate)));
                reinterpret_cast<__nv_bfloat16*>(p.d)[row] = __float2bfloat16_rn(__fmul_rn(up, m_));
            }
        } else if (WPT == 1) {
            // every warp owns a whole tile: finish it from registers, no CTA-level synchronisation at all
            if (tile < tiles) {
                if (p.kslices == 1) {
                    if (t == 0) {
                        epilogue_store(tile * 16u + g, ra);
                        epilogue_store(tile * 16u + g + 8, rb);
                    }
                } else {
                    float* wst = p.ws + ((size_t)tile * p.kslices + slice) * 16;
                    if (t == 0) { wst[g] = ra; wst[g + 8] = rb; }
                    __threadfence();
                    __syncwarp();
                    unsigned int ticket = 0;
                    if (lane == 0) ticket = atomicAdd(&p.counters[tile], 1u);
                    ticket = __shfl_sync(0xffffffffu, ticket, 0);
                    if (ticket == p.kslices - 1) {
                        __threadfence();
                        if (lane < 16) {
                            const float* wt = p.ws + (size_t)tile * p.kslices * 16;
                            float total = 0.0f;
                            for (uint32_t sl = 0; sl < p.kslices; ++sl) total += __ldcg(wt + (size_t)sl * 16 + lane);
                            epilogue_store(tile * 16u + lane, total);
                        }
                        __syncwarp();
                        if (lane == 0) p.counters[tile] = 0;
                    }
                }
            }
        } else if (warp < TPC) {
            // warp w < TPC finishes tile_local = w: lanes 0..15 own one output row each
            const uint32_t my_tile = (item / p.kslices) * TPC + warp;
            if (my_tile < tiles) {
                float sum = 0.0f;
                if (lane < 16)
                    for (uint32_t kp = 0; kp < WPT; ++kp) sum += redp[(warp * WPT + kp) * 16 + lane];
                if (p.kslices == 1) {
                    if (lane < 16) epilogue_store(my_tile * 16u + lane, sum);
                } else {
                    float* wst = p.ws + ((size_t)my_tile * p.kslices + slice) * 16;
                    if (lane < 16) wst[lane] = sum;
                    __threadfence();
                    __syncwarp();
                    unsigned int ticket = 0;
                    if (lane == 0) ticket = atomicAdd(&p.counters[my_tile], 1u);
                    ticket = __shfl_sync(0xffffffffu, ticket, 0);
                    if (ticket == p.kslices - 1) {
                        __threadfence();
                        if (lane < 16) {
                            const float* wt = p.ws + (size_t)my_tile * p.kslices * 16;
                            float total = 0.0f;
                            for (uint32_t sl = 0; sl < p.kslices; ++sl) total += __ldcg(wt + (size_t)sl * 16 + lane);
                            epilogue_store(my_tile * 16u + lane, total);
                        }
                        __syncwarp();
                        if (lane == 0) p.counters[my_tile] = 0;
                    }
                }
            }
        }
        (void)slice;
    }
}

// -------------------------------------------------------------------------------------------------
// Rows kernel: 2..16 activation rows share ONE pass over the weights (speculation passes over a trie, multi-sequence batched decode,
// the tail rows of a prefill chunk). Same weight stream and dot engine as the decode kernel (per-warp cp.async rings, mma.sync
// m16n8k16 over (128 + code) bf16 pairs, hoisted per-group affine), but the 8 MMA columns carry 8 ACTIVATION ROWS instead of being
// steered to quantisation groups. For that the groups must be separable without spending columns: a lane copies bytes [8t, 8t+8) and
// [32 + 8t, 32 + 8t + 8) of every 64-byte chunk of a weight row (two 8-byte cp.async instead of one 16-byte one), so words 0-1 of
// every lane lie in the first 64 nibbles of the chunk and words 2-3 in the second: with 64-nibble groups the MMAs of words 0-1
// accumulate group 2j and those of words 2-3 group 2j+1 of chunk j (128-nibble groups: the whole chunk is one group). Per super-chunk
// a warp issues 32 x MB MMAs (MB = 1 for <= 8 rows, 2 for <= 16) where the older streaming kernel needed 32 x 4 for 16 rows.
// A CTA (8 warps, one per SM) is pinned to one k-slice (grid is a multiple of kslices), stages the m x slice activations once
// (bf16 B-fragment layout + per-group sums, row stride padded to an odd number of 16-byte items so the 8 rows of a fragment load hit
// 8 different bank groups) and walks tile groups; the WPT warps of a tile meet in shared memory, k-slices in the split-k workspace.
// -------------------------------------------------------------------------------------------------
#ifndef UZU_QMV_ROWS_DEFAULT_WARPS
#define UZU_QMV_ROWS_DEFAULT_WARPS 16
#endif
constexpr int RW_WARPS = 8;      // warps per CTA of the default variant (the host sizes rings / reduction buffers with the launch's own count)

template <int NPG, int STAGES, int METHOD, int BITS, int MB, int WARPS, int FEED>
__global__ void __launch_bounds__(WARPS * 32, 1) qmv_rows_kernel(const QmvParams p) {
    constexpr uint32_t method = (uint32_t)METHOD, bits = BITS;
    static_assert(NPG == 64 || NPG == 128, "rows kernel covers int4 gs64 / gs128 and int8 gs64");
    constexpr int GPS = 512 / NPG;          // groups per super-chunk: 8 or 4
    constexpr int MROWS = MB * 8;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const uint32_t ngroups = p.groups_per_row;
    const uint32_t kslices = p.kslices, cps = p.chunks_per_slice;
    const uint32_t slice = blockIdx.x % kslices;                 // constant over this CTA's items (host: gridDim.x % kslices == 0)
    const uint32_t cb = slice * cps, ce = min(p.chunks_total, cb + cps);
    const uint32_t slice_items = cps * 16u;                      // 16-byte B-fragment items per activation row and slice
    const uint32_t xstride = slice_items + 1u;                   // odd: rows g and g+1 of a fragment load fall on different bank groups
    const uint32_t sgroups = (cps * 128u) / NPG;                 // groups per slice
    uint4* xs = reinterpret_cast<uint4*>(smem_raw);                                              // [MROWS][xstride]
    float* sx = reinterpret_cast<float*>(smem_raw + (size_t)MROWS * xstride * 16);               // [MROWS][sgroups]
    const uint32_t sxs = sgroups | 1u;                           // odd row stride: the four activation-row pairs of a quad read different banks
    float* red = sx + (size_t)MROWS * sxs;                                                   // [WARPS][MROWS][16]
    uint8_t* ring_base = reinterpret_cast<uint8_t*>(red + WARPS * MROWS * 16) + (size_t)warp * STAGES * QA_STAGE_BYTES;
    const uint32_t ring_s = (uint32_t)__cvta_generic_to_shared(ring_base);
    uint32_t magic;
    asm volatile("mov.b32 %0, 0x43004300;" : "=r"(magic));

    const float mult128 = bits == 4 ? 128.0f : 128.0f * 17.0f;
    const float sym_mid = bits == 4 ? 8.0f : 128.0f;
    const bool lane_has_groups = 2 * t < GPS;

    const uint32_t WPT = p.warps_per_tile, TPC = WARPS / WPT;
    const uint32_t tile_local = warp / WPT, kpart = warp % WPT;
    const uint32_t tiles = (p.n + 15u) / 16u;
    const uint32_t tile_groups = (tiles + TPC - 1) / TPC;
    const uint32_t items_cta = tile_groups * kslices;

    struct Item {
        uint32_t tile;
        const uint8_t *wa_base, *wb_base;
        const __nv_bfloat16 *sa_base, *sb_base;
        const uint8_t *za_base, *zb_base;
        uint32_t za_row, zb_row;               // byte offset of the row's first zero-point byte
    };
    auto setup = [&](Item& it, uint32_t item) {
        it.tile = (item / kslices) * TPC + tile_local;
        const uint32_t tl = min(it.tile, tiles - 1u);
        const uint32_t row_a = min(tl * 16u + (uint32_t)g, p.n - 1), row_b = min(tl * 16u + (uint32_t)g + 8u, p.n - 1);
        it.wa_base = p.w + (size_t)row_a * p.row_bytes + (size_t)t * (FEED ? 16 : 8);
        it.wb_base = p.w + (size_t)row_b * p.row_bytes + (size_t)t * (FEED ? 16 : 8);
        it.sa_base = p.scales + (size_t)row_a * ngroups + 2 * t;
        it.sb_base = p.scales + (size_t)row_b * ngroups + 2 * t;
        if (method == UZU_QMETHOD_SCALE_BIAS) {
            it.za_base = reinterpret_cast<const uint8_t*>(p.biases + (size_t)row_a * ngroups + 2 * t);
            it.zb_base = reinterpret_cast<const uint8_t*>(p.biases + (size_t)row_b * ngroups + 2 * t);
        } else {
            it.za_base = it.zb_base = nullptr;
        }
        it.za_row = row_a * p.zp_stride;
        it.zb_row = row_b * p.zp_stride;
    };
    // one stage = one super-chunk of this warp's tile: per lane 4 chunks x 2 rows x (8 + 8) bytes of codes, then the coefficient words
    auto issue_stage = [&](const Item& it, uint32_t c0, uint32_t slot) {
        const uint32_t sbase = ring_s + slot * QA_STAGE_BYTES + (uint32_t)lane * 16u;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint8_t* sa_ = it.wa_base + (size_t)(c0 + j) * 64;
            const uint8_t* sb_ = it.wb_base + (size_t)(c0 + j) * 64;
            if (FEED) {      // 16 contiguous bytes per lane (L2-only cp.async.cg); the consumer regroups the halves with one quad exchange
                cp_async16(sbase + (uint32_t)j * 512u, sa_);
                cp_async16(sbase + (uint32_t)(4 + j) * 512u, sb_);
            } else {
                cp_async8(sbase + (uint32_t)j * 512u, sa_);
                cp_async8(sbase + (uint32_t)j * 512u + 8u, sa_ + 32);
                cp_async8(sbase + (uint32_t)(4 + j) * 512u, sb_);
                cp_async8(sbase + (uint32_t)(4 + j) * 512u + 8u, sb_ + 32);
            }
        }
        if (lane_has_groups) {
            const uint32_t wbase = ring_s + slot * QA_STAGE_BYTES + 4096u + (uint32_t)lane * 4u;
            const uint32_t gi = (c0 * 128u) / NPG;
            cp_async4(wbase, it.sa_base + gi);
            cp_async4(wbase + 128u, it.sb_base + gi);
            if (method == UZU_QMETHOD_SCALE_ZERO_POINT) {
                const uint32_t goff = (bits == 4 ? gi / 2 + (uint32_t)t : gi + 2u * (uint32_t)t);
                cp_async4(wbase + 256u, p.zero_points + ((it.za_row + goff) & ~3u));
                cp_async4(wbase + 384u, p.zero_points + ((it.zb_row + goff) & ~3u));
            } else if (method == UZU_QMETHOD_SCALE_BIAS) {
                cp_async4(wbase + 256u, it.za_base + (size_t)gi * 2);
                cp_async4(wbase + 384u, it.zb_base + (size_t)gi * 2);
            }
        }
    };
    auto first_sc = [&]() { return cb + kpart * 4u; };

    struct Cursor { uint32_t item, c0; Item it; };
    auto cursor_begin = [&](Cursor& c, uint32_t item) {
        c.item = item;
        if (item < items_cta) { setup(c.it, item); c.c0 = first_sc(); }
    };
    auto cursor_has = [&](const Cursor& c) { return c.item < items_cta && c.it.tile < tiles && c.c0 < ce; };
    Cursor ic;
    cursor_begin(ic, blockIdx.x);
    uint32_t issued = 0;
    auto issue_next = [&]() {
        while (ic.item < items_cta && !cursor_has(ic)) cursor_begin(ic, ic.item + gridDim.x);
        if (ic.item < items_cta) {
            issue_stage(ic.it, ic.c0, issued % STAGES);
            ic.c0 += 4u * WPT;
        }
        cp_async_commit();
        ++issued;
    };
#pragma unroll
    for (int s_ = 0; s_ < STAGES - 1; ++s_) issue_next();      // weights do not depend on the producer of the activations
    pdl_launch_dependents();
    pdl_wait();

    {   // ---- one-time staging of this CTA's k-slice of the m activation rows (rows >= m: zeros). All global loads of a batch are
        // issued before any is consumed: the loop is L2-latency-bound otherwise (one dependent round trip per item) ----------------
        constexpr uint32_t IPG = NPG / 8;
        constexpr int BATCH = 8;
        const uint32_t total = MROWS * slice_items;
        for (uint32_t base = tid; base < total; base += blockDim.x * BATCH) {
            uint4 v[BATCH];
#pragma unroll
            for (int u = 0; u < BATCH; ++u) {
                const uint32_t it = base + u * blockDim.x;
                v[u] = make_uint4(0, 0, 0, 0);
                if (it < total) {
                    const uint32_t r = it / slice_items, li = it % slice_items;
                    const uint32_t pos = (cb * 16u + li) * 8u;          // nibble position in the row
                    if (r < p.m && pos < p.np) {
                        if (bits == 4) v[u] = *reinterpret_cast<const uint4*>(p.x + (size_t)r * p.k + pos);
                        else {
                            const uint2 h = *reinterpret_cast<const uint2*>(p.x + (size_t)r * p.k + pos / 2);
                            v[u].x = h.x; v[u].y = h.y;
                        }
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < BATCH; ++u) {
                const uint32_t it = base + u * blockDim.x;
                if (it >= total) break;                  // warp-uniform: total and blockDim.x are multiples of 32
                const uint32_t r = it / slice_items, li = it % slice_items;
                uint4 out;
                float part = 0.0f;
                if (bits == 4) {
                    out.x = __byte_perm(v[u].x, v[u].z, 0x5410);
                    out.y = __byte_perm(v[u].x, v[u].z, 0x7632);
                    out.z = __byte_perm(v[u].y, v[u].w, 0x5410);
                    out.w = __byte_perm(v[u].y, v[u].w, 0x7632);
                    const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&v[u]);
#pragma unroll
                    for (int e = 0; e < 4; ++e) { part += __low2float(h2[e]); part += __high2float(h2[e]); }
                } else {
                    __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&v[u].x);
                    __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&v[u].y);
                    const float x0 = __low2float(a), x1 = __high2float(a), x2 = __low2float(b), x3 = __high2float(b);
                    __nv_bfloat162 o0 = __floats2bfloat162_rn(x0, x2);
                    __nv_bfloat162 o1 = __floats2bfloat162_rn(16.0f * x0, 16.0f * x2);
                    __nv_bfloat162 o2 = __floats2bfloat162_rn(x1, x3);
                    __nv_bfloat162 o3 = __floats2bfloat162_rn(16.0f * x1, 16.0f * x3);
                    out.x = *reinterpret_cast<uint32_t*>(&o0);
                    out.y = *reinterpret_cast<uint32_t*>(&o1);
                    out.z = *reinterpret_cast<uint32_t*>(&o2);
                    out.w = *reinterpret_cast<uint32_t*>(&o3);
                    part = ((x0 + x1) + x2) + x3;
                }
                xs[(size_t)r * xstride + li] = out;
#pragma unroll
                for (uint32_t o = 1; o < IPG; o <<= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
                if ((li & (IPG - 1)) == 0) sx[(size_t)r * sxs + li / IPG] = part;
            }
        }
    }
    __syncthreads();

    uint32_t consumed = 0;
    for (uint32_t item = blockIdx.x; item < items_cta; item += gridDim.x) {
        Item cur;
        setup(cur, item);
        const uint32_t tile = cur.tile;
        float acc[MB][4];
#pragma unroll
        for (int c = 0; c < MB; ++c) { acc[c][0] = acc[c][1] = acc[c][2] = acc[c][3] = 0.0f; }

        if (tile < tiles) {
            for (uint32_t c0 = first_sc(); c0 < ce; c0 += 4u * WPT) {
                // the coefficient words of a stage are read by all four lanes of a quad: nobody refills the slot consumed last iteration
                // before every lane is done with it, and nobody reads the new stage before every lane's copies have landed
                __syncwarp();
                issue_next();
                cp_async_wait<STAGES - 1>();
                __syncwarp();
                const uint32_t slot = consumed % STAGES;
                ++consumed;
                const uint4* sw = reinterpret_cast<const uint4*>(ring_base + (size_t)slot * QA_STAGE_BYTES) + lane;
                const uint4* swc = reinterpret_cast<const uint4*>(ring_base + (size_t)slot * QA_STAGE_BYTES + 4096);
                // coefficient words of rows g / g+8 for all groups of the super-chunk: the four lanes of this quad hold them (word [w][g*4 + t'])
                const uint4 sa4 = swc[g], sb4 = swc[8 + g];
                uint4 za4 = make_uint4(0, 0, 0, 0), zb4 = make_uint4(0, 0, 0, 0);
                if (method != UZU_QMETHOD_SCALE_SYMMETRIC) { za4 = swc[16 + g]; zb4 = swc[24 + g]; }
                const uint32_t sav[4] = {sa4.x, sa4.y, sa4.z, sa4.w}, sbv[4] = {sb4.x, sb4.y, sb4.z, sb4.w};
                const uint32_t zav[4] = {za4.x, za4.y, za4.z, za4.w}, zbv[4] = {zb4.x, zb4.y, zb4.z, zb4.w};
                const uint32_t gi0 = (c0 * 128u) / NPG;                     // first group of the super-chunk (row-global)
                const uint32_t gl0 = gi0 - (cb * 128u) / NPG;               // ... and slice-local (index into sx)

                // affine of one finished group fragment: out[wrow][arow] += s[wrow] * (d + k[wrow] * Sx[arow]); gl = group within the super-chunk
                auto affine = [&](const float (&d)[MB][4], int gl) {
                    const int tq = gl >> 1, hi = gl & 1;                    // quad lane t' holding the pair (2t', 2t'+1), which half
                    const uint32_t bsa = sav[tq], bsb = sbv[tq];
                    const float sa_ = hi ? __uint_as_float(bsa & 0xffff0000u) : __uint_as_float(bsa << 16);
                    const float sb_ = hi ? __uint_as_float(bsb & 0xffff0000u) : __uint_as_float(bsb << 16);
                    float ka, kb, ba = 0.0f, bb = 0.0f;
                    if (method == UZU_QMETHOD_SCALE_ZERO_POINT) {
                        const uint32_t goff = bits == 4 ? gi0 / 2 + (uint32_t)tq : gi0 + 2u * (uint32_t)tq;
                        const uint32_t wa_ = zav[tq] >> (((cur.za_row + goff) & 3u) * 8u), wb_ = zbv[tq] >> (((cur.zb_row + goff) & 3u) * 8u);
                        if (bits == 4) {
                            ka = -((float)((wa_ >> (hi * 4)) & 15u) + mult128);
                            kb = -((float)((wb_ >> (hi * 4)) & 15u) + mult128);
                        } else {
                            ka = -((float)((wa_ >> (hi * 8)) & 255u) + mult128);
                            kb = -((float)((wb_ >> (hi * 8)) & 255u) + mult128);
                        }
                    } else if (method == UZU_QMETHOD_SCALE_BIAS) {
                        ka = kb = -mult128;
                        ba = hi ? __uint_as_float(zav[tq] & 0xffff0000u) : __uint_as_float(zav[tq] << 16);
                        bb = hi ? __uint_as_float(zbv[tq] & 0xffff0000u) : __uint_as_float(zbv[tq] << 16);
                    } else {
                        ka = kb = -(sym_mid + mult128);
                    }
#pragma unroll
                    for (int c = 0; c < MB; ++c) {
                        const float s0 = sx[(size_t)(c * 8 + 2 * t) * sxs + gl0 + gl], s1 = sx[(size_t)(c * 8 + 2 * t + 1) * sxs + gl0 + gl];
                        if (method == UZU_QMETHOD_SCALE_BIAS) {
                            acc[c][0] += sa_ * (d[c][0] + ka * s0) + ba * s0;
                            acc[c][1] += sa_ * (d[c][1] + ka * s1) + ba * s1;
                            acc[c][2] += sb_ * (d[c][2] + kb * s0) + bb * s0;
                            acc[c][3] += sb_ * (d[c][3] + kb * s1) + bb * s1;
                        } else {
                            acc[c][0] += sa_ * (d[c][0] + ka * s0);
                            acc[c][1] += sa_ * (d[c][1] + ka * s1);
                            acc[c][2] += sb_ * (d[c][2] + kb * s0);
                            acc[c][3] += sb_ * (d[c][3] + kb * s1);
                        }
                    }
                };

#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    uint4 va = sw[j * 32], vb = sw[(4 + j) * 32];
                    if (p.xor_mask) {
                        va.x ^= p.xor_mask; va.y ^= p.xor_mask; va.z ^= p.xor_mask; va.w ^= p.xor_mask;
                        vb.x ^= p.xor_mask; vb.y ^= p.xor_mask; vb.z ^= p.xor_mask; vb.w ^= p.xor_mask;
                    }
                    if (FEED) {
                        // lane t holds nibbles [32t, 32t + 32) of the chunk: lanes t < 2 keep their first two words and take the first two of
                        // lane t + 2, lanes t >= 2 keep their last two and take the last two of lane t - 2 -> words 0-1 lie in the first
                        // 64 nibbles of the chunk for every lane, words 2-3 in the second
                        const bool lo = t < 2;
                        const uint32_t s0 = lo ? va.z : va.x, s1 = lo ? va.w : va.y, s2 = lo ? vb.z : vb.x, s3 = lo ? vb.w : vb.y;
                        const uint32_t r0 = __shfl_xor_sync(0xffffffffu, s0, 2), r1 = __shfl_xor_sync(0xffffffffu, s1, 2);
                        const uint32_t r2 = __shfl_xor_sync(0xffffffffu, s2, 2), r3 = __shfl_xor_sync(0xffffffffu, s3, 2);
                        if (lo) { va.z = r0; va.w = r1; vb.z = r2; vb.w = r3; }
                        else { va.x = r0; va.y = r1; vb.x = r2; vb.y = r3; }
                    }
                    const uint32_t wav[4] = {va.x, va.y, va.z, va.w};
                    const uint32_t wbv[4] = {vb.x, vb.y, vb.z, vb.w};
                    // four independent accumulator chains per group half (column block x {first, second} MMA of a word): asm volatile keeps
                    // the issue order, so consecutive MMAs must not share an accumulator (the HMMA result latency is ~10 issue slots)
                    float dA[MB][4], dB[MB][4], eA[MB][4], eB[MB][4];
#pragma unroll
                    for (int c = 0; c < MB; ++c) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) { dA[c][i] = 0.0f; dB[c][i] = 0.0f; eA[c][i] = 0.0f; eB[c][i] = 0.0f; }
                    }
                    // B fragment of word w_: activation items of chunk (c0 + j): first half items [8 * half + 2t + (w_ & 1)]
                    // (FEED: after the exchange lane t's words 0-1 are items 4 (t & 1) + 2 (t >> 1) + {0, 1} of the chunk's first half)
                    const uint32_t xi = ((c0 + j) - cb) * 16u + (FEED ? 4u * (uint32_t)(t & 1) + 2u * (uint32_t)(t >> 1) : 2u * (uint32_t)t);
#pragma unroll
                    for (int w_ = 0; w_ < 4; ++w_) {
                        const uint32_t a0 = nib_pair_fast(wav[w_], 0, magic), a1 = nib_pair_fast(wav[w_], 4, magic), a2 = nib_pair_fast(wav[w_], 8, magic), a3 = nib_pair_fast(wav[w_], 12, magic);
                        const uint32_t b0 = nib_pair_fast(wbv[w_], 0, magic), b1 = nib_pair_fast(wbv[w_], 4, magic), b2 = nib_pair_fast(wbv[w_], 8, magic), b3 = nib_pair_fast(wbv[w_], 12, magic);
                        const uint32_t li = xi + (uint32_t)(w_ >> 1) * 8u + (uint32_t)(w_ & 1);
                        uint4 xb[MB];
#pragma unroll
                        for (int c = 0; c < MB; ++c) xb[c] = xs[(size_t)(c * 8 + g) * xstride + li];
                        if (NPG == 64 && w_ >= 2) {
#pragma unroll
                            for (int c = 0; c < MB; ++c) mma_16816(dB[c], a0, b0, a1, b1, xb[c].x, xb[c].y);
#pragma unroll
                            for (int c = 0; c < MB; ++c) mma_16816(eB[c], a2, b2, a3, b3, xb[c].z, xb[c].w);
                        } else {
#pragma unroll
                            for (int c = 0; c < MB; ++c) mma_16816(dA[c], a0, b0, a1, b1, xb[c].x, xb[c].y);
#pragma unroll
                            for (int c = 0; c < MB; ++c) mma_16816(eA[c], a2, b2, a3, b3, xb[c].z, xb[c].w);
                        }
                    }
#pragma unroll
                    for (int c = 0; c < MB; ++c) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) { dA[c][i] += eA[c][i]; dB[c][i] += eB[c][i]; }
                    }
                    if (NPG == 64) {
                        affine(dA, 2 * j);
                        affine(dB, 2 * j + 1);
                    } else {
                        affine(dA, j);
                    }
                }
            }
        }

        // ---- the WPT warps of a tile meet in shared memory; k-slices meet in the split-k workspace ------------------------------
        {
            float* rw = red + (size_t)warp * (MROWS * 16);
#pragma unroll
            for (int c = 0; c < MB; ++c) {
                rw[(c * 8 + 2 * t) * 16 + g] = acc[c][0];
                rw[(c * 8 + 2 * t + 1) * 16 + g] = acc[c][1];
                rw[(c * 8 + 2 * t) * 16 + g + 8] = acc[c][2];
                rw[(c * 8 + 2 * t + 1) * 16 + g + 8] = acc[c][3];
            }
        }
        // WPT == 1: a warp owns its tile outright -- the shared-memory round trip is only the fragment -> row-major transposition, and the
        // warps of the CTA never wait for each other (no bubble at item boundaries); WPT > 1: CTA barrier, warp w < TPC finishes tile w
        if (WPT == 1) __syncwarp(); else __syncthreads();
        if ((uint32_t)warp < TPC) {
            const uint32_t my_tile = (item / kslices) * TPC + (uint32_t)warp;
            if (my_tile < tiles) {
                constexpr int OUTS = MROWS * 16 / 32;      // outputs per lane: o = lane + 32 i -> activation row o / 16, weight row o % 16
                float sum[OUTS];
#pragma unroll
                for (int i = 0; i < OUTS; ++i) {
                    sum[i] = 0.0f;
                    for (uint32_t kp = 0; kp < WPT; ++kp) sum[i] += red[(size_t)((uint32_t)warp * WPT + kp) * (MROWS * 16) + lane + 32 * i];
                }
                auto store = [&](int i, float v) {
                    const uint32_t o = (uint32_t)lane + 32u * (uint32_t)i;
                    const uint32_t arow = o / 16u, row = my_tile * 16u + (o % 16u);
                    if (arow >= p.m || row >= p.n) return;
                    const size_t idx = (size_t)arow * p.n + row;
                    float value = p.ab_scale * v;
                    if (p.accumulate) value += p.d_is_f32 ? reinterpret_cast<float*>(p.d)[idx] : __bfloat162float(reinterpret_cast<__nv_bfloat16*>(p.d)[idx]);
                    if (p.bias) value += __bfloat162float(p.bias[row]);
                    if (p.has_soft_cap) value = p.soft_cap * tanhf(value / p.soft_cap);
                    if (p.d_is_f32) reinterpret_cast<float*>(p.d)[idx] = value;
                    else reinterpret_cast<__nv_bfloat16*>(p.d)[idx] = __float2bfloat16_rn(value);
                };
                if (kslices == 1) {
#pragma unroll
                    for (int i = 0; i < OUTS; ++i) store(i, sum[i]);
                } else {
                    float* wst = p.ws + ((size_t)my_tile * kslices + slice) * (MROWS * 16);
#pragma unroll
                    for (int i = 0; i < OUTS; ++i) wst[lane + 32 * i] = sum[i];
                    __threadfence();
                    __syncwarp();
                    unsigned int ticket = 0;
                    if (lane == 0) ticket = atomicAdd(&p.counters[my_tile], 1u);
                    ticket = __shfl_sync(0xffffffffu, ticket, 0);
                    if (ticket == kslices - 1) {
                        __threadfence();
                        const float* wt = p.ws + (size_t)my_tile * kslices * (MROWS * 16);
#pragma unroll
                        for (int i = 0; i < OUTS; ++i) {
                            float total = 0.0f;
                            for (uint32_t sl = 0; sl < kslices; ++sl) total += __ldcg(wt + (size_t)sl * (MROWS * 16) + lane + 32 * i);
                            store(i, total);
                        }
                        __syncwarp();
                        if (lane == 0) p.counters[my_tile] = 0;
                    }
                }
            }
        }
        if (WPT == 1) __syncwarp(); else __syncthreads();       // `red` is single-buffered
    }
}

// -------------------------------------------------------------------------------------------------
// Generic kernel: one warp per output element, the reference loop verbatim (any dtype / layout / gather).
// -------------------------------------------------------------------------------------------------
struct GenericParams {
    uzu_matmul_args a;
};

__device__ __forceinline__ float load_as_f32(const void* base, uint32_t dt, size_t i) {
    return dt == UZU_DT_F32 ? reinterpret_cast<const float*>(base)[i] : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(base)[i]);
}

__global__ void __launch_bounds__(256) generic_kernel(const uzu_matmul_args a) {
    const int lane = threadIdx.x & 31;
    const size_t wid = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const size_t total = (size_t)a.m * a.n;
    if (wid >= total) return;
    const uint32_t row = (uint32_t)(wid / a.n), col = (uint32_t)(wid % a.n);
    const size_t b_col = a.gather_indices ? reinterpret_cast<const uint32_t*>(a.gather_indices)[(size_t)row * a.n + col] : col;
    const bool quant = a.b_prologue != UZU_B_FULL_PRECISION;
    const uint32_t bits = a.b_mode == UZU_QMODE_U4 ? 4 : 8;
    const uint32_t groups = quant ? (a.k + a.b_group_size - 1) / a.b_group_size : 0;
    const uint32_t zp_stride = bits == 4 ? (groups + 1) / 2 : groups;
    float acc = 0.0f;
    for (uint32_t k = lane; k < a.k; k += 32) {
        float av = load_as_f32((const void*)a.a, a.input_dt, (size_t)row * a.k + k);
        float bv;
        if (!quant) {
            size_t ld = a.b_leading_dimension ? a.b_leading_dimension : (a.b_transpose ? a.k : a.n);
            size_t idx = a.b_transpose ? b_col * ld + k : (size_t)k * ld + b_col;
            bv = load_as_f32((const void*)a.b, a.weights_dt, idx);
        } else {
            size_t lin = b_col * (size_t)a.k + k;
            uint32_t code;
            const uint8_t* wbytes = reinterpret_cast<const uint8_t*>(a.b);
            if (bits == 4) {
                uint8_t byte = wbytes[lin >> 1];
                code = (lin & 1) ? (byte >> 4) : (byte & 15);
                if (a.b_signed_codes) code ^= 8u;
            } else {
                code = wbytes[lin];
                if (a.b_signed_codes) code ^= 128u;
            }
            uint32_t gi = k / a.b_group_size;
            float scale = load_as_f32((const void*)a.b_scales, a.weights_dt, b_col * groups + gi);
            float corr;
            if (a.b_prologue == UZU_B_SCALE_ZERO_POINT_DEQUANT) {
                const uint8_t* zp = reinterpret_cast<const uint8_t*>(a.b_zero_points);
                float z;
                if (bits == 4) {
                    uint8_t byte = zp[b_col * zp_stride + (gi >> 1)];
                    z = (float)((gi & 1) ? (byte >> 4) : (byte & 15));
                } else {
                    z = (float)zp[b_col * zp_stride + gi];
                }
                corr = -scale * z;
            } else if (a.b_prologue == UZU_B_SCALE_BIAS_DEQUANT) {
                corr = load_as_f32((const void*)a.b_biases, a.weights_dt, b_col * groups + gi);
            } else {
                corr = -scale * (float)(1u << (bits - 1));
            }
            bv = scale * (float)code + corr;
        }
        acc += av * bv;
    }
    acc = warp_sum(acc);
    if (lane != 0) return;
    const size_t oi = (size_t)row * a.n + col;
    float value = ((a.d_transform & UZU_D_SCALE) ? a.ab_scale : 1.0f) * acc;
    if (a.d_transform & UZU_D_ACCUMULATE) value += load_as_f32((const void*)a.d, a.output_dt, oi);
    if ((a.d_transform & UZU_D_BIAS) && a.bias) value += load_as_f32((const void*)a.bias, a.weights_dt, col);
    if (a.d_transform & UZU_D_SOFT_CAP) value = a.soft_cap * tanhf(value / a.soft_cap);
    if (a.output_dt == UZU_DT_F32) reinterpret_cast<float*>(a.d)[oi] = value;
    else reinterpret_cast<__nv_bfloat16*>(a.d)[oi] = __float2bfloat16_rn(value);
}

// -------------------------------------------------------------------------------------------------
// host dispatch
// -------------------------------------------------------------------------------------------------
static bool dt_ok(uint32_t dt) { return dt == UZU_DT_BF16 || dt == UZU_DT_F32; }

static const char* validate(const uzu_matmul_args* a) {
    if (!a) return "null arguments";
    if (!dt_ok(a->weights_dt) || !dt_ok(a->input_dt) || !dt_ok(a->output_dt)) return "unsupported data type (bf16 | f32 only)";
    if (a->m == 0 || a->n == 0 || a->k == 0) return "empty shape";
    if (!a->a || !a->b || !a->d) return "null operand";
    if (a->d_transform & UZU_D_RHT) {
        if (!a->rht_factors) return "UZU_D_RHT without rht_factors";
        if (a->n % 32 != 0) return "output RHT needs n to be a multiple of HADAMARD_TRANSFORM_BLOCK_SIZE (32)";
        if (a->gather_indices) return "output RHT with gather_indices is undefined (the transform mixes 32 neighbouring columns)";
    }
    if (a->b_prologue > UZU_B_SCALE_SYMMETRIC_DEQUANT) return "bad b_prologue";
    if (a->b_prologue != UZU_B_FULL_PRECISION) {
        if (!a->b_transpose) return "quantized B must be [n,k] (b_transpose)";
        if (!a->b_scales) return "quantized B requires scales";
        if (a->b_mode != UZU_QMODE_U4 && a->b_mode != UZU_QMODE_U8) return "quantization mode must be U4 or U8";
        if (a->b_group_size == 0) return "group size is zero";
        if (a->b_prologue == UZU_B_SCALE_ZERO_POINT_DEQUANT && !a->b_zero_points) return "missing zero points";
        if (a->b_prologue == UZU_B_SCALE_BIAS_DEQUANT && !a->b_biases) return "missing biases";
        if (a->b_mode == UZU_QMODE_U4 && (a->k & 1)) return "4-bit rows need an even k";
        if (a->weights_dt != UZU_DT_BF16 && a->weights_dt != UZU_DT_F32) return "bad scale dtype";
    }
    return nullptr;
}

static size_t qmv_smem_bytes(uint32_t m, uint32_t nc, uint32_t ngl, int mt) {
    const uint32_t coef_stride = ((ngl + 15u) & ~15u) + 2u;
    return (size_t)m * (((size_t)nc * 16 + 31) & ~(size_t)31) * 16 + (size_t)16 * coef_stride * 8 + (size_t)m * ngl * 4 + (size_t)QMV_WARPS * mt * 4 * 32 * 4 + 64;
}

template <int NPG, int MT>
static void launch_qmv(uzu_command_buffer* cmd, const QmvParams& p, uint32_t tiles) {
    constexpr int CPM = NPG >= 128 ? 1 : 128 / NPG;
    constexpr int MROWS = MT * (8 / CPM);
    const uint32_t nc = p.chunks_per_slice;
    const uint32_t ngl = (nc * 128u + NPG - 1) / NPG + 1;
    (void)MROWS;
    size_t smem = qmv_smem_bytes(p.m, nc, ngl, MT);
    static std::atomic<uint64_t> smem_opt_in{0};
    opt_in_dynamic_smem(cmd, qmv_kernel<NPG, MT>, (int)(200 * 1024), smem_opt_in);
    qmv_kernel<NPG, MT><<<tiles * p.kslices, QMV_WARPS * 32, smem, cmd->ctx->stream>>>(p);
    after_launch(cmd, "qmv_kernel");
}

template <int NPG, int STAGES, int METHOD, int BITS, int PRO, bool EPI>
static void launch_qmv_decode_async_s(uzu_command_buffer* cmd, const QmvParams& p, uint32_t grid, size_t smem) {
    static std::atomic<uint64_t> smem_opt_in{0};
    opt_in_dynamic_smem(cmd, qmv_decode_async_kernel<NPG, STAGES, METHOD, BITS, PRO, EPI>, (int)(200 * 1024), smem_opt_in);
    launch(cmd, "qmv_decode_async_kernel", qmv_decode_async_kernel<NPG, STAGES, METHOD, BITS, PRO, EPI>, dim3(grid), dim3(QS_WARPS * 32), smem, p);
}
template <int NPG, int METHOD, int BITS, int PRO, bool EPI>
static void launch_qmv_decode_async_i(uzu_command_buffer* cmd, const QmvParams& p, uint32_t grid, size_t smem) {
    switch (p.stages) {
        case 4: launch_qmv_decode_async_s<NPG, 4, METHOD, BITS, PRO, EPI>(cmd, p, grid, smem); break;
        case 3: launch_qmv_decode_async_s<NPG, 3, METHOD, BITS, PRO, EPI>(cmd, p, grid, smem); break;
        default: launch_qmv_decode_async_s<NPG, 2, METHOD, BITS, PRO, EPI>(cmd, p, grid, smem); break;
    }
}
template <int NPG, int PRO, bool EPI>
static void launch_qmv_decode_async_f(uzu_command_buffer* cmd, const QmvParams& p, uint32_t grid, size_t smem) {
    if (p.bits == 4) {
        switch (p.method) {
            case QMV_TMA: launch_qmv_decode_async_i<NPG, QMV_TMA, 4, PRO, EPI>(cmd, p, grid, smem); break;
            case UZU_QMETHOD_SCALE_ZERO_POINT: launch_qmv_decode_async_i<NPG, UZU_QMETHOD_SCALE_ZERO_POINT, 4, PRO, EPI>(cmd, p, grid, smem); break;
            case UZU_QMETHOD_SCALE_BIAS: launch_qmv_decode_async_i<NPG, UZU_QMETHOD_SCALE_BIAS, 4, PRO, EPI>(cmd, p, grid, smem); break;
            default: launch_qmv_decode_async_i<NPG, UZU_QMETHOD_SCALE_SYMMETRIC, 4, PRO, EPI>(cmd, p, grid, smem); break;
        }
    } else {
        if constexpr (PRO == 0 && !EPI) {     // the fused variants are instantiated for 4-bit weights only (host: fused_linear_supported)
            switch (p.method) {
                case UZU_QMETHOD_SCALE_ZERO_POINT: launch_qmv_decode_async_i<NPG, UZU_QMETHOD_SCALE_ZERO_POINT, 8, 0, false>(cmd, p, grid, smem); break;
                case UZU_QMETHOD_SCALE_BIAS: launch_qmv_decode_async_i<NPG, UZU_QMETHOD_SCALE_BIAS, 8, 0, false>(cmd, p, grid, smem); break;
                default: launch_qmv_decode_async_i<NPG, UZU_QMETHOD_SCALE_SYMMETRIC, 8, 0, false>(cmd, p, grid, smem); break;
            }
        }
    }
}
template <int NPG>
static void launch_qmv_decode_async(uzu_command_buffer* cmd, const QmvParams& p, uint32_t grid, size_t smem) {
    const uint32_t key = p.prologue * 2u + (p.epi_gated ? 1u : 0u);
    switch (key) {
        case 0: launch_qmv_decode_async_f<NPG, 0, false>(cmd, p, grid, smem); break;
        case 1: launch_qmv_decode_async_f<NPG, 0, true>(cmd, p, grid, smem); break;
        case 2: launch_qmv_decode_async_f<NPG, 1, false>(cmd, p, grid, smem); break;
        case 3: launch_qmv_decode_async_f<NPG, 1, true>(cmd, p, grid, smem); break;
        default: launch_qmv_decode_async_f<NPG, 3, false>(cmd, p, grid, smem); break;   // 6: sigmoid gate
    }
}

template <int NPG, int STAGES, int METHOD, int BITS, int MB, int WARPS, int FEED>
static void launch_qmv_rows_f(uzu_command_buffer* cmd, const QmvParams& p, uint32_t grid, size_t smem) {
    static std::atomic<uint64_t> smem_opt_in{0};
    opt_in_dynamic_smem(cmd, qmv_rows_kernel<NPG, STAGES, METHOD, BITS, MB, WARPS, FEED>, (int)(227 * 1024), smem_opt_in);
    launch(cmd, "qmv_rows_kernel", qmv_rows_kernel<NPG, STAGES, METHOD, BITS, MB, WARPS, FEED>, dim3(grid), dim3(WARPS * 32), smem, p);
}
template <int NPG, int STAGES, int METHOD, int BITS, int MB, int WARPS>
static void launch_qmv_rows_s(uzu_command_buffer* cmd, const QmvParams& p, uint32_t grid, size_t smem) {
    // FEED 1 (16-byte cp.async.cg + one quad exchange) measured 7-8 % faster than FEED 0 (two 8-byte cp.async.ca per chunk and row) on the
    // Llama-3-8B shapes (profiles/r2y_trie_probe_f{0,1}.json); only FEED 1 is instantiated
    launch_qmv_rows_f<NPG, STAGES, METHOD, BITS, MB, WARPS, 1>(cmd, p, grid, smem);
}
// (warps, stages): 16 warps x 2 stages (more warps to hide the latency of the ~1000-instruction stage) or 8 warps x 3 / 4 stages
template <int NPG, int METHOD, int BITS>
static void launch_qmv_rows_m(uzu_command_buffer* cmd, const QmvParams& p, uint32_t grid, size_t smem, int mb, int warps) {
    if (warps == 16) { if (mb == 1) launch_qmv_rows_s<NPG, 2, METHOD, BITS, 1, 16>(cmd, p, grid, smem); else launch_qmv_rows_s<NPG, 2, METHOD, BITS, 2, 16>(cmd, p, grid, smem); }
    else if (p.stages >= 4) { if (mb == 1) launch_qmv_rows_s<NPG, 4, METHOD, BITS, 1, 8>(cmd, p, grid, smem); else launch_qmv_rows_s<NPG, 4, METHOD, BITS, 2, 8>(cmd, p, grid, smem); }
    else { if (mb == 1) launch_qmv_rows_s<NPG, 3, METHOD, BITS, 1, 8>(cmd, p, grid, smem); else launch_qmv_rows_s<NPG, 3, METHOD, BITS, 2, 8>(cmd, p, grid, smem); }
}
template <int NPG>
static void launch_qmv_rows(uzu_command_buffer* cmd, const QmvParams& p, uint32_t grid, size_t smem, int mb, int warps) {
    if (p.bits == 4) {
        switch (p.method) {
            case UZU_QMETHOD_SCALE_ZERO_POINT: launch_qmv_rows_m<NPG, UZU_QMETHOD_SCALE_ZERO_POINT, 4>(cmd, p, grid, smem, mb, warps); break;
            case UZU_QMETHOD_SCALE_BIAS: launch_qmv_rows_m<NPG, UZU_QMETHOD_SCALE_BIAS, 4>(cmd, p, grid, smem, mb, warps); break;
            default: launch_qmv_rows_m<NPG, UZU_QMETHOD_SCALE_SYMMETRIC, 4>(cmd, p, grid, smem, mb, warps); break;
        }
    } else if constexpr (NPG == 128) {     // int8: 64-element groups = 128 nibbles (int8 gs32 would be NPG 64: streaming kernel)
        switch (p.method) {
            case UZU_QMETHOD_SCALE_ZERO_POINT: launch_qmv_rows_m<NPG, UZU_QMETHOD_SCALE_ZERO_POINT, 8>(cmd, p, grid, smem, mb, warps); break;
            case UZU_QMETHOD_SCALE_BIAS: launch_qmv_rows_m<NPG, UZU_QMETHOD_SCALE_BIAS, 8>(cmd, p, grid, smem, mb, warps); break;
            default: launch_qmv_rows_m<NPG, UZU_QMETHOD_SCALE_SYMMETRIC, 8>(cmd, p, grid, smem, mb, warps); break;
        }
    }
}

// Tuning overrides for sweeps (0 = heuristic). Not part of the reference-facing API; set through uzu_debug_set_qmv_tuning.
struct QmvTuning { int wpt = 0, dks = 0, per_sm = 0, stages = 0; };   // stages: 2..4 forces the cp.async ring depth
static QmvTuning g_tune;

template <int NPG, int MT>
static void launch_qmv_stream(uzu_command_buffer* cmd, const QmvParams& p, uint32_t grid, size_t smem) {
    static std::atomic<uint64_t> smem_opt_in{0};
    opt_in_dynamic_smem(cmd, qmv_stream_kernel<NPG, MT>, (int)(200 * 1024), smem_opt_in);
    launch(cmd, "qmv_stream_kernel", qmv_stream_kernel<NPG, MT>, dim3(grid), dim3(QS_WARPS * 32), smem, p);
}

template <int NPG>
static void launch_qmv_stream_mt(uzu_command_buffer* cmd, const QmvParams& p, uint32_t grid, size_t smem, int mt) {
    switch (mt) {
        case 1: launch_qmv_stream<NPG, 1>(cmd, p, grid, smem); break;
        case 2: launch_qmv_stream<NPG, 2>(cmd, p, grid, smem); break;
        default: launch_qmv_stream<NPG, 4>(cmd, p, grid, smem); break;
    }
}

template <int NPG>
static void launch_qmv_mt(uzu_command_buffer* cmd, const QmvParams& p, uint32_t tiles, int mt) {
    switch (mt) {
        case 1: launch_qmv<NPG, 1>(cmd, p, tiles); break;
        case 2: launch_qmv<NPG, 2>(cmd, p, tiles); break;
        default: launch_qmv<NPG, 4>(cmd, p, tiles); break;
    }
}

static void launch_generic(uzu_command_buffer* cmd, const uzu_matmul_args& a) {
    size_t total = (size_t)a.m * a.n;
    uint32_t blocks = (uint32_t)((total + 7) / 8);
    generic_kernel<<<blocks, 256, 0, cmd->ctx->stream>>>(a);
    after_launch(cmd, "matmul generic_kernel");
}

// returns true when the fused fast path (decode kernel) handled the call; with `fused` set and `dry_run`, only reports applicability
static bool encode_matmul(uzu_command_buffer* cmd, uzu_context* ctx_for_query, const uzu_matmul_args& a, const uzu_fused_linear_args* fused = nullptr, bool dry_run = false) {
    // prefill (m >= 64 tokens): tensor-core GEMM with an in-kernel dequant stage (prefill_gemm.cu)
    if (!fused && !dry_run && cmd && prefill_gemm_applicable(a)) {
        encode_prefill_gemm(cmd, a);
        return true;
    }
    const bool quant = a.b_prologue != UZU_B_FULL_PRECISION;
    const uint32_t bits = a.b_mode == UZU_QMODE_U4 ? 4 : 8;
    const uint32_t np = quant ? a.k * bits / 4 : 0;          // nibbles per row
    const uint32_t npg = quant ? a.b_group_size * bits / 4 : 0;
    const bool fast = quant && !a.gather_indices && a.input_dt == UZU_DT_BF16 && a.weights_dt == UZU_DT_BF16 &&
                      (np % 32 == 0) && (npg == 32 || npg == 64 || npg == 128 || npg == 256) &&
                      (a.k % a.b_group_size == 0 || true) && ((a.a & 15) == 0) && ((a.b & 15) == 0) && ((a.k * 2) % 16 == 0);
    if (!fast) {
        if (fused) return false;
        launch_generic(cmd, a);
        return true;
    }
    uzu_context* ctx = cmd ? cmd->ctx : ctx_for_query;
    if (fused && (a.m != 1 || (npg != 64 && npg != 128) || a.output_dt != UZU_DT_BF16 && a.output_dt != UZU_DT_F32)) return false;
    const int cpm = npg >= 128 ? 1 : 128 / npg;
    const int mpm = 8 / cpm;
    const uint32_t tiles = (a.n + 15) / 16;
    const uint32_t chunks_total = (np + 127) / 128;
    const uint32_t chunk_align = npg > 128 ? npg / 128 : 1;
    // rows kernel (2..16 activation rows per weight pass): int4 gs64 / gs128 and int8 gs64, whole super-chunks, aligned operands
    static const bool rows_off = [] { const char* v = getenv("UZU_QMV_ROWS"); return v && atoi(v) == 0; }();
    const bool rows_ok = !fused && !rows_off && a.m >= 2 && (npg == 128 || (npg == 64 && bits == 4)) && np % 512 == 0 && a.k % a.b_group_size == 0 &&
                         (a.k / a.b_group_size) % (512u / npg) == 0 && ((a.b_scales & 15) == 0) && ((a.b_zero_points & 15) == 0) &&
                         ((a.b_biases & 15) == 0) && (a.k * 2) % 16 == 0 && (np / 2) % 8 == 0;
    const uint32_t max_rows = rows_ok ? 16u : 4u * mpm;  // MT = 4

    for (uint32_t m0 = 0; m0 < a.m; m0 += max_rows) {
        const uint32_t mb = std::min(max_rows, a.m - m0);
        if (rows_ok && mb >= 2) {
            const int MBk = mb <= 8 ? 1 : 2;
            const uint32_t mrows_k = 8u * MBk, gps_k = 512u / npg;
            const uint32_t scs = chunks_total / QS_SC;                       // super-chunks per row
            // CTA shape: 16 warps x 2 ring stages or 8 warps x 3 / 4 stages (UZU_QMV_ROWS_WARPS = 8 | 16 overrides the default)
            static const int warps_env = [] { const char* v = getenv("UZU_QMV_ROWS_WARPS"); return v ? atoi(v) : 0; }();
            const uint32_t nw = warps_env == 8 ? 8u : (warps_env == 16 ? 16u : (uint32_t)UZU_QMV_ROWS_DEFAULT_WARPS);
            // shared memory: rings (warps x stages x 4608 B) + reduction buffer + the m x slice activations (B-fragment items + group sums)
            const size_t fixed = (size_t)nw * mrows_k * 16 * 4 + 256 + mrows_k * 4;
            const size_t per_sc = (size_t)mrows_k * (QS_SC * 16 * 16) + (size_t)mrows_k * gps_k * 4;     // per super-chunk of slice (+ one pad word per row, in `fixed`)
            const size_t budget = 225u * 1024u;
            uint32_t stages = nw == 16 ? 2u : 4u;
            auto dsc_max_for = [&](uint32_t st) {
                const size_t rings = (size_t)nw * st * QA_STAGE_BYTES;
                if (rings + fixed + mrows_k * 16 >= budget) return 0u;
                return (uint32_t)((budget - rings - fixed - mrows_k * 16) / per_sc);
            };
            uint32_t dsc_max = dsc_max_for(stages);
            if (nw == 8 && dsc_max < std::min(scs, 4u)) { stages = 3; dsc_max = dsc_max_for(3); }
            if (dsc_max >= 1) {
                // (k-slices, warps per tile): a CTA item = (nw / wpt) row tiles x one k-slice, every warp walks dsc / wpt super-chunks ("stages",
                // ~1000 instructions each) and then pays one epilogue. With so few stages per warp the launch is a handful of dependent
                // stage times long, so the choice minimises   rounds x (stages per warp + split-k epilogue + CTA barrier)   in stage units
                // (measured on B200: a split-k epilogue -- workspace write, fence, ticket -- ~0.6 stage, a CTA barrier + shared reduction ~0.5).
                const uint32_t ks_min = (scs + dsc_max - 1) / dsc_max;
                uint32_t ks = ks_min, dsc = (scs + ks_min - 1) / ks_min, wpt = 1;
                float best = 1e30f;
                for (uint32_t ks_c = ks_min; ks_c <= std::min(scs, ks_min * 4u); ++ks_c) {
                    const uint32_t dsc_c = (scs + ks_c - 1) / ks_c;
                    const uint32_t ks_r = (scs + dsc_c - 1) / dsc_c;
                    if (ks_r != ks_c) continue;                                   // same slice length as a smaller count
                    if (ks_c > 1 && ((size_t)tiles * ks_c * mrows_k * 16 * 4 > ctx->splitk_ws_bytes || tiles > ctx->splitk_counter_count)) break;
                    for (uint32_t w = 1; w <= 4u && w <= dsc_c; w *= 2) {
                        const uint32_t tpc_c = nw / w;
                        const uint32_t items_c = ((tiles + tpc_c - 1) / tpc_c) * ks_c;
                        const uint32_t amax_c = std::max(ks_c, ((uint32_t)ctx->sm_count / ks_c) * ks_c);
                        const uint32_t rounds_c = (items_c + amax_c - 1) / amax_c;
                        const float est = (float)rounds_c * ((float)((dsc_c + w - 1) / w) + (ks_c > 1 ? 0.6f : 0.1f) + (w > 1 ? 0.5f : 0.0f));
                        if (est < best) { best = est; ks = ks_c; dsc = dsc_c; wpt = w; }
                    }
                }
                const uint32_t tpc = nw / wpt;
                const uint32_t tgroups = (tiles + tpc - 1) / tpc;
                const size_t ws_need = (size_t)tiles * ks * mrows_k * 16 * 4;
                if ((ks == 1 || (ws_need <= ctx->splitk_ws_bytes && tiles <= ctx->splitk_counter_count))) {
                    QmvParams p{};
                    p.w = (const uint8_t*)a.b;
                    p.scales = (const __nv_bfloat16*)a.b_scales;
                    p.zero_points = (const uint8_t*)a.b_zero_points;
                    p.biases = (const __nv_bfloat16*)a.b_biases;
                    p.x = (const __nv_bfloat16*)a.a + (size_t)m0 * a.k;
                    p.d = (void*)(a.d + (size_t)m0 * a.n * (a.output_dt == UZU_DT_F32 ? 4 : 2));
                    p.bias = (a.d_transform & UZU_D_BIAS) ? (const __nv_bfloat16*)a.bias : nullptr;
                    p.ws = ctx->splitk_ws;
                    p.counters = ctx->splitk_counters;
                    p.m = mb; p.n = a.n; p.k = a.k;
                    p.np = np;
                    p.row_bytes = np / 2;
                    p.groups_per_row = a.k / a.b_group_size;
                    p.zp_stride = bits == 4 ? (p.groups_per_row + 1) / 2 : p.groups_per_row;
                    p.group_size = a.b_group_size;
                    p.chunks_total = chunks_total; p.chunks_per_slice = dsc * QS_SC; p.kslices = ks;
                    p.warps_per_tile = wpt;
                    p.stages = stages;
                    p.method = a.b_prologue == UZU_B_SCALE_BIAS_DEQUANT ? UZU_QMETHOD_SCALE_BIAS
                               : a.b_prologue == UZU_B_SCALE_ZERO_POINT_DEQUANT ? UZU_QMETHOD_SCALE_ZERO_POINT : UZU_QMETHOD_SCALE_SYMMETRIC;
                    p.bits = bits;
                    p.xor_mask = a.b_signed_codes ? (bits == 4 ? 0x88888888u : 0x80808080u) : 0u;
                    p.d_is_f32 = a.output_dt == UZU_DT_F32;
                    p.accumulate = (a.d_transform & UZU_D_ACCUMULATE) != 0;
                    p.has_soft_cap = (a.d_transform & UZU_D_SOFT_CAP) != 0;
                    p.ab_scale = (a.d_transform & UZU_D_SCALE) ? a.ab_scale : 1.0f;
                    p.soft_cap = a.soft_cap;
                    const uint32_t items = tgroups * ks;
                    // one CTA per SM; every CTA keeps one k-slice (grid multiple of ks) and gets (almost) the same number of items
                    const uint32_t amax = std::max(ks, ((uint32_t)ctx->sm_count / ks) * ks);
                    const uint32_t rounds = (items + amax - 1) / amax;
                    uint32_t grid = (items + rounds - 1) / rounds;
                    grid = std::min(amax, ((grid + ks - 1) / ks) * ks);
                    const size_t xs_bytes = (size_t)mrows_k * ((size_t)dsc * QS_SC * 16 + 1) * 16;
                    const size_t smem = xs_bytes + (size_t)mrows_k * ((dsc * gps_k) | 1u) * 4 + (size_t)nw * mrows_k * 16 * 4 + (size_t)nw * stages * QA_STAGE_BYTES;
                    if (smem <= 227u * 1024u) {
                        if (npg == 64) launch_qmv_rows<64>(cmd, p, grid, smem, MBk, (int)nw);
                        else launch_qmv_rows<128>(cmd, p, grid, smem, MBk, (int)nw);
                        continue;
                    }
                }
            }
        }
        int mt = (int)((mb + mpm - 1) / mpm);
        mt = mt <= 1 ? 1 : (mt == 2 ? 2 : 4);
        const uint32_t mrows = mt * mpm;
        // ---- streaming kernel (decode fast path) when the whole activation row set fits shared memory ----------
        static const bool force_tile = getenv("UZU_QMV_TILE") != nullptr;
        const uint32_t gps = 512u / npg;
        const size_t stream_smem = (size_t)mb * (((size_t)chunks_total * 16 + 31) & ~(size_t)31) * 16 +
                                   (size_t)mb * ((a.k + a.b_group_size - 1) / a.b_group_size) * 4 + 64 + 64;
        if (!force_tile && np % 512 == 0 && a.k % a.b_group_size == 0 && (a.k / a.b_group_size) % gps == 0 && stream_smem <= 150u * 1024u &&
            ((a.b_scales & 15) == 0) && ((a.b_zero_points & 15) == 0) && ((a.b_biases & 15) == 0)) {
            // items = (tile, k-slice); aim for >= 2 items per resident warp, slices of >= 2 super-chunks
            const uint32_t ctas_per_sm = stream_smem <= 40u * 1024u ? 4u : (stream_smem <= 70u * 1024u ? 3u : (stream_smem <= 100u * 1024u ? 2u : 1u));
            const uint32_t grid = (uint32_t)ctx->sm_count * ctas_per_sm;
            const uint32_t warps = grid * QS_WARPS;
            const uint32_t scs = chunks_total / QS_SC;            // super-chunks per row
            uint32_t want_slices = (2u * warps + tiles - 1) / tiles;
            uint32_t max_slices = std::max(1u, scs / 2u);
            uint32_t ks = std::max(1u, std::min(want_slices, max_slices));
            uint32_t sc_per_slice = (scs + ks - 1) / ks;
            ks = (scs + sc_per_slice - 1) / sc_per_slice;
            while (ks > 1 && ((size_t)tiles * ks * 16 * mrows * 4 > ctx->splitk_ws_bytes || tiles > ctx->splitk_counter_count)) {
                sc_per_slice *= 2;
                ks = (scs + sc_per_slice - 1) / sc_per_slice;
            }
            QmvParams p{};
            p.w = (const uint8_t*)a.b;
            p.scales = (const __nv_bfloat16*)a.b_scales;
            p.zero_points = (const uint8_t*)a.b_zero_points;
            p.biases = (const __nv_bfloat16*)a.b_biases;
            p.x = (const __nv_bfloat16*)a.a + (size_t)m0 * a.k;
            p.d = (void*)(a.d + (size_t)m0 * a.n * (a.output_dt == UZU_DT_F32 ? 4 : 2));
            p.bias = (a.d_transform & UZU_D_BIAS) ? (const __nv_bfloat16*)a.bias : nullptr;
            p.ws = ctx->splitk_ws;
            p.counters = ctx->splitk_counters;
            p.m = mb; p.n = a.n; p.k = a.k;
            p.np = np;
            p.row_bytes = np / 2;
            p.groups_per_row = a.k / a.b_group_size;
            p.zp_stride = bits == 4 ? (p.groups_per_row + 1) / 2 : p.groups_per_row;
            p.group_size = a.b_group_size;
            p.chunks_total = chunks_total; p.chunks_per_slice = sc_per_slice * QS_SC; p.kslices = ks;
            p.method = a.b_prologue == UZU_B_SCALE_BIAS_DEQUANT ? UZU_QMETHOD_SCALE_BIAS
                       : a.b_prologue == UZU_B_SCALE_ZERO_POINT_DEQUANT ? UZU_QMETHOD_SCALE_ZERO_POINT : UZU_QMETHOD_SCALE_SYMMETRIC;
            p.bits = bits;
            p.xor_mask = a.b_signed_codes ? (bits == 4 ? 0x88888888u : 0x80808080u) : 0u;
            p.d_is_f32 = a.output_dt == UZU_DT_F32;
            p.accumulate = (a.d_transform & UZU_D_ACCUMULATE) != 0;
            p.has_soft_cap = (a.d_transform & UZU_D_SOFT_CAP) != 0;
            p.ab_scale = (a.d_transform & UZU_D_SCALE) ? a.ab_scale : 1.0f;
            p.soft_cap = a.soft_cap;
            const uint32_t items = tiles * ks;
            const uint32_t use_grid = std::min(grid, std::max(1u, (items + QS_WARPS - 1) / QS_WARPS));
            static const bool no_decode_kernel = getenv("UZU_QMV_NO_DECODE") != nullptr;
            // the m == 1 decode kernel needs the activation row + its rings in shared memory; longer rows take the streaming kernel below
            const size_t decode_smem = stream_smem + 2 * QS_WARPS * 16 * 4 + 64 + (size_t)QS_WARPS * QA_STAGES * QA_STAGE_BYTES;
            if (mb == 1 && (npg == 64 || npg == 128) && !no_decode_kernel && decode_smem <= 200u * 1024u) {
                // CTA items: TPC tiles x one global k-slice; k split over the CTA's warps first, over CTAs only for few-row shapes
                uint32_t dks = 1;
                {
                    const uint32_t wpt0 = scs >= 4 ? 4u : (scs >= 2 ? 2u : 1u);
                    const uint32_t tg0 = (tiles + (4 / wpt0) - 1) / (4 / wpt0);
                    if (tg0 < (uint32_t)ctx->sm_count && scs >= 8)
                        dks = std::max(1u, std::min((2u * (uint32_t)ctx->sm_count + tg0 - 1) / tg0, scs / 4u));
                }
                uint32_t dsc = (scs + dks - 1) / dks;
                dks = (scs + dsc - 1) / dsc;
                while (dks > 1 && ((size_t)tiles * dks * 16 * 4 > ctx->splitk_ws_bytes || tiles > ctx->splitk_counter_count)) {
                    dsc *= 2;
                    dks = (scs + dsc - 1) / dsc;
                }
                const bool want_paired = fused && fused->epilogue == 1;
                if (g_tune.dks > 0 && !want_paired) {
                    dks = std::min((uint32_t)g_tune.dks, scs);
                    dsc = (scs + dks - 1) / dks;
                    dks = (scs + dsc - 1) / dsc;
                    if ((size_t)tiles * dks * 16 * 4 > ctx->splitk_ws_bytes || tiles > ctx->splitk_counter_count) { dks = 1; dsc = scs; }
                }
                uint32_t wpt = dsc >= 4 ? 4u : (dsc >= 2 ? 2u : 1u);
                // plenty of row tiles: give every warp whole rows (no cross-warp reduction, no CTA sync); else split k in the CTA
                const uint32_t resident_warps = 4u * 3u * (uint32_t)ctx->sm_count;
                if (tiles >= 2u * resident_warps) wpt = 1;
                else if (tiles * 2u >= 2u * resident_warps && wpt > 2) wpt = 2;
                if (g_tune.wpt > 0) wpt = std::min((uint32_t)g_tune.wpt, dsc >= 4 ? 4u : (dsc >= 2 ? 2u : 1u));
                const uint32_t per_sm_cap = g_tune.per_sm > 0 ? (uint32_t)g_tune.per_sm : 4u;
                // Ring depth: 2 stages let 4 CTAs share an SM (best when the grid fills the machine); when the grid is smaller than
                // what deeper rings still leave room for, the spare shared memory buys a longer prefetch per warp instead.
                auto pick_stages = [&](size_t smem2, uint32_t grid2, size_t& smem_out) {
                    int st = 2;
                    smem_out = smem2;
                    if (g_tune.stages >= 2 && g_tune.stages <= 4) {
                        st = g_tune.stages;
                    } else {
                        for (int cand = 4; cand > 2; --cand) {
                            const size_t sm = smem2 + (size_t)QS_WARPS * (cand - 2) * QA_STAGE_BYTES;
                            if (sm > 200u * 1024u) continue;
                            const uint32_t per = std::min(per_sm_cap, (uint32_t)((220u * 1024u) / (sm + 1024u)));
                            if (per >= 1 && per * (uint32_t)ctx->sm_count >= grid2) { st = cand; break; }
                        }
                    }
                    smem_out = smem2 + (size_t)QS_WARPS * (st - 2) * QA_STAGE_BYTES;
                    if (smem_out > 200u * 1024u) { st = 2; smem_out = smem2; }
                    return st;
                };
                const uint32_t tgroups = (tiles + (4 / wpt) - 1) / (4 / wpt);
                p.chunks_per_slice = dsc * QS_SC;
                p.kslices = dks;
                p.warps_per_tile = wpt;
                const size_t dsmem = stream_smem + 2 * QS_WARPS * 16 * 4 + 64;
                const size_t asmem = dsmem + (size_t)QS_WARPS * QA_STAGES * QA_STAGE_BYTES;
                const size_t fsmem = asmem + (fused && fused->prologue ? (size_t)a.k * 2 + 64 : 0);
                if (fused) {
                    if (fsmem > 200u * 1024u || (a.k % 8) != 0 || bits != 4) return false;
                    // decode-stream copy available: one TMA bulk copy per ring stage (zero-point / symmetric weights, unsigned codes)
                    static const bool tma_off = [] { const char* v = getenv("UZU_QMV_TMA"); return v && atoi(v) == 0; }();
                    if (fused->decode_stream && !tma_off && p.xor_mask == 0 && p.method != UZU_QMETHOD_SCALE_BIAS && (npg == 64 || npg == 128)) {
                        p.stream = (const uint8_t*)fused->decode_stream;
                        p.stream_C = (np + 511u) / 512u;
                        p.method = uzu::QMV_TMA;
                    }
                    uint32_t pair_groups = 0;
                    if (fused->epilogue == 1) {
                        // gated-act epilogue: rows [0,F) up, [F,2F) gate; every CTA owns whole pairs, so no global k split
                        if ((a.n & 1u) || ((a.n / 2) % 16u) != 0 || a.output_dt != UZU_DT_BF16 || (a.d_transform & (UZU_D_ACCUMULATE | UZU_D_SOFT_CAP))) return false;
                        dks = 1; dsc = scs;
                        // both halves of a pair meet in one CTA: at most 2 warps per tile (TPC >= 2)
                        wpt = dsc >= 2 ? 2u : 1u;
                        if (tiles >= 2u * resident_warps) wpt = 1;
                        if (g_tune.wpt > 0) wpt = std::min((uint32_t)g_tune.wpt, dsc >= 2 ? 2u : 1u);
                        p.chunks_per_slice = dsc * QS_SC; p.kslices = 1; p.warps_per_tile = wpt;
                        p.epi_gated = 1; p.epi_act = fused->act_type; p.pair_tiles = a.n / 32u;
                        pair_groups = (p.pair_tiles + (2 / wpt) - 1) / (2 / wpt);
                    } else if (fused->epilogue != 0) return false;
                    if (dry_run) return true;
                    p.prologue = fused->prologue;
                    if (fused->prologue == 0) p.x = (const __nv_bfloat16*)a.a;
                    if (fused->prologue == 1) {
                        p.pro_a = (const __nv_bfloat16*)fused->norm_input;
                        p.pro_b = (const __nv_bfloat16*)fused->norm_shortcut_in;
                        p.pro_shortcut_out = (__nv_bfloat16*)fused->shortcut_out;
                        p.pro_scales = (const float*)fused->norm_scales;
                        p.pro_eps = fused->norm_epsilon; p.pro_scale_offset = fused->norm_scale_offset;
                        p.pro_residual_add = fused->norm_residual_add; p.pro_full_layer = fused->norm_full_layer;
                    } else if (fused->prologue == 3) {
                        p.pro_a = (const __nv_bfloat16*)fused->sg_attn;
                        p.pro_b = (const __nv_bfloat16*)fused->sg_gate;
                    }
                    const uint32_t per_sm = std::max(1u, std::min(per_sm_cap, (uint32_t)((220u * 1024u) / (fsmem + 1024u))));
                    const uint32_t aitems = pair_groups ? pair_groups : std::max(1u, tgroups * dks), amax = per_sm * (uint32_t)ctx->sm_count;
                    const uint32_t rounds = (aitems + amax - 1) / amax;
                    const uint32_t agrid = (aitems + rounds - 1) / rounds;
                    size_t lsmem = fsmem;
                    p.stages = pick_stages(fsmem, agrid, lsmem);
                    if (npg == 64) launch_qmv_decode_async<64>(cmd, p, agrid, lsmem);
                    else launch_qmv_decode_async<128>(cmd, p, agrid, lsmem);
                    return true;
                }
                {
                    const uint32_t per_sm = std::max(1u, std::min(per_sm_cap, (uint32_t)((220u * 1024u) / (asmem + 1024u))));
                    const uint32_t aitems = std::max(1u, tgroups * dks), amax = per_sm * (uint32_t)ctx->sm_count;
                    const uint32_t rounds = (aitems + amax - 1) / amax;
                    const uint32_t agrid = (aitems + rounds - 1) / rounds;     // same number of items for (almost) every CTA
                    size_t lsmem = asmem;
                    p.stages = pick_stages(asmem, agrid, lsmem);
                    if (npg == 64) launch_qmv_decode_async<64>(cmd, p, agrid, lsmem);
                    else launch_qmv_decode_async<128>(cmd, p, agrid, lsmem);
                    continue;
                }
            }
            if (fused) return false;
            switch (npg) {
                case 32: launch_qmv_stream_mt<32>(cmd, p, use_grid, stream_smem, mt); break;
                case 64: launch_qmv_stream_mt<64>(cmd, p, use_grid, stream_smem, mt); break;
                case 128: launch_qmv_stream_mt<128>(cmd, p, use_grid, stream_smem, mt); break;
                default: launch_qmv_stream_mt<256>(cmd, p, use_grid, stream_smem, mt); break;
            }
            continue;
        }
        if (fused) return false;
        // k slicing: enough CTAs to fill the machine, >= 1 chunk per warp, activations must fit shared memory
        uint32_t ks = 1;
        const uint32_t target = 4u * (uint32_t)ctx->sm_count;
        if (tiles < target) ks = (target + tiles - 1) / tiles;
        uint32_t max_ks = std::max(1u, chunks_total / QMV_WARPS);
        ks = std::min(ks, max_ks);
        uint32_t cps = (chunks_total + ks - 1) / ks;
        const uint32_t smem_cap_chunks = std::max(1u, (150u * 1024u) / (mb * 256u + mb * 16u + 64u));
        cps = std::min(cps, smem_cap_chunks);
        cps = (cps + chunk_align - 1) / chunk_align * chunk_align;
        ks = (chunks_total + cps - 1) / cps;
        // workspace limits
        while (ks > 1 && ((size_t)tiles * ks * 16 * mrows * 4 > ctx->splitk_ws_bytes || tiles > ctx->splitk_counter_count)) {
            cps *= 2;
            ks = (chunks_total + cps - 1) / cps;
        }
        const size_t smem_need = qmv_smem_bytes(mb, cps, (cps * 128u + npg - 1) / npg + 1, mt);
        if (smem_need > 200u * 1024u) {  // cannot satisfy both limits: generic fallback
            uzu_matmul_args b = a;
            b.a = a.a + (size_t)m0 * a.k * 2;
            b.d = a.d + (size_t)m0 * a.n * (a.output_dt == UZU_DT_F32 ? 4 : 2);
            b.m = mb;
            launch_generic(cmd, b);
            continue;
        }
        QmvParams p{};
        p.w = (const uint8_t*)a.b;
        p.scales = (const __nv_bfloat16*)a.b_scales;
        p.zero_points = (const uint8_t*)a.b_zero_points;
        p.biases = (const __nv_bfloat16*)a.b_biases;
        p.x = (const __nv_bfloat16*)a.a + (size_t)m0 * a.k;
        p.d = (void*)(a.d + (size_t)m0 * a.n * (a.output_dt == UZU_DT_F32 ? 4 : 2));
        p.bias = (a.d_transform & UZU_D_BIAS) ? (const __nv_bfloat16*)a.bias : nullptr;
        p.ws = ctx->splitk_ws;
        p.counters = ctx->splitk_counters;
        p.m = mb; p.n = a.n; p.k = a.k;
        p.np = np;
        p.row_bytes = np / 2;
        p.groups_per_row = (a.k + a.b_group_size - 1) / a.b_group_size;
        p.zp_stride = bits == 4 ? (p.groups_per_row + 1) / 2 : p.groups_per_row;
        p.group_size = a.b_group_size;
        p.chunks_total = chunks_total; p.chunks_per_slice = cps; p.kslices = ks;
        p.method = a.b_prologue == UZU_B_SCALE_BIAS_DEQUANT ? UZU_QMETHOD_SCALE_BIAS
                   : a.b_prologue == UZU_B_SCALE_ZERO_POINT_DEQUANT ? UZU_QMETHOD_SCALE_ZERO_POINT : UZU_QMETHOD_SCALE_SYMMETRIC;
        p.bits = bits;
        p.xor_mask = a.b_signed_codes ? (bits == 4 ? 0x88888888u : 0x80808080u) : 0u;
        p.d_is_f32 = a.output_dt == UZU_DT_F32;
        p.accumulate = (a.d_transform & UZU_D_ACCUMULATE) != 0;
        p.has_soft_cap = (a.d_transform & UZU_D_SOFT_CAP) != 0;
        p.ab_scale = (a.d_transform & UZU_D_SCALE) ? a.ab_scale : 1.0f;
        p.soft_cap = a.soft_cap;
        switch (npg) {
            case 32: launch_qmv_mt<32>(cmd, p, tiles, mt); break;
            case 64: launch_qmv_mt<64>(cmd, p, tiles, mt); break;
            case 128: launch_qmv_mt<128>(cmd, p, tiles, mt); break;
            default: launch_qmv_mt<256>(cmd, p, tiles, mt); break;
        }
    }
    return true;
}

}  // namespace uzu

extern "C" {

uzu_status uzu_matmul_validate(const uzu_matmul_args* args) {
    const char* err = uzu::validate(args);
    if (err) return uzu::fail(UZU_ERROR_INVALID_ARGUMENT, std::string("matmul: ") + err);
    return UZU_OK;
}

void uzu_matmul_encode(uzu_command_buffer* cmd, const uzu_matmul_args* args) {
    if (!uzu::encodable(cmd, "matmul")) return;
    const char* err = uzu::validate(args);
    if (err) {
        cmd->record_error(UZU_ERROR_INVALID_ARGUMENT, std::string("matmul: ") + err);
        return;
    }
    if (args->d_transform & UZU_D_RHT) {
        // MatmulDOps::rht_factors (cpu/kernel/matmul/kernel.rs:64,162,285,297-303): the product (scale / accumulate / soft-cap epilogue, NO
        // bias) lands in D, ActivationTransform(OutputRht) runs over D in place, then TensorAddBias adds the bias (bias_after_rht).
        uzu_matmul_args inner = *args;
        inner.d_transform &= ~(uint32_t)(UZU_D_RHT | UZU_D_BIAS);
        inner.bias = 0;
        inner.rht_factors = 0;
        uzu::encode_matmul(cmd, cmd->ctx, inner);
        uzu_activation_transform_args t{};
        t.fp_out = args->d; t.rht_factors = args->rht_factors;
        t.batch_size = args->m; t.element_count = args->n;
        t.ops = UZU_ACTIVATION_TRANSFORM_OUTPUT_RHT; t.in_place = 1; t.data_type = args->output_dt;
        uzu_activation_transform_encode(cmd, &t);
        if ((args->d_transform & UZU_D_BIAS) && args->bias) {
            if (args->output_dt != UZU_DT_BF16 || args->weights_dt != UZU_DT_BF16) {
                cmd->record_error(UZU_ERROR_UNSUPPORTED, "matmul: bias after output RHT is bf16-only (TensorAddBias)");
                return;
            }
            uzu_tensor_add_bias_encode(cmd, 0, args->bias, args->d, args->n, args->m * args->n);
        }
        return;
    }
    uzu::encode_matmul(cmd, cmd->ctx, *args);
}

void uzu_debug_set_qmv_tuning(int warps_per_tile, int k_slices, int ctas_per_sm, int stages) {
    uzu::g_tune.wpt = warps_per_tile; uzu::g_tune.dks = k_slices; uzu::g_tune.per_sm = ctas_per_sm; uzu::g_tune.stages = stages;
}

int uzu_fused_linear_supported(uzu_context* ctx, const uzu_fused_linear_args* args) {
    if (!ctx || !args || uzu::validate(&args->matmul)) {
        // validate() wants matmul.a non-null; the fused path ignores it, so tolerate a == 0 here
        if (!ctx || !args) return 0;
        uzu_matmul_args m = args->matmul;
        if (!m.a) m.a = m.d;
        if (uzu::validate(&m)) return 0;
    }
    // prologue 0 + epilogue 0 = a plain GEMV: only meaningful through this entry point when it brings a decode stream (TMA-fed rings)
    if (args->matmul.d_transform & UZU_D_RHT) return 0;   // the output transform needs the whole row first: unfused sequence
    if (args->prologue > 3 || args->prologue == 2 || args->epilogue > 1 || (args->prologue == 0 && args->epilogue == 0 && !args->decode_stream)) return 0;
    if (args->prologue == 3 && args->epilogue) return 0;
    if (args->prologue == 0 && !args->matmul.a) return 0;
    if (args->prologue == 1 && (!args->norm_input || !args->norm_scales || (args->norm_residual_add && !args->norm_shortcut_in) ||
                                (args->shortcut_out && args->shortcut_out == args->norm_shortcut_in))) return 0;
    if (args->prologue == 3 && (!args->sg_attn || !args->sg_gate)) return 0;
    uzu_matmul_args m = args->matmul;
    if (args->prologue != 0) m.a = 0;
    return uzu::encode_matmul(nullptr, ctx, m, args, true) ? 1 : 0;
}

void uzu_fused_linear_encode(uzu_command_buffer* cmd, const uzu_fused_linear_args* args) {
    if (!uzu::encodable(cmd, "fused_linear")) return;
    if (!uzu_fused_linear_supported(cmd->ctx, args)) {
        cmd->record_error(UZU_ERROR_UNSUPPORTED, "fused_linear: shape / format not covered by the fused decode kernel (encode the unfused sequence)");
        return;
    }
    uzu_matmul_args m = args->matmul;
    if (args->prologue != 0) m.a = 0;
    uzu::encode_matmul(cmd, cmd->ctx, m, args, false);
}

}  // extern "C"
