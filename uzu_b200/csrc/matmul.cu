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
    uint32_t method;           // uzu_quantization_method
    uint32_t bits;
    uint32_t xor_mask;         // signed_codes
    uint32_t d_is_f32;
    uint32_t accumulate, has_soft_cap;
    float ab_scale, soft_cap;
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
    if (a->d_transform & UZU_D_RHT) return "output RHT (Mirai HybridSpec) is not supported by this backend";
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
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(qmv_kernel<NPG, MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        attr_set = true;
    }
    qmv_kernel<NPG, MT><<<tiles * p.kslices, QMV_WARPS * 32, smem, cmd->ctx->stream>>>(p);
    after_launch(cmd, "qmv_kernel");
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

static void encode_matmul(uzu_command_buffer* cmd, const uzu_matmul_args& a) {
    const bool quant = a.b_prologue != UZU_B_FULL_PRECISION;
    const uint32_t bits = a.b_mode == UZU_QMODE_U4 ? 4 : 8;
    const uint32_t np = quant ? a.k * bits / 4 : 0;          // nibbles per row
    const uint32_t npg = quant ? a.b_group_size * bits / 4 : 0;
    const bool fast = quant && !a.gather_indices && a.input_dt == UZU_DT_BF16 && a.weights_dt == UZU_DT_BF16 &&
                      (np % 32 == 0) && (npg == 32 || npg == 64 || npg == 128 || npg == 256) &&
                      (a.k % a.b_group_size == 0 || true) && ((a.a & 15) == 0) && ((a.b & 15) == 0) && ((a.k * 2) % 16 == 0);
    if (!fast) {
        launch_generic(cmd, a);
        return;
    }
    uzu_context* ctx = cmd->ctx;
    const int cpm = npg >= 128 ? 1 : 128 / npg;
    const int mpm = 8 / cpm;
    const uint32_t max_rows = 4 * mpm;  // MT = 4
    const uint32_t tiles = (a.n + 15) / 16;
    const uint32_t chunks_total = (np + 127) / 128;
    const uint32_t chunk_align = npg > 128 ? npg / 128 : 1;

    for (uint32_t m0 = 0; m0 < a.m; m0 += max_rows) {
        const uint32_t mb = std::min(max_rows, a.m - m0);
        int mt = (int)((mb + mpm - 1) / mpm);
        mt = mt <= 1 ? 1 : (mt == 2 ? 2 : 4);
        const uint32_t mrows = mt * mpm;
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
    uzu::encode_matmul(cmd, *args);
}

}  // extern "C"
