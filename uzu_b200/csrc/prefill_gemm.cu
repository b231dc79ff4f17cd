// Prefill MatmulKernel (m >= 64 tokens) for sm_100a: D[m,n] = epilogue(A[m,k] * dequant(B[n,k])^T) on the 5th-generation
// tensor cores (tcgen05.mma, accumulators in tensor memory), fed by an in-kernel dequantisation stage.
//
// Semantic spec: backends/cpu/kernel/matmul/kernel.rs:164-295 (GEMM branch of MatmulKernel; caller
// encodable_block/linear/matmul.rs:122-148, m = prefill chunk <= 1024). SURVEY 8, north_star row "prefill batched-GEMM".
//
// Shape of the kernel (one CTA = 128 output features x (MT x 128) tokens, K walked in blocks of 64):
//   * warps 0-3 (128 threads) are the PRODUCERS. Per K block they
//       - copy the activation tile (bf16, K contiguous = "K-major") into shared memory with 16 B loads/stores, and
//       - dequantise one weight row each: 64 codes -> f32 `scale*code + corr` (exactly the reference's expression; the product
//         is exact, so the FMA equals the reference's separate multiply and add) -> split into TWO bf16 planes
//         hi = bf16(w), lo = bf16(w - hi). For ZeroPoint / Symmetric weights w = scale*(code - zp) has <= 16 significant bits, so
//         hi + lo == w EXACTLY (8 + 8 bits); for MLX scale/bias the residual is < 2^-17 |w|. Feeding a single rounded bf16 plane
//         would put a 2^-9 relative error on every weight (about 1e-3 of the output rms, i.e. the whole parity budget).
//     Tiles are written in the UMMA canonical K-major SWIZZLE_128B layout (8-row x 128-byte atoms, 16-byte chunk index XORed
//     with the row index inside the atom), made visible to the async proxy with fence.proxy.async, and handed to the MMA warp
//     through an mbarrier ("full", 128 arrivals).
//   * warp 4, one lane, is the MMA ISSUER: per K block 4 (K=16 steps) x MT x 2 (hi, lo plane) tcgen05.mma.cta_group::1.kind::f16
//     128x128x16 instructions accumulate into MT x 128 TMEM columns; bf16 x bf16 products are exact in the f32 accumulator, so
//     the result is sum_k x_k*w_k with f32 accumulation, the reference's arithmetic up to summation order. tcgen05.commit
//     releases the shared-memory stage ("empty" mbarrier) and, after the last block, signals the epilogue.
//   * warps 0-3 then run the EPILOGUE: tcgen05.ld (32 lanes x 32 columns per instruction; warp w owns TMEM lanes 32w..32w+31 =
//     token rows), ab_scale / accumulate / bias / soft-cap in the reference's order, bf16 (RNE) or f32 store.
// The weight tile is dequantised once per 128 x MT tokens; with MT = 2 the tensor pipe (2 planes x 2 sub-tiles) and the
// dequantising ALU work per K block are about balanced (~1000 vs ~450 issue cycles), so the kernel is tensor-bound by design.
//
// Bound: tensor pipe (2*m*n*k flop useful, 2x that issued because of the hi/lo planes). Algorithmic HBM bytes per launch are the
// packed weights once per token tile + activations once per feature tile (both L2 resident for the shapes of SURVEY 8).
#include <algorithm>
#include <cstdlib>

#include "common.cuh"

namespace uzu {

struct QmmParams {
    const uint8_t* w;
    const __nv_bfloat16* scales;
    const uint8_t* zero_points;
    const __nv_bfloat16* biases;
    const __nv_bfloat16* x;    // [m, k]
    void* d;                   // [m, n]
    const __nv_bfloat16* bias; // [n] or null
    uint32_t m, n, k;
    uint32_t row_bytes, groups_per_row, zp_stride, group_size;
    uint32_t method, xor_mask, d_is_f32, accumulate, has_soft_cap;
    float ab_scale, soft_cap;
    // UMMA encodings, chosen by the host so that tools/umma_probe.py can sweep alternatives without recompiling
    uint32_t layout;   // 0 = SWIZZLE_128B K-major atoms, 1 = SWIZZLE_NONE core matrices (8 rows x 16 B)
    uint32_t desc_hi;  // upper 32 bits of the shared-memory matrix descriptor (SBO, version, layout type)
    uint32_t desc_lbo; // leading-dimension byte offset field (>> 4)
    uint32_t k_step;   // start-address increment (>> 4) per UMMA_K = 16 elements
    uint32_t idesc;    // instruction descriptor (kind::f16, bf16 x bf16 -> f32, M = 128, N = 128, both K-major)
};

constexpr int QMM_PRODUCERS = 128;
constexpr int QMM_THREADS = 160;
constexpr uint32_t QMM_TILE_BYTES = 128 * 64 * 2;  // one 128-row x 64-k bf16 tile

// ---- PTX wrappers ----------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("{ .reg .b64 st; mbarrier.arrive.shared::cta.b64 st, [%0]; }" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// Spins on the barrier phase; a protocol error (wrong descriptor -> MMA never commits) must not hang the GPU, so after ~2 s of
// waiting the kernel traps and the launch surfaces as a CUDA error on the command buffer.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    uint64_t t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if ((++spins & 1023u) == 0) {
            uint64_t t1;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t1 - t0 > 2000000000ull) __trap();
        }
    }
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t slot_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]^T, issued by ONE thread for the CTA
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p; }"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// mbarrier arrive once every tcgen05.mma issued so far by this thread has completed (implies fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// 32 TMEM lanes (one per thread of the warp) x 32 consecutive 32-bit columns
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, "
        "[%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// byte offset of the 16-byte chunk `ch` (8 bf16 along k) of row `row` inside a 128-row x 64-k tile
__device__ __forceinline__ uint32_t tile_off(uint32_t row, uint32_t ch, uint32_t layout) {
    return layout == 0 ? row * 128u + ((ch ^ (row & 7u)) << 4)                     // SWIZZLE_128B: Swizzle<3,4,3> on byte addresses
                       : (row >> 3) * 1024u + ch * 128u + (row & 7u) * 16u;         // core matrices: 8 rows x 16 B contiguous
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo_elem, float hi_elem) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo_elem, hi_elem);   // .x (low 16 bits) = first element in memory
    return *reinterpret_cast<uint32_t*>(&v);
}

// Two weights -> packed hi plane and lo plane words. w = fmaf(s, code, c): s*code is exact (<= 16 significant bits), so this is the
// reference's `scale * code + corr` (kernel.rs:267-276) bit for bit.
__device__ __forceinline__ void dequant_pair(uint32_t code0, uint32_t code1, float s, float c, uint32_t& hi, uint32_t& lo) {
    const float f0 = __uint_as_float(0x4B000000u | code0) - 8388608.0f;   // exact int -> float
    const float f1 = __uint_as_float(0x4B000000u | code1) - 8388608.0f;
    const float w0 = fmaf(s, f0, c), w1 = fmaf(s, f1, c);
    hi = pack_bf16x2(w0, w1);
    const float h0 = __uint_as_float(hi << 16), h1 = __uint_as_float(hi & 0xffff0000u);
    lo = pack_bf16x2(w0 - h0, w1 - h1);
}

template <int BITS, int MT, int STAGES>
__global__ void __launch_bounds__(QMM_THREADS, 1) qmm_umma_kernel(const QmmParams p) {
    constexpr uint32_t A_BYTES = QMM_TILE_BYTES * MT;
    constexpr uint32_t STAGE_BYTES = A_BYTES + 2 * QMM_TILE_BYTES;
    constexpr uint32_t TMEM_COLS = 128 * MT;
    constexpr int WORDS = BITS == 4 ? 8 : 16;   // 32-bit words of packed codes per row per K block (64 weights)

    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[2 * STAGES + 1];
    __shared__ uint32_t tmem_slot;

    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint32_t smem_base = smem_u32(smem_raw);
    const uint32_t pad = (1024u - (smem_base & 1023u)) & 1023u;    // SWIZZLE_128B atoms must sit on 1024-byte boundaries
    uint8_t* smem = smem_raw + pad;
    smem_base += pad;
    const uint32_t bar0 = smem_u32(bars);
    auto full_bar = [&](uint32_t s) { return bar0 + 8u * s; };
    auto empty_bar = [&](uint32_t s) { return bar0 + 8u * (STAGES + s); };
    const uint32_t accum_bar = bar0 + 8u * (2 * STAGES);

    if (warp == 0) {
        if (lane == 0) {
            for (int s = 0; s < STAGES; ++s) {
                mbar_init(full_bar(s), QMM_PRODUCERS);
                mbar_init(empty_bar(s), 1);
            }
            mbar_init(accum_bar, 1);
            fence_mbar_init();
        }
        __syncwarp();
        tmem_alloc(smem_u32(&tmem_slot), TMEM_COLS);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;

    const uint32_t n0 = blockIdx.x * 128u, m0 = blockIdx.y * (128u * MT);
    const uint32_t nkb = p.k >> 6;
    const uint32_t layout = p.layout;

    if (warp < 4) {
        // ================================ producers ================================
        const uint32_t brow = min(n0 + tid, p.n - 1);      // rows past n are computed from a clamped row and never stored
        const uint8_t* wrow = p.w + (size_t)brow * p.row_bytes;
        const __nv_bfloat16* srow = p.scales + (size_t)brow * p.groups_per_row;
        const uint8_t* zrow = p.zero_points ? p.zero_points + (size_t)brow * p.zp_stride : nullptr;
        const __nv_bfloat16* crow = p.biases ? p.biases + (size_t)brow * p.groups_per_row : nullptr;
        pdl_wait();                                         // activations come from the previous kernel
        for (uint32_t kb = 0; kb < nkb; ++kb) {
            const uint32_t s = kb % STAGES, it = kb / STAGES;
            // ---- global loads first (in flight while we wait for the stage to drain) ----
            uint4 av[MT * 8];
#pragma unroll
            for (int i = 0; i < MT * 8; ++i) {
                const uint32_t c = (uint32_t)i * QMM_PRODUCERS + tid, row = c >> 3, ch = c & 7u;
                const uint32_t grow = m0 + row;
                av[i] = make_uint4(0u, 0u, 0u, 0u);
                if (grow < p.m) av[i] = *reinterpret_cast<const uint4*>(p.x + (size_t)grow * p.k + (size_t)kb * 64u + ch * 8u);
            }
            uint32_t q[WORDS];
#pragma unroll
            for (int i = 0; i < WORDS / 4; ++i) {
                const uint4 t = ldg_stream_u4(wrow + (size_t)kb * (WORDS * 4) + i * 16);
                q[4 * i] = t.x ^ p.xor_mask; q[4 * i + 1] = t.y ^ p.xor_mask; q[4 * i + 2] = t.z ^ p.xor_mask; q[4 * i + 3] = t.w ^ p.xor_mask;
            }
            float sc[2], cc[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const uint32_t g = (kb * 64u + (uint32_t)h * 32u) / p.group_size;
                const float s_ = bf2f(srow[g]);
                float c_;
                if (p.method == UZU_QMETHOD_SCALE_ZERO_POINT) {
                    uint32_t z;
                    if (BITS == 4) {
                        const uint32_t zb = zrow[g >> 1];
                        z = (g & 1u) ? (zb >> 4) : (zb & 15u);
                    } else {
                        z = zrow[g];
                    }
                    c_ = -s_ * (float)z;
                } else if (p.method == UZU_QMETHOD_SCALE_BIAS) {
                    c_ = bf2f(crow[g]);
                } else {
                    c_ = -s_ * (float)(1u << (BITS - 1));
                }
                sc[h] = s_;
                cc[h] = c_;
            }
            // ---- wait for the MMA warp to release the stage, then fill it ----
            mbar_wait(empty_bar(s), (it & 1u) ^ 1u);
            uint8_t* st = smem + (size_t)s * STAGE_BYTES;
#pragma unroll
            for (int i = 0; i < MT * 8; ++i) {
                const uint32_t c = (uint32_t)i * QMM_PRODUCERS + tid, row = c >> 3, ch = c & 7u;
                *reinterpret_cast<uint4*>(st + (row >> 7) * QMM_TILE_BYTES + tile_off(row & 127u, ch, layout)) = av[i];
            }
            uint8_t* bhi = st + A_BYTES;
            uint8_t* blo = bhi + QMM_TILE_BYTES;
#pragma unroll
            for (int ch = 0; ch < 8; ++ch) {               // chunk = 8 consecutive k
                const float s_ = sc[ch >> 2], c_ = cc[ch >> 2];
                uint4 hi4, lo4;
                if constexpr (BITS == 4) {
                    const uint32_t w = q[ch];
                    dequant_pair(w & 15u, (w >> 4) & 15u, s_, c_, hi4.x, lo4.x);
                    dequant_pair((w >> 8) & 15u, (w >> 12) & 15u, s_, c_, hi4.y, lo4.y);
                    dequant_pair((w >> 16) & 15u, (w >> 20) & 15u, s_, c_, hi4.z, lo4.z);
                    dequant_pair((w >> 24) & 15u, w >> 28, s_, c_, hi4.w, lo4.w);
                } else {
                    const uint32_t w0 = q[2 * ch], w1 = q[2 * ch + 1];
                    dequant_pair(w0 & 255u, (w0 >> 8) & 255u, s_, c_, hi4.x, lo4.x);
                    dequant_pair((w0 >> 16) & 255u, w0 >> 24, s_, c_, hi4.y, lo4.y);
                    dequant_pair(w1 & 255u, (w1 >> 8) & 255u, s_, c_, hi4.z, lo4.z);
                    dequant_pair((w1 >> 16) & 255u, w1 >> 24, s_, c_, hi4.w, lo4.w);
                }
                const uint32_t off = tile_off(tid, (uint32_t)ch, layout);
                *reinterpret_cast<uint4*>(bhi + off) = hi4;
                *reinterpret_cast<uint4*>(blo + off) = lo4;
            }
            fence_proxy_async_smem();                       // generic-proxy stores -> visible to tcgen05.mma (async proxy)
            mbar_arrive(full_bar(s));
        }

        // ================================ epilogue ================================
        mbar_wait(accum_bar, 0);
        tc_fence_after();
        const bool vec_ok = p.d_is_f32 ? ((p.n & 3u) == 0 && ((uintptr_t)p.d & 15u) == 0) : ((p.n & 7u) == 0 && ((uintptr_t)p.d & 15u) == 0);
#pragma unroll 1
        for (int mt = 0; mt < MT; ++mt) {
            const uint32_t row = m0 + (uint32_t)mt * 128u + warp * 32u + lane;
#pragma unroll 1
            for (int cb = 0; cb < 4; ++cb) {
                uint32_t v[32];
                tmem_ld_32x32(tmem_base + ((warp * 32u) << 16) + (uint32_t)mt * 128u + (uint32_t)cb * 32u, v);
                const uint32_t col0 = n0 + (uint32_t)cb * 32u;
                if (row >= p.m || col0 >= p.n) continue;
                const size_t obase = (size_t)row * p.n + col0;
                float f[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    float val = __fmul_rn(p.ab_scale, __uint_as_float(v[j]));   // separate roundings, like the reference
                    const uint32_t col = col0 + (uint32_t)j;
                    if (col < p.n) {
                        if (p.accumulate)
                            val = __fadd_rn(val, p.d_is_f32 ? reinterpret_cast<const float*>(p.d)[obase + j]
                                                            : bf2f(reinterpret_cast<const __nv_bfloat16*>(p.d)[obase + j]));
                        if (p.bias) val = __fadd_rn(val, bf2f(p.bias[col]));
                        if (p.has_soft_cap) val = __fmul_rn(p.soft_cap, tanhf(__fdiv_rn(val, p.soft_cap)));
                    }
                    f[j] = val;
                }
                if (vec_ok && col0 + 32u <= p.n) {
                    if (p.d_is_f32) {
                        float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.d) + obase);
#pragma unroll
                        for (int j = 0; j < 8; ++j) o[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
                    } else {
                        uint4* o = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.d) + obase);
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            o[j] = make_uint4(pack_bf16x2(f[8 * j], f[8 * j + 1]), pack_bf16x2(f[8 * j + 2], f[8 * j + 3]),
                                              pack_bf16x2(f[8 * j + 4], f[8 * j + 5]), pack_bf16x2(f[8 * j + 6], f[8 * j + 7]));
                    }
                } else {
#pragma unroll 1
                    for (int j = 0; j < 32; ++j) {
                        if (col0 + (uint32_t)j >= p.n) break;
                        if (p.d_is_f32) reinterpret_cast<float*>(p.d)[obase + j] = f[j];
                        else reinterpret_cast<__nv_bfloat16*>(p.d)[obase + j] = f2bf(f[j]);
                    }
                }
            }
        }
    } else {
        // ================================ MMA issuer (one lane) ================================
        if (lane == 0) {
            for (uint32_t kb = 0; kb < nkb; ++kb) {
                const uint32_t s = kb % STAGES, it = kb / STAGES;
                mbar_wait(full_bar(s), it & 1u);
                tc_fence_after();
                const uint32_t a_addr = smem_base + s * STAGE_BYTES;
                const uint32_t bh_addr = a_addr + A_BYTES, bl_addr = bh_addr + QMM_TILE_BYTES;
                const uint64_t hi = (uint64_t)p.desc_hi << 32;
                const uint32_t lbo = p.desc_lbo << 16;
#pragma unroll
                for (uint32_t kk = 0; kk < 4; ++kk) {
                    const uint64_t dbh = hi | (uint64_t)((((bh_addr >> 4) + kk * p.k_step) & 0x3FFFu) | lbo);
                    const uint64_t dbl = hi | (uint64_t)((((bl_addr >> 4) + kk * p.k_step) & 0x3FFFu) | lbo);
#pragma unroll
                    for (uint32_t mt = 0; mt < (uint32_t)MT; ++mt) {
                        const uint64_t da = hi | (uint64_t)(((((a_addr + mt * QMM_TILE_BYTES) >> 4) + kk * p.k_step) & 0x3FFFu) | lbo);
                        umma_bf16(tmem_base + mt * 128u, da, dbh, p.idesc, (kb | kk) != 0u ? 1u : 0u);
                        umma_bf16(tmem_base + mt * 128u, da, dbl, p.idesc, 1u);
                    }
                }
                umma_commit(empty_bar(s));                  // stage reusable once these MMAs have read it
            }
            umma_commit(accum_bar);                         // accumulators complete
        }
        __syncwarp();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// ---- host side -------------------------------------------------------------------------------------------------------
struct UmmaTuning {
    int layout = -1;               // -1 = default (SWIZZLE_128B)
    uint32_t desc_hi = 0, desc_lbo = 0, k_step = 0, idesc = 0;
    int mt = 0;                    // 0 = heuristic
    bool custom = false;
};
static UmmaTuning g_umma;

// kind::f16 instruction descriptor (cute/arch/mma_sm100_desc.hpp InstrDescriptor bit layout): c_format F32 = 1 @ [4,6),
// a_format BF16 = 1 @ [7,10), b_format BF16 = 1 @ [10,13), a_major = b_major = 0 (K-major) @ 15 / 16, n_dim = N >> 3 @ [17,23),
// m_dim = M >> 4 @ [24,29)
static constexpr uint32_t umma_idesc_bf16(uint32_t m, uint32_t n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((n >> 3) << 17) | ((m >> 4) << 24);
}

template <int BITS, int MT, int STAGES>
static void launch_qmm(uzu_command_buffer* cmd, const QmmParams& p, dim3 grid) {
    constexpr size_t smem = (size_t)STAGES * (QMM_TILE_BYTES * MT + 2 * QMM_TILE_BYTES) + 1024;
    static bool attr_done = false;
    if (!attr_done) {
        cudaFuncSetAttribute(qmm_umma_kernel<BITS, MT, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr_done = true;
    }
    launch(cmd, "matmul qmm_umma_kernel", qmm_umma_kernel<BITS, MT, STAGES>, grid, dim3(QMM_THREADS), smem, p);
}

bool prefill_gemm_applicable(const uzu_matmul_args& a) {
    static const int min_m = [] {
        const char* e = getenv("UZU_PREFILL_GEMM_MIN_M");   // 0 disables the tensor-core path
        return e ? atoi(e) : 64;
    }();
    if (min_m <= 0 || a.m < (uint32_t)min_m) return false;
    if (a.b_prologue == UZU_B_FULL_PRECISION || a.gather_indices || !a.b_transpose) return false;
    if (a.input_dt != UZU_DT_BF16 || a.weights_dt != UZU_DT_BF16) return false;
    if (a.output_dt != UZU_DT_BF16 && a.output_dt != UZU_DT_F32) return false;
    if (a.b_mode != UZU_QMODE_U4 && a.b_mode != UZU_QMODE_U8) return false;
    if (a.k % 64u != 0) return false;
    const uint32_t gs = a.b_group_size;
    if (gs != 32 && (gs % 64u) != 0) return false;         // a 32-k half block never straddles a group
    if (a.k % gs != 0) return false;
    if ((a.a & 15u) || (a.b & 15u)) return false;
    return true;
}

void encode_prefill_gemm(uzu_command_buffer* cmd, const uzu_matmul_args& a) {
    const uint32_t bits = a.b_mode == UZU_QMODE_U4 ? 4 : 8;
    QmmParams p{};
    p.w = (const uint8_t*)a.b;
    p.scales = (const __nv_bfloat16*)a.b_scales;
    p.zero_points = (const uint8_t*)a.b_zero_points;
    p.biases = (const __nv_bfloat16*)a.b_biases;
    p.x = (const __nv_bfloat16*)a.a;
    p.d = (void*)a.d;
    p.bias = ((a.d_transform & UZU_D_BIAS) && a.bias) ? (const __nv_bfloat16*)a.bias : nullptr;
    p.m = a.m; p.n = a.n; p.k = a.k;
    p.row_bytes = a.k * bits / 8;
    p.groups_per_row = a.k / a.b_group_size;
    p.zp_stride = bits == 4 ? (p.groups_per_row + 1) / 2 : p.groups_per_row;
    p.group_size = a.b_group_size;
    p.method = a.b_prologue == UZU_B_SCALE_BIAS_DEQUANT ? UZU_QMETHOD_SCALE_BIAS
               : a.b_prologue == UZU_B_SCALE_ZERO_POINT_DEQUANT ? UZU_QMETHOD_SCALE_ZERO_POINT : UZU_QMETHOD_SCALE_SYMMETRIC;
    p.xor_mask = a.b_signed_codes ? (bits == 4 ? 0x88888888u : 0x80808080u) : 0u;
    p.d_is_f32 = a.output_dt == UZU_DT_F32;
    p.accumulate = (a.d_transform & UZU_D_ACCUMULATE) != 0;
    p.has_soft_cap = (a.d_transform & UZU_D_SOFT_CAP) != 0;
    p.ab_scale = (a.d_transform & UZU_D_SCALE) ? a.ab_scale : 1.0f;
    p.soft_cap = a.soft_cap;
    // shared-memory matrix descriptor (cute/arch/mma_sm100_desc.hpp SmemDescriptor): start address >> 4 @ [0,14), LBO >> 4 @ [16,30),
    // SBO >> 4 @ [32,46), version = 1 @ [46,48), layout type @ [61,64) (2 = SWIZZLE_128B, 0 = none).
    // SWIZZLE_128B, K-major: atoms of 8 rows x 128 B, SBO = 1024 B between 8-row groups, LBO unused (1); UMMA_K = 16 bf16 = 32 B -> +2.
    p.layout = 0;
    p.desc_hi = (1024u >> 4) | (1u << 14) | (2u << 29);
    p.desc_lbo = 1;
    p.k_step = 2;
    p.idesc = umma_idesc_bf16(128, 128);
    if (g_umma.custom) {
        p.layout = (uint32_t)g_umma.layout;
        p.desc_hi = g_umma.desc_hi; p.desc_lbo = g_umma.desc_lbo; p.k_step = g_umma.k_step;
        if (g_umma.idesc) p.idesc = g_umma.idesc;
    }
    const uint32_t nt = (a.n + 127u) / 128u;
    int mt = (nt * ((a.m + 255u) / 256u) >= (uint32_t)cmd->ctx->sm_count || a.m > 4096u) ? 2 : 1;
    if (a.m <= 128u) mt = 1;
    if (g_umma.mt == 1 || g_umma.mt == 2) mt = g_umma.mt;
    const dim3 grid(nt, (a.m + 128u * mt - 1) / (128u * mt));
    if (bits == 4) {
        if (mt == 2) launch_qmm<4, 2, 3>(cmd, p, grid); else launch_qmm<4, 1, 4>(cmd, p, grid);
    } else {
        if (mt == 2) launch_qmm<8, 2, 3>(cmd, p, grid); else launch_qmm<8, 1, 4>(cmd, p, grid);
    }
}

}  // namespace uzu

extern "C" void uzu_debug_set_umma(int layout, uint32_t desc_hi, uint32_t desc_lbo, uint32_t k_step, uint32_t idesc, int mt) {
    uzu::g_umma.custom = layout >= 0;
    uzu::g_umma.layout = layout;
    uzu::g_umma.desc_hi = desc_hi; uzu::g_umma.desc_lbo = desc_lbo; uzu::g_umma.k_step = k_step; uzu::g_umma.idesc = idesc;
    uzu::g_umma.mt = mt;
}
