// Prefill MatmulKernel (m >= 64 tokens) for sm_100a: D[m,n] = epilogue(A[m,k] * dequant(B[n,k])^T) on the 5th-generation
// tensor cores (tcgen05.mma, accumulators in tensor memory), fed by an in-kernel dequantisation stage.
//
// Semantic spec: backends/cpu/kernel/matmul/kernel.rs:164-295 (GEMM branch of MatmulKernel; caller
// encodable_block/linear/matmul.rs:122-148, m = prefill chunk <= 1024). SURVEY 8, north_star row "prefill batched-GEMM".
//
// Shape of the kernel (one CTA = 128 output features x (MT x 128) tokens, K walked in blocks of 64):
//   * warps 0 .. NP/32-1 (NP = 256 threads by default) are the PRODUCERS. Per K block they
//       - copy the activation tile (bf16, K contiguous = "K-major") with cp.async straight into its UMMA stage, and
//       - dequantise their share of a weight row (NP/128 threads per row): codes -> `scale*code + corr` (the reference's
//         expression, bit for bit) -> TWO bf16 planes hi = bf16(w), lo = w - hi. For ZeroPoint / Symmetric weights
//         w = scale*(code - zp) has <= 16 significant bits, so hi + lo == w EXACTLY (8 + 8 bits); for MLX scale/bias the residual is
//         < 2^-17 |w|. Feeding a single rounded bf16 plane would put a 2^-9 relative error on every weight (about 1e-3 of the output
//         rms, i.e. the whole parity budget). 4-bit ZeroPoint / Symmetric runs in packed bf16x2 arithmetic (HADD2 / HMUL2 / HFMA2).
//     Tiles are written in the UMMA canonical K-major SWIZZLE_128B layout (8-row x 128-byte atoms, 16-byte chunk index XORed
//     with the row index inside the atom), made visible to the async proxy with fence.proxy.async, and handed to the MMA warp
//     through an mbarrier ("full", NP arrivals). The packed codes arrive through cooperative, coalesced cp.async "super-blocks"
//     (128 rows x 128 contiguous bytes) two buffers deep; a named barrier per super-block hands them from loaders to consumers.
//   * the last warp, one lane, is the MMA ISSUER: per K block 4 (K=16 steps) x MT x 2 (hi, lo plane) tcgen05.mma.cta_group::1.kind::f16
//     128x128x16 instructions accumulate into MT x 128 TMEM columns; bf16 x bf16 products are exact in the f32 accumulator, so
//     the result is sum_k x_k*w_k with f32 accumulation, the reference's arithmetic up to summation order. tcgen05.commit
//     releases the shared-memory stage ("empty" mbarrier) and, after the last block, signals the epilogue.
//   * the producer warps then run the EPILOGUE: tcgen05.ld (32 lanes x 32 columns per instruction; warp w owns TMEM lanes
//     32*(w mod 4) .. +31 = token rows, w / 4 picks the column blocks), ab_scale / accumulate / bias / soft-cap in the reference's
//     order, bf16 (RNE) or f32 store.
// Measured on B200 (profiles/README.md): 452 TFLOP/s useful at m = 2048 on the Llama-3-8B up projection, tensor pipe 43.5 % active;
// the limiter is the producers' instruction issue, the ceiling 50 % useful because of the two planes (DESIGN.md 4.5 has the history).
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
    uint32_t gshift;   // log2(group_size): quantisation group of element k = k >> gshift
    uint32_t desc_hi;  // upper 32 bits of the shared-memory matrix descriptor (SBO, version, layout type)
    uint32_t desc_lbo; // leading-dimension byte offset field (>> 4)
    uint32_t k_step;   // start-address increment (>> 4) per UMMA_K = 16 elements
    uint32_t idesc;    // instruction descriptor (kind::f16, bf16 x bf16 -> f32, M = 128, N = 128, both K-major)
    uint32_t packed_path;  // 4-bit ZeroPoint / Symmetric: bf16x2 dequant, the lo plane is stored negated and its MMAs negate B (idesc bit 14)
};

constexpr uint32_t QMM_TILE_BYTES = 128 * 64 * 2;  // one 128-row x 64-k bf16 tile
constexpr int QMM_DEFAULT_PRODUCERS = 256;          // producer threads (B200 sweep, profiles/README.md)
constexpr bool QMM_DEFAULT_COAL = true;             // cooperative coalesced code loads

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

// 16-byte asynchronous global -> shared copy (L2 only); src_bytes = 0 zero-fills the destination
__device__ __forceinline__ void cp_async16_zfill(uint32_t dst_smem, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst_smem), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit_group() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async8(uint32_t dst_smem, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst_smem), "l"(src) : "memory");
}
template <int N>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// byte offset of the 16-byte chunk `ch` (8 bf16 along k) of row `row` inside a 128-row x 64-k tile. (The no-swizzle core-matrix
// layout, LBO = 128 / SBO = 1024, was also verified with tools/umma_probe.py in an earlier revision; one layout is kept so the offset
// arithmetic is compile-time.)
__device__ __forceinline__ uint32_t tile_off(uint32_t row, uint32_t ch) {
    return row * 128u + ((ch ^ (row & 7u)) << 4);   // SWIZZLE_128B: Swizzle<3,4,3> on byte addresses (8-row x 128-byte atoms)
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo_elem, float hi_elem) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo_elem, hi_elem);   // .x (low 16 bits) = first element in memory
    return *reinterpret_cast<uint32_t*>(&v);
}

// Code -> f32 without a conversion instruction: the code is shifted into the mantissa of 128.0f (4-bit) / 256.0f (8-bit), giving
// BASE + code exactly; the BASE is folded into the per-group constant c' = corr - scale*BASE (exact for ZeroPoint / Symmetric, where
// corr = -scale*zp: scale*(zp + BASE) has <= 17 significant bits), so w = fmaf(scale, BASE + code, c') == scale*code + corr, the
// reference's expression (kernel.rs:267-276), bit for bit. `in_place_shift` = position of the code inside `word`.
template <int BITS>
__device__ __forceinline__ float code_to_f32(uint32_t word, int bit_pos) {
    constexpr int target = BITS == 4 ? 16 : 15;                       // mantissa bit that weighs 1.0 under the BASE exponent
    constexpr uint32_t mask = BITS == 4 ? 0x000F0000u : 0x007F8000u;
    constexpr uint32_t base = BITS == 4 ? 0x43000000u : 0x43800000u;  // 128.0f / 256.0f
    const uint32_t v = bit_pos <= target ? (word << (target - bit_pos)) : (word >> (bit_pos - target));
    return __uint_as_float((v & mask) | base);
}
__device__ __forceinline__ __nv_bfloat162 u32_as_bf162(uint32_t v) { return *reinterpret_cast<__nv_bfloat162*>(&v); }
__device__ __forceinline__ uint32_t bf162_as_u32(__nv_bfloat162 v) { return *reinterpret_cast<uint32_t*>(&v); }
// Two f32 weights -> packed hi plane word (RNE bf16) and lo plane word (the exact remainders, RNE bf16).
__device__ __forceinline__ void split_pair(float w0, float w1, uint32_t& hi, uint32_t& lo) {
    hi = pack_bf16x2(w0, w1);
    const float h0 = __uint_as_float(hi << 16), h1 = __uint_as_float(hi & 0xffff0000u);
    lo = pack_bf16x2(w0 - h0, w1 - h1);
}

// NP = producer threads (128 / 256 / 512): NP / 128 threads share a weight row, each dequantising 8 / (NP / 128) 16-byte chunks per
// K block. One producer warp per scheduler (NP = 128) is latency-bound (measured: 0.29 IPC, 3000 clk per K block against 1024 clk
// of MMA work); 4 warps per scheduler (NP = 512) hide the ALU / shared-memory latencies of the dequant chain.
// COAL = cooperative, coalesced loading of the packed codes: measured with ncu, per-thread loads of "my row's 32 bytes" cost one
// LSU wavefront per lane (32 rows = 32 different 128-byte lines per warp instruction, 256-512 wavefronts per K block, L1TEX 79 % busy).
// With COAL the producers fetch a SUPER-BLOCK of 128 rows x 128 contiguous bytes (4 K blocks of 4-bit / 2 of 8-bit codes) with 8
// consecutive lanes per row (4 lines per warp instruction, 8x fewer wavefronts) into one of two 16 KB buffers, XOR-swizzled by row
// so that the per-row reads of the dequantising threads are bank-conflict free; a named barrier among the producer warps per
// super-block hands the buffer from the loading threads to the consuming threads and recycles the other buffer.
template <int BITS, int MT, int STAGES, int NP, bool COAL>
__global__ void __launch_bounds__(NP + 32, 1) qmm_umma_kernel(const QmmParams p) {
    constexpr uint32_t A_BYTES = QMM_TILE_BYTES * MT;
    constexpr uint32_t STAGE_BYTES = A_BYTES + 2 * QMM_TILE_BYTES;
    constexpr uint32_t TMEM_COLS = 128 * MT;
    constexpr int TPR = NP / 128;                  // threads per weight row
    constexpr int CPT = 8 / TPR;                   // 16-byte bf16 chunks (8 k each) per thread per K block
    constexpr int WORDS = CPT * BITS / 4;          // 32-bit words of packed codes per thread per K block
    constexpr int PIECE = WORDS >= 4 ? 16 : WORDS * 4;   // cp.async granule for the raw codes (16 or 8 bytes)
    constexpr int PIECES = WORDS * 4 / PIECE;
    constexpr int A_PER_THREAD = MT * 1024 / NP;   // activation chunks copied per thread per K block
    static_assert(NP == 128 || NP == 256 || NP == 512, "producer threads");
    static_assert(PIECE == 16 || PIECE == 8, "raw code granule");

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
                mbar_init(full_bar(s), NP);
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

    if (tid < (uint32_t)NP) {
        // ================================ producers ================================
        const uint32_t trow = tid & 127u, part = tid >> 7;          // `part` is warp-uniform: which CPT chunks of the row
        const uint32_t brow = min(n0 + trow, p.n - 1);              // rows past n are computed from a clamped row and never stored
        const uint8_t* wrow = p.w + (size_t)brow * p.row_bytes + part * (WORDS * 4);
        const __nv_bfloat16* srow = p.scales + (size_t)brow * p.groups_per_row;
        const uint8_t* zrow = p.zero_points ? p.zero_points + (size_t)brow * p.zp_stride : nullptr;
        const __nv_bfloat16* crow = p.biases ? p.biases + (size_t)brow * p.groups_per_row : nullptr;
        // Software pipeline (all copies are cp.async, i.e. need no registers while in flight):
        //   weights: raw packed codes of K block j go to a per-thread slot of a RAW-deep ring PW = RAW - 2 blocks ahead (HBM latency),
        //   activations: the bf16 tile of K block j is copied straight into its swizzled UMMA stage PA = STAGES - 2 blocks ahead
        //                (the stage must have been released by the MMA warp first; PA < STAGES - 1 leaves the MMA one block of slack),
        //   scales / zero points: plain loads one block ahead (same cache line for 32+ consecutive blocks).
        // Iteration t issues {A(t), W(t + DW)} as cp.async group t and then dequantises block t - PA.
        constexpr int PA = STAGES - 2;
        constexpr int RAW = BITS == 4 ? 8 : 4;
        constexpr int PW = RAW - 2;
        constexpr int DW = PW - PA;
        static_assert(PA >= 1 && DW >= 0, "pipeline distances");
        constexpr uint32_t ROW_CODE_BYTES = BITS * 8u;                          // packed bytes of one row per K block (64 weights)
        constexpr uint32_t RAW_SLOT_BYTES = ROW_CODE_BYTES * 128u;
        constexpr uint32_t CODE_BASE = BITS == 4 ? 128u : 256u;
        constexpr uint32_t SB = 128u / ROW_CODE_BYTES;                          // K blocks per super-block (COAL): 4 (4-bit) / 2 (8-bit)
        constexpr uint32_t SB_BYTES = 128u * 128u;                              // 128 rows x 128 bytes
        constexpr int SB_PIECES = 1024 / NP;                                    // 16-byte pieces per thread per super-block
        const uint32_t raw_base = smem_base + (uint32_t)STAGES * STAGE_BYTES;
        const uint8_t* raw_ptr = smem + (size_t)STAGES * STAGE_BYTES;
        auto issue_w = [&](uint32_t j) {
            const uint32_t slot = raw_base + (j % RAW) * RAW_SLOT_BYTES + tid * PIECE;   // [piece][thread][PIECE]: conflict-free reads
#pragma unroll
            for (int i = 0; i < PIECES; ++i) {
                if constexpr (PIECE == 16) cp_async16_zfill(slot + (uint32_t)i * (NP * 16u), wrow + (size_t)j * ROW_CODE_BYTES + i * 16, 16u);
                else cp_async8(slot + (uint32_t)i * (NP * 8u), wrow + (size_t)j * ROW_CODE_BYTES + i * 8);
            }
        };
        // COAL: super-block `sb` = K blocks [sb*SB, sb*SB + SB): piece (row, seg) = 16 bytes at byte seg*16 of the row's 128, stored at
        // row*128 + ((seg ^ (row & 7)) << 4). Rows past n re-read row n-1 (never stored); a tail super-block shorter than 128 bytes
        // zero-fills the K blocks that do not exist (never consumed).
        const uint32_t k_bytes = p.row_bytes;
        auto issue_sb = [&](uint32_t sb) {
            const uint32_t buf = raw_base + (sb & 1u) * SB_BYTES;
#pragma unroll
            for (int i = 0; i < SB_PIECES; ++i) {
                const uint32_t pi = (uint32_t)i * NP + tid, row = pi >> 3, seg = pi & 7u;
                const uint32_t grow = min(n0 + row, p.n - 1);
                const uint32_t off = sb * 128u + seg * 16u;                     // byte offset inside the weight row
                const bool ok = off < k_bytes;
                cp_async16_zfill(buf + row * 128u + ((seg ^ (row & 7u)) << 4), p.w + (size_t)grow * k_bytes + (ok ? off : 0u), ok ? 16u : 0u);
            }
        };
        auto issue_a = [&](uint32_t j, uint32_t stage) {
            const uint32_t st = smem_base + stage * STAGE_BYTES;
#pragma unroll
            for (int i = 0; i < A_PER_THREAD; ++i) {
                const uint32_t c = (uint32_t)i * NP + tid, row = c >> 3, ch = c & 7u;
                const uint32_t grow = m0 + row;
                const bool ok = grow < p.m;                                            // rows past m are zero-filled (src-size 0)
                const __nv_bfloat16* src = p.x + (size_t)(ok ? grow : 0u) * p.k + (size_t)j * 64u + ch * 8u;
                cp_async16_zfill(st + (row >> 7) * QMM_TILE_BYTES + tile_off(row & 127u, ch), src, ok ? 16u : 0u);
            }
        };
        // the thread's CPT chunks cover k in [k0, k0 + 8*CPT) of the block; two halves so that group size 32 is handled at NP = 128
        struct ScRaw { uint32_t s[2], c[2]; };
        const uint32_t gshift = p.gshift;
        auto group_of = [&](uint32_t j, int h) { return (j * 64u + part * (CPT * 8u) + (uint32_t)h * (CPT * 4u)) >> gshift; };
        auto load_sc = [&](uint32_t j, ScRaw& r) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const uint32_t g = group_of(j, h);
                r.s[h] = reinterpret_cast<const uint16_t*>(srow)[g];
                if (p.method == UZU_QMETHOD_SCALE_ZERO_POINT) r.c[h] = BITS == 4 ? zrow[g >> 1] : zrow[g];
                else if (p.method == UZU_QMETHOD_SCALE_BIAS) r.c[h] = reinterpret_cast<const uint16_t*>(crow)[g];
                else r.c[h] = 0u;
            }
        };
        if constexpr (COAL) {
            issue_sb(0);                                     // joins cp.async group 0; weights do not depend on the previous kernel
        } else {
#pragma unroll
            for (int j = 0; j < DW; ++j)
                if ((uint32_t)j < nkb) issue_w((uint32_t)j);
        }
        ScRaw nxt;
        load_sc(0, nxt);
        pdl_wait();                                         // activations come from the previous kernel
        uint32_t a_stage = 0, a_parity = 1;                 // issue side: stage of K block t, parity its `empty` wait needs
        uint32_t c_stage = 0;                               // consume side: stage of K block t - PA
        uint32_t c_sub = 0, c_buf = 0;                      // COAL: K block index inside its super-block, buffer of that super-block
        for (uint32_t t = 0; t < nkb + PA; ++t) {
            if constexpr (COAL) {
                // first K block of a super-block: its codes were issued SB iterations ago (group t - SB; super-block 0: before the
                // loop) -> wait for this thread's pieces, then the named barrier publishes everyone's pieces and proves that every
                // thread has finished the previous super-block, whose buffer the next super-block's loads may now overwrite.
                if (t >= (uint32_t)PA && c_sub == 0u) {
                    if (t == (uint32_t)PA) cp_async_wait_group<0>();
                    else cp_async_wait_group<(int)SB - 1>();
                    asm volatile("bar.sync 1, %0;" ::"n"(NP) : "memory");
                    const uint32_t nsb = (t - PA) / SB + 1u;            // SB is a power of two
                    if (nsb * SB < nkb) issue_sb(nsb);
                }
            }
            if (t < nkb) {
                mbar_wait(empty_bar(a_stage), a_parity);
                issue_a(t, a_stage);
                if (++a_stage == (uint32_t)STAGES) { a_stage = 0; a_parity ^= 1u; }
                if constexpr (!COAL) {
                    if (t + DW < nkb) issue_w(t + DW);
                }
            }
            cp_async_commit_group();
            if (t < (uint32_t)PA) continue;
            const uint32_t kb = t - PA, s = c_stage;
            if (++c_stage == (uint32_t)STAGES) c_stage = 0;
            const ScRaw cur = nxt;
            if (kb + 1 < nkb) load_sc(kb + 1, nxt);
            float sc[2], cc[2];
            uint32_t sc2[2], zb2[2];                         // packed bf16x2 (scale, scale) and (128 + zp, 128 + zp)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const uint32_t g = group_of(kb, h);
                const float s_ = __uint_as_float(cur.s[h] << 16);
                sc2[h] = cur.s[h] | (cur.s[h] << 16);
                zb2[h] = 0u;
                float c_;
                if (p.method == UZU_QMETHOD_SCALE_ZERO_POINT) {
                    const uint32_t z = BITS == 4 ? ((g & 1u) ? (cur.c[h] >> 4) : (cur.c[h] & 15u)) : cur.c[h];
                    c_ = -s_ * (float)(z + CODE_BASE);              // corr - scale*BASE, exact
                    zb2[h] = (0x4300u | z) * 0x00010001u;           // bf16 bits of 128 + z (z < 128), both halves
                } else if (p.method == UZU_QMETHOD_SCALE_BIAS) {
                    c_ = __uint_as_float(cur.c[h] << 16);
                } else {
                    c_ = -s_ * (float)((1u << (BITS - 1)) + CODE_BASE);
                    zb2[h] = (0x4300u | 8u) * 0x00010001u;
                }
                sc[h] = s_;
                cc[h] = c_;
            }
            cp_async_wait_group<PA>();                      // groups <= kb have landed: this thread's A chunks and its raw codes
            uint32_t q[WORDS];
            if constexpr (COAL) {
                // this thread's WORDS*4 bytes sit at byte (kb % SB)*ROW_CODE_BYTES + part*WORDS*4 of its row's 128-byte super-block line
                const uint8_t* line = raw_ptr + (size_t)c_buf * SB_BYTES + trow * 128u;
                const uint32_t o = c_sub * ROW_CODE_BYTES + part * (WORDS * 4u);
                if (++c_sub == SB) { c_sub = 0; c_buf ^= 1u; }
#pragma unroll
                for (int i = 0; i < PIECES; ++i) {
                    const uint32_t ob = o + (uint32_t)i * PIECE;
                    const uint8_t* src = line + (((ob >> 4) ^ (trow & 7u)) << 4) + (ob & 15u);
                    if constexpr (PIECE == 16) {
                        const uint4 v4 = *reinterpret_cast<const uint4*>(src);
                        q[4 * i] = v4.x ^ p.xor_mask; q[4 * i + 1] = v4.y ^ p.xor_mask; q[4 * i + 2] = v4.z ^ p.xor_mask; q[4 * i + 3] = v4.w ^ p.xor_mask;
                    } else {
                        const uint2 v2 = *reinterpret_cast<const uint2*>(src);
                        q[2 * i] = v2.x ^ p.xor_mask; q[2 * i + 1] = v2.y ^ p.xor_mask;
                    }
                }
            } else {
#pragma unroll
                for (int i = 0; i < PIECES; ++i) {
                    const uint8_t* src = raw_ptr + (size_t)(kb % RAW) * RAW_SLOT_BYTES + (size_t)i * (NP * PIECE) + tid * PIECE;
                    if constexpr (PIECE == 16) {
                        const uint4 v4 = *reinterpret_cast<const uint4*>(src);
                        q[4 * i] = v4.x ^ p.xor_mask; q[4 * i + 1] = v4.y ^ p.xor_mask; q[4 * i + 2] = v4.z ^ p.xor_mask; q[4 * i + 3] = v4.w ^ p.xor_mask;
                    } else {
                        const uint2 v2 = *reinterpret_cast<const uint2*>(src);
                        q[2 * i] = v2.x ^ p.xor_mask; q[2 * i + 1] = v2.y ^ p.xor_mask;
                    }
                }
            }
            uint8_t* st = smem + (size_t)s * STAGE_BYTES;
            uint8_t* bhi = st + A_BYTES;
            uint8_t* blo = bhi + QMM_TILE_BYTES;
            if (BITS == 4 && p.packed_path) {
                // ---- packed bf16x2 path (4-bit ZeroPoint / Symmetric): w = s*(q - z) ----
                // d = (128 + q) - (128 + z) is an exact small integer; hi = RNE_bf16(s*d) is ONE HMUL2; -lo = fma(-s, d, hi) is exact
                // (s*d has <= 13 significant bits, the remainder <= 8) -> hi + lo == s*(q - z) == the reference's scale*code + corr.
                // Nibbles p and p+4 of a word are 16 bits apart, so one shift + one LOP3 yields the pair (e_p, e_p+4); the planes are
                // put back into k order with PRMT before the 16-byte stores. The lo plane is stored NEGATED (no separate negation);
                // its MMAs run with the instruction descriptor's negate-B bit.
#pragma unroll
                for (int ch = 0; ch < CPT; ++ch) {
                    const int h = ch / (CPT / 2);
                    const uint32_t s2 = sc2[h], ns2 = s2 ^ 0x80008000u, z2 = zb2[h];
                    uint32_t H[4], L[4];
#pragma unroll
                    for (int pr = 0; pr < 4; ++pr) {
                        const uint32_t pq = ((q[ch] >> (4 * pr)) & 0x000F000Fu) | 0x43004300u;      // bf16x2 (128 + e_p, 128 + e_p+4)
                        const __nv_bfloat162 d = __hsub2(u32_as_bf162(pq), u32_as_bf162(z2));
                        const __nv_bfloat162 hh = __hmul2(u32_as_bf162(s2), d);
                        const __nv_bfloat162 nl = __hfma2(u32_as_bf162(ns2), d, hh);
                        H[pr] = bf162_as_u32(hh);
                        L[pr] = bf162_as_u32(nl);
                    }
                    uint4 hi4, lo4;
                    hi4.x = __byte_perm(H[0], H[1], 0x5410); hi4.y = __byte_perm(H[2], H[3], 0x5410);
                    hi4.z = __byte_perm(H[0], H[1], 0x7632); hi4.w = __byte_perm(H[2], H[3], 0x7632);
                    lo4.x = __byte_perm(L[0], L[1], 0x5410); lo4.y = __byte_perm(L[2], L[3], 0x5410);
                    lo4.z = __byte_perm(L[0], L[1], 0x7632); lo4.w = __byte_perm(L[2], L[3], 0x7632);
                    const uint32_t off = tile_off(trow, part * CPT + (uint32_t)ch);
                    *reinterpret_cast<uint4*>(bhi + off) = hi4;
                    *reinterpret_cast<uint4*>(blo + off) = lo4;
                }
            } else {
                // ---- f32 path (8-bit codes, MLX scale/bias): w = fmaf(s, code, c) in f32, then the hi / lo split ----
                const bool mlx = p.method == UZU_QMETHOD_SCALE_BIAS;
#pragma unroll
                for (int ch = 0; ch < CPT; ++ch) {
                    const int h = ch / (CPT / 2);
                    const float s_ = sc[h], c_ = cc[h];
                    float wv[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        float f;
                        if constexpr (BITS == 4) f = code_to_f32<4>(q[ch], 4 * e);
                        else f = code_to_f32<8>(q[(2 * ch + (e >> 2)) % WORDS], 8 * (e & 3));
                        if (mlx) f = __fadd_rn(f, -(float)CODE_BASE);   // MLX: keep the reference's scale*code + bias rounding
                        wv[e] = fmaf(s_, f, c_);
                    }
                    uint4 hi4, lo4;
                    split_pair(wv[0], wv[1], hi4.x, lo4.x);
                    split_pair(wv[2], wv[3], hi4.y, lo4.y);
                    split_pair(wv[4], wv[5], hi4.z, lo4.z);
                    split_pair(wv[6], wv[7], hi4.w, lo4.w);
                    const uint32_t off = tile_off(trow, part * CPT + (uint32_t)ch);
                    *reinterpret_cast<uint4*>(bhi + off) = hi4;
                    *reinterpret_cast<uint4*>(blo + off) = lo4;
                }
            }
            fence_proxy_async_smem();                       // generic-proxy writes (st.shared, cp.async) -> visible to tcgen05.mma
            mbar_arrive(full_bar(s));
        }

        // ================================ epilogue ================================
        // warp w reads TMEM lanes 32*(w % 4) .. +31 (token rows); with more than 4 producer warps the 32-column blocks are split by w / 4
        mbar_wait(accum_bar, 0);
        tc_fence_after();
        const bool vec_ok = p.d_is_f32 ? ((p.n & 3u) == 0 && ((uintptr_t)p.d & 15u) == 0) : ((p.n & 7u) == 0 && ((uintptr_t)p.d & 15u) == 0);
        const uint32_t lane_grp = warp & 3u;
#pragma unroll 1
        for (int mt = 0; mt < MT; ++mt) {
            const uint32_t row = m0 + (uint32_t)mt * 128u + lane_grp * 32u + lane;
#pragma unroll 1
            for (uint32_t cb = warp >> 2; cb < 4u; cb += (uint32_t)TPR) {
                uint32_t v[32];
                tmem_ld_32x32(tmem_base + ((lane_grp * 32u) << 16) + (uint32_t)mt * 128u + cb * 32u, v);
                const uint32_t col0 = n0 + cb * 32u;
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
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        if (col0 + (uint32_t)j < p.n) {
                            if (p.d_is_f32) reinterpret_cast<float*>(p.d)[obase + j] = f[j];
                            else reinterpret_cast<__nv_bfloat16*>(p.d)[obase + j] = f2bf(f[j]);
                        }
                    }
                }
            }
        }
    } else {
        // ================================ MMA issuer (one lane) ================================
        if (lane == 0) {
            uint32_t s = 0, parity = 0;
            for (uint32_t kb = 0; kb < nkb; ++kb) {
                mbar_wait(full_bar(s), parity);
                tc_fence_after();
                const uint32_t a_addr = smem_base + s * STAGE_BYTES;
                const uint32_t bh_addr = a_addr + A_BYTES, bl_addr = bh_addr + QMM_TILE_BYTES;
                const uint64_t hi = (uint64_t)p.desc_hi << 32;
                const uint32_t lbo = p.desc_lbo << 16;
                const uint32_t idesc_lo = (BITS == 4 && p.packed_path) ? (p.idesc | (1u << 14)) : p.idesc;
#pragma unroll
                for (uint32_t kk = 0; kk < 4; ++kk) {
                    const uint64_t dbh = hi | (uint64_t)((((bh_addr >> 4) + kk * p.k_step) & 0x3FFFu) | lbo);
                    const uint64_t dbl = hi | (uint64_t)((((bl_addr >> 4) + kk * p.k_step) & 0x3FFFu) | lbo);
#pragma unroll
                    for (uint32_t mt = 0; mt < (uint32_t)MT; ++mt) {
                        const uint64_t da = hi | (uint64_t)(((((a_addr + mt * QMM_TILE_BYTES) >> 4) + kk * p.k_step) & 0x3FFFu) | lbo);
                        umma_bf16(tmem_base + mt * 128u, da, dbh, p.idesc, (kb | kk) != 0u ? 1u : 0u);
                        umma_bf16(tmem_base + mt * 128u, da, dbl, idesc_lo, 1u);
                    }
                }
                umma_commit(empty_bar(s));                  // stage reusable once these MMAs have read it
                if (++s == (uint32_t)STAGES) { s = 0; parity ^= 1u; }
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
    int np = 0;                    // producer threads, 0 = default
    bool custom = false;
};
static UmmaTuning g_umma;

// kind::f16 instruction descriptor (cute/arch/mma_sm100_desc.hpp InstrDescriptor bit layout): c_format F32 = 1 @ [4,6),
// a_format BF16 = 1 @ [7,10), b_format BF16 = 1 @ [10,13), a_major = b_major = 0 (K-major) @ 15 / 16, n_dim = N >> 3 @ [17,23),
// m_dim = M >> 4 @ [24,29)
static constexpr uint32_t umma_idesc_bf16(uint32_t m, uint32_t n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((n >> 3) << 17) | ((m >> 4) << 24);
}

template <int BITS, int MT, int STAGES, int NP, bool COAL>
static void launch_qmm(uzu_command_buffer* cmd, const QmmParams& p, dim3 grid) {
    // stages + raw-code ring (32 KB: 8 / 4 per-thread K-block slots, or two 16 KB super-block buffers) + alignment slack
    constexpr size_t smem = (size_t)STAGES * (QMM_TILE_BYTES * MT + 2 * QMM_TILE_BYTES) + 32768 + 1024;
    static std::atomic<uint64_t> smem_opt_in{0};
    opt_in_dynamic_smem(cmd, qmm_umma_kernel<BITS, MT, STAGES, NP, COAL>, (int)((int)smem), smem_opt_in);
    launch(cmd, "matmul qmm_umma_kernel", qmm_umma_kernel<BITS, MT, STAGES, NP, COAL>, grid, dim3(NP + 32), smem, p);
}
template <int BITS, int MT, int STAGES>
static void launch_qmm_np(uzu_command_buffer* cmd, const QmmParams& p, dim3 grid, int np, bool coal) {
    if (coal) {
        if (np == 128) launch_qmm<BITS, MT, STAGES, 128, true>(cmd, p, grid);
        else if (np == 256) launch_qmm<BITS, MT, STAGES, 256, true>(cmd, p, grid);
        else launch_qmm<BITS, MT, STAGES, 512, true>(cmd, p, grid);
    } else {
        if (np == 128) launch_qmm<BITS, MT, STAGES, 128, false>(cmd, p, grid);
        else if (np == 256) launch_qmm<BITS, MT, STAGES, 256, false>(cmd, p, grid);
        else launch_qmm<BITS, MT, STAGES, 512, false>(cmd, p, grid);
    }
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
    if (gs < 32 || (gs & (gs - 1)) != 0) return false;      // power-of-two groups >= 32: a thread's span never straddles a group, index = k >> log2(gs)
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
    p.gshift = 0;
    while ((1u << p.gshift) < a.b_group_size) ++p.gshift;
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
    p.desc_hi = (1024u >> 4) | (1u << 14) | (2u << 29);
    p.desc_lbo = 1;
    p.k_step = 2;
    p.idesc = umma_idesc_bf16(128, 128);
    static const bool packed_ok = [] { const char* e = getenv("UZU_QMM_PACKED"); return !e || atoi(e) != 0; }();   // 0: f32 dequant path for A/B runs
    p.packed_path = (packed_ok && bits == 4 && p.method != UZU_QMETHOD_SCALE_BIAS) ? 1u : 0u;
    if (g_umma.custom) {
        p.desc_hi = g_umma.desc_hi; p.desc_lbo = g_umma.desc_lbo; p.k_step = g_umma.k_step;
        if (g_umma.idesc) p.idesc = g_umma.idesc;
    }
    const uint32_t nt = (a.n + 127u) / 128u;
    int mt = (nt * ((a.m + 255u) / 256u) >= (uint32_t)cmd->ctx->sm_count || a.m > 4096u) ? 2 : 1;
    if (a.m <= 128u) mt = 1;
    if (g_umma.mt == 1 || g_umma.mt == 2) mt = g_umma.mt;
    const dim3 grid(nt, (a.m + 128u * mt - 1) / (128u * mt));
    static const int np_env = [] { const char* e = getenv("UZU_QMM_PRODUCERS"); return e ? atoi(e) : 0; }();   // 128 | 256 | 512 (A/B runs)
    const int np = g_umma.np ? g_umma.np : ((np_env == 128 || np_env == 256 || np_env == 512) ? np_env : QMM_DEFAULT_PRODUCERS);
    static const int coal_env = [] { const char* e = getenv("UZU_QMM_COAL"); return e ? atoi(e) : -1; }();       // 0 | 1 (A/B runs)
    const bool coal = coal_env < 0 ? QMM_DEFAULT_COAL : coal_env != 0;
    if (bits == 4) {
        if (mt == 2) launch_qmm_np<4, 2, 3>(cmd, p, grid, np, coal); else launch_qmm_np<4, 1, 4>(cmd, p, grid, np, coal);
    } else {
        if (mt == 2) launch_qmm_np<8, 2, 3>(cmd, p, grid, np, coal); else launch_qmm_np<8, 1, 4>(cmd, p, grid, np, coal);
    }
}

}  // namespace uzu

extern "C" void uzu_debug_set_umma(int layout, uint32_t desc_hi, uint32_t desc_lbo, uint32_t k_step, uint32_t idesc, int mt) {
    uzu::g_umma.custom = layout >= 0;
    uzu::g_umma.layout = layout;
    uzu::g_umma.desc_hi = desc_hi; uzu::g_umma.desc_lbo = desc_lbo; uzu::g_umma.k_step = k_step; uzu::g_umma.idesc = idesc;
    uzu::g_umma.mt = mt & 0xff;                      // low byte: tokens-per-CTA factor; bits 8.. : producer threads (128 | 256 | 512)
    const int np = mt >> 8;
    uzu::g_umma.np = (np == 128 || np == 256 || np == 512) ? np : 0;
}
