// Persistent whole-token decode kernel for B200 (sm_100a): ONE cooperative launch runs the complete per-token forward pass of
// LanguageModelStream::next() (engine/language_model/stream/stream.rs:363-782; op order of encodable_block/{decoder.rs:138-203,
// transformer.rs:226-329, transformer_layer.rs:194-238}) as a program of phases separated by grid barriers.
//
// Why: the per-kernel decode step is launch- and ramp-bound (round 1: 137-197 dependent launches of 4-8 us, each restarting the HBM
// stream from empty). Here the weight stream never stops:
//   * one CTA per SM, MK_NCW consumer warps + 1 producer warp. The producer warp walks the CTA's static schedule of weight units
//     (16 rows x 512 nibbles = 4608 B incl. per-group coefficients) and feeds per-warp shared-memory rings with cp.async.bulk (TMA bulk
//     copies, UBLKCP in SASS) completing on mbarriers. It does NOT take part in the grid barriers: weights never depend on
//     activations, so while the consumers wait at a barrier / stage the next activation row the rings already hold the next
//     phase's first units (the whole ring of all SMs is ~22 MB: for Qwen3.5-0.8B that is more than a full layer).
//   * weights are read from a decode-stream copy built once at load (mega_repack_kernel): unit-major, each unit exactly in the order
//     the consumer lanes read it (conflict-free LDS.128), so a warp's whole range of a matrix is ONE contiguous byte range of HBM.
//   * work partition = "stream-K": the units of a phase are cut into gridDim * MK_NCW equal contiguous ranges (+-1 unit), one per
//     consumer warp, regardless of tile boundaries. Every (tile, range) intersection leaves a 16-row partial sum ("piece") in a
//     small f32 workspace; whoever consumes the matmul output sums a tile's pieces in piece order (deterministic), rounds once to
//     bf16 (= the reference's matmul output rounding) and continues with the reference's arithmetic (norm / gate / act / RoPE).
//   * the dot products use the validated inner loop of qmv_decode_async_kernel (matmul.cu): nibbles -> exact bf16 (128 + q) with one
//     LOP3 per pair, mma.sync m16n8k16 as the f32-accumulating dot engine, affine part hoisted per quantisation group.
//   * attention / DeltaNet / activation / argmax phases run on the same consumer warps with plain loads.
// Every spin loop has a clock watchdog: on a timeout the kernel raises error_flag and drains instead of hanging the GPU.
#include <algorithm>
#include <cstdlib>

#include "decode_mega.cuh"

namespace uzu {

// ---------------------------------------------------------------------------------------------------------------------------------
// PTX helpers
// ---------------------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t mk_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mk_mbar_init(uint32_t addr, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(addr), "r"(count) : "memory");
}
__device__ __forceinline__ void mk_mbar_expect_tx(uint32_t addr, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(addr), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mk_mbar_arrive(uint32_t addr) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(addr) : "memory"); }
__device__ __forceinline__ bool mk_mbar_try_wait(uint32_t addr, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok)
                 : "r"(addr), "r"(parity)
                 : "memory");
    return ok != 0;
}
__device__ __forceinline__ bool mk_mbar_test_wait(uint32_t addr, uint32_t parity) {     // non-blocking
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok)
                 : "r"(addr), "r"(parity)
                 : "memory");
    return ok != 0;
}
// the weight stream is read exactly once per token: evict-first keeps the small hot data (pieces, residual rows, norm scales) in L2
__device__ __forceinline__ void mk_bulk_copy(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t mbar, uint64_t policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst_smem), "l"(src),
                 "r"(bytes), "r"(mbar), "l"(policy)
                 : "memory");
}
__device__ __forceinline__ uint64_t mk_evict_first_policy() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void mk_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
__device__ __forceinline__ unsigned int mk_ld_acquire(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void mk_st_release(unsigned int* p, unsigned int v) { asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

constexpr long long MK_TIMEOUT_CLOCKS = 3000000000ll;   // ~1.5 s at 1.9 GHz

struct MkWatch {
    unsigned int* flag;        // device word polled by the spin loops (never host memory: a poll over PCIe from every thread costs milliseconds)
    unsigned int* host_flag;   // pinned host mirror, written only on failure
    bool dead;
    __device__ __forceinline__ void fail(unsigned int code) {
        *reinterpret_cast<volatile unsigned int*>(flag) = code;     // plain stores (a racing second code is harmless)
        *reinterpret_cast<volatile unsigned int*>(host_flag) = code;
        dead = true;
    }
    __device__ __forceinline__ bool poll_dead() {
        if (!dead && *reinterpret_cast<volatile unsigned int*>(flag) != 0u) dead = true;
        return dead;
    }
};

__device__ __forceinline__ void mk_wait(uint32_t mbar, uint32_t parity, MkWatch& w, unsigned int code) {
    if (w.dead) return;
    if (mk_mbar_try_wait(mbar, parity)) return;
    const long long t0 = clock64();
    uint32_t spins = 0;
    while (!mk_mbar_try_wait(mbar, parity)) {
        if ((++spins & 255u) == 0u) {
            if (w.poll_dead()) return;
            if (clock64() - t0 > MK_TIMEOUT_CLOCKS) { w.fail(code); return; }
        }
    }
}

// Grid barrier over the consumer warps of all CTAs (the producer warp never waits here). One monotonic 64-bit arrival counter that is
// never reset: barrier number `index` of a launch completes when the counter reaches barrier_base + (index + 1) * gridDim (the host
// advances barrier_base by barriers_per_launch * gridDim per launch). Arrive = red.release (fire and forget: no round trip before the
// wait starts), wait = ld.acquire polling by one thread. Called by all consumer threads.
__device__ __forceinline__ void mk_grid_sync(const MkParams& p, uint32_t index, int ctid, int nct, MkWatch& w) {
    mk_bar_sync(1, nct);
    if (ctid == 0 && !w.dead) {
        const unsigned long long target = p.barrier_base + (unsigned long long)(index + 1) * gridDim.x;
        asm volatile("red.release.gpu.global.add.u64 [%0], 1;" ::"l"(p.barrier) : "memory");
        const long long t0 = clock64();
        uint32_t spins = 0;
        for (;;) {
            unsigned long long v;
            asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p.barrier) : "memory");   // acquire once, after the loop
            if (v >= target) break;
            if ((++spins & 63u) == 0u) {
                if (w.poll_dead()) break;
                if (clock64() - t0 > MK_TIMEOUT_CLOCKS) { w.fail(0x100u); break; }
            }
        }
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
    }
    mk_bar_sync(1, nct);
}

__device__ __forceinline__ void mk_mma_16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
// (128 + q_lo, 128 + q_hi) as two exact bf16 from nibbles `shift` and `shift + 16` of w: one shift + one LOP3 (magic kept in a register).
// The GEMV body is bound by the integer ALU pipe (ncu r2: ~75 % active inside the bodies, FMA pipe ~6 %): UZU_MK_SHIFT_MODE moves shifts to
// the FMA pipe as multiply-high (IMAD.HI). 1: plain shifts; 2: shifts by 4 and 12 on the FMA pipe, by 8 on the ALU pipe; 0: all on FMA.
#ifndef UZU_MK_SHIFT_MODE
#define UZU_MK_SHIFT_MODE 1
#endif
__device__ __forceinline__ uint32_t mk_nib_pair(uint32_t w, int shift, uint32_t magic) {
    uint32_t r, s;
    if (shift == 0) s = w;
    else if (UZU_MK_SHIFT_MODE == 1 || (UZU_MK_SHIFT_MODE == 2 && shift == 8)) s = w >> shift;
    else s = __umulhi(w, 1u << (32 - shift));
    asm("lop3.b32 %0, %1, 0x000f000f, %2, 0xea;" : "=r"(r) : "r"(s), "r"(magic));   // (s & mask) | magic
    return r;
}

// ---------------------------------------------------------------------------------------------------------------------------------
// pieces
// ---------------------------------------------------------------------------------------------------------------------------------
// bf16-rounded matmul output rows [row, row + 4) (row % 4 == 0): sum of the tile's P piece slots in slot order (unused slots hold zeros),
// one RNE rounding. All P loads are independent: one L2 round trip.
__device__ __forceinline__ float4 mk_rows4_raw(const MkPieces& pc, uint32_t row) {
    const uint32_t tile = row >> 4, r = row & 15u;
    const float4* base = reinterpret_cast<const float4*>(pc.pieces + ((size_t)tile * pc.P) * 16 + r);
    float4 s = __ldcg(base);
#pragma unroll 4
    for (uint32_t q = 1; q < pc.P; ++q) {
        const float4 v = __ldcg(base + (size_t)q * 4);
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    return s;
}
__device__ __forceinline__ float4 mk_rows4(const MkPieces& pc, uint32_t row) {
    float4 s = mk_rows4_raw(pc, row);
    s.x = round_bf16(s.x); s.y = round_bf16(s.y); s.z = round_bf16(s.z); s.w = round_bf16(s.w);
    return s;
}
// N rows at once, slot-outer / row-inner: every iteration of the slot loop issues N independent loads, so N rows cost P round trips
// instead of N * P (measured: the serialised form made the attention prepare ~13 us). Row sums keep the slot order.
template <int N>
__device__ __forceinline__ void mk_rows4_n(const MkPieces& pc, const uint32_t (&rows)[N], float4 (&out)[N]) {
    const float4* base[N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
        base[i] = reinterpret_cast<const float4*>(pc.pieces + ((size_t)(rows[i] >> 4) * pc.P) * 16 + (rows[i] & 15u));
        out[i] = __ldcg(base[i]);
    }
    for (uint32_t q = 1; q < pc.P; ++q) {
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const float4 v = __ldcg(base[i] + (size_t)q * 4);
            out[i].x += v.x; out[i].y += v.y; out[i].z += v.z; out[i].w += v.w;
        }
    }
#pragma unroll
    for (int i = 0; i < N; ++i) { out[i].x = round_bf16(out[i].x); out[i].y = round_bf16(out[i].y); out[i].z = round_bf16(out[i].z); out[i].w = round_bf16(out[i].w); }
}
template <int N>
__device__ __forceinline__ void mk_rows1_n(const MkPieces& pc, const uint32_t (&rows)[N], const bool (&live)[N], float (&out)[N]) {
    const float* base[N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
        base[i] = pc.pieces + ((size_t)(rows[i] >> 4) * pc.P) * 16 + (rows[i] & 15u);
        out[i] = live[i] ? __ldcg(base[i]) : 0.0f;
    }
    for (uint32_t q = 1; q < pc.P; ++q) {
#pragma unroll
        for (int i = 0; i < N; ++i)
            if (live[i]) out[i] += __ldcg(base[i] + (size_t)q * 16);
    }
#pragma unroll
    for (int i = 0; i < N; ++i) out[i] = round_bf16(out[i]);
}
__device__ __forceinline__ uint4 mk_pack8(const float4& a, const float4& b) {
    uint4 o;
    __nv_bfloat162 t;
    t = __floats2bfloat162_rn(a.x, a.y); o.x = *reinterpret_cast<uint32_t*>(&t);
    t = __floats2bfloat162_rn(a.z, a.w); o.y = *reinterpret_cast<uint32_t*>(&t);
    t = __floats2bfloat162_rn(b.x, b.y); o.z = *reinterpret_cast<uint32_t*>(&t);
    t = __floats2bfloat162_rn(b.z, b.w); o.w = *reinterpret_cast<uint32_t*>(&t);
    return o;
}
__device__ __forceinline__ float mk_row1(const MkPieces& pc, uint32_t row) {
    const uint32_t tile = row >> 4, r = row & 15u;
    const float* base = pc.pieces + ((size_t)tile * pc.P) * 16 + r;
    float s = __ldcg(base);
#pragma unroll 4
    for (uint32_t q = 1; q < pc.P; ++q) s += __ldcg(base + (size_t)q * 16);
    return round_bf16(s);
}
// eight consecutive bf16-rounded rows packed as 4 x bf16x2 (row % 8 == 0)
__device__ __forceinline__ uint4 mk_rows8_bf16(const MkPieces& pc, uint32_t row) {
    const float4 a = mk_rows4(pc, row), b = mk_rows4(pc, row + 4);
    uint4 o;
    __nv_bfloat162 t;
    t = __floats2bfloat162_rn(a.x, a.y); o.x = *reinterpret_cast<uint32_t*>(&t);
    t = __floats2bfloat162_rn(a.z, a.w); o.y = *reinterpret_cast<uint32_t*>(&t);
    t = __floats2bfloat162_rn(b.x, b.y); o.z = *reinterpret_cast<uint32_t*>(&t);
    t = __floats2bfloat162_rn(b.z, b.w); o.w = *reinterpret_cast<uint32_t*>(&t);
    return o;
}

// index of the warp range that contains unit u when U units are cut into W ranges with boundaries floor(i * U / W)
__device__ __forceinline__ uint32_t mk_range_of(uint64_t u, uint64_t U, uint64_t W) { return (uint32_t)(((u + 1) * W - 1) / U); }

// Two-level partition of a phase with U units. Level 1: the units are cut into geff = min(gridDim, U) equal contiguous CTA ranges
// (boundaries floor(c * U / geff)): every SM pulls the same number of bytes. Level 2: a CTA's range is cut into min(NCW, length) equal
// contiguous warp ranges. Contiguous warps of ONE CTA therefore share the tiles that straddle warp boundaries: their partial sums meet in
// shared memory (fixed warp order) and a tile leaves one piece per CTA that touches it -- at most two for every BASELINE shape -- instead
// of one per warp (the first version: up to ten pieces per tile, i.e. ten dependent L2 round trips in the consumer's staging).
__device__ __forceinline__ void mk_my_range(uint32_t cta, uint32_t warp, uint32_t ncw, uint32_t U, uint32_t& geff, uint32_t& ub, uint32_t& ue) {
    geff = min(gridDim.x, U);
    ub = ue = 0;
    if (cta >= geff) return;
    // 32-bit arithmetic (the host rejects phases with U * gridDim >= 2^32): a 64-bit division is ~100 instructions and every warp does six per phase
    const uint32_t cb = (cta * U) / geff, ce = ((cta + 1) * U) / geff;
    const uint32_t len = ce - cb, nw = min(ncw, len);
    if (warp >= nw) return;
    ub = cb + (warp * len) / nw;
    ue = cb + ((warp + 1) * len) / nw;
}

__device__ __forceinline__ unsigned long long mk_pack_key(float v, uint32_t i) {   // sampling.cu pack_key: value desc, index asc
    uint32_t u = __float_as_uint(v);
    if (v != v) u = 0u;
    else u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
    return ((unsigned long long)u << 32) | (unsigned long long)(0xFFFFFFFFu - i);
}

// ---------------------------------------------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------------------------------------------
// NPG: nibbles per quantisation group (64: int4 gs64; 128: int4 gs128 / int8 gs64). BITS 4 | 8. NCW consumer warps, S ring stages per warp.
template <int NPG, int BITS, int NCW, int S>
__global__ void __launch_bounds__(NCW * 32, 1) decode_mega_kernel(const MkParams p) {
    constexpr int NCT = NCW * 32;                 // consumer threads
    constexpr int CPM = NPG >= 128 ? 1 : 2;       // quantisation groups per 128-nibble chunk
    constexpr int GPS = 512 / NPG;                // groups per unit
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* scratch = smem;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.scratch_bytes);           // full[NCW][S]
    uint8_t* ring = smem + p.scratch_bytes + ((NCW * S * 8u + 127u) & ~127u);
    __shared__ float red[NCW + 4];
    __shared__ unsigned long long redk[NCW];
    __shared__ unsigned int sm_ticket;
    // phase descriptors are staged in shared memory one phase ahead: every field read of a phase is a shared-memory access instead of a
    // chain of first-touch global loads (measured: ~3 us per phase before this)
    __shared__ MkOp sops[2];
    __shared__ float fragv[NCW][2][16];        // partial tiles of a GEMV phase: [warp][head / tail][row]
    __shared__ uint32_t fragt[NCW][2], fragc[NCW][2];   // their (matrix, tile) key and the first super-chunk they cover

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
        MkWatch watch{reinterpret_cast<unsigned int*>(p.barrier + 1), p.error_flag, false};

    if (tid == 0) {
        for (int i = 0; i < NCW * S; ++i) mk_mbar_init(mk_smem_u32(bars + i), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // ===================================================== consumer warps =====================================================
    const int g = lane >> 2, t = lane & 3;
    const uint32_t gw = warp * gridDim.x + blockIdx.x;      // consecutive work items sit on different SMs
    const uint32_t position = p.state->position;            // prefix length == position of the token being fed
    const uint32_t token = p.token_ids[0];
    uint32_t q = 0;                                          // unit sequence number (matches the producer's)
    const uint32_t full0 = mk_smem_u32(bars + warp * S);
    const uint8_t* ring_w = ring + (size_t)warp * S * MK_STAGE_BYTES;
    const uint32_t ring_s = mk_smem_u32(ring_w);
    // ---- weight stream: every warp feeds its own ring. A prefetch cursor walks this warp's unit sequence over ALL GEMV phases, S units
    // ahead of the unit being consumed: the slot a unit is read from is refilled by the same warp the moment it is done with it (TMA bulk
    // copy completing on the slot's mbarrier; no producer warp, no hand-off latency), and because weights never depend on activations the
    // ring already holds the next phase's first units while the warps stage the next activation row or wait at the grid barrier.
    const uint64_t policy = mk_evict_first_policy();
    uint32_t pf_g = 0, pf_u = 0, pf_ue = 0, pf_split = 0, pf_q = 0;
    const uint8_t* pf_s0 = nullptr;
    const uint8_t* pf_s1 = nullptr;
    bool pf_done = false, pf_hold = false;
    auto pf_load = [&](uint32_t from) {          // first GEMV phase at index >= from in which this warp owns units
        for (pf_g = from; pf_g < p.nstreams; ++pf_g) {
            const uint4 a = __ldg(reinterpret_cast<const uint4*>(p.streams + pf_g));
            const uint4 b = __ldg(reinterpret_cast<const uint4*>(p.streams + pf_g) + 1);
            uint32_t geff;
            mk_my_range(blockIdx.x, warp, NCW, b.x, geff, pf_u, pf_ue);
            if (pf_u < pf_ue) {
                pf_s0 = reinterpret_cast<const uint8_t*>(((uint64_t)a.y << 32) | a.x);
                pf_s1 = reinterpret_cast<const uint8_t*>(((uint64_t)a.w << 32) | a.z);
                pf_split = b.y;
                pf_hold = b.z != 0u;
                return;
            }
        }
        pf_done = true;
    };
    auto pf_issue = [&]() {                      // next unit of the cursor -> ring slot pf_q % S (all lanes keep the cursor, lane 0 issues)
        if (pf_done || pf_hold) return;
        const uint8_t* src = pf_u < pf_split ? pf_s0 + (size_t)pf_u * MK_STAGE_BYTES : pf_s1 + (size_t)(pf_u - pf_split) * MK_STAGE_BYTES;
        const uint32_t s_ = pf_q % S;
        if (lane == 0) {
            mk_mbar_expect_tx(full0 + s_ * 8u, MK_STAGE_BYTES);
            mk_bulk_copy(ring_s + s_ * MK_STAGE_BYTES, src, MK_STAGE_BYTES, full0 + s_ * 8u, policy);
        }
        ++pf_q;
        if (++pf_u >= pf_ue) pf_load(pf_g + 1);
    };
    // called by the phase in front of a held stream once its own HBM loads are in flight: the ring is empty then, refill it
    auto pf_release = [&]() {
        if (!pf_hold) return;
        pf_hold = false;
#pragma unroll 1
        for (int i = 0; i < S; ++i) pf_issue();
    };
    pf_load(0);
    pf_hold = false;                              // nothing precedes the first phase
#pragma unroll 1
    for (int i = 0; i < S; ++i) pf_issue();
    uint32_t magic;
    asm volatile("mov.b32 %0, 0x43004300;" : "=r"(magic));
    const float mult128 = BITS == 4 ? 128.0f : 128.0f * 17.0f;
    const bool lane_has_groups = 2 * t < GPS;
    const int b_chunk = CPM == 2 ? (g >> 1) : g;
    const bool b_lane = CPM == 2 ? ((g & 1) == (t >> 1)) : true;

#define MK_TRACE(k) do { if (p.trace && blockIdx.x == p.trace_cta && tid == 0) p.trace[(size_t)oi * 8 + (k)] = (unsigned long long)clock64(); } while (0)
    constexpr int OPV = (int)(sizeof(MkOp) / 16);
    if (tid < OPV) reinterpret_cast<uint4*>(&sops[0])[tid] = __ldg(reinterpret_cast<const uint4*>(p.ops) + tid);
    mk_bar_sync(2, NCT);
    for (uint32_t oi = 0; oi < p.nops; ++oi) {
        const MkOp& op = sops[oi & 1];
        // next phase's descriptor: its buffer was last read in phase oi - 1, which every thread left before the previous grid barrier;
        // the CTA-wide sync inside this phase's grid barrier publishes it
        if (oi + 1 < p.nops && tid >= NCT - OPV) reinterpret_cast<uint4*>(&sops[(oi + 1) & 1])[tid - (NCT - OPV)] = __ldg(reinterpret_cast<const uint4*>(p.ops + oi + 1) + (tid - (NCT - OPV)));
        MK_TRACE(0);
        switch (op.kind) {
        case MK_GEMV: {
            // ---------------- stage the activation row (every CTA: the row is <= 32 KB) ----------------------------------------
            // scratch: xs [items] uint4 (B-operand order, see matmul.cu) + zero block (4 uint4) + sx [groups] f32
            const uint32_t K = op.k;
            const uint32_t C = op.mat[0].C;                              // all matrices of a phase share K, hence C
            const uint32_t items_all = C * 64u;                           // uint4 items of 8 nibbles
            uint4* xs = reinterpret_cast<uint4*>(scratch);
            float* sx = reinterpret_cast<float*>(scratch + (size_t)(items_all + 4) * 16);
            const uint32_t ngroups = C * GPS;
            constexpr uint32_t EPO = 8;                                   // elements per thread step (one "octet")
            const uint32_t octets_all = BITS == 4 ? items_all : items_all / 2;   // octets covering the padded row
            const uint32_t in_kind = op.in_kind, src_kind = op.src_kind;
            if (tid < 4) xs[items_all + tid] = make_uint4(0, 0, 0, 0);
            if (op.dn_commit) {
                // side job: advance the rolling conv state of the DeltaNet layer whose update phase just finished (every reader of the
                // old state is behind the grid barrier): state[t - 1] = state[t], state[last] = x (conv_update.rs:40-55)
                const uint32_t conv_dim = 2 * op.dn_key_dim + op.dn_value_dim, taps = op.dn_kernel_size - 1;
                for (uint32_t ch = blockIdx.x + gridDim.x * tid; ch < conv_dim; ch += gridDim.x * NCT) {
                    const float x = mk_row1(op.dn_in_pc, ch);
                    float* st = op.dn_conv_state + (size_t)ch * taps;
                    float prev[7];
#pragma unroll
                    for (uint32_t tp = 1; tp < 7; ++tp) prev[tp] = tp < taps ? st[tp] : 0.0f;
#pragma unroll
                    for (uint32_t tp = 1; tp < 7; ++tp)
                        if (tp < taps) st[tp - 1] = prev[tp];
                    st[taps - 1] = x;
                }
            }
            // Rolled loops over the LIVE octets only (an octet = 8 consecutive elements, one thread step). The first version unrolled five
            // octets per thread for every input kind: ~1000 instructions per warp per phase, i.e. ~3 us of pure issue time per phase.
            const uint32_t octs = (K + 7u) / 8u;                           // live octets
            const uint32_t octs_pad = (octs + 31u) & ~31u;                  // warp-uniform loop bound (the group sums use shuffles)
            // B-operand order + per-group activation sums of one octet (`v` = 8 bf16; dead octets carry zeros)
            auto emit = [&](uint32_t o, uint4 v) {
                const __nv_bfloat162* a2 = reinterpret_cast<const __nv_bfloat162*>(&v);
                float part = 0.0f;
                if (BITS == 4) {
                    uint4 out;
                    out.x = __byte_perm(v.x, v.z, 0x5410);
                    out.y = __byte_perm(v.x, v.z, 0x7632);
                    out.z = __byte_perm(v.y, v.w, 0x5410);
                    out.w = __byte_perm(v.y, v.w, 0x7632);
#pragma unroll
                    for (int i = 0; i < 4; ++i) { part += __low2float(a2[i]); part += __high2float(a2[i]); }
                    xs[o] = out;
                } else {
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {
                        const float x0 = __low2float(a2[2 * hh]), x1 = __high2float(a2[2 * hh]), x2 = __low2float(a2[2 * hh + 1]), x3 = __high2float(a2[2 * hh + 1]);
                        __nv_bfloat162 o0 = __floats2bfloat162_rn(x0, x2), o1 = __floats2bfloat162_rn(16.0f * x0, 16.0f * x2);
                        __nv_bfloat162 o2 = __floats2bfloat162_rn(x1, x3), o3 = __floats2bfloat162_rn(16.0f * x1, 16.0f * x3);
                        uint4 out;
                        out.x = *reinterpret_cast<uint32_t*>(&o0); out.y = *reinterpret_cast<uint32_t*>(&o1);
                        out.z = *reinterpret_cast<uint32_t*>(&o2); out.w = *reinterpret_cast<uint32_t*>(&o3);
                        xs[2 * o + hh] = out;
                        part += ((x0 + x1) + x2) + x3;
                    }
                }
                // group sums: a group is NPG nibbles = (BITS == 4 ? NPG : NPG / 2) elements = that / 8 adjacent octets
                constexpr uint32_t OPG = (BITS == 4 ? NPG : NPG / 2) / 8;
#pragma unroll
                for (uint32_t off = 1; off < OPG; off <<= 1) part += __shfl_xor_sync(0xffffffffu, part, off);
                const uint32_t gl = o / OPG;
                if ((o & (OPG - 1)) == 0 && gl < ngroups) sx[gl] = part;
            };
            // padded tail of the row (beyond the last live warp step): zeros
#pragma unroll 1
            for (uint32_t o = octs_pad + tid; o < octets_all; o += NCT) emit(o, make_uint4(0, 0, 0, 0));

            if (in_kind == MK_IN_NORM) {
                // RMSNorm (+ residual add) in two passes over a bf16 copy of the row in shared memory (K <= 8192: behind xs / sx)
                uint4* xrow = reinterpret_cast<uint4*>(scratch + (((size_t)(items_all + 4) * 16 + (size_t)ngroups * 4 + 15u) & ~(size_t)15u));
                float ssq = 0.0f;
#pragma unroll 1
                for (uint32_t o = tid; o < octs; o += NCT) {
                    const uint32_t e0 = o * EPO;
                    uint4 v;
                    if (src_kind == MK_SRC_BF16) v = __ldcg(reinterpret_cast<const uint4*>(op.src_vec + e0));
                    else if (src_kind == MK_SRC_PIECES) {
                        const uint32_t rws[2] = {op.src_row0 + e0, op.src_row0 + e0 + 4};
                        float4 rv[2];
                        mk_rows4_n<2>(op.src_pc, rws, rv);
                        v = mk_pack8(rv[0], rv[1]);
                    } else {
                        // embedding row of the input token (quant_embedding_lookup_kernel / fp_embedding_lookup_kernel arithmetic)
                        const MkEmbed& E = op.embed;
                        float f[8];
                        if (token >= E.vocab) {
#pragma unroll
                            for (int i = 0; i < 8; ++i) f[i] = 0.0f;
                        } else if (E.full_precision) {
                            const uint4 raw = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(E.weights) + (size_t)token * K + e0);
                            const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
                            for (int i = 0; i < 4; ++i) { f[2 * i] = __fmul_rn(__low2float(h2[i]), E.input_scale); f[2 * i + 1] = __fmul_rn(__high2float(h2[i]), E.input_scale); }
                        } else {
                            const uint32_t ng = (K + E.group_size - 1) / E.group_size;
                            const uint32_t gi = e0 / E.group_size;
                            const float scale = bf2f(E.scales[(size_t)token * ng + gi]);
                            float bias;
                            if (E.method == UZU_QMETHOD_SCALE_BIAS) bias = bf2f(E.biases[(size_t)token * ng + gi]);
                            else if (E.method == UZU_QMETHOD_SCALE_ZERO_POINT) {
                                uint32_t zp;
                                if (E.mode == UZU_QMODE_U4) { const uint8_t pk = E.zero_points[(size_t)token * ((ng + 1) / 2) + gi / 2]; zp = (gi & 1) ? (pk >> 4) : (pk & 15u); }
                                else zp = E.zero_points[(size_t)token * ng + gi];
                                bias = __fmul_rn(-scale, (float)zp);
                            } else bias = __fmul_rn(-scale, E.mode == UZU_QMODE_U4 ? 8.0f : 128.0f);
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                int32_t qv;
                                if (E.mode == UZU_QMODE_U4) { const uint8_t pk = E.weights[(size_t)token * (K / 2) + (e0 + i) / 2]; qv = (i & 1) ? (pk >> 4) : (pk & 15); }
                                else if (E.mode == UZU_QMODE_I8) qv = reinterpret_cast<const int8_t*>(E.weights)[(size_t)token * K + e0 + i];
                                else qv = E.weights[(size_t)token * K + e0 + i];
                                f[i] = __fmul_rn(__fadd_rn(__fmul_rn(scale, (float)qv), bias), E.input_scale);
                            }
                        }
                        __nv_bfloat162 tt;
                        tt = __floats2bfloat162_rn(f[0], f[1]); v.x = *reinterpret_cast<uint32_t*>(&tt);
                        tt = __floats2bfloat162_rn(f[2], f[3]); v.y = *reinterpret_cast<uint32_t*>(&tt);
                        tt = __floats2bfloat162_rn(f[4], f[5]); v.z = *reinterpret_cast<uint32_t*>(&tt);
                        tt = __floats2bfloat162_rn(f[6], f[7]); v.w = *reinterpret_cast<uint32_t*>(&tt);
                    }
                    __nv_bfloat162* a2 = reinterpret_cast<__nv_bfloat162*>(&v);
                    if (op.norm_residual_add) {
                        const uint4 sb = __ldcg(reinterpret_cast<const uint4*>(op.shortcut_in + e0));
                        const __nv_bfloat162* b2 = reinterpret_cast<const __nv_bfloat162*>(&sb);
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            a2[i] = __floats2bfloat162_rn(__fadd_rn(__low2float(a2[i]), __low2float(b2[i])), __fadd_rn(__high2float(a2[i]), __high2float(b2[i])));
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float lo = __low2float(a2[i]), hi = __high2float(a2[i]);
                        ssq = __fadd_rn(ssq, __fmul_rn(lo, lo));
                        ssq = __fadd_rn(ssq, __fmul_rn(hi, hi));
                    }
                    if (op.shortcut_out && blockIdx.x == 0) *reinterpret_cast<uint4*>(op.shortcut_out + e0) = v;
                    xrow[o] = v;
                }
                ssq = warp_sum(ssq);
                if (lane == 0) red[warp] = ssq;
                mk_bar_sync(2, NCT);
                float tot = 0.0f;
#pragma unroll
                for (int w_ = 0; w_ < NCW; ++w_) tot = __fadd_rn(tot, red[w_]);
                const float rms_inv = __frcp_rn(__fsqrt_rn(__fadd_rn(__fdiv_rn(tot, (float)K), op.norm_eps)));
#pragma unroll 1
                for (uint32_t o = tid; o < octs_pad; o += NCT) {
                    uint4 v = make_uint4(0, 0, 0, 0);
                    if (o < octs) {
                        const uint32_t e0 = o * EPO;
                        v = xrow[o];                                       // written by this very thread in pass 1
                        __nv_bfloat162* a2 = reinterpret_cast<__nv_bfloat162*>(&v);
                        const float4 s0 = __ldg(reinterpret_cast<const float4*>(op.norm_scales + e0)), s1 = __ldg(reinterpret_cast<const float4*>(op.norm_scales + e0 + 4));
                        const float sc[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            float x[2] = {__low2float(a2[i]), __high2float(a2[i])};
#pragma unroll
                            for (int h = 0; h < 2; ++h) {
                                const float normalized = __fmul_rn(x[h], rms_inv);
                                const float so = __fadd_rn(sc[2 * i + h], op.norm_scale_offset);
                                if (op.norm_full_layer) x[h] = __fmul_rn(normalized, so);
                                else x[h] = __fmul_rn(round_bf16(normalized), round_bf16(so));
                            }
                            a2[i] = __floats2bfloat162_rn(x[0], x[1]);
                        }
                    }
                    emit(o, v);
                }
            } else {
#pragma unroll 1
                for (uint32_t o = tid; o < octs_pad; o += NCT) {
                    const uint32_t e0 = o * EPO;
                    const bool live = o < octs;                            // warp-uniform whenever K % 256 == 0 (required for MK_IN_DELTA)
                    uint4 v = make_uint4(0, 0, 0, 0);
                    if (live && in_kind == MK_IN_PLAIN) {
                        v = __ldcg(reinterpret_cast<const uint4*>(op.src_vec + e0));
                    } else if (live && in_kind == MK_IN_GATED) {
                        // GatedActMul folded into the consumer: rows [0, F) of the fused up projection are `up`, rows [F, 2F) `gate`
                        const uint32_t rws[4] = {e0, e0 + 4, K + e0, K + e0 + 4};
                        float4 rv[4];
                        mk_rows4_n<4>(op.gated_pc, rws, rv);
                        const float uu[8] = {rv[0].x, rv[0].y, rv[0].z, rv[0].w, rv[1].x, rv[1].y, rv[1].z, rv[1].w};
                        const float gg[8] = {rv[2].x, rv[2].y, rv[2].z, rv[2].w, rv[3].x, rv[3].y, rv[3].z, rv[3].w};
                        float hh[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) hh[i] = __fmul_rn(uu[i], round_bf16(act_f32_nofma(op.gated_act, gg[i])));
                        __nv_bfloat162 tt;
                        tt = __floats2bfloat162_rn(hh[0], hh[1]); v.x = *reinterpret_cast<uint32_t*>(&tt);
                        tt = __floats2bfloat162_rn(hh[2], hh[3]); v.y = *reinterpret_cast<uint32_t*>(&tt);
                        tt = __floats2bfloat162_rn(hh[4], hh[5]); v.z = *reinterpret_cast<uint32_t*>(&tt);
                        tt = __floats2bfloat162_rn(hh[6], hh[7]); v.w = *reinterpret_cast<uint32_t*>(&tt);
                    } else if (live && in_kind == MK_IN_SIGMOID) {
                        v = __ldcg(reinterpret_cast<const uint4*>(op.src_vec + e0));
                        __nv_bfloat162* a2 = reinterpret_cast<__nv_bfloat162*>(&v);
                        const uint32_t rws[2] = {e0, e0 + 4};
                        float4 rv[2];
                        mk_rows4_n<2>(op.gate_pc, rws, rv);
                        const float gt[8] = {rv[0].x, rv[0].y, rv[0].z, rv[0].w, rv[1].x, rv[1].y, rv[1].z, rv[1].w};
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const float m0 = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-gt[2 * i])));
                            const float m1 = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-gt[2 * i + 1])));
                            a2[i] = __floats2bfloat162_rn(__fmul_rn(__low2float(a2[i]), m0), __fmul_rn(__high2float(a2[i]), m1));
                        }
                    } else if (live && in_kind == MK_IN_DELTA) {
                        // raw DeltaNet output -> RMS over the head (Dv elements = Dv / 8 adjacent threads) * norm_weight * silu(z)
                        __nv_bfloat162* a2 = reinterpret_cast<__nv_bfloat162*>(&v);
                        const float4 r0 = __ldcg(reinterpret_cast<const float4*>(op.dn_raw + e0)), r1 = __ldcg(reinterpret_cast<const float4*>(op.dn_raw + e0 + 4));
                        const float rv[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
                        float ss = 0.0f;
#pragma unroll
                        for (int i = 0; i < 8; ++i) ss = __fadd_rn(ss, __fmul_rn(rv[i], rv[i]));
                        const uint32_t Dv = op.dn_head_v_dim;
                        for (uint32_t off = 1; off < Dv / 8; off <<= 1) ss = __fadd_rn(ss, __shfl_xor_sync(0xffffffffu, ss, off));
                        const float inv_rms = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(__fdiv_rn(ss, (float)Dv), op.dn_eps)));
                        const uint32_t i0 = e0 % Dv;
                        const uint32_t rws[2] = {op.dn_z_row0 + e0, op.dn_z_row0 + e0 + 4};
                        float4 zr[2];
                        mk_rows4_n<2>(op.dn_z_pc, rws, zr);
                        const float zv[8] = {zr[0].x, zr[0].y, zr[0].z, zr[0].w, zr[1].x, zr[1].y, zr[1].z, zr[1].w};
                        float ov[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const float nw = __ldg(op.dn_norm_weight + i0 + i);
                            const float zs = act_f32_nofma(UZU_ACT_SILU, zv[i]);
                            ov[i] = __fmul_rn(__fmul_rn(__fmul_rn(rv[i], inv_rms), nw), zs);
                        }
#pragma unroll
                        for (int i = 0; i < 4; ++i) a2[i] = __floats2bfloat162_rn(ov[2 * i], ov[2 * i + 1]);
                    }
                    emit(o, v);
                }
            }
            mk_bar_sync(2, NCT);
            MK_TRACE(1);

            // ---------------- this warp's contiguous range of units -------------------------------------------------------------
            const uint64_t U = op.units;
            uint32_t geff, u, ue;
            mk_my_range(blockIdx.x, warp, NCW, op.units, geff, u, ue);
            uint32_t nfrag = 0;                                   // partial tiles of this warp (head and / or tail of its range): <= 2
            if (lane < 2) fragt[warp][lane] = 0xffffffffu;
            while (u < ue) {
                const int mi = (op.nmat > 1 && u >= op.mat[1].unit0) ? 1 : 0;
                const MkMat M = op.mat[mi];               // by value: the hot loop must not re-read the descriptor
                const uint32_t v0 = u - M.unit0;
                uint32_t tile = v0 / C, c = v0 % C;
                uint32_t c_first = c;
                const uint32_t mend = min(ue, M.unit0 + M.tiles * C);
                float acc0 = 0.0f, acc1 = 0.0f, acc2 = 0.0f, acc3 = 0.0f;
                for (; u < mend; ++u, ++q) {
                    const uint32_t s = q % S, ph = (q / S) & 1u;
                    mk_wait(full0 + s * 8u, ph, watch, 0x300u);
                    const uint4* sw = reinterpret_cast<const uint4*>(ring_w + (size_t)s * MK_STAGE_BYTES) + lane;
                    const uint32_t* swd = reinterpret_cast<const uint32_t*>(ring_w + (size_t)s * MK_STAGE_BYTES + 4096) + lane;
                    // four independent accumulator fragments: the 32 MMAs of a unit form four dependent chains of eight, not two of sixteen
                    float d[4] = {0.0f, 0.0f, 0.0f, 0.0f}, d2[4] = {0.0f, 0.0f, 0.0f, 0.0f}, d3[4] = {0.0f, 0.0f, 0.0f, 0.0f}, d4[4] = {0.0f, 0.0f, 0.0f, 0.0f};
                    const uint32_t c0 = c * 4u;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const uint4 va = sw[j * 32], vb = sw[(4 + j) * 32];
                        const uint32_t wav[4] = {va.x, va.y, va.z, va.w};
                        const uint32_t wbv[4] = {vb.x, vb.y, vb.z, vb.w};
                        const bool feeds = b_lane && b_chunk == j;
                        const uint4* xrow = feeds ? xs + ((size_t)(c0 + j) * 4 + t) * 4 : xs + items_all;
#pragma unroll
                        for (int w_ = 0; w_ < 4; ++w_) {
                            const uint4 xb = xrow[w_];
                            const uint32_t a0 = mk_nib_pair(wav[w_], 0, magic), a1 = mk_nib_pair(wav[w_], 4, magic), a2 = mk_nib_pair(wav[w_], 8, magic), a3 = mk_nib_pair(wav[w_], 12, magic);
                            const uint32_t b0 = mk_nib_pair(wbv[w_], 0, magic), b1 = mk_nib_pair(wbv[w_], 4, magic), b2 = mk_nib_pair(wbv[w_], 8, magic), b3 = mk_nib_pair(wbv[w_], 12, magic);
                            if (w_ & 1) {
                                mk_mma_16816(d3, a0, b0, a1, b1, xb.x, xb.y);
                                mk_mma_16816(d4, a2, b2, a3, b3, xb.z, xb.w);
                            } else {
                                mk_mma_16816(d, a0, b0, a1, b1, xb.x, xb.y);
                                mk_mma_16816(d2, a2, b2, a3, b3, xb.z, xb.w);
                            }
                        }
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i) d[i] = (d[i] + d3[i]) + (d2[i] + d4[i]);
                    if (lane_has_groups) {
                        // this lane's two D columns (2t, 2t + 1) hold groups gi + 2t, gi + 2t + 1 of rows g (d0, d1) and g + 8 (d2, d3)
                        const uint32_t gi = c * GPS + 2 * t;
                        const float2 sxv = *reinterpret_cast<const float2*>(sx + gi);
                        const uint32_t bsa = swd[0], bsb = swd[32], bca = swd[64], bcb = swd[96];
                        const float sa0 = __uint_as_float(bsa << 16), sa1 = __uint_as_float(bsa & 0xffff0000u);
                        const float sb0 = __uint_as_float(bsb << 16), sb1 = __uint_as_float(bsb & 0xffff0000u);
                        const float ca0 = __uint_as_float(bca << 16), ca1 = __uint_as_float(bca & 0xffff0000u);
                        const float cb0 = __uint_as_float(bcb << 16), cb1 = __uint_as_float(bcb & 0xffff0000u);
                        if (!M.bias_form) {       // value = s * (q - zp): s * (d - (128m + zp) * Sx)
                            acc0 += sa0 * (d[0] - (ca0 + mult128) * sxv.x);
                            acc1 += sa1 * (d[1] - (ca1 + mult128) * sxv.y);
                            acc2 += sb0 * (d[2] - (cb0 + mult128) * sxv.x);
                            acc3 += sb1 * (d[3] - (cb1 + mult128) * sxv.y);
                        } else {                  // value = s * q + b
                            acc0 += sa0 * (d[0] - mult128 * sxv.x) + ca0 * sxv.x;
                            acc1 += sa1 * (d[1] - mult128 * sxv.y) + ca1 * sxv.y;
                            acc2 += sb0 * (d[2] - mult128 * sxv.x) + cb0 * sxv.x;
                            acc3 += sb1 * (d[3] - mult128 * sxv.y) + cb1 * sxv.y;
                        }
                    }
                    // every lane is done with slot s: refill it with this warp's unit S positions ahead (generic-proxy reads before the
                    // async-proxy write: warp sync + proxy fence)
                    __syncwarp();
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    pf_issue();
                    // ---- end of a tile or of the range ---------------------------------------------------------------------------
                    if (c + 1 == C || u + 1 == mend) {
                        float ra = acc0 + acc1, rb = acc2 + acc3;
                        ra += __shfl_xor_sync(0xffffffffu, ra, 1);
                        rb += __shfl_xor_sync(0xffffffffu, rb, 1);
                        ra += __shfl_xor_sync(0xffffffffu, ra, 2);
                        rb += __shfl_xor_sync(0xffffffffu, rb, 2);
                        if (c_first == 0 && c + 1 == C) {
                            // the whole tile was summed by this warp: it lies inside this CTA's range -> slot 0, final
                            float* dst = M.pieces + ((size_t)tile * M.P) * 16;
                            if (t == 0) { dst[g] = ra; dst[g + 8] = rb; }
                        } else {
                            // partial tile: meets the neighbouring warps' parts in shared memory after the CTA-wide sync below
                            if (t == 0) { fragv[warp][nfrag][g] = ra; fragv[warp][nfrag][g + 8] = rb; }
                            if (lane == 0) {
                                fragt[warp][nfrag] = ((uint32_t)mi << 28) | tile;
                                fragc[warp][nfrag] = c_first;
                            }
                            ++nfrag;
                        }
                        acc0 = acc1 = acc2 = acc3 = 0.0f;
                        c_first = 0;
                    }
                    if (++c == C) { c = 0; ++tile; }
                }
            }
            mk_bar_sync(2, NCT);
            // ---------------- combine the partial tiles of this CTA (fixed warp order), one piece per (tile, CTA) ----------------------
            // A partial tile is owned by the first warp of this CTA that touches it: a fragment that starts at the tile's first unit, or
            // warp 0's head fragment (the earlier units belong to the previous CTA). The following warps' first fragments continue it.
            for (uint32_t f = 0; f < nfrag; ++f) {
                const uint32_t key = fragt[warp][f];
                const bool owner = fragc[warp][f] == 0 || warp == 0;
                if (!owner) continue;
                if (lane < 16) {
                    float sum = fragv[warp][f][lane];
                    for (int w2 = warp + 1; w2 < NCW && fragt[w2][0] == key; ++w2) sum += fragv[w2][0][lane];
                    const int mi = (int)(key >> 28);
                    const uint32_t tile = key & 0x0fffffffu;
                    const MkMat& M = op.mat[mi];
                    // slot = index of this CTA among the CTAs that touch the tile
                    uint32_t slot = 0;
                    if (fragc[warp][f] != 0) slot = blockIdx.x - mk_range_of((uint64_t)M.unit0 + (uint64_t)tile * C, U, geff);
                    M.pieces[((size_t)tile * M.P + slot) * 16 + lane] = sum;
                }
            }
            break;
        }
        case MK_ATTN: {
            // QKVNorm + RoPE + KV append (qkv_norm.rs:36-76, attention_prepare.rs:34-126) folded into decode attention over keys
            // [0, position] (attention_single_pass.rs:49-126 arithmetic: bf16 q pre-scaled in f32, expf, f32 accumulators; plain causal
            // decode: every cached key is visible). CTAs per kv head = gridDim / Hkv; a CTA owns one contiguous key range of one kv head
            // for all G query heads; lane groups of D / 8 lanes own one key row each. The CTA that owns the LAST range appends the new
            // K / V rows to the cache before its key loop (it is the only reader of that row).
            const uint32_t Hq = op.num_q_heads, Hkv = op.num_kv_heads, D = op.head_dim, G = Hq / Hkv;
            const uint32_t seq = position + 1;
            const uint32_t cph = gridDim.x / Hkv;
            const uint32_t LPK = D / 8, KPW = 32 / LPK;
            const uint32_t step_keys = NCW * KPW;
            // keys per CTA part: whole CTA steps, at least four per warp (one full batch of loads in flight): short contexts use few CTAs
            // per kv head instead of many tiny parts whose merge would dominate
            const uint32_t kp = max(((seq + cph - 1) / cph + step_keys - 1) / step_keys, 4u) * step_keys;
            const uint32_t nparts = (seq + kp - 1) / kp;
            const uint32_t kvh = blockIdx.x / cph, part = blockIdx.x % cph;
            if (blockIdx.x >= cph * Hkv || part >= nparts) { pf_release(); break; }
            float* sq = reinterpret_cast<float*>(scratch);               // [G][D] scaled queries, later reused by the warp merge
            {
                // warp w < G: query head kvh * G + w; the last part's warps G and G + 1: the new key / value rows
                const bool last_part = part == nparts - 1;
                const uint32_t role = (uint32_t)warp;
                if (role < G || (last_part && role < G + 2)) {
                    const bool is_q = role < G, is_k = role == G;
                    const uint32_t h = is_q ? kvh * G + role : (is_k ? Hq + kvh : Hq + Hkv + kvh);
                    const bool normed = (is_q && op.qnorm_present) || (is_k && op.knorm_present);
                    const float* nscales = is_q ? op.qnorm_scales : op.knorm_scales;
                    const float eps = is_q ? op.qnorm_eps : op.knorm_eps, offs = is_q ? op.qnorm_offset : op.knorm_offset;
                    const uint32_t full = is_q ? op.qnorm_full_layer : op.knorm_full_layer, has_sc = is_q ? op.qnorm_has_scales : op.knorm_has_scales;
                    const bool roped = op.rope_cos != nullptr && (is_q || is_k);
                    const uint32_t rd = op.rope_dim, half = rd / 2;
                    constexpr int EPT = 8;                               // D <= 256: elements dd = lane + 32 i
                    float ev[EPT], pv[EPT], cs[EPT], sn[EPT];
                    const float* cosr = roped ? op.rope_cos + (size_t)position * rd : nullptr;
                    const float* sinr = roped ? op.rope_sin + (size_t)position * rd : nullptr;
#pragma unroll
                    for (int i = 0; i < EPT; ++i) {                      // the table rows do not depend on this token's data: request them first
                        const uint32_t dd = lane + 32 * i;
                        cs[i] = (roped && dd < rd) ? __ldg(cosr + dd) : 1.0f;
                        sn[i] = (roped && dd < rd) ? __ldg(sinr + dd) : 0.0f;
                    }
                    // every load of this head first (one L2 round trip), then the arithmetic
                    {
                        uint32_t rws[2 * EPT];
                        bool live[2 * EPT];
                        float got[2 * EPT];
#pragma unroll
                        for (int i = 0; i < EPT; ++i) {
                            const uint32_t dd = lane + 32 * i;
                            live[i] = dd < D;
                            rws[i] = h * D + (live[i] ? dd : 0u);
                            live[EPT + i] = live[i] && roped && dd < rd;
                            rws[EPT + i] = h * D + (live[EPT + i] ? (dd < half ? dd + half : dd - half) : 0u);
                        }
                        mk_rows1_n<2 * EPT>(op.qkv_pc, rws, live, got);
#pragma unroll
                        for (int i = 0; i < EPT; ++i) { ev[i] = got[i]; pv[i] = got[EPT + i]; }
                    }
                    float rms = 0.0f;
                    if (normed) {
                        float total = 0.0f;
#pragma unroll
                        for (int i = 0; i < EPT; ++i) total = __fadd_rn(total, __fmul_rn(ev[i], ev[i]));   // lane-strided, then the xor tree: the standalone kernel's order
                        total = warp_sum(total);
                        rms = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(__fdiv_rn(total, (float)D), eps)));
                    }
                    auto normalise = [&](float r, uint32_t dd) {
                        const float normalized = __fmul_rn(r, rms);
                        if (!has_sc) return round_bf16(normalized);
                        if (full) return round_bf16(__fmul_rn(normalized, __fadd_rn(nscales[dd], offs)));
                        return round_bf16(__fmul_rn(round_bf16(normalized), round_bf16(__fadd_rn(nscales[dd], offs))));
                    };
#pragma unroll
                    for (int i = 0; i < EPT; ++i) {
                        const uint32_t dd = lane + 32 * i;
                        if (dd >= D) continue;
                        float e = ev[i];
                        if (normed) e = normalise(e, dd);
                        if (roped && dd < rd) {
                            float pr = pv[i];
                            if (normed) pr = normalise(pr, dd < half ? dd + half : dd - half);
                            const float signed_p = dd < half ? -pr : pr;
                            e = round_bf16(__fadd_rn(__fmul_rn(e, cs[i]), __fmul_rn(signed_p, sn[i])));
                        }
                        if (is_q) sq[role * D + dd] = __fmul_rn(op.attn_scale, e);
                        else if (is_k) op.keys[((size_t)position * Hkv + kvh) * D + dd] = f2bf(e);
                        else op.values[((size_t)position * Hkv + kvh) * D + dd] = f2bf(e);
                    }
                }
            }
            mk_bar_sync(2, NCT);
            MK_TRACE(1);
            // scratch: sq [G][D] | sc [G][kp] scores -> probabilities | sst [2 G] (max, sum) | so [ceil(NCW / 2)][G][D] partial outputs
            const uint32_t sub = lane / LPK, li = lane % LPK, d0 = li * 8;
            const uint32_t kbeg = part * kp, kend = min(seq, kbeg + kp), nkeys = kend - kbeg;
            float* sc = sq + G * D;
            float* sst = sc + G * kp;
            float* so = sst + 16;                                     // keeps the float4 accesses below 16-byte aligned (G * kp is a multiple of 4)
            constexpr int MAXG = 4;       // query heads per kv head held in registers (the host rejects larger groups)
            float qf[MAXG][8];
#pragma unroll
            for (int h = 0; h < MAXG; ++h) {
#pragma unroll
                for (int e = 0; e < 8; ++e) qf[h][e] = 0.0f;
                if ((uint32_t)h < G) {
                    const float4 qa = *reinterpret_cast<const float4*>(sq + h * D + d0), qb = *reinterpret_cast<const float4*>(sq + h * D + d0 + 4);
                    qf[h][0] = qa.x; qf[h][1] = qa.y; qf[h][2] = qa.z; qf[h][3] = qa.w; qf[h][4] = qb.x; qf[h][5] = qb.y; qf[h][6] = qb.z; qf[h][7] = qb.w;
                }
            }
            const __nv_bfloat16* kbase = op.keys + (size_t)kvh * D + d0;
            const __nv_bfloat16* vbase = op.values + (size_t)kvh * D + d0;
            const size_t rstride = (size_t)Hkv * D;
            constexpr int UU = 4;     // key rows in flight per lane group (all loads of a batch are issued before any is consumed)
            // ---- (1) scores: s[h][key] = q_h . k_key for every key of this CTA's range (one softmax per CTA instead of one per key and
            // lane: the per-key online softmax of the first version was instruction-bound, 16 lanes repeating the same expf). The loop
            // bounds are warp-uniform (the shuffles need every lane); rows past the end are masked.
            uint4 v0r[UU];                                            // V rows of the first batch travel while the scores are computed
#pragma unroll
            for (int u_ = 0; u_ < UU; ++u_) {
                const uint32_t ki = kbeg + warp * KPW + sub + u_ * step_keys;
                v0r[u_] = ki < kend ? __ldcg(reinterpret_cast<const uint4*>(vbase + (size_t)ki * rstride)) : make_uint4(0, 0, 0, 0);
            }
            bool released = false;
            for (uint32_t kb = kbeg + warp * KPW; kb < kend; kb += step_keys * UU) {
                uint4 kr[UU];
#pragma unroll
                for (int u_ = 0; u_ < UU; ++u_) {
                    const uint32_t ki = kb + sub + u_ * step_keys;
                    kr[u_] = ki < kend ? __ldcg(reinterpret_cast<const uint4*>(kbase + (size_t)ki * rstride)) : make_uint4(0, 0, 0, 0);
                }
                if (!released) { pf_release(); released = true; }     // the KV rows are on their way: now the next phase's weights may queue
#pragma unroll
                for (int u_ = 0; u_ < UU; ++u_) {
                    if (kb + u_ * step_keys >= kend) break;           // warp-uniform: no lane group has a key in this slot
                    const uint32_t ki = kb + sub + u_ * step_keys;
                    float kf[8];
                    const __nv_bfloat162* k2 = reinterpret_cast<const __nv_bfloat162*>(&kr[u_]);
#pragma unroll
                    for (int e = 0; e < 4; ++e) { kf[2 * e] = __low2float(k2[e]); kf[2 * e + 1] = __high2float(k2[e]); }
#pragma unroll
                    for (int h = 0; h < MAXG; ++h) {
                        if ((uint32_t)h >= G) break;
                        float sdot = 0.0f;
#pragma unroll
                        for (int e = 0; e < 8; ++e) sdot = fmaf(qf[h][e], kf[e], sdot);
                        for (uint32_t off = LPK / 2; off > 0; off >>= 1) sdot += __shfl_xor_sync(0xffffffffu, sdot, off);
                        if (li == 0 && ki < kend) sc[h * kp + (ki - kbeg)] = sdot;
                    }
                }
            }
            if (!released) pf_release();
            mk_bar_sync(2, NCT);
            MK_TRACE(2);
            // ---- (2) softmax statistics of the range, one warp per head: max, p = expf(s - max) in place, sum
            if ((uint32_t)warp < G) {
                float* row = sc + warp * kp;
                float mloc = -INFINITY;
                for (uint32_t kl = lane; kl < nkeys; kl += 32) mloc = fmaxf(mloc, row[kl]);
                const float M = warp_max(mloc);
                float lsum = 0.0f;
                for (uint32_t kl = lane; kl < nkeys; kl += 32) {
                    const float pv_ = expf(row[kl] - M);
                    row[kl] = pv_;
                    lsum += pv_;
                }
                lsum = warp_sum(lsum);
                if (lane == 0) { sst[2 * warp] = M; sst[2 * warp + 1] = lsum; }
            }
            mk_bar_sync(2, NCT);
            // ---- (3) o[h] = sum_key p[h][key] * v_key (no rescaling: every warp uses the range's max)
            float o[MAXG][8];
#pragma unroll
            for (int h = 0; h < MAXG; ++h)
#pragma unroll
                for (int e = 0; e < 8; ++e) o[h][e] = 0.0f;
            bool first_batch = true;
            for (uint32_t kb = kbeg + warp * KPW; kb < kend; kb += step_keys * UU) {
                uint4 vr[UU];
#pragma unroll
                for (int u_ = 0; u_ < UU; ++u_) {
                    const uint32_t ki = kb + sub + u_ * step_keys;
                    if (first_batch) vr[u_] = v0r[u_];
                    else vr[u_] = ki < kend ? __ldcg(reinterpret_cast<const uint4*>(vbase + (size_t)ki * rstride)) : make_uint4(0, 0, 0, 0);
                }
                first_batch = false;
#pragma unroll
                for (int u_ = 0; u_ < UU; ++u_) {
                    if (kb + u_ * step_keys >= kend) break;           // warp-uniform
                    const uint32_t ki = kb + sub + u_ * step_keys;
                    if (ki >= kend) continue;
                    float vf[8];
                    const __nv_bfloat162* v2 = reinterpret_cast<const __nv_bfloat162*>(&vr[u_]);
#pragma unroll
                    for (int e = 0; e < 4; ++e) { vf[2 * e] = __low2float(v2[e]); vf[2 * e + 1] = __high2float(v2[e]); }
#pragma unroll
                    for (int h = 0; h < MAXG; ++h) {
                        if ((uint32_t)h >= G) break;
                        const float pv_ = sc[h * kp + (ki - kbeg)];
#pragma unroll
                        for (int e = 0; e < 8; ++e) o[h][e] = fmaf(pv_, vf[e], o[h][e]);
                    }
                }
            }
            // lane groups of a warp (plain sums), then the warps of the CTA through shared memory in two rounds (half the buffer)
#pragma unroll
            for (int h = 0; h < MAXG; ++h) {
                if ((uint32_t)h >= G) break;
                for (uint32_t off = LPK; off < 32; off <<= 1) {
#pragma unroll
                    for (int e = 0; e < 8; ++e) o[h][e] += __shfl_xor_sync(0xffffffffu, o[h][e], off);
                }
            }
            constexpr int HALF = (NCW + 1) / 2;
            auto park = [&](int slot) {
                if (sub == 0) {
#pragma unroll
                    for (int h = 0; h < MAXG; ++h) {
                        if ((uint32_t)h >= G) break;
                        float* dst = so + ((size_t)slot * G + h) * D + d0;
                        *reinterpret_cast<float4*>(dst) = make_float4(o[h][0], o[h][1], o[h][2], o[h][3]);
                        *reinterpret_cast<float4*>(dst + 4) = make_float4(o[h][4], o[h][5], o[h][6], o[h][7]);
                    }
                }
            };
            if (warp >= HALF) park(warp - HALF);
            mk_bar_sync(2, NCT);
            if (warp < HALF && warp + HALF < NCW && sub == 0) {
#pragma unroll
                for (int h = 0; h < MAXG; ++h) {
                    if ((uint32_t)h >= G) break;
                    const float* src = so + ((size_t)warp * G + h) * D + d0;
                    const float4 a4 = *reinterpret_cast<const float4*>(src), b4 = *reinterpret_cast<const float4*>(src + 4);
                    o[h][0] += a4.x; o[h][1] += a4.y; o[h][2] += a4.z; o[h][3] += a4.w; o[h][4] += b4.x; o[h][5] += b4.y; o[h][6] += b4.z; o[h][7] += b4.w;
                }
            }
            mk_bar_sync(2, NCT);
            if (warp < HALF) park(warp);
            mk_bar_sync(2, NCT);
            float* pbase = op.attn_part + ((size_t)kvh * cph + part) * G * (D + 2);
            for (uint32_t idx = tid; idx < G * D; idx += NCT) {
                const uint32_t h = idx / D, dd = idx % D;
                float O = 0.0f;
#pragma unroll
                for (int w_ = 0; w_ < HALF; ++w_) O += so[((size_t)w_ * G + h) * D + dd];
                if (nparts == 1) {
                    op.attn_out[((size_t)(kvh * G + h)) * D + dd] = f2bf(O / sst[2 * h + 1]);      // a single part: no cross-CTA merge
                } else {
                    pbase[(size_t)h * (D + 2) + dd] = O;
                    if (dd == 0) { pbase[(size_t)h * (D + 2) + D] = sst[2 * h]; pbase[(size_t)h * (D + 2) + D + 1] = sst[2 * h + 1]; }
                }
            }
            if (nparts == 1) break;
            MK_TRACE(3);
            // the last CTA of this kv head merges the parts (fixed order) and writes the bf16 attention output
            __threadfence();
            mk_bar_sync(2, NCT);
            if (tid == 0) sm_ticket = atomicAdd(op.attn_tickets + kvh, 1u);
            mk_bar_sync(2, NCT);
            MK_TRACE(4);
            if (sm_ticket == nparts - 1) {
                __threadfence();
                const float* hb = op.attn_part + (size_t)kvh * cph * G * (D + 2);
                float* sf = reinterpret_cast<float*>(scratch);              // [G][nparts] weight of every part: exp(m_p - M) / L
                // step 1: warp h computes the weights of head h (parts strided over the lanes, fixed combination order)
                if ((uint32_t)warp < G) {
                    const uint32_t h = warp;
                    float mloc = -INFINITY;
                    for (uint32_t pi = lane; pi < nparts; pi += 32) mloc = fmaxf(mloc, __ldcg(hb + ((size_t)pi * G + h) * (D + 2) + D));
                    const float M = warp_max(mloc);
                    float lsum = 0.0f;
                    for (uint32_t pi = lane; pi < nparts; pi += 32) {
                        const float* pp = hb + ((size_t)pi * G + h) * (D + 2);
                        const float f = expf(__ldcg(pp + D) - M);
                        sf[h * nparts + pi] = f;
                        lsum += __ldcg(pp + D + 1) * f;
                    }
                    const float L = warp_sum(lsum);
                    if (lane == 0) sf[G * nparts + h] = L;
                }
                mk_bar_sync(2, NCT);
                // step 2: every output element: independent loads over the parts
                for (uint32_t idx = tid; idx < G * D; idx += NCT) {
                    const uint32_t h = idx / D, dd = idx % D;
                    float O = 0.0f;
#pragma unroll 8
                    for (uint32_t pi = 0; pi < nparts; ++pi) O += __ldcg(hb + ((size_t)pi * G + h) * (D + 2) + dd) * sf[h * nparts + pi];
                    op.attn_out[((size_t)(kvh * G + h)) * D + dd] = f2bf(O / sf[G * nparts + h]);
                }
                if (tid == 0) op.attn_tickets[kvh] = 0u;
            }
            break;
        }
        case MK_ACT: {
            // GatedActMul (gated_act_mul/mod.rs:5-12): hidden[j] = bf16(bf16(up_j) * bf16(act(bf16(gate_j)))), rows [0, F) up, [F, 2F) gate
            const uint32_t F = op.act_dim;
            for (uint32_t j = (blockIdx.x + gridDim.x * tid) * 4u; j < F; j += gridDim.x * NCT * 4u) {
                const uint32_t rws[2] = {j, F + j};
                float4 rv[2];
                mk_rows4_n<2>(op.up_pc, rws, rv);
                const float4 up = rv[0], gt = rv[1];
                const float m0 = round_bf16(act_f32_nofma(op.act_type, gt.x)), m1 = round_bf16(act_f32_nofma(op.act_type, gt.y));
                const float m2 = round_bf16(act_f32_nofma(op.act_type, gt.z)), m3 = round_bf16(act_f32_nofma(op.act_type, gt.w));
                __nv_bfloat162 a = __floats2bfloat162_rn(__fmul_rn(up.x, m0), __fmul_rn(up.y, m1)), b = __floats2bfloat162_rn(__fmul_rn(up.z, m2), __fmul_rn(up.w, m3));
                uint2 out;
                out.x = *reinterpret_cast<uint32_t*>(&a); out.y = *reinterpret_cast<uint32_t*>(&b);
                *reinterpret_cast<uint2*>(op.hidden + j) = out;
            }
            break;
        }
        case MK_DN_UPDATE: {
            // DeltaNetConvUpdate (read-only: the rolling state is committed by the next phase's staging) + L2-normalised q / k + gated delta
            // rule (gdn/conv_update.rs:8-55, gdn/update.rs:13-118). CTAs per v head = gridDim / Hv; a CTA owns a block of the head's Dv
            // state rows (one warp per row); the raw output goes to dn_out_raw, the head RMS norm * silu(z) is applied by the consumer
            // (MK_IN_DELTA staging of the out projection).
            constexpr uint32_t DK = 128;
            const uint32_t Hv = op.dn_num_v_heads, Dv = op.dn_hv_dim, Hk = op.dn_num_k_heads;
            const uint32_t conv_dim = 2 * op.dn_key_dim + op.dn_value_dim, taps = op.dn_kernel_size - 1;
            const uint32_t cph = gridDim.x / Hv;
            const uint32_t hv = blockIdx.x / cph, part = blockIdx.x % cph;
            if (blockIdx.x >= cph * Hv) { pf_release(); break; }
            const uint32_t rows_per = (Dv + cph - 1) / cph;
            const uint32_t r0 = part * rows_per, r1 = min(Dv, r0 + rows_per);
            if (r0 >= r1) { pf_release(); break; }
            // the recurrent state does not depend on this token: request this warp's first row before anything else, then let the next
            // phase's weights queue behind it
            float4 s_first = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r0 + warp < r1) s_first = *(reinterpret_cast<const float4*>(op.dn_state + ((size_t)hv * Dv + r0 + warp) * 128) + lane);
            pf_release();
            const uint32_t hk = hv / (Hv / Hk);
            float* sqk = reinterpret_cast<float*>(scratch);          // [2][128] conv outputs, then normalised q / k
            float* sv = sqk + 2 * DK;                                // [rows of this CTA]
            float* sred = sv + 256;                                  // [16] partial sums
            auto conv = [&](uint32_t ch) {                           // bf16-rounded SiLU output of channel ch; the state is only read
                const float x = mk_row1(op.dn_in_pc, ch);
                const float* wv = op.dn_conv_weight + (size_t)ch * op.dn_kernel_size;
                const float* st = op.dn_conv_state + (size_t)ch * taps;
                float acc = op.dn_conv_bias ? op.dn_conv_bias[ch] : 0.0f;
                const float wlast = wv[taps];
                float sv_[7], wt[7];
#pragma unroll
                for (uint32_t tp = 0; tp < 7; ++tp) { sv_[tp] = tp < taps ? st[tp] : 0.0f; wt[tp] = tp < taps ? wv[tp] : 0.0f; }
#pragma unroll
                for (uint32_t tp = 0; tp < 7; ++tp)
                    if (tp < taps) acc = __fadd_rn(acc, __fmul_rn(sv_[tp], wt[tp]));
                acc = __fadd_rn(acc, __fmul_rn(x, wlast));
                return round_bf16(act_f32_nofma(UZU_ACT_SILU, acc));
            };
            float mine = 0.0f;
            if (tid < 2 * (int)DK) {
                mine = conv(tid < (int)DK ? hk * DK + tid : op.dn_key_dim + hk * DK + (tid - DK));
                sqk[tid] = mine;
            } else if (tid - 2 * DK < r1 - r0) {
                sv[tid - 2 * DK] = conv(2 * op.dn_key_dim + hv * Dv + r0 + (tid - 2 * DK));
            }
            // L2 norms over the 128 q values (warps 0..3) and the 128 k values (warps 4..7)
            float sqv = warp_sum(__fmul_rn(mine, mine));
            if (lane == 0 && warp < 8) sred[warp] = sqv;
            mk_bar_sync(2, NCT);
            const float qn = __fadd_rn(__fadd_rn(sred[0], sred[1]), __fadd_rn(sred[2], sred[3]));
            const float kn = __fadd_rn(__fadd_rn(sred[4], sred[5]), __fadd_rn(sred[6], sred[7]));
            const float qi = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(qn, 1e-6f))), ki = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(kn, 1e-6f)));
            const float qscale = __fdiv_rn(1.0f, __fsqrt_rn((float)DK));
            mk_bar_sync(2, NCT);                                     // everyone has read sred before it is reused
            float kqp = 0.0f;
            if (tid < (int)DK) {
                const float qq = __fmul_rn(__fmul_rn(sqk[tid], qi), qscale), kk = __fmul_rn(sqk[DK + tid], ki);
                kqp = __fmul_rn(kk, qq);
                sqk[tid] = qq; sqk[DK + tid] = kk;
            }
            kqp = warp_sum(kqp);
            if (lane == 0 && warp < 4) sred[8 + warp] = kqp;
            mk_bar_sync(2, NCT);
            const float kq = __fadd_rn(__fadd_rn(sred[8], sred[9]), __fadd_rn(sred[10], sred[11]));
            const float beta_raw = mk_row1(op.dn_in_pc, conv_dim + op.dn_value_dim + hv);
            const float a_raw = mk_row1(op.dn_in_pc, conv_dim + op.dn_value_dim + Hv + hv);
            const float beta = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-beta_raw)));
            const float sp_in = __fadd_rn(a_raw, op.dn_dt_bias[hv]);
            const float sp = sp_in > 20.0f ? sp_in : logf(__fadd_rn(1.0f, expf(sp_in)));
            const float gdec = __fmul_rn(-expf(op.dn_a_log[hv]), sp);
            const float decay = expf(gdec);
            const float4 q4 = *(reinterpret_cast<const float4*>(sqk) + lane), k4 = *(reinterpret_cast<const float4*>(sqk + DK) + lane);
            for (uint32_t r = r0 + warp; r < r1; r += NCW) {
                const uint32_t row = hv * Dv + r;
                float* srow = op.dn_state + (size_t)row * DK;
                const float4 s = (r == r0 + warp) ? s_first : *(reinterpret_cast<const float4*>(srow) + lane);
                const float v_i = sv[r - r0];
                float sqa = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(s.x, q4.x), __fmul_rn(s.y, q4.y)), __fmul_rn(s.z, q4.z)), __fmul_rn(s.w, q4.w));
                float ska = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(s.x, k4.x), __fmul_rn(s.y, k4.y)), __fmul_rn(s.z, k4.z)), __fmul_rn(s.w, k4.w));
                sqa = warp_sum(sqa);
                ska = warp_sum(ska);
                const float retrieved = __fmul_rn(decay, ska);
                const float delta = __fmul_rn(beta, __fadd_rn(v_i, -retrieved));
                if (lane == 0) op.dn_out_raw[row] = __fadd_rn(__fmul_rn(decay, sqa), __fmul_rn(delta, kq));
                float4 ns;
                ns.x = __fadd_rn(__fmul_rn(decay, s.x), __fmul_rn(k4.x, delta));
                ns.y = __fadd_rn(__fmul_rn(decay, s.y), __fmul_rn(k4.y, delta));
                ns.z = __fadd_rn(__fmul_rn(decay, s.z), __fmul_rn(k4.z, delta));
                ns.w = __fadd_rn(__fmul_rn(decay, s.w), __fmul_rn(k4.w, delta));
                *(reinterpret_cast<float4*>(srow) + lane) = ns;
            }
            break;
        }
        case MK_LOGITS: {
            // logits = bf16(sum of readout pieces); greedy argmax (value desc, index asc: unified_sampling.rs:34-98 with no filters)
            const uint32_t V = op.vocab;
            unsigned long long best = 0ull;
            for (uint32_t n0 = (blockIdx.x * NCT + tid) * 4u; n0 < V; n0 += gridDim.x * NCT * 4u) {
                const float4 v = mk_rows4(op.logits_pc, n0);
                const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    if (n0 + i < V) {
                        op.logits[n0 + i] = f2bf(vv[i]);
                        const unsigned long long key = mk_pack_key(vv[i], n0 + i);
                        best = key > best ? key : best;
                    }
                }
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, off);
                best = other > best ? other : best;
            }
            if (lane == 0) redk[warp] = best;
            mk_bar_sync(2, NCT);
            if (tid == 0) {
                for (int w_ = 1; w_ < NCW; ++w_) best = redk[w_] > best ? redk[w_] : best;
                op.argmax_keys[blockIdx.x] = best;
            }
            break;
        }
        case MK_FINISH: {
            if (blockIdx.x == 0 && warp == 0) {
                unsigned long long best = 0ull;
                for (uint32_t i = lane; i < gridDim.x; i += 32) {
                    const unsigned long long key = __ldcg(op.argmax_keys + i);
                    best = key > best ? key : best;
                }
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) {
                    const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, off);
                    best = other > best ? other : best;
                }
                if (lane == 0) {
                    const uint32_t tok = 0xFFFFFFFFu - (uint32_t)(best & 0xFFFFFFFFull);
                    const uint32_t step = p.state->step;
                    p.sampled[0] = tok;
                    p.token_out[0] = tok;                                   // device-side chaining of the next input (stream.rs:611-615)
                    p.host_ring[step % p.token_ring] = tok;
                    if (p.dev_out) p.dev_out[step - p.dev_out_base_step] = tok;
                    p.state->position = position + 1;
                    p.state->step = step + 1;
                }
            }
            break;
        }
        default: break;
        }
        MK_TRACE(6);
        if (op.kind != MK_FINISH) mk_grid_sync(p, oi, tid, NCT, watch);
        MK_TRACE(7);
    }
#undef MK_TRACE
}

// ---------------------------------------------------------------------------------------------------------------------------------
// decode-stream repack (once per matrix at load)
// ---------------------------------------------------------------------------------------------------------------------------------
// One thread per (unit, lane, vector): writes the unit exactly as the consumer reads it:
//   [j < 4][lane] uint4 = row tile*16 + g,     bytes [(c*4 + j)*64 + t*16, +16)      (g = lane / 4, t = lane % 4)
//   [4 + j][lane] uint4 = row tile*16 + g + 8, same bytes
//   [4096 + w*128 + lane*4], w = 0..3: scales (row g), scales (row g + 8), c pair (row g), c pair (row g + 8); a pair = two bf16
//   for groups gi + 2t, gi + 2t + 1 (gi = c * GPS); c = zero point (exact in bf16) / 8 or 128 (symmetric) / MLX bias.
__global__ void __launch_bounds__(256) mega_repack_kernel(const uint8_t* w, const __nv_bfloat16* scales, const uint8_t* zero_points,
                                                         const __nv_bfloat16* biases, uint32_t n, uint32_t k, uint32_t bits, uint32_t group_size,
                                                         uint32_t method, uint32_t tiles, uint32_t C, uint8_t* out) {
    const uint32_t row_bytes = k * bits / 8;
    const uint32_t npg = group_size * bits / 4;             // nibbles per group
    const uint32_t gps = 512 / npg;
    const uint32_t ngroups = (k + group_size - 1) / group_size;
    const uint32_t zp_stride = bits == 4 ? (ngroups + 1) / 2 : ngroups;
    const size_t total = (size_t)tiles * C * 288;            // 256 weight vectors + 32 coefficient quads per unit
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const uint32_t unit = (uint32_t)(idx / 288), r = (uint32_t)(idx % 288);
        const uint32_t tile = unit / C, c = unit % C;
        uint8_t* ub = out + (size_t)unit * MK_STAGE_BYTES;
        if (r < 256) {
            const uint32_t vec = r / 32, lane = r % 32, g = lane / 4, t = lane % 4;
            const uint32_t j = vec & 3, row = tile * 16 + g + (vec >= 4 ? 8 : 0);
            const uint32_t off = (c * 4 + j) * 64 + t * 16;
            uint4 v = make_uint4(0, 0, 0, 0);
            if (row < n && off + 16 <= row_bytes) v = *reinterpret_cast<const uint4*>(w + (size_t)row * row_bytes + off);
            else if (row < n && off < row_bytes) {
                uint8_t tmp[16];
                for (int i = 0; i < 16; ++i) tmp[i] = off + i < row_bytes ? w[(size_t)row * row_bytes + off + i] : 0;
                v = *reinterpret_cast<uint4*>(tmp);
            }
            *reinterpret_cast<uint4*>(ub + (size_t)vec * 512 + lane * 16) = v;
        } else {
            const uint32_t lane = r - 256, g = lane / 4, t = lane % 4;
            uint32_t words[4] = {0, 0, 0, 0};
            if (2 * t < gps) {
                for (int hb = 0; hb < 2; ++hb) {
                    const uint32_t row = tile * 16 + g + hb * 8;
                    uint32_t spair = 0, cpair = 0;
                    for (int e = 0; e < 2; ++e) {
                        const uint32_t gi = c * gps + 2 * t + e;
                        uint16_t sbits = 0, cbits = 0;
                        if (row < n && gi < ngroups) {
                            sbits = reinterpret_cast<const uint16_t*>(scales)[(size_t)row * ngroups + gi];
                            if (method == UZU_QMETHOD_SCALE_BIAS) cbits = reinterpret_cast<const uint16_t*>(biases)[(size_t)row * ngroups + gi];
                            else {
                                float zp;
                                if (method == UZU_QMETHOD_SCALE_ZERO_POINT) {
                                    if (bits == 4) { const uint8_t pk = zero_points[(size_t)row * zp_stride + gi / 2]; zp = (float)((gi & 1) ? (pk >> 4) : (pk & 15u)); }
                                    else zp = (float)zero_points[(size_t)row * zp_stride + gi];
                                } else zp = bits == 4 ? 8.0f : 128.0f;
                                const __nv_bfloat16 zb = __float2bfloat16_rn(zp);      // integers <= 256 are exact in bf16
                                cbits = *reinterpret_cast<const uint16_t*>(&zb);
                            }
                        }
                        spair |= (uint32_t)sbits << (16 * e);
                        cpair |= (uint32_t)cbits << (16 * e);
                    }
                    words[hb] = spair;
                    words[2 + hb] = cpair;
                }
            }
            for (int wi = 0; wi < 4; ++wi) *reinterpret_cast<uint32_t*>(ub + 4096 + wi * 128 + lane * 4) = words[wi];
        }
    }
}

size_t mega_stream_bytes(uint32_t n, uint32_t k, uint32_t bits) {
    const uint32_t tiles = (n + 15) / 16, C = (k * bits / 4 + 511) / 512;
    return (size_t)tiles * C * MK_STAGE_BYTES;
}

void mega_repack(uzu_context* ctx, const uint8_t* w, const __nv_bfloat16* scales, const uint8_t* zero_points, const __nv_bfloat16* biases,
                 uint32_t n, uint32_t k, uint32_t bits, uint32_t group_size, uint32_t method, uint8_t* out) {
    const uint32_t tiles = (n + 15) / 16, C = (k * bits / 4 + 511) / 512;
    make_current(ctx);
    const size_t total = (size_t)tiles * C * 288;
    const uint32_t blocks = (uint32_t)std::min<size_t>((total + 255) / 256, 65535u * 8u);
    mega_repack_kernel<<<blocks, 256, 0, ctx->stream>>>(w, scales, zero_points, biases, n, k, bits, group_size, method, tiles, C, out);
}

// ---------------------------------------------------------------------------------------------------------------------------------
// launch
// ---------------------------------------------------------------------------------------------------------------------------------
template <int NPG, int BITS, int NCW, int S>
static const char* launch_variant(uzu_context* ctx, const MegaConfig& cfg, const MkParams& p) {
    static std::atomic<uint64_t> done{0};
    const uint64_t bit = 1ull << (ctx->device & 63);
    make_current(ctx);
    if (!(done.load(std::memory_order_acquire) & bit)) {
        cudaError_t e = cudaFuncSetAttribute(decode_mega_kernel<NPG, BITS, NCW, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 6144);
        if (e != cudaSuccess) return cudaGetErrorString(e);
        done.fetch_or(bit, std::memory_order_release);
    }
    cudaLaunchConfig_t lc{};
    lc.gridDim = dim3(cfg.grid);
    lc.blockDim = dim3(NCW * 32);
    lc.dynamicSmemBytes = cfg.smem_bytes;
    lc.stream = ctx->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;       // co-residency of all CTAs is guaranteed (or the launch fails): the grid barrier cannot deadlock
    attr[0].val.cooperative = 1;
    lc.attrs = attr;
    lc.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&lc, decode_mega_kernel<NPG, BITS, NCW, S>, p);
    return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

static int mega_env(const char* name, int dflt) {
    const char* v = getenv(name);
    return v ? atoi(v) : dflt;
}

bool mega_config(uzu_context* ctx, uint32_t npg, uint32_t bits, uint32_t scratch_bytes, MegaConfig* out) {
    if (!((npg == 64 && bits == 4) || (npg == 128 && (bits == 4 || bits == 8)))) return false;
    MegaConfig c{};
    c.npg = npg; c.bits = bits;
    // (warps, ring stages per warp): the first configuration whose ring + scratch fits in shared memory. In flight per SM = (S - 1) / S
    // of the ring; 14 x 3 keeps ~129 KB in flight with 146 registers per thread.
    static const uint32_t table[][2] = {{14, 3}, {12, 3}, {16, 2}, {8, 4}};
    const uint32_t want = (uint32_t)mega_env("UZU_MEGA_WARPS", 0);
    c.grid = (uint32_t)ctx->sm_count;
    c.scratch_bytes = (scratch_bytes + 127u) & ~127u;
    bool found = false;
    for (auto& tb : table) {
        if (want && tb[0] != want) continue;
        c.ncw = tb[0]; c.stages = tb[1];
        c.smem_bytes = c.scratch_bytes + ((c.ncw * c.stages * 8u + 127u) & ~127u) + (size_t)c.ncw * c.stages * MK_STAGE_BYTES;
        if (c.smem_bytes <= 227u * 1024u - 6144u) { found = true; break; }
    }
    if (!found) return false;
    *out = c;
    return true;
}

const char* mega_launch(uzu_context* ctx, const MegaConfig& cfg, const MkParams& p) {
#define UZU_MK(NPG_, BITS_)                                                                   \
    if (cfg.npg == NPG_ && cfg.bits == BITS_) {                                               \
        if (cfg.ncw == 14) return launch_variant<NPG_, BITS_, 14, 3>(ctx, cfg, p);            \
        if (cfg.ncw == 12) return launch_variant<NPG_, BITS_, 12, 3>(ctx, cfg, p);            \
        if (cfg.ncw == 16) return launch_variant<NPG_, BITS_, 16, 2>(ctx, cfg, p);            \
        return launch_variant<NPG_, BITS_, 8, 4>(ctx, cfg, p);                                \
    }
    UZU_MK(64, 4)
    UZU_MK(128, 4)
    UZU_MK(128, 8)
#undef UZU_MK
    return "decode_mega: unsupported quantisation geometry";
}

}  // namespace uzu
