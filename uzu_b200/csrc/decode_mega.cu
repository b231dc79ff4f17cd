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
            asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p.barrier) : "memory");
            if (v >= target) break;
            if ((++spins & 63u) == 0u) {
                if (w.poll_dead()) break;
                if (clock64() - t0 > MK_TIMEOUT_CLOCKS) { w.fail(0x100u); break; }
            }
        }
    }
    mk_bar_sync(1, nct);
}

__device__ __forceinline__ void mk_mma_16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
// (128 + q_lo, 128 + q_hi) as two exact bf16 from nibbles `shift` and `shift + 16` of w: one shift + one LOP3 (magic kept in a register)
__device__ __forceinline__ uint32_t mk_nib_pair(uint32_t w, int shift, uint32_t magic) {
    uint32_t r;
    const uint32_t s = shift ? (w >> shift) : w;
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

// Range r of a phase with U units: consecutive ranges sit on different SMs (r = warp * gridDim + cta) and only min(W, U) ranges exist,
// so every range holds at least one unit and the pieces of a tile are numbered without gaps.
__device__ __forceinline__ void mk_my_range(uint32_t cta, uint32_t warp, uint32_t U, uint32_t W, uint32_t& ri, uint32_t& weff, uint32_t& ub, uint32_t& ue) {
    ri = warp * gridDim.x + cta;
    weff = min(W, U);
    if (ri >= weff) { ub = ue = 0; return; }
    ub = (uint32_t)(((uint64_t)ri * U) / weff);
    ue = (uint32_t)(((uint64_t)(ri + 1) * U) / weff);
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
__global__ void __launch_bounds__((NCW + 1) * 32, 1) decode_mega_kernel(const MkParams p) {
    constexpr int NCT = NCW * 32;                 // consumer threads
    constexpr int CPM = NPG >= 128 ? 1 : 2;       // quantisation groups per 128-nibble chunk
    constexpr int GPS = 512 / NPG;                // groups per unit
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* scratch = smem;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.scratch_bytes);           // full[NCW][S], empty[NCW][S]
    uint8_t* ring = smem + p.scratch_bytes + ((2u * NCW * S * 8u + 127u) & ~127u);
    __shared__ float red[NCW + 4];
    __shared__ unsigned long long redk[NCW];
    __shared__ unsigned int sm_ticket;
    // phase descriptors are staged in shared memory one phase ahead: every field read of a phase is a shared-memory access instead of a
    // chain of first-touch global loads (measured: ~3 us per phase before this)
    __shared__ MkOp sops[2];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t W = gridDim.x * NCW;
    MkWatch watch{reinterpret_cast<unsigned int*>(p.barrier + 1), p.error_flag, false};

    if (tid == 0) {
        for (int i = 0; i < NCW * S; ++i) {
            mk_mbar_init(mk_smem_u32(bars + i), 1);
            mk_mbar_init(mk_smem_u32(bars + NCW * S + i), 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == NCW) {
        // ================================================= producer warp ====================================================
        // lane l feeds consumer warp l: the same walk over (GEMV phase, range) the consumer does, S stages ahead at most. The lanes
        // stay converged in ONE polling loop with a non-blocking mbarrier test: a lane whose ring is full (its consumer is parked at a
        // grid barrier) must not stall the lanes whose rings have room -- that is exactly when the next phase's weights are prefetched.
        const bool active = lane < NCW;
        uint32_t q = 0;                                       // unit sequence number of this consumer warp (across phases)
        const uint32_t full0 = mk_smem_u32(bars + (active ? lane : 0) * S), empty0 = mk_smem_u32(bars + NCW * S + (active ? lane : 0) * S);
        const uint32_t ring0 = mk_smem_u32(ring + (size_t)(active ? lane : 0) * S * MK_STAGE_BYTES);
        uint32_t oi = 0, u = 0, ue = 0, mend = 0;
        const uint8_t* src = nullptr;
        bool done = !active;
        auto load_op = [&](uint32_t from) {       // first GEMV phase at index >= from in which this consumer warp owns units
            for (oi = from; oi < p.nops; ++oi) {
                if (p.ops[oi].kind != MK_GEMV) continue;
                uint32_t ri, weff;
                mk_my_range(blockIdx.x, lane, p.ops[oi].units, W, ri, weff, u, ue);
                if (u < ue) return true;
            }
            return false;
        };
        auto load_segment = [&]() {               // the part of [u, ue) that lies in one matrix: one contiguous byte range of its stream
            const MkOp* op = p.ops + oi;
            const int mi = (op->nmat > 1 && u >= op->mat[1].unit0) ? 1 : 0;
            mend = min(ue, op->mat[mi].unit0 + op->mat[mi].tiles * op->mat[mi].C);
            src = op->mat[mi].stream + (size_t)(u - op->mat[mi].unit0) * MK_STAGE_BYTES;
        };
        if (active) {
            if (load_op(0)) load_segment();
            else done = true;
        }
        const uint64_t policy = mk_evict_first_policy();
        const long long t0 = clock64();
        uint32_t spins = 0;
        while (__any_sync(0xffffffffu, !done)) {
            if (!done) {
                const uint32_t s = q % S, ph = (q / S) & 1u;
                if (mk_mbar_test_wait(empty0 + s * 8u, ph ^ 1u)) {
                    mk_mbar_expect_tx(full0 + s * 8u, MK_STAGE_BYTES);
                    mk_bulk_copy(ring0 + s * MK_STAGE_BYTES, src, MK_STAGE_BYTES, full0 + s * 8u, policy);
                    ++q; ++u; src += MK_STAGE_BYTES;
                    if (u >= mend) {
                        if (u >= ue && !load_op(oi + 1)) done = true;
                        if (!done) load_segment();
                    }
                }
            }
            if ((++spins & 1023u) == 0u) {
                if (watch.poll_dead()) done = true;
                else if (clock64() - t0 > 8 * MK_TIMEOUT_CLOCKS) { watch.fail(0x200u); done = true; }   // a whole step never takes this long
            }
        }
        return;
    }

    // ===================================================== consumer warps =====================================================
    const int g = lane >> 2, t = lane & 3;
    const uint32_t gw = warp * gridDim.x + blockIdx.x;      // consecutive work items sit on different SMs
    const uint32_t position = p.state->position;            // prefix length == position of the token being fed
    const uint32_t token = p.token_ids[0];
    uint32_t q = 0;                                          // unit sequence number (matches the producer's)
    const uint32_t full0 = mk_smem_u32(bars + warp * S), empty0 = mk_smem_u32(bars + NCW * S + warp * S);
    const uint8_t* ring_w = ring + (size_t)warp * S * MK_STAGE_BYTES;
    uint32_t magic;
    asm volatile("mov.b32 %0, 0x43004300;" : "=r"(magic));
    const float mult128 = BITS == 4 ? 128.0f : 128.0f * 17.0f;
    const bool lane_has_groups = 2 * t < GPS;
    const int b_chunk = CPM == 2 ? (g >> 1) : g;
    const bool b_lane = CPM == 2 ? ((g & 1) == (t >> 1)) : true;

#define MK_TRACE(k) do { if (p.trace && blockIdx.x == p.trace_cta && tid == 0) p.trace[(size_t)oi * 4 + (k)] = (unsigned long long)clock64(); } while (0)
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
            constexpr int MAXO = (2048 + NCT - 1) / NCT;                  // octets per thread: rows of up to 16384 elements
            uint4 vals[MAXO];
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
            // -- pass 1: the row before the (optional) normalisation, 8 elements per thread step
            float ssq = 0.0f;
#pragma unroll
            for (int r = 0; r < MAXO; ++r) {
                const uint32_t o = tid + r * NCT, e0 = o * EPO;
                vals[r] = make_uint4(0, 0, 0, 0);
                if (o >= octets_all || e0 >= K) continue;
                uint4 v;
                if (in_kind == MK_IN_GATED) {
                    // GatedActMul folded into the consumer: rows [0, F) of the fused up projection are `up`, rows [F, 2F) `gate`
                    const float4 u0 = mk_rows4(op.gated_pc, e0), u1 = mk_rows4(op.gated_pc, e0 + 4);
                    const float4 g0 = mk_rows4(op.gated_pc, K + e0), g1 = mk_rows4(op.gated_pc, K + e0 + 4);
                    const float uu[8] = {u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w}, gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
                    float hh[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) hh[i] = __fmul_rn(uu[i], round_bf16(act_f32_nofma(op.gated_act, gg[i])));
                    __nv_bfloat162 tt;
                    tt = __floats2bfloat162_rn(hh[0], hh[1]); v.x = *reinterpret_cast<uint32_t*>(&tt);
                    tt = __floats2bfloat162_rn(hh[2], hh[3]); v.y = *reinterpret_cast<uint32_t*>(&tt);
                    tt = __floats2bfloat162_rn(hh[4], hh[5]); v.z = *reinterpret_cast<uint32_t*>(&tt);
                    tt = __floats2bfloat162_rn(hh[6], hh[7]); v.w = *reinterpret_cast<uint32_t*>(&tt);
                } else if (src_kind == MK_SRC_BF16) v = __ldcg(reinterpret_cast<const uint4*>(op.src_vec + e0));
                else if (src_kind == MK_SRC_PIECES) v = mk_rows8_bf16(op.src_pc, op.src_row0 + e0);
                else {
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
                if (in_kind == MK_IN_NORM) {
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
                } else if (in_kind == MK_IN_SIGMOID) {
                    const float4 ga = mk_rows4(op.gate_pc, e0), gb = mk_rows4(op.gate_pc, e0 + 4);
                    const float gt[8] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, gb.z, gb.w};
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float m0 = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-gt[2 * i])));
                        const float m1 = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-gt[2 * i + 1])));
                        a2[i] = __floats2bfloat162_rn(__fmul_rn(__low2float(a2[i]), m0), __fmul_rn(__high2float(a2[i]), m1));
                    }
                }
                vals[r] = v;
            }
            float rms_inv = 0.0f;
            if (in_kind == MK_IN_NORM) {
                ssq = warp_sum(ssq);
                if (lane == 0) red[warp] = ssq;
                mk_bar_sync(2, NCT);
                float tot = 0.0f;
#pragma unroll
                for (int w_ = 0; w_ < NCW; ++w_) tot = __fadd_rn(tot, red[w_]);
                rms_inv = __frcp_rn(__fsqrt_rn(__fadd_rn(__fdiv_rn(tot, (float)K), op.norm_eps)));
            }
            // -- pass 2: final bf16 row -> B-operand order + per-group sums
#pragma unroll
            for (int r = 0; r < MAXO; ++r) {
                const uint32_t o = tid + r * NCT, e0 = o * EPO;
                if (o >= octets_all) continue;          // warp-uniform: octets_all is a multiple of 32
                uint4 v = vals[r];
                __nv_bfloat162* a2 = reinterpret_cast<__nv_bfloat162*>(&v);
                const bool live = e0 < K;
                if (live && in_kind == MK_IN_NORM) {
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
                } else if (live && in_kind == MK_IN_DELTA) {
                    // raw DeltaNet output -> RMS over the head (Dv elements = Dv / 8 adjacent threads) * norm_weight * silu(z)
                    const float4 r0 = __ldcg(reinterpret_cast<const float4*>(op.dn_raw + e0)), r1 = __ldcg(reinterpret_cast<const float4*>(op.dn_raw + e0 + 4));
                    const float rv[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
                    float ss = 0.0f;
#pragma unroll
                    for (int i = 0; i < 8; ++i) ss = __fadd_rn(ss, __fmul_rn(rv[i], rv[i]));
                    const uint32_t Dv = op.dn_head_v_dim;
                    for (uint32_t off = 1; off < Dv / 8; off <<= 1) ss = __fadd_rn(ss, __shfl_xor_sync(0xffffffffu, ss, off));
                    const float inv_rms = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(__fdiv_rn(ss, (float)Dv), op.dn_eps)));
                    const uint32_t i0 = e0 % Dv;
                    const float4 z0 = mk_rows4(op.dn_z_pc, op.dn_z_row0 + e0), z1 = mk_rows4(op.dn_z_pc, op.dn_z_row0 + e0 + 4);
                    const float zv[8] = {z0.x, z0.y, z0.z, z0.w, z1.x, z1.y, z1.z, z1.w};
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
            }
            mk_bar_sync(2, NCT);
            MK_TRACE(1);

            // ---------------- this warp's contiguous range of units -------------------------------------------------------------
            const uint64_t U = op.units;
            uint32_t ri, weff, u, ue;
            mk_my_range(blockIdx.x, warp, op.units, W, ri, weff, u, ue);
            while (u < ue) {
                const int mi = (op.nmat > 1 && u >= op.mat[1].unit0) ? 1 : 0;
                const MkMat M = op.mat[mi];               // by value: the hot loop must not re-read the descriptor from global memory
                const uint32_t v0 = u - M.unit0;
                uint32_t tile = v0 / C, c = v0 % C;
                const uint32_t mend = min(ue, M.unit0 + M.tiles * C);
                float acc0 = 0.0f, acc1 = 0.0f, acc2 = 0.0f, acc3 = 0.0f;
                for (; u < mend; ++u, ++q) {
                    const uint32_t s = q % S, ph = (q / S) & 1u;
                    mk_wait(full0 + s * 8u, ph, watch, 0x300u);
                    const uint4* sw = reinterpret_cast<const uint4*>(ring_w + (size_t)s * MK_STAGE_BYTES) + lane;
                    const uint32_t* swd = reinterpret_cast<const uint32_t*>(ring_w + (size_t)s * MK_STAGE_BYTES + 4096) + lane;
                    float d[4] = {0.0f, 0.0f, 0.0f, 0.0f}, d2[4] = {0.0f, 0.0f, 0.0f, 0.0f};
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
                            mk_mma_16816(d, a0, b0, a1, b1, xb.x, xb.y);
                            mk_mma_16816(d2, a2, b2, a3, b3, xb.z, xb.w);
                        }
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i) d[i] += d2[i];
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
                    __syncwarp();
                    if (lane == 0) mk_mbar_arrive(empty0 + s * 8u);
                    // ---- end of a tile or of the range: leave a piece -----------------------------------------------------------
                    if (c + 1 == C || u + 1 == mend) {
                        float ra = acc0 + acc1, rb = acc2 + acc3;
                        ra += __shfl_xor_sync(0xffffffffu, ra, 1);
                        rb += __shfl_xor_sync(0xffffffffu, rb, 1);
                        ra += __shfl_xor_sync(0xffffffffu, ra, 2);
                        rb += __shfl_xor_sync(0xffffffffu, rb, 2);
                        const uint32_t first = mk_range_of((uint64_t)M.unit0 + (uint64_t)tile * C, U, weff);
                        float* dst = M.pieces + ((size_t)tile * M.P + (ri - first)) * 16;
                        if (t == 0) { dst[g] = ra; dst[g + 8] = rb; }
                        acc0 = acc1 = acc2 = acc3 = 0.0f;
                    }
                    if (++c == C) { c = 0; ++tile; }
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
            if (blockIdx.x >= cph * Hkv || part >= nparts) break;
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
                    float ev[EPT], pv[EPT];
                    // every load of this head first (one L2 round trip), then the arithmetic
#pragma unroll
                    for (int i = 0; i < EPT; ++i) {
                        const uint32_t dd = lane + 32 * i;
                        ev[i] = 0.0f; pv[i] = 0.0f;
                        if (dd < D) {
                            ev[i] = mk_row1(op.qkv_pc, h * D + dd);
                            if (roped && dd < rd) pv[i] = mk_row1(op.qkv_pc, h * D + (dd < half ? dd + half : dd - half));
                        }
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
                    const float* cosr = roped ? op.rope_cos + (size_t)position * rd : nullptr;
                    const float* sinr = roped ? op.rope_sin + (size_t)position * rd : nullptr;
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
                            e = round_bf16(__fadd_rn(__fmul_rn(e, cosr[dd]), __fmul_rn(signed_p, sinr[dd])));
                        }
                        if (is_q) sq[role * D + dd] = __fmul_rn(op.attn_scale, e);
                        else if (is_k) op.keys[((size_t)position * Hkv + kvh) * D + dd] = f2bf(e);
                        else op.values[((size_t)position * Hkv + kvh) * D + dd] = f2bf(e);
                    }
                }
            }
            mk_bar_sync(2, NCT);
            const uint32_t sub = lane / LPK, li = lane % LPK, d0 = li * 8;
            const uint32_t kbeg = part * kp, kend = min(seq, kbeg + kp);
            constexpr int MAXG = 4;       // query heads per kv head held in registers (the host rejects larger groups)
            float qf[MAXG][8], o[MAXG][8], mrun[MAXG], lrun[MAXG];
#pragma unroll
            for (int h = 0; h < MAXG; ++h) {
                mrun[h] = -INFINITY; lrun[h] = 0.0f;
#pragma unroll
                for (int e = 0; e < 8; ++e) { o[h][e] = 0.0f; qf[h][e] = 0.0f; }
                if ((uint32_t)h < G) {
                    const float4 qa = *reinterpret_cast<const float4*>(sq + h * D + d0), qb = *reinterpret_cast<const float4*>(sq + h * D + d0 + 4);
                    qf[h][0] = qa.x; qf[h][1] = qa.y; qf[h][2] = qa.z; qf[h][3] = qa.w; qf[h][4] = qb.x; qf[h][5] = qb.y; qf[h][6] = qb.z; qf[h][7] = qb.w;
                }
            }
            mk_bar_sync(2, NCT);                                       // sq is reused below
            const __nv_bfloat16* kbase = op.keys + (size_t)kvh * D + d0;
            const __nv_bfloat16* vbase = op.values + (size_t)kvh * D + d0;
            const size_t rstride = (size_t)Hkv * D;
            constexpr int UU = 4;     // key rows in flight per lane group (all loads of a batch are issued before any is consumed)
            // the loop bound is warp-uniform (the lane groups of a warp run in lockstep: the shuffles below need every lane);
            // rows past the end of the range are masked with ok[]
            for (uint32_t kb = kbeg + warp * KPW; kb < kend; kb += step_keys * UU) {
                const uint32_t k0 = kb + sub;
                uint4 kr[UU], vr[UU];
                bool ok[UU];
#pragma unroll
                for (int u_ = 0; u_ < UU; ++u_) {
                    const uint32_t ki = k0 + u_ * step_keys;
                    ok[u_] = ki < kend;
                    kr[u_] = ok[u_] ? __ldcg(reinterpret_cast<const uint4*>(kbase + (size_t)ki * rstride)) : make_uint4(0, 0, 0, 0);
                    vr[u_] = ok[u_] ? __ldcg(reinterpret_cast<const uint4*>(vbase + (size_t)ki * rstride)) : make_uint4(0, 0, 0, 0);
                }
#pragma unroll
                for (int u_ = 0; u_ < UU; ++u_) {
                    float kf[8], vf[8];
                    {
                        const __nv_bfloat162* k2 = reinterpret_cast<const __nv_bfloat162*>(&kr[u_]);
                        const __nv_bfloat162* v2 = reinterpret_cast<const __nv_bfloat162*>(&vr[u_]);
#pragma unroll
                        for (int e = 0; e < 4; ++e) { kf[2 * e] = __low2float(k2[e]); kf[2 * e + 1] = __high2float(k2[e]); vf[2 * e] = __low2float(v2[e]); vf[2 * e + 1] = __high2float(v2[e]); }
                    }
#pragma unroll
                    for (int h = 0; h < MAXG; ++h) {
                        if ((uint32_t)h >= G) break;
                        float sdot = 0.0f;
#pragma unroll
                        for (int e = 0; e < 8; ++e) sdot = fmaf(qf[h][e], kf[e], sdot);
                        for (uint32_t off = LPK / 2; off > 0; off >>= 1) sdot += __shfl_xor_sync(0xffffffffu, sdot, off);
                        if (ok[u_]) {
                            const float mnew = fmaxf(mrun[h], sdot);
                            const float factor = (mrun[h] == -INFINITY) ? 0.0f : expf(mrun[h] - mnew);
                            const float pv_ = expf(sdot - mnew);
                            lrun[h] = lrun[h] * factor + pv_;
                            mrun[h] = mnew;
#pragma unroll
                            for (int e = 0; e < 8; ++e) o[h][e] = fmaf(pv_, vf[e], o[h][e] * factor);
                        }
                    }
                }
            }
            // merge the KPW lane groups of the warp (xor LPK, 2 LPK, ...), then the warps of the CTA through shared memory
#pragma unroll
            for (int h = 0; h < MAXG; ++h) {
                if ((uint32_t)h >= G) break;
                for (uint32_t off = LPK; off < 32; off <<= 1) {
                    const float mo = __shfl_xor_sync(0xffffffffu, mrun[h], off), lo_ = __shfl_xor_sync(0xffffffffu, lrun[h], off);
                    const float mn = fmaxf(mrun[h], mo);
                    const float fa = (mrun[h] == -INFINITY) ? 0.0f : expf(mrun[h] - mn), fb = (mo == -INFINITY) ? 0.0f : expf(mo - mn);
                    lrun[h] = lrun[h] * fa + lo_ * fb;
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const float oo = __shfl_xor_sync(0xffffffffu, o[h][e], off);
                        o[h][e] = o[h][e] * fa + oo * fb;
                    }
                    mrun[h] = mn;
                }
            }
            float* so = reinterpret_cast<float*>(scratch);              // [NCW][G][D]
            float* sml = so + (size_t)NCW * G * D;                      // [NCW][G][2]
            if (sub == 0) {
#pragma unroll
                for (int h = 0; h < MAXG; ++h) {
                    if ((uint32_t)h >= G) break;
                    float* dst = so + ((size_t)warp * G + h) * D + d0;
                    *reinterpret_cast<float4*>(dst) = make_float4(o[h][0], o[h][1], o[h][2], o[h][3]);
                    *reinterpret_cast<float4*>(dst + 4) = make_float4(o[h][4], o[h][5], o[h][6], o[h][7]);
                    if (li == 0) { sml[((size_t)warp * G + h) * 2] = mrun[h]; sml[((size_t)warp * G + h) * 2 + 1] = lrun[h]; }
                }
            }
            mk_bar_sync(2, NCT);
            float* pbase = op.attn_part + ((size_t)kvh * cph + part) * G * (D + 2);
            for (uint32_t idx = tid; idx < G * D; idx += NCT) {
                const uint32_t h = idx / D, dd = idx % D;
                float M = -INFINITY;
#pragma unroll
                for (int w_ = 0; w_ < NCW; ++w_) M = fmaxf(M, sml[((size_t)w_ * G + h) * 2]);
                float L = 0.0f, O = 0.0f;
#pragma unroll
                for (int w_ = 0; w_ < NCW; ++w_) {
                    const float mw = sml[((size_t)w_ * G + h) * 2];
                    const float f = (mw == -INFINITY) ? 0.0f : expf(mw - M);
                    L += sml[((size_t)w_ * G + h) * 2 + 1] * f;
                    O += so[((size_t)w_ * G + h) * D + dd] * f;
                }
                if (nparts == 1) {
                    op.attn_out[((size_t)(kvh * G + h)) * D + dd] = f2bf(O / L);       // a single part: no cross-CTA merge
                } else {
                    pbase[(size_t)h * (D + 2) + dd] = O;
                    if (dd == 0) { pbase[(size_t)h * (D + 2) + D] = M; pbase[(size_t)h * (D + 2) + D + 1] = L; }
                }
            }
            if (nparts == 1) break;
            // the last CTA of this kv head merges the parts (fixed order) and writes the bf16 attention output
            __threadfence();
            mk_bar_sync(2, NCT);
            if (tid == 0) sm_ticket = atomicAdd(op.attn_tickets + kvh, 1u);
            mk_bar_sync(2, NCT);
            if (sm_ticket == nparts - 1) {
                __threadfence();
                const float* hb = op.attn_part + (size_t)kvh * cph * G * (D + 2);
                for (uint32_t idx = tid; idx < G * D; idx += NCT) {
                    const uint32_t h = idx / D, dd = idx % D;
                    float M = -INFINITY;
#pragma unroll 8
                    for (uint32_t pi = 0; pi < nparts; ++pi) M = fmaxf(M, __ldcg(hb + ((size_t)pi * G + h) * (D + 2) + D));
                    float L = 0.0f, O = 0.0f;
#pragma unroll 8
                    for (uint32_t pi = 0; pi < nparts; ++pi) {
                        const float* pp = hb + ((size_t)pi * G + h) * (D + 2);
                        const float f = expf(__ldcg(pp + D) - M);
                        L += __ldcg(pp + D + 1) * f;
                        O += __ldcg(pp + dd) * f;
                    }
                    op.attn_out[((size_t)(kvh * G + h)) * D + dd] = f2bf(O / L);
                }
                if (tid == 0) op.attn_tickets[kvh] = 0u;
            }
            break;
        }
        case MK_ACT: {
            // GatedActMul (gated_act_mul/mod.rs:5-12): hidden[j] = bf16(bf16(up_j) * bf16(act(bf16(gate_j)))), rows [0, F) up, [F, 2F) gate
            const uint32_t F = op.act_dim;
            for (uint32_t j = (blockIdx.x + gridDim.x * tid) * 4u; j < F; j += gridDim.x * NCT * 4u) {
                const float4 up = mk_rows4(op.up_pc, j), gt = mk_rows4(op.up_pc, F + j);
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
            if (blockIdx.x >= cph * Hv) break;
            const uint32_t rows_per = (Dv + cph - 1) / cph;
            const uint32_t r0 = part * rows_per, r1 = min(Dv, r0 + rows_per);
            if (r0 >= r1) break;
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
                const float4 s = *(reinterpret_cast<const float4*>(srow) + lane);
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
        MK_TRACE(2);
        if (op.kind != MK_FINISH) mk_grid_sync(p, oi, tid, NCT, watch);
        MK_TRACE(3);
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
        cudaError_t e = cudaFuncSetAttribute(decode_mega_kernel<NPG, BITS, NCW, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 4096);
        if (e != cudaSuccess) return cudaGetErrorString(e);
        done.fetch_or(bit, std::memory_order_release);
    }
    cudaLaunchConfig_t lc{};
    lc.gridDim = dim3(cfg.grid);
    lc.blockDim = dim3((NCW + 1) * 32);
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
    // consumer warps + 1 producer warp = a multiple of 4 warps (register allocation granularity): 16 warps -> 128 registers per thread
    c.ncw = (uint32_t)mega_env("UZU_MEGA_WARPS", 15);
    c.stages = c.ncw == 15 ? 2u : c.ncw == 11 ? 3u : 5u;
    if (c.ncw != 15 && c.ncw != 11 && c.ncw != 7) return false;
    c.grid = (uint32_t)ctx->sm_count;
    c.scratch_bytes = (scratch_bytes + 127u) & ~127u;
    c.smem_bytes = c.scratch_bytes + ((2u * c.ncw * c.stages * 8u + 127u) & ~127u) + (size_t)c.ncw * c.stages * MK_STAGE_BYTES;
    if (c.smem_bytes > 227u * 1024u - 4096u) return false;
    *out = c;
    return true;
}

const char* mega_launch(uzu_context* ctx, const MegaConfig& cfg, const MkParams& p) {
#define UZU_MK(NPG_, BITS_)                                                                   \
    if (cfg.npg == NPG_ && cfg.bits == BITS_) {                                               \
        if (cfg.ncw == 15) return launch_variant<NPG_, BITS_, 15, 2>(ctx, cfg, p);            \
        if (cfg.ncw == 11) return launch_variant<NPG_, BITS_, 11, 3>(ctx, cfg, p);            \
        return launch_variant<NPG_, BITS_, 7, 5>(ctx, cfg, p);                                \
    }
    UZU_MK(64, 4)
    UZU_MK(128, 4)
    UZU_MK(128, 8)
#undef UZU_MK
    return "decode_mega: unsupported quantisation geometry";
}

}  // namespace uzu
