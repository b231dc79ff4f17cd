// Persistent whole-token decode kernel ("mega" path): descriptors shared by decode_mega.cu (kernel + launch) and engine.cu
// (which compiles a model into a phase program). See decode_mega.cu for the design.
#pragma once

#include <stdint.h>

#include "common.cuh"

namespace uzu {

constexpr uint32_t MK_STAGE_BYTES = 4608;     // one unit: 16 rows x 256 packed bytes (4096 B) + 512 B of per-group coefficients
constexpr uint32_t MK_MAX_MATS = 2;

enum MkKind : uint32_t { MK_GEMV = 1, MK_ATTN = 3, MK_ACT = 4, MK_DN_UPDATE = 6, MK_LOGITS = 7, MK_FINISH = 8 };
enum MkInput : uint32_t { MK_IN_PLAIN = 0, MK_IN_NORM = 1, MK_IN_SIGMOID = 2, MK_IN_DELTA = 3, MK_IN_GATED = 4 };
enum MkSource : uint32_t { MK_SRC_BF16 = 0, MK_SRC_PIECES = 1, MK_SRC_EMBED = 2 };

// Output of a streamed GEMV: every (tile, warp range) intersection leaves one partial sum of the tile's 16 rows ("piece").
// Row r of the matmul output = bf16( sum_{p < P} pieces[((r / 16) * P + p) * 16 + r % 16] ), summed in slot order. The partition is static,
// so the slots a tile does not use are never written: they keep the zeros of the initial memset and the consumers need no per-tile count
// (one dependent load less on the critical path of every phase).
struct MkPieces {
    const float* pieces;
    uint32_t P;               // piece slots per tile
    uint32_t pad;
};

struct MkMat {
    const uint8_t* stream;    // [tiles][C][MK_STAGE_BYTES] decode-stream layout (mega_repack_kernel)
    float* pieces;            // [tiles][P][16]
    uint32_t n, k, tiles, C;  // C = 512-nibble super-chunks per row
    uint32_t P, bias_form;    // bias_form: MLX scale/bias (value = s*q + b) instead of s*(q - zp)
    uint32_t unit0, pad;      // first unit of this matrix in the phase's unit space
};

struct MkEmbed {              // quantised / full-precision input embedding row of the current token (embedding.rs:345-372)
    const uint8_t* weights;
    const __nv_bfloat16* scales;
    const uint8_t* zero_points;
    const __nv_bfloat16* biases;
    uint32_t full_precision, mode, method, group_size, vocab;
    float input_scale;
};

struct alignas(16) MkOp {
    uint32_t kind, pad0;
    // ---- MK_GEMV -----------------------------------------------------------------------------------------------------------
    uint32_t nmat, units;
    MkMat mat[MK_MAX_MATS];
    uint32_t k;                    // input length (all matrices of a phase share the input row)
    uint32_t in_kind, src_kind, pad1;
    const __nv_bfloat16* src_vec;  // MK_SRC_BF16: the input row (plain) / pre-norm row
    MkPieces src_pc;               // MK_SRC_PIECES: matmul output pieces, row offset src_row0
    uint32_t src_row0, pad2;
    MkEmbed embed;                 // MK_SRC_EMBED
    // MK_IN_NORM (normalization.rs:50-125; shortcut add before the norm, transformer_layer.rs:95-150)
    const __nv_bfloat16* shortcut_in;
    __nv_bfloat16* shortcut_out;
    const float* norm_scales;
    float norm_eps, norm_scale_offset;
    uint32_t norm_residual_add, norm_full_layer;
    // MK_IN_SIGMOID (sigmoid_gate.rs:7-22): x = src_vec * sigmoid(gate row), gate row from pieces
    MkPieces gate_pc;
    // MK_IN_GATED (gated_act_mul/mod.rs:5-12): x[j] = bf16(bf16(up_j) * bf16(act(bf16(gate_j)))), rows [0, F) up / [F, 2F) gate of gated_pc
    MkPieces gated_pc;
    uint32_t gated_act, pad5;
    // optional side job of the staging: commit the rolling conv state of the previous MK_DN_UPDATE phase (dn_* fields below)
    uint32_t dn_commit, pad6;
    // MK_IN_DELTA (gdn/update.rs:120-144): x = raw * inv_rms(head) * norm_weight * silu(z)
    const float* dn_raw;           // [Hv * Dv] f32 written by MK_DN_UPDATE
    const float* dn_norm_weight;   // [Dv]
    MkPieces dn_z_pc;              // in_proj pieces; z rows start at dn_z_row0
    uint32_t dn_z_row0, dn_heads, dn_head_v_dim;
    float dn_eps;
    // ---- MK_PREP / MK_ATTN -------------------------------------------------------------------------------------------------
    MkPieces qkv_pc;
    __nv_bfloat16* queries;        // [Hq][D]
    __nv_bfloat16* keys;           // KV cache, token-major [T][Hkv][D]
    __nv_bfloat16* values;
    __nv_bfloat16* attn_out;       // [Hq][D]
    const float* rope_cos;         // [positions][rope_dim] or null
    const float* rope_sin;
    const float* qnorm_scales;
    const float* knorm_scales;
    float qnorm_eps, qnorm_offset, knorm_eps, knorm_offset;
    uint32_t qnorm_full_layer, knorm_full_layer, qnorm_has_scales, knorm_has_scales;
    uint32_t qnorm_present, knorm_present;
    uint32_t num_q_heads, num_kv_heads, head_dim, rope_dim;
    float attn_scale;
    uint32_t pad3;
    float* attn_part;              // [Hkv][parts][G][D + 2] CTA partials (o, then m, l)
    unsigned int* attn_tickets;    // [Hkv]
    // ---- MK_ACT (gated_act_mul/mod.rs:5-12) ---------------------------------------------------------------------------------
    MkPieces up_pc;
    __nv_bfloat16* hidden;         // [F]
    uint32_t act_dim, act_type;
    // ---- MK_DN_CONV / MK_DN_UPDATE (gdn/conv_update.rs:8-55, gdn/update.rs:13-144) --------------------------------------------
    MkPieces dn_in_pc;             // in_proj pieces: [q | k | v | z | beta | a]
    const float* dn_conv_weight;   // [conv_dim][kernel]
    const float* dn_conv_bias;     // or null
    float* dn_conv_state;          // [conv_dim][kernel - 1]
    const float* dn_a_log;
    const float* dn_dt_bias;
    float* dn_state;               // [Hv][Dv][128]
    float* dn_out_raw;             // [Hv * Dv]
    uint32_t dn_kernel_size, dn_key_dim, dn_value_dim, dn_num_k_heads, dn_num_v_heads, dn_hv_dim;
    // ---- MK_LOGITS / MK_FINISH ------------------------------------------------------------------------------------------------
    MkPieces logits_pc;
    __nv_bfloat16* logits;         // [V]
    uint32_t vocab, pad4;
    unsigned long long* argmax_keys;   // [grid]
};

static_assert(sizeof(MkOp) % 16 == 0 && sizeof(MkOp) <= 1024, "MkOp is copied to shared memory in 16-byte pieces");

// One entry per GEMV phase, in program order: what a warp needs to walk its own weight stream ahead of the phases (32 bytes, __ldg)
struct alignas(16) MkStream {
    const uint8_t* stream0;        // matrix 0 of the phase
    const uint8_t* stream1;        // matrix 1 (or null)
    uint32_t units, split;         // units of the phase; first unit of matrix 1 (== units when there is one matrix)
    uint32_t hold;                 // the phase before this one issues latency-critical HBM loads (KV rows, recurrent state): the prefetch
                                   // cursor parks here until that phase has issued them, so they do not queue behind ~20 MB of weights
    uint32_t pad;
};

struct MkStepState {               // == engine.cu's DecodeState
    uint32_t position, step;
    uint64_t base_seed;
};

struct MkParams {
    const MkOp* ops;
    const MkStream* streams;       // the GEMV phases' weight streams, in program order
    uint32_t nstreams, pad_s;
    uint32_t nops, ncw;            // consumer warps per CTA the program was partitioned for
    MkStepState* state;
    unsigned long long* barrier;   // [0] monotonic arrival counter (never reset: launch k's barrier b completes at barrier_base + (b + 1) * grid), [1] low word = device-side error flag polled by the spin loops
    unsigned long long barrier_base;
    unsigned int* error_flag;      // pinned host mirror of the error flag, written only when a spin loop times out (the kernel drains instead of hanging)
    const uint32_t* token_ids;     // [1] input token (device-chained)
    uint32_t* token_out;           // == token_ids (next input)
    uint32_t* sampled;
    volatile uint32_t* host_ring;
    uint32_t* dev_out;
    uint32_t dev_out_base_step, token_ring;
    uint32_t scratch_bytes;        // shared-memory scratch (activation row / attention merge) carved in front of the rings
    uint32_t stages;
    unsigned long long* trace;     // debug: [nops][4] SM clock stamps of CTA trace_cta (op start, after staging, after body, after barrier) or null
    uint32_t trace_cta, pad;
};

// decode_mega.cu
size_t mega_stream_bytes(uint32_t n, uint32_t k, uint32_t bits);
// Repack one quantised matrix [n, k] (uzu layout) into the decode-stream layout. method: UZU_QMETHOD_*; bits 4 / 8; group_size 64 (int4 / int8) or 128 (int4).
void mega_repack(uzu_context* ctx, const uint8_t* w, const __nv_bfloat16* scales, const uint8_t* zero_points, const __nv_bfloat16* biases,
                 uint32_t n, uint32_t k, uint32_t bits, uint32_t group_size, uint32_t method, uint8_t* out);
struct MegaConfig { uint32_t npg, bits, ncw, stages, grid; size_t smem_bytes; uint32_t scratch_bytes; };
bool mega_config(uzu_context* ctx, uint32_t npg, uint32_t bits, uint32_t scratch_bytes, MegaConfig* out);
const char* mega_launch(uzu_context* ctx, const MegaConfig& cfg, const MkParams& p);

}  // namespace uzu
