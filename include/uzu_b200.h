/*
 * uzu_b200.h -- C ABI of libuzu_b200.so: a B200 (sm_100a) CUDA backend for uzu's transformer
 * decode hot path.
 *
 * This is the drop-in boundary. The reference has no FFI: its backends are
 * Rust types implementing the `Backend` / `Context` / `CommandBuffer*` / `DenseBuffer` /
 * `SparseBuffer` / `Kernels` / `<Name>Kernel` traits of crates/backend-uzu. A `backends/cuda`
 * Rust module implements those traits by calling the entry points below one-to-one
 * (INTEGRATION.md shows the binding); each declaration cites the trait item it replaces.
 * Paths are relative to /root/reference/crates/backend-uzu/src/.
 *
 * Conventions
 *   - plain C: opaque handles, raw device pointers (Buffer::gpu_ptr() + byte offset) and sizes;
 *   - `#[repr(C)]` gpu_types are mirrored byte-for-byte (32-bit enums, u32 bit-flags);
 *   - constructors return uzu_status; `*_encode` never blocks and, like the generated Rust
 *     `encode`, cannot fail synchronously: an invalid call records a sticky error on the command
 *     buffer, surfaced by uzu_command_buffer_wait_until_completed (command_buffer.rs:113-118);
 *   - one thread encodes into a command buffer; submit/wait may run on another thread;
 *   - submission order == execution order (one CUDA stream per context; backends/cpu/context.rs:21-27
 *     is the model), so HazardTracker barriers are no-ops as on the CPU backend;
 *   - no CPU fallback anywhere: without a CUDA device uzu_context_create fails.
 */
#ifndef UZU_B200_H
#define UZU_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UZU_API __attribute__((visibility("default")))

/* ---- Backend consts: backends/common/backend.rs:5-18 ------------------------------------ */
#define UZU_BACKEND_NAME "cuda-b200"
#define UZU_MIN_ALLOCATION_ALIGNMENT 256u        /* 128-bit vector loads + TMA need >= 16 B; 256 B = cudaMalloc grain */
#define UZU_MAX_ALLOCATION_ALIGNMENT 256u
#define UZU_ALLOCATION_GRANULARITY (8u << 20)    /* same 8 MiB pools as cpu/metal */
#define UZU_MAX_INLINE_BYTES 4096u               /* kernel-parameter inline constants (Metal: 4096) */

typedef enum uzu_status {
    UZU_OK = 0,
    UZU_ERROR_CUDA = 1,              /* a CUDA runtime/driver call failed (message has the code) */
    UZU_ERROR_INVALID_ARGUMENT = 2,
    UZU_ERROR_UNSUPPORTED = 3,       /* valid request this backend does not implement (like CpuError::NotSupported) */
    UZU_ERROR_OUT_OF_MEMORY = 4,
    UZU_ERROR_IO = 5,
    UZU_ERROR_NO_DEVICE = 6
} uzu_status;

/* Thread-local message for the last failing call on this thread. */
UZU_API const char* uzu_last_error(void);
UZU_API const char* uzu_version(void);
/* sizeof() of an ABI struct by name ("uzu_matmul_args", ...), 0 if unknown: lets a binding verify its mirror. */
UZU_API size_t uzu_abi_struct_size(const char* name);

/* ---- gpu_types mirrors ------------------------------------------------------------------- */
typedef enum uzu_data_type { UZU_DT_BF16 = 0, UZU_DT_F32 = 1, UZU_DT_F16 = 2 } uzu_data_type; /* data_type.rs:7-36 subset */
/* gpu_types/quantization.rs:9 */
typedef enum uzu_quantization_mode { UZU_QMODE_U4 = 0, UZU_QMODE_I8 = 1, UZU_QMODE_U8 = 2 } uzu_quantization_mode;
/* gpu_types/quantization_method.rs:5 */
typedef enum uzu_quantization_method { UZU_QMETHOD_SCALE_BIAS = 0, UZU_QMETHOD_SCALE_ZERO_POINT = 1, UZU_QMETHOD_SCALE_SYMMETRIC = 2 } uzu_quantization_method;
/* gpu_types/gemm.rs:13 */
typedef enum uzu_gemm_b_prologue_kind {
    UZU_B_FULL_PRECISION = 0, UZU_B_SCALE_BIAS_DEQUANT = 1, UZU_B_SCALE_ZERO_POINT_DEQUANT = 2, UZU_B_SCALE_SYMMETRIC_DEQUANT = 3
} uzu_gemm_b_prologue_kind;
/* gpu_types/gemm.rs:23-29 (bitflags, u32) */
enum { UZU_D_SCALE = 1u << 0, UZU_D_ACCUMULATE = 1u << 1, UZU_D_BIAS = 1u << 2, UZU_D_RHT = 1u << 3, UZU_D_SOFT_CAP = 1u << 4 };
/* gpu_types/activation_type.rs:8-14 */
typedef enum uzu_activation_type { UZU_ACT_SILU = 0, UZU_ACT_GELU_APPROX = 1, UZU_ACT_GELU_EXACT = 2, UZU_ACT_IDENTITY = 3, UZU_ACT_SOFTPLUS = 4 } uzu_activation_type;
/* gpu_types/ring.rs, trie.rs, kv_cache_update.rs */
typedef struct uzu_ring_params { uint32_t ring_offset, ring_length; } uzu_ring_params;
typedef struct uzu_trie_node { uint32_t trie_start, trie_end, height; } uzu_trie_node;
typedef struct uzu_kv_copy { uint32_t source, destination; } uzu_kv_copy;

/* ---- Context: backends/common/context.rs:5-48 ------------------------------------------- */
typedef struct uzu_context uzu_context;
/* Context::new(); the reference opens "the" device -- the ordinal is this backend's extension for
 * one-process-per-GPU multi-GPU runs (negative = CUDA current device / UZU_DEVICE env var). */
UZU_API uzu_status uzu_context_create(int device_ordinal, uzu_context** out);
UZU_API void uzu_context_destroy(uzu_context* ctx);
UZU_API uzu_status uzu_context_synchronize(uzu_context* ctx);
/* Context::peak_memory_usage */
UZU_API uzu_status uzu_context_peak_memory_usage(uzu_context* ctx, size_t* out_bytes);
/* Context::device_capabilities (device_capabilities.rs): bit 0 = SPARSE_BUFFERS */
enum { UZU_CAP_SPARSE_BUFFERS = 1u << 0 };
UZU_API uint32_t uzu_context_device_capabilities(uzu_context* ctx);
/* Context::start_capture / stop_capture -> cudaProfilerStart/Stop (ncu --profile-from-start off) */
UZU_API uzu_status uzu_context_start_capture(uzu_context* ctx, const char* trace_path);
UZU_API uzu_status uzu_context_stop_capture(uzu_context* ctx);
UZU_API int uzu_context_device(uzu_context* ctx);
UZU_API int uzu_context_sm_count(uzu_context* ctx);
/* The context's CUDA stream (cudaStream_t), for callers that time with CUDA events. */
UZU_API void* uzu_context_stream(uzu_context* ctx);

/* ---- Buffers: buffer/mod.rs:11-17, dense.rs:5-7, sparse.rs:5-19 ------------------------- */
typedef struct uzu_buffer uzu_buffer;
typedef enum uzu_buffer_kind {
    /* Context::create_buffer: CPU-addressable (DenseBuffer::cpu_ptr) device memory. CUDA managed memory
     * with preferred location = device; the loader writes through cpu_ptr (parameters/loader.rs:162-179),
     * then uzu_buffer_make_resident migrates the pages to HBM once. */
    UZU_BUFFER_MANAGED = 0,
    /* cpu_available pool allocations (encoder.rs:77-88: token ids, seeds, RoPE tables, sampled tokens):
     * pinned, device-mapped host memory; no page migration on the per-token path. */
    UZU_BUFFER_PINNED_HOST = 1,
    /* scratch that never needs a CPU pointer: plain cudaMalloc; cpu_ptr is NULL. */
    UZU_BUFFER_DEVICE = 2
} uzu_buffer_kind;
UZU_API uzu_status uzu_buffer_create(uzu_context* ctx, size_t size, uzu_buffer_kind kind, uzu_buffer** out);
UZU_API void uzu_buffer_destroy(uzu_buffer* buf);
UZU_API uint64_t uzu_buffer_gpu_ptr(const uzu_buffer* buf);   /* Buffer::gpu_ptr */
UZU_API void* uzu_buffer_cpu_ptr(const uzu_buffer* buf);      /* DenseBuffer::cpu_ptr */
UZU_API size_t uzu_buffer_size(const uzu_buffer* buf);        /* Buffer::size */
UZU_API uzu_status uzu_buffer_make_resident(uzu_context* ctx, uzu_buffer* buf);

typedef struct uzu_sparse_buffer uzu_sparse_buffer;           /* CUDA VMM: cuMemAddressReserve + cuMemMap */
UZU_API uzu_status uzu_sparse_buffer_create(uzu_context* ctx, size_t capacity, uzu_sparse_buffer** out);
UZU_API void uzu_sparse_buffer_destroy(uzu_sparse_buffer* buf);
UZU_API uint64_t uzu_sparse_buffer_gpu_ptr(const uzu_sparse_buffer* buf);
UZU_API size_t uzu_sparse_buffer_size(const uzu_sparse_buffer* buf);
UZU_API size_t uzu_sparse_buffer_page_size_bytes(const uzu_sparse_buffer* buf);
UZU_API uzu_status uzu_sparse_buffer_map(uzu_sparse_buffer* buf, const uint32_t* pages, size_t page_count);
UZU_API uzu_status uzu_sparse_buffer_unmap(uzu_sparse_buffer* buf, const uint32_t* pages, size_t page_count);

/* ---- Command buffer typestate: command_buffer.rs:15-125 ---------------------------------- *
 * Initial -> (start_encoding) Encoding -> (end_encoding) Executable -> (submit) Pending ->
 * (wait_until_completed) Completed. Encoding == enqueueing on the context stream. */
typedef struct uzu_command_buffer uzu_command_buffer;
UZU_API uzu_status uzu_command_buffer_create(uzu_context* ctx, const char* name, uzu_command_buffer** out);
UZU_API void uzu_command_buffer_destroy(uzu_command_buffer* cmd);
UZU_API uzu_status uzu_command_buffer_start_encoding(uzu_command_buffer* cmd);
UZU_API void uzu_command_buffer_encode_copy(uzu_command_buffer* cmd, uint64_t src, uint64_t dst, size_t bytes);
UZU_API void uzu_command_buffer_encode_fill(uzu_command_buffer* cmd, uint64_t dst, size_t bytes, uint8_t value);
UZU_API void uzu_command_buffer_encode_barrier(uzu_command_buffer* cmd, uint32_t after, uint32_t before); /* no-op */
UZU_API void uzu_command_buffer_push_debug_group(uzu_command_buffer* cmd, const char* name);
UZU_API void uzu_command_buffer_pop_debug_group(uzu_command_buffer* cmd);
UZU_API uzu_status uzu_command_buffer_end_encoding(uzu_command_buffer* cmd);
UZU_API uzu_status uzu_command_buffer_submit(uzu_command_buffer* cmd);
UZU_API uzu_status uzu_command_buffer_wait_until_completed(uzu_command_buffer* cmd);
/* CommandBufferCompleted::gpu_execution_time (seconds, CUDA events around the encoded work) */
UZU_API uzu_status uzu_command_buffer_gpu_execution_time(uzu_command_buffer* cmd, double* out_seconds);
/* number of CUDA kernels this command buffer launched (bench.py's gpu_launches evidence) */
UZU_API uint64_t uzu_command_buffer_launch_count(const uzu_command_buffer* cmd);

/* ---- Kernels ------------------------------------------------------------------------------ *
 * One `<name>_args` struct per generated `<Name>Kernel` trait. Fields appear in the declaration
 * order of the CPU `#[kernel]` fn (= the generated `encode` argument order,
 * build/common/traitgen.rs:72-81); `#[specialize]` parameters (moved to `new` by the codegen) are the
 * trailing block. `#[optional]` pointers are 0 when absent. All pointers are device addresses
 * (gpu_ptr + offset). T is bf16 unless a dtype field says otherwise. */

/* MatmulKernel: common/kernel/matmul/{kernel.rs:12-43, arguments.rs:4-15, d_ops.rs:3-9};
 * semantic spec backends/cpu/kernel/matmul/kernel.rs:164-295. */
typedef struct uzu_matmul_args {
    uint64_t a;                    /* MatmulA::FullPrecision values (+offset applied by caller): [m,k] input_dt */
    uint64_t b;                    /* weights: packed codes [n, k*bits/8] or full precision [n,ld]/[k,ld] weights_dt */
    uint64_t b_scales;             /* [n, ceil(k/group_size)] weights_dt */
    uint64_t b_zero_points;        /* ScaleZeroPoint: u8, 4-bit nibble-packed [n, ceil(groups/2)], 8-bit [n, groups] */
    uint64_t b_biases;             /* ScaleBias (MLX): [n, groups] weights_dt */
    uint64_t d;                    /* [m,n] output_dt */
    uint64_t bias;                 /* MatmulDOps::bias [n] weights_dt, 0 = none */
    uint64_t gather_indices;       /* optional u32 [m,n]: output column c of row r reads B row gather[r*n+c] */
    uint64_t rht_factors;          /* MatmulDOps::rht_factors, i32 [n] of +-1 (UZU_D_RHT): D is output-RHT-transformed in place after the
                                    * epilogue and `bias` is added AFTER the transform (cpu/kernel/matmul/kernel.rs:162,285,297-303) */
    uint32_t b_prologue;           /* uzu_gemm_b_prologue_kind */
    uint32_t b_mode;               /* uzu_quantization_mode */
    uint32_t b_group_size;
    uint32_t b_signed_codes;       /* XOR the top code bit (weight_matrix.rs:226-241) */
    uint32_t b_leading_dimension;  /* 0 = default (k if transposed else n) */
    uint32_t b_transpose;          /* quantized B requires 1 (linear/matmul.rs:136) */
    uint32_t d_transform;          /* UZU_D_* mask */
    float ab_scale;
    float soft_cap;
    uint32_t m, n, k;
    uint32_t weights_dt, input_dt, output_dt;   /* MatmulKernel::new arguments */
} uzu_matmul_args;
UZU_API void uzu_matmul_encode(uzu_command_buffer* cmd, const uzu_matmul_args* args);
UZU_API uzu_status uzu_matmul_validate(const uzu_matmul_args* args);   /* MatmulKernel::new + encode error paths */

/* Fused decode linear (extension; m = 1): Matmul whose activation row is produced inside the GEMV, replacing a separate launch of
 *   prologue 1: NormalizationKernel (RMS norm, optional residual add; the updated residual goes to `shortcut_out`, which must differ
 *               from `norm_shortcut_in`; 0 = do not write, for a second linear sharing the same norm),
 *   prologue 2: reserved (GatedActMul moved to the producing GEMV's epilogue, see `epilogue`),
 *   prologue 3: SigmoidGateKernel (x = attn * sigmoid(gate)),
 *   prologue 0: x = matmul.a (only meaningful together with an epilogue).
 * Same arithmetic and rounding points as the standalone kernels. uzu_fused_linear_supported() tells whether the fast path applies;
 * if not, the caller encodes the unfused sequence. `matmul.a` is ignored for prologue != 0. */
typedef struct uzu_fused_linear_args {
    uzu_matmul_args matmul;
    uint32_t prologue;
    uint64_t norm_input, norm_shortcut_in, norm_scales, shortcut_out;
    float norm_epsilon, norm_scale_offset;
    uint32_t norm_residual_add, norm_full_layer;
    uint64_t act_operand; uint32_t act_type;
    uint64_t sg_attn, sg_gate;
    /* epilogue 1: the matmul's n = 2F rows are [up | gate] (DenseMlp's fused up projection, mlp/dense.rs:32-48); instead of the
     * 2F row the kernel stores hidden[j] = GatedActMul(up_j, gate_j) (act_type) as bf16 [F] into matmul.d. prologue may be 0. */
    uint32_t epilogue, reserved0;
    /* optional decode-stream copy of matmul.b (unit-major 4608-byte units, built by the engine at load; 0 = none): the decode GEMV then feeds
     * its shared-memory rings with one TMA bulk copy per stage (UBLKCP) instead of per-lane cp.async. With a stream, prologue 0 + epilogue 0
     * (a plain GEMV over matmul.a) is accepted too. */
    uint64_t decode_stream;
} uzu_fused_linear_args;
/* Tuning sweeps only (tools/): override the decode GEMV's work split; 0 = heuristic. Process-wide, not thread-safe. */
UZU_API void uzu_debug_set_qmv_tuning(int warps_per_tile, int k_slices, int ctas_per_sm, int stages);
/* Prefill path (m >= 64, quantized B, bf16 A): uzu_matmul_encode runs the tcgen05 tensor-core GEMM with an in-kernel dequant stage
 * (csrc/prefill_gemm.cu); UZU_PREFILL_GEMM_MIN_M=<m> moves the threshold, 0 disables it. Probe hook (tools/umma_probe.py only):
 * override the UMMA shared-memory layout (0 = SWIZZLE_128B, 1 = no swizzle; < 0 = built-in default), descriptor words and the
 * instruction descriptor (0 = default), and force the tokens-per-CTA factor mt (0 = heuristic). Process-wide, not thread-safe. */
UZU_API void uzu_debug_set_umma(int layout, uint32_t desc_hi, uint32_t desc_lbo, uint32_t k_step, uint32_t idesc, int mt);
UZU_API int uzu_fused_linear_supported(uzu_context* ctx, const uzu_fused_linear_args* args);
UZU_API void uzu_fused_linear_encode(uzu_command_buffer* cmd, const uzu_fused_linear_args* args);

/* ---- Tensor parallelism (extension; the reference is single-device) -------------------------------------------
 * One process per GPU. Rank 0 creates a 128-byte NCCL unique id, the host program distributes it (bench.py: torch.distributed
 * broadcast), every rank calls uzu_context_tp_init on its context. The engine then loads the checkpoint shard written by
 * uzu_b200/tp.py (config.json "tensor_parallel" block) and inserts the two exchange steps below. NCCL is dlopen'ed at first use
 * (UZU_NCCL_LIB overrides the soname), so the library itself has no link-time dependency on it. */
UZU_API uzu_status uzu_tp_get_unique_id(uint8_t* out128);
UZU_API uzu_status uzu_context_tp_init(uzu_context* ctx, uint32_t rank, uint32_t size, const uint8_t* unique_id128);
UZU_API void uzu_context_tp_destroy(uzu_context* ctx);
UZU_API uint32_t uzu_context_tp_size(const uzu_context* ctx);
UZU_API uint32_t uzu_context_tp_rank(const uzu_context* ctx);
/* Optional peer-memory exchange for small (decode) messages -- one kernel pushes the partials to every peer over NVLink, waits, reduces in
 * rank order and rounds (csrc/tp.cu: tp_p2p_all_reduce_kernel). Every rank: export (allocates the exchange buffer, returns a 64-byte CUDA IPC
 * handle), all-gather the handles with the host program, import (handles of all ranks, rank order), then a host-side barrier. After that
 * uzu_tp_all_reduce_encode uses the peer path for count <= capacity_f32 and NCCL otherwise. NOT yet run on hardware (round 1). */
UZU_API uzu_status uzu_tp_p2p_export(uzu_context* ctx, uint32_t capacity_f32, uint8_t* handle_out64);
UZU_API uzu_status uzu_tp_p2p_import(uzu_context* ctx, const uint8_t* handles_size_x_64);
/* Row-parallel projection epilogue: sum the f32 partials [count] over the ranks in place (ncclAllReduce on the context stream),
 * then round to bf16 once into out_bf16. With a 1-rank context only the rounding runs. */
UZU_API void uzu_tp_all_reduce_encode(uzu_command_buffer* cmd, uint64_t partial_f32, uint32_t count, uint64_t out_bf16);
/* Vocab-parallel readout: every rank's [rows, cols_local] 2-byte logits -> [rows, tp_size * cols_local] on every rank.
 * `scratch` (tp_size * rows * cols_local * 2 bytes) is needed for rows > 1 only. */
typedef struct uzu_tp_all_gather_args { uint64_t src, dst, scratch; uint32_t rows, cols_local; } uzu_tp_all_gather_args;
UZU_API void uzu_tp_all_gather_encode(uzu_command_buffer* cmd, const uzu_tp_all_gather_args* args);

/* NormalizationKernel: backends/cpu/kernel/normalization/normalization.rs:7-49 */
typedef struct uzu_normalization_args {
    uint64_t input;                /* optional(!in_place) */
    uint64_t scales;               /* optional(has_scales), f32 [element_count] */
    uint64_t biases;               /* optional(has_biases), f32 */
    uint64_t output;
    uint64_t shortcut;             /* optional(copy_to_shortcut) */
    uint64_t hadamard_factors;     /* optional(use_hadamard): unsupported */
    uint32_t batch_size, element_count;
    float epsilon, scale_offset, post_layer_scalar;
    /* #[specialize] */
    uint32_t in_place, subtract_mean, full_layer, copy_to_shortcut, residual_add, use_hadamard;
    uint32_t scale_residual_sum, scale_output, has_biases, has_scales;
} uzu_normalization_args;
UZU_API void uzu_normalization_encode(uzu_command_buffer* cmd, const uzu_normalization_args* args);

/* QKVNormKernel: backends/cpu/kernel/attention/qkv_norm.rs:7-35 */
typedef struct uzu_qkv_norm_args {
    uint64_t qkv_input;            /* optional(!in_place) */
    uint64_t scales;               /* optional(has_scales), f32 [head_dim] */
    uint64_t qkv_output;
    uint32_t batch_size, total_heads, head_dim;
    float epsilon, scale_offset;
    uint32_t head_offset, head_count, full_layer;
    uint32_t in_place, has_scales;
} uzu_qkv_norm_args;
UZU_API void uzu_qkv_norm_encode(uzu_command_buffer* cmd, const uzu_qkv_norm_args* args);

/* AttentionPrepareKernel: backends/cpu/kernel/attention/attention_prepare.rs:34-52 */
typedef struct uzu_attention_prepare_args {
    uint64_t qkv, queries;
    uint64_t keys, values;         /* optional(has_kv): cache base; row kv_token_offset + t is written */
    uint64_t cosines, sines;       /* optional(has_rope): f32 [batch_dim, rope_dim] */
    uint32_t num_q_heads, num_kv_heads, head_dim, rope_dim, kv_token_offset, batch_dim;
    uint32_t has_kv, has_rope;
    /* Extension for CUDA-graph replay (0 = reference semantics): device u32* holding the prefix length P. When set,
     * kv_token_offset is read as P and cosines/sines are indexed from row P of a position-major table
     * ([max_positions, rope_dim], the reference's host table evaluated once for every position). */
    uint64_t dynamic_position;
} uzu_attention_prepare_args;
UZU_API void uzu_attention_prepare_encode(uzu_command_buffer* cmd, const uzu_attention_prepare_args* args);

/* Extension: QKVNorm (queries), QKVNorm (keys) and AttentionPrepare in ONE launch (Attention::encode runs them back to back on the
 * same qkv rows, mixer/attention/mode.rs; qkv_norm.rs:7-35 + attention_prepare.rs:34-52). Same arithmetic and rounding points as the
 * three kernels (the normalised row is rounded to bf16 before RoPE exactly as the in-place QKVNorm stores it); the qkv buffer itself is
 * left untouched. head_dim <= 256. A norm with `present == 0` is skipped. */
typedef struct uzu_qk_norm_config {
    uint64_t scales;               /* f32 [head_dim] (has_scales) */
    float epsilon, scale_offset;
    uint32_t present, full_layer, has_scales, reserved0;
} uzu_qk_norm_config;
typedef struct uzu_attention_prepare_norm_args {
    uzu_attention_prepare_args prepare;
    uzu_qk_norm_config q_norm, k_norm;
} uzu_attention_prepare_norm_args;
UZU_API void uzu_attention_prepare_norm_encode(uzu_command_buffer* cmd, const uzu_attention_prepare_norm_args* args);

/* AttentionSinglePassKernel / AttentionTwoPass1Kernel share this argument block:
 * backends/cpu/kernel/attention/attention_single_pass.rs:13-37, attention_two_pass.rs:15-40. */
typedef struct uzu_attention_args {
    uint64_t queries;              /* [num_heads, suffix_length, head_dim] */
    uint64_t keys, values;
    uint64_t out;                  /* single pass: T [suffix, num_heads, D]; two-pass-1: f32 partials [suffix, num_heads, 32, D] */
    uint64_t sums, maxs;           /* two-pass-1 only: f32 [suffix, num_heads, 32] */
    uint32_t gqa_factor, sequence_length;
    uint32_t k_head_stride, k_seq_stride, v_head_stride, v_seq_stride;
    uzu_ring_params ring_params;   /* optional(is_kv_cache_ring) */
    float scale;
    uint64_t trie;                 /* optional(is_trie): uzu_trie_node[suffix] */
    uint32_t sliding_window_size;  /* optional(is_sliding_window) */
    uint64_t sinks;                /* optional(has_sinks): T [num_heads] */
    uint32_t num_heads, suffix_length;
    uint32_t head_dim;             /* HEAD_DIM variant: 64 | 128 | 256 */
    uint32_t has_sinks, is_kv_cache_ring, is_causal, is_trie, is_sliding_window;
    /* Extension for CUDA-graph replay (0 = reference semantics): device u32* holding the prefix length P; when set,
     * sequence_length is read as P + suffix_length and `sequence_length` is only a sizing hint for the KV split. */
    uint64_t dynamic_position;
} uzu_attention_args;
UZU_API void uzu_attention_single_pass_encode(uzu_command_buffer* cmd, const uzu_attention_args* args);
UZU_API void uzu_attention_two_pass1_encode(uzu_command_buffer* cmd, const uzu_attention_args* args);
/* Opt-in tensor-core prefill attention (csrc/attention_prefill.cu; suffix >= 16, plain causal / non-causal, head_dim 64 | 128): taken by
 * uzu_attention_single_pass_encode when UZU_PREFILL_ATTN=1 or after uzu_debug_set_prefill_attention(1); -1 = follow the environment.
 * Round 1: probed on hardware against a float64 softmax (profiles/r1_prefill_attention_probe.txt), oracle tests pending -- the default stays
 * the split-KV kernel. */
UZU_API void uzu_debug_set_prefill_attention(int mode);
/* AttentionTwoPass2Kernel: attention_two_pass.rs:139-149 */
typedef struct uzu_attention_two_pass2_args {
    uint64_t partials, sums, maxs, out;
    uint32_t num_heads, suffix_length, head_dim;
} uzu_attention_two_pass2_args;
UZU_API void uzu_attention_two_pass2_encode(uzu_command_buffer* cmd, const uzu_attention_two_pass2_args* args);

/* KVCacheUpdateKernel: backends/cpu/kernel/attention/kv_cache_update.rs:7-15. `copies` is an inline
 * constant slice in the reference (&[Copy]); here a host pointer copied into the launch (<= 512 copies
 * per call). The flat decode path always passes copy_count = 0 (state.rs:181-199). */
typedef struct uzu_kv_cache_update_args {
    uint64_t in_place_keys, in_place_values;
    const uzu_kv_copy* copies;     /* HOST pointer */
    uint32_t copy_count, element_dim;
} uzu_kv_cache_update_args;
UZU_API void uzu_kv_cache_update_encode(uzu_command_buffer* cmd, const uzu_kv_cache_update_args* args);

/* ActivationTransformKernel (Mirai RHT): backends/cpu/kernel/activation_transform/activation_transform.rs:44-63 (declaration
 * order), semantics :64-136 and mod.rs:9-44; ops = gpu_types ActivationTransformOp. InputRht: out = H32 (s o x) per 32-wide stripe,
 * OutputRht: out = s o (H32 x), H32 = Sylvester Walsh-Hadamard / sqrt(32), s = rht_factors (i32 +-1, [element_count]). Quantize ops run
 * the InputRht transform and emit symmetric int8 codes per activation group (divisor = max|t| / 127), the f32 divisors and -- with
 * group sums -- the integer code sums per sum group. Callers: RHTLinearWrapper::encode_input (encodable_block/linear/rht_wrapper.rs:
 * 214-297) and the matmul's output-RHT epilogue. */
typedef enum uzu_activation_transform_op {
    UZU_ACTIVATION_TRANSFORM_INPUT_RHT = 0, UZU_ACTIVATION_TRANSFORM_OUTPUT_RHT = 1,
    UZU_ACTIVATION_TRANSFORM_QUANTIZE = 2, UZU_ACTIVATION_TRANSFORM_QUANTIZE_WITH_GROUP_SUMS = 3
} uzu_activation_transform_op;
typedef struct uzu_activation_transform_args {
    uint64_t input;                /* optional(!in_place): [batch, element_count] T */
    uint64_t fp_out;               /* optional(InputRht | OutputRht): [batch, element_count] T (the operand itself when in_place) */
    uint64_t q_out;                /* optional(Quantize*): i8 [batch, element_count] */
    uint64_t scales_out;           /* optional(Quantize*): f32 [batch, element_count / activation_scale_group_size] */
    uint64_t group_sums_out;       /* optional(QuantizeWithGroupSums): i32 [batch, element_count / sum_group_size] */
    uint64_t rht_factors;          /* i32 [element_count] */
    uint32_t batch_size, element_count;
    uint32_t ops;                  /* specialize: uzu_activation_transform_op */
    uint32_t in_place;             /* specialize */
    uint32_t activation_scale_group_size, sum_group_size;   /* specialize (multiples of 32, <= 256, nested) */
    uint32_t data_type;            /* T: UZU_DT_BF16 | UZU_DT_F32 */
} uzu_activation_transform_args;
UZU_API void uzu_activation_transform_encode(uzu_command_buffer* cmd, const uzu_activation_transform_args* args);
UZU_API uzu_status uzu_activation_transform_validate(const uzu_activation_transform_args* args);

/* SigmoidGateKernel: backends/cpu/kernel/attention/sigmoid_gate.rs:7-13 */
UZU_API void uzu_sigmoid_gate_encode(uzu_command_buffer* cmd, uint64_t gate, uint64_t output, uint32_t total_elements);

/* GatedActMulKernel (FullPrecision op): backends/cpu/kernel/gated_act_mul/gated_act_mul.rs:13-42 */
typedef struct uzu_gated_act_mul_args {
    uint64_t act_operand;
    uint64_t value_operand;        /* optional(!interleaved) */
    uint64_t fp_out;
    uint32_t gated_dim, batch_dim, value_offset, value_row_stride;
    uint32_t act_type;             /* uzu_activation_type */
    uint32_t interleaved;          /* Quantize ops / Hadamard: unsupported */
} uzu_gated_act_mul_args;
UZU_API void uzu_gated_act_mul_encode(uzu_command_buffer* cmd, const uzu_gated_act_mul_args* args);

/* QuantizedEmbeddingLookupKernel: backends/cpu/kernel/embedding/quant_embedding.rs:11-34 */
typedef struct uzu_quantized_embedding_lookup_args {
    uint64_t token_ids;            /* u32 [batch] */
    uint64_t weights, scales, zero_points, biases, output;
    uint32_t batch_size, vocab_size, model_dim;
    float input_scale;
    uint32_t group_size, quantization_mode, quantization_method;
} uzu_quantized_embedding_lookup_args;
UZU_API void uzu_quantized_embedding_lookup_encode(uzu_command_buffer* cmd, const uzu_quantized_embedding_lookup_args* args);

/* FullPrecisionEmbeddingLookupKernel: embedding/full_precision_embedding.rs:7-17 */
UZU_API void uzu_full_precision_embedding_lookup_encode(uzu_command_buffer* cmd, uint64_t token_ids, uint64_t weights,
                                                        uint64_t output, uint32_t batch_size, uint32_t vocab_size,
                                                        uint32_t model_dim, float input_scale);

/* LogitTransformKernel: logit_transform/logit_transform.rs:7-15 */
UZU_API void uzu_logit_transform_encode(uzu_command_buffer* cmd, uint64_t logits, uint32_t length, float scale,
                                        float soft_cap, uint32_t has_soft_cap);

/* TensorAddScale / TensorCopy / TensorAddBias / TensorAddSwap: backends/cpu/kernel/tensor_{add_scale,copy,add_bias,add_swap} */
UZU_API void uzu_tensor_add_scale_encode(uzu_command_buffer* cmd, uint64_t input /* 0 = in place */, uint64_t bias,
                                         uint64_t output, uint32_t num_cols, uint32_t length, float scale);
UZU_API void uzu_tensor_copy_encode(uzu_command_buffer* cmd, uint64_t src, uint64_t dst, uint32_t length);
UZU_API void uzu_tensor_add_bias_encode(uzu_command_buffer* cmd, uint64_t input /* 0 = in place */, uint64_t bias,
                                        uint64_t output, uint32_t num_cols, uint32_t length);
UZU_API void uzu_tensor_add_swap_encode(uzu_command_buffer* cmd, uint64_t skip_buffer, uint64_t main_buffer, uint32_t length);

/* UnifiedSamplingKernel: backends/cpu/kernel/sampling/unified_sampling.rs:13-33. Token ids are
 * bit-identical to the CPU kernel (Philox4x32-10 stream layout of sampling/gumbel.rs:34-81). */
typedef struct uzu_unified_sampling_args {
    uint64_t logits;               /* T [batch, vocab] */
    uint64_t output;               /* u32 [batch] */
    uint64_t seeds;                /* optional(is_stochastic): u64 [batch] */
    uint64_t bitmask;              /* optional(has_bitmask): u32 [batch, ceil(vocab/32)] */
    float temperature;
    uint32_t top_k;
    float top_p, min_p;
    uint32_t vocab_size, batch_size;
    uint32_t is_stochastic, has_bitmask, has_temperature, has_top_k, has_top_p, has_min_p;
} uzu_unified_sampling_args;
UZU_API void uzu_unified_sampling_encode(uzu_command_buffer* cmd, const uzu_unified_sampling_args* args);

/* DeltaNetConvUpdateKernel / DeltaNetUpdateKernel (Qwen3.5 hybrid layers, decode branch):
 * backends/cpu/kernel/gdn/conv_update.rs:8-20, update.rs:13-30; parameter/state dtypes are f32 as
 * the engine allocates them and Metal declares them. */
typedef struct uzu_delta_net_conv_update_args {
    uint64_t conv_weight, bias /* optional(has_bias) */, in_out, state;
    uint32_t kernel_size, conv_dim, state_stride, has_bias;
} uzu_delta_net_conv_update_args;
UZU_API void uzu_delta_net_conv_update_encode(uzu_command_buffer* cmd, const uzu_delta_net_conv_update_args* args);
typedef struct uzu_delta_net_update_args {
    uint64_t in_proj, a_log, dt_bias, norm_weight, state, out;
    uint32_t num_v_heads, num_k_heads, head_v_dim, key_dim, value_dim;
    float norm_epsilon;
    uint32_t head_k_dim;           /* HEAD_K_DIM variant: 128 */
} uzu_delta_net_update_args;
UZU_API void uzu_delta_net_update_encode(uzu_command_buffer* cmd, const uzu_delta_net_update_args* args);
/* Extension: DeltaNetConvUpdate folded into DeltaNetUpdate (DeltaNet::encode runs them back to back, mixer/delta_net.rs). `update.in_proj`
 * is the RAW projection row (== conv.in_out, which is NOT rewritten); every CTA convolves the q / k / v channels it consumes with the
 * conv kernel's arithmetic (bf16-rounded SiLU output) and the rolling conv state is advanced once per channel. Requires
 * num_v_heads == num_k_heads (each k head is read by exactly one v head's CTA cluster), kernel_size <= 8. */
typedef struct uzu_delta_net_fused_update_args {
    uzu_delta_net_update_args update;
    uzu_delta_net_conv_update_args conv;
} uzu_delta_net_fused_update_args;
UZU_API int uzu_delta_net_fused_update_supported(const uzu_delta_net_fused_update_args* args);
UZU_API void uzu_delta_net_fused_update_encode(uzu_command_buffer* cmd, const uzu_delta_net_fused_update_args* args);
/* Opt-in (UZU_DELTA_PREFILL_KERNEL=1 or uzu_debug_set_delta_prefill(1); -1 = follow the environment): the engine's batched hybrid prefill runs
 * the m-token DeltaNet recurrence of a layer in ONE launch (csrc/deltanet_prefill.cu: one CTA per v-head, state resident in shared memory)
 * instead of one decode-kernel launch per token. NOT yet run on hardware (round 1). */
UZU_API void uzu_debug_set_delta_prefill(int mode);

/* ---- Host engine (C++ mirror of the reference's backend-agnostic Rust host code) ---------- *
 * No Rust toolchain exists in the build image, so the layer above the kernels -- Engine /
 * LanguageModel / LanguageModelStream / Decoder / Transformer / TransformerLayer (engine/mod.rs:15-43,
 * engine/language_model/mod.rs:58-116, engine/language_model/stream/stream.rs:131-782,
 * encodable_block/{decoder,transformer,transformer_layer}.rs) -- is mirrored in C++ inside this
 * library and drives exactly the entry points above, in the reference's op order. */
typedef struct uzu_engine uzu_engine;                 /* Engine<B> + one loaded LanguageModel + its state */
typedef struct uzu_engine_options {
    uint32_t max_context_length;   /* LanguageModel::create_empty_state(max_context_length) */
    uint32_t use_cuda_graph;       /* capture the per-token command list once and replay it */
    uint32_t fused_decode;         /* use the fused decode kernels (norm->GEMV prologue, act->GEMV prologue, ...) */
    uint32_t tp_rank, tp_size;     /* tensor-parallel shard of this process (1 = none) */
    uint64_t reserved[4];
} uzu_engine_options;
typedef enum uzu_sampling_kind { UZU_SAMPLING_GREEDY = 0, UZU_SAMPLING_STOCHASTIC = 1 } uzu_sampling_kind;
typedef struct uzu_sampling_method {      /* encodable_block/sampling/mod.rs:44-56 */
    uint32_t kind;
    uint32_t has_temperature; float temperature;
    uint32_t has_top_k; uint32_t top_k;
    uint32_t has_top_p; float top_p;
    uint32_t has_min_p; float min_p;
    uint64_t seed;
} uzu_sampling_method;
typedef struct uzu_model_info {
    uint32_t model_dim, hidden_dim, vocab_size, num_layers, num_attention_layers, num_delta_net_layers;
    uint64_t weight_bytes_per_token;      /* algorithmic bytes of quantised weights one decode step streams */
    uint64_t kv_bytes_per_token_per_ctx;  /* K+V bytes read per context token per decode step */
    uint64_t state_bytes_per_token;       /* DeltaNet state read+write bytes per decode step */
} uzu_model_info;

UZU_API uzu_status uzu_engine_create(uzu_context* ctx, const char* model_dir, const uzu_engine_options* opts, uzu_engine** out);
UZU_API void uzu_engine_destroy(uzu_engine* e);
UZU_API uzu_status uzu_engine_info(const uzu_engine* e, uzu_model_info* out);
/* LanguageModelState reset (new conversation) */
UZU_API uzu_status uzu_engine_reset(uzu_engine* e);
UZU_API uint32_t uzu_engine_context_length(const uzu_engine* e);
/* Rewind the state to `context_length` tokens (benchmark helper: repeat a decode window). Only valid for
 * attention-only models or when rewinding to the DeltaNet snapshot taken by uzu_engine_snapshot. */
UZU_API uzu_status uzu_engine_snapshot(uzu_engine* e);
UZU_API uzu_status uzu_engine_restore(uzu_engine* e);
/* LanguageModelStream::new: prefill `tokens` (HOST pointer) in chunks of <= 1024, sample the first token.
 * Blocks until done; returns the sampled token. */
UZU_API uzu_status uzu_engine_prefill(uzu_engine* e, const uint32_t* tokens, uint32_t count,
                                      const uzu_sampling_method* sampling, uint32_t* out_token);
/* LanguageModelStream::next(): one decode step. The input token is the previously sampled one, chained on the
 * device (stream.rs:611-615); returns the token sampled by the *previous* pass after waiting for it, keeping
 * one pass in flight (ForwardPassChaining::InFlight, stream.rs:31-79). */
UZU_API uzu_status uzu_engine_next(uzu_engine* e, uint32_t* out_token);
/* Drain the in-flight pass (Drop of the stream). */
UZU_API uzu_status uzu_engine_flush(uzu_engine* e, uint32_t* out_token);
/* Enqueue `steps` chained decode passes without host reads; tokens land in `out_tokens_dev` (u32[steps], device)
 * if non-zero. Used by bench.py's device-resident `value` measurement. Does not synchronize. */
UZU_API uzu_status uzu_engine_decode_device(uzu_engine* e, uint32_t steps, uint64_t out_tokens_dev);
/* Teacher-forced single pass for parity tests: run `count` tokens (HOST) as one flat batch from the current
 * state, copy bf16 logits of rows [row_begin,row_end) to `out_logits` (HOST, u16 bits) and synchronize. */
UZU_API uzu_status uzu_engine_forward(uzu_engine* e, const uint32_t* tokens, uint32_t count, uint32_t row_begin,
                                      uint32_t row_end, uint16_t* out_logits);
/* Multi-sequence batched decode (extension -- the reference decodes ONE sequence; BASELINE config 4 "batch = 8"): `sequences` (<= 16)
 * independent sequence states (own KV caches / DeltaNet states / positions) advance by one token per step and share one pass over the
 * weights (every linear runs with m = sequences rows); attention and the DeltaNet recurrence run per sequence. Greedy sampling.
 *   batch_begin        allocate / reset the sequence states (the engine's own single-sequence state is untouched)
 *   batch_prefill      ordinary chunked prefill of `tokens` (HOST) into sequence `sequence`; returns its first sampled token
 *   batch_step         tokens_in[sequences] (HOST) -> one step for all sequences -> tokens_out[sequences] (HOST)
 *   batch_decode_timed `steps` device-chained steps from first_tokens[sequences] between two CUDA events; seconds for all of them
 *   batch_logits       bf16 logits [sequences, vocab] of the last step (parity tests)
 * NOT yet run on hardware (round 1): orchestration of the parity-tested kernels, plain stream-ordered launches (no CUDA graph yet). */
UZU_API uzu_status uzu_engine_batch_begin(uzu_engine* e, uint32_t sequences);
UZU_API uzu_status uzu_engine_batch_prefill(uzu_engine* e, uint32_t sequence, const uint32_t* tokens, uint32_t count, uint32_t* out_token);
UZU_API uzu_status uzu_engine_batch_step(uzu_engine* e, const uint32_t* tokens_in, uint32_t* tokens_out);
UZU_API uzu_status uzu_engine_batch_decode_timed(uzu_engine* e, const uint32_t* first_tokens, uint32_t steps, double* out_seconds);
UZU_API uzu_status uzu_engine_batch_logits(uzu_engine* e, uint16_t* out_logits);
UZU_API uint32_t uzu_engine_batch_context_length(const uzu_engine* e, uint32_t sequence);
UZU_API uint64_t uzu_engine_launch_count(const uzu_engine* e);   /* kernels launched (incl. graph nodes replayed) */
/* Measurement helpers (bench.py). All time with CUDA events recorded on the engine's stream.
 *  decode_timed: `steps` device-chained decode passes between two events; returns seconds.
 *  step_host:    one decode pass fed from and read back to HOST memory: token_in is copied host->device from pinned
 *                memory, the pass runs, the sampled token is read back (4 B each way) and the call returns it.
 *  time_linears: every weight-streaming GEMV of one decode step (all layers + readout, m = 1) replayed back to back
 *                `iters` times between two events; returns seconds and the number of kernel launches. */
UZU_API uzu_status uzu_engine_decode_timed(uzu_engine* e, uint32_t steps, double* out_seconds);
UZU_API uzu_status uzu_engine_step_host(uzu_engine* e, uint32_t token_in, uint32_t* token_out);
/* Persistent decode kernel (extension; uzu_b200/csrc/decode_mega.cu): when the model is covered (uniform int4 / int8 block quantisation,
 * greedy sampling, one GPU) a decode step is ONE cooperative launch that runs the whole per-token pass (stream.rs:363-782) with the weight
 * stream fed by TMA bulk copies across phase boundaries. decode_mode: 1 = persistent kernel, 0 = per-kernel path (reason says why).
 * set_decode_mode(0) forces the per-kernel path (A/B runs, parity tests); last_logits copies the [vocab] bf16 logits of the latest step. */
UZU_API int uzu_engine_decode_mode(const uzu_engine* e);
UZU_API const char* uzu_engine_decode_mode_reason(const uzu_engine* e);
UZU_API uzu_status uzu_engine_set_decode_mode(uzu_engine* e, int persistent);
UZU_API uzu_status uzu_engine_last_logits(uzu_engine* e, uint16_t* out_logits);

/* Speculative (trie) decode, the verify half. Replaces what LanguageModelStream::generate does between "propose a trie" and
 * "accept a path" (engine/language_model/stream/stream.rs:550-657 and :436-520): the host (trie.rs, unchanged) linearizes the proposal into
 * `count` <= 16 nodes -- token ids, gpu_types::trie::TrieNode {trie_start, trie_end, height} from FlatTrie::token_subtrie_ranges
 * (trie.rs:211-222) and per-node seeds (token_seeds, :224-226; NULL = PRng::derive(context + height)) -- and
 *  trie_pass:   runs Decoder::encode over the nodes from the current state (BatchTopology::new(nodes, full_accept = false)): RoPE position =
 *               context + height (transformer.rs:248), K/V of node i appended at row context + i, attention masked by subtrie range
 *               (mask.rs:21-29), logits and one sampled token for EVERY node (stream.rs:640). out_tokens[count] (HOST) receives the sampled
 *               ids, out_logits (HOST, optional) the bf16 logits [count, vocab]. Nothing is accepted: the context length does not move and
 *               every other engine call fails until trie_accept.
 *  trie_accept: TransformerState::encode_accept(accepted_indices) (transformer.rs:56-75 -> mixer/attention/state.rs:174-237): indices
 *               strictly increasing (FlatTrie::accept, trie.rs:262-296, yields them that way); KV row context + accepted[i] moves to
 *               context + i through KVCacheUpdateKernel, context grows by `count`; `next_token` (the last verified node's sampled token)
 *               becomes the device-chained input of the next pass (ForwardPassChaining::Constant).
 *  speculation_supported: Mixer::speculation_supported (mixer/mod.rs, delta_net.rs:442-444): 0 for hybrid (DeltaNet) models -- the tree-
 *               verify core of the DeltaNet mixer is not built -- 1 for attention-only models. */
UZU_API int uzu_engine_speculation_supported(const uzu_engine* e);
UZU_API uzu_status uzu_engine_trie_pass(uzu_engine* e, const uint32_t* tokens, const uzu_trie_node* nodes, const uint64_t* seeds, uint32_t count,
                                        const uzu_sampling_method* sampling, uint32_t* out_tokens, uint16_t* out_logits);
UZU_API uzu_status uzu_engine_trie_accept(uzu_engine* e, const uint32_t* accepted_indices, uint32_t count, uint32_t next_token);
/* measurement helper: run ONE persistent decode step and return, for CTA `cta`, eight SM-clock stamps per phase ([0] phase start, [1] activation
 * row staged / attention prepared, [2..5] phase-specific, [6] phase body done, [7] grid barrier passed; 0 = not stamped) and the phase kinds (1 GEMV, 2 prepare, 3 attention, 4 act, 5/6 DeltaNet, 7 logits, 8 finish) */
UZU_API uzu_status uzu_engine_debug_decode_trace(uzu_engine* e, uint32_t cta, uint32_t capacity, uint32_t* out_kinds, uint64_t* out_cycles, uint32_t* out_nops);
UZU_API uzu_status uzu_engine_time_linears(uzu_engine* e, uint32_t iters, double* out_seconds, uint64_t* out_launches);
/*  time_prefill_linears: every linear of one PREFILL pass over m rows (all layers, no readout) back to back; returns seconds per pass
 *                and the useful flops of one pass (2*m*N*K summed): the tensor-core GEMM's achieved TFLOP/s = flops / seconds. */
UZU_API uzu_status uzu_engine_time_prefill_linears(uzu_engine* e, uint32_t m, uint32_t iters, double* out_seconds, double* out_flops);
/* same, restricted to a subset: bit 0 mixer input projections (qkv / gate / in_proj), 1 mixer output projection, 2 MLP up, 3 MLP down,
 * 4 readout, 5 MLP up with the GatedActMul epilogue, 6 attention mix (q/k norm, RoPE + KV append, attention core, gate) at the
 * current context length, 7 DeltaNet conv + state update; bit 31: plain stream-ordered launches (no programmatic dependent launch) */
UZU_API uzu_status uzu_engine_time_linears_select(uzu_engine* e, uint32_t iters, uint32_t select, double* out_seconds, uint64_t* out_launches);

#ifdef __cplusplus
}
#endif
#endif /* UZU_B200_H */
