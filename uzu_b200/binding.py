"""ctypes binding of libuzu_b200.so (include/uzu_b200.h) for tests and bench.py.

This is the kind of stub a host-language binding adds: structs mirrored field for field
(verified at import against uzu_abi_struct_size), opaque handles, status codes turned into exceptions.
There is no fallback: a missing library raises at load(), a missing GPU raises at Context().
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

LIB_PATH = Path(os.environ.get("UZU_B200_LIB") or Path(__file__).resolve().parent / "lib" / "libuzu_b200.so")   # override: kernel experiments

u32, u64, f32 = C.c_uint32, C.c_uint64, C.c_float

# enums (include/uzu_b200.h)
DT_BF16, DT_F32 = 0, 1
QMODE_U4, QMODE_I8, QMODE_U8 = 0, 1, 2
QMETHOD_SCALE_BIAS, QMETHOD_SCALE_ZERO_POINT, QMETHOD_SCALE_SYMMETRIC = 0, 1, 2
B_FULL_PRECISION, B_SCALE_BIAS, B_SCALE_ZERO_POINT, B_SCALE_SYMMETRIC = 0, 1, 2, 3
D_SCALE, D_ACCUMULATE, D_BIAS, D_RHT, D_SOFT_CAP = 1, 2, 4, 8, 16
ACT_SILU, ACT_GELU_APPROX, ACT_GELU_EXACT, ACT_IDENTITY, ACT_SOFTPLUS = 0, 1, 2, 3, 4
BUFFER_MANAGED, BUFFER_PINNED_HOST, BUFFER_DEVICE = 0, 1, 2
SAMPLING_GREEDY, SAMPLING_STOCHASTIC = 0, 1


class UzuError(RuntimeError):
    pass


def _struct(name, fields):
    return type(name, (C.Structure,), {"_fields_": fields})


RingParams = _struct("uzu_ring_params", [("ring_offset", u32), ("ring_length", u32)])
TrieNode = _struct("uzu_trie_node", [("trie_start", u32), ("trie_end", u32), ("height", u32)])
KvCopy = _struct("uzu_kv_copy", [("source", u32), ("destination", u32)])

MatmulArgs = _struct("uzu_matmul_args", [
    ("a", u64), ("b", u64), ("b_scales", u64), ("b_zero_points", u64), ("b_biases", u64), ("d", u64), ("bias", u64),
    ("gather_indices", u64), ("rht_factors", u64), ("b_prologue", u32), ("b_mode", u32), ("b_group_size", u32), ("b_signed_codes", u32),
    ("b_leading_dimension", u32), ("b_transpose", u32), ("d_transform", u32), ("ab_scale", f32), ("soft_cap", f32),
    ("m", u32), ("n", u32), ("k", u32), ("weights_dt", u32), ("input_dt", u32), ("output_dt", u32)])

ActivationTransformArgs = _struct("uzu_activation_transform_args", [
    ("input", u64), ("fp_out", u64), ("q_out", u64), ("scales_out", u64), ("group_sums_out", u64), ("rht_factors", u64),
    ("batch_size", u32), ("element_count", u32), ("ops", u32), ("in_place", u32), ("activation_scale_group_size", u32),
    ("sum_group_size", u32), ("data_type", u32)])

NormalizationArgs = _struct("uzu_normalization_args", [
    ("input", u64), ("scales", u64), ("biases", u64), ("output", u64), ("shortcut", u64), ("hadamard_factors", u64),
    ("batch_size", u32), ("element_count", u32), ("epsilon", f32), ("scale_offset", f32), ("post_layer_scalar", f32),
    ("in_place", u32), ("subtract_mean", u32), ("full_layer", u32), ("copy_to_shortcut", u32), ("residual_add", u32),
    ("use_hadamard", u32), ("scale_residual_sum", u32), ("scale_output", u32), ("has_biases", u32), ("has_scales", u32)])

QkvNormArgs = _struct("uzu_qkv_norm_args", [
    ("qkv_input", u64), ("scales", u64), ("qkv_output", u64), ("batch_size", u32), ("total_heads", u32), ("head_dim", u32),
    ("epsilon", f32), ("scale_offset", f32), ("head_offset", u32), ("head_count", u32), ("full_layer", u32),
    ("in_place", u32), ("has_scales", u32)])

AttentionPrepareArgs = _struct("uzu_attention_prepare_args", [
    ("qkv", u64), ("queries", u64), ("keys", u64), ("values", u64), ("cosines", u64), ("sines", u64),
    ("num_q_heads", u32), ("num_kv_heads", u32), ("head_dim", u32), ("rope_dim", u32), ("kv_token_offset", u32),
    ("batch_dim", u32), ("has_kv", u32), ("has_rope", u32), ("dynamic_position", u64)])

QkNormConfig = _struct("uzu_qk_norm_config", [
    ("scales", u64), ("epsilon", f32), ("scale_offset", f32), ("present", u32), ("full_layer", u32), ("has_scales", u32), ("reserved0", u32)])

AttentionPrepareNormArgs = _struct("uzu_attention_prepare_norm_args", [
    ("prepare", AttentionPrepareArgs), ("q_norm", QkNormConfig), ("k_norm", QkNormConfig)])

AttentionArgs = _struct("uzu_attention_args", [
    ("queries", u64), ("keys", u64), ("values", u64), ("out", u64), ("sums", u64), ("maxs", u64),
    ("gqa_factor", u32), ("sequence_length", u32), ("k_head_stride", u32), ("k_seq_stride", u32), ("v_head_stride", u32),
    ("v_seq_stride", u32), ("ring_params", RingParams), ("scale", f32), ("trie", u64), ("sliding_window_size", u32),
    ("sinks", u64), ("num_heads", u32), ("suffix_length", u32), ("head_dim", u32), ("has_sinks", u32),
    ("is_kv_cache_ring", u32), ("is_causal", u32), ("is_trie", u32), ("is_sliding_window", u32), ("dynamic_position", u64)])

AttentionTwoPass2Args = _struct("uzu_attention_two_pass2_args", [
    ("partials", u64), ("sums", u64), ("maxs", u64), ("out", u64), ("num_heads", u32), ("suffix_length", u32), ("head_dim", u32)])

KvCacheUpdateArgs = _struct("uzu_kv_cache_update_args", [
    ("in_place_keys", u64), ("in_place_values", u64), ("copies", C.POINTER(KvCopy)), ("copy_count", u32), ("element_dim", u32)])

GatedActMulArgs = _struct("uzu_gated_act_mul_args", [
    ("act_operand", u64), ("value_operand", u64), ("fp_out", u64), ("gated_dim", u32), ("batch_dim", u32),
    ("value_offset", u32), ("value_row_stride", u32), ("act_type", u32), ("interleaved", u32)])

QuantizedEmbeddingLookupArgs = _struct("uzu_quantized_embedding_lookup_args", [
    ("token_ids", u64), ("weights", u64), ("scales", u64), ("zero_points", u64), ("biases", u64), ("output", u64),
    ("batch_size", u32), ("vocab_size", u32), ("model_dim", u32), ("input_scale", f32), ("group_size", u32),
    ("quantization_mode", u32), ("quantization_method", u32)])

UnifiedSamplingArgs = _struct("uzu_unified_sampling_args", [
    ("logits", u64), ("output", u64), ("seeds", u64), ("bitmask", u64), ("temperature", f32), ("top_k", u32),
    ("top_p", f32), ("min_p", f32), ("vocab_size", u32), ("batch_size", u32), ("is_stochastic", u32), ("has_bitmask", u32),
    ("has_temperature", u32), ("has_top_k", u32), ("has_top_p", u32), ("has_min_p", u32)])

DeltaNetConvUpdateArgs = _struct("uzu_delta_net_conv_update_args", [
    ("conv_weight", u64), ("bias", u64), ("in_out", u64), ("state", u64), ("kernel_size", u32), ("conv_dim", u32),
    ("state_stride", u32), ("has_bias", u32)])

DeltaNetUpdateArgs = _struct("uzu_delta_net_update_args", [
    ("in_proj", u64), ("a_log", u64), ("dt_bias", u64), ("norm_weight", u64), ("state", u64), ("out", u64),
    ("num_v_heads", u32), ("num_k_heads", u32), ("head_v_dim", u32), ("key_dim", u32), ("value_dim", u32),
    ("norm_epsilon", f32), ("head_k_dim", u32)])

DeltaNetFusedUpdateArgs = _struct("uzu_delta_net_fused_update_args", [("update", DeltaNetUpdateArgs), ("conv", DeltaNetConvUpdateArgs)])

EngineOptions = _struct("uzu_engine_options", [
    ("max_context_length", u32), ("use_cuda_graph", u32), ("fused_decode", u32), ("tp_rank", u32), ("tp_size", u32),
    ("reserved", u64 * 4)])

SamplingMethod = _struct("uzu_sampling_method", [
    ("kind", u32), ("has_temperature", u32), ("temperature", f32), ("has_top_k", u32), ("top_k", u32), ("has_top_p", u32),
    ("top_p", f32), ("has_min_p", u32), ("min_p", f32), ("seed", u64)])

ModelInfo = _struct("uzu_model_info", [
    ("model_dim", u32), ("hidden_dim", u32), ("vocab_size", u32), ("num_layers", u32), ("num_attention_layers", u32),
    ("num_delta_net_layers", u32), ("weight_bytes_per_token", u64), ("kv_bytes_per_token_per_ctx", u64),
    ("state_bytes_per_token", u64)])

TpAllGatherArgs = _struct("uzu_tp_all_gather_args", [("src", u64), ("dst", u64), ("scratch", u64), ("rows", u32), ("cols_local", u32)])

FusedLinearArgs = _struct("uzu_fused_linear_args", [
    ("matmul", MatmulArgs), ("prologue", u32), ("norm_input", u64), ("norm_shortcut_in", u64), ("norm_scales", u64), ("shortcut_out", u64),
    ("norm_epsilon", f32), ("norm_scale_offset", f32), ("norm_residual_add", u32), ("norm_full_layer", u32), ("act_operand", u64),
    ("act_type", u32), ("sg_attn", u64), ("sg_gate", u64), ("epilogue", u32), ("reserved0", u32), ("decode_stream", u64)])

ABI_STRUCTS = [DeltaNetFusedUpdateArgs, QkNormConfig, AttentionPrepareNormArgs, FusedLinearArgs, RingParams, TrieNode, KvCopy, MatmulArgs, NormalizationArgs, QkvNormArgs, AttentionPrepareArgs, AttentionArgs,
               AttentionTwoPass2Args, KvCacheUpdateArgs, GatedActMulArgs, QuantizedEmbeddingLookupArgs, UnifiedSamplingArgs,
               DeltaNetConvUpdateArgs, DeltaNetUpdateArgs, EngineOptions, SamplingMethod, ModelInfo, TpAllGatherArgs]

# every symbol include/uzu_b200.h declares (tests/test_abi.py checks the library exports all of them)
EXPORTS = """uzu_last_error uzu_version uzu_abi_struct_size uzu_context_create uzu_context_destroy uzu_context_synchronize
uzu_context_peak_memory_usage uzu_context_device_capabilities uzu_context_start_capture uzu_context_stop_capture
uzu_context_device uzu_context_sm_count uzu_context_stream uzu_buffer_create uzu_buffer_destroy uzu_buffer_gpu_ptr
uzu_buffer_cpu_ptr uzu_buffer_size uzu_buffer_make_resident uzu_sparse_buffer_create uzu_sparse_buffer_destroy
uzu_sparse_buffer_gpu_ptr uzu_sparse_buffer_size uzu_sparse_buffer_page_size_bytes uzu_sparse_buffer_map
uzu_sparse_buffer_unmap uzu_command_buffer_create uzu_command_buffer_destroy uzu_command_buffer_start_encoding
uzu_command_buffer_encode_copy uzu_command_buffer_encode_fill uzu_command_buffer_encode_barrier
uzu_command_buffer_push_debug_group uzu_command_buffer_pop_debug_group uzu_command_buffer_end_encoding
uzu_command_buffer_submit uzu_command_buffer_wait_until_completed uzu_command_buffer_gpu_execution_time
uzu_command_buffer_launch_count uzu_matmul_encode uzu_matmul_validate uzu_normalization_encode uzu_qkv_norm_encode
uzu_attention_prepare_encode uzu_attention_prepare_norm_encode uzu_attention_single_pass_encode uzu_attention_two_pass1_encode uzu_attention_two_pass2_encode
uzu_kv_cache_update_encode uzu_sigmoid_gate_encode uzu_gated_act_mul_encode uzu_quantized_embedding_lookup_encode
uzu_full_precision_embedding_lookup_encode uzu_logit_transform_encode uzu_tensor_add_scale_encode uzu_tensor_copy_encode
uzu_tensor_add_bias_encode uzu_tensor_add_swap_encode uzu_unified_sampling_encode uzu_delta_net_conv_update_encode
uzu_delta_net_update_encode uzu_engine_create uzu_engine_destroy uzu_engine_info uzu_engine_reset uzu_engine_context_length
uzu_engine_snapshot uzu_engine_restore uzu_engine_prefill uzu_engine_next uzu_engine_flush uzu_engine_decode_device
uzu_engine_forward uzu_engine_batch_begin uzu_engine_batch_prefill uzu_engine_batch_step uzu_engine_batch_decode_timed
uzu_engine_batch_logits uzu_engine_batch_context_length uzu_engine_launch_count uzu_engine_decode_timed uzu_engine_step_host
uzu_delta_net_fused_update_supported uzu_delta_net_fused_update_encode uzu_engine_time_linears uzu_engine_time_prefill_linears uzu_engine_time_linears_select uzu_debug_set_qmv_tuning uzu_debug_set_delta_prefill uzu_debug_set_prefill_attention uzu_debug_set_umma uzu_tp_get_unique_id uzu_context_tp_init uzu_context_tp_destroy uzu_context_tp_size
uzu_context_tp_rank uzu_tp_p2p_export uzu_tp_p2p_import uzu_tp_all_reduce_encode uzu_tp_all_gather_encode uzu_fused_linear_supported uzu_fused_linear_encode
uzu_engine_decode_mode uzu_engine_decode_mode_reason uzu_engine_set_decode_mode uzu_engine_last_logits uzu_engine_debug_decode_trace
uzu_engine_speculation_supported uzu_engine_trie_pass uzu_engine_trie_accept uzu_activation_transform_encode
uzu_activation_transform_validate""".split()

_lib = None


def load() -> C.CDLL:
    """Load libuzu_b200.so (building nothing: run uzu_b200.build / __graft_entry__.build first)."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise UzuError(f"{LIB_PATH} is missing: build it with `python -m uzu_b200.build` (there is no CPU fallback)")
    lib = C.CDLL(str(LIB_PATH))
    lib.uzu_last_error.restype = C.c_char_p
    lib.uzu_version.restype = C.c_char_p
    lib.uzu_abi_struct_size.restype = C.c_size_t
    lib.uzu_abi_struct_size.argtypes = [C.c_char_p]
    for st in ABI_STRUCTS:
        n = lib.uzu_abi_struct_size(st.__name__.encode())
        if n != C.sizeof(st):
            raise UzuError(f"ABI mismatch: {st.__name__} is {n} bytes in the library, {C.sizeof(st)} in the binding")
    vp = C.c_void_p
    sigs = {
        "uzu_context_create": (C.c_int, [C.c_int, C.POINTER(vp)]),
        "uzu_context_destroy": (None, [vp]),
        "uzu_context_synchronize": (C.c_int, [vp]),
        "uzu_context_peak_memory_usage": (C.c_int, [vp, C.POINTER(C.c_size_t)]),
        "uzu_context_device_capabilities": (u32, [vp]),
        "uzu_context_start_capture": (C.c_int, [vp, C.c_char_p]),
        "uzu_context_stop_capture": (C.c_int, [vp]),
        "uzu_context_device": (C.c_int, [vp]),
        "uzu_context_sm_count": (C.c_int, [vp]),
        "uzu_context_stream": (vp, [vp]),
        "uzu_buffer_create": (C.c_int, [vp, C.c_size_t, C.c_int, C.POINTER(vp)]),
        "uzu_buffer_destroy": (None, [vp]),
        "uzu_buffer_gpu_ptr": (u64, [vp]),
        "uzu_buffer_cpu_ptr": (vp, [vp]),
        "uzu_buffer_size": (C.c_size_t, [vp]),
        "uzu_buffer_make_resident": (C.c_int, [vp, vp]),
        "uzu_sparse_buffer_create": (C.c_int, [vp, C.c_size_t, C.POINTER(vp)]),
        "uzu_sparse_buffer_destroy": (None, [vp]),
        "uzu_sparse_buffer_gpu_ptr": (u64, [vp]),
        "uzu_sparse_buffer_size": (C.c_size_t, [vp]),
        "uzu_sparse_buffer_page_size_bytes": (C.c_size_t, [vp]),
        "uzu_sparse_buffer_map": (C.c_int, [vp, C.POINTER(u32), C.c_size_t]),
        "uzu_sparse_buffer_unmap": (C.c_int, [vp, C.POINTER(u32), C.c_size_t]),
        "uzu_command_buffer_create": (C.c_int, [vp, C.c_char_p, C.POINTER(vp)]),
        "uzu_command_buffer_destroy": (None, [vp]),
        "uzu_command_buffer_start_encoding": (C.c_int, [vp]),
        "uzu_command_buffer_encode_copy": (None, [vp, u64, u64, C.c_size_t]),
        "uzu_command_buffer_encode_fill": (None, [vp, u64, C.c_size_t, C.c_uint8]),
        "uzu_command_buffer_encode_barrier": (None, [vp, u32, u32]),
        "uzu_command_buffer_push_debug_group": (None, [vp, C.c_char_p]),
        "uzu_command_buffer_pop_debug_group": (None, [vp]),
        "uzu_command_buffer_end_encoding": (C.c_int, [vp]),
        "uzu_command_buffer_submit": (C.c_int, [vp]),
        "uzu_command_buffer_wait_until_completed": (C.c_int, [vp]),
        "uzu_command_buffer_gpu_execution_time": (C.c_int, [vp, C.POINTER(C.c_double)]),
        "uzu_command_buffer_launch_count": (u64, [vp]),
        "uzu_matmul_encode": (None, [vp, C.POINTER(MatmulArgs)]),
        "uzu_matmul_validate": (C.c_int, [C.POINTER(MatmulArgs)]),
        "uzu_activation_transform_encode": (None, [vp, C.POINTER(ActivationTransformArgs)]),
        "uzu_activation_transform_validate": (C.c_int, [C.POINTER(ActivationTransformArgs)]),
        "uzu_normalization_encode": (None, [vp, C.POINTER(NormalizationArgs)]),
        "uzu_qkv_norm_encode": (None, [vp, C.POINTER(QkvNormArgs)]),
        "uzu_attention_prepare_encode": (None, [vp, C.POINTER(AttentionPrepareArgs)]),
        "uzu_attention_prepare_norm_encode": (None, [vp, C.POINTER(AttentionPrepareNormArgs)]),
        "uzu_delta_net_fused_update_supported": (C.c_int, [C.POINTER(DeltaNetFusedUpdateArgs)]),
        "uzu_delta_net_fused_update_encode": (None, [vp, C.POINTER(DeltaNetFusedUpdateArgs)]),
        "uzu_attention_single_pass_encode": (None, [vp, C.POINTER(AttentionArgs)]),
        "uzu_attention_two_pass1_encode": (None, [vp, C.POINTER(AttentionArgs)]),
        "uzu_attention_two_pass2_encode": (None, [vp, C.POINTER(AttentionTwoPass2Args)]),
        "uzu_kv_cache_update_encode": (None, [vp, C.POINTER(KvCacheUpdateArgs)]),
        "uzu_sigmoid_gate_encode": (None, [vp, u64, u64, u32]),
        "uzu_gated_act_mul_encode": (None, [vp, C.POINTER(GatedActMulArgs)]),
        "uzu_quantized_embedding_lookup_encode": (None, [vp, C.POINTER(QuantizedEmbeddingLookupArgs)]),
        "uzu_full_precision_embedding_lookup_encode": (None, [vp, u64, u64, u64, u32, u32, u32, f32]),
        "uzu_logit_transform_encode": (None, [vp, u64, u32, f32, f32, u32]),
        "uzu_tensor_add_scale_encode": (None, [vp, u64, u64, u64, u32, u32, f32]),
        "uzu_tensor_copy_encode": (None, [vp, u64, u64, u32]),
        "uzu_tensor_add_bias_encode": (None, [vp, u64, u64, u64, u32, u32]),
        "uzu_tensor_add_swap_encode": (None, [vp, u64, u64, u32]),
        "uzu_unified_sampling_encode": (None, [vp, C.POINTER(UnifiedSamplingArgs)]),
        "uzu_delta_net_conv_update_encode": (None, [vp, C.POINTER(DeltaNetConvUpdateArgs)]),
        "uzu_delta_net_update_encode": (None, [vp, C.POINTER(DeltaNetUpdateArgs)]),
        "uzu_engine_create": (C.c_int, [vp, C.c_char_p, C.POINTER(EngineOptions), C.POINTER(vp)]),
        "uzu_engine_destroy": (None, [vp]),
        "uzu_engine_info": (C.c_int, [vp, C.POINTER(ModelInfo)]),
        "uzu_engine_reset": (C.c_int, [vp]),
        "uzu_engine_context_length": (u32, [vp]),
        "uzu_engine_snapshot": (C.c_int, [vp]),
        "uzu_engine_restore": (C.c_int, [vp]),
        "uzu_engine_prefill": (C.c_int, [vp, C.POINTER(u32), u32, C.POINTER(SamplingMethod), C.POINTER(u32)]),
        "uzu_engine_next": (C.c_int, [vp, C.POINTER(u32)]),
        "uzu_engine_flush": (C.c_int, [vp, C.POINTER(u32)]),
        "uzu_engine_decode_device": (C.c_int, [vp, u32, u64]),
        "uzu_engine_forward": (C.c_int, [vp, C.POINTER(u32), u32, u32, u32, C.POINTER(C.c_uint16)]),
        "uzu_engine_launch_count": (u64, [vp]),
        "uzu_engine_batch_begin": (C.c_int, [vp, u32]),
        "uzu_engine_batch_prefill": (C.c_int, [vp, u32, C.POINTER(u32), u32, C.POINTER(u32)]),
        "uzu_engine_batch_step": (C.c_int, [vp, C.POINTER(u32), C.POINTER(u32)]),
        "uzu_engine_batch_decode_timed": (C.c_int, [vp, C.POINTER(u32), u32, C.POINTER(C.c_double)]),
        "uzu_engine_batch_logits": (C.c_int, [vp, C.POINTER(C.c_uint16)]),
        "uzu_engine_batch_context_length": (u32, [vp, u32]),
        "uzu_engine_decode_timed": (C.c_int, [vp, u32, C.POINTER(C.c_double)]),
        "uzu_engine_step_host": (C.c_int, [vp, u32, C.POINTER(u32)]),
        "uzu_engine_decode_mode": (C.c_int, [vp]),
        "uzu_engine_decode_mode_reason": (C.c_char_p, [vp]),
        "uzu_engine_set_decode_mode": (C.c_int, [vp, C.c_int]),
        "uzu_engine_last_logits": (C.c_int, [vp, C.POINTER(C.c_uint16)]),
        "uzu_engine_speculation_supported": (C.c_int, [vp]),
        "uzu_engine_trie_pass": (C.c_int, [vp, C.POINTER(u32), vp, C.POINTER(u64), u32, C.POINTER(SamplingMethod), C.POINTER(u32), C.POINTER(C.c_uint16)]),
        "uzu_engine_trie_accept": (C.c_int, [vp, C.POINTER(u32), u32, u32]),
        "uzu_engine_debug_decode_trace": (C.c_int, [vp, u32, u32, C.POINTER(u32), C.POINTER(u64), C.POINTER(u32)]),
        "uzu_engine_time_linears": (C.c_int, [vp, u32, C.POINTER(C.c_double), C.POINTER(u64)]),
        "uzu_engine_time_prefill_linears": (C.c_int, [vp, u32, u32, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
        "uzu_engine_time_linears_select": (C.c_int, [vp, u32, u32, C.POINTER(C.c_double), C.POINTER(u64)]),
        "uzu_debug_set_qmv_tuning": (None, [C.c_int, C.c_int, C.c_int, C.c_int]),
        "uzu_debug_set_prefill_attention": (None, [C.c_int]),
        "uzu_debug_set_delta_prefill": (None, [C.c_int]),
        "uzu_debug_set_umma": (None, [C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int]),
        "uzu_tp_get_unique_id": (C.c_int, [C.POINTER(C.c_uint8)]),
        "uzu_context_tp_init": (C.c_int, [vp, u32, u32, C.POINTER(C.c_uint8)]),
        "uzu_context_tp_destroy": (None, [vp]),
        "uzu_context_tp_size": (u32, [vp]),
        "uzu_context_tp_rank": (u32, [vp]),
        "uzu_tp_p2p_export": (C.c_int, [vp, u32, C.POINTER(C.c_uint8)]),
        "uzu_tp_p2p_import": (C.c_int, [vp, C.POINTER(C.c_uint8)]),
        "uzu_tp_all_reduce_encode": (None, [vp, u64, u32, u64]),
        "uzu_tp_all_gather_encode": (None, [vp, C.POINTER(TpAllGatherArgs)]),
        "uzu_fused_linear_supported": (C.c_int, [vp, C.POINTER(FusedLinearArgs)]),
        "uzu_fused_linear_encode": (None, [vp, C.POINTER(FusedLinearArgs)]),
    }
    for name, (res, args) in sigs.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def tp_unique_id() -> bytes:
    """ncclGetUniqueId through the library (rank 0); hand the 128 bytes to every rank's Context.tp_init."""
    buf = (C.c_uint8 * 128)()
    _check(load().uzu_tp_get_unique_id(buf))
    return bytes(buf)


def _check(status: int):
    if status != 0:
        raise UzuError(f"uzu status {status}: {load().uzu_last_error().decode(errors='replace')}")


class Context:
    def __init__(self, device: int = -1):
        self.lib = load()
        h = C.c_void_p()
        _check(self.lib.uzu_context_create(device, C.byref(h)))
        self.h = h
        self._buffers = []

    def close(self):
        if self.h:
            for b in self._buffers:
                b.close()
            self._buffers = []
            self.lib.uzu_context_destroy(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def synchronize(self):
        _check(self.lib.uzu_context_synchronize(self.h))

    def tp_p2p_export(self, capacity_f32: int) -> bytes:
        """Allocate this rank's peer-memory exchange buffer; returns its 64-byte CUDA IPC handle (all-gather it, then tp_p2p_import)."""
        buf = (C.c_uint8 * 64)()
        _check(self.lib.uzu_tp_p2p_export(self.h, capacity_f32, buf))
        return bytes(buf)

    def tp_p2p_import(self, handles: list):
        """`handles` = the 64-byte handles of all ranks in rank order. Barrier all ranks afterwards, before the first exchange."""
        blob = b"".join(bytes(h) for h in handles)
        buf = (C.c_uint8 * len(blob)).from_buffer_copy(blob)
        _check(self.lib.uzu_tp_p2p_import(self.h, buf))

    def tp_init(self, rank: int, size: int, unique_id: bytes):
        """Join the tensor-parallel group: `unique_id` = tp_unique_id() of rank 0, distributed by the host program."""
        buf = (C.c_uint8 * 128).from_buffer_copy(bytes(unique_id))
        _check(self.lib.uzu_context_tp_init(self.h, rank, size, buf))

    @property
    def sm_count(self):
        return self.lib.uzu_context_sm_count(self.h)

    @property
    def stream(self):
        return self.lib.uzu_context_stream(self.h)

    def capabilities(self):
        return self.lib.uzu_context_device_capabilities(self.h)

    def buffer(self, nbytes: int, kind=BUFFER_MANAGED) -> "Buffer":
        b = Buffer(self, nbytes, kind)
        self._buffers.append(b)
        return b

    def upload(self, arr: np.ndarray, kind=BUFFER_MANAGED) -> "Buffer":
        """New CPU-addressable buffer initialised from `arr` (written through DenseBuffer::cpu_ptr)."""
        arr = np.ascontiguousarray(arr)
        b = self.buffer(max(arr.nbytes, 16), kind)
        if arr.nbytes:
            C.memmove(b.cpu_ptr, arr.ctypes.data, arr.nbytes)
        if kind == BUFFER_MANAGED:
            _check(self.lib.uzu_buffer_make_resident(self.h, b.h))
        return b

    def command_buffer(self, name="cmd") -> "CommandBuffer":
        return CommandBuffer(self, name)


class Buffer:
    def __init__(self, ctx: Context, nbytes: int, kind):
        self.ctx = ctx
        h = C.c_void_p()
        _check(ctx.lib.uzu_buffer_create(ctx.h, nbytes, kind, C.byref(h)))
        self.h = h
        self.nbytes = nbytes
        self.ptr = ctx.lib.uzu_buffer_gpu_ptr(h)
        self.cpu_ptr = ctx.lib.uzu_buffer_cpu_ptr(h)

    def numpy(self, dtype, shape=None) -> np.ndarray:
        """Copy of the contents, read through cpu_ptr (synchronise first)."""
        dtype = np.dtype(dtype)
        count = self.nbytes // dtype.itemsize if shape is None else int(np.prod(shape))
        out = np.empty(count, dtype=dtype)
        C.memmove(out.ctypes.data, self.cpu_ptr, count * dtype.itemsize)
        return out if shape is None else out.reshape(shape)

    def close(self):
        if self.h:
            self.ctx.lib.uzu_buffer_destroy(self.h)
            self.h = None


class CommandBuffer:
    """Initial -> Encoding -> Executable -> Pending -> Completed (command_buffer.rs typestate)."""

    def __init__(self, ctx: Context, name: str):
        self.ctx, self.lib = ctx, ctx.lib
        h = C.c_void_p()
        _check(self.lib.uzu_command_buffer_create(ctx.h, name.encode(), C.byref(h)))
        self.h = h

    def __enter__(self):
        _check(self.lib.uzu_command_buffer_start_encoding(self.h))
        return self

    def __exit__(self, exc_type, exc, tb):
        try:
            if exc_type is None:
                self.commit_and_wait()
        finally:
            self.lib.uzu_command_buffer_destroy(self.h)
            self.h = None

    def commit_and_wait(self):
        _check(self.lib.uzu_command_buffer_end_encoding(self.h))
        _check(self.lib.uzu_command_buffer_submit(self.h))
        _check(self.lib.uzu_command_buffer_wait_until_completed(self.h))
        t = C.c_double()
        _check(self.lib.uzu_command_buffer_gpu_execution_time(self.h, C.byref(t)))
        self.gpu_seconds = t.value
        self.launches = self.lib.uzu_command_buffer_launch_count(self.h)

    def encode(self, fn_name: str, *args):
        getattr(self.lib, fn_name)(self.h, *args)


class Engine:
    """Engine + LanguageModel + LanguageModelState + stream (engine/language_model/*)."""

    def __init__(self, ctx: Context, model_dir, max_context_length=8192, use_cuda_graph=True, fused_decode=True, tp_rank=0, tp_size=1):
        self.ctx, self.lib = ctx, ctx.lib
        opts = EngineOptions(max_context_length=max_context_length, use_cuda_graph=int(use_cuda_graph),
                             fused_decode=int(fused_decode), tp_rank=tp_rank, tp_size=tp_size)
        h = C.c_void_p()
        _check(self.lib.uzu_engine_create(ctx.h, str(model_dir).encode(), C.byref(opts), C.byref(h)))
        self.h = h
        info = ModelInfo()
        _check(self.lib.uzu_engine_info(h, C.byref(info)))
        self.info = info

    def close(self):
        if self.h:
            self.lib.uzu_engine_destroy(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def reset(self):
        _check(self.lib.uzu_engine_reset(self.h))

    @property
    def context_length(self):
        return self.lib.uzu_engine_context_length(self.h)

    def snapshot(self):
        _check(self.lib.uzu_engine_snapshot(self.h))

    def restore(self):
        _check(self.lib.uzu_engine_restore(self.h))

    @staticmethod
    def sampling(seed=None, temperature=None, top_k=None, top_p=None, min_p=None) -> SamplingMethod:
        if seed is None:
            return SamplingMethod(kind=SAMPLING_GREEDY)
        return SamplingMethod(kind=SAMPLING_STOCHASTIC, has_temperature=int(temperature is not None),
                              temperature=temperature or 0.0, has_top_k=int(top_k is not None), top_k=top_k or 0,
                              has_top_p=int(top_p is not None), top_p=top_p or 0.0, has_min_p=int(min_p is not None),
                              min_p=min_p or 0.0, seed=seed)

    def prefill(self, tokens, sampling: SamplingMethod | None = None) -> int:
        tokens = np.ascontiguousarray(tokens, dtype=np.uint32)
        out = u32()
        sm = sampling or SamplingMethod(kind=SAMPLING_GREEDY)
        _check(self.lib.uzu_engine_prefill(self.h, tokens.ctypes.data_as(C.POINTER(u32)), len(tokens), C.byref(sm), C.byref(out)))
        return out.value

    def next(self) -> int:
        out = u32()
        _check(self.lib.uzu_engine_next(self.h, C.byref(out)))
        return out.value

    def flush(self) -> int:
        out = u32()
        _check(self.lib.uzu_engine_flush(self.h, C.byref(out)))
        return out.value

    def generate(self, prompt, steps, sampling=None):
        """prefill + `steps` tokens through next()/flush() (the host-visible streaming API)."""
        toks = [self.prefill(prompt, sampling)]
        for _ in range(steps - 1):
            t = self.next()
            if t != 0xFFFFFFFF:
                toks.append(t)
        while len(toks) < steps:
            toks.append(self.flush())
        return toks

    def decode_device(self, steps: int, out_tokens_dev: int = 0):
        _check(self.lib.uzu_engine_decode_device(self.h, steps, out_tokens_dev))

    def forward(self, tokens, row_begin=None, row_end=None) -> np.ndarray:
        tokens = np.ascontiguousarray(tokens, dtype=np.uint32)
        n = len(tokens)
        row_begin = n - 1 if row_begin is None else row_begin
        row_end = n if row_end is None else row_end
        out = np.zeros((row_end - row_begin, self.info.vocab_size), dtype=np.uint16)
        _check(self.lib.uzu_engine_forward(self.h, tokens.ctypes.data_as(C.POINTER(u32)), n, row_begin, row_end,
                                           out.ctypes.data_as(C.POINTER(C.c_uint16))))
        return out

    @property
    def launch_count(self):
        return self.lib.uzu_engine_launch_count(self.h)

    def decode_timed(self, steps: int) -> float:
        t = C.c_double()
        _check(self.lib.uzu_engine_decode_timed(self.h, steps, C.byref(t)))
        return t.value

    def step_host(self, token: int) -> int:
        out = u32()
        _check(self.lib.uzu_engine_step_host(self.h, int(token), C.byref(out)))
        return out.value

    # ---- persistent decode kernel (extension, see include/uzu_b200.h) ----
    @property
    def persistent_decode(self) -> bool:
        return bool(self.lib.uzu_engine_decode_mode(self.h))

    @property
    def persistent_decode_reason(self) -> str:
        return (self.lib.uzu_engine_decode_mode_reason(self.h) or b"").decode(errors="replace")

    def set_persistent_decode(self, on: bool):
        _check(self.lib.uzu_engine_set_decode_mode(self.h, int(on)))

    def decode_trace(self, cta: int = 0):
        """One persistent decode step with per-phase SM-clock stamps of CTA `cta`: (kinds [n], cycles [n, 8])."""
        cap = 4096
        kinds = np.zeros(cap, dtype=np.uint32)
        cyc = np.zeros((cap, 8), dtype=np.uint64)
        n = u32()
        _check(self.lib.uzu_engine_debug_decode_trace(self.h, cta, cap, kinds.ctypes.data_as(C.POINTER(u32)), cyc.ctypes.data_as(C.POINTER(u64)), C.byref(n)))
        return kinds[:n.value], cyc[:n.value]

    def last_logits(self) -> np.ndarray:
        out = np.zeros((1, self.info.vocab_size), dtype=np.uint16)
        _check(self.lib.uzu_engine_last_logits(self.h, out.ctypes.data_as(C.POINTER(C.c_uint16))))
        return out

    # ---- speculative (trie) decode: the verify half of stream.rs:550-657 (host trie logic in uzu_b200/trie.py) ----
    @property
    def speculation_supported(self) -> bool:
        return bool(self.lib.uzu_engine_speculation_supported(self.h))

    def trie_pass(self, tokens, nodes, seeds=None, sampling: SamplingMethod | None = None, want_logits=False):
        """One Decoder::encode over a linearized trie (<= 16 nodes) from the current state. Returns the token sampled at every node
        (and the bf16 logits [count, vocab] when want_logits). Nothing is accepted until trie_accept()."""
        tokens = np.ascontiguousarray(tokens, dtype=np.uint32)
        nodes = np.ascontiguousarray(nodes, dtype=np.uint32).reshape(len(tokens), 3)
        n = len(tokens)
        out = np.zeros(n, dtype=np.uint32)
        logits = np.zeros((n, self.info.vocab_size), dtype=np.uint16) if want_logits else None
        sd = np.ascontiguousarray(seeds, dtype=np.uint64) if seeds is not None else None
        _check(self.lib.uzu_engine_trie_pass(self.h, tokens.ctypes.data_as(C.POINTER(u32)), nodes.ctypes.data_as(C.c_void_p),
                                             sd.ctypes.data_as(C.POINTER(u64)) if sd is not None else None, n,
                                             C.byref(sampling) if sampling is not None else None, out.ctypes.data_as(C.POINTER(u32)),
                                             logits.ctypes.data_as(C.POINTER(C.c_uint16)) if want_logits else None))
        toks = [int(t) for t in out]
        return (toks, logits) if want_logits else toks

    def trie_accept(self, accepted_indices, next_token: int):
        idx = np.ascontiguousarray(accepted_indices, dtype=np.uint32)
        _check(self.lib.uzu_engine_trie_accept(self.h, idx.ctypes.data_as(C.POINTER(u32)), len(idx), int(next_token)))

    def generate_speculative(self, prompt, steps, proposer, sampling: SamplingMethod | None = None, stats: dict | None = None):
        """prefill + `steps` tokens through speculation passes, the way LanguageModelStream::generate drives a speculator
        (stream.rs:550-657, 436-520): `proposer(history, root_token, budget)` returns a uzu_b200.trie.TrieNode tree rooted at `root_token`
        (a draft model in the reference; any callable here); every pass verifies the whole tree with one sweep over the weights and keeps
        the path the model itself would have sampled, so the output equals plain decode whatever the proposer suggests."""
        from .trie import PRng, TrieNode
        seed = sampling.seed if sampling is not None and sampling.kind == SAMPLING_STOCHASTIC else 0
        prng = PRng(seed)
        history = [int(t) for t in prompt]
        root_token = self.prefill(prompt, sampling)
        out = [root_token]
        passes = 0
        while len(out) < steps:
            ctx = self.context_length
            trie = proposer(history + out[:-1], root_token, 16) if proposer is not None else None
            if trie is None:
                trie = TrieNode(root_token, prng.derive(ctx))
            assert trie.token == root_token, "the proposal must be rooted at the last sampled token"
            trie.prune_to_budget(16)
            flat = trie.linearize()
            seeds = [prng.derive(ctx + h) for h in flat.heights()]          # dflash_tfm.rs:267,304
            sampled = self.trie_pass(flat.token_ids(), flat.nodes(), seeds, sampling)
            full = flat.accept(sampled)
            root_token = full[-1][2]
            self.trie_accept([i for i, _, _ in full], root_token)
            out.extend(t for _, _, t in full)
            passes += 1
        if stats is not None:
            stats.update(passes=passes, tokens=len(out) - 1, tokens_per_pass=(len(out) - 1) / max(passes, 1))
        return out[:steps]

    # ---- multi-sequence batched decode (extension, see include/uzu_b200.h) ----
    def batch_begin(self, sequences: int):
        _check(self.lib.uzu_engine_batch_begin(self.h, sequences))
        self._batch = sequences

    def batch_prefill(self, sequence: int, tokens) -> int:
        arr = np.ascontiguousarray(tokens, dtype=np.uint32)
        out = u32()
        _check(self.lib.uzu_engine_batch_prefill(self.h, sequence, arr.ctypes.data_as(C.POINTER(u32)), len(arr), C.byref(out)))
        return out.value

    def batch_step(self, tokens_in) -> list:
        arr = np.ascontiguousarray(tokens_in, dtype=np.uint32)
        assert len(arr) == self._batch
        out = np.zeros(self._batch, np.uint32)
        _check(self.lib.uzu_engine_batch_step(self.h, arr.ctypes.data_as(C.POINTER(u32)), out.ctypes.data_as(C.POINTER(u32))))
        return [int(t) for t in out]

    def batch_decode_timed(self, first_tokens, steps: int) -> float:
        arr = np.ascontiguousarray(first_tokens, dtype=np.uint32)
        t = C.c_double()
        _check(self.lib.uzu_engine_batch_decode_timed(self.h, arr.ctypes.data_as(C.POINTER(u32)), steps, C.byref(t)))
        return t.value

    def batch_logits(self) -> np.ndarray:
        out = np.zeros((self._batch, self.info.vocab_size), np.uint16)
        _check(self.lib.uzu_engine_batch_logits(self.h, out.ctypes.data_as(C.POINTER(C.c_uint16))))
        return out

    def time_prefill_linears(self, m: int, iters: int = 3):
        """(seconds per pass, useful flops per pass) of every linear of one prefill pass over m rows (tensor-core GEMM for m >= 64)."""
        t, f = C.c_double(), C.c_double()
        _check(self.lib.uzu_engine_time_prefill_linears(self.h, m, iters, C.byref(t), C.byref(f)))
        return t.value, f.value

    def time_linears(self, iters: int, select: int = 31):
        t, n = C.c_double(), u64()
        _check(self.lib.uzu_engine_time_linears_select(self.h, iters, select, C.byref(t), C.byref(n)))
        return t.value, n.value
