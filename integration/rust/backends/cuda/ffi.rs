//! `extern "C"` block over include/uzu_b200.h (hand-written; `bindgen include/uzu_b200.h` produces the same items).
//! Every `#[repr(C)]` struct mirrors the header field for field; `abi_self_check()` compares sizes with the library at start-up.
#![allow(non_camel_case_types)]
use std::ffi::{c_char, c_int, c_void};

#[repr(C)] pub struct uzu_context { _p: [u8; 0] }
#[repr(C)] pub struct uzu_buffer { _p: [u8; 0] }
#[repr(C)] pub struct uzu_sparse_buffer { _p: [u8; 0] }
#[repr(C)] pub struct uzu_command_buffer { _p: [u8; 0] }

pub const UZU_BUFFER_MANAGED: c_int = 0;
pub const UZU_BUFFER_PINNED_HOST: c_int = 1;
pub const UZU_BUFFER_DEVICE: c_int = 2;

#[repr(C)] #[derive(Default, Clone, Copy)]
pub struct uzu_matmul_args {
    pub a: u64, pub b: u64, pub b_scales: u64, pub b_zero_points: u64, pub b_biases: u64, pub d: u64, pub bias: u64, pub gather_indices: u64,
    pub b_prologue: u32, pub b_mode: u32, pub b_group_size: u32, pub b_signed_codes: u32, pub b_leading_dimension: u32, pub b_transpose: u32,
    pub d_transform: u32, pub ab_scale: f32, pub soft_cap: f32, pub m: u32, pub n: u32, pub k: u32,
    pub weights_dt: u32, pub input_dt: u32, pub output_dt: u32,
}

#[repr(C)] #[derive(Default, Clone, Copy)]
pub struct uzu_normalization_args {
    pub input: u64, pub scales: u64, pub biases: u64, pub output: u64, pub shortcut: u64, pub hadamard_factors: u64,
    pub batch_size: u32, pub element_count: u32, pub epsilon: f32, pub scale_offset: f32, pub post_layer_scalar: f32,
    pub in_place: u32, pub subtract_mean: u32, pub full_layer: u32, pub copy_to_shortcut: u32, pub residual_add: u32, pub use_hadamard: u32,
    pub scale_residual_sum: u32, pub scale_output: u32, pub has_biases: u32, pub has_scales: u32,
}

#[repr(C)] #[derive(Default, Clone, Copy)]
pub struct uzu_tp_all_gather_args { pub src: u64, pub dst: u64, pub scratch: u64, pub rows: u32, pub cols_local: u32 }

#[link(name = "uzu_b200")]
extern "C" {
    pub fn uzu_last_error() -> *const c_char;
    pub fn uzu_abi_struct_size(name: *const c_char) -> usize;

    pub fn uzu_context_create(device_ordinal: c_int, out: *mut *mut uzu_context) -> c_int;
    pub fn uzu_context_destroy(ctx: *mut uzu_context);
    pub fn uzu_context_peak_memory_usage(ctx: *mut uzu_context, out_bytes: *mut usize) -> c_int;
    pub fn uzu_context_device_capabilities(ctx: *mut uzu_context) -> u32;
    pub fn uzu_context_start_capture(ctx: *mut uzu_context, trace_path: *const c_char) -> c_int;
    pub fn uzu_context_stop_capture(ctx: *mut uzu_context) -> c_int;

    pub fn uzu_buffer_create(ctx: *mut uzu_context, size: usize, kind: c_int, out: *mut *mut uzu_buffer) -> c_int;
    pub fn uzu_buffer_destroy(buf: *mut uzu_buffer);
    pub fn uzu_buffer_gpu_ptr(buf: *const uzu_buffer) -> u64;
    pub fn uzu_buffer_cpu_ptr(buf: *const uzu_buffer) -> *mut c_void;
    pub fn uzu_buffer_size(buf: *const uzu_buffer) -> usize;
    pub fn uzu_buffer_make_resident(ctx: *mut uzu_context, buf: *mut uzu_buffer) -> c_int;

    pub fn uzu_sparse_buffer_create(ctx: *mut uzu_context, capacity: usize, out: *mut *mut uzu_sparse_buffer) -> c_int;
    pub fn uzu_sparse_buffer_destroy(buf: *mut uzu_sparse_buffer);
    pub fn uzu_sparse_buffer_gpu_ptr(buf: *const uzu_sparse_buffer) -> u64;
    pub fn uzu_sparse_buffer_size(buf: *const uzu_sparse_buffer) -> usize;
    pub fn uzu_sparse_buffer_page_size_bytes(buf: *const uzu_sparse_buffer) -> usize;
    pub fn uzu_sparse_buffer_map(buf: *mut uzu_sparse_buffer, pages: *const u32, page_count: usize) -> c_int;
    pub fn uzu_sparse_buffer_unmap(buf: *mut uzu_sparse_buffer, pages: *const u32, page_count: usize) -> c_int;

    pub fn uzu_command_buffer_create(ctx: *mut uzu_context, name: *const c_char, out: *mut *mut uzu_command_buffer) -> c_int;
    pub fn uzu_command_buffer_destroy(cmd: *mut uzu_command_buffer);
    pub fn uzu_command_buffer_start_encoding(cmd: *mut uzu_command_buffer) -> c_int;
    pub fn uzu_command_buffer_encode_copy(cmd: *mut uzu_command_buffer, src: u64, dst: u64, bytes: usize);
    pub fn uzu_command_buffer_encode_fill(cmd: *mut uzu_command_buffer, dst: u64, bytes: usize, value: u8);
    pub fn uzu_command_buffer_push_debug_group(cmd: *mut uzu_command_buffer, name: *const c_char);
    pub fn uzu_command_buffer_pop_debug_group(cmd: *mut uzu_command_buffer);
    pub fn uzu_command_buffer_end_encoding(cmd: *mut uzu_command_buffer) -> c_int;
    pub fn uzu_command_buffer_submit(cmd: *mut uzu_command_buffer) -> c_int;
    pub fn uzu_command_buffer_wait_until_completed(cmd: *mut uzu_command_buffer) -> c_int;
    pub fn uzu_command_buffer_gpu_execution_time(cmd: *mut uzu_command_buffer, out_seconds: *mut f64) -> c_int;

    pub fn uzu_matmul_validate(args: *const uzu_matmul_args) -> c_int;
    pub fn uzu_matmul_encode(cmd: *mut uzu_command_buffer, args: *const uzu_matmul_args);
    pub fn uzu_normalization_encode(cmd: *mut uzu_command_buffer, args: *const uzu_normalization_args);
    // ... one declaration per UZU_API item of include/uzu_b200.h (qkv_norm, attention_prepare, attention_single_pass / two_pass1 / two_pass2,
    // kv_cache_update, sigmoid_gate, gated_act_mul, quantized / full-precision embedding lookup, logit_transform, tensor_*, unified_sampling,
    // delta_net_conv_update / update): same shape as the two above.

    // tensor parallelism (extension)
    pub fn uzu_tp_get_unique_id(out128: *mut u8) -> c_int;
    pub fn uzu_context_tp_init(ctx: *mut uzu_context, rank: u32, size: u32, unique_id128: *const u8) -> c_int;
    pub fn uzu_tp_all_reduce_encode(cmd: *mut uzu_command_buffer, partial_f32: u64, count: u32, out_bf16: u64);
    pub fn uzu_tp_all_gather_encode(cmd: *mut uzu_command_buffer, args: *const uzu_tp_all_gather_args);
}

/// Call once at start-up: a binding compiled against a different header revision must not run.
pub fn abi_self_check() {
    macro_rules! chk { ($t:ty) => {{
        let name = std::ffi::CString::new(stringify!($t)).unwrap();
        let lib = unsafe { uzu_abi_struct_size(name.as_ptr()) };
        assert_eq!(lib, std::mem::size_of::<$t>(), concat!("ABI mismatch: ", stringify!($t)));
    }}; }
    chk!(uzu_matmul_args);
    chk!(uzu_normalization_args);
    chk!(uzu_tp_all_gather_args);
}
