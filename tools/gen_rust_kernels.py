#!/usr/bin/env python
"""Generate integration/rust/backends/cuda/kernel/generated.rs: one `impl <Name>Kernel for Cuda<Name>Kernel` per `#[kernel]` of the
reference -- what a `build/cuda/compiler.rs` arm would emit next to build/cpu/compiler.rs:176-613.

Inputs: integration/rust/kernel_table.json (the declaration table, tools/extract_kernel_table.py) and include/uzu_b200.h (the C ABI).
For the kernels this backend implements, `new` stores the generics / `#[specialize]` block and `encode` fills `uzu_<name>_args` (or the
positional C parameters) FIELD BY NAME: a reference argument with no C counterpart, or a C field nothing feeds, stops the generator
unless COVERED below says what to do with it. That makes the header/reference agreement mechanical instead of reviewed by eye.
Kernels outside the decode hot path get a stub whose `new` returns CudaError::NotSupported (the CPU backend does the same for sparse
buffers), so `impl Kernels for CudaKernels` names all 72 associated types.

Signatures follow build/common/traitgen.rs:15-81: new(context, <generics: type -> DataType, const -> its type>, <specialize args>);
encode<'buffer.., 'encoder>(&self, <arguments>, encoder: &'encoder mut Encoder<Self::Backend>), `#[optional]` -> Option<_>, `*const` ->
impl BufferArg, `*mut` -> impl BufferArgMut, `&[T]` -> constant slice. No Rust toolchain exists in this image: the output has not been
through rustc; tests/test_rust_kernels_gen.py checks its structure against the table and the header.

    python tools/gen_rust_kernels.py [--check]"""
from __future__ import annotations

import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent))
import gen_rust_ffi as FFI  # noqa: E402

ROOT = Path(__file__).resolve().parents[1]
TABLE = ROOT / "integration" / "rust" / "kernel_table.json"
OUT = ROOT / "integration" / "rust" / "backends" / "cuda" / "kernel" / "generated.rs"

# kernel -> C entry point stem + what to do with the pieces that do not map one-to-one by name.
#   bf16:      type generics that must be DataType::BF16 (others are accepted as declared: f32 affine / rope / accumulation types)
#   dtype:     {C field: generic} -- C field receives dt(generic) (kernels with a bf16 | f32 switch)
#   reject:    {specialize arg: (rust condition over the arg, message)} -> NotSupported from `new`
#   drop:      reference arguments the C ABI has no slot for (only legal when `reject` makes them dead)
#   rename:    {reference argument: C name}
#   zero:      C fields left 0 (CUDA-side extensions the reference path never sets)
#   custom:    {C field: rust expression}
COVERED = {
    "Normalization": dict(c="uzu_normalization", bf16=["InputT", "OutputT"], reject={"use_hadamard": ("use_hadamard", "in-norm Hadamard (Mirai RHT)")}),
    "QKVNorm": dict(c="uzu_qkv_norm", bf16=["InputT", "OutputT"]),
    "AttentionPrepare": dict(c="uzu_attention_prepare", bf16=["ElementT"], zero=["dynamic_position"]),
    "AttentionSinglePass": dict(c="uzu_attention_single_pass", bf16=["T"], zero=["sums", "maxs", "dynamic_position"],
                                reject={"HEAD_DIM": ("HEAD_DIM == 512", "head_dim 512 (AttentionFallback path)")},
                                custom={"ring_params": "ring_params.map(|r| ffi::uzu_ring_params { ring_offset: r.ring_offset, ring_length: r.ring_length }).unwrap_or(ffi::uzu_ring_params { ring_offset: 0, ring_length: 0 })"}),
    "AttentionTwoPass1": dict(c="uzu_attention_two_pass1", bf16=["T"], zero=["dynamic_position"],
                              reject={"HEAD_DIM": ("HEAD_DIM == 512", "head_dim 512 (AttentionFallback path)")},
                              rename={"partials": "out"},
                              custom={"ring_params": "ring_params.map(|r| ffi::uzu_ring_params { ring_offset: r.ring_offset, ring_length: r.ring_length }).unwrap_or(ffi::uzu_ring_params { ring_offset: 0, ring_length: 0 })"}),
    "AttentionTwoPass2": dict(c="uzu_attention_two_pass2", bf16=["T"]),
    "KVCacheUpdate": dict(c="uzu_kv_cache_update", bf16=["T"], custom={"copies": "copies.as_ptr() as *const ffi::uzu_kv_copy"}),
    "ActivationTransform": dict(c="uzu_activation_transform", dtype={"data_type": "T"}),
    "SigmoidGate": dict(c="uzu_sigmoid_gate", bf16=["T"]),
    "GatedActMul": dict(c="uzu_gated_act_mul", bf16=["T"],
                        reject={"use_hadamard": ("use_hadamard", "GatedActMul with a fused Hadamard"),
                                "ops": ("ops != crate::backends::common::gpu_types::GatedActMulOp::FullPrecision", "GatedActMul with int8 activation outputs (A8)")},
                        drop=["q_out", "scales_out", "group_sums_out", "hadamard_factors", "ops", "use_hadamard", "activation_scale_group_size", "sum_group_size"]),
    "QuantizedEmbeddingLookup": dict(c="uzu_quantized_embedding_lookup", bf16=["T"],
                                     reject={"use_hadamard": ("use_hadamard", "quantized embedding with an output Hadamard")},
                                     drop=["output_hadamard_factors", "use_hadamard"]),
    "FullPrecisionEmbeddingLookup": dict(c="uzu_full_precision_embedding_lookup", bf16=["T"]),
    "LogitTransform": dict(c="uzu_logit_transform", bf16=["T"]),
    "TensorAddScale": dict(c="uzu_tensor_add_scale", bf16=["T"], drop=["in_place"]),
    "TensorCopy": dict(c="uzu_tensor_copy", bf16=["T"], rename={"src_buffer": "src", "dst_buffer": "dst"}),
    "TensorAddBias": dict(c="uzu_tensor_add_bias", bf16=["T", "BiasT"], drop=["in_place"]),
    "TensorAddSwap": dict(c="uzu_tensor_add_swap", bf16=["T"]),
    "UnifiedSampling": dict(c="uzu_unified_sampling", bf16=["T"]),
    "DeltaNetConvUpdate": dict(c="uzu_delta_net_conv_update", bf16=["T"]),
    "DeltaNetUpdate": dict(c="uzu_delta_net_update", bf16=["T"]),
}

RUST_TO_C_SCALAR = {"u32": "u32", "f32": "f32", "bool": "u32"}


def rust_param_type(arg_or_generic) -> str:
    if arg_or_generic.get("kind") == "type":
        return "DataType"
    return arg_or_generic["type"]


def encode_arg_type(a) -> tuple[str | None, str]:
    """(lifetime or None, rust type) of an encode argument, as build/common/traitgen.rs:33-62 spells it"""
    k = a["kind"]
    if k == "buffer_read":
        lt, ty = f"'{a['name']}", f"impl BufferArg<'{a['name']}, Cuda>"
    elif k == "buffer_read_write":
        lt, ty = f"'{a['name']}", f"impl BufferArgMut<'{a['name']}, Cuda>"
    elif k == "constant_slice":
        lt, ty = None, f"&[{a['type']}]"
    elif k == "constant_array":
        lt, ty = None, f"&[{a['type']}]"
    else:
        lt, ty = None, a["type"]
    if a["optional"] is not None:
        ty = f"Option<{ty}>"
    return lt, ty


def c_target(cfg, structs, funcs):
    """('struct', struct name, [(field, ctype)]) or ('positional', fn, [(param, ctype)])"""
    fn = cfg["c"] + "_encode"
    _, params = funcs[fn]
    if len(params) == 2 and params[1][1].rstrip().endswith("*"):
        sname = params[1][1].replace("const", "").replace("*", "").strip()
        return "struct", sname, [(f, ct) for f, ct, _ in structs[sname]]
    return "positional", fn, params[1:]


def value_expr(name, a, ctype) -> str:
    """rust expression for reference encode argument `a` feeding a C slot of type `ctype`"""
    k = a["kind"]
    opt = a["optional"] is not None
    if k == "buffer_read":
        return f"{name}.map(addr).unwrap_or(0)" if opt else f"addr({name})"
    if k == "buffer_read_write":
        return f"{name}.map(addr_mut).unwrap_or(0)" if opt else f"addr_mut({name})"
    if k == "scalar":
        if a["type"] in ("u32", "f32"):
            return f"{name}.unwrap_or_default()" if opt else name
        if a["type"] == "bool":
            return f"{name} as u32"
        return f"{name} as u32"            # #[repr(C)] gpu_types enums
    raise ValueError(f"no automatic mapping for {name} ({k})")


def stored_expr(name, ty) -> str:
    return f"self.{name}" if ty in ("u32", "f32") else (f"dt(self.{name})" if ty == "DataType" else f"self.{name} as u32")


def gen_kernel(k, cfg, structs, funcs) -> list[str]:
    name = k["name"]
    sname = f"Cuda{name}Kernel"
    generics = k["generics"]
    spec = [a for a in k["arguments"] if a["kind"] == "specialize"]
    enc = [a for a in k["arguments"] if a["kind"] != "specialize"]
    new_params = [(g["name"], rust_param_type(g)) for g in generics] + [(a["name"], a["type"]) for a in spec]
    out = [f"/// {k['file'].split('src/')[-1]}:{k['line']}"]
    lifetimes = [lt for lt, _ in map(encode_arg_type, enc) if lt] + ["'encoder"]
    enc_sig = ", ".join([f"{a['name']}: {encode_arg_type(a)[1]}" for a in enc] + ["encoder: &'encoder mut Encoder<Cuda>"])
    new_sig = ", ".join(["context: &CudaContext"] + [f"{n}: {t}" for n, t in new_params])
    if cfg is None:
        out += [f"pub struct {sname};",
                f"impl {name}Kernel for {sname} {{", "    type Backend = Cuda;",
                "    #[allow(non_snake_case, unused_variables, clippy::too_many_arguments)]",
                f"    fn new({new_sig}) -> Result<Self, CudaError> {{",
                f"        Err(CudaError::NotSupported(\"{name}: outside the CUDA decode hot path (SURVEY 8 scope)\"))", "    }",
                "    #[allow(unused_variables, clippy::too_many_arguments)]",
                f"    fn encode<{', '.join(lifetimes)}>(&self, {enc_sig}) {{",
                f"        unreachable!(\"{sname} cannot be constructed\")", "    }", "}", ""]
        return out
    style, target, slots = c_target(cfg, structs, funcs)
    rename = cfg.get("rename", {})
    drop = set(cfg.get("drop", []))
    custom = cfg.get("custom", {})
    zero = set(cfg.get("zero", []))
    dtype = cfg.get("dtype", {})
    stored = {n.lower() if n.isupper() else n: (n, t) for n, t in new_params}      # HEAD_DIM feeds `head_dim`
    by_c = {rename.get(a["name"], a["name"]): a for a in enc}
    # every reference item must land somewhere
    cnames = [c for c, _ in slots]
    for a in enc:
        cn = rename.get(a["name"], a["name"])
        if cn not in cnames and a["name"] not in drop:
            raise SystemExit(f"{name}: reference argument {a['name']!r} has no slot in {target}")
    for n, t in new_params:
        key = n.lower() if n.isupper() else n
        if t != "DataType" and key not in cnames and n not in drop and n not in cfg.get("reject", {}):
            raise SystemExit(f"{name}: specialization {n!r} has no slot in {target}")
    fields = []
    for cn, ct in slots:
        if cn in custom:
            expr = custom[cn]
        elif cn in dtype:
            expr = f"dt(self.{dtype[cn]})"
        elif cn in by_c:
            expr = value_expr(by_c[cn]["name"], by_c[cn], ct)
        elif cn in stored:
            expr = stored_expr(*stored[cn])
        elif cn in zero:
            expr = "0"
        else:
            raise SystemExit(f"{name}: nothing feeds C slot {cn!r} of {target}")
        fields.append((cn, expr))
    keep = [(n, t) for n, t in new_params]
    out += ["#[allow(non_snake_case, dead_code)]",
            "pub struct " + sname + " { " + ", ".join(f"{n}: {t}" for n, t in keep) + " }" if keep else f"pub struct {sname};",
            f"impl {name}Kernel for {sname} {{", "    type Backend = Cuda;",
            "    #[allow(non_snake_case, unused_variables, clippy::too_many_arguments)]",
            f"    fn new({new_sig}) -> Result<Self, CudaError> {{"]
    for g in cfg.get("bf16", []):
        out.append(f"        if {g} != DataType::BF16 {{ return Err(CudaError::NotSupported(\"{name}: {g} must be bf16\")); }}")
    for g in dtype.values():
        out.append(f"        if {g} != DataType::BF16 && {g} != DataType::F32 {{ return Err(CudaError::NotSupported(\"{name}: {g} must be bf16 or f32\")); }}")
    for arg, (cond, msg) in cfg.get("reject", {}).items():
        out.append(f"        if {cond} {{ return Err(CudaError::NotSupported(\"{msg}\")); }}")
    out.append("        Ok(Self" + (" { " + ", ".join(n for n, _ in keep) + " }" if keep else "") + ")")
    out.append("    }")
    out.append("    #[allow(unused_variables, clippy::too_many_arguments)]")
    out.append(f"    fn encode<{', '.join(lifetimes)}>(&self, {enc_sig}) {{")
    if style == "struct":
        out.append(f"        let args = ffi::{target} {{")
        for cn, expr in fields:
            out.append(f"            {cn}: {expr},")
        out.append("        };")
        out.append(f"        unsafe {{ ffi::{cfg['c']}_encode(encoder.as_command_buffer_mut().raw(), &args) }};")
    else:
        out.append(f"        unsafe {{ ffi::{target}(encoder.as_command_buffer_mut().raw(), " + ", ".join(e for _, e in fields) + ") };")
    out += ["    }", "}", ""]
    return out


def generate() -> str:
    table = json.loads(TABLE.read_text())["kernels"]
    opaque, enum_types, enums, structs, funcs = FFI.parse(FFI.HEADER.read_text())
    funcs = {n: (r, a) for n, r, a in funcs}
    names = sorted(k["name"] for k in table)
    unknown = set(COVERED) - set(names)
    if unknown:
        raise SystemExit(f"COVERED names missing from the kernel table: {sorted(unknown)}")
    out = ["//! GENERATED by tools/gen_rust_kernels.py from integration/rust/kernel_table.json + include/uzu_b200.h -- do not edit.",
           f"//! {len(table)} kernel traits of backends/common/kernel (traitgen.rs:15-81): {len(COVERED)} implemented over libuzu_b200.so, the rest",
           "//! NotSupported stubs. Never compiled here (no rustc in the build image); structure checked by tests/test_rust_kernels_gen.py.",
           "#![allow(clippy::style, clippy::complexity, clippy::perf)]",
           "use crate::{", "    backends::common::{kernel::*, BufferArg, BufferArgMut, Encoder},", "    data_type::DataType,", "};", "",
           "use super::{addr, addr_mut, dt};", "use crate::backends::cuda::{ffi_generated as ffi, Cuda, CudaContext, CudaError};", ""]
    for k in sorted(table, key=lambda k: k["name"]):
        out += gen_kernel(k, COVERED.get(k["name"]), structs, funcs)
    out.append("/// `autogen_kernels!()` of this backend (build/cpu/compiler.rs:632-658 emits the CPU one): every generated associated type of `Kernels`.")
    out.append("macro_rules! autogen_cuda_kernels {")
    out.append("    () => {")
    for n in names:
        out.append(f"        type {n}Kernel = crate::backends::cuda::kernel::generated::Cuda{n}Kernel;")
    out += ["    };", "}", "pub(crate) use autogen_cuda_kernels;", ""]
    return "\n".join(out)


if __name__ == "__main__":
    text = generate()
    if "--check" in sys.argv:
        if not OUT.exists() or OUT.read_text() != text:
            print(f"{OUT} is stale: run python tools/gen_rust_kernels.py")
            sys.exit(1)
        print("up to date")
    else:
        OUT.write_text(text)
        print(f"wrote {OUT} ({len(text.splitlines())} lines)")
