#!/usr/bin/env python
"""Generate integration/rust/backends/cuda/ffi_generated.rs from include/uzu_b200.h: every opaque handle, every enum constant, every
`#[repr(C)]` argument struct and every `extern "C"` entry point, mechanically (the hand-written ffi.rs elided most of them).

    python tools/gen_rust_ffi.py [--check]

`--check` regenerates in memory and fails if the committed file differs (tests/test_rust_ffi_gen.py also checks that the struct layouts
computed from the Rust field types equal the sizes the library reports through uzu_abi_struct_size). No Rust toolchain exists in this
image, so the output has never been through rustc; it is plain `extern "C"` / `#[repr(C)]` with primitive types only."""
from __future__ import annotations

import re
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
HEADER = ROOT / "include" / "uzu_b200.h"
OUT = ROOT / "integration" / "rust" / "backends" / "cuda" / "ffi_generated.rs"

PRIM = {"uint64_t": "u64", "uint32_t": "u32", "int32_t": "i32", "uint16_t": "u16", "uint8_t": "u8", "float": "f32", "double": "f64",
        "int": "c_int", "unsigned int": "u32", "size_t": "usize", "char": "c_char", "void": "c_void", "uzu_status": "c_int"}
SIZES = {"u64": 8, "u32": 4, "i32": 4, "u16": 2, "u8": 1, "f32": 4, "f64": 8, "c_int": 4, "usize": 8}


def strip_comments(text: str) -> str:
    return re.sub(r"/\*.*?\*/", " ", text, flags=re.S)


def parse(text: str):
    text = strip_comments(text)
    opaque = re.findall(r"typedef\s+struct\s+(uzu_\w+)\s+\1\s*;", text)
    enums = {}        # constant -> value expression
    enum_types = set()
    for m in re.finditer(r"(?:typedef\s+)?enum\s*(uzu_\w+)?\s*\{(.*?)\}\s*(uzu_\w+)?\s*;", text, flags=re.S):
        if m.group(3):
            enum_types.add(m.group(3))
        for item in m.group(2).split(","):
            item = item.strip()
            if not item:
                continue
            name, _, val = item.partition("=")
            enums[name.strip()] = val.strip()
    structs = {}      # name -> [(field, ctype, array_len)]
    for m in re.finditer(r"typedef\s+struct\s+(uzu_\w+)\s*\{(.*?)\}\s*\1\s*;", text, flags=re.S):
        fields = []
        for decl in m.group(2).split(";"):
            decl = " ".join(decl.split())
            if not decl:
                continue
            mm = re.match(r"((?:const\s+)?(?:unsigned\s+)?\w+)\s*(.*)", decl)
            ctype, rest = mm.group(1), mm.group(2)
            for var in rest.split(","):
                var = var.strip()
                stars = var.count("*")
                var = var.replace("*", "").strip()
                am = re.match(r"(\w+)\[(\d+)\]", var)
                if am:
                    fields.append((am.group(1), ctype + "*" * stars, int(am.group(2))))
                else:
                    fields.append((var, ctype + "*" * stars, 0))
        structs[m.group(1)] = fields
    funcs = []        # (name, ret, [(argname, ctype_with_ptrs)])
    for m in re.finditer(r"UZU_API\s+([\w\s\*]+?)\b(uzu_\w+)\s*\((.*?)\)\s*;", text, flags=re.S):
        ret, name, args = " ".join(m.group(1).split()), m.group(2), " ".join(m.group(3).split())
        arglist = []
        if args and args != "void":
            for i, a in enumerate(args.split(",")):
                a = a.strip()
                mm = re.match(r"(.*?)(\w+)$", a)
                t, n = mm.group(1).strip(), mm.group(2)
                if not t:                      # unnamed parameter
                    t, n = a, f"arg{i}"
                arglist.append((n, t))
        funcs.append((name, ret, arglist))
    return opaque, enum_types, enums, structs, funcs


def rust_type(ctype: str, structs, opaque, enum_types) -> str:
    t = ctype.replace("const ", "const@").strip()
    stars = t.count("*")
    base = t.replace("*", "").replace("const@", "").strip()
    is_const = "const@" in t
    if base in PRIM:
        r = PRIM[base]
    elif base in enum_types:
        r = "c_int"
    elif base in structs or base in opaque:
        r = base
    else:
        raise ValueError(f"unknown C type {ctype!r}")
    for _ in range(stars):
        r = ("*const " if is_const else "*mut ") + r
        is_const = False if stars > 1 else is_const
    return r


def layout(fields, structs, opaque, enum_types):
    """(size, align) of a #[repr(C)] struct with these fields."""
    off, align = 0, 1
    for _, ctype, n in fields:
        rt = rust_type(ctype, structs, opaque, enum_types)
        if rt.startswith("*"):
            sz, al = 8, 8
        elif rt in SIZES:
            sz = al = SIZES[rt]
        else:
            sz, al = layout(structs[rt], structs, opaque, enum_types)
        off = (off + al - 1) // al * al
        off += sz * max(1, n)
        align = max(align, al)
    return (off + align - 1) // align * align, align


def eval_const(expr: str) -> int:
    return int(eval(re.sub(r"(\d+)u\b", r"\1", expr), {"__builtins__": {}}))


def generate() -> str:
    opaque, enum_types, enums, structs, funcs = parse(HEADER.read_text())
    out = ["//! GENERATED by tools/gen_rust_ffi.py from include/uzu_b200.h -- do not edit. Complete `extern \"C\"` surface of libuzu_b200.so:",
           "//! opaque handles, enum constants, `#[repr(C)]` argument structs (field for field) and every entry point. Never compiled here",
           "//! (no rustc in the build image); the struct layouts are checked against the library by tests/test_rust_ffi_gen.py.",
           "#![allow(non_camel_case_types, dead_code)]", "use std::ffi::{c_char, c_int, c_void};", ""]
    for o in opaque:
        out.append(f"#[repr(C)] pub struct {o} {{ _p: [u8; 0] }}")
    out.append("")
    for name, val in enums.items():
        out.append(f"pub const {name}: u32 = {eval_const(val)};")
    out.append("")
    for name, fields in structs.items():
        size, _ = layout(fields, structs, opaque, enum_types)
        out.append(f"/// {size} bytes")
        has_ptr = any(rust_type(ct, structs, opaque, enum_types).startswith("*") for _, ct, _ in fields)
        nested = any(rust_type(ct, structs, opaque, enum_types) in structs and
                     any(rust_type(c2, structs, opaque, enum_types).startswith("*") for _, c2, _ in structs[rust_type(ct, structs, opaque, enum_types)])
                     for _, ct, _ in fields)
        out.append("#[repr(C)] #[derive(Clone, Copy)]" if has_ptr or nested else "#[repr(C)] #[derive(Default, Clone, Copy)]")
        out.append(f"pub struct {name} {{")
        for f, ctype, n in fields:
            rt = rust_type(ctype, structs, opaque, enum_types)
            out.append(f"    pub {f}: " + (f"[{rt}; {n}]," if n else f"{rt},"))
        out.append("}")
        out.append("")
    out.append("/// (C name, size_of) of every argument struct: ffi::abi_self_check() compares them with uzu_abi_struct_size() of the loaded library")
    out.append("pub const STRUCT_SIZES: &[(&str, usize)] = &[")
    for name in structs:
        out.append(f"    (\"{name}\", std::mem::size_of::<{name}>()),")
    out.append("];")
    out.append("")
    out.append('#[link(name = "uzu_b200")]')
    out.append('extern "C" {')
    for name, ret, args in funcs:
        a = ", ".join(f"{('type_' if n == 'type' else n)}: {rust_type(t, structs, opaque, enum_types)}" for n, t in args)
        r = "" if ret == "void" else f" -> {rust_type(ret, structs, opaque, enum_types)}"
        out.append(f"    pub fn {name}({a}){r};")
    out.append("}")
    return "\n".join(out) + "\n"


if __name__ == "__main__":
    text = generate()
    if "--check" in sys.argv:
        sys.exit(0 if OUT.exists() and OUT.read_text() == text else 1)
    OUT.write_text(text)
    print(f"wrote {OUT} ({len(text.splitlines())} lines)")
