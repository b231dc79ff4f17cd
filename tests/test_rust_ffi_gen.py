"""The generated Rust FFI layer (integration/rust/backends/cuda/ffi_generated.rs, tools/gen_rust_ffi.py) cannot be compiled here (no rustc),
so it is checked structurally: it is up to date with include/uzu_b200.h, it declares every exported entry point, and the size of every
`#[repr(C)]` struct computed from its Rust field types equals what the library itself reports (uzu_abi_struct_size)."""
import importlib.util
import re
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
spec = importlib.util.spec_from_file_location("gen_rust_ffi", ROOT / "tools" / "gen_rust_ffi.py")
gen = importlib.util.module_from_spec(spec)
spec.loader.exec_module(gen)


def test_generated_file_is_current():
    assert gen.OUT.exists(), "run python tools/gen_rust_ffi.py"
    assert gen.OUT.read_text() == gen.generate(), "include/uzu_b200.h changed: run python tools/gen_rust_ffi.py"


def test_every_export_is_declared():
    from uzu_b200 import binding
    text = gen.OUT.read_text()
    declared = set(re.findall(r"pub fn (uzu_\w+)\(", text))
    assert declared == set(binding.EXPORTS), declared ^ set(binding.EXPORTS)


def test_struct_layouts_match_the_library(lib):
    opaque, enum_types, enums, structs, funcs = gen.parse(gen.HEADER.read_text())
    checked = 0
    for name, fields in structs.items():
        n = lib.uzu_abi_struct_size(name.encode())
        if n == 0:
            continue                      # the library does not register this one
        size, _ = gen.layout(fields, structs, opaque, enum_types)
        assert size == n, (name, size, n)
        checked += 1
    assert checked >= 20, checked
