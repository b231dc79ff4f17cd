"""The generated kernel trait impls (integration/rust/backends/cuda/kernel/generated.rs, tools/gen_rust_kernels.py) cannot be
compiled here (no rustc), so they are checked structurally: current with the kernel table + header, one impl per reference `#[kernel]`,
every covered kernel's struct literal names exactly the fields of its C struct in header order, the C ABI registers every struct the Rust
start-up check asks about, and -- when the reference checkout is present (build container only) -- the committed table is what
tools/extract_kernel_table.py extracts from it."""
import importlib.util
import json
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


def _load(name):
    spec = importlib.util.spec_from_file_location(name, ROOT / "tools" / f"{name}.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


gen = _load("gen_rust_kernels")
ffi = _load("gen_rust_ffi")
TABLE = json.loads((ROOT / "integration" / "rust" / "kernel_table.json").read_text())["kernels"]


def test_generated_file_is_current():
    assert gen.OUT.exists(), "run python tools/gen_rust_kernels.py"
    assert gen.OUT.read_text() == gen.generate(), "kernel table / header / generator changed: run python tools/gen_rust_kernels.py"


def test_one_impl_per_reference_kernel_and_the_macro_names_them_all():
    text = gen.OUT.read_text()
    names = sorted(k["name"] for k in TABLE)
    assert len(names) == 72 and len(set(names)) == 72
    impls = re.findall(r"^impl (\w+)Kernel for Cuda(\w+)Kernel \{", text, flags=re.M)
    assert sorted(a for a, _ in impls) == names and all(a == b for a, b in impls)
    macro = re.findall(r"type (\w+)Kernel = crate::backends::cuda::kernel::generated::Cuda(\w+)Kernel;", text)
    assert sorted(a for a, _ in macro) == names and all(a == b for a, b in macro)
    stubs = text.count("outside the CUDA decode hot path")
    assert stubs == 72 - len(gen.COVERED) and len(gen.COVERED) >= 20


def test_covered_kernels_fill_their_c_struct_field_for_field():
    text = gen.OUT.read_text()
    opaque, enum_types, enums, structs, funcs = ffi.parse(ffi.HEADER.read_text())
    funcs = {n: (r, a) for n, r, a in funcs}
    for name, cfg in gen.COVERED.items():
        style, target, slots = gen.c_target(cfg, structs, funcs)
        block = text[text.index(f"impl {name}Kernel for Cuda{name}Kernel"):]
        block = block[:block.index("\n}\n")]
        if style == "struct":
            lit = block[block.index(f"ffi::{target} {{"):]
            lit = lit[:lit.index("};")]
            got = re.findall(r"^\s{12}(\w+):", lit, flags=re.M)
            assert got == [f for f, _ in slots], (name, got)
            assert f"ffi::{cfg['c']}_encode(encoder.as_command_buffer_mut().raw(), &args)" in block
        else:
            call = block[block.index(f"ffi::{target}("):]
            assert call.count(",") >= len(slots), name
        # signature shape of build/common/traitgen.rs: one lifetime per buffer argument + 'encoder
        k = next(k for k in TABLE if k["name"] == name)
        buffers = [a["name"] for a in k["arguments"] if a["kind"].startswith("buffer")]
        sig = re.search(r"fn encode<([^>]*)>", block).group(1).split(", ")
        assert sig == [f"'{b}" for b in buffers] + ["'encoder"], (name, sig)


def test_library_registers_every_struct_the_rust_start_up_check_asks_about(lib):
    opaque, enum_types, enums, structs, funcs = ffi.parse(ffi.HEADER.read_text())
    for name, fields in structs.items():
        n = lib.uzu_abi_struct_size(name.encode())
        assert n == ffi.layout(fields, structs, opaque, enum_types)[0], name


@pytest.mark.skipif(not Path("/root/reference/crates/backend-uzu").exists(), reason="reference checkout not present (GPU box)")
def test_committed_kernel_table_is_what_the_reference_declares():
    ext = _load("extract_kernel_table")
    assert ext.extract(Path("/root/reference")) == TABLE, "run python tools/extract_kernel_table.py"


def _strip_rust(text: str) -> str:
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    text = re.sub(r"//[^\n]*", " ", text)
    text = re.sub(r'"(\\.|[^"\\])*"', '""', text)
    return re.sub(r"'(\\.|[^\\'])'", "' '", text)          # char literals; lifetimes ('a) have no closing quote and stay


def test_every_rust_file_has_balanced_delimiters():
    """No rustc in the image: at least the cheapest class of slip (an unbalanced brace / bracket / parenthesis in hand-written or generated
    glue) is caught here."""
    pairs = {")": "(", "]": "[", "}": "{"}
    files = sorted((ROOT / "integration" / "rust").rglob("*.rs"))
    assert len(files) >= 9
    for f in files:
        stack = []
        for i, ch in enumerate(_strip_rust(f.read_text())):
            if ch in "([{":
                stack.append(ch)
            elif ch in pairs:
                assert stack and stack[-1] == pairs[ch], f"{f.name}: unbalanced {ch!r} at offset {i}"
                stack.pop()
        assert not stack, f"{f.name}: unclosed {stack[-1]!r}"
