#!/usr/bin/env python
"""Extract the kernel declaration table the reference's build scripts walk (build/cpu/compiler.rs:176-330: every `#[kernel(Name)]` fn under
src/backends/cpu/kernel/**, its `#[variants]` generics, its arguments with `#[optional(..)]` / `#[specialize]`) into
integration/rust/kernel_table.json, so that tools/gen_rust_kernels.py can emit the `impl <Name>Kernel for Cuda<Name>Kernel` blocks
 on a machine that has neither the reference checkout nor a Rust toolchain.

    python tools/extract_kernel_table.py [/root/reference]

Only declarations (names, argument kinds, type texts, file:line) are recorded -- the facts a backend has to agree with; no kernel bodies.
Run in the build container, where the reference is mounted; the JSON is committed."""
from __future__ import annotations

import json
import re
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
OUT = ROOT / "integration" / "rust" / "kernel_table.json"
OPEN, CLOSE = "([{<", ")]}>"


def balanced(text: str, start: int) -> int:
    """index just past the bracket group opening at text[start]; `->` and comparison operators do not occur inside signatures' brackets"""
    depth = 0
    i = start
    while i < len(text):
        c = text[i]
        if c in OPEN:
            depth += 1
        elif c in CLOSE:
            if c == ">" and text[i - 1] == "-":
                i += 1
                continue
            depth -= 1
            if depth == 0:
                return i + 1
        i += 1
    raise ValueError("unbalanced")


def split_top(text: str) -> list[str]:
    parts, depth, cur = [], 0, []
    for i, c in enumerate(text):
        if c in OPEN:
            depth += 1
        elif c in CLOSE and not (c == ">" and text[i - 1] == "-"):
            depth -= 1
        if c == "," and depth == 0:
            parts.append("".join(cur).strip())
            cur = []
        else:
            cur.append(c)
    tail = "".join(cur).strip()
    if tail:
        parts.append(tail)
    return parts


def use_map(src: str) -> dict:
    """identifier -> full path for the `use` trees of a file (enough for the gpu_types the signatures mention)"""
    out = {}

    def walk(prefix, tree):
        tree = tree.strip()
        if tree.startswith("{"):
            for part in split_top(tree[1:-1]):
                walk(prefix, part)
            return
        m = re.match(r"([\w:]+?)::\{(.*)\}$", tree, flags=re.S)
        if m:
            for part in split_top(m.group(2)):
                walk(prefix + [m.group(1)], part)
            return
        m = re.match(r"([\w:]+)(?:\s+as\s+(\w+))?$", tree)
        if m:
            path = "::".join(prefix + [m.group(1)]) if prefix else m.group(1)
            if path.endswith("::self"):
                path = path[:-6]
            out[m.group(2) or path.split("::")[-1]] = path

    for m in re.finditer(r"^use\s+(.*?);", src, flags=re.S | re.M):
        walk([], " ".join(m.group(1).split()))
    return out


def classify(ty: str):
    ty = " ".join(ty.split())
    if ty.startswith("*const "):
        return "buffer_read", ty[7:]
    if ty.startswith("*mut "):
        return "buffer_read_write", ty[5:]
    m = re.match(r"&\[(.+);\s*(\w+)\]$", ty)
    if m:
        return "constant_array", f"{m.group(1).strip()}; {m.group(2)}"
    m = re.match(r"&\[(.+)\]$", ty)
    if m:
        return "constant_slice", m.group(1).strip()
    return "scalar", ty


def extract(ref: Path) -> list:
    base = ref / "crates" / "backend-uzu" / "src" / "backends" / "cpu" / "kernel"
    kernels = []
    for path in sorted(base.rglob("*.rs")):
        src = path.read_text()
        uses = use_map(src)
        for m in re.finditer(r"#\[kernel\((\w+)\)\]", src):
            name = m.group(1)
            line = src.count("\n", 0, m.start()) + 1
            fn = re.compile(r"\bfn\s+(\w+)").search(src, m.end())
            attrs = src[m.end():fn.start()]
            variants = [(v.group(1), [x.strip() for x in v.group(2).split(",") if x.strip()])
                        for v in re.finditer(r"#\[variants\((\w+)\s*,([^\]]*)\)\]", attrs)]
            i = fn.end()
            generics = []
            if src[i] == "<":
                j = balanced(src, i)
                for g in split_top(src[i + 1:j - 1]):
                    cm = re.match(r"const\s+(\w+)\s*:\s*(.+)$", g)
                    if cm:
                        generics.append({"name": cm.group(1), "kind": "value", "type": cm.group(2).strip()})
                    else:
                        generics.append({"name": re.match(r"(\w+)", g).group(1), "kind": "type"})
                i = j
            assert src[i] == "(", (path, name)
            j = balanced(src, i)
            args = []
            for a in split_top(src[i + 1:j - 1]):
                optional, specialize = None, False
                while a.startswith("#["):
                    k = balanced(a, 1)
                    attr = a[2:k - 1].strip()
                    if attr == "specialize":
                        specialize = True
                    elif attr.startswith("optional"):
                        optional = attr[len("optional"):].strip()[1:-1].strip()
                    a = a[k:].strip()
                an, _, ty = a.partition(":")
                ty = " ".join(ty.split())
                if optional is not None:
                    om = re.match(r"Option<(.*)>$", ty)
                    assert om, (path, name, a)
                    ty = om.group(1).strip()
                kind, inner = ("specialize", ty) if specialize else classify(ty)
                # full path of imported non-primitive scalar / constant element types (what canonicalize_type_text produces for enums)
                head = re.match(r"[\w:]+", inner)
                resolved = inner
                if head and "::" not in head.group(0) and head.group(0) in uses and kind != "buffer_read" and kind != "buffer_read_write":
                    resolved = uses[head.group(0)] + inner[head.end():]
                args.append({"name": an.strip(), "kind": kind, "type": resolved, "optional": optional})
            assert len(generics) == len(variants) and all(g["name"] == v[0] for g, v in zip(generics, variants)), (path, name)
            for g, v in zip(generics, variants):
                g["variants"] = v[1]
            kernels.append({"name": name, "function": fn.group(1), "file": str(path.relative_to(ref)), "line": line,
                            "module": "::".join(path.relative_to(base).with_suffix("").parts), "generics": generics, "arguments": args})
    return kernels


if __name__ == "__main__":
    ref = Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
    table = extract(ref)
    OUT.write_text(json.dumps({"source": "trymirai/uzu crates/backend-uzu/src/backends/cpu/kernel/**", "kernels": table}, indent=1) + "\n")
    print(f"wrote {OUT} ({len(table)} kernels)")
