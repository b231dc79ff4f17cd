"""Builds uzu_b200/lib/libuzu_b200.so with nvcc for sm_100a (in-tree; the .so travels to the GPU box).

No torch, no cmake: plain `nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo` per source, then one
link step. Objects are cached by (source mtime, flags). `python -m uzu_b200.build [--force] [--verbose]`.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
OBJ = HERE / "build"
LIB = HERE / "lib" / "libuzu_b200.so"
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
HOST_CXX = "/usr/bin/g++"

COMMON = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
          "-Xcompiler", "-fvisibility=hidden", "-ccbin", HOST_CXX, "--expt-relaxed-constexpr", "-I", str(HERE.parent / "include")]
# Per-file flags. Elementwise/normalisation code keeps separate multiply and add roundings (-fmad=false) so it
# rounds exactly where the reference CPU kernels round; the streaming GEMV / attention inner loops may contract.
SOURCES = {
    "runtime.cu": [],
    "matmul.cu": [],
    "prefill_gemm.cu": [],
    "tp.cu": [],
    "attention.cu": ["-fmad=false"],
    "attention_prefill.cu": [],
    "norm.cu": ["-fmad=false"],
    "elementwise.cu": ["-fmad=false"],
    "sampling.cu": ["-fmad=false"],
    "deltanet.cu": ["-fmad=false"],
    "deltanet_prefill.cu": ["-fmad=false"],
    "decode_mega.cu": [],
    "engine.cu": [],
}


def _deps_mtime() -> float:
    hdrs = list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + [HERE.parent / "include" / "uzu_b200.h"]
    return max(h.stat().st_mtime for h in hdrs)


def _compile(src: Path, extra, verbose: bool, force: bool):
    key = hashlib.sha1((" ".join(COMMON + extra)).encode()).hexdigest()[:8]
    obj = OBJ / f"{src.stem}.{key}.o"
    newest = max(src.stat().st_mtime, _deps_mtime())
    if not force and obj.exists() and obj.stat().st_mtime >= newest:
        return obj, None
    cmd = [NVCC, *COMMON, *extra, "-c", str(src), "-o", str(obj)]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
    return obj, (r.stdout + r.stderr)


def build(force: bool = False, verbose: bool = False) -> Path:
    OBJ.mkdir(exist_ok=True)
    LIB.parent.mkdir(exist_ok=True)
    # kernel experiments: UZU_B200_EXTRA_NVCC="-DFOO=1" adds flags to matmul.cu, UZU_B200_LIB_OUT names the output library
    extra = os.environ.get("UZU_B200_EXTRA_NVCC", "").split()
    lib = Path(os.environ["UZU_B200_LIB_OUT"]) if os.environ.get("UZU_B200_LIB_OUT") else LIB
    srcs = [(CSRC / name, flags + (extra if name == "matmul.cu" else [])) for name, flags in SOURCES.items() if (CSRC / name).exists()]
    with ThreadPoolExecutor(max_workers=8) as ex:
        results = list(ex.map(lambda sf: _compile(sf[0], sf[1], verbose, force), srcs))
    objs = [o for o, _ in results]
    logs = [l for _, l in results if l]
    if verbose:
        for l in logs:
            print(l)
    relink = force or bool(logs) or bool(extra) or not lib.exists() or any(o.stat().st_mtime > lib.stat().st_mtime for o in objs)
    if relink:
        cmd = [NVCC, "-shared", "-o", str(lib), *map(str, objs), "-ccbin", HOST_CXX, "-gencode", "arch=compute_100a,code=sm_100a",
               "-cudart", "static", "-Xlinker", "-z,defs", "-lpthread", "-ldl", "-lrt"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return lib


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(p)
