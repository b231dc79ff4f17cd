"""Builds uzu_b200/lib/libuzu_b200.so with nvcc for sm_100a (in-tree; the .so travels to the GPU box).

No torch, no cmake: plain `nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo` per source, then one
link step. Objects are cached by the content hash of (source, headers, flags). `python -m uzu_b200.build [--force] [--verbose]`.
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
    "activation_transform.cu": ["-fmad=false"],
    "sampling.cu": ["-fmad=false"],
    "deltanet.cu": ["-fmad=false"],
    "deltanet_prefill.cu": ["-fmad=false"],
    "decode_mega.cu": [],
    "engine.cu": [],
}


def _deps_digest() -> bytes:
    hdrs = sorted(list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + [HERE.parent / "include" / "uzu_b200.h"])
    h = hashlib.sha1()
    for f in hdrs:
        h.update(f.name.encode())
        h.update(f.read_bytes())
    return h.digest()


def _compile(src: Path, extra, verbose: bool, force: bool):
    # the cache key is the CONTENT of the source, of every header and the flags (not mtimes: the gpurun snapshot does not keep them, and a
    # stale-looking object would trigger a multi-minute rebuild on the GPU box)
    flags = " ".join(f for f in COMMON + extra if not f.startswith(str(HERE.parent)))      # the -I path is a location, not content
    key = hashlib.sha1(flags.encode() + src.read_bytes() + _deps_digest()).hexdigest()[:12]
    obj = OBJ / f"{src.stem}.{key}.o"
    if not force and obj.exists():
        return obj, None
    if not os.environ.get("UZU_B200_LIB_OUT"):          # experiment builds keep the default objects
        for old in OBJ.glob(f"{src.stem}.*.o"):
            old.unlink()
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
    srcs = [(CSRC / name, flags + (extra if name in ("matmul.cu", "decode_mega.cu") else [])) for name, flags in SOURCES.items() if (CSRC / name).exists()]
    with ThreadPoolExecutor(max_workers=8) as ex:
        results = list(ex.map(lambda sf: _compile(sf[0], sf[1], verbose, force), srcs))
    objs = [o for o, _ in results]
    logs = [l for _, l in results if l]
    if verbose:
        for l in logs:
            print(l)
    manifest = lib.with_suffix(".objects")
    want = "\n".join(o.name for o in objs)
    relink = force or bool(extra) or not lib.exists() or not manifest.exists() or manifest.read_text() != want
    if relink:
        cmd = [NVCC, "-shared", "-o", str(lib), *map(str, objs), "-ccbin", HOST_CXX, "-gencode", "arch=compute_100a,code=sm_100a",
               "-cudart", "static", "-Xlinker", "-z,defs", "-lpthread", "-ldl", "-lrt"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
        manifest.write_text(want)
    return lib


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(p)
