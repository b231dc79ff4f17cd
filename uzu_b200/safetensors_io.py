"""Minimal safetensors reader/writer (numpy only) for uzu-format checkpoints.

Format as consumed by the reference loader (parameters/safetensors_metadata.rs:93-127,
parameters/loader.rs:64-101,219-228): 8-byte little-endian header length, JSON header
{name: {dtype, shape, data_offsets}, "__metadata__": {str: str}}, then raw tensor bytes.
Per-matrix quantisation specs are JSON *strings* under `<prefix>.spec` in `__metadata__`.
bf16 tensors are carried as numpy uint16 with dtype tag "BF16".
"""
from __future__ import annotations

import json
import struct
from pathlib import Path

import numpy as np

_NP2ST = {np.dtype("float32"): "F32", np.dtype("uint8"): "U8", np.dtype("int8"): "I8", np.dtype("int32"): "I32",
          np.dtype("uint32"): "U32", np.dtype("int64"): "I64", np.dtype("uint64"): "U64", np.dtype("float16"): "F16"}
_ST2NP = {v: k for k, v in _NP2ST.items()}
_ST2NP["BF16"] = np.dtype("uint16")


class BF16(np.ndarray):
    """Marker subclass: a uint16 array that is to be stored as BF16."""


def as_bf16(bits: np.ndarray) -> np.ndarray:
    assert bits.dtype == np.uint16
    return bits.view(BF16)


def save(path, tensors: dict, metadata: dict | None = None) -> None:
    header = {}
    offset = 0
    order = sorted(tensors)
    for name in order:
        t = tensors[name]
        tag = "BF16" if isinstance(t, BF16) else _NP2ST[t.dtype]
        nbytes = t.size * t.dtype.itemsize
        header[name] = {"dtype": tag, "shape": list(t.shape), "data_offsets": [offset, offset + nbytes]}
        offset += nbytes
    if metadata:
        header["__metadata__"] = {k: (v if isinstance(v, str) else json.dumps(v)) for k, v in metadata.items()}
    blob = json.dumps(header, separators=(",", ":")).encode()
    blob += b" " * ((8 - len(blob) % 8) % 8)
    with open(path, "wb") as f:
        f.write(struct.pack("<Q", len(blob)))
        f.write(blob)
        for name in order:
            f.write(np.ascontiguousarray(tensors[name]).tobytes())


def load(path):
    """Returns (tensors: dict[name -> ndarray (bf16 as uint16)], dtypes: dict[name -> tag], metadata)."""
    path = Path(path)
    with open(path, "rb") as f:
        (hlen,) = struct.unpack("<Q", f.read(8))
        header = json.loads(f.read(hlen))
    data = np.memmap(path, dtype=np.uint8, mode="r", offset=8 + hlen)
    meta = header.pop("__metadata__", {})
    tensors, dtypes = {}, {}
    for name, info in header.items():
        b, e = info["data_offsets"]
        dt = _ST2NP[info["dtype"]]
        tensors[name] = np.frombuffer(data[b:e], dtype=dt).reshape(info["shape"])
        dtypes[name] = info["dtype"]
    return tensors, dtypes, meta
