"""uzu_b200: a B200-native CUDA backend for uzu's transformer decode hot path.

The product is `libuzu_b200.so` (C ABI in include/uzu_b200.h: hand-written sm_100a kernels plus
the C++ host engine mirroring crates/backend-uzu's Engine/LanguageModel/Decoder call pattern).
This Python package is a thin ctypes binding used by tests and bench.py, plus the synthetic
checkpoint generator. There is no CPU fallback: every compute entry point fails loudly when
the shared library or a CUDA device is missing.
"""
__all__ = ["synth", "safetensors_io"]
