"""Regenerates tests/golden/*.npz with the CPU oracle (run from the repo root: python tests/golden/make_golden.py).

These are ORACLE-generated regression vectors (the reference stores no outputs for this path and cannot be run
here -- see DESIGN.md section 2): they freeze the oracle's behaviour and give the GPU tests fixed targets that do
not depend on the oracle library being rebuilt."""
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import oracle as O  # noqa: E402
from oracle.model import OracleModel  # noqa: E402
from uzu_b200 import synth  # noqa: E402

OUT = Path(__file__).resolve().parent


def quant_case(seed, m, n, k, bits, gs, method):
    rng = np.random.default_rng(seed)
    groups = -(-k // gs)
    w = rng.integers(0, 256, (n, k // 2 if bits == 4 else k), dtype=np.uint8)
    sc = O.f32_to_bf16(rng.uniform(0.01, 0.3, (n, groups)).astype(np.float32))
    zp = rng.integers(0, 256, (n, -(-groups // 2) if bits == 4 else groups), dtype=np.uint8) if method == O.QM_ZERO_POINT else None
    bi = O.f32_to_bf16(rng.uniform(-0.03, 0.03, (n, groups)).astype(np.float32)) if method == O.QM_SCALE_BIAS else None
    x = O.f32_to_bf16(rng.uniform(-0.3, 0.3, (m, k)).astype(np.float32))
    return x, w, sc, zp, bi


def main():
    cases = {}
    for name, (seed, m, n, k, bits, gs, method) in {
        "int4_gs64_zp": (1, 1, 96, 1024, 4, 64, O.QM_ZERO_POINT), "int8_gs64_zp": (2, 3, 64, 512, 8, 64, O.QM_ZERO_POINT),
        "mlx4_gs32": (3, 2, 80, 512, 4, 32, O.QM_SCALE_BIAS), "sym4_gs128": (4, 1, 48, 1024, 4, 128, O.QM_SYMMETRIC),
        # prefill row counts (m >= 64: the tcgen05 GEMM with the in-kernel dequant stage on the GPU side)
        "prefill_int4_gs64_zp": (5, 96, 136, 320, 4, 64, O.QM_ZERO_POINT), "prefill_int8_gs32_mlx": (6, 70, 96, 256, 8, 32, O.QM_SCALE_BIAS)}.items():
        x, w, sc, zp, bi = quant_case(seed, m, n, k, bits, gs, method)
        d = O.matmul(x, w, m=m, n=n, k=k, scales=sc, zero_points=zp, biases=bi, method=method, bits=bits, group_size=gs, d_f32=True)
        cases[f"matmul_{name}_seed"] = np.array([seed, m, n, k, bits, gs, method])
        cases[f"matmul_{name}_out"] = d
    np.savez_compressed(OUT / "matmul.npz", **cases)

    models = {}
    for kind in ("llama", "qwen-hybrid"):
        spec = synth.tiny(kind)
        with tempfile.TemporaryDirectory() as td:
            path = synth.write_model(spec, Path(td) / "m", seed=21)
            prompt = (np.arange(14) * 53 + 7) % spec.vocab_size
            toks, logits = OracleModel(path, max_context=64).generate(prompt, 4)
            models[f"{kind}_prompt"] = prompt.astype(np.uint32)
            models[f"{kind}_tokens"] = np.array(toks, np.uint32)
            models[f"{kind}_logits"] = np.stack([l[0] for l in logits])   # bf16 bits, [4, vocab]
    np.savez_compressed(OUT / "tiny_models.npz", **models)
    print("wrote", [p.name for p in OUT.glob("*.npz")])


if __name__ == "__main__":
    main()
