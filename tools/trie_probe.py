"""Speculation-pass cost on a BASELINE workload: python tools/trie_probe.py <workload> <prefill>
Times uzu_engine_trie_pass (one Decoder::encode over a linearized trie, readout + sampling for every node) at 1/2/4/8/16 nodes against
the plain decode step, wall clock around the synchronous call (each pass ends with the D2H read of the sampled ids, like the stream).
The ratio pass(m) / step(1) is what an accepted-tokens-per-pass figure has to beat for speculation to pay on this GPU."""
import json
import sys
import time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import bench
from uzu_b200 import binding as B
from uzu_b200.trie import PRng, TrieNode

workload, prefill = sys.argv[1], int(sys.argv[2])
ctx = B.Context(0)
eng = B.Engine(ctx, bench.model_dir_for(workload), max_context_length=max(1024, prefill + 256))
rng = np.random.default_rng(0)
V = eng.info.vocab_size
tok = eng.prefill(rng.integers(0, V, prefill).astype(np.uint32))
out = {"workload": workload, "prefill": prefill, "decode_path": "persistent" if eng.persistent_decode else "kernels"}
for _ in range(8):
    tok = eng.step_host(tok)
t0 = time.perf_counter()
for _ in range(32):
    tok = eng.step_host(tok)
out["step_host_ms"] = (time.perf_counter() - t0) / 32 * 1e3
out["decode_timed_ms"] = eng.decode_timed(64) / 64 * 1e3
eng.flush()
rows = []
for m in (1, 2, 4, 8, 16):
    # a bushy trie: root + chains, the shape a draft model proposes; the cost depends on the node count, not on the shape
    root = TrieNode(tok, 0)
    node, made = root, 1
    while made < m:
        child = TrieNode(int(rng.integers(0, V)), 0)
        try:
            node.add(child)
        except Exception:
            continue
        made += 1
        if made % 3 == 0:
            node = child
    flat = root.linearize()
    ms = []
    for it in range(12):
        t0 = time.perf_counter()
        sampled = eng.trie_pass(flat.token_ids(), flat.nodes())
        t1 = time.perf_counter()
        eng.trie_accept([0], sampled[0])      # keep the root only: context grows by one per pass, like a rejected proposal
        ms.append((t1 - t0) * 1e3)
    rows.append({"nodes": m, "pass_ms": float(np.median(ms[2:])), "accept_ms_excluded": True})
    tok = sampled[0]
out["trie_pass"] = rows
for r in rows:
    r["vs_plain_step"] = r["pass_ms"] / out["decode_timed_ms"]
print(json.dumps(out))
eng.close(); ctx.close()
