"""One speculation pass of a BASELINE workload (ncu target): python tools/trie_short.py <workload> <prefill> <nodes>"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import bench
from uzu_b200 import binding as B
from uzu_b200.trie import PRng, TrieNode
workload, prefill, nodes = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
ctx = B.Context(0)
eng = B.Engine(ctx, bench.model_dir_for(workload), max_context_length=max(1024, prefill + 64))
rng = np.random.default_rng(0)
tok = eng.prefill(rng.integers(0, eng.info.vocab_size, prefill).astype(np.uint32))
flat = TrieNode.flat(prefill, [tok] + [int(t) for t in rng.integers(0, eng.info.vocab_size, nodes - 1)], PRng(0)).linearize()
for _ in range(2):
    sampled = eng.trie_pass(flat.token_ids(), flat.nodes())
    eng.trie_accept([0], sampled[0])
print("ok", sampled[:4])
eng.close(); ctx.close()
