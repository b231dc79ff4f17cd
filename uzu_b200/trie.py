"""Host-side speculation trie: the mirror of the reference's `TrieNode` / `FlatTrie` (crates/backend-uzu/src/trie.rs) and `PRng`
(src/encodable_block/sampling/prng.rs:12-23) that the decode stream runs on the CPU around a speculation pass
(src/engine/language_model/stream/stream.rs:550-657). In the Rust drop-in this logic stays the reference's own; the Python host of this
repository needs it to drive `Engine.trie_pass` / `Engine.trie_accept` the way the stream does:

    trie = proposer(...)                       # a TrieNode tree rooted at the last sampled token
    flat = trie.linearize()                    # depth-first order, (trie_start, trie_end, height) per node
    sampled = engine.trie_pass(flat.token_ids(), flat.nodes(), flat.seeds())
    full = flat.accept(sampled)                # [(index, input token, sampled token)] along the verified path
    engine.trie_accept([i for i, _, _ in full])

Behaviour is pinned by the reference's own unit tests (tests/unit/trie_test.rs), restated in tests/test_trie_cpu.py."""
from __future__ import annotations

import numpy as np

_M64 = (1 << 64) - 1


class PRng:
    """PRng::derive (prng.rs:12-23): a murmur3 finaliser over seed + index."""

    def __init__(self, seed: int):
        self.seed = seed & _M64

    def derive(self, index: int) -> int:
        h = (self.seed + index) & _M64
        h ^= h >> 33
        h = (h * 0xff51afd7ed558ccd) & _M64
        h ^= h >> 33
        h = (h * 0xc4ceb9fe1a85ec53) & _M64
        h ^= h >> 33
        return h


class DuplicateTokenId(ValueError):
    """TrieError::DuplicateTokenId (trie.rs:12-16)."""


class TrieNode:
    """trie.rs:25-31."""

    __slots__ = ("token", "seed", "logprob", "next")

    def __init__(self, token: int, seed: int, logprob: float = 0.0):
        self.token, self.seed, self.logprob = int(token), int(seed) & _M64, float(logprob)
        self.next: list[TrieNode] = []

    def add(self, node: "TrieNode") -> int:
        """trie.rs:60-70."""
        if any(n.token == node.token for n in self.next):
            raise DuplicateTokenId("child with the same token id is already present")
        self.next.append(node)
        return len(self.next) - 1

    def get(self, token: int):
        """trie.rs:72-77."""
        for n in self.next:
            if n.token == token:
                return n
        return None

    def node_count(self) -> int:
        return 1 + sum(n.node_count() for n in self.next)

    def prune_to_budget(self, budget: int) -> None:
        """trie.rs:94-137: keep the `budget` nodes with the largest cumulative log-probability (stable order on ties, so a parent --
        visited first, never below its child -- is kept before the child); pruned children drop with their subtrees."""
        assert budget > 0, "budget must keep at least the root"
        logprobs: list[float] = []

        def collect(node, parent):
            lp = float(np.float32(parent) + np.float32(node.logprob))
            logprobs.append(lp)
            for c in node.next:
                collect(c, lp)

        collect(self, 0.0)
        if budget >= len(logprobs):
            return
        order = sorted(range(len(logprobs)), key=lambda i: -logprobs[i])    # stable, descending (sort_by total_cmp b vs a)
        kept = [False] * len(logprobs)
        for i in order[:budget]:
            kept[i] = True
        cursor = [0]

        def prune(node):
            cursor[0] += 1
            survivors = []
            for c in node.next:
                index = cursor[0]
                prune(c)
                if kept[index]:
                    survivors.append(c)
            node.next = survivors

        prune(self)

    @staticmethod
    def flat(prefix_length: int, tokens, prng: PRng) -> "TrieNode":
        """trie.rs:139-156: a chain (prefill chunks / plain decode), node i seeded with derive(prefix_length + i)."""
        tokens = [int(t) for t in tokens]
        assert tokens, "need seed node"
        root = TrieNode(tokens[0], prng.derive(prefix_length))
        leaf = root
        for i, t in enumerate(tokens[1:], start=1):
            leaf.add(TrieNode(t, prng.derive(prefix_length + i)))
            leaf = leaf.next[0]
        return root

    def linearize(self) -> "FlatTrie":
        """trie.rs:158-178: depth-first pre-order; subtrie_range = [own index, index of the last node of the subtree]."""
        entries = [[self, 0, 0, 0]]        # node, start, end, height
        stack = [[0, 0]]
        while stack:
            cur, child = stack[-1]
            nxt = entries[cur][0].next
            if child >= len(nxt):
                entries[cur][2] = len(entries) - 1
                stack.pop()
                continue
            stack[-1][1] += 1
            node = nxt[child]
            entries.append([node, len(entries), len(entries), len(stack)])
            if node.next:
                stack.append([len(entries) - 1, 0])
        return FlatTrie(entries)


class FlatTrie:
    """trie.rs:39-42, 196-296."""

    def __init__(self, entries):
        self._entries = entries

    def __len__(self) -> int:
        return len(self._entries)

    def token_ids(self) -> list[int]:
        return [e[0].token for e in self._entries]

    def token_seeds(self) -> list[int]:
        return [e[0].seed for e in self._entries]

    seeds = token_seeds

    def heights(self) -> list[int]:
        return [e[3] for e in self._entries]

    def nodes(self) -> np.ndarray:
        """token_subtrie_ranges (trie.rs:211-222): u32 [len, 3] rows of gpu_types::trie::TrieNode {trie_start, trie_end, height}."""
        return np.array([[e[1], e[2], e[3]] for e in self._entries], dtype=np.uint32).reshape(len(self._entries), 3)

    def parents(self) -> list[int]:
        """BatchTopology::new (encodable_block/batch_topology.rs:15-27)."""
        stack: list[int] = []
        out = []
        for i, e in enumerate(self._entries):
            del stack[e[3]:]
            out.append(stack[-1] if stack else -1)
            stack.append(i)
        return out

    def is_flat(self) -> bool:
        return all(e[3] == i for i, e in enumerate(self._entries))

    def root(self):
        return self._entries[0][0] if self._entries else None

    def index(self, node) -> int | None:
        """Identity (pointer) lookup, trie.rs:255-260."""
        for i, e in enumerate(self._entries):
            if e[0] is node:
                return i
        return None

    def accept(self, sampled_tokens) -> list[tuple[int, int, int]]:
        """trie.rs:262-296 without a grammar: walk from the root; at each node take what the model sampled there, and descend only if the
        trie proposed exactly that token. Returns (flat index, input token, sampled token) per verified node; the last entry's sampled
        token is the fresh one no proposal covered."""
        cur = self.root()
        out = []
        while True:
            i = self.index(cur)
            tok = int(sampled_tokens[i])
            out.append((i, cur.token, tok))
            nxt = cur.get(tok)
            if nxt is None:
                return out
            cur = nxt
