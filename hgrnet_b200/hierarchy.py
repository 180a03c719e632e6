"""Hierarchy index structures, O(N + E).

Produces exactly what the reference's ``gen_tree`` returns (utils.py:39-72) -- ``p2c``,
``c2p``, ``d2n`` (keys in first-occurrence order), ``nodes``, ``start_up`` -- from the same
edge-list JSON (``[[parent_wnid, child_wnid], ...]``, root ``'fall11'``; format written by
data/hierarchical.py:45 / data/remove_irrelevant.py:34), but with dictionary lookups instead
of ``list.index`` (O(N^2) string compares at N ~ 18k in the reference, utils.py:16-20) and
one BFS instead of N ``nx.shortest_path`` calls.  It also derives what the CUDA kernels
consume: per-node depth, and CSR rows / weights for the class-bank aggregation kernel.
"""
from __future__ import annotations

import json
import math
from collections import OrderedDict, defaultdict, deque
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

ROOT = "fall11"  # utils.py:45


class Hierarchy:
    def __init__(self, edges: Sequence[Sequence[str]]):
        succ: "OrderedDict[str, List[str]]" = OrderedDict()
        seen_edge = set()
        for u, v in edges:
            # nx.DiGraph.add_edges_from: nodes enter in first-seen order, duplicate edges collapse
            if u not in succ:
                succ[u] = []
            if v not in succ:
                succ[v] = []
            if (u, v) not in seen_edge:
                seen_edge.add((u, v))
                succ[u].append(v)
        if ROOT not in succ:
            raise ValueError("edge list has no root %r" % ROOT)
        self.nodes: List[str] = [n for n in succ if n != ROOT]          # utils.py:44-45
        self.index: Dict[str, int] = {n: i for i, n in enumerate(self.nodes)}
        self.start_up: List[int] = [self.index[c] for c in succ[ROOT]]  # utils.py:46
        self.p2c: List[List[int]] = [[self.index[c] for c in succ[n]] for n in self.nodes]  # utils.py:48-51

        # one BFS from the root: shortest root->node chain (unique on a tree); on a DAG the
        # first-discovered parent wins (networkx's own tie-break is version dependent)
        parent = {ROOT: None}
        q = deque([ROOT])
        while q:
            u = q.popleft()
            for v in succ[u]:
                if v not in parent:
                    parent[v] = u
                    q.append(v)
        missing = [n for n in self.nodes if n not in parent]
        if missing:
            raise ValueError("%d nodes unreachable from the root (e.g. %s)" % (len(missing), missing[0]))
        N = len(self.nodes)
        self.parent_id = np.full(N, -1, dtype=np.int32)
        for n in self.nodes:
            p = parent[n]
            if p != ROOT:
                self.parent_id[self.index[n]] = self.index[p]
        self.c2p: List[List[int]] = [None] * N  # type: ignore
        # chains root-side first (utils.py:53-56); built top-down so each is parent's chain + parent
        order = []
        q = deque(self.start_up)
        visited = set(self.start_up)
        while q:
            i = q.popleft()
            order.append(i)
            for c in self.p2c[i]:
                if c not in visited and self.parent_id[c] == i:
                    visited.add(c)
                    q.append(c)
        for i in order:
            p = int(self.parent_id[i])
            self.c2p[i] = [] if p < 0 else self.c2p[p] + [p]
        self.depth = np.array([len(c) for c in self.c2p], dtype=np.int32)
        self.d2n: "defaultdict[int, List[int]]" = defaultdict(list)     # utils.py:66-70
        for i in range(N):
            self.d2n[int(self.depth[i])].append(i)
        self.max_depth = int(self.depth.max()) if N else 0

    # ------------------------------------------------------------------ constructors
    @classmethod
    def from_json(cls, path: str) -> "Hierarchy":
        with open(path, "r") as f:
            return cls(json.load(f))

    def as_tuple(self):
        """``(p2c, c2p, d2n, nodes, start_up)`` -- the return value of utils.py:72."""
        return self.p2c, self.c2p, self.d2n, self.nodes, self.start_up

    def __len__(self):
        return len(self.nodes)

    # ------------------------------------------------------------------ CSR for kernel (1)
    def identity_csr(self) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        N = len(self.nodes)
        return (np.arange(N + 1, dtype=np.int32), np.arange(N, dtype=np.int32), np.ones(N, dtype=np.float32))

    def chain_csr(self, ratio: float, level_weights, include_children: bool = False,
                  child_weight: float = 0.0) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        """CSR rows for the hierarchy-aggregated class bank.

        Row ``c`` lists the last ``ceil(ratio * len(chain))`` nodes of ``c2p[c] + [c]`` deepest
        first -- the node set the OM loop walks (model/clip_tree.py:232-237 / :246-251) -- with
        ``level_weights(n)[position]`` as weights (``get_weights``, :198-219).  Optionally the
        direct children are added with a uniform share ``child_weight`` (descendant side of the
        DGP-style aggregation).  ``ratio = 0`` keeps only the node itself (k = 1, :235-236):
        with weight 1 that is the reference's own class bank.
        """
        rowptr = [0]
        col: List[int] = []
        w: List[float] = []
        for c in range(len(self.nodes)):
            chain = self.c2p[c] + [c]
            k = math.ceil(ratio * len(chain))
            if k == 0:
                k = 1
            sel = chain[::-1][:k]
            lw = np.asarray(level_weights(len(sel)), dtype=np.float32)
            col.extend(sel)
            w.extend(float(x) for x in lw)
            if include_children and self.p2c[c] and child_weight != 0.0:
                kids = self.p2c[c]
                col.extend(kids)
                w.extend([child_weight / len(kids)] * len(kids))
            rowptr.append(len(col))
        return np.asarray(rowptr, np.int32), np.asarray(col, np.int32), np.asarray(w, np.float32)


# ---------------------------------------------------------------------- synthetic hierarchies
WORDNET_LIKE_21841 = [20, 150, 900, 3000, 5500, 5500, 3500, 1800, 900, 400, 120, 51]  # SURVEY.md section 8d cfg 2


def scaled_levels(total: int, template: Sequence[int] = WORDNET_LIKE_21841) -> List[int]:
    """Scale a level-size template so that it sums to ``total`` (every level >= 1)."""
    s = float(sum(template))
    sizes = [max(1, int(round(x * total / s))) for x in template]
    diff = total - sum(sizes)
    big = max(range(len(sizes)), key=lambda i: sizes[i])
    sizes[big] += diff
    assert sizes[big] >= 1 and sum(sizes) == total
    return sizes


def wnid_of(i: int) -> str:
    return "n%08d" % (i + 1)


def synthetic_tree_edges(level_sizes: Sequence[int], seed: int = 0) -> List[List[str]]:
    """A random tree with the given number of nodes per depth, as a reference-format edge list.

    Nodes are numbered level by level (so ``nodes`` order == id order); every node of level
    ``d > 0`` gets a parent drawn from level ``d-1`` such that each parent has >= 1 child
    whenever the level sizes allow it.
    """
    rng = np.random.RandomState(seed)
    edges: List[List[str]] = []
    start = 0
    prev: List[int] = []
    for d, n in enumerate(level_sizes):
        ids = list(range(start, start + n))
        if d == 0:
            edges.extend([ROOT, wnid_of(i)] for i in ids)
        else:
            parents = list(prev[: min(len(prev), n)])          # every parent gets one child first
            if n > len(parents):
                parents += list(rng.choice(prev, size=n - len(parents)))
            rng.shuffle(parents)
            edges.extend([wnid_of(int(p)), wnid_of(i)] for p, i in zip(parents, ids))
        prev = ids
        start += n
    return edges


def synthetic_hierarchy(level_sizes: Sequence[int], seed: int = 0) -> Hierarchy:
    return Hierarchy(synthetic_tree_edges(level_sizes, seed))
