"""Hierarchy index structures, O(N + E).

Produces what the reference's ``gen_tree`` returns (utils.py:39-72) -- ``p2c``,
``c2p``, ``d2n`` (keys in first-occurrence order), ``nodes``, ``start_up`` -- from the same
edge-list JSON (``[[parent_wnid, child_wnid], ...]``, root ``'fall11'``; format written by
data/hierarchical.py:45 / data/remove_irrelevant.py:34), but with dictionary lookups instead
of ``list.index`` (O(N^2) string compares at N ~ 18k in the reference, utils.py:16-20) and
one BFS instead of N ``nx.shortest_path`` calls (nodes with several shortest root paths replay networkx's
own bidirectional search, so multi-parent DAGs give the reference's chains as well).  It also derives what the CUDA kernels
consume: per-node depth, and CSR rows / weights for the class-bank aggregation kernel.
"""
from __future__ import annotations

import json
import math
from collections import OrderedDict, defaultdict, deque
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

ROOT = "fall11"  # utils.py:45


def _nx_bidirectional_path(succ, pred_of, source, target):
    """The path `networkx.shortest_path(G, source, target)` returns on an unweighted DiGraph: a restatement of
    networkx 3.6.1 `_bidirectional_pred_succ` / `bidirectional_shortest_path` (the dependency the reference calls at
    utils.py:55; not vendored in the reference).  Two BFS fronts, the smaller one (forward on a tie) advances a whole
    level; neighbours are visited in adjacency insertion order; the first node seen by both fronts joins the halves.
    Pinned against networkx itself on random DAGs by tests/test_cpu_oracle_and_host.py."""
    if source == target:
        return [source]
    pred = {source: None}
    nxt = {target: None}
    forward, reverse = [source], [target]
    meet = None
    while forward and reverse and meet is None:
        if len(forward) <= len(reverse):
            level, forward = forward, []
            for v in level:
                for w in succ[v]:
                    if w not in pred:
                        forward.append(w)
                        pred[w] = v
                    if w in nxt:
                        meet = w
                        break
                if meet is not None:
                    break
        else:
            level, reverse = reverse, []
            for v in level:
                for w in pred_of[v]:
                    if w not in nxt:
                        nxt[w] = v
                        reverse.append(w)
                    if w in pred:
                        meet = w
                        break
                if meet is not None:
                    break
    if meet is None:
        raise ValueError("no path between %r and %r" % (source, target))
    path = []
    w = meet
    while w is not None:
        path.append(w)
        w = pred[w]
    path.reverse()
    w = nxt[path[-1]]
    while w is not None:
        path.append(w)
        w = nxt[w]
    return path


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

        # Ancestor chains (utils.py:53-56): `nx.shortest_path(G, 'fall11', node)[1:-1]`.  One BFS from the root gives
        # every node's distance and the NUMBER of shortest root paths.  Where that path is unique (every node of a
        # tree) the chain is the BFS parent chain.  A node with several shortest paths (the real graph is a DAG:
        # multi-parent wnids, data/remove_irrelevant.py) gets the path networkx itself would return: its
        # bidirectional search is replayed for that node (`_nx_bidirectional_path`), so the OM anchor chains and the
        # TOR/POR chains equal the reference's on real data too.
        # nx.DiGraph keeps a node's predecessors in edge-insertion order
        pred_of: Dict[str, List[str]] = {n: [] for n in succ}
        seen_edge2 = set()
        for u, v in edges:
            if (u, v) not in seen_edge2:
                seen_edge2.add((u, v))
                pred_of[v].append(u)
        parent = {ROOT: None}
        dist = {ROOT: 0}
        npaths = {ROOT: 1}
        q = deque([ROOT])
        while q:
            u = q.popleft()
            for v in succ[u]:
                if v not in parent:
                    parent[v] = u
                    dist[v] = dist[u] + 1
                    npaths[v] = npaths[u]
                    q.append(v)
                elif dist[v] == dist[u] + 1:
                    npaths[v] = min(npaths[v] + npaths[u], 2)
        missing = [n for n in self.nodes if n not in parent]
        if missing:
            raise ValueError("%d nodes unreachable from the root (e.g. %s)" % (len(missing), missing[0]))
        N = len(self.nodes)
        self.c2p: List[List[int]] = [None] * N  # type: ignore
        self.n_multipath = 0
        for n in sorted(self.nodes, key=lambda x: dist[x]):        # shallow first: a unique chain extends its parent's
            i = self.index[n]
            if npaths[n] == 1:
                p = parent[n]
                self.c2p[i] = [] if p == ROOT else self.c2p[self.index[p]] + [self.index[p]]
            else:
                self.n_multipath += 1
                path = _nx_bidirectional_path(succ, pred_of, ROOT, n)
                self.c2p[i] = [self.index[x] for x in path[1:-1]]
        self.parent_id = np.full(N, -1, dtype=np.int32)            # last node of the chain (the chain's own parent)
        for i in range(N):
            if self.c2p[i]:
                self.parent_id[i] = self.c2p[i][-1]
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
