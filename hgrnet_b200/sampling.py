"""Negative-class sampling of the training step (host side, integer set arithmetic).

Mirrors ``tree_model.get_contra`` (model/clip_tree.py:80-196) for the strategies that need no
encoder call.  ``--sample_strategy topk`` is NOT a top-k over logits: it is the "TopM" set of
nodes in the ``--k`` levels above the anchor's depth, minus the anchor's own chain, randomly
sub-sampled to ``--num_compare`` with Python's ``random.sample``.  The
draw order of Python's global ``random`` module is preserved so that a seeded run picks the
same classes as the reference.
"""
from __future__ import annotations

import copy
import math
import random as _random
from typing import List, Optional, Sequence, Tuple


_LIST_CACHE_ENTRIES = 512


def contra_topk(d2n, target: int, depth: int, parents: Sequence[int], k: int, num_compare: int,
                rng=_random, cache: Optional[dict] = None) -> Tuple[List[int], int]:
    """model/clip_tree.py:116-141.  Returns ``(compare_idx, label_position)``.

    ``cache`` (optional dict owned by the caller) memoises (a) the candidate SET of a depth window: it only depends
    on ``(low, depth)``, and a set built from the same insertion sequence has the same iteration order; (b) the
    candidate LIST after the anchor chain has been removed, keyed by ``(low, depth, chain)`` -- a class comes back
    for several batches per epoch and for every epoch, and the set difference over thousands of nodes is what is
    left of the step's host time (bounded to ``_LIST_CACHE_ENTRIES`` lists, oldest evicted first).  Either way the
    list handed to ``random.sample`` -- and therefore the draw -- is exactly what rebuilding it every call (as the
    reference does, :125-131) would give: the same objects built by the same operations.
    """
    low = min(d2n.keys())
    if depth - k > low:
        low = depth - k
    key = (low, depth)
    candi_set = cache.get(key) if cache is not None else None
    if candi_set is None:
        candi: List[int] = []
        for d in range(low, depth):
            candi.extend(d2n[d])
        if depth == 0:
            candi.extend(d2n[depth])
        candi_set = set(candi)
        if cache is not None:
            cache[key] = candi_set
    lkey = ("list", low, depth, tuple(parents))
    base = cache.get(lkey) if cache is not None else None
    if base is None:
        base = list(candi_set - set(parents))
        if cache is not None:
            order = cache.setdefault("_list_keys", [])
            if len(order) >= _LIST_CACHE_ENTRIES:
                cache.pop(order.pop(0), None)
            order.append(lkey)
            cache[lkey] = base
    if len(base) > num_compare:
        compare_idx = rng.sample(base, num_compare)   # a new list; the cached one is never modified
    else:
        compare_idx = list(base)
    if target not in compare_idx:
        compare_idx.append(target)
    return compare_idx, compare_idx.index(target)


def contra_random(train_ids: Sequence[int], target: int, num_compare: int, rng=_random) -> Tuple[List[int], int]:
    """model/clip_tree.py:81-89."""
    compare_idx = rng.sample(list(train_ids), num_compare)
    if target not in compare_idx:
        compare_idx.append(target)
    return compare_idx, compare_idx.index(target)


def contra_brothers(p2c, start_up, target: int, depth: int, parents: Sequence[int], num_compare: int,
                    rng=_random) -> Tuple[List[int], int]:
    """model/clip_tree.py:180-196: siblings under the chain node one level up (or the top level)."""
    if len(parents) > 1 and depth > 0:
        compare_idx = copy.copy(p2c[parents[depth - 1]])
    else:
        compare_idx = copy.copy(start_up)
    if len(compare_idx) > num_compare:
        compare_idx = rng.sample(compare_idx, num_compare)
    if target not in compare_idx:
        compare_idx.append(target)
    return compare_idx, compare_idx.index(target)


def om_schedule(c2p, target: int, out_ratio: float, in_ratio: float):
    """The (outer anchor, inner level) iteration space of the OM step, model/clip_tree.py:228-256.

    Returns a list of ``(k_loop, m_loop, p_out, depth, parents_in, n_out, n_in)``: the positive
    class is always the OUTER anchor ``p_out``; ``depth`` (position of the inner node in
    ``parents_in``) only selects the depth window negatives are drawn from.
    """
    parents = list(c2p[target]) + [target]
    k = math.ceil(out_ratio * len(parents))
    if k == 0:
        k = 1
    p_loop_out = parents[::-1][:k]
    out = []
    for k_loop, p_out in enumerate(p_loop_out):
        parents_in = list(c2p[p_out]) + [p_out]
        m = math.ceil(in_ratio * len(parents_in))
        if m == 0:
            m = 1
        p_loop_in = parents_in[::-1][:m]
        for m_loop, p_in in enumerate(p_loop_in):
            out.append((k_loop, m_loop, p_out, parents_in.index(p_in), parents_in, len(p_loop_out), len(p_loop_in)))
    return out


def hierarchical_schedule(c2p, target: int):
    """Iteration space of ``training_method == 'hierarchical'`` (model/clip_tree.py:283-312):
    one iteration per chain level, positive = the label itself, weight index = level."""
    parents = list(c2p[target]) + [target]
    return [(j, target, j, parents, len(parents)) for j in range(len(parents))]
