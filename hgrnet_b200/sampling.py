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


# ---- draw-identical fast `random.sample` ------------------------------------------------------------------------
# The reference draws its negatives with Python's `random.sample` (clip_tree.py:134), ~0.15 ms per call at
# ImageNet-21K level sizes -- 17 calls per OM step, by far the largest item of the step once the GPU side is fused.
# A `SampleStream` replays those calls several times faster with IDENTICAL results: the Mersenne-Twister words the
# reference would consume one `getrandbits` call at a time are fetched in bulk (a k-bit `getrandbits` concatenates
# consecutive 32-bit outputs, least significant first), CPython's two selection algorithms run on that word array
# (numpy for the set-based one, a tight loop for the pool-based one), and when the stream is closed the generator is
# rewound and advanced by exactly the number of words the reference calls consume -- so `random` is in the same state
# afterwards as after the reference's step.  Restates CPython's `Random.sample` / `_randbelow_with_getrandbits`
# (3.8 - 3.13); `_fast_sample_ok` checks the restatement against the running interpreter once and falls back to
# `rng.sample` on any mismatch.
def _sample_setsize(k: int) -> int:
    setsize = 21
    if k > 5:
        setsize += 4 ** math.ceil(math.log(k * 3, 4))
    return setsize


class SampleStream:
    """Bulk view of the words `rng` will produce.  Use as a context manager around a run of `sample` calls; nothing
    else may draw from `rng` in between."""

    def __init__(self, rng):
        self.rng = rng
        self.state = None
        self.words = None
        self.pos = 0          # words consumed so far
        self._words_ptr = 0
        self._scratch = None
        self._scratch_ptr = 0

    def __enter__(self):
        import numpy as np
        self.state = self.rng.getstate()
        self.words = np.empty(0, dtype="<u4")
        self.pos = 0
        return self

    def ensure(self, m: int):
        """At least `m` unconsumed words in the buffer."""
        import numpy as np
        have = self.words.shape[0] - self.pos
        if have < m:
            extra = max(m - have, 8192)
            new = np.frombuffer(self.rng.getrandbits(32 * extra).to_bytes(4 * extra, "little"), dtype="<u4")
            self.words = np.concatenate([self.words[self.pos:], new])
            self._base = getattr(self, "_base", 0) + self.pos
            self.pos = 0
        self._words_ptr = self.words.__array_interface__["data"][0]

    def consumed(self) -> int:
        return getattr(self, "_base", 0) + self.pos

    def __exit__(self, *exc):
        # the generator is ahead by everything fetched: rewind, then take exactly what the reference calls consume
        used = self.consumed()
        self.rng.setstate(self.state)
        if used:
            self.rng.getrandbits(32 * used)
        return False

    # -- the selected POSITIONS of random.sample(range(n), k), by the library's host helper (hgr_sample_replay: the same
    #    two algorithms as below in C, ~2 us per call instead of ~40)
    def positions(self, n: int, k: int):
        import numpy as np
        if not 0 <= k <= n:
            raise ValueError("Sample larger than population or is negative")
        lib = _host_lib()
        setsize = _sample_setsize(k)
        out = np.empty(k, dtype=np.int32)
        if self._scratch is None or self._scratch.shape[0] < n:
            self._scratch = np.empty(max(n, 4096), dtype=np.int32)
            self._scratch_ptr = self._scratch.__array_interface__["data"][0]
        m = 2 * k + 64 if n <= setsize else int(k * (1 << n.bit_length()) / max(n, 1) * 1.25) + 64
        while True:
            self.ensure(m)
            avail = self.words.shape[0] - self.pos
            # (`ndarray.ctypes` costs ~10 us per access; the array interface gives the same address for ~1 us)
            used = lib.hgr_sample_replay(self._words_ptr + 4 * self.pos, avail, n, k, setsize,
                                         out.__array_interface__["data"][0], self._scratch_ptr)
            if used >= 0:
                self.pos += int(used)
                return out
            if used != -1:
                raise ValueError("hgr_sample_replay rejected n = %d, k = %d" % (n, k))
            m = 2 * max(m, avail)

    def sample_arrays(self, populations, ks):
        """A run of `random.sample(populations[c], ks[c])` calls (numpy populations) in ONE library call."""
        import numpy as np
        lib = _host_lib()
        if lib is None:
            return [np.asarray(self._sample_py(p.tolist(), int(k)), dtype=p.dtype) for p, k in zip(populations, ks)]
        cnt = len(populations)
        if cnt == 0:
            return []
        ns = np.fromiter((p.shape[0] for p in populations), dtype=np.int64, count=cnt)
        kk = np.asarray(ks, dtype=np.int64)
        if (kk > ns).any() or (kk < 0).any():
            raise ValueError("Sample larger than population or is negative")
        total = int(kk.sum())
        out = np.empty(max(total, 1), dtype=np.int32)
        nmax = int(ns.max())
        if self._scratch is None or self._scratch.shape[0] < nmax:
            self._scratch = np.empty(max(nmax, 4096), dtype=np.int32)
            self._scratch_ptr = self._scratch.__array_interface__["data"][0]
        m = int(1.6 * total) + 32 * cnt        # ~1.5 words per draw at ImageNet-21K level sizes; doubled when short
        while True:
            self.ensure(m)
            avail = self.words.shape[0] - self.pos
            used = lib.hgr_sample_replay_many(self._words_ptr + 4 * self.pos, avail, cnt, ns.__array_interface__["data"][0],
                                              kk.__array_interface__["data"][0], out.__array_interface__["data"][0],
                                              self._scratch_ptr)
            if used >= 0:
                self.pos += int(used)
                break
            if used != -1:
                raise ValueError("hgr_sample_replay_many rejected its arguments")
            m = 2 * max(m, avail)
        res, off = [], 0
        for p, k in zip(populations, kk.tolist()):
            res.append(p[out[off:off + k]])
            off += k
        return res

    def plan(self, cands, anchors, num_compare: int, n_nodes: int, ptrs=None):
        """The host-side plan of one OM step in ONE library call (`hgr_om_plan`): every candidate array longer than
        `num_compare` is sub-sampled exactly like `random.sample`, anchors are appended where the draw missed them, the
        union of the sets is formed in ascending node id and the sets are rewritten as positions in it.  Returns
        ``(set_ptr [T+1], set_col [n_col], label_pos [T], union [n_union])`` int32 arrays, or None without the library."""
        import numpy as np
        lib = _host_lib()
        if lib is None:
            return None
        T = len(cands)
        ns = np.fromiter((c.shape[0] for c in cands), dtype=np.int64, count=T)
        if ptrs is None:
            ptrs = [c.__array_interface__["data"][0] for c in cands]
        ptrs = np.asarray(ptrs, dtype=np.uint64)
        anc = np.asarray(anchors, dtype=np.int64)
        cap = T * (num_compare + 1) + 1
        set_ptr = np.empty(T + 1, dtype=np.int32)
        set_col = np.empty(cap, dtype=np.int32)
        label = np.empty(max(T, 1), dtype=np.int32)
        union = np.empty(cap, dtype=np.int32)
        counts = np.zeros(2, dtype=np.int64)
        need = n_nodes + (int(ns.max()) if T else 0) + num_compare + 8
        if self._scratch is None or self._scratch.shape[0] < need:
            self._scratch = np.empty(max(need, 4096), dtype=np.int32)
            self._scratch_ptr = self._scratch.__array_interface__["data"][0]
        draws = int((ns > num_compare).sum())
        m = int(1.6 * draws * num_compare) + 32 * draws + 8
        addr = lambda a: a.__array_interface__["data"][0]
        while True:
            self.ensure(m)
            avail = self.words.shape[0] - self.pos
            used = lib.hgr_om_plan(self._words_ptr + 4 * self.pos, avail, T, addr(ptrs), addr(ns), addr(anc), num_compare,
                                   n_nodes, addr(set_ptr), addr(set_col), addr(label), addr(union), addr(counts),
                                   self._scratch_ptr)
            if used >= 0:
                self.pos += int(used)
                break
            if used != -1:
                raise ValueError("hgr_om_plan rejected its arguments")
            m = 2 * max(m, avail)
        return set_ptr, set_col[:int(counts[0])], label[:T], union[:int(counts[1])]

    def sample_array(self, population, k: int):
        """`random.sample(population, k)` for a numpy `population`, as a numpy array (same elements, same order)."""
        return population[self.positions(population.shape[0], k)]

    # -- random.sample(population, k) on the stream
    def sample(self, population, k: int):
        import numpy as np
        n = len(population)
        if not 0 <= k <= n:
            raise ValueError("Sample larger than population or is negative")
        if _host_lib() is not None:
            pos = self.positions(n, k).tolist()
            return [population[j] for j in pos]
        return self._sample_py(population, k)

    # pure-Python restatement (no library): what `_fast_sample_ok` pins the C helper and itself against
    def _sample_py(self, population, k: int):
        import numpy as np
        n = len(population)
        if k == 0:
            return []          # random.sample(population, 0) draws nothing
        if n <= _sample_setsize(k):
            # pool algorithm: j = randbelow(n - i); result[i] = pool[j]; pool[j] = pool[n - i - 1]
            m = 2 * k + 64
            while True:
                self.ensure(m)
                words = self.words[self.pos:self.pos + m].tolist()
                pool = list(population)
                result = [None] * k
                p = 0
                try:
                    for i in range(k):
                        left = n - i
                        sh = 32 - left.bit_length()
                        r = words[p] >> sh
                        p += 1
                        while r >= left:
                            r = words[p] >> sh
                            p += 1
                        result[i] = pool[r]
                        pool[r] = pool[left - 1]
                except IndexError:        # more rejections than the window covers: look further ahead
                    m *= 2
                    continue
                self.pos += p
                return result
        # set algorithm: j = randbelow(n), redrawn while already selected; result[i] = population[j]
        sh = 32 - n.bit_length()
        m = int(k * (1 << n.bit_length()) / n * 1.25) + 64
        while True:
            self.ensure(m)
            r = self.words[self.pos:self.pos + m] >> np.uint32(sh)
            pos = np.flatnonzero(r < n)                      # words that survive randbelow's rejection
            vals = r[pos]
            # first occurrence of every distinct index: scatter the positions in reverse, so the earliest one sticks
            order = np.arange(vals.shape[0])
            owner = np.empty(n, dtype=np.intp)
            owner[vals[::-1]] = order[::-1]
            first = np.flatnonzero(owner[vals] == order)
            if first.shape[0] >= k:
                first = first[:k]
                self.pos += int(pos[first[-1]]) + 1
                return [population[j] for j in vals[first].tolist()]
            m *= 2


_FAST_SAMPLE = None
_HOST_LIB = False        # False: not looked for yet; None: unavailable


def _host_lib():
    """libhgr_b200.so for its host helper `hgr_sample_replay` (None when it is not built -- the pure-Python restatement
    of the same algorithms runs instead; this is host-side integer logic, not a fallback of any device computation)."""
    global _HOST_LIB
    if _HOST_LIB is False:
        try:
            from . import _cabi
            _HOST_LIB = _cabi.load()
        except Exception:
            _HOST_LIB = None
    return _HOST_LIB


def _fast_sample_ok() -> bool:
    """Both replays (C helper, pure Python) against the running interpreter's `random.sample`, once per process."""
    global _FAST_SAMPLE, _HOST_LIB
    if _FAST_SAMPLE is None:
        shapes = ((300, 256), (1045, 256), (1046, 256), (5500, 256), (40, 7), (4097, 64), (21841, 256), (7, 7), (1, 1), (3, 0))

        def agrees(use_lib) -> bool:
            a, b = _random.Random(11), _random.Random(11)
            ok = True
            with SampleStream(b) as st:
                for (n, k) in shapes:
                    pop = list(range(1000, 1000 + n))
                    ok = ok and a.sample(pop, k) == (st.sample(pop, k) if use_lib else st._sample_py(pop, k))
            return ok and a.getstate() == b.getstate()
        try:
            if _host_lib() is not None and not agrees(True):
                _HOST_LIB = None          # never trust a helper that disagrees with the interpreter
            _FAST_SAMPLE = agrees(False) if _host_lib() is None else True
        except Exception:
            _FAST_SAMPLE = False
    return _FAST_SAMPLE


def sample_stream(rng):
    """A `SampleStream` over `rng` when the restatement matches this interpreter, else a pass-through whose `sample` is
    `rng.sample` -- either way: same lists, same generator state after the `with` block."""
    if hasattr(rng, "getstate") and hasattr(rng, "getrandbits") and _fast_sample_ok():
        return SampleStream(rng)
    return _PassThrough(rng)


class _PassThrough:
    def __init__(self, rng):
        self.rng = rng

    def __enter__(self):
        return self.rng

    def __exit__(self, *exc):
        return False


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
    base = contra_topk_candidates(d2n, depth, parents, k, cache)
    if len(base) > num_compare:
        compare_idx = rng.sample(base, num_compare)   # a new list; the cached one is never modified (rng may be a SampleStream)
    else:
        compare_idx = list(base)
    if target not in compare_idx:
        compare_idx.append(target)
    return compare_idx, compare_idx.index(target)


def contra_topk_candidates(d2n, depth: int, parents: Sequence[int], k: int, cache: Optional[dict] = None,
                           as_array: bool = False):
    """The list `random.sample` draws from in `contra_topk` (clip_tree.py:118-131): nodes of the depth window minus the
    anchor chain, in the iteration order of the reference's set difference.  ``as_array``: the same ids as a (cached)
    int64 numpy array, for `SampleStream.sample_arrays`."""
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
    if not as_array:
        return base
    import numpy as np
    akey = ("array",) + lkey[1:]
    arr = cache.get(akey) if cache is not None else None
    if arr is None:
        arr = np.asarray(base, dtype=np.int64)
        if cache is not None:
            order = cache.setdefault("_list_keys", [])
            order.append(akey)          # evicted with the lists, oldest first
            cache[akey] = arr
    return arr


def contra_topk_many(d2n, requests, k: int, num_compare: int, rng, cache: Optional[dict] = None):
    """`contra_topk` for a run of (target, depth, parents) requests -- the T iterations of one OM step -- with all the
    `random.sample` calls of the run replayed in ONE library call (`SampleStream.sample_arrays`): same draws, same
    order, same generator state.  Returns [(compare_idx int64 numpy array, label position)]."""
    import numpy as np
    bases = [contra_topk_candidates(d2n, depth, parents, k, cache, as_array=True) for (_, depth, parents) in requests]
    need = [i for i, b in enumerate(bases) if b.shape[0] > num_compare]
    if hasattr(rng, "sample_arrays"):
        drawn = rng.sample_arrays([bases[i] for i in need], [num_compare] * len(need)) if need else []
    else:       # plain `random` (the restatement did not match this interpreter): call by call
        drawn = [np.asarray(rng.sample(bases[i].tolist(), num_compare), dtype=np.int64) for i in need]
    for i, d in zip(need, drawn):
        bases[i] = d
    out = []
    for (target, _, _), ids in zip(requests, bases):
        hit = np.flatnonzero(ids == target)
        if hit.shape[0] == 0:
            ids = np.append(ids, target)
            pos = ids.shape[0] - 1
        else:
            pos = int(hit[0])
        out.append((ids, pos))
    return out


def om_plan(d2n, requests, k: int, num_compare: int, rng, n_nodes: int, cache: Optional[dict] = None):
    """`contra_topk_many` + the union / inverse of `train_batch` in one call of the library's host helper
    (`SampleStream.plan`); None when `rng` is not a SampleStream over the library (the caller then takes the numpy path)."""
    if not hasattr(rng, "plan"):
        return None
    cands = [contra_topk_candidates(d2n, depth, parents, k, cache, as_array=True) for (_, depth, parents) in requests]
    ptrs = None
    if cache is not None:      # the candidate arrays are cached objects: so are their addresses (`__array_interface__` is slow)
        known = cache.setdefault("_ptrs", {})
        if len(known) > 4 * _LIST_CACHE_ENTRIES:
            known.clear()
        ptrs = []
        for c in cands:
            e = known.get(id(c))
            if e is None or e[0] is not c:
                e = known[id(c)] = (c, c.__array_interface__["data"][0])
            ptrs.append(e[1])
    return rng.plan(cands, [t for (t, _, _) in requests], num_compare, n_nodes, ptrs)


def contra_random(train_ids: Sequence[int], target: int, num_compare: int, rng=_random) -> Tuple[List[int], int]:
    """model/clip_tree.py:81-89."""
    compare_idx = rng.sample(list(train_ids), num_compare)          # rng: `random` or a SampleStream over it
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
