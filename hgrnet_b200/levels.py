"""Level-weight vectors of the OM loss (host side; tiny).

Mirrors ``tree_model.get_weights`` (model/clip_tree.py:198-219) and the ``layer_weight``
initialisation (:70-74).  The vectors have at most ~13 entries, are needed on the HOST to
build the per-iteration weights of the fused cross-entropy kernel, and are therefore computed
with torch-CPU ops in exactly the reference's formulas.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn.functional as F

METHODS = ("equal", "increasing", "decreasing", "adaptive", "nl_increasing", "nl_decreasing")


def layer_weight_init(d2n, scale: float) -> torch.Tensor:
    """``1 / |level|`` in ``d2n`` insertion order, times ``--scale`` (model/clip_tree.py:72-74)."""
    num_layer = [len(d2n[layer]) for layer in d2n.keys()]
    return (1.0 / torch.tensor(num_layer)) * scale


def level_weights(method: str, n: int, layer_weight: Optional[torch.Tensor] = None) -> torch.Tensor:
    """fp32 weight vector of length ``n`` on the device of ``layer_weight`` (CPU if None)."""
    if method == "equal":
        return torch.ones(n) / n
    if method == "decreasing":
        w = torch.arange(start=n, end=0, step=-1)
        return w / w.sum()
    if method == "increasing":
        w = torch.arange(start=1, end=n + 1)
        return w / w.sum()
    if method == "adaptive":
        if layer_weight is None:
            raise ValueError("adaptive weights need layer_weight")
        return F.softmax(100 ** layer_weight[:n], dim=0)
    if method == "nl_increasing":
        w = torch.arange(start=1, end=n + 1) ** 3
        return w / w.sum()
    if method == "nl_decreasing":
        w = torch.arange(start=n, end=0, step=-1) ** 3
        return w / w.sum()
    raise ValueError("unknown --weights %r (expected one of %s)" % (method, ", ".join(METHODS)))


# ---- numpy twin for the training step's host side ---------------------------------------------------------------------
# `train_batch` needs the T per-iteration weights (products of level weights, clip_tree.py:265-273) and, when the
# adaptive `layer_weight` trains, d loss / d layer_weight.  A handful of 13-element torch ops plus a torch-autograd pass
# cost ~0.25 ms of host time per step; the same formulas in numpy, with the softmax Jacobian written out, ~40 us.
# Pinned against `level_weights` + autograd in the CPU tests.
def level_weights_np(method: str, n: int, layer_weight=None):
    """`level_weights` in numpy float32 (same formulas, model/clip_tree.py:198-219)."""
    import numpy as np
    if method == "equal":
        return np.full(n, 1.0 / n, dtype=np.float32)
    if method in ("decreasing", "increasing", "nl_decreasing", "nl_increasing"):
        w = np.arange(n, 0, -1, dtype=np.float64) if method.endswith("decreasing") else np.arange(1, n + 1, dtype=np.float64)
        if method.startswith("nl_"):
            w = w ** 3
        return (w / w.sum()).astype(np.float32)
    if method == "adaptive":
        if layer_weight is None:
            raise ValueError("adaptive weights need layer_weight")
        z = np.power(np.float32(100.0), np.asarray(layer_weight[:n], dtype=np.float32))
        e = np.exp(z - z.max())
        return (e / e.sum()).astype(np.float32)
    raise ValueError("unknown --weights %r (expected one of %s)" % (method, ", ".join(METHODS)))


def iteration_weights_np(recipes, layer_weight=None):
    """w_t = product over the recipe's (method, n, position) factors -> ``(w [T] float32, context for the gradient)``."""
    import numpy as np
    vecs = {}
    for rec in recipes:
        for (method, n, _) in rec:
            if (method, n) not in vecs:
                vecs[(method, n)] = level_weights_np(method, n, layer_weight)
    w = np.ones(len(recipes), dtype=np.float32)
    for t, rec in enumerate(recipes):
        for (method, n, pos) in rec:
            w[t] *= vecs[(method, n)][pos]
    return w, (recipes, vecs)


def iteration_weights_grad_np(ctx, coeff, layer_weight):
    """d (sum_t coeff_t * w_t) / d layer_weight for the adaptive vectors (softmax(100 ** lw[:n])): the softmax Jacobian
    and d 100**x / dx = ln(100) * 100**x written out.  ``coeff`` [T] float; returns float32 [len(layer_weight)]."""
    import numpy as np
    recipes, vecs = ctx
    lw = np.asarray(layer_weight, dtype=np.float32)
    g_vec = {k: np.zeros(v.shape[0], dtype=np.float64) for k, v in vecs.items() if k[0] == "adaptive"}
    for t, rec in enumerate(recipes):
        for f, (method, n, pos) in enumerate(rec):
            if method != "adaptive":
                continue
            other = 1.0
            for f2, (m2, n2, p2) in enumerate(rec):
                if f2 != f:
                    other *= float(vecs[(m2, n2)][p2])
            g_vec[(method, n)][pos] += float(coeff[t]) * other
    g = np.zeros(lw.shape[0], dtype=np.float64)
    for (method, n), gv in g_vec.items():
        # the softmax over 100 ** lw is sharply peaked (its largest entry is 1 - 1e-8 and less): the Jacobian is taken on
        # a float64 softmax, not on the rounded float32 weights -- 1 - v would be all rounding error there
        z = np.power(100.0, lw[:n].astype(np.float64))
        e = np.exp(z - z.max())
        v = e / e.sum()
        g_z = v * (gv - float(np.dot(gv, v)))
        g[:n] += g_z * np.log(100.0) * z
    return g.astype(np.float32)
