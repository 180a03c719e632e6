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
