"""Synthetic stand-ins for the upstream producers (CLIP encoders, image loader).

There is no network, no CLIP checkpoint and no ImageNet-21K in this environment, and the
encoders / JPEG loaders are out of scope: benchmarks and tests feed the head
with seeded synthetic embeddings of the named shapes.
"""
from __future__ import annotations

import types
from typing import List, Sequence

import numpy as np
import torch
import torch.nn as nn


def bf16_valued(t: torch.Tensor) -> torch.Tensor:
    """Round to bf16 and return as fp32 (the precision rule of SURVEY.md section 8c)."""
    return t.to(torch.bfloat16).float()


def synthetic_embeddings(n: int, d: int, seed: int, device="cpu", normalize: bool = True) -> torch.Tensor:
    """N(0,1) per element, row-normalised, bf16-valued fp32."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, d, generator=g)
    if normalize:
        x = x / x.norm(dim=-1, keepdim=True)
    return bf16_valued(x).to(device)


def clustered_bank(n: int, d: int, seed: int, fan: int = 64, sigma: float = 0.15) -> torch.Tensor:
    """A class bank whose ROW ORDER follows a hierarchy: child = parent + sigma * noise, siblings adjacent -- like the
    reference's `nodes` (graph order, utils.py:44-45) with real CLIP text embeddings.  An image near a leaf then has
    its whole top-20 in a few adjacent bank rows.  Row-normalised, bf16-valued fp32, raw (unnormalised) when used as a
    text table is fine too."""
    g = torch.Generator().manual_seed(seed)
    centers = torch.randn((n + fan - 1) // fan, d, generator=g)
    w = centers.repeat_interleave(fan, 0)[:n] + sigma * torch.randn(n, d, generator=g)
    return bf16_valued(w / w.norm(dim=-1, keepdim=True))


def near_leaf_features(bank: torch.Tensor, b: int, seed: int, noise: float = 0.3) -> torch.Tensor:
    """Image features drawn near randomly chosen rows of `bank` (unnormalised fp32)."""
    g = torch.Generator().manual_seed(seed)
    pick = torch.randint(0, bank.shape[0], (b,), generator=g)
    return bank[pick].float() + noise * torch.randn(b, bank.shape[1], generator=g) / bank.shape[1] ** 0.5


class TableEncoder(nn.Module):
    """Duck-typed CLIP: ``encode_text`` gathers rows of a learnable table by node id (token column 0),
    ``encode_image`` passes pre-computed features through a unit gain.  ``logit_scale`` initialises to
    ln(1/0.07) like CLIP (clip/model.py:291)."""

    def __init__(self, text_table: torch.Tensor, log_scale: float = float(np.log(1 / 0.07))):
        super().__init__()
        self.text_table = nn.Parameter(text_table.clone().float())
        self.image_gain = nn.Parameter(torch.ones(()))
        self.logit_scale = nn.Parameter(torch.tensor(float(log_scale)))
        self.visual = types.SimpleNamespace(input_resolution=224)

    def encode_text(self, tokens):
        return self.text_table[tokens[:, 0]]

    def encode_image(self, x):
        return x * self.image_gain


def node_id_tokens(n: int, context_length: int = 77) -> torch.Tensor:
    toks = torch.zeros(n, context_length, dtype=torch.long)
    toks[:, 0] = torch.arange(n)
    return toks


class FeatureLoader:
    """Iterable of the reference's test-batch dicts ``{'img': [1,B,D], 'label': [1,B]}`` with ONE label per
    batch (dataset/imagenet_group_test.py:150-163), built from pinned host feature batches."""

    def __init__(self, feats: Sequence[torch.Tensor], labels: Sequence[int], pin: bool = False):
        self.feats = [f.pin_memory() if pin else f for f in feats]
        self.labels = list(labels)
        self.batch_sampler = types.SimpleNamespace(num_batch=len(self.feats))

    def __len__(self):
        return len(self.feats)

    def __iter__(self):
        for f, l in zip(self.feats, self.labels):
            yield {"img": f[None], "label": torch.full((1, f.shape[0]), int(l), dtype=torch.long)}
