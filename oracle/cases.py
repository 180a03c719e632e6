"""Seeded synthetic inputs of the golden cases.  TEST INFRASTRUCTURE (shared by
``oracle/gen_golden.py`` and ``tests/``): fixtures store only outputs, inputs are regenerated
here from the recorded seeds.  Embeddings are N(0,1), optionally row-normalised, and
bf16-VALUED fp32.
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import numpy as np
import torch

ROOT = "fall11"


def wnid(i: int) -> str:
    return "n%08d" % (i + 1)


def tree_edges(level_sizes: Sequence[int], seed: int) -> List[List[str]]:
    """Random tree, nodes numbered level by level, reference edge-list format (root 'fall11')."""
    rng = np.random.RandomState(seed)
    edges: List[List[str]] = []
    start, prev = 0, []
    for d, n in enumerate(level_sizes):
        ids = list(range(start, start + n))
        if d == 0:
            edges += [[ROOT, wnid(i)] for i in ids]
        else:
            par = [prev[j % len(prev)] for j in range(n)]
            rng.shuffle(par)
            edges += [[wnid(int(p)), wnid(i)] for p, i in zip(par, ids)]
        prev, start = ids, start + n
    return edges


# node order b, c, a, d, e, f -> depths 1, 2, 0, 1, 2, 3: d2n keys come out as [1, 2, 0, 3]
QUIRKY_EDGES = [[wnid(1), wnid(2)], [wnid(0), wnid(1)], [ROOT, wnid(0)], [wnid(0), wnid(3)], [wnid(3), wnid(4)],
                [wnid(4), wnid(5)], [ROOT, wnid(6)], [wnid(6), wnid(7)]]


def bf16_valued(t: torch.Tensor) -> torch.Tensor:
    return t.to(torch.bfloat16).float()


def _randn(n, d, seed):
    return torch.randn(n, d, generator=torch.Generator().manual_seed(seed))


def text_table(spec: Dict, n_nodes: int, normalize: bool = True) -> torch.Tensor:
    x = _randn(n_nodes, spec["D"], spec["text_seed"])
    if normalize:
        x = x / x.norm(dim=-1, keepdim=True)
    else:
        x = x * 0.05  # raw encoder-like scale; normalisation happens inside the head
    return bf16_valued(x)


def image_feats(spec: Dict, batch: int = 0) -> torch.Tensor:
    return bf16_valued(_randn(spec["B"], spec["D"], spec["img_seed"] + batch))


def test_ids(spec: Dict, n_nodes: int) -> List[int]:
    """Test classes: the deepest level (leaves), like the 'rest' split; or every node."""
    if spec.get("test_all"):
        return list(range(n_nodes))
    last = spec["levels"][-1]
    return list(range(n_nodes - last, n_nodes))


def eval_batches(spec: Dict, test_id_list: Sequence[int]):
    """Single-label batches.  Half of each batch is pulled towards its label's text embedding so that the
    Hit@k counters are non-trivial."""
    n_nodes = sum(spec["levels"])
    table = text_table(spec, n_nodes)
    rng = np.random.RandomState(spec["label_seed"])
    out = []
    for b in range(spec["batches"]):
        label = int(test_id_list[rng.randint(len(test_id_list))])
        f = _randn(spec["B"], spec["D"], spec["img_seed"] + b)
        f = f / f.norm(dim=-1, keepdim=True)
        mix = torch.linspace(0.0, spec.get("signal", 0.25), spec["B"])[:, None]
        f = bf16_valued(f + mix * table[label][None, :])
        out.append((f, label))
    return out


EVAL_CASES = [
    # cfg 1 of BASELINE.json: 3-level hierarchy 10/100/1000, leaves are the test set, batch 64, RN50 dim 1024
    dict(name="eval_cfg1", levels=[10, 100, 1000], tree_seed=11, D=1024, B=64, batches=3, text_seed=101,
         img_seed=201, label_seed=301),
    # deeper, ragged: 6 levels, every node is a test class, batch not a multiple of anything
    dict(name="eval_deep", levels=[3, 7, 19, 41, 83, 160], tree_seed=12, D=256, B=37, batches=4, text_seed=102,
         img_seed=202, label_seed=302, test_all=True, signal=0.5),
]

OM_CASES = [
    dict(name="om_3level_equal", levels=[4, 20, 200], tree_seed=3, D=64, B=16, text_seed=111, img_seed=211,
         target=4 + 20 + 57, sample_seed=5,
         opts=dict(weights="equal", out_ratio=0.25, in_ratio=0.5, k=1, num_compare=256, weighting="both")),
    dict(name="om_3level_adaptive", levels=[4, 20, 200], tree_seed=3, D=64, B=16, text_seed=111, img_seed=211,
         target=4 + 20 + 57, sample_seed=5,
         opts=dict(weights="adaptive", out_ratio=0.25, in_ratio=0.5, k=1, num_compare=256, weighting="both", scale=1.0)),
    # deep chain, sub-sampling active (level sizes > num_compare), k = 2 levels of negatives
    dict(name="om_deep_sampled", levels=[3, 9, 40, 120, 300, 500], tree_seed=4, D=128, B=24, text_seed=112,
         img_seed=212, target=3 + 9 + 40 + 120 + 300 + 123, sample_seed=6,
         opts=dict(weights="increasing", out_ratio=0.5, in_ratio=0.75, k=2, num_compare=64, weighting="out")),
]
