"""A few calls of hgr_hier_metrics_fused at the cfg-2 shape (B = 512, all 21,841 nodes as train classes, 12 levels) for ncu."""
import sys

import torch

sys.path.insert(0, ".")
from hgrnet_b200 import ops
from hgrnet_b200.hierarchy import WORDNET_LIKE_21841, synthetic_hierarchy

h = synthetic_hierarchy(WORDNET_LIKE_21841, seed=1)
N, D, B = len(h), 1024, 512
g = torch.Generator().manual_seed(1)
bank = torch.randn(N, D, generator=g)
bank = (bank / bank.norm(dim=-1, keepdim=True)).to(torch.bfloat16).cuda()
x = ops.normalize_rows(torch.randn(B, D, generator=g).cuda())
depth = torch.from_numpy(h.depth).long()
n_levels = int(depth.max()) + 1
order = torch.sort(depth, stable=True).indices
bank_sorted = bank[order.cuda()].contiguous()
level_end = torch.cumsum(torch.bincount(depth, minlength=n_levels), 0).tolist()
first_out = torch.tensor([int((depth != l).nonzero()[0]) for l in range(n_levels)], dtype=torch.int32).cuda()
parents = list(h.c2p[N - 1]) + [N - 1]
chain = torch.tensor(parents, dtype=torch.int32).cuda()
chain_level = torch.tensor([len(h.c2p[p]) for p in parents], dtype=torch.int32).cuda()
counts = torch.zeros(3, dtype=torch.int64).cuda()
for _ in range(4):
    ops.hier_metrics_fused(x, bank_sorted, level_end, order.to(torch.int32).cuda(), first_out, chain, chain_level, counts)
torch.cuda.synchronize()
print("counts", counts.tolist())
