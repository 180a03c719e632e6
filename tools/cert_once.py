"""One process, G logical ranks: two batches of the class-sharded head in global-certificate mode (narrow lists + bounds
scattered to the row owners, owners certify) -- the launch sequence tools/ncu_capture_r02b.sh profiles.
usage: cert_once.py B C_total D G [lists-only]"""
import sys

import torch

sys.path.insert(0, ".")
from hgrnet_b200 import _cabi, ops
from hgrnet_b200.dist import PeerExchange, exchange_layout, shard_bounds

B, C, D, G = (int(v) for v in sys.argv[1:5])
lists_only = len(sys.argv) > 5
K = 20
g = torch.Generator().manual_seed(1)


def emb(n):
    x = torch.randn(n, D, generator=g)
    return (x / x.norm(dim=-1, keepdim=True)).to(torch.bfloat16).cuda()


dev = torch.device("cuda", 0)
x = emb(B)
bounds = shard_bounds(C, G)
if lists_only:     # the scoring kernel of rank 0's shard alone, list length of the global certificate
    banks = [emb(bounds[0][1]) for _ in range(3)]
    dummy = [x.data_ptr()] * G
    for i in range(6):
        ops.score_topk_scatter(x, banks[i % 3], dummy, dummy, (B + G - 1) // G, K=K,
                               impl=ops.HGR_IMPL_TCGEN05 | _cabi.HGR_IMPL_FLAG_NO_MERGE, bound_block_ptrs=dummy, C_total=C)
    torch.cuda.synchronize()
    print("list length", ops.global_list_len(B, bounds[0][1], D, K, C))
    sys.exit(0)
w = emb(C)
shards = [w[lo:hi].contiguous() for lo, hi in bounds]
lay = exchange_layout(B, K, G, 4)
bufs = [ops.peer_alloc(lay["total"])[0] for _ in range(G)]
ranks = [PeerExchange(B, K, dev, slots=4, _bases=bufs, _rank=r, _world=G) for r in range(G)]
table = ops.shard_table([(s.data_ptr(), 0, s.shape[0], lo) for s, (lo, _) in zip(shards, bounds)], dev)
repairs = torch.zeros(1, dtype=torch.int32, device=dev)
targets = torch.randint(0, C, (B,), generator=g).int().cuda()
hits = ops.new_hits(dev)
for rep in range(2):
    for r, px in enumerate(ranks):
        px.scatter(x, shards[r], bounds[r][0], rep, C_total=C)
    for px in ranks:
        px.merge(rep, targets, hits, certify=(x, table, repairs))
torch.cuda.synchronize()
print("list length", ops.global_list_len(B, shards[0].shape[0], D, K, C), "repaired rows", int(repairs.item()), "hits", hits.tolist())
