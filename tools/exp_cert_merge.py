"""Producer side of the global certificate at B = 4096: scoring kernel alone vs scoring + local merge + scatter (to local
buffers), at the N = 8 / 4 / 2 shard sizes -- what the merge of the narrow lists costs (library in HGR_LIB)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from hgrnet_b200 import _cabi, ops
from sweep import emb, timeit

NM = _cabi.HGR_IMPL_FLAG_NO_MERGE
B, D, K, CT = 4096, 1024, 20, 21841
tag = os.path.basename(os.environ.get("HGR_LIB", "stock"))
for G in (8, 4, 2):
    C = (CT + G - 1) // G
    banks = [emb(C, D, 2 + i).cuda() for i in range(6)]
    xs = [emb(B, D, 10 + i).cuda() for i in range(4)]
    blk = B // G
    val = [torch.empty(blk, K, device="cuda") for _ in range(G)]
    idx = [torch.empty(blk, K, dtype=torch.int32, device="cuda") for _ in range(G)]
    bnd = [torch.empty(blk, device="cuda") for _ in range(G)]
    vp, ip, bp = [t.data_ptr() for t in val], [t.data_ptr() for t in idx], [t.data_ptr() for t in bnd]
    t0 = timeit(lambda i: ops.score_topk_scatter(xs[i % 4], banks[i % 6], vp, ip, blk, K=K, impl=ops.HGR_IMPL_TCGEN05 | NM,
                                                 bound_block_ptrs=bp, C_total=CT))
    t1 = timeit(lambda i: ops.score_topk_scatter(xs[i % 4], banks[i % 6], vp, ip, blk, K=K, bound_block_ptrs=bp, C_total=CT))
    print("%s N=%d shard C=%d list length %d: lists only %.2f us, + merge / scatter %.2f us (merge %.2f)" % (
        tag, G, C, ops.global_list_len(B, C, D, K, CT), t0, t1, t1 - t0), flush=True)
