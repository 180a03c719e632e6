"""Per-rank cost of the class-sharded pipeline on ONE GPU: a world-size-1 ShardedEvalStream over a bank shard of
the size a rank holds at N = 2 / 4 / 8 (B = 4096, D = 1024), next to its kernels timed alone."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from hgrnet_b200 import _cabi, ops
from hgrnet_b200.dist import ShardedEvalStream
from sweep import emb, timeit

B, D, K = 4096, 1024, 20
for C in (10921, 5461, 2731):
    nb = 6
    banks = [emb(C, D, 2 + i).cuda() for i in range(nb)]
    feats = [torch.randn(B, D, device="cuda") for _ in range(4)]
    xs = [ops.normalize_rows(f) for f in feats]
    for exchange, ch in (("p2p", 1), ("p2p", 2), ("p2p", 3), ("p2p", 4), ("nccl", 1)):
        ses = ShardedEvalStream(banks[0], 0, batch=B, K=K, steps=8, banks=banks, exchange=exchange, channels=ch)
        for s in range(8):
            ses.dev_feats[s].copy_(feats[s % 4])
        for _ in range(3):
            ses.run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            ses.run()
        e1.record()
        torch.cuda.synchronize()
        print("C=%d %s x%d pipeline: %.2f us/step" % (C, exchange, ch, e0.elapsed_time(e1) / 160 * 1e3), flush=True)
    NM = _cabi.HGR_IMPL_FLAG_NO_MERGE
    print("  normalize      %.2f us" % timeit(lambda i: ops.normalize_rows(feats[i % 4])))
    print("  score no-merge %.2f us" % timeit(lambda i: ops.score_topk(xs[i % 4], banks[i % nb], K=K, impl=ops.HGR_IMPL_TCGEN05 | NM)))
    print("  score + merge  %.2f us" % timeit(lambda i: ops.score_topk(xs[i % 4], banks[i % nb], K=K)))
    print("  null main loop %.2f us" % timeit(lambda i: ops.score_topk(xs[i % 4], banks[i % nb], K=K, impl=ops.HGR_IMPL_TCGEN05_NULL)))
    pv = torch.randn(8, 512, K, device="cuda").sort(dim=-1, descending=True).values.contiguous()
    pi = torch.randint(0, 21841, (8, 512, K), device="cuda", dtype=torch.int32)
    print("  owner merge 8 lists x 512 rows %.2f us" % timeit(lambda i: ops.topk_merge(pv, pi)))
    del banks, ses
