"""Kernel (2) alone at B = 4096 over the bank sizes a rank holds at N = 8 / 4 / 2 / 1: production lists, exact
20-entry lists, with / without the merge, and the null-epilogue main loop (CUDA-graph timed, rotating banks)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from hgrnet_b200 import _cabi, ops
from sweep import emb, timeit

NM = _cabi.HGR_IMPL_FLAG_NO_MERGE
for (B, C) in ((4096, 2731), (4096, 5461), (4096, 10921), (4096, 21841)):
    nb = 6 if C < 20000 else 5
    banks = [emb(C, 1024, 2 + i).cuda() for i in range(nb)]
    xs = [emb(B, 1024, 10 + i).cuda() for i in range(4)]
    print(B, C, "prod %.2f  +merge %.2f  exact %.2f  +merge %.2f  null %.2f" % (
        timeit(lambda i: ops.score_topk(xs[i % 4], banks[i % nb], K=20, impl=ops.HGR_IMPL_TCGEN05 | NM)),
        timeit(lambda i: ops.score_topk(xs[i % 4], banks[i % nb], K=20)),
        timeit(lambda i: ops.score_topk(xs[i % 4], banks[i % nb], K=20, impl=ops.HGR_IMPL_TCGEN05_EXACT | NM)),
        timeit(lambda i: ops.score_topk(xs[i % 4], banks[i % nb], K=20, impl=ops.HGR_IMPL_TCGEN05_EXACT)),
        timeit(lambda i: ops.score_topk(xs[i % 4], banks[i % nb], K=20, impl=ops.HGR_IMPL_TCGEN05_NULL))), flush=True)
    del banks
