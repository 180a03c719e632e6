"""Scoring kernel (production lists) with / without its merge at the BASELINE shapes, library in HGR_LIB."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hgrnet_b200 import _cabi, ops
from sweep import emb, timeit

NM = _cabi.HGR_IMPL_FLAG_NO_MERGE
tag = os.path.basename(os.environ.get("HGR_LIB", "stock"))
for (B, C, D) in ((512, 21841, 1024), (1024, 10450, 512), (256, 21841, 1024), (4096, 21841, 1024)):
    nb = 6 if C < 20000 else 5
    banks = [emb(C, D, 2 + i).cuda() for i in range(nb)]
    xs = [emb(B, D, 10 + i).cuda() for i in range(4)]
    print(tag, B, C, D, "list length %d: kernel %.2f us  + merge %.2f us  null %.2f us" % (
        ops.score_topk_plan(B, C, D)["list_len"],
        timeit(lambda i: ops.score_topk(xs[i % 4], banks[i % nb], K=20, impl=ops.HGR_IMPL_TCGEN05 | NM)),
        timeit(lambda i: ops.score_topk(xs[i % 4], banks[i % nb], K=20)),
        timeit(lambda i: ops.score_topk(xs[i % 4], banks[i % nb], K=20, impl=ops.HGR_IMPL_TCGEN05_NULL))), flush=True)
    del banks
