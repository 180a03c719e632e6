"""Scoring kernel alone (no merge) at B = 4096 over the shard sizes of N = 8 / 4 / 2 with the list length of the library
in HGR_LIB (tools/build_variant.sh klN "-DHGR_FORCE_KL=N" score_launch.cu): what narrow lists would buy there."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hgrnet_b200 import _cabi, ops
from sweep import emb, timeit

NM = _cabi.HGR_IMPL_FLAG_NO_MERGE
tag = os.path.basename(os.environ.get("HGR_LIB", "stock"))
for (B, C) in ((4096, 2731), (4096, 5461), (4096, 10921), (512, 21841)):
    nb = 6 if C < 20000 else 5
    banks = [emb(C, 1024, 2 + i).cuda() for i in range(nb)]
    xs = [emb(B, 1024, 10 + i).cuda() for i in range(4)]
    print(tag, B, C, "lists-only %.2f us   exact %.2f us" % (
        timeit(lambda i: ops.score_topk(xs[i % 4], banks[i % nb], K=20, impl=ops.HGR_IMPL_TCGEN05 | NM)),
        timeit(lambda i: ops.score_topk(xs[i % 4], banks[i % nb], K=20, impl=ops.HGR_IMPL_TCGEN05_EXACT | NM))), flush=True)
    del banks
