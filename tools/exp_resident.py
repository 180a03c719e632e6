"""Experiment driver for the resident-A kernel: times the null-epilogue main loop and the production kernel for a
few shapes under the environment knobs given on the command line (each configuration needs its own process:
the library reads the knobs once)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from hgrnet_b200 import _cabi, ops
from sweep import emb, timeit

NM = _cabi.HGR_IMPL_FLAG_NO_MERGE
shapes = [(512, 21841, 1024), (1024, 10450, 512), (4096, 2731, 1024)]
if len(sys.argv) > 1:
    shapes = [tuple(int(v) for v in a.split(",")) for a in sys.argv[1:]]
tag = " ".join("%s=%s" % (k, v) for k, v in sorted(os.environ.items()) if k.startswith("HGR_"))
for (B, C, D) in shapes:
    nb = min(6, max(2, int(1.6 * 126e6 / (C * D * 2)) + 1))
    banks = [emb(C, D, 2 + i).cuda() for i in range(nb)]
    xs = [emb(B, D, 10 + i).cuda() for i in range(4)]
    null = timeit(lambda i: ops.score_topk(xs[i % 4], banks[i % nb], K=20, impl=ops.HGR_IMPL_TCGEN05_NULL))
    prod = timeit(lambda i: ops.score_topk(xs[i % 4], banks[i % nb], K=20, impl=ops.HGR_IMPL_TCGEN05 | NM))
    flops = 2.0 * B * C * D
    snull = timeit(lambda i: ops.score_topk(xs[i % 4], banks[i % nb], K=20, impl=_cabi.HGR_IMPL_TCGEN05_STREAM_NULL))
    sprod = timeit(lambda i: ops.score_topk(xs[i % 4], banks[i % nb], K=20, impl=_cabi.HGR_IMPL_TCGEN05_STREAM | NM))
    print("[%s] B=%d C=%d D=%d  null %.2f us (%.3f)  prod %.2f us (%.3f) | stream null %.2f prod %.2f" % (
        tag, B, C, D, null, flops / null / 1e6 / 1658.5, prod, flops / prod / 1e6 / 1658.5, snull, sprod), flush=True)
    del banks
    if os.environ.get("HGR_TIMELINE"):
        for w_ in ops._workspaces.values():
            tl = w_[64:64 + 256 * 32 * 8].view(torch.int64).reshape(256, 32)[:148].cpu()
            if (tl[:, 13] > 0).any():
                life_ns = (tl[:, 13] - tl[:, 0]).float()
                cyc = tl[:, 16:21].sum(1).float()
                ok = cyc > 0
                print("   in-kernel clock: lifetime %.2f us, epilogue-warp cycles %.0f -> %.3f GHz" % (
                    life_ns[ok].median() / 1e3, cyc[ok].median(), (cyc[ok] / life_ns[ok]).median()), flush=True)
