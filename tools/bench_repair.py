"""Cost of the exact repair path (speculative lists that cannot be certified), to calibrate the constant of the
speculation rule in csrc/score_launch.cu (`t_repair`, currently 0.6 us per bank row of 1024 elements).

An adversarial bank (a block of adjacent rows close to the mean image direction) overflows one speculative list of
many image rows at once; the script reports the number of repaired rows, the time of the fused call with and
without repairs, and the implied microseconds per re-scanned bank row.  Run on the GPU box:
    python tools/bench_repair.py [B C]          (default: cfg 2, 512 x 21841)
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from hgrnet_b200 import ops
from sweep import emb, timeit

B, C = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) >= 3 else (512, 21841)
D, K = 1024, 20
plan = ops.score_topk_plan(B, C, D, K)
print("plan:", plan)
if plan["list_len"] >= K:
    print("this shape runs exact lists: nothing to repair (set HGR_SPEC_COST=1 to force speculation)")
x = torch.randn(B, D, generator=torch.Generator().manual_seed(1))
w = emb(C, D, 2).float()
xn = ops.normalize_rows(x.cuda())
clean = w.cuda().bfloat16()
mean_dir = x.mean(0)
mean_dir = mean_dir / mean_dir.norm()
noise = emb(40, D, 3).float()
adv = w.clone()
r0 = C // 4 + 37
blk = mean_dir[None, :] * 3 + noise
adv[r0:r0 + 40] = (blk / blk.norm(dim=-1, keepdim=True)).to(torch.bfloat16).float()
adv = adv.cuda().bfloat16()

ops.score_topk(xn, clean, K=K)
torch.cuda.synchronize()
n0 = ops.last_rescan_count("cuda:0")
ops.score_topk(xn, adv, K=K)
torch.cuda.synchronize()
n1 = ops.last_rescan_count("cuda:0")
t_clean = timeit(lambda i: ops.score_topk(xn, clean, K=K), n=100)
t_adv = timeit(lambda i: ops.score_topk(xn, adv, K=K), n=100)
cols = plan["cols_per_worker"]
print("repaired rows: clean %d, adversarial %d" % (n0, n1))
print("fused call: clean %.1f us, adversarial %.1f us" % (t_clean, t_adv))
if n1 > 0:
    # repaired rows run concurrently (one warp each, 4 per CTA); the call waits for the slowest: one range of `cols` rows
    print("=> ~%.2f us per re-scanned bank row (range of %d rows)" % ((t_adv - t_clean) / cols, cols))
