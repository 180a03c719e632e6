"""Per-CTA timeline of the CTA-pair kernel from %globaltimer stamps (run with HGR_TIMELINE=1)."""
import os
import sys

os.environ["HGR_TIMELINE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from hgrnet_b200 import _cabi, ops


def emb(n, d, seed):
    x = torch.randn(n, d, generator=torch.Generator().manual_seed(seed))
    return (x / x.norm(dim=-1, keepdim=True)).to(torch.bfloat16)


names = ["entry", "setup", "first_full", "last_mma_issued", "acc0", "acc1", "acc2", "acc3", "epi0", "epi1", "epi2", "epi3",
         "epi_done", "exit"] + ["-"] * 18
NM = int(os.environ.get('HGR_TL_IMPL', ops.HGR_IMPL_TCGEN05)) | _cabi.HGR_IMPL_FLAG_NO_MERGE
CASES = ((512, 21841, 1024, NM, "prod"), (512, 21841, 1024, ops.HGR_IMPL_TCGEN05_NULL, "null"),
         (512, 2731, 1024, ops.HGR_IMPL_TCGEN05_NULL, "null-small"))
if len(sys.argv) >= 3:   # python tools/timeline.py B C  -> production and null epilogue at that shape
    b_, c_ = int(sys.argv[1]), int(sys.argv[2])
    d_ = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
    CASES = ((b_, c_, d_, NM, "prod"), (b_, c_, d_, ops.HGR_IMPL_TCGEN05_NULL, "null"))
for (B, C, D, impl, tag) in CASES:
    banks = [emb(C, D, 2).cuda() for _ in range(5)]
    x = emb(B, D, 1).cuda()
    for i in range(6):
        ops.score_topk(x, banks[i % 5], K=20, impl=impl)
    torch.cuda.synchronize()
    ws = ops._workspaces[("cuda", 0, torch.cuda.current_stream().cuda_stream)]
    print("   stats words:", ws[:64].view(torch.int32).tolist()[:12])
    tl = ws[64:64 + 256 * 32 * 8].view(torch.int64).reshape(256, 32)[:148].cpu()
    tl = tl[tl[:, 13] > 0]                                   # CTAs of this launch
    t0 = tl[:, 0].min()
    rel = (tl - t0).float() / 1e3
    print("== %s B=%d C=%d: all times in us relative to the first CTA entry" % (tag, B, C))
    for j, n in enumerate(names):
        if n == "-" or (tl[:, j] == 0).all():
            continue
        col = rel[:, j][tl[:, j] > 0]
        print("  %-16s min %7.2f  median %7.2f  max %7.2f" % (n, col.min(), col.median(), col.max()))
    cyc = tl[:, 16:21].float()
    for j, n in enumerate(["wait acc", "warm-up pass", "tmem ld", "scan", "drain"]):
        print("  cycles %-14s median %8.0f  (%.2f us at 1.9 GHz)" % (n, cyc[:, j].median(), cyc[:, j].median() / 1900))
    if (tl[:, 24] > 0).any():
        for j in range(3):
            print("  scan loop of sub-tile %d: median %8.0f cycles" % (j, tl[:, 24 + j].float().median()))
        for j, n in ((27, "select cycles"), (28, "compact cycles"), (29, "crowded chunks"), (30, "bisection passes"), (31, "selections")):
            print("  warp 2: %-18s median %8.0f  mean %8.1f" % (n, tl[:, j].float().median(), tl[:, j].float().mean()))
    d = rel[:, 13] - rel[:, 0]
    print("  CTA lifetime     min %7.2f  median %7.2f  max %7.2f" % (d.min(), d.median(), d.max()))
