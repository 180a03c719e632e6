"""Development check of the floor-sketch epilogue: parity against fp32 logits on ordinary and hostile banks, then timing."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from hgrnet_b200 import _cabi, ops
from sweep import emb, timeit
from tests.util import compare_topk

torch.backends.cuda.matmul.allow_tf32 = False
SK = _cabi.HGR_IMPL_TCGEN05_SKETCH
NM = _cabi.HGR_IMPL_FLAG_NO_MERGE


def banks(C, D, kind, seed=3):
    g = torch.Generator().manual_seed(seed)
    if kind == "iid":
        return emb(C, D, seed)
    if kind == "clustered":      # siblings adjacent and similar: child = parent + noise, rows in tree order
        centers = torch.randn((C + 63) // 64, D, generator=g)
        w = centers.repeat_interleave(64, 0)[:C] + 0.15 * torch.randn(C, D, generator=g)
        return (w / w.norm(dim=-1, keepdim=True)).to(torch.bfloat16)
    if kind == "ascending":      # every row's logits rise along the bank: each column beats all before it
        base = torch.randn(1, D, generator=g)
        t = torch.linspace(0.0, 1.0, C).unsqueeze(1)
        w = t * base + (1 - t) * torch.randn(1, D, generator=g) + 0.01 * torch.randn(C, D, generator=g)
        return (w / w.norm(dim=-1, keepdim=True)).to(torch.bfloat16)
    if kind == "equal":          # all bank rows identical: every logit of a row ties
        w = torch.randn(1, D, generator=g).repeat(C, 1)
        return (w / w.norm(dim=-1, keepdim=True)).to(torch.bfloat16)
    if kind == "dups":           # 40 copies of every distinct row: ties at the cut
        w = emb((C + 39) // 40, D, seed).repeat_interleave(40, 0)[:C]
        return w.contiguous()
    if kind == "zeros":          # mostly zero rows (padding) and a few real ones
        w = torch.zeros(C, D)
        w[::97] = torch.randn(len(range(0, C, 97)), D, generator=g)
        n = w.norm(dim=-1, keepdim=True)
        return (w / n.clamp_min(1e-12)).to(torch.bfloat16)
    raise ValueError(kind)


def check(B, C, D, kind):
    w = banks(C, D, kind).cuda()
    if kind == "ascending":
        x = (w[-1:].float() + 0.02 * torch.randn(B, D, device="cuda"))
        x = (x / x.norm(dim=-1, keepdim=True)).to(torch.bfloat16)
    elif kind == "clustered":
        pick = torch.randint(0, C, (B,), device="cuda")
        x = w[pick].float() + 0.3 * torch.randn(B, D, device="cuda") / D ** 0.5
        x = (x / x.norm(dim=-1, keepdim=True)).to(torch.bfloat16)
    else:
        x = emb(B, D, 11).cuda()
    val, idx = ops.score_topk(x, w, K=20, impl=SK)
    torch.cuda.synchronize()
    sel = ops.last_rescan_count()
    ref_v, ref_i = ops.score_topk(x, w, K=20, impl=ops.HGR_IMPL_SIMT)
    logits = (x.float() @ w.float().T).cpu()
    ties = compare_topk(val, idx, logits, list(range(C)), 20)
    same = bool((idx == ref_i).all())
    print("ok   B=%d C=%d D=%d %-10s selections %d  tie-rule rows %d  ids identical to the CUDA-core kernel: %s" % (
        B, C, D, kind, sel, ties, same), flush=True)


if __name__ == "__main__":
    if "time" not in sys.argv:
        for (B, C, D) in [(64, 1000, 1024), (512, 21841, 1024), (1024, 10450, 512), (4096, 2731, 1024), (300, 5000, 512),
                          (1, 17, 64), (130, 700, 256), (257, 33, 128), (4096, 21841, 1024)]:
            kinds = ["iid", "clustered", "ascending", "equal", "dups", "zeros"] if B * C < 3e7 else ["iid", "clustered"]
            for kind in kinds:
                check(B, C, D, kind)
    for (B, C, D) in [(512, 21841, 1024), (4096, 21841, 1024), (1024, 10450, 512), (4096, 2731, 1024), (4096, 5461, 1024)][:int(os.environ.get("HGR_EXP_SHAPES", "5"))]:
        nb = min(6, max(2, int(1.6 * 126e6 / (C * D * 2)) + 1))
        flops = 2.0 * B * C * D
        for kind in ("iid", "clustered"):
            bk = [banks(C, D, kind, 2 + i).cuda() for i in range(nb)]
            xs = [emb(B, D, 10 + i).cuda() for i in range(4)]
            res = []
            for name, impl in (("sketch", SK | NM), ("sketch+merge", SK), ("defer", ops.HGR_IMPL_TCGEN05 | NM),
                               ("defer+merge", ops.HGR_IMPL_TCGEN05), ("null", ops.HGR_IMPL_TCGEN05_NULL)):
                us = timeit(lambda i: ops.score_topk(xs[i % 4], bk[i % nb], K=20, impl=impl))
                res.append("%s %.2f us (%.3f)" % (name, us, flops / us / 1e6 / 1658.5))
            print("[time] B=%d C=%d D=%d %-9s %s" % (B, C, D, kind, " | ".join(res)), flush=True)
            del bk
