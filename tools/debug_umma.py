"""GPU-side diagnostic for the tcgen05 path: dense logits vs the CUDA-core kernel and torch, with an
error-pattern report (which rows / columns / K blocks are off) -- one run should localise a descriptor bug."""
import sys
import os
import json

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from hgrnet_b200 import ops


def emb(n, d, seed):
    x = torch.randn(n, d, generator=torch.Generator().manual_seed(seed))
    return (x / x.norm(dim=-1, keepdim=True)).to(torch.bfloat16)


def report(B, C, D):
    x, w = emb(B, D, 1).cuda(), emb(C, D, 2).cuda()
    ref = x.float() @ w.float().T
    out = {}
    for name, impl in (("simt", ops.HGR_IMPL_SIMT), ("tcgen05", ops.HGR_IMPL_TCGEN05)):
        try:
            got = ops.logits_dense(x, w, impl=impl)
            torch.cuda.synchronize()
        except Exception as e:  # noqa
            out[name] = "EXC %r" % (e,)
            continue
        err = (got - ref).abs()
        bad = err > 1e-3
        info = {"max_err": float(err.max()), "bad_frac": float(bad.float().mean())}
        if bad.any():
            rows = bad.any(1).nonzero().squeeze(1)
            cols = bad.any(0).nonzero().squeeze(1)
            info["bad_rows"] = rows[:16].tolist() + ["n=%d" % rows.numel()]
            info["bad_cols"] = cols[:16].tolist() + ["n=%d" % cols.numel()]
            info["sample_got"] = got[:2, :8].tolist()
            info["sample_ref"] = ref[:2, :8].tolist()
            # does the result look like a permutation of K chunks / rows?
            info["ratio_mean"] = float((got[bad] / ref[bad]).mean())
        out[name] = info
    return out


if __name__ == "__main__":
    res = {}
    for shape in ((128, 256, 64), (128, 256, 128), (128, 256, 1024), (128, 16, 64), (64, 80, 64), (256, 512, 256),
                  (512, 21841, 1024)):
        try:
            res[str(shape)] = report(*shape)
        except Exception as e:  # noqa
            res[str(shape)] = "EXC %r" % (e,)
            break
        print(shape, json.dumps(res[str(shape)])[:600], flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/debug_umma.json", "w"), indent=1)
