"""GPU-side timing sweep of the scoring-head kernels (CUDA events, rotating inputs larger than L2).
Prints one line per (workload, variant): microseconds per call and the tensor-roofline fraction."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from hgrnet_b200 import _cabi, ops

PEAK = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))[
    "bf16_tflops"] if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) else 1590.0


def emb(n, d, seed):
    x = torch.randn(n, d, generator=torch.Generator().manual_seed(seed))
    return (x / x.norm(dim=-1, keepdim=True)).to(torch.bfloat16)


def timeit(fn, n=200, warm=10, variants=20):
    """Time `fn(i)` per call with the interpreter off the path: `variants` rotation indices are captured into one
    CUDA graph each (the library launches on the capture stream) and replayed."""
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    graphs = []
    with torch.cuda.stream(side):
        for i in range(variants):
            fn(i)
        side.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            for i in range(variants):
                fn(i)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    # clocks: the GPU idles at 120 MHz between measurements and needs tens of ms of load to reach its boost clock,
    # so warm up for >= 0.25 s and time >= 0.2 s of back-to-back replays
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    one = max(e0.elapsed_time(e1), 1e-3)           # ms per replay (cold clocks: an over-estimate)
    for _ in range(int(250.0 / one) + 1):
        g.replay()
    torch.cuda.synchronize()
    reps = max(n // variants, int(200.0 / one) + 1)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (reps * variants) * 1e3


def main():
    NM = _cabi.HGR_IMPL_FLAG_NO_MERGE
    variants = [("prod+merge", ops.HGR_IMPL_TCGEN05), ("prod", ops.HGR_IMPL_TCGEN05 | NM),
                ("exact", ops.HGR_IMPL_TCGEN05_EXACT | NM), ("null", ops.HGR_IMPL_TCGEN05_NULL),
                ("sketch", ops.HGR_IMPL_TCGEN05_SKETCH | NM), ("sketch+merge", ops.HGR_IMPL_TCGEN05_SKETCH)]
    out = []
    for (B, C, D) in ((512, 21841, 1024), (4096, 21841, 1024), (1024, 10450, 512), (4096, 2731, 1024), (512, 2731, 1024)):
        nb = max(2, int(1.6 * 126e6 / (C * D * 2)) + 1)
        banks = [emb(C, D, 2).cuda() for _ in range(min(nb, 6))]
        xs = [emb(B, D, 10 + i).cuda() for i in range(4)]
        xraw = [torch.randn(B, D, device="cuda") for i in range(4)]
        flops = 2.0 * B * C * D
        for name, impl in variants:
            us = timeit(lambda i: ops.score_topk(xs[i % 4], banks[i % len(banks)], K=20, impl=impl))
            out.append((B, C, D, name, us, flops / (us * 1e-6) / 1e12 / PEAK))
            print("B=%d C=%d D=%d %-10s %8.2f us  %.3f of bf16 peak" % out[-1], flush=True)
        us = timeit(lambda i: ops.normalize_rows(xraw[i % 4]))
        print("B=%d C=%d D=%d %-10s %8.2f us" % (B, C, D, "normalize", us), flush=True)
        dense_out = torch.empty((B, C), dtype=torch.float32, device="cuda")
        us = timeit(lambda i: ops.logits_dense(xs[i % 4], banks[i % len(banks)], out=dense_out), n=40)
        print("B=%d C=%d D=%d %-10s %8.2f us  %.3f of bf16 peak" % (B, C, D, "dense", us, flops / (us * 1e-6) / 1e12 / PEAK),
              flush=True)
        del banks
    # python-side overhead of one call (no GPU work to wait for): tiny problem
    x, w = emb(8, 64, 1).cuda(), emb(64, 64, 2).cuda()
    print("graphed tiny problem (simt): %.2f us/call" % timeit(lambda i: ops.score_topk(x, w, K=20, impl=ops.HGR_IMPL_SIMT)))


if __name__ == "__main__":
    main()
