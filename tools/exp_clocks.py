"""SM clock and board power under back-to-back launches of the scoring kernel: production epilogue vs null epilogue
(the main loop alone).  Is the MMA slow-down next to a busy epilogue a clock effect of the power cap?"""
import os
import subprocess
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from hgrnet_b200 import _cabi, ops
from sweep import emb

NM = _cabi.HGR_IMPL_FLAG_NO_MERGE


def sample(fn, seconds=2.0):
    rows = []
    proc = subprocess.Popen(["nvidia-smi", "-i", "0", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits",
                             "-lms", "100"], stdout=subprocess.PIPE, text=True)
    th = threading.Thread(target=lambda: [rows.append(l) for l in proc.stdout], daemon=True)
    th.start()
    g = torch.cuda.CUDAGraph()
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        fn(0)
        st.synchronize()
        with torch.cuda.graph(g, stream=st):
            for i in range(20):
                fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    n = 0
    e0.record()
    while time.time() - t0 < seconds:
        for _ in range(50):
            g.replay()
        n += 50
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / (n * 20) * 1e3
    time.sleep(0.15)
    proc.terminate()
    vals = [[float(x) for x in r.split(",")] for r in rows[5:] if r.strip()]
    vals.sort()
    med = vals[len(vals) // 2] if vals else [0, 0]
    return us, med[0], sorted(v[1] for v in vals)[len(vals) // 2] if vals else 0


for (B, C, D) in ((512, 21841, 1024), (4096, 2731, 1024)):
    banks = [emb(C, D, 2 + i).cuda() for i in range(6)]
    xs = [emb(B, D, 10 + i).cuda() for i in range(4)]
    for name, impl in (("null epilogue", ops.HGR_IMPL_TCGEN05_NULL), ("production lists", ops.HGR_IMPL_TCGEN05 | NM),
                       ("exact lists", ops.HGR_IMPL_TCGEN05_EXACT | NM)):
        us, mhz, watts = sample(lambda i: ops.score_topk(xs[i % 4], banks[i % 6], K=20, impl=impl))
        print("B=%d C=%d %-17s %.2f us per launch, median SM clock %.0f MHz, median board power %.0f W" % (B, C, name, us, mhz, watts), flush=True)
