mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_configs.py tests/test_gpu_peer.py -q -m gpu -x 2>&1 | tail -6
cat > /tmp/ab.py <<'PY'
import os, sys
sys.path.insert(0, 'tools'); sys.path.insert(0, '.')
import torch
from hgrnet_b200 import ops, _cabi
from sweep import emb, timeit
NM = _cabi.HGR_IMPL_FLAG_NO_MERGE
tag = "no-gf" if os.environ.get("HGR_NO_GLOBAL_FLOOR") else "gf"
for (B, C, D) in [(512, 21841, 1024), (4096, 21841, 1024), (1024, 10450, 512), (4096, 2731, 1024), (4096, 5461, 1024), (4096, 10921, 1024)]:
    nb = min(6, max(2, int(1.6 * 126e6 / (C * D * 2)) + 1))
    banks = [emb(C, D, 2 + i).cuda() for i in range(nb)]
    xs = [emb(B, D, 10 + i).cuda() for i in range(4)]
    res = []
    for name, impl in (("prod", ops.HGR_IMPL_TCGEN05 | NM), ("prod+merge", ops.HGR_IMPL_TCGEN05), ("exact", ops.HGR_IMPL_TCGEN05_EXACT | NM), ("null", ops.HGR_IMPL_TCGEN05_NULL)):
        us = timeit(lambda i: ops.score_topk(xs[i % 4], banks[i % nb], K=20, impl=impl))
        res.append("%s %.2f" % (name, us))
    print("[%s] B=%d C=%d D=%d  %s" % (tag, B, C, D, " | ".join(res)), flush=True)
PY
timeout 400 python /tmp/ab.py > gpurun_out/r2j_ab.log 2>&1
HGR_NO_GLOBAL_FLOOR=1 timeout 400 python /tmp/ab.py >> gpurun_out/r2j_ab.log 2>&1
grep "^\[" gpurun_out/r2j_ab.log | sort -k2,4 ; grep -v "^\[" gpurun_out/r2j_ab.log | tail -3
