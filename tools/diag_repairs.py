import sys
sys.path.insert(0, "/root/repo/tools"); sys.path.insert(0, "/root/repo")
import torch
from hgrnet_b200 import ops
from sweep import emb
for C in (2731, 21841):
    w = emb(C, 1024, 2).cuda(); x = emb(4096, 1024, 10).cuda()
    for it in range(3):
        v, i = ops.score_topk(x, w, K=20)
        torch.cuda.synchronize()
        print(C, "repairs", ops.last_rescan_count("cuda:0"))
    ve, ie = ops.score_topk(x, w, K=20, impl=ops.HGR_IMPL_TCGEN05_EXACT)
    print("equal to exact:", bool(torch.equal(i, ie)), bool(torch.equal(v, ve)))
