"""cProfile of tree_model.train_batch at cfg 3 (host side: where the wall clock of an OM step goes)."""
import cProfile
import os
import pstats
import random
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from hgrnet_b200.flags import parse_args
from hgrnet_b200.head import tree_model
from hgrnet_b200.hierarchy import WORDNET_LIKE_21841, synthetic_hierarchy
from hgrnet_b200.synthetic import TableEncoder, node_id_tokens, synthetic_embeddings

DEV = "cuda:0"
hier = synthetic_hierarchy(WORDNET_LIKE_21841, seed=1)
N, D, B = len(hier), 1024, 256
table = synthetic_embeddings(N, D, 1, normalize=False)
opts = parse_args([])
opts.device, opts.folder = 0, "/tmp/hgr_om_prof"
model = tree_model(opts, hier.nodes, hier.nodes, clip_model=TableEncoder(table).to(DEV), hierarchy=hier,
                   node_tokens=node_id_tokens(N))
target = N - 1
img = synthetic_embeddings(B, D, 5, normalize=False).to(DEV).requires_grad_(True)
targets = torch.full((B,), target, dtype=torch.long, device=DEV)
random.seed(0)
for _ in range(5):
    model.train_batch(img, targets, "OM", "topk")
torch.cuda.synchronize()
ts = []
for _ in range(30):
    t0 = time.perf_counter()
    model.train_batch(img, targets, "OM", "topk")
    torch.cuda.synchronize()
    ts.append(time.perf_counter() - t0)
ts.sort()
print("median wall clock per step: %.3f ms (best %.3f)" % (ts[15] * 1e3, ts[0] * 1e3))
pr = cProfile.Profile()
pr.enable()
for _ in range(30):
    model.train_batch(img, targets, "OM", "topk")
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(30)
st.sort_stats("tottime").print_stats(25)
