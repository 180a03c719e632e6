"""Per-stage device timeline of the host-fed class-sharded evaluator (run under torch.distributed.run).

One exchange channel, eager issue (no graph), CUDA events around every stage of a batch on this rank's stream:
H2D of the rank's row block -> normalise + broadcast over NVLink -> flag signal -> wait for all blocks -> scoring kernel +
scatter of the lists -> flag signal -> wait for all lists -> merge of my rows -> D2H of the hit counters.  The waits
contain the skew between ranks; everything is per batch, averaged over the batches of the run, maximum over ranks."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from hgrnet_b200 import ops
from hgrnet_b200.dist import ShardedEvalStream, shard_bounds
from hgrnet_b200.synthetic import synthetic_embeddings

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
B, C, D, K = 4096, 21841, 1024, 20
lo, hi = shard_bounds(C, world)[rank]
bank = synthetic_embeddings(C, D, 1)[lo:hi].to(dev).to(torch.bfloat16)

marks = []          # (stage name, start event, end event)


def timed(name, fn):
    def wrapper(*a, **kw):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn(*a, **kw)
        e1.record()
        marks.append((name, e0, e1))
        return out
    return wrapper


state = {"n": 0}


def wait_name(fn):          # peer_wait is used twice per batch: features first, lists second
    def wrapper(*a, **kw):
        state["n"] += 1
        return timed("wait for all feature blocks" if state["n"] % 2 == 1 else "wait for all lists", fn)(*a, **kw)
    return wrapper


def signal_name(fn):
    def wrapper(*a, **kw):
        return timed("flag signal", fn)(*a, **kw)
    return wrapper


for dtype in (torch.float32, torch.float16):
    ses = ShardedEvalStream(bank, lo, batch=B, K=K, steps=8, channels=1, host_io=True, use_graph=False, feat_dtype=dtype)
    g = torch.Generator().manual_seed(3)
    for s in range(8):
        ses.host_feats[s].copy_(torch.randn(ses.row_hi - ses.row_lo, D, generator=g).to(dtype))
        ses.host_labels[s].fill_(7)
    saved = (ops.normalize_rows_bcast, ops.peer_signal, ops.peer_wait, ops.score_topk_scatter, ops.topk_merge_raw)
    ops.normalize_rows_bcast = timed("normalise + NVLink broadcast of my rows", saved[0])
    ops.peer_signal = signal_name(saved[1])
    ops.peer_wait = wait_name(saved[2])
    ops.score_topk_scatter = timed("scoring kernel + scatter of the lists", saved[3])
    ops.topk_merge_raw = timed("merge of my rows", saved[4])
    for rep in range(6):
        if rep == 2:
            marks.clear()
            state["n"] = 0
        ses.run()
        torch.cuda.synchronize()
        dist.barrier()
    ops.normalize_rows_bcast, ops.peer_signal, ops.peer_wait, ops.score_topk_scatter, ops.topk_merge_raw = saved
    agg = {}
    for name, e0, e1 in marks:
        agg.setdefault(name, []).append(e0.elapsed_time(e1) * 1e3)
    names = list(agg)
    t = torch.tensor([sum(agg[n]) / len(agg[n]) for n in names], device=dev)
    tmax = t.clone()
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    if rank == 0:
        print("N=%d host-fed, %s features, one channel, eager: us per batch and stage (mean over ranks / slowest rank)" % (
            world, str(dtype).replace("torch.", "")))
        for n, a, m in zip(names, (t / world).tolist(), tmax.tolist()):
            print("  %-42s %8.1f / %8.1f" % (n, a, m))
        print("  (H2D of the row block: %.2f MB per rank and batch; not bracketed -- it is a copy node on the same stream)" % (
            (ses.row_hi - ses.row_lo) * D * torch.empty((), dtype=dtype).element_size() / 1e6), flush=True)
    del ses
    torch.cuda.synchronize()
    dist.barrier()
sys.stdout.flush()
os._exit(0)
