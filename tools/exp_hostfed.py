"""Host-fed class-sharded evaluator: per-step time for a few pipeline shapes (run under torch.distributed.run)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from hgrnet_b200.dist import ShardedEvalStream, shard_bounds
from hgrnet_b200.synthetic import synthetic_embeddings

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
B, C, D, K = 4096, 21841, 1024, 20
lo, hi = shard_bounds(C, world)[rank]
bank = synthetic_embeddings(C, D, 1)[lo:hi].to(dev).to(torch.bfloat16)
banks = [bank, bank.clone()]


def run(steps, channels, host_io, dtype, reps=30):
    ses = ShardedEvalStream(banks[0], lo, batch=B, K=K, steps=steps, banks=banks, host_io=host_io, channels=channels,
                            feat_dtype=dtype)
    g = torch.Generator().manual_seed(3)
    for s in range(steps):
        if host_io:
            ses.host_feats[s].copy_(torch.randn(ses.row_hi - ses.row_lo, D, generator=g).to(dtype))
            ses.host_labels[s].fill_(7)
        else:
            ses.dev_feats[s].copy_(torch.randn(B, D, generator=g).to(dtype))
    for _ in range(3):
        ses.run()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        ses.run()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / (reps * steps)], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print("N=%d steps=%2d channels=%d host_io=%d %-8s  %.1f us/step  %.1f M img/s" % (
            world, steps, ses.channels, host_io, str(dtype).replace("torch.", ""), float(t) * 1e3, B / float(t) / 1e3), flush=True)
    del ses
    torch.cuda.synchronize()
    dist.barrier()


for (steps, ch) in ((10, 4), (16, 8), (12, 6), (8, 2)):
    run(steps, ch, False, torch.float32)
    run(steps, ch, True, torch.float32)
    run(steps, ch, True, torch.float16)
sys.stdout.flush()
torch.cuda.synchronize()
dist.barrier()
os._exit(0)
