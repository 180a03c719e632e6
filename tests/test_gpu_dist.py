"""Multi-GPU (NCCL) parity of the class-sharded head: needs >= 2 visible GPUs, skipped otherwise.
Each rank scores the whole batch against its bank shard with kernel (2); one all-gather; merge + Hit@k."""
from __future__ import annotations

import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, B, C, D, K, out):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from hgrnet_b200 import ops
        from hgrnet_b200.dist import ShardedScorer, shard_bounds
        g = torch.Generator().manual_seed(9)
        x = torch.randn(B, D, generator=g)
        w = torch.randn(C, D, generator=g)
        w = (w / w.norm(dim=-1, keepdim=True)).to(torch.bfloat16)
        targets = torch.randint(0, C, (B,), generator=g).int()
        xn = ops.normalize_rows(x.to(dev))
        lo, hi = shard_bounds(C, world)[rank]
        sc = ShardedScorer(w[lo:hi].to(dev).contiguous(), None, id_base=lo, K=K)
        hits = ops.new_hits(dev)
        val, idx = sc.score(xn, targets.to(dev), hits)
        # pipelined API: two batches in flight
        sc.submit(xn, targets.to(dev))
        sc.submit(xn, targets.to(dev))
        v1, i1 = sc.collect(hits)
        v2, i2 = sc.collect(hits)
        torch.cuda.synchronize()
        # reference: the same head on one GPU (whole bank)
        h1 = ops.new_hits(dev)
        rv, ri = ops.score_topk(xn, w.to(dev), targets=targets.to(dev), K=K, hits=h1)
        assert torch.equal(ri, idx) and torch.equal(rv, val), "sharded result differs from the single-GPU result"
        assert torch.equal(i1, idx) and torch.equal(i2, idx)
        assert hits.tolist() == [3 * h for h in h1.tolist()]
        # graph-captured streaming evaluator, both exchanges: peer-memory scatter (default) and NCCL all-gather
        from hgrnet_b200.dist import ShardedEvalStream
        for exchange in ("p2p", "nccl"):
            ses = ShardedEvalStream(w[lo:hi].to(dev).contiguous(), lo, batch=B, K=K, steps=8, exchange=exchange)
            for s_ in range(8):
                ses.dev_feats[s_].copy_(x.to(dev))
                ses.dev_labels[s_].copy_(targets.to(dev))
            for _ in range(3):
                ses.run()
            torch.cuda.synchronize()
            hs = ses.all_reduce_hits()
            assert hs.tolist() == [24 * h for h in h1.tolist()], (exchange, hs.tolist(), h1.tolist())
            for s_ in (0, 3, 7):
                assert torch.equal(ses.idx[s_], ri[ses.row_lo:ses.row_hi]), exchange
                assert torch.equal(ses.val[s_], rv[ses.row_lo:ses.row_hi]), exchange
            dist.barrier()
        # hostile bank (siblings adjacent and similar, queries near a leaf): a shard's narrow lists overflow, the row
        # owner must notice (bound >= global K-th value) and re-scan that shard -- a PEER's bank, over NVLink
        from tests.util import hostile_bank, hostile_queries
        wh = hostile_bank(C, D, "clustered")
        xh = hostile_queries(wh.to(dev), B, "clustered")
        ses = ShardedEvalStream(wh[lo:hi].to(dev).contiguous(), lo, batch=B, K=K, steps=2)
        for s_ in range(2):
            ses.dev_feats[s_].copy_(xh.float())
        ses.run()
        torch.cuda.synchronize()
        hv, hi_ = ops.score_topk(ops.normalize_rows(xh.float()), wh.to(dev), K=K, impl=ops.HGR_IMPL_SIMT)
        rep = ses.repairs.clone()
        dist.all_reduce(rep)
        if ops.global_list_len(B, hi - lo, D, K, C) < K:
            assert int(rep.item()) > 0, "the clustered bank should defeat the narrow lists of some rows"
        mine_v, mine_i = ses.val[1], ses.idx[1]
        assert torch.allclose(mine_v, hv[ses.row_lo:ses.row_hi], rtol=1e-3, atol=1e-5)
        assert (mine_i != hi_[ses.row_lo:ses.row_hi]).any(1).float().mean().item() < 0.1   # near-ties may swap
        dist.barrier()
        # host-fed evaluator: every rank copies only its row block from pinned memory, NVLink replicates the rest
        ses = ShardedEvalStream(w[lo:hi].to(dev).contiguous(), lo, batch=B, K=K, steps=8, host_io=True)
        for s_ in range(8):
            ses.host_feats[s_].copy_(x[ses.row_lo:ses.row_hi])
            ses.host_labels[s_].copy_(targets[ses.row_lo:ses.row_hi])
        for _ in range(2):
            ses.run()
        torch.cuda.synchronize()
        assert ses.all_reduce_hits().tolist() == [16 * h for h in h1.tolist()]
        # every batch reads the counters back; channels run concurrently, so a snapshot may miss a neighbour's
        # batch, but the snapshot of the batch that finished last has seen them all
        snaps = torch.stack(ses.host_hits)
        assert torch.equal(snaps.max(dim=0).values, ses.hits.cpu()), (snaps.tolist(), ses.hits.tolist())
        for s_ in (0, 5):
            assert torch.equal(ses.idx[s_], ri[ses.row_lo:ses.row_hi]) and torch.equal(ses.val[s_], rv[ses.row_lo:ses.row_hi])
        dist.barrier()
        out.put(("ok", rank, h1.tolist()))
    except BaseException as e:  # noqa: BLE001 -- report, then leave without tearing NCCL down
        import traceback
        out.put(("fail", rank, traceback.format_exc()[-1500:]))
    finally:
        # CUDA graphs that hold NCCL kernels make destroy_process_group block: results are already in the queue
        os._exit(0)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("B,C", [(512, 21841), (130, 1000)])
def test_class_sharded_head_matches_single_gpu(B, C):
    import torch.multiprocessing as mp
    world = min(torch.cuda.device_count(), 8)
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, B, C, 1024, 20, out)) for r in range(world)]
    for p in procs:
        p.start()
    stuck = False
    for p in procs:
        p.join(150)
        if p.exitcode is None:      # a peer that failed leaves the others waiting in a collective: report what we have
            p.kill()
            stuck = True
    res = []
    while not out.empty():
        res.append(out.get())
    assert not stuck and len(res) == world and all(r[0] == "ok" for r in res), res
