"""Peer-memory exchange of the class-sharded head (hgr_score_topk_scatter / hgr_peer_signal / hgr_peer_wait):
G logical ranks inside ONE process on ONE GPU -- the kernels cannot tell a local pointer from a peer mapping, so
the whole data path (row-block scatter, flags, owner-side merge, Hit@k per row block) is covered without NVLink.
The real multi-process / multi-GPU run is tests/test_gpu_dist.py."""
from __future__ import annotations

import pytest
import torch

pytestmark = pytest.mark.gpu


def _emb(n, d, seed):
    x = torch.randn(n, d, generator=torch.Generator().manual_seed(seed))
    return (x / x.norm(dim=-1, keepdim=True)).to(torch.bfloat16)


@pytest.mark.parametrize("B,C,D,G", [(512, 21841, 1024, 8), (130, 1000, 256, 3), (64, 40, 64, 4), (5, 300, 128, 8)])
def test_scatter_exchange_matches_single_gpu(B, C, D, G):
    from hgrnet_b200 import ops
    from hgrnet_b200.dist import PeerExchange, exchange_layout, shard_bounds
    dev = torch.device("cuda", 0)
    K = 20
    x = _emb(B, D, 1).to(dev)
    w = _emb(C, D, 2).to(dev)
    targets = torch.randint(0, C, (B,), generator=torch.Generator().manual_seed(3)).int().to(dev)
    h_ref = ops.new_hits(dev)
    rv, ri = ops.score_topk(x, w, targets=targets, K=K, hits=h_ref)

    lay = exchange_layout(B, K, G, 4)
    bufs = [ops.peer_alloc(lay["total"])[0] for _ in range(G)]
    try:
        ranks = [PeerExchange(B, K, dev, slots=4, _bases=bufs, _rank=r, _world=G) for r in range(G)]
        bounds = shard_bounds(C, G)
        hits = ops.new_hits(dev)
        for rep in range(6):                      # more batches than slots: the sequence counters keep advancing
            slot = rep % 4
            for r, px in enumerate(ranks):        # every rank scores its class shard and scatters to the owners
                lo, hi = bounds[r]
                px.scatter(x, w[lo:hi].contiguous(), lo, slot)
            outs = [px.merge(slot, targets, hits) for px in ranks]
        torch.cuda.synchronize()
        val = torch.cat([o[0] for o in outs if o is not None])
        idx = torch.cat([o[1] for o in outs if o is not None])
        assert torch.equal(idx, ri), "row-block scatter + owner merge differs from the single-GPU result"
        assert torch.equal(val, rv)
        assert hits.tolist() == [6 * h for h in h_ref.tolist()]
        assert [int(px.seq[0]) for px in ranks] == [6] * G and [int(px.seq[1]) for px in ranks] == [6] * G
    finally:
        torch.cuda.synchronize()
        for b in bufs:
            ops.peer_free(b)


@pytest.mark.parametrize("B,D,G,dtype", [(512, 1024, 8, torch.float32), (130, 256, 3, torch.float16), (5, 64, 8, torch.float32)])
def test_feature_ingest_replicates_normalised_rows(B, D, G, dtype):
    """Every logical rank normalises its own row block and broadcasts it; afterwards every replica equals the
    single-GPU normalise of the whole batch, bit for bit."""
    from hgrnet_b200 import ops
    from hgrnet_b200.dist import PeerExchange, exchange_layout
    dev = torch.device("cuda", 0)
    feats = torch.randn(B, D, generator=torch.Generator().manual_seed(4)).to(dtype).to(dev)
    ref = ops.normalize_rows(feats)
    lay = exchange_layout(B, 20, G, 4, D)
    bufs = [ops.peer_alloc(lay["total"])[0] for _ in range(G)]
    try:
        ranks = [PeerExchange(B, 20, dev, slots=4, D=D, _bases=bufs, _rank=r, _world=G) for r in range(G)]
        for xslot in (0, 1, 0):
            # producers first (a single stream: a consumer's wait must not precede the producers it waits for)
            for px in ranks:
                if px.hi > px.lo:
                    ops.normalize_rows_bcast(feats[px.lo:px.hi], px.lo, px.x_ptrs(xslot))
                ops.peer_signal(px.xflag_ptrs, px.seq[2:3])
            for px in ranks:
                ops.peer_wait(px.bases[px.rank] + 64, G, px.seq[3:4])
                assert torch.equal(px.x_view(xslot), ref)
    finally:
        torch.cuda.synchronize()
        for b in bufs:
            ops.peer_free(b)


def test_peer_wait_passes_only_after_all_signals():
    """The wait kernel of a consumer must not complete before every producer has signalled."""
    from hgrnet_b200 import ops
    from hgrnet_b200.dist import PeerExchange, exchange_layout
    dev = torch.device("cuda", 0)
    G, B, K = 3, 12, 20
    lay = exchange_layout(B, K, G, 4)
    bufs = [ops.peer_alloc(lay["total"])[0] for _ in range(G)]
    try:
        ranks = [PeerExchange(B, K, dev, _bases=bufs, _rank=r, _world=G) for r in range(G)]
        side = torch.cuda.Stream()
        done = torch.cuda.Event()
        with torch.cuda.stream(side):
            ops.peer_wait(bufs[0], G, ranks[0].seq[1:2])      # consumer 0 waits on a side stream
            done.record()
        ops.peer_signal(ranks[0].flag_ptrs, ranks[0].seq[0:1])
        ops.peer_signal(ranks[1].flag_ptrs, ranks[1].seq[0:1])
        torch.cuda.current_stream().synchronize()
        assert not done.query(), "wait returned although producer 2 has not signalled"
        ops.peer_signal(ranks[2].flag_ptrs, ranks[2].seq[0:1])
        side.synchronize()
        assert done.query()
    finally:
        torch.cuda.synchronize()
        for b in bufs:
            ops.peer_free(b)


@pytest.mark.parametrize("host_io", [False, True])
def test_sharded_eval_stream_single_rank(host_io):
    """The graph-captured class-sharded evaluator on ONE rank (its own exchange buffer, 4 channels): same lists and
    hit counters as plain score_topk; host-fed mode reads features / labels from pinned memory and copies the
    counters back after every batch."""
    from hgrnet_b200 import ops
    from hgrnet_b200.dist import ShardedEvalStream
    dev = torch.device("cuda", 0)
    B, C, D, K = 300, 5000, 512, 20
    w = _emb(C, D, 2).to(dev)
    feats = [torch.randn(B, D, generator=torch.Generator().manual_seed(10 + s)) for s in range(8)]
    labels = [torch.randint(0, C, (B,), generator=torch.Generator().manual_seed(30 + s)).int() for s in range(8)]
    ses = ShardedEvalStream(w, 0, batch=B, K=K, steps=8, host_io=host_io)
    for s in range(8):
        if host_io:
            ses.host_feats[s].copy_(feats[s])
            ses.host_labels[s].copy_(labels[s])
        else:
            ses.dev_feats[s].copy_(feats[s])
            ses.dev_labels[s].copy_(labels[s])
    ses.run()
    ses.run()
    torch.cuda.synchronize()
    ref_hits = ops.new_hits(dev)
    for s in range(8):
        rv, ri = ops.score_topk(ops.normalize_rows(feats[s].to(dev)), w, targets=labels[s].to(dev), K=K, hits=ref_hits)
        assert torch.equal(ses.idx[s], ri) and torch.equal(ses.val[s], rv)
    assert ses.all_reduce_hits().tolist() == [2 * h for h in ref_hits.tolist()]
    if host_io:
        assert torch.equal(torch.stack(ses.host_hits).max(dim=0).values, ses.hits.cpu())


def _run_certified(x, w, G, K=20, targets=None, reps=1):
    """G logical ranks, global-certificate mode: narrow lists + bounds scattered to the owners, owners certify / repair
    against the shard table.  Returns (val, idx, hits, repaired rows, list length of the first shard)."""
    from hgrnet_b200 import ops
    from hgrnet_b200.dist import PeerExchange, exchange_layout, shard_bounds
    dev = x.device
    B, D = x.shape
    C = w.shape[0]
    lay = exchange_layout(B, K, G, 4)
    bufs = [ops.peer_alloc(lay["total"])[0] for _ in range(G)]
    try:
        ranks = [PeerExchange(B, K, dev, slots=4, _bases=bufs, _rank=r, _world=G) for r in range(G)]
        bounds = shard_bounds(C, G)
        shards = [w[lo:hi].contiguous() for lo, hi in bounds]
        table = ops.shard_table([(s.data_ptr() if s.numel() else 0, 0, s.shape[0], lo) for s, (lo, _) in zip(shards, bounds)], dev)
        hits = ops.new_hits(dev)
        repairs = torch.zeros(1, dtype=torch.int32, device=dev)
        for rep in range(reps):
            slot = rep % 4
            for r, px in enumerate(ranks):
                px.scatter(x, shards[r], bounds[r][0], slot, C_total=C)
            outs = [px.merge(slot, targets, hits if targets is not None else None, certify=(x, table, repairs)) for px in ranks]
        torch.cuda.synchronize()
        val = torch.cat([o[0] for o in outs if o is not None])
        idx = torch.cat([o[1] for o in outs if o is not None])
        kl = ops.global_list_len(B, shards[0].shape[0], D, K, C) if shards[0].shape[0] else K
        return val, idx, hits, int(repairs.item()), kl
    finally:
        torch.cuda.synchronize()
        for b in bufs:
            ops.peer_free(b)


@pytest.mark.parametrize("B,C,D,G", [(4096, 21841, 1024, 8), (4096, 21841, 1024, 4), (4096, 21841, 1024, 2),
                                     (512, 21841, 1024, 8), (130, 1000, 256, 3), (5, 300, 128, 8)])
def test_global_certificate_matches_single_gpu(B, C, D, G):
    """Class shards keep NARROW lists sized for the row's global stream (10 entries at the N = 8 shard of cfg 5) and
    report a bound of what they dropped; the owner certifies against the global K-th value.  On an i.i.d. bank nothing
    needs a repair and the result is bit-identical to the single-GPU head (main.py:136-147)."""
    from hgrnet_b200 import ops
    dev = torch.device("cuda", 0)
    K = 20
    x = _emb(B, D, 11).to(dev)
    w = _emb(C, D, 12).to(dev)
    targets = torch.randint(0, C, (B,), generator=torch.Generator().manual_seed(3)).int().to(dev)
    h_ref = ops.new_hits(dev)
    rv, ri = ops.score_topk(x, w, targets=targets, K=K, hits=h_ref, impl=ops.HGR_IMPL_TCGEN05_EXACT)
    val, idx, hits, repaired, kl = _run_certified(x, w, G, K, targets, reps=3)
    if B == 4096:
        assert kl < K, "the cfg-5 shards are expected to run narrow lists (got %d)" % kl
        assert repaired == 0
    assert torch.equal(idx, ri) and torch.equal(val, rv)
    assert hits.tolist() == [3 * h for h in h_ref.tolist()]


@pytest.mark.parametrize("kind", ["clustered", "ascending", "equal", "dups"])
@pytest.mark.parametrize("B,C,D,G", [(512, 21841, 1024, 8), (130, 2000, 256, 4)])
def test_global_certificate_repairs_hostile_banks(kind, B, C, D, G):
    """Banks whose best classes sit in adjacent rows of ONE shard overflow that shard's narrow lists: its bound reaches
    the global K-th value, the owner must notice and re-scan that shard exactly (here the 'peer' banks are local)."""
    from hgrnet_b200 import ops
    from tests.util import compare_topk, hostile_bank, hostile_queries
    dev = torch.device("cuda", 0)
    w = hostile_bank(C, D, kind).to(dev)
    xn = hostile_queries(w, B, kind)
    logits = (xn.float() @ w.float().T).cpu()
    val, idx, _, repaired, kl = _run_certified(xn, w, G)
    compare_topk(val, idx, logits, torch.arange(C), 20)
    if kl < 20:
        assert repaired > 0, "a hostile bank should defeat the narrow lists of at least one row"
