"""world_size-2 gloo test (CPU) of the host-side logic of the class-sharded head: shard bounds, the packed
candidate record exchanged by the single all-gather, and that merging the gathered per-rank lists equals the
oracle's global top-K.  (The CUDA merge kernel itself is covered by the -m gpu tests.)"""
from __future__ import annotations

import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import hgr_oracle as orc


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, B, C, D, K, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from hgrnet_b200.dist import pack_candidates, shard_bounds, unpack_gathered
        g = torch.Generator().manual_seed(5)
        x = torch.randn(B, D, generator=g)
        w = torch.randn(C, D, generator=g)
        logits = orc.forward_logits(x, orc.normalize_rows(w))          # every rank holds the full oracle logits
        lo, hi = shard_bounds(C, world)[rank]
        k_loc = min(K, hi - lo)
        val = torch.full((B, K), float("-inf"))
        idx = torch.full((B, K), -1, dtype=torch.int32)
        if k_loc > 0:                                                   # what kernel (2) returns for this shard
            v, i = logits[:, lo:hi].topk(k_loc, 1, True, True)
            val[:, :k_loc], idx[:, :k_loc] = v, (i + lo).int()
        send = pack_candidates(val, idx)
        recv = torch.empty((world,) + tuple(send.shape), dtype=torch.int32)
        dist.all_gather_into_tensor(recv.view(-1), send.view(-1))       # THE one collective on the data path
        pv, pi = unpack_gathered(recv)
        assert pv.shape == (world, B, K) and pi.dtype == torch.int32
        assert pv.stride(0) == pi.stride(0) == 2 * B * K and pv.stride(1) == K    # merge-in-place layout
        flat_v = pv.permute(1, 0, 2).reshape(B, world * K)
        flat_i = pi.permute(1, 0, 2).reshape(B, world * K)
        mv, mp_ = flat_v.topk(K, 1, True, True)
        mi = flat_i.gather(1, mp_)
        ov, oi = logits.topk(K, 1, True, True)
        assert torch.equal(mv, ov) and torch.equal(mi.long(), oi)
        if rank == 0:
            out.put("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("B,C,K", [(33, 1001, 20), (4, 30, 20)])
def test_sharded_candidates_merge_to_global_topk(B, C, K):
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, B, C, 64, K, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert out.get() == "ok"


def test_shard_bounds_cover_the_bank_once():
    from hgrnet_b200.dist import shard_bounds
    for C, G in ((21841, 8), (21841, 3), (5, 8), (16, 2), (0, 4)):
        b = shard_bounds(C, G)
        assert len(b) == G and b[0][0] == 0 and b[-1][1] == C
        assert all(b[i][1] == b[i + 1][0] for i in range(G - 1))
        assert all(lo <= hi for lo, hi in b)
    assert shard_bounds(21841, 8)[0] == (0, 2731) and shard_bounds(21841, 8)[7] == (19117, 21841)


def test_exchange_layout_and_row_blocks():
    """Host arithmetic of the peer-memory exchange: regions do not overlap, feature slots are 256-byte aligned, row
    blocks cover the batch exactly once (ragged last block, ranks without rows)."""
    from hgrnet_b200.dist import X_SLOTS, exchange_layout
    for B, K, world, slots, D in [(4096, 20, 8, 4, 1024), (130, 20, 3, 4, 256), (5, 20, 8, 4, 64), (512, 5, 1, 2, 0)]:
        lay = exchange_layout(B, K, world, slots, D)
        rows = lay["block_rows"]
        assert rows * world >= B and (rows - 1) * world < B       # smallest block size that covers the batch
        assert lay["part_bytes"] == rows * K * 4
        # value lists, id lists, then the [world, block_rows] fp32 bounds of the global certificate (16-byte padded)
        assert lay["slot_bytes"] % 16 == 0 and lay["slot_bytes"] >= 2 * world * lay["part_bytes"] + world * rows * 4
        assert lay["header"] >= 64 + 4 * 16                      # two flag sets of up to 16 ranks
        assert lay["x_off"] % 256 == 0 and lay["x_off"] >= lay["header"] + slots * lay["slot_bytes"]
        assert lay["total"] == lay["x_off"] + (X_SLOTS * lay["x_bytes"] if D else 0)
        if D:
            assert lay["x_bytes"] % 256 == 0 and lay["x_bytes"] >= B * D * 2
        covered = []
        for r in range(world):
            lo = min(B, r * rows)
            hi = min(B, lo + rows)
            covered.extend(range(lo, hi))
        assert covered == list(range(B))


def _peer_setup_worker(rank, world, port, failing_rank, stage, out):
    """PeerExchange set-up with faked allocation / mapping calls: the rank `failing_rank` fails at `stage`."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import contextlib

        from hgrnet_b200 import dist as hd
        from hgrnet_b200 import ops
        freed = []

        def fake_alloc(nbytes):
            if stage == "alloc" and rank == failing_rank:
                raise RuntimeError("out of memory (fake)")
            return 0x1000 * (rank + 1), bytes([rank]) * 64

        def fake_open(handle):
            if stage == "open" and rank == failing_rank:
                raise RuntimeError("cudaIpcOpenMemHandle: peer access unsupported (fake)")
            return 0x100000 + handle[0]

        ops.peer_alloc, ops.peer_open = fake_alloc, fake_open
        ops.peer_close = lambda p: freed.append(("close", p))
        ops.peer_free = lambda p: freed.append(("free", p))
        torch.cuda.device = lambda d: contextlib.nullcontext()         # no CUDA on this box
        px = hd.PeerExchange.__new__(hd.PeerExchange)
        px.device, px.world, px.rank, px._opened, px._own = "cpu", world, rank, [], None
        px.lay = hd.exchange_layout(64, 20, world, 4)
        try:
            px._map_peers(None)
            out.put((rank, "mapped", px.bases))
        except hd.PeerMemoryUnavailable as e:
            out.put((rank, "unavailable", str(e), freed))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("stage", ["alloc", "open", "none"])
def test_peer_exchange_setup_failure_is_agreed_on_by_all_ranks(stage):
    """If ANY rank cannot allocate or map the exchange buffers, EVERY rank must raise PeerMemoryUnavailable (and free
    what it holds) -- otherwise the healthy ranks would sit in the next collective forever."""
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_peer_setup_worker, args=(r, 2, port, 1, stage, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    res = sorted(out.get() for _ in range(2))
    if stage == "none":
        assert [r[1] for r in res] == ["mapped", "mapped"]
        assert res[0][2] == [0x1000, 0x100001] and res[1][2] == [0x100000, 0x2000]
    else:
        assert [r[1] for r in res] == ["unavailable", "unavailable"], res
        if stage == "open":
            assert ("free", 0x1000) in res[0][3] and ("close", 0x100001) in res[0][3]    # the healthy rank cleaned up


def test_exchange_slot_rings_never_reuse_a_slot_back_to_back():
    """ShardedEvalStream: a producer may overwrite a list slot of batch j only after the owner's next signal, so two
    consecutive batches of a channel -- including the last of one graph replay and the first of the next -- must use
    different slots (hgrnet_b200.dist._ring_len)."""
    from hgrnet_b200.dist import X_SLOTS, _ring_len
    for slots in (4, X_SLOTS):
        for n_c in range(2, 13):
            ring = _ring_len(n_c, slots)
            assert 2 <= ring <= slots
            seq = [j % ring for j in range(n_c)] * 3            # three replays back to back
            assert all(a != b for a, b in zip(seq, seq[1:])), (n_c, ring)
    with pytest.raises(ValueError):
        _ring_len(13, 4)
