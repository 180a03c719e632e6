"""Class-sharded multi-GPU eval head (one process per GPU, ``torch.distributed`` / NCCL plumbing).

The reference is single-GPU (main.py:226).  Here the class bank is cut row-wise into ``G``
contiguous shards; every rank scores the WHOLE image batch against its shard
with the fused kernel (2), the per-rank ``[B, K]`` candidate lists (value, node id) are exchanged
with ONE ``all_gather`` per batch over NVLink, and every rank merges the ``G`` lists with
``hgr_topk_merge`` (which also counts Hit@k -- a label hit needs the GLOBAL rank, so hits cannot
be counted per shard).  No other collective is on the data path.

``ShardedScorer.score`` is synchronous; ``ShardedScorer.submit`` / ``collect`` pipeline batches so
that the all-gather + merge of batch i overlap the GEMM of batch i+1 (the exchange is latency
bound: 2*B*K*4 bytes per rank).
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist

from . import ops


def shard_bounds(C: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous row ranges ``[lo, hi)`` of ``ceil(C / world)`` rows (the last ones may be short or empty)."""
    per = (C + world - 1) // world
    return [(min(C, r * per), min(C, (r + 1) * per)) for r in range(world)]


def pack_candidates(val: torch.Tensor, idx: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """One ``[2, B, K]`` int32 record per rank: fp32 value bits, then node ids -> a single all-gather."""
    if out is None:
        out = torch.empty((2,) + tuple(val.shape), dtype=torch.int32, device=val.device)
    out[0].copy_(val.view(torch.int32))
    out[1].copy_(idx)
    return out


def unpack_gathered(buf: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """``[G, 2, B, K]`` int32 -> strided ``[G, B, K]`` views (fp32 values, int32 ids), no copy."""
    return buf[:, 0].view(torch.float32), buf[:, 1]


class ShardedScorer:
    def __init__(self, bank_shard: torch.Tensor, col_id_shard: Optional[torch.Tensor], id_base: int = 0, K: int = 20,
                 group=None, depth: int = 2):
        self.bank = bank_shard
        self.col_id = col_id_shard
        self.id_base = id_base
        self.K = K
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.depth = depth
        self._slots = []          # per in-flight batch: (send buffer, gather buffer, work handle, targets, event)
        self._free = []
        self._merge_stream = None

    # ------------------------------------------------------------------ synchronous
    def score(self, x_norm: torch.Tensor, targets: Optional[torch.Tensor], hits: Optional[torch.Tensor] = None):
        self.submit(x_norm, targets)
        return self.collect(hits)

    # ------------------------------------------------------------------ pipelined
    def _buffers(self, B, device):
        while self._free:
            send, recv = self._free.pop()
            if send.shape[1] == B:
                return send, recv
        send = torch.empty((2, B, self.K), dtype=torch.int32, device=device)
        recv = torch.empty((self.world, 2, B, self.K), dtype=torch.int32, device=device)
        return send, recv

    def submit(self, x_norm: torch.Tensor, targets: Optional[torch.Tensor]):
        """Enqueue local scoring + the all-gather of one batch; returns immediately."""
        B = x_norm.shape[0]
        send, recv = self._buffers(B, x_norm.device)
        if self.bank.shape[0] > 0:
            val, idx = ops.score_topk(x_norm, self.bank, col_id=self.col_id, id_base=self.id_base, K=self.K)
        else:  # empty shard (more ranks than class chunks)
            val = torch.full((B, self.K), float("-inf"), device=x_norm.device)
            idx = torch.full((B, self.K), -1, dtype=torch.int32, device=x_norm.device)
        pack_candidates(val, idx, send)
        if self.world > 1:
            work = dist.all_gather_into_tensor(recv.view(-1), send.view(-1), group=self.group, async_op=True)
        else:
            recv[0].copy_(send)
            work = None
        self._slots.append((send, recv, work, targets))

    def collect(self, hits: Optional[torch.Tensor] = None):
        """Merge the oldest in-flight batch -> ``(val [B,K], idx [B,K])`` (+ hits)."""
        send, recv, work, targets = self._slots.pop(0)
        if work is not None:
            work.wait()  # orders the current stream after the collective; no host block
        pv, pi = unpack_gathered(recv)
        t = targets.to(torch.int32) if targets is not None else None
        out = ops.topk_merge(pv, pi, targets=t, hits=hits)
        self._free.append((send, recv))
        return out


class ShardedEvalStream:
    """Software-pipelined, CUDA-graph captured class-sharded eval steps (one process per GPU).

    One graph holds ``steps`` consecutive batches.  Inside it the compute stream runs
    ``normalise -> fused score/top-K on the local shard -> pack`` of batch i+1 BEFORE it waits for the all-gather
    of batch i (issued on NCCL's stream), so the exchange and the merge of a batch overlap the GEMM of the next;
    collectives stay in batch order on NCCL's single stream, on every rank.  A replay is one ``cudaGraphLaunch``
    for ``steps`` batches: no interpreter and no per-batch launch latency on the path.
    """

    def __init__(self, bank_shard: torch.Tensor, id_base: int, *, batch: int, K: int = 20, steps: int = 8,
                 feat_dtype=torch.float32, banks=None, group=None, use_graph: bool = True):
        self.device = bank_shard.device
        self.banks = list(banks) if banks is not None else [bank_shard]
        self.id_base, self.K, self.B, self.steps, self.group = id_base, K, batch, steps, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        D = bank_shard.shape[1]
        self.dev_feats = [torch.empty((batch, D), dtype=feat_dtype, device=self.device) for _ in range(steps)]
        self.dev_labels = [torch.zeros((batch,), dtype=torch.int32, device=self.device) for _ in range(steps)]
        self.send = [torch.empty((2, batch, K), dtype=torch.int32, device=self.device) for _ in range(steps)]
        self.recv = [torch.empty((self.world, 2, batch, K), dtype=torch.int32, device=self.device) for _ in range(steps)]
        self.hits = ops.new_hits(self.device)
        self.val = [None] * steps
        self.idx = [None] * steps
        self.graph = None
        self.stream = torch.cuda.Stream(device=self.device)
        self.stream.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self.stream):
            self._issue()                      # warm-up: NCCL communicator, workspaces
            self.stream.synchronize()
            self.hits.zero_()
            if use_graph:
                try:
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, stream=self.stream):
                        self._issue()
                    self.graph = g
                except Exception as e:  # pragma: no cover - depends on the NCCL / torch build
                    print("ShardedEvalStream: graph capture of the NCCL pipeline failed (%r); running eagerly" % (e,))
                    self.graph = None
        torch.cuda.synchronize(self.device)
        self.hits.zero_()

    def _local(self, s: int):
        x = ops.normalize_rows(self.dev_feats[s])
        bank = self.banks[s % len(self.banks)]
        if bank.shape[0] > 0:
            # results go straight into the send record (no pack copies)
            ops.score_topk(x, bank, id_base=self.id_base, K=self.K, out=(self.send[s][0].view(torch.float32), self.send[s][1]))
        else:
            self.send[s][0].copy_(torch.full((self.B, self.K), float("-inf"), device=self.device).view(torch.int32))
            self.send[s][1].fill_(-1)
        if self.world > 1:
            return dist.all_gather_into_tensor(self.recv[s].view(-1), self.send[s].view(-1), group=self.group,
                                               async_op=True)
        self.recv[s][0].copy_(self.send[s])
        return None

    def _merge(self, s: int, work):
        if work is not None:
            work.wait()
        pv, pi = unpack_gathered(self.recv[s])
        self.val[s], self.idx[s] = ops.topk_merge(pv, pi, targets=self.dev_labels[s], hits=self.hits)

    def _issue(self):
        prev = None
        for s in range(self.steps):
            work = self._local(s)
            if prev is not None:
                self._merge(*prev)
            prev = (s, work)
        self._merge(*prev)

    def run(self):
        """Score the ``steps`` batches currently in ``dev_feats`` / ``dev_labels`` (enqueue only)."""
        if self.graph is not None:
            self.graph.replay()
        else:
            self._issue()
