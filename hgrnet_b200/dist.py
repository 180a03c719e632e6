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
