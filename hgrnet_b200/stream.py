"""Streaming eval step with static buffers, replayed as CUDA graphs on a few round-robin streams.

The reference's eval loop (main.py:131-148) launches ~10 small torch ops per batch from Python;
at B200 speeds one batch of the fused head is ~20 us of GPU time, less than the CPU cost of
issuing it, and a large part of those 20 us is per-kernel fixed latency (launch, TMEM/barrier
set-up, first HBM touch, grid tail).  ``EvalStream`` removes both from the per-batch path: for a
fixed batch shape it owns pinned host staging buffers (the place a ``pin_memory=True`` loader
writes into, dataset/imagenet_group_test.py:83), device buffers and one captured CUDA graph per
slot that does

    H2D(features, labels) -> row-normalise (kernel 1) -> fused logits/top-K/Hit@k (kernel 2 + merge)
    -> D2H(hit counters)

A step is one ``cudaGraphLaunch``.  Consecutive slots are replayed on different streams (batches
are independent), so the copy of batch i+1 and the small kernels of neighbouring batches overlap
the GEMM of batch i and CTAs of the next GEMM start as soon as SMs drain.  Results (top-K values /
node ids) stay in the slot's device buffers; the Hit@{1,2,5,10,20} counters accumulate on the
device (atomics commute) and are mirrored to pinned host memory every step.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import ops


class EvalStream:
    def __init__(self, bank: torch.Tensor, col_id: Optional[torch.Tensor] = None, id_base: int = 0, *, batch: int,
                 feat_dtype=torch.float32, K: int = 20, slots: int = 4, streams: int = 2, banks=None,
                 host_io: bool = True):
        """``bank`` [C, D] bf16 on the device.  ``banks``: optional list of bank tensors to rotate over per slot
        (benchmarks use it to keep the working set larger than L2).  ``host_io=False`` skips the H2D / D2H nodes
        (features are then written straight into ``dev_feats`` / ``dev_labels``)."""
        self.device = bank.device
        self.K = K
        self.B = batch
        D = bank.shape[1]
        self.banks = list(banks) if banks is not None else [bank]
        self.col_id, self.id_base = col_id, id_base
        self.slots = slots
        self.host_io = host_io
        # features and labels of a slot share ONE staging buffer on each side, so a batch is a single DMA (a second,
        # 2 KB copy costs a descriptor's fixed latency on the same copy engine that carries the 2 MB of features)
        fbytes = batch * D * torch.empty((), dtype=feat_dtype).element_size()
        nbytes = fbytes + batch * 4

        def views(pack):
            return pack[:fbytes].view(feat_dtype).view(batch, D), pack[fbytes:].view(torch.int32)

        self.dev_pack = [torch.zeros(nbytes, dtype=torch.uint8, device=self.device) for _ in range(slots)]
        self.dev_feats, self.dev_labels = map(list, zip(*[views(p) for p in self.dev_pack]))
        if host_io:
            self.host_pack = [torch.zeros(nbytes, dtype=torch.uint8).pin_memory() for _ in range(slots)]
            self.host_feats, self.host_labels = map(list, zip(*[views(p) for p in self.host_pack]))
            self.host_hits = torch.zeros(ops.HGR_NUM_HITS, dtype=torch.int64).pin_memory()
        self.hits = ops.new_hits(self.device)
        self.val = [None] * slots
        self.idx = [None] * slots
        self.streams = [torch.cuda.Stream(device=self.device) for _ in range(max(1, min(streams, slots)))]
        self.graphs = []
        self.num_samples = 0
        self._capture()

    def _stream_of(self, slot: int):
        return self.streams[slot % len(self.streams)]

    def _body(self, s: int):
        if self.host_io:
            self.dev_pack[s].copy_(self.host_pack[s], non_blocking=True)
        x = ops.normalize_rows(self.dev_feats[s])                                    # clip_tree.py:330
        self.val[s], self.idx[s] = ops.score_topk(x, self.banks[s % len(self.banks)], col_id=self.col_id,
                                                  id_base=self.id_base, targets=self.dev_labels[s], K=self.K,
                                                  hits=self.hits)                    # clip_tree.py:331 + main.py:136-147
        if self.host_io:
            self.host_hits.copy_(self.hits, non_blocking=True)

    def _capture(self):
        cur = torch.cuda.current_stream(self.device)
        for st in self.streams:
            st.wait_stream(cur)
        # a slot is always captured and replayed on the same stream: the kernel workspace is per stream
        for s in range(self.slots):
            with torch.cuda.stream(self._stream_of(s)):
                self._body(s)                             # warm-up: workspace allocation, lazy module loading
        for st in self.streams:
            st.synchronize()
        self.hits.zero_()
        torch.cuda.synchronize(self.device)
        for s in range(self.slots):
            st = self._stream_of(s)
            with torch.cuda.stream(st):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=st):
                    self._body(s)
                self.graphs.append(g)
        torch.cuda.synchronize(self.device)

    # ------------------------------------------------------------------ per-batch API
    def begin(self):
        """Order the worker streams after everything already queued on the caller's current stream."""
        cur = torch.cuda.current_stream(self.device)
        for st in self.streams:
            st.wait_stream(cur)

    def step(self, slot: int):
        """Score the batch held in ``host_feats[slot]`` / ``host_labels[slot]`` (one graph launch; returns at once)."""
        with torch.cuda.stream(self._stream_of(slot)):
            self.graphs[slot].replay()
        self.num_samples += self.B

    def end(self):
        """Make the caller's current stream wait for all steps issued so far."""
        cur = torch.cuda.current_stream(self.device)
        for st in self.streams:
            cur.wait_stream(st)

    def synchronize(self):
        for st in self.streams:
            st.synchronize()

    def hit_counts(self):
        """Hit@{1,2,5,10,20} so far (synchronises)."""
        self.synchronize()
        # (the per-step D2H copies of several streams land in the same 40 bytes in no particular order: read the
        # counters once more now that every stream is idle)
        return self.hits.tolist()
