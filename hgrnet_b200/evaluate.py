"""Re-hosted eval loop of the reference (``main.test``, main.py:104-222).

Per batch the reference materialises logits [B,N], gathers the test columns, runs
``topk(20)``, maps ids, compares with the label and sums Hit@{1,2,5,10,20} (main.py:135-147).
Here that whole chain is ONE fused call, ``model.score_topk`` (kernel (1) + kernel (2)); the
hit counters stay on the device until the loop ends, as in the reference.  The hierarchical
metrics TOR / POR (``hit_ratio`` / ``path_ratio`` / ``point_ratio``, main.py:143,152-191) are the
first "next" row of SURVEY.md section 8f: they are computed from the dense logits of
``hgr_logits_dense`` with vectorised device ops (no per-sample python loop, no D2H copies).
Output format: ``count_acc`` (utils.py:135-146) and the log files of main.py:205-222.
"""
from __future__ import annotations

import copy
from typing import Iterable, Optional

import torch

from . import ops
from ._cabi import HIT_CUTS


def count_acc(hits_dict, num_tot):
    """utils.py:135-146."""
    out_str = ""
    acc_dict = dict()
    keys = list(hits_dict.keys())
    for key, value in hits_dict.items():
        acc = value / num_tot * 100.0
        acc_dict[key] = acc
        out_str += "Top@{}(%):{:.2f}".format(key, acc)
        out_str += ", " if key != keys[-1] else "."
    return out_str, acc_dict


class HierMetrics:
    """TOR / POR accumulators (main.py:123-127,152-191).  One fused pass per batch (``hgr_hier_metrics``): per-level
    arg-max over the train columns, chain matching and counting on the device; the per-batch divisions by the chain
    length stay device-side float ops, nothing is copied to the host before the ratios are printed."""

    def __init__(self, model):
        self.model = model
        dev = model.device
        self.hits_all = torch.zeros((), dtype=torch.float64, device=dev)
        self.path_all = torch.zeros((), dtype=torch.float64, device=dev)
        self.point_all = torch.zeros((), dtype=torch.float64, device=dev)
        self.path_all_count = 0
        depth = torch.from_numpy(model.hierarchy.depth).to(torch.int64)
        self.n_levels = int(depth.max()) + 1 if depth.numel() else 1
        # Only the TRAIN columns enter TOR / POR (main.py:143,155,174): the dense logits are taken against the train
        # rows of the bank alone ([B, M], M = 983 of 18,278 classes on ImageNet-21K-D) instead of all N nodes, and the
        # kernel works in train POSITIONS: level per position, the label's chain as positions (-1: not a train class).
        train = model.train_index.cpu()
        dt = depth[train]
        M = dt.numel()
        self._level = dt.to(torch.int8).to(dev)
        self._pos_of = {int(n): j for j, n in enumerate(train.tolist())}
        # first position of train_index that is NOT at level l: where the reference's -1 fill (main.py:171) sits first
        first_out = []
        for l in range(self.n_levels):
            out = (dt != l).nonzero()
            first_out.append(int(out[0]) if out.numel() else M)
        self._first_out = torch.tensor(first_out, dtype=torch.int32, device=dev)
        # fused path (hgr_hier_metrics_fused): the train rows of the bank sorted by level -- stable, so that inside a
        # level the sorted order is the train_index order and "first position among equal values" is preserved
        order = torch.sort(dt, stable=True).indices
        self._sorted_to_pos = order.to(torch.int32).to(dev)
        self._level_end = torch.cumsum(torch.bincount(dt, minlength=self.n_levels), 0).tolist()
        self._bank_sorted = None      # built on first use from model.bank_train (update_classifier must have run)
        self._bank_version = None

    def update(self, logits_train: torch.Tensor, target: int):
        """``logits_train`` [B, M]: cosine logits against ``model.bank_train`` (column j = node train_index[j])."""
        m = self.model
        B = logits_train.shape[0]
        parents = list(m.c2p[target]) + [target]
        L = len(parents)
        dev = logits_train.device
        chain = torch.tensor([self._pos_of.get(p, -1) for p in parents], dtype=torch.int32).to(dev, non_blocking=True)
        chain_level = torch.tensor([len(m.c2p[p]) for p in parents], dtype=torch.int32).to(dev, non_blocking=True)
        counts = torch.zeros(3, dtype=torch.int64, device=dev)
        ops.hier_metrics(logits_train, None, self._level, self.n_levels, self._first_out, chain, chain_level, counts)
        c = counts.double()
        self.hits_all += c[0]                                               # main.py:158-160
        self.path_all += c[2] if L == 1 else c[2] / (L - 1)                 # main.py:179-190
        self.point_all += c[1] / L                                          # main.py:191
        self.path_all_count += B

    def update_fused(self, x_norm: torch.Tensor, target: int):
        """``update`` without the dense logits: ``x_norm`` [B, D] bf16 normalised image features; the per-level arg-max
        over the train classes runs in the epilogue of the GEMM against the level-sorted train bank."""
        m = self.model
        bank = m.bank_train
        stamp = (bank.data_ptr(), bank._version)       # a refreshed bank (new tensor or in-place update) is re-sorted
        if self._bank_sorted is None or self._bank_version != stamp:
            self._bank_sorted = bank[self._sorted_to_pos.long()].contiguous()
            self._bank_version = stamp
        B = x_norm.shape[0]
        parents = list(m.c2p[target]) + [target]
        L = len(parents)
        dev = x_norm.device
        chain = torch.tensor([self._pos_of.get(p, -1) for p in parents], dtype=torch.int32).to(dev, non_blocking=True)
        chain_level = torch.tensor([len(m.c2p[p]) for p in parents], dtype=torch.int32).to(dev, non_blocking=True)
        counts = torch.zeros(3, dtype=torch.int64, device=dev)
        ops.hier_metrics_fused(x_norm, self._bank_sorted, self._level_end, self._sorted_to_pos, self._first_out, chain,
                               chain_level, counts)
        c = counts.double()
        self.hits_all += c[0]
        self.path_all += c[2] if L == 1 else c[2] / (L - 1)
        self.point_all += c[1] / L
        self.path_all_count += B

    def ratios(self, num_sample):
        return (float(self.hits_all) / num_sample * 100.0, float(self.path_all) / self.path_all_count * 100.0,
                float(self.point_all) / num_sample * 100.0)


def test(opts, model, device, splits=None, loader: Optional[Iterable] = None, log: bool = True):
    """Drop-in for main.py:104-222.  ``loader`` yields the reference's batch dicts
    ``{'img': [1,B,...], 'label': [1,B]}`` (single-label batches, imagenet_group_test.py)."""
    print("out", opts.out_ratio)
    print("in", opts.in_ratio)
    model.eval()
    model.update_classifier()
    if loader is None:
        raise ValueError("hgrnet_b200.evaluate.test needs a loader: image I/O (dataset/imagenet_group_test.py) is "
                         "an upstream component; pass loader=DataManager_test(...).get_data_loader()")
    print("Running.", flush=True)
    hier = HierMetrics(model) if getattr(opts, "hgr_hier_metrics", True) else None
    hits = ops.new_hits(model.device)
    num_sample = 0
    out_str = ""
    with torch.no_grad():
        for i, data in enumerate(loader):
            imgs, targets = data["img"].to(device, non_blocking=True)[0], data["label"].to(device, non_blocking=True)[0]
            x = model.encode_image_normalized(imgs)
            model.score_topk(None, targets, hits=hits, feats_normalized=x)   # main.py:135-147, fused
            num_sample += len(targets)
            if hier is not None:
                if getattr(opts, "hgr_hier_dense", False):                   # cross-check path: dense [B, M] logits
                    logits_train = ops.logits_dense(x, model.bank_train)     # clip_tree.py:331, train columns only
                    hier.update(logits_train, int(data["label"][0][0]))
                else:                                                        # per-level arg-max in the GEMM epilogue
                    hier.update_fused(x, int(data["label"][0][0]))
            if i % opts.print_freq == 0:
                out_str = _format(hits, num_sample, hier)
                print(out_str, flush=True)
    print("End of testing.")
    out_str = _format(hits, num_sample, hier)
    print(out_str, flush=True)
    if log:
        with open(model.save_path + "arugements.log", "a") as f:       # [sic] main.py:217
            f.writelines(out_str + "\n")
        with open("{}.txt".format(opts.weights), "a") as f:            # main.py:219-222
            method = "{},{},{}:".format(opts.weights, opts.out_ratio, opts.in_ratio)
            f.writelines(method + "\n" + out_str + "\n")
    return out_str


def _format(hits, num_sample, hier):
    h = hits.tolist()
    hits_dict = dict(zip(HIT_CUTS, h))
    out_str = "\n"
    tmp_str, _ = count_acc(hits_dict, num_sample)
    out_str += tmp_str
    if hier is not None:
        hit_ratio, path_ratio, point_ratio = hier.ratios(num_sample)
        out_str += " hit_ratio(%):{:.2f}".format(hit_ratio)
        out_str += " path_ratio(%):{:.2f}".format(path_ratio)
        out_str += " point_ratio(%):{:.2f}".format(point_ratio)
    return out_str
