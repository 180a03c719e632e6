"""Generate the golden fixtures under tests/golden/ FROM THE REFERENCE ITSELF.  TEST INFRASTRUCTURE.

Run in the build container (where /root/reference is mounted):

    python -m oracle.gen_golden

For every case it (1) regenerates seeded synthetic inputs, (2) runs the unmodified reference
code through ``oracle/ref_harness.py``, (3) asserts that the fp32 restatement in
``oracle/hgr_oracle.py`` reproduces the reference (this is what pins the oracle), and
(4) freezes the reference's outputs.  Inputs are never stored -- tests regenerate them from
the seeds recorded in each fixture with ``oracle.cases``.
"""
from __future__ import annotations

import json
import os
import random
import sys

import numpy as np
import torch

from . import cases, hgr_oracle as orc, ref_harness as rh

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _splits(nodes, test_ids):
    return {"train": nodes, "rest": [nodes[i] for i in test_ids], "all": nodes}


def gen_tree_case():
    """gen_tree on the quirky edge list (d2n key order != depth order) and on a 3-level tree."""
    out = {}
    for name, edges in (("quirky", cases.QUIRKY_EDGES), ("tree_4_20_200", cases.tree_edges([4, 20, 200], 3))):
        nodes = orc.gen_tree(edges)[3]
        with rh.reference_session(edges, _splits(nodes, range(len(nodes))), torch.zeros(len(nodes), 8), 0.0) as ns:
            import utils as ref_utils  # the reference's utils.py
            opts = ns.main.opts
            p2c, c2p, d2n, rnodes, start_up = ref_utils.gen_tree(opts)
        o = orc.gen_tree(edges)
        assert (o[0], o[1], dict(o[2]), o[3], o[4]) == (p2c, c2p, dict(d2n), rnodes, start_up), name
        assert list(o[2].keys()) == list(d2n.keys()), name
        out[name] = {"p2c": p2c, "c2p": c2p, "d2n_keys": list(d2n.keys()), "d2n": {str(k): v for k, v in d2n.items()},
                     "nodes": rnodes, "start_up": start_up}
    return out


def flags_case():
    """Flag surface of the reference's main.py (main.py:14-68): dest -> [option strings, default, nargs/type tag]."""
    edges = cases.QUIRKY_EDGES
    nodes = orc.gen_tree(edges)[3]
    out = {}
    with rh.reference_session(edges, _splits(nodes, range(len(nodes))), torch.zeros(len(nodes), 8), 0.0) as ns:
        for a in ns.main.parser._actions:
            if a.dest == "help":
                continue
            kind = "flag" if a.nargs == 0 else getattr(a.type, "__name__", str(a.type))
            out[a.dest] = {"opts": list(a.option_strings), "default": a.default, "kind": kind}
    return out


def weights_case():
    """get_weights known answers for every method (clip_tree.py:198-219) incl. adaptive on (4,20,200)."""
    edges = cases.tree_edges([4, 20, 200], 3)
    nodes = orc.gen_tree(edges)[3]
    splits = _splits(nodes, range(24, 224))
    out = {}
    with rh.reference_session(edges, splits, torch.zeros(len(nodes), 8), 0.0) as ns:
        model = rh.build_tree_model(ns, splits, weights="adaptive", scale=1.0)
        for method in ("equal", "increasing", "decreasing", "adaptive", "nl_increasing", "nl_decreasing"):
            for n in (1, 2, 3):
                ref = model.get_weights(method, n).float()
                lw = orc.layer_weight_init(orc.gen_tree(edges)[2], 1.0)
                mine = orc.get_weights(method, n, lw).float()
                assert torch.allclose(ref, mine, rtol=0, atol=0), (method, n)
                out["%s_%d" % (method, n)] = ref.tolist()
        out["layer_weight"] = model.layer_weight.tolist()
    return out


def eval_case(spec):
    """update_classifier + forward + the unmodified main.test loop (Top@k / hit / path / point ratios)."""
    edges = cases.tree_edges(spec["levels"], spec["tree_seed"])
    p2c, c2p, d2n, nodes, _ = orc.gen_tree(edges)
    test_ids = cases.test_ids(spec, len(nodes))
    splits = _splits(nodes, test_ids)
    table = cases.text_table(spec, len(nodes))
    batches = cases.eval_batches(spec, test_ids)
    res = {}
    with rh.reference_session(edges, splits, table, float(np.log(1 / 0.07)), argv=["--train", "False"]) as ns:
        model = rh.build_tree_model(ns, splits, weights="equal")
        model.update_classifier()
        bank = model.zsl_weights.detach().clone()
        with torch.no_grad():
            logits0 = model(batches[0][0].clone(), None).clone()
        line = rh.run_reference_test_loop(ns, model, [(f.clone(), l) for f, l in batches], splits)
    # ---- pin the restatement on the reference
    obank = orc.normalize_rows(table)
    assert torch.equal(obank, bank)
    assert torch.equal(orc.forward_logits(batches[0][0], obank), logits0)
    test_index = torch.tensor(test_ids)
    train_index = torch.arange(len(nodes))
    hits = {k: 0 for k in orc.TOPK}
    tor = path = point = 0.0
    n = 0
    preds, vals = [], []
    for feats, label in batches:
        lg = orc.forward_logits(feats, obank)
        tg = torch.full((feats.shape[0],), label, dtype=torch.long)
        pred, val, h = orc.eval_hits(lg, test_index, tg)
        preds.append(pred.t().contiguous().numpy().astype(np.int32))
        vals.append(val.numpy())
        for k in hits:
            hits[k] += h[k]
        a, b, c = orc.tor_por(lg, train_index, c2p, d2n, len(nodes), label)
        tor, path, point = tor + a, path + b, point + c
        n += feats.shape[0]
    s, _ = orc.count_acc(hits, n)
    mine = s + " hit_ratio(%):{:.2f}".format(tor / n * 100.0) + " path_ratio(%):{:.2f}".format(path / n * 100.0) \
        + " point_ratio(%):{:.2f}".format(point / n * 100.0)
    assert mine == line, (mine, line)
    res["line"] = line
    res["hits"] = {str(k): v for k, v in hits.items()}
    res["num_sample"] = n
    np.savez_compressed(os.path.join(GOLDEN, spec["name"] + ".npz"),
                        pred=np.stack(preds), val=np.stack(vals),
                        bank_rows=bank[:: max(1, len(nodes) // 16)].numpy(),
                        logits0_rows=logits0[:4].numpy())
    return res


def om_case(spec):
    """One OM training step through the unmodified tree_model.train_batch (clip_tree.py:222-281)."""
    edges = cases.tree_edges(spec["levels"], spec["tree_seed"])
    p2c, c2p, d2n, nodes, _ = orc.gen_tree(edges)
    splits = _splits(nodes, cases.test_ids(spec, len(nodes)))
    table = cases.text_table(spec, len(nodes), normalize=False)
    img = cases.image_feats(spec)
    target = spec["target"]
    log_scale = float(np.log(1 / 0.07))
    o = spec["opts"]
    with rh.reference_session(edges, splits, table, log_scale) as ns:
        model = rh.build_tree_model(ns, splits, **o)
        rec = {"ids": [], "ce": []}
        orig_contra = model.get_contra

        def contra(*a, **kw):
            ci, lab = orig_contra(*a, **kw)
            rec["ids"].append(ci.tolist())
            return ci, lab

        model.get_contra = contra
        ce = model.loss

        class _RecordingLoss(torch.nn.Module):
            def forward(self, lg, lab):
                v = ce(lg, lab)
                rec["ce"].append(float(v.detach()))
                return v

        model.loss = _RecordingLoss()
        x = img.clone().requires_grad_(True)
        random.seed(spec["sample_seed"])
        loss = model.train_batch(x, torch.full((img.shape[0],), target, dtype=torch.long), o.get("training_method", "OM"),
                                 o.get("sample_strategy", "topk"))
        d_text = ns.clip_model.text_table.grad.clone()
        d_ls = ns.clip_model.logit_scale.grad.clone()
        d_x = x.grad.clone()
        lw = model.layer_weight.clone() if o.get("weights") == "adaptive" else None
    # ---- pin the restatement
    random.seed(spec["sample_seed"])
    mine = orc.om_step(img, table, torch.tensor(log_scale), c2p, d2n, target, out_ratio=o["out_ratio"],
                       in_ratio=o["in_ratio"], weights=o["weights"], weighting=o.get("weighting", "both"),
                       k=o.get("k", 1), num_compare=o.get("num_compare", 256), layer_weight=lw)
    assert mine["compare_idx"] == rec["ids"]
    assert abs(mine["loss"] - loss) <= 1e-6 * abs(loss), (mine["loss"], loss)
    assert torch.allclose(mine["d_text_raw"], d_text, rtol=1e-5, atol=1e-8)
    assert torch.allclose(mine["d_log_scale"], d_ls, rtol=1e-5, atol=1e-8)
    # d_x of the reference is the gradient w.r.t. the RAW image features (through the normalisation)
    xn = img / img.norm(dim=-1, keepdim=True)
    g = mine["d_img_n"]
    d_raw = (g - xn * (xn * g).sum(-1, keepdim=True)) / img.norm(dim=-1, keepdim=True)
    assert torch.allclose(d_raw, d_x, rtol=1e-4, atol=1e-7)
    rows = d_text.abs().sum(1).nonzero().squeeze(1)
    np.savez_compressed(os.path.join(GOLDEN, spec["name"] + ".npz"), d_x=d_x.numpy(), d_text_rows=rows.numpy(),
                        d_text=d_text[rows].numpy(), d_log_scale=d_ls.numpy())
    return {"loss": loss, "losses": mine["losses"], "ce": rec["ce"], "compare_idx": rec["ids"],
            "labels": mine["labels"], "weights": mine["weights"], "T": len(rec["ids"])}


def main():
    assert rh.reference_available(), "run in the build container with /root/reference mounted"
    os.makedirs(GOLDEN, exist_ok=True)
    torch.set_num_threads(1)  # bit-reproducible CPU reductions
    meta = {"torch": torch.__version__, "generator": "oracle/gen_golden.py"}
    meta["gen_tree"] = gen_tree_case()
    meta["get_weights"] = weights_case()
    meta["flags"] = flags_case()
    meta["eval"] = {s["name"]: eval_case(s) for s in cases.EVAL_CASES}
    meta["om"] = {s["name"]: om_case(s) for s in cases.OM_CASES}
    with open(os.path.join(GOLDEN, "golden.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    print("wrote", GOLDEN, sorted(os.listdir(GOLDEN)))


if __name__ == "__main__":
    sys.exit(main())
