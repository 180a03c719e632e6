"""Drive the UNMODIFIED reference (``/root/reference``) on synthetic inputs.  TEST INFRASTRUCTURE.

Used only by ``oracle/gen_golden.py`` (build container) and by the ``-m "not gpu"`` test
that re-validates the restatement when ``/root/reference`` happens to be mounted.  Nothing
on the GPU box reads ``/root/reference``.

The reference cannot be imported as shipped: it needs ``ipdb``, ``ftfy`` and
``nltk.corpus.wordnet`` (not installed), its data JSONs and CLIP weights (not present) and
it parses ``sys.argv`` at import (main.py:70).  All of that is incidental to the arithmetic
of the scoring head, so the harness supplies:

* ``sys.modules`` stubs for ``ipdb`` / ``ftfy`` / ``nltk.corpus.wordnet``;
* a temp working directory holding a synthetic ``graph_edges_cls.json`` /
  ``splits_for_tree.json`` (the reference opens them by relative path, utils.py:40,
  main.py:227);
* ``clip.load`` / ``clip.tokenize`` replaced by a duck-typed table-lookup "CLIP" whose
  ``encode_text`` returns rows of a given text table and whose ``encode_image`` is the
  identity on pre-computed image features -- the encoders are upstream producers, out of
  scope, so the head sees exactly the embeddings we choose;
* for ``main.test``: a fake ``DataManager_test`` yielding feature batches, and on CPU
  ``Tensor.cuda`` -> identity because of the hard-coded ``.cuda()`` at main.py:169.

No reference source is copied; the reference modules are imported from where they lie.
"""
from __future__ import annotations

import contextlib
import io
import json
import os
import sys
import tempfile
import types
from typing import Dict, List, Sequence

import torch
import torch.nn as nn

REFERENCE_ROOT = os.environ.get("HGR_REFERENCE_ROOT", "/root/reference")
_REF_MODULES = ("main", "model", "dataset", "utils", "clip", "data")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "model", "clip_tree.py"))


def wnid_of(i: int) -> str:
    """Synthetic wnid of node ``i`` (offset i+1, so that offset 0 is never used)."""
    return "n%08d" % (i + 1)


class TableCLIP(nn.Module):
    """Duck-typed stand-in for ``clip.model.CLIP`` (clip/model.py:239-368).

    ``encode_text(tokens)`` gathers rows of ``text_table`` by the node id carried in token
    column 0; ``encode_image(x)`` returns ``x * image_gain`` (gain 1, so that the features
    are a non-leaf and ``img_feats.backward`` at clip_tree.py:280 is legal).
    ``logit_scale`` initialises like clip/model.py:291.
    """

    def __init__(self, text_table: torch.Tensor, log_scale: float):
        super().__init__()
        self.text_table = nn.Parameter(text_table.clone().float())
        self.image_gain = nn.Parameter(torch.ones(()))
        self.logit_scale = nn.Parameter(torch.tensor(float(log_scale)))
        self.visual = types.SimpleNamespace(input_resolution=224)

    def encode_text(self, tokens):
        return self.text_table[tokens[:, 0]]

    def encode_image(self, x):
        return x * self.image_gain


def _install_stubs():
    ipdb = types.ModuleType("ipdb")
    ipdb.set_trace = lambda *a, **k: None
    ftfy = types.ModuleType("ftfy")
    ftfy.fix_text = lambda s: s
    nltk = types.ModuleType("nltk")
    corpus = types.ModuleType("nltk.corpus")
    wordnet = types.ModuleType("nltk.corpus.wordnet")

    class _Syn:
        def __init__(self, off):
            self.off = off

        def name(self):
            return "thing_%d.n.01" % self.off

    wordnet.synset_from_pos_and_offset = lambda pos, off: _Syn(off)
    corpus.wordnet = wordnet
    nltk.corpus = corpus
    sys.modules.update({"ipdb": ipdb, "ftfy": ftfy, "nltk": nltk, "nltk.corpus": corpus,
                        "nltk.corpus.wordnet": wordnet})


def _purge_reference_modules():
    for name in list(sys.modules):
        if name.split(".")[0] in _REF_MODULES:
            mod = sys.modules[name]
            f = getattr(mod, "__file__", "") or ""
            if name in _REF_MODULES and not f.startswith(REFERENCE_ROOT) and f:
                continue  # somebody else's module of the same name (e.g. our own main.py)
            del sys.modules[name]


@contextlib.contextmanager
def reference_session(edges: Sequence[Sequence[str]], splits: Dict[str, List[str]],
                      text_table: torch.Tensor, log_scale: float, argv: Sequence[str] = ()):
    """Context manager yielding a namespace with the imported reference modules.

    ``ns.main`` (main.py, argv parsed from ``argv``), ``ns.tree_model`` and ``ns.clip_model``
    (the TableCLIP every ``clip.load`` call returns).
    """
    assert reference_available(), "reference checkout not mounted at %s" % REFERENCE_ROOT
    old_cwd, old_argv, old_path = os.getcwd(), list(sys.argv), list(sys.path)
    saved = {n: m for n, m in sys.modules.items() if n.split(".")[0] in _REF_MODULES}
    for n in saved:
        del sys.modules[n]
    tmp = tempfile.mkdtemp(prefix="hgr_ref_")
    try:
        os.makedirs(os.path.join(tmp, "data", "process_results"))
        with open(os.path.join(tmp, "data", "process_results", "graph_edges_cls.json"), "w") as f:
            json.dump([list(e) for e in edges], f)
        with open(os.path.join(tmp, "data", "process_results", "splits_for_tree.json"), "w") as f:
            json.dump(splits, f)
        os.chdir(tmp)
        sys.argv = ["main.py"] + list(argv)
        sys.path.insert(0, REFERENCE_ROOT)
        _install_stubs()
        import clip as ref_clip  # noqa: the reference's vendored package

        fake = TableCLIP(text_table, log_scale)
        ref_clip.load = lambda name, device=None, download_root=None, **kw: (fake, None)

        def _tokenize(names, context_length=77):
            # prompt is "a photo of a thing <offset>." (clip_tree.py:52-58 + the wordnet stub)
            toks = torch.zeros(len(names), context_length, dtype=torch.long)
            for r, s in enumerate(names):
                toks[r, 0] = int(s.rstrip(".").split(" ")[-1]) - 1
            return toks

        ref_clip.tokenize = _tokenize
        with contextlib.redirect_stdout(io.StringIO()):
            import main as ref_main
        ref_main.opts.device = "cpu"  # main.py:226 builds 'cuda:{device}' itself; tree_model reads opts.device
        ns = types.SimpleNamespace(main=ref_main, tree_model=ref_main.tree_model, clip_model=fake,
                                   tmp=tmp, clip=ref_clip)
        yield ns
    finally:
        os.chdir(old_cwd)
        sys.argv = old_argv
        sys.path[:] = old_path
        _purge_reference_modules()
        for n in ("ipdb", "ftfy", "nltk", "nltk.corpus", "nltk.corpus.wordnet"):
            sys.modules.pop(n, None)
        sys.modules.update(saved)


def build_tree_model(ns, splits, **opt_overrides):
    """Instantiate the reference ``tree_model`` (clip_tree.py:20) on CPU."""
    opts = ns.main.opts
    for k, v in opt_overrides.items():
        setattr(opts, k, v)
    opts.device = "cpu"
    opts.folder = os.path.join(ns.tmp, "out")
    with contextlib.redirect_stdout(io.StringIO()):
        model = ns.tree_model(opts, candidates_train=splits[opts.model_train],
                              candidates_test=splits[opts.model_test])
    if opts.weights == "adaptive":
        # clip_tree.py:74 yields a non-leaf tensor; repeated backward through it fails on
        # torch>=2.  Detaching keeps the value, touches no reference file.
        model.layer_weight = model.layer_weight.detach()
    return model


def run_reference_test_loop(ns, model, batches, splits) -> str:
    """Run the unmodified ``main.test`` (main.py:104-222) over feature batches.

    ``batches``: list of ``(feats [B,D] fp32, label int)`` single-label batches.  Returns the
    final ``Top@1...point_ratio`` line the reference prints and logs.
    """
    ref_main = ns.main

    class _Loader:
        def __init__(self):
            self.batch_sampler = types.SimpleNamespace(num_batch=len(batches))

        def __iter__(self):
            for feats, label in batches:
                yield {"img": feats[None], "label": torch.full((1, feats.shape[0]), label, dtype=torch.long)}

    class _DM:
        def __init__(self, **kw):
            pass

        def get_data_loader(self):
            return _Loader()

    ref_main.DataManager_test = _DM
    old_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self  # main.py:169 hard-codes .cuda()
    buf = io.StringIO()
    cwd = os.getcwd()
    try:
        with contextlib.redirect_stdout(buf):
            ref_main.test(ref_main.opts, model, "cpu", splits)
    finally:
        torch.Tensor.cuda = old_cuda
        os.chdir(cwd)
    lines = [l for l in buf.getvalue().splitlines() if l.startswith("Top@1")]
    return lines[-1]
