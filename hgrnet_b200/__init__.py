"""hgrnet_b200 -- B200-native (sm_100a) hierarchical zero-shot scoring head of HGR-Net.

Public surface mirrors the reference's ``tree_model`` / ``main.py``;
the compute lives in ``libhgr_b200.so`` behind the C ABI of ``include/hgr_b200.h``.
"""
from .hierarchy import Hierarchy, synthetic_hierarchy, synthetic_tree_edges, scaled_levels, WORDNET_LIKE_21841  # noqa: F401
from .levels import level_weights, layer_weight_init  # noqa: F401
from .flags import build_parser, parse_args  # noqa: F401

__all__ = ["Hierarchy", "synthetic_hierarchy", "synthetic_tree_edges", "scaled_levels", "level_weights",
           "layer_weight_init", "build_parser", "parse_args", "tree_model", "ops"]


def __getattr__(name):  # torch-facing pieces load lazily so that `import hgrnet_b200` stays cheap
    if name in ("tree_model",):
        from .head import tree_model
        return tree_model
    if name in ("ops", "evaluate", "synthetic", "dist", "head", "_cabi"):
        import importlib
        return importlib.import_module("." + name, __name__)
    raise AttributeError(name)
