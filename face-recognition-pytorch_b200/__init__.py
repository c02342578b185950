"""B200-native margin-softmax head (ArcFace + PartialFC) and pair-verification scorer.

Drop-in for the head of aanna0701/face-recognition-pytorch: Python host code with the reference's module
interface over hand-written sm_100a CUDA reached through the C ABI in include/pfc.h (libpfc_b200.so).
Importing the package loads (building it first if needed) the CUDA library; there is no CPU fallback.
"""
from . import _lib                                    # noqa: F401  (fails loudly if the extension is unavailable)
from .arcface import ArcFace, CosFace, CombinedMarginLoss
from .partial_fc import PartialFC, PartialFCAdamW, shard_range
from .eval import pair_score, cross_score, performance_roc, performance_acc, kfold_accuracy
from .graph import GraphedHeadStep
from .checkpoint import head_shard_state, load_head_shard, reshard

__all__ = ["ArcFace", "CosFace", "CombinedMarginLoss", "PartialFC", "PartialFCAdamW", "shard_range", "pair_score",
           "cross_score", "performance_roc", "performance_acc", "kfold_accuracy", "GraphedHeadStep", "head_shard_state",
           "load_head_shard", "reshard"]
__version__ = "0.1.0"
