"""Margin modules with the reference's constructor signatures and attribute names (nets/ArcFace.py).

Inside PartialFC the margin is not a separate pass: the head reads `margin_spec()` and the forward GEMM's
epilogue applies it on the target column.  Called on their own -- `margin(logits, labels) -> logits`, the
contract of nets/ArcFace.py:76 -- they run the stand-alone CUDA kernel pfc_margin_apply.
"""
import math

import torch

from . import kernels as K


class _MarginApply(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, labels, spec):
        kind, s, m2, m3, thr = spec
        logits = logits.contiguous()
        lab = labels.reshape(-1).to(torch.int64).contiguous()
        out = torch.empty_like(logits)
        gate = torch.empty_like(logits)
        K.margin_apply(logits, lab, kind, s, m2, m3, thr, out, gate)
        ctx.save_for_backward(gate)
        return out

    @staticmethod
    def backward(ctx, g):
        (gate,) = ctx.saved_tensors
        return g * gate, None, None


class _Margin(torch.nn.Module):
    def margin_spec(self):
        """(kind, s, m2, m3, interclass_filtering_threshold) consumed by the fused head."""
        raise NotImplementedError

    def forward(self, logits: torch.Tensor, labels: torch.Tensor):
        if logits.dtype != torch.float32:
            raise TypeError("margin modules take fp32 logits (nets/PartialFC.py:202-204 casts before the margin)")
        return _MarginApply.apply(logits, labels, self.margin_spec())


class ArcFace(_Margin):
    """ArcFace additive angular margin, nets/ArcFace.py:63-91: cos(theta + m) on the target, then * s."""

    def __init__(self, s=64.0, margin=0.5):
        super().__init__()
        self.scale = s
        self.margin = margin
        self.cos_m = math.cos(margin)
        self.sin_m = math.sin(margin)
        self.theta = math.cos(math.pi - margin)
        self.sinmm = math.sin(math.pi - margin) * margin
        self.easy_margin = False

    def margin_spec(self):
        if self.easy_margin:
            raise NotImplementedError("easy_margin=True is not wired in the reference either (nets/ArcFace.py:73)")
        return (K.MARGIN_ARCFACE, float(self.scale), float(self.margin), 0.0, 0.0)


class CosFace(_Margin):
    """CosFace additive cosine margin, nets/ArcFace.py:94-106: target - m, then * s."""

    def __init__(self, s=64.0, m=0.40):
        super().__init__()
        self.s = s
        self.m = m

    def margin_spec(self):
        return (K.MARGIN_COSFACE, float(self.s), 0.0, float(self.m), 0.0)


class CombinedMarginLoss(_Margin):
    """nets/ArcFace.py:5-61: s*(cos(m1*theta + m2) - m3) for the two cases the reference implements
    (m1 == 1 and m3 == 0 -> ArcFace with m2;  m3 > 0 -> CosFace with m3) plus inter-class filtering."""

    def __init__(self, s, m1, m2, m3, interclass_filtering_threshold=0):
        super().__init__()
        self.s = s
        self.m1 = m1
        self.m2 = m2
        self.m3 = m3
        self.interclass_filtering_threshold = interclass_filtering_threshold
        self.cos_m = math.cos(self.m2)
        self.sin_m = math.sin(self.m2)
        self.theta = math.cos(math.pi - self.m2)
        self.sinmm = math.sin(math.pi - self.m2) * self.m2
        self.easy_margin = False

    def margin_spec(self):
        thr = float(self.interclass_filtering_threshold)
        if self.m1 == 1.0 and self.m3 == 0.0:
            return (K.MARGIN_ARCFACE, float(self.s), float(self.m2), 0.0, thr)
        if self.m3 > 0:
            return (K.MARGIN_COSFACE, float(self.s), 0.0, float(self.m3), thr)
        raise RuntimeError("unsupported CombinedMarginLoss parameters (the reference raises here too, nets/ArcFace.py:58-59)")
