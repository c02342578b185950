"""CPU oracle for the margin-softmax head (ArcFace + PartialFC).  TEST INFRASTRUCTURE ONLY.

This is a restatement of the reference algorithm in plain torch-on-CPU tensor ops (fp64 by default, fp32 when
used as the timed CPU baseline), written as closed-form functions rather than autograd modules.  It exists to
check the CUDA path; nothing in the product package may import it (only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs do).

Pinning: tests/test_oracle_golden.py checks every function here against fixtures under tests/golden/ that were
produced by running the UNMODIFIED reference modules imported from /root/reference (tests/golden/make_golden.py:
gloo world sizes 1 and 2, sample_rate 1.0 and < 1, SGD steps, scorer).  Behaviour the reference leaves undefined
(margin derivative at |t| = 1, top-k ties straddling the k-th value) is unpinned and documented in DESIGN.md.

Each function cites the reference lines it follows (paths relative to the reference root).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import torch


# ----------------------------------------------------------------------------------------------- shard arithmetic
def shard_range(num_classes: int, rank: int, world_size: int) -> Tuple[int, int]:
    """(num_local, class_start) of a rank's contiguous class shard.  nets/PartialFC.py:57-62."""
    num_local = num_classes // world_size + int(rank < num_classes % world_size)
    class_start = num_classes // world_size * rank + min(rank, num_classes % world_size)
    return num_local, class_start


def num_sample(sample_rate: float, num_local: int) -> int:
    """nets/PartialFC.py:63."""
    return int(sample_rate * num_local)


def localize_labels(labels: torch.Tensor, class_start: int, num_local: int) -> torch.Tensor:
    """Global labels -> shard-local ids, -1 where another rank owns the class.  nets/PartialFC.py:188-193."""
    labels = labels.reshape(-1).long()
    owned = (labels >= class_start) & (labels < class_start + num_local)
    return torch.where(owned, labels - class_start, torch.full_like(labels, -1))


# ----------------------------------------------------------------------------------------------- sampling
def sample_indices(perm: torch.Tensor, labels_local: torch.Tensor, n_sample: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Negative-class sampling.  nets/PartialFC.py:108-118.

    perm: the uniform draw [num_local] the reference obtains from torch.rand (:110).
    Returns (index ascending int64, labels remapped into the index list, -1 kept).
    Ties at the k-th value: lowest index first (the reference leaves this to torch.topk).
    """
    labels_local = labels_local.reshape(-1).long()
    owned = labels_local >= 0
    positive = torch.unique(labels_local[owned], sorted=True)
    if n_sample - positive.numel() >= 0:
        p = perm.clone().float()
        p[positive] = 2.0
        # k largest with the lowest-index tie rule == stable descending sort
        order = torch.sort(p, descending=True, stable=True).indices[:n_sample]
        index = torch.sort(order).values
    else:
        index = positive
    remapped = labels_local.clone()
    remapped[owned] = torch.searchsorted(index, labels_local[owned])
    return index, remapped


# ----------------------------------------------------------------------------------------------- margins
@dataclass
class Margin:
    """kind 'arcface': nets/ArcFace.py:63-91 (and CombinedMarginLoss m1 == 1, m3 == 0, :42-52);
    kind 'cosface': nets/ArcFace.py:94-106 (and CombinedMarginLoss m3 > 0, :54-57)."""
    kind: str = "arcface"
    s: float = 64.0
    m: float = 0.5
    filter_thr: float = 0.0   # CombinedMarginLoss.interclass_filtering_threshold, :30-38

    def consts(self):
        return (math.cos(self.m), math.sin(self.m), math.cos(math.pi - self.m), math.sin(math.pi - self.m) * self.m)

    def apply(self, t: torch.Tensor) -> torch.Tensor:
        """final target cosine as a function of the (clamped) target cosine t."""
        if self.kind == "cosface":
            return t - self.m
        cos_m, sin_m, theta, sinmm = self.consts()
        sin_t = torch.sqrt(1.0 - t * t)
        return torch.where(t > theta, t * cos_m - sin_t * sin_m, t - sinmm)

    def derivative(self, t: torch.Tensor) -> torch.Tensor:
        """d final / d t -- what autograd produces through nets/ArcFace.py:80-87 for |t| < 1."""
        if self.kind == "cosface":
            return torch.ones_like(t)
        cos_m, sin_m, theta, _ = self.consts()
        d = cos_m + sin_m * t / torch.sqrt(torch.clamp(1.0 - t * t, min=1e-12))
        return torch.where(t > theta, d, torch.ones_like(t))


def normalize_rows(x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """F.normalize (nets/PartialFC.py:199-200): x / max(||x||, 1e-12); also returns 1/max(||x||, 1e-12)."""
    denom = torch.clamp(torch.linalg.vector_norm(x, dim=1, keepdim=True), min=1e-12)
    return x / denom, (1.0 / denom).reshape(-1)


def normalize_backward(g: torch.Tensor, xn: torch.Tensor, inv_norm: torch.Tensor) -> torch.Tensor:
    """Gradient of F.normalize: (g - xn (xn . g)) / ||x||."""
    return (g - xn * (xn * g).sum(dim=1, keepdim=True)) * inv_norm.reshape(-1, 1)


# ----------------------------------------------------------------------------------------------- one rank, one step
@dataclass
class RankForward:
    """Everything one rank produces in the forward pass before the collectives."""
    xn: torch.Tensor            # [B, d] normalised global batch
    wn: torch.Tensor            # [n, d] normalised active classes
    inv_w: torch.Tensor         # [n]
    raw: torch.Tensor           # [B, n] unclamped cosines
    z: torch.Tensor             # [B, n] scaled logits after clamp + margin (+ filtering)
    labels: torch.Tensor        # [B] local (remapped) labels, -1 elsewhere
    grad_gate: torch.Tensor     # [B, n] d z / d raw  (s * clamp gate * margin derivative on the target * filter)


def rank_logits(xn: torch.Tensor, w_active: torch.Tensor, labels: torch.Tensor, margin: Margin) -> RankForward:
    """normalise -> cosine GEMM -> clamp -> margin -> scale.  nets/PartialFC.py:199-206, nets/ArcFace.py:76-91."""
    wn, inv_w = normalize_rows(w_active)
    raw = xn @ wn.t()
    cl = raw.clamp(-1.0, 1.0)
    gate = ((raw >= -1.0) & (raw <= 1.0)).to(raw.dtype)            # clamp backward
    rows = torch.nonzero(labels >= 0).reshape(-1)
    cols = labels[rows]
    if margin.filter_thr > 0:                                        # nets/ArcFace.py:30-38
        dirty = cl > margin.filter_thr
        dirty[rows, cols] = False
        cl = torch.where(dirty, torch.zeros_like(cl), cl)
        gate = torch.where(dirty, torch.zeros_like(gate), gate)
    t = cl[rows, cols]
    z = cl.clone()
    z[rows, cols] = margin.apply(t)
    z = z * margin.s
    gg = gate * margin.s
    gg[rows, cols] = gg[rows, cols] * margin.derivative(t)
    return RankForward(xn=xn, wn=wn, inv_w=inv_w, raw=raw, z=z, labels=labels, grad_gate=gg)


def dist_cross_entropy(zs: Sequence[torch.Tensor], labels_per_rank: Sequence[torch.Tensor]):
    """Model-parallel softmax cross-entropy over class shards.  nets/PartialFC.py:442-461.

    Returns (loss, [p_r]) with p_r the per-rank probability blocks the reference saves for backward.
    """
    B = zs[0].shape[0]
    gmax = torch.stack([z.max(dim=1).values for z in zs]).max(dim=0).values.reshape(-1, 1)   # all_reduce MAX
    es = [torch.exp(z - gmax) for z in zs]
    denom = sum(e.sum(dim=1, keepdim=True) for e in es)                                      # all_reduce SUM
    ps = [e / denom for e in es]
    pt = torch.zeros(B, 1, dtype=zs[0].dtype)
    for p, lab in zip(ps, labels_per_rank):                                                  # all_reduce SUM
        rows = torch.nonzero(lab >= 0).reshape(-1)
        pt[rows, 0] = p[rows, lab[rows]]
    loss = -(pt.clamp_min(1e-30).log().mean())
    return loss, ps


@dataclass
class StepResult:
    loss: torch.Tensor
    dx_local: List[torch.Tensor]            # per rank [b, d]: world_size * d loss / d local_embeddings
    dw: List[torch.Tensor]                  # per rank [n_r, d]: d loss / d weight_activated
    index: List[Optional[torch.Tensor]]     # per rank sampled class index (None when sample_rate == 1)
    labels_local: List[torch.Tensor]        # per rank remapped labels [B]


def head_step(local_embeddings: Sequence[torch.Tensor], local_labels: Sequence[torch.Tensor],
              weights: Sequence[torch.Tensor], num_classes: int, margin: Margin, sample_rate: float = 1.0,
              perms: Optional[Sequence[torch.Tensor]] = None, grad_scale: float = 1.0,
              dtype: torch.dtype = torch.float64) -> StepResult:
    """One PartialFC forward + backward for a whole world, simulated rank by rank in one process.

    local_embeddings[r]: [b, d]; local_labels[r]: [b] global ids; weights[r]: the rank's FULL shard [num_local_r, d].
    Follows PartialFC.forward (nets/PartialFC.py:146-208) and the backward of DistCrossEntropyFunc (:464-484),
    ArcFace, clamp, linear, normalize and AllGatherFunc (:505-522) in closed form.
    """
    W = len(weights)
    x = torch.cat([e.to(dtype) for e in local_embeddings])            # all_gather, rank order (:182-186)
    labels = torch.cat([l.reshape(-1).long() for l in local_labels])
    B, b = x.shape[0], local_embeddings[0].shape[0]
    xn, inv_x = normalize_rows(x)
    fwd: List[RankForward] = []
    idxs: List[Optional[torch.Tensor]] = []
    for r in range(W):
        nl, cs = shard_range(num_classes, r, W)
        assert weights[r].shape[0] == nl
        lab = localize_labels(labels, cs, nl)
        w_full = weights[r].to(dtype)
        if sample_rate < 1:
            index, lab = sample_indices(perms[r], lab, num_sample(sample_rate, nl))
            w_act = w_full[index]                                     # :120
            idxs.append(index)
        else:
            w_act = w_full
            idxs.append(None)
        fwd.append(rank_logits(xn, w_act, lab, margin))
    loss, ps = dist_cross_entropy([f.z for f in fwd], [f.labels for f in fwd])
    dxn = torch.zeros_like(xn)
    dws = []
    for f, p in zip(fwd, ps):
        dz = p.clone()                                                # (p - onehot) / B * g   (:478-484)
        rows = torch.nonzero(f.labels >= 0).reshape(-1)
        dz[rows, f.labels[rows]] -= 1.0
        dz = dz / B * grad_scale
        draw = dz * f.grad_gate                                       # margin, scale, clamp backward
        dxn = dxn + draw @ f.wn                                       # reduce over ranks (:510-519)
        dwn = draw.t() @ xn
        dws.append(normalize_backward(dwn, f.wn, f.inv_w))
    dx = normalize_backward(dxn, xn, inv_x) * W                       # grad_out *= len(grad_list)   (:521)
    dx_local = [dx[r * b:(r + 1) * b] for r in range(W)]
    return StepResult(loss=loss, dx_local=dx_local, dw=dws, index=idxs, labels_local=[f.labels for f in fwd])


def sgd_update(w: torch.Tensor, buf: torch.Tensor, grad: torch.Tensor, lr: float, momentum: float,
               weight_decay: float) -> Tuple[torch.Tensor, torch.Tensor]:
    """torch.optim.SGD (dampening 0, no nesterov) as driven by model/FR_PartialFC.py:188; zero buffer == first step."""
    g = grad + weight_decay * w
    buf = momentum * buf + g if momentum != 0 else g
    return w - lr * buf, buf


def adamw_update(w, exp_avg, exp_avg_sq, grad, step, lr, beta1, beta2, eps, weight_decay, decoupled=True):
    """torch.optim.AdamW / Adam single-tensor update."""
    if decoupled:
        w = w * (1 - lr * weight_decay)
    else:
        grad = grad + weight_decay * w
    exp_avg = beta1 * exp_avg + (1 - beta1) * grad
    exp_avg_sq = beta2 * exp_avg_sq + (1 - beta2) * grad * grad
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    denom = exp_avg_sq.sqrt() / math.sqrt(bc2) + eps
    return w - (lr / bc1) * exp_avg / denom, exp_avg, exp_avg_sq


class PartialFCOracle:
    """Stateful single-process restatement of PartialFC + SGD over several steps for a whole world.

    Mirrors the reference's buffer semantics: with sample_rate < 1 each step gathers the sampled rows
    (nets/PartialFC.py:120-121), the optimizer updates only those, and they are scattered back at the start of the
    NEXT forward (update(), :133-143) -- so `weight` lags by one step, exactly like the reference.
    """

    def __init__(self, weights: Sequence[torch.Tensor], num_classes: int, margin: Margin, sample_rate: float,
                 lr: float, momentum: float, weight_decay: float, dtype=torch.float64, optimizer: str = "sgd",
                 betas=(0.9, 0.999), eps: float = 1e-8):
        """optimizer "sgd": PartialFC + torch.optim.SGD; "adamw" / "adam": PartialFCAdamW + torch.optim.AdamW / Adam
        (nets/PartialFC.py:235-432, :320), `mom` then holds exp_avg and `mom2` exp_avg_sq."""
        self.W = len(weights)
        self.num_classes, self.margin, self.sample_rate = num_classes, margin, sample_rate
        self.lr, self.momentum, self.wd, self.dtype = lr, momentum, weight_decay, dtype
        self.optimizer, self.betas, self.eps = optimizer, betas, eps
        self.t = 0            # forward calls so far == PartialFCAdamW.step (:306)
        self.weight = [w.clone().to(dtype) for w in weights]
        self.mom = [torch.zeros_like(w) for w in self.weight]
        self.mom2 = [torch.zeros_like(w) for w in self.weight]
        self.pending = None   # (index per rank, activated weights, activated state...) awaiting scatter-back

    def _flush(self):
        if self.pending is None:
            return
        for r, (idx, w_act, m_act, v_act) in enumerate(self.pending):
            if idx is None:
                self.weight[r], self.mom[r] = w_act, m_act
                if v_act is not None:
                    self.mom2[r] = v_act
            else:
                self.weight[r][idx] = w_act
                self.mom[r][idx] = m_act
                if v_act is not None:
                    self.mom2[r][idx] = v_act
        self.pending = None

    def step(self, local_embeddings, local_labels, perms=None) -> StepResult:
        self._flush()                                                          # update() (:166)
        res = head_step(local_embeddings, local_labels, self.weight, self.num_classes, self.margin,
                        self.sample_rate, perms, dtype=self.dtype)
        self.t += 1
        pend = []
        for r in range(self.W):
            idx = res.index[r]
            w_act = self.weight[r] if idx is None else self.weight[r][idx]
            m_act = self.mom[r] if idx is None else self.mom[r][idx]
            if self.optimizer in ("adamw", "adam"):
                v_act = self.mom2[r] if idx is None else self.mom2[r][idx]
                # Sampled: sample() writes state["step"] = self.step (= t) BEFORE optimizer.step(), which increments it
                # once more -- the bias correction of forward call t uses t + 1 (nets/PartialFC.py:306, :327 + torch's
                # `step += 1`).  Full (sample_rate == 1): sample() never runs and the optimizer counts by itself: t.
                step = self.t + 1 if idx is not None else self.t
                w_new, m_new, v_new = adamw_update(w_act, m_act, v_act, res.dw[r], step, self.lr, self.betas[0],
                                                   self.betas[1], self.eps, self.wd,
                                                   decoupled=self.optimizer == "adamw")
                pend.append((idx, w_new, m_new, v_new))
            else:
                w_new, m_new = sgd_update(w_act, m_act, res.dw[r], self.lr, self.momentum, self.wd)
                pend.append((idx, w_new, m_new, None))
        self.pending = pend
        return res

    def full_weights(self):
        self._flush()
        return self.weight, self.mom

    def full_adam_state(self):
        self._flush()
        return self.mom, self.mom2


# ----------------------------------------------------------------------------------------------- timed CPU baseline
def cpu_reference_step(x: torch.Tensor, labels: torch.Tensor, w: torch.nn.Parameter, opt: torch.optim.Optimizer,
                       margin: Margin) -> float:
    """The reference head's op sequence on one rank, autograd and all, for timing on the host cores
    (bench.py cpu_baseline / --impl reference).  Same ops in the same order as nets/PartialFC.py:199-207,
    nets/ArcFace.py:77-90, nets/PartialFC.py:446-461 and :478-484, then optimizer.step()."""
    opt.zero_grad(set_to_none=True)
    lab = labels.view(-1, 1)
    xn = torch.nn.functional.normalize(x)
    wn = torch.nn.functional.normalize(w)
    logits = torch.nn.functional.linear(xn, wn).clamp(-1, 1)
    index = torch.where(lab != -1)[0]
    tl = logits[index, lab[index].view(-1)]
    cos_m, sin_m, theta, sinmm = margin.consts()
    sin_theta = torch.sqrt(1.0 - torch.pow(tl, 2))
    ctm = tl * cos_m - sin_theta * sin_m
    final = torch.where(tl > theta, ctm, tl - sinmm)
    logits[index, lab[index].view(-1)] = final
    logits = logits * margin.s
    loss = _DistCE.apply(logits, lab)
    loss.backward()
    opt.step()
    return float(loss.detach())


class _DistCE(torch.autograd.Function):
    """World-size-1 restatement of DistCrossEntropyFunc (nets/PartialFC.py:435-484) used by cpu_reference_step."""

    @staticmethod
    def forward(ctx, logits, label):
        B = logits.size(0)
        mx, _ = torch.max(logits, dim=1, keepdim=True)
        logits = logits - mx
        logits.exp_()
        logits.div_(logits.sum(dim=1, keepdim=True))
        index = torch.where(label != -1)[0]
        loss = torch.zeros(B, 1, dtype=logits.dtype)
        loss[index] = logits[index].gather(1, label[index])
        ctx.save_for_backward(index, logits, label)
        return loss.clamp_min_(1e-30).log_().mean() * (-1)

    @staticmethod
    def backward(ctx, g):
        index, p, label = ctx.saved_tensors
        B = p.size(0)
        one_hot = torch.zeros(index.size(0), p.size(1), dtype=p.dtype)
        one_hot.scatter_(1, label[index], 1)
        p[index] -= one_hot
        p.div_(B)
        return p * g.item(), None
