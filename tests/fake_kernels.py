"""CPU stand-in for face_recognition_pytorch_b200.kernels used ONLY by tests/test_dist_gloo.py.

It implements the contract of include/pfc.h on CPU tensors with torch ops (and the oracle's sampler), so that the
HOST logic of the head -- shard arithmetic, all-gather order, label localisation, the single [B,2] statistics
all-reduce, the dX reduce-scatter and its x world_size, optimizer patching, update()/scatter-back -- can run
under gloo with world_size 2 on a machine without a GPU.  The product never imports this file.
"""
import math

import torch

from oracle import head_oracle as ho

LOG2E = 1.4426950408889634


class FakeKernels:
    MARGIN_ARCFACE, MARGIN_COSFACE = 0, 1

    def __init__(self, real):
        self._real = real
        # host-only helpers come from the real library (no GPU needed)
        for name in ("exp_top", "padded_classes", "padded_batch", "num_class_tiles", "part_sum_cols", "dx_splits",
                     "dx_max_splits", "sample_workspace_bytes", "hist_bins"):
            setattr(self, name, getattr(real, name))

    # ---- rows
    def l2norm_rows(self, x, index, rows, xn, inv_norm):
        src = x[index[:rows]] if index is not None else x[:rows]
        denom = src.norm(dim=1, keepdim=True).clamp_min(1e-12)
        xn[:rows] = (src / denom).to(xn.dtype)
        inv_norm[:rows] = (1.0 / denom).reshape(-1)

    def l2norm_rows_localize(self, x, rows, xn, inv_norm, labels, class_start, num_local, labels_local):
        self.l2norm_rows(x, None, rows, xn, inv_norm)
        self.localize_labels(labels, class_start, num_local, labels_local)

    def row_stats_loss(self, part_sum, n_tiles, B, labels, tgt_e, stats, row_L, out, ticket):
        self.row_stats(part_sum, n_tiles, B, labels, tgt_e, stats)
        self.loss(stats, B, row_L, out)

    def row_stats_loss_prepare(self, part_sum, n_tiles, B, labels, tgt_e, stats, row_L, out, ticket, grad_loss, s, d,
                               tgt_raw, kind, m2, xn, xs, coef, E, n_pad):
        self.row_stats_loss(part_sum, n_tiles, B, labels, tgt_e, stats, row_L, out, ticket)
        self.backward_prepare(stats, row_L, grad_loss, s, B, d, labels, tgt_raw, kind, m2, xn, xs, coef, E, n_pad)

    def localize_labels(self, labels, class_start, num_local, out):
        out.copy_(ho.localize_labels(labels, class_start, num_local).to(torch.int32))

    def sample(self, perm, labels_local, num_local, num_sample, index_out, n_out, labels_remapped, workspace):
        index, remapped = ho.sample_indices(perm, labels_local.long(), num_sample)
        index_out[: index.numel()] = index
        n_out[0] = index.numel()
        labels_remapped.copy_(remapped.to(torch.int32))

    def gather_rows(self, srcs, dsts, index, rows):
        for s, t in zip(srcs, dsts):
            t[:rows] = s[index[:rows]]

    def scatter_rows(self, srcs, dsts, index, rows):
        for s, t in zip(srcs, dsts):
            t[index[:rows]] = s[:rows]

    # ---- forward
    @staticmethod
    def _margin(kind, t, m2, m3):
        if kind == 1:
            return t - m3, torch.ones_like(t)
        theta = math.cos(math.pi - m2)
        sinmm = math.sin(math.pi - m2) * m2
        st = torch.sqrt((1 - t * t).clamp_min(0))
        fin = torch.where(t > theta, t * math.cos(m2) - st * math.sin(m2), t - sinmm)
        dm = torch.where(t > theta, math.cos(m2) + math.sin(m2) * t / torch.sqrt((1 - t * t).clamp_min(1e-12)),
                         torch.ones_like(t))
        return fin, dm

    def forward(self, xn, wn, labels, B, n, d, s, kind, m2, m3, thr, E, n_pad, part_sum, tgt_raw, tgt_e, tgt_z):
        raw = xn[:B].float() @ wn[:n].float().t()
        cl = raw.clamp(-1, 1)
        keep = torch.ones_like(raw, dtype=torch.bool)   # the clamp gate is applied on the target column only (prepare)
        rows = torch.nonzero(labels[:B] >= 0).reshape(-1)
        cols = labels[rows].long()
        if thr > 0:
            dirty = cl > thr
            dirty[rows, cols] = False
            cl = torch.where(dirty, torch.zeros_like(cl), cl)
            keep = keep & ~dirty
        k1 = s * LOG2E
        e = torch.exp2(cl * k1 - (k1 - self.exp_top()))
        t = cl[rows, cols]
        fin, _ = self._margin(kind, t, m2, m3)
        tgt_raw[rows] = raw[rows, cols]
        tgt_e[rows] = torch.exp2(fin * k1 - (k1 - self.exp_top()))
        tgt_z[rows] = fin * s
        e[rows, cols] = 0
        Ev = E[: B * n_pad].view(B, n_pad)
        Ev[:, :n] = torch.where(keep, e, torch.zeros_like(e)).to(torch.bfloat16)
        nt, Bp = self.num_class_tiles(n), self.padded_batch(B)
        ps = part_sum[: nt * Bp].view(nt, Bp)
        cw = self.part_sum_cols()
        for tix in range(nt):
            ps[tix, :B] = e[:, tix * cw:(tix + 1) * cw].sum(1)       # one slab per `cw`-class column group

    def row_stats(self, part_sum, n_tiles, B, labels, tgt_e, stats):
        Bp = self.padded_batch(B)
        stats[:, 0] = part_sum[: n_tiles * Bp].view(n_tiles, Bp)[:, :B].sum(0)
        stats[:, 1] = torch.where(labels[:B] >= 0, tgt_e[:B], torch.zeros_like(tgt_e[:B]))

    def loss(self, stats, B, row_L, out):
        L = stats[:, 0] + stats[:, 1]
        row_L.copy_(L)
        out[0] = -(stats[:, 1] / L).clamp_min(1e-30).log().mean()

    # ---- backward
    def backward_prepare(self, stats, row_L, grad_loss, s, B, d, labels, tgt_raw, kind, m2, xn, xs, coef, E, n_pad):
        g = float(grad_loss[0]) if grad_loss is not None else 1.0
        c = g * s / (B * row_L)
        coef.copy_(c)
        xs[:B] = (xn[:B].float() * c.reshape(-1, 1)).to(torch.bfloat16)
        rows = torch.nonzero(labels[:B] >= 0).reshape(-1)
        raw = tgt_raw[rows]
        t = raw.clamp(-1, 1)
        _, dm = self._margin(kind, t, m2, 0.0)
        mask = (raw.abs() <= 1).float()
        Ev = E[: B * n_pad].view(B, n_pad)
        Ev[rows, labels[rows].long()] = (-dm * mask * stats[rows, 0]).to(torch.bfloat16)

    def cast_f16_to_bf16(self, src, dst, elems):
        dst.reshape(-1)[:elems] = src.reshape(-1)[:elems].to(torch.bfloat16)

    def backward_dx(self, E, n_pad, wn, B, n, d, partial, splits):
        p = partial[: splits * B * d].view(splits, B, d)
        p.zero_()
        p[0] = E[: B * n_pad].view(B, n_pad)[:, :n].float() @ wn[:n].float()

    def dx_finalize(self, partial, splits, coef, x, inv_norm, scale, rows, rows_total, d, out):
        g = partial.reshape(-1)[: splits * rows_total * d].view(splits, rows_total, d)[:, :rows].sum(0)
        if coef is not None:
            g = g * coef[:rows].reshape(-1, 1)
        if x is not None:
            xn = x[:rows] * inv_norm[:rows].reshape(-1, 1)
            g = (g - xn * (xn * g).sum(1, keepdim=True)) * inv_norm[:rows].reshape(-1, 1)
        out[:rows] = g * scale

    def backward_dw(self, E, n_pad, xs, B, n, d, dwn, keep_in_l2=True):
        dwn[:n] = (E[: B * n_pad].view(B, n_pad)[:, :n].float().t() @ xs[:B].float()).to(dwn.dtype)

    def dw_finalize(self, dwn, w, inv_norm_w, rows, d, inv_grad_scale, dw):
        wn = w[:rows] * inv_norm_w[:rows].reshape(-1, 1)
        g = dwn[:rows]
        dw[:rows] = (g - wn * (wn * g).sum(1, keepdim=True)) * inv_norm_w[:rows].reshape(-1, 1) * inv_grad_scale

    def _twin(self, wn_next, wn_next_b, rows):
        """bf16 twin of an fp16 wn_next (what the real update kernels write next to it in AMP mode)."""
        if wn_next_b is not None and wn_next_b is not wn_next:
            assert wn_next.dtype == torch.float16 and wn_next_b.dtype == torch.bfloat16
            wn_next_b[:rows] = wn_next[:rows].to(torch.bfloat16)

    def dw_sgd(self, dwn, w, mom, inv_norm_w, rows, d, lr, momentum, wd, grad_scale, wn_next, inv_norm_next, index=None,
               wn_next_b=None):
        sel = slice(0, rows) if index is None else index[:rows].long()
        g = torch.empty(rows, d)
        inv_grad_scale = 1.0 if grad_scale is None else 1.0 / float(grad_scale[0])
        self.dw_finalize(dwn.float(), w[sel], inv_norm_w, rows, d, inv_grad_scale, g)
        w_new, m_new = ho.sgd_update(w[sel], mom[sel], g, lr, momentum, wd)
        w[sel] = w_new
        mom[sel] = m_new
        if wn_next is not None:
            self.l2norm_rows(w_new, None, rows, wn_next, inv_norm_next)
            self._twin(wn_next, wn_next_b, rows)

    def dw_adam(self, dwn, w, exp_avg, exp_avg_sq, inv_norm_w, rows, d, lr, beta1, beta2, eps, wd, step, decoupled,
                grad_scale, wn_next, inv_norm_next, step_dev=None, index=None, wn_next_b=None):
        sel = slice(0, rows) if index is None else index[:rows].long()
        if step_dev is not None:
            step = int(step_dev[0]) + 1
        inv_grad_scale = 1.0 if grad_scale is None else 1.0 / float(grad_scale[0])
        g = torch.empty(rows, d)
        self.dw_finalize(dwn, w[sel], inv_norm_w, rows, d, inv_grad_scale, g)
        w_new, m_new, v_new = ho.adamw_update(w[sel], exp_avg[sel], exp_avg_sq[sel], g, step, lr, beta1, beta2,
                                              eps, wd, decoupled=bool(decoupled))
        w[sel] = w_new
        exp_avg[sel] = m_new
        exp_avg_sq[sel] = v_new
        if wn_next is not None:
            self.l2norm_rows(w_new, None, rows, wn_next, inv_norm_next)
            self._twin(wn_next, wn_next_b, rows)
