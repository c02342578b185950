"""Torch-tensor level wrappers over the C ABI (include/pfc.h).

Every function takes CUDA tensors, passes raw device pointers plus the current CUDA stream to libpfc_b200 and
returns nothing (outputs are caller-allocated).  No host synchronisation, no allocation, no fallback: a CPU tensor
is an error.
"""
import ctypes

import torch

from . import _lib
from ._lib import check, lib

MARGIN_ARCFACE, MARGIN_COSFACE = 0, 1


def _p(t, dtype=None):
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("libpfc_b200 kernels need CUDA tensors (there is no CPU fallback)")
    if not t.is_contiguous():
        raise RuntimeError("libpfc_b200 kernels need contiguous tensors")
    if dtype is not None and t.dtype != dtype:
        raise TypeError(f"expected {dtype}, got {t.dtype}")
    return ctypes.c_void_p(t.data_ptr())


def _op(t):
    """A normalised GEMM operand (Xn / Wn): bf16 by default, fp16 in the reference's AMP mode; -> (pointer, fp16 flag)."""
    if t is None:
        return None, 0
    if t.dtype not in (BF16, F16):
        raise TypeError(f"expected a bf16 or fp16 operand, got {t.dtype}")
    return _p(t), int(t.dtype == F16)


def _copy_b(wn_next, wn_next_b):
    """Pointer of the optional bf16 copy of an fp16 `wn_next` (None -> NULL)."""
    if wn_next_b is None or wn_next_b is wn_next:
        return None
    if wn_next is None or wn_next.dtype != F16 or wn_next_b.shape != wn_next.shape:
        raise TypeError("wn_next_b is the bf16 twin of an fp16 wn_next of the same shape")
    return _p(wn_next_b, BF16)


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


# ---- accounting used by bench.py: how many kernels were launched, and (optionally) their device time
_LAUNCHES_PER_CALL = {}
_count = 0
_timing = None      # name -> list of (start_event, end_event) while enabled


def launch_count():
    return _count


def enable_timing(on):
    global _timing
    _timing = {} if on else None


def collect_timing():
    """{name: {"calls", "ms_total", "ms_avg"}} for the calls made since enable_timing(True); synchronises."""
    torch.cuda.synchronize()
    out = {}
    for name, evs in (_timing or {}).items():
        ms = [a.elapsed_time(b) for a, b in evs]
        out[name] = {"calls": len(ms), "ms_total": sum(ms), "ms_avg": sum(ms) / max(1, len(ms))}
    return out


def _timed(name):
    def deco(fn):
        def wrapper(*a, **k):
            global _count
            _count += _LAUNCHES_PER_CALL.get(name, 1)
            if _timing is None:
                return fn(*a, **k)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = fn(*a, **k)
            e1.record()
            _timing.setdefault(name, []).append((e0, e1))
            return r
        wrapper.__name__ = fn.__name__
        wrapper.__doc__ = fn.__doc__
        return wrapper
    return deco


F32, BF16, I32, I64, U8, F64 = torch.float32, torch.bfloat16, torch.int32, torch.int64, torch.uint8, torch.float64
F16 = torch.float16

exp_top = lib.pfc_exp_top
padded_classes = lib.pfc_padded_classes
padded_batch = lib.pfc_padded_batch
num_class_tiles = lib.pfc_num_class_tiles
part_sum_cols = lib.pfc_part_sum_cols
dx_splits = lib.pfc_dx_splits
dx_max_splits = lib.pfc_dx_max_splits
sample_workspace_bytes = lib.pfc_sample_workspace_bytes
sample_launches = lib.pfc_sample_launches
hist_bins = lib.pfc_eval_hist_bins


def spill_to_rowmajor(E, B, n_pad):
    """The class-blocked spill E'[n_pad/64][B][64] (include/pfc.h, pfc_forward) as a row-major [B, n_pad] matrix
    (a copy; for tests and tools -- the kernels only ever see the blocked layout)."""
    return E.reshape(-1)[: B * n_pad].view(n_pad // 64, B, 64).permute(1, 0, 2).reshape(B, n_pad)


def spill_from_rowmajor(M):
    """Inverse of spill_to_rowmajor: a row-major [B, n_pad] bf16 matrix -> flat class-blocked buffer."""
    B, n_pad = M.shape
    return M.view(B, n_pad // 64, 64).permute(1, 0, 2).contiguous().reshape(-1)


@_timed("pfc_cast_f16_to_bf16")
def cast_f16_to_bf16(src, dst, elems):
    """AMP mode: bf16 copy of the fp16 shard for the dX contraction (tcgen05 takes no mixed bf16 x fp16 operands)."""
    check(lib.pfc_cast_f16_to_bf16(_p(src, F16), _p(dst, BF16), elems, _stream()), "pfc_cast_f16_to_bf16")


@_timed("pfc_l2norm_rows")
def l2norm_rows(x, index, rows, xn, inv_norm):
    d = x.shape[1]
    xp, f16 = _op(xn)
    check(lib.pfc_l2norm_rows(_p(x, F32), _p(index, I64), rows, d, xp, _p(inv_norm, F32), f16, _stream()),
          "pfc_l2norm_rows")


@_timed("pfc_l2norm_rows_localize")
def l2norm_rows_localize(x, rows, xn, inv_norm, labels, class_start, num_local, labels_local):
    d = x.shape[1]
    xp, f16 = _op(xn)
    check(lib.pfc_l2norm_rows_localize(_p(x, F32), rows, d, xp, _p(inv_norm, F32), _p(labels, I64),
                                       class_start, num_local, _p(labels_local, I32), f16, _stream()),
          "pfc_l2norm_rows_localize")


@_timed("pfc_localize_labels")
def localize_labels(labels, class_start, num_local, out):
    check(lib.pfc_localize_labels(_p(labels, I64), labels.numel(), class_start, num_local, _p(out, I32), _stream()),
          "pfc_localize_labels")


@_timed("pfc_sample")
def sample(perm, labels_local, num_local, num_sample, index_out, n_out, labels_remapped, workspace):
    global _count
    _count += lib.pfc_sample_launches(num_local) - 1        # one cluster kernel, or six tiled ones for huge shards
    check(lib.pfc_sample(_p(perm, F32), _p(labels_local, I32), labels_local.numel(), num_local, num_sample,
                         _p(index_out, I64), _p(n_out, I32), _p(labels_remapped, I32), _p(workspace, U8),
                         workspace.numel(), _stream()), "pfc_sample")


@_timed("pfc_gather_rows")
def gather_rows(srcs, dsts, index, rows):
    d = srcs[0].shape[1]
    check(lib.pfc_gather_rows(_lib.ptr_array([_p(s, F32).value for s in srcs]),
                              _lib.ptr_array([_p(t, F32).value for t in dsts]), len(srcs), _p(index, I64), rows, d,
                              _stream()), "pfc_gather_rows")


@_timed("pfc_scatter_rows")
def scatter_rows(srcs, dsts, index, rows):
    d = srcs[0].shape[1]
    check(lib.pfc_scatter_rows(_lib.ptr_array([_p(s, F32).value for s in srcs]),
                               _lib.ptr_array([_p(t, F32).value for t in dsts]), len(srcs), _p(index, I64), rows, d,
                               _stream()), "pfc_scatter_rows")


@_timed("pfc_forward")
def forward(xn, wn, labels_local, B, n, d, s, margin_kind, m2, m3, filter_thr, E, n_pad, part_sum, tgt_raw, tgt_e,
            tgt_z):
    (xp, f16), (wp, wf16) = _op(xn), _op(wn)
    if f16 != wf16:
        raise TypeError("pfc_forward: Xn and Wn must have the same operand format")
    check(lib.pfc_forward(xp, wp, _p(labels_local, I32), B, n, d, s, margin_kind, m2, m3,
                          filter_thr, _p(E, BF16), n_pad, _p(part_sum, F32), _p(tgt_raw, F32), _p(tgt_e, F32),
                          _p(tgt_z, F32), f16, _stream()), "pfc_forward")


@_timed("pfc_margin_apply")
def margin_apply(logits, labels, margin_kind, s, m2, m3, filter_thr, out, gate):
    B, n = logits.shape
    check(lib.pfc_margin_apply(_p(logits, F32), _p(labels, I64), B, n, margin_kind, s, m2, m3, filter_thr,
                               _p(out, F32), _p(gate, F32), _stream()), "pfc_margin_apply")


@_timed("pfc_row_stats")
def row_stats(part_sum, n_tiles, B, labels_local, tgt_e, stats):
    check(lib.pfc_row_stats(_p(part_sum, F32), n_tiles, B, _p(labels_local, I32), _p(tgt_e, F32), _p(stats, F32),
                            _stream()), "pfc_row_stats")


@_timed("pfc_row_stats_loss")
def row_stats_loss(part_sum, n_tiles, B, labels_local, tgt_e, stats, row_L, out, ticket):
    check(lib.pfc_row_stats_loss(_p(part_sum, F32), n_tiles, B, _p(labels_local, I32), _p(tgt_e, F32), _p(stats, F32),
                                 _p(row_L, F32), _p(out, F32), _p(ticket, I32), _stream()), "pfc_row_stats_loss")


@_timed("pfc_row_stats_loss_prepare")
def row_stats_loss_prepare(part_sum, n_tiles, B, labels_local, tgt_e, stats, row_L, out, ticket, grad_loss, s, d, tgt_raw,
                           margin_kind, m2, xn, xs, coef, E, n_pad):
    """row_stats_loss + backward_prepare in one launch (one GPU, no-autograd step)."""
    xp, f16 = _op(xn)
    check(lib.pfc_row_stats_loss_prepare(_p(part_sum, F32), n_tiles, B, _p(labels_local, I32), _p(tgt_e, F32),
                                         _p(stats, F32), _p(row_L, F32), _p(out, F32), _p(ticket, I32),
                                         _p(grad_loss, F32), s, d, _p(tgt_raw, F32), margin_kind, m2, xp,
                                         _p(xs, BF16), _p(coef, F32), _p(E, BF16), n_pad, f16, _stream()),
          "pfc_row_stats_loss_prepare")


@_timed("pfc_loss")
def loss(stats, B, row_L, out):
    check(lib.pfc_loss(_p(stats, F32), B, _p(row_L, F32), _p(out, F32), _stream()), "pfc_loss")


@_timed("pfc_backward_prepare")
def backward_prepare(stats, row_L, grad_loss, s, B, d, labels_local, tgt_raw, margin_kind, m2, xn, xs, coef, E,
                     n_pad):
    xp, f16 = _op(xn)
    check(lib.pfc_backward_prepare(_p(stats, F32), _p(row_L, F32), _p(grad_loss, F32), s, B, d, _p(labels_local, I32),
                                   _p(tgt_raw, F32), margin_kind, m2, xp, _p(xs, BF16), _p(coef, F32),
                                   _p(E, BF16), n_pad, f16, _stream()), "pfc_backward_prepare")


@_timed("pfc_backward_dx")
def backward_dx(E, n_pad, wn, B, n, d, partial, splits):
    check(lib.pfc_backward_dx(_p(E, BF16), n_pad, _p(wn, BF16), B, n, d, _p(partial, F32), splits, _stream()),
          "pfc_backward_dx")


@_timed("pfc_dx_finalize")
def dx_finalize(partial, splits, coef, x, inv_norm, scale, rows, rows_total, d, out):
    check(lib.pfc_dx_finalize(_p(partial, F32), splits, _p(coef, F32), _p(x, F32), _p(inv_norm, F32), scale, rows,
                              rows_total, d, _p(out, F32), _stream()), "pfc_dx_finalize")


@_timed("pfc_backward_dw")
def backward_dw(E, n_pad, xs, B, n, d, dwn, keep_in_l2=True):
    """dwn: fp32 [n,d], or bf16 [n,d] (fused-SGD spill; keep_in_l2: the update kernel runs right behind this GEMM)."""
    is_bf16 = dwn.dtype == BF16
    check(lib.pfc_backward_dw(_p(E, BF16), n_pad, _p(xs, BF16), B, n, d, _p(dwn, BF16 if is_bf16 else F32),
                              (1 if keep_in_l2 else 2) if is_bf16 else 0, _stream()), "pfc_backward_dw")


@_timed("pfc_dw_finalize")
def dw_finalize(dwn, w, inv_norm_w, rows, d, inv_grad_scale, dw):
    check(lib.pfc_dw_finalize(_p(dwn, F32), _p(w, F32), _p(inv_norm_w, F32), rows, d, inv_grad_scale, _p(dw, F32),
                              _stream()), "pfc_dw_finalize")


@_timed("pfc_dw_sgd")
def dw_sgd(dwn, w, mom, inv_norm_w, rows, d, lr, momentum, weight_decay, grad_scale, wn_next, inv_norm_next, index=None,
           wn_next_b=None):
    """grad_scale: device scalar holding the loss scale the gradient carries (divided out first), or None.
    index (int64 [rows], ascending): w / mom are the FULL shard arrays and row r of dwn updates row index[r] in place.
    wn_next_b (fp16 wn_next only): bf16 copy of the same rows for the next step's dX contraction."""
    is_bf16 = dwn.dtype == BF16
    wp, f16 = _op(wn_next)
    wb = _copy_b(wn_next, wn_next_b)
    check(lib.pfc_dw_sgd(_p(dwn, BF16 if is_bf16 else F32), int(is_bf16), _p(w, F32), _p(mom, F32),
                         _p(inv_norm_w, F32), rows, d, lr, momentum, weight_decay, _p(grad_scale, F32),
                         wp, _p(inv_norm_next, F32), _p(index, I64), f16, wb, _stream()), "pfc_dw_sgd")


@_timed("pfc_dw_adam")
def dw_adam(dwn, w, exp_avg, exp_avg_sq, inv_norm_w, rows, d, lr, beta1, beta2, eps, weight_decay, step, decoupled,
            grad_scale, wn_next, inv_norm_next, step_dev=None, index=None, wn_next_b=None):
    """step_dev (int32 device scalar): the update is step step_dev[0] + 1 (CUDA-graph replay), `step` is ignored.
    index, wn_next_b: as in dw_sgd (in-place update of a sampled shard; bf16 copy of an fp16 wn_next)."""
    wp, f16 = _op(wn_next)
    wb = _copy_b(wn_next, wn_next_b)
    check(lib.pfc_dw_adam(_p(dwn, F32), _p(w, F32), _p(exp_avg, F32), _p(exp_avg_sq, F32), _p(inv_norm_w, F32), rows,
                          d, lr, beta1, beta2, eps, weight_decay, step, int(decoupled), _p(grad_scale, F32),
                          wp, _p(inv_norm_next, F32), _p(step_dev, I32), _p(index, I64), f16, wb, _stream()),
          "pfc_dw_adam")


# ---- peer-memory exchanges (peer_* are ctypes arrays of W mapped device pointers, see partial_fc._PeerExchange)
@_timed("pfc_peer_barrier")
def peer_barrier(peer_flags, counter, rank, W):
    check(lib.pfc_peer_barrier(peer_flags, _p(counter, I32), rank, W, _stream()), "pfc_peer_barrier")


@_timed("pfc_peer_l2norm_gather")
def peer_l2norm_gather(x, labels, rank, W, peer_xn_all, peer_labels_all, inv_norm, fp16=False):
    """fp16: the peers' xn_all buffers hold fp16 instead of bf16 rows (the reference's AMP operands)."""
    b, d = x.shape
    check(lib.pfc_peer_l2norm_gather(_p(x, F32), _p(labels, I64), b, d, rank, W, peer_xn_all, peer_labels_all,
                                     _p(inv_norm, F32), int(bool(fp16)), _stream()), "pfc_peer_l2norm_gather")


@_timed("pfc_peer_row_stats")
def peer_row_stats(part_sum, n_tiles, B, labels_local, tgt_e, rank, W, peer_slots):
    check(lib.pfc_peer_row_stats(_p(part_sum, F32), n_tiles, B, _p(labels_local, I32), _p(tgt_e, F32), rank, W,
                                 peer_slots, _stream()), "pfc_peer_row_stats")


@_timed("pfc_peer_loss")
def peer_loss(peer_flags, state, rank, slots, W, B, stats, row_L, out):
    """peer_flags / state None: no barrier (the caller ran pfc_peer_barrier); else the kernel takes it at its start."""
    check(lib.pfc_peer_loss(peer_flags, _p(state, I32), rank, _p(slots, F32), W, B, _p(stats, F32), _p(row_L, F32),
                            _p(out, F32), _stream()), "pfc_peer_loss")


@_timed("pfc_peer_loss_prepare")
def peer_loss_prepare(peer_flags, state, rank, slots, W, B, stats, row_L, out, ticket, grad_loss, s, d, labels_local,
                      tgt_raw, margin_kind, m2, xn, xs, coef, E, n_pad):
    """barrier + peer_loss + backward_prepare in one launch (no-autograd step)."""
    xp, f16 = _op(xn)
    check(lib.pfc_peer_loss_prepare(peer_flags, _p(state, I32), rank, _p(slots, F32), W, B, _p(stats, F32),
                                    _p(row_L, F32), _p(out, F32), _p(ticket, I32), _p(grad_loss, F32), s, d,
                                    _p(labels_local, I32), _p(tgt_raw, F32), margin_kind, m2, xp, _p(xs, BF16),
                                    _p(coef, F32), _p(E, BF16), n_pad, f16, _stream()), "pfc_peer_loss_prepare")


@_timed("pfc_peer_localize_labels")
def peer_localize_labels(peer_flags, state, rank, W, labels, class_start, num_local, out):
    check(lib.pfc_peer_localize_labels(peer_flags, _p(state, I32), rank, W, _p(labels, I64), labels.numel(), class_start,
                                       num_local, _p(out, I32), _stream()), "pfc_peer_localize_labels")


@_timed("pfc_peer_dx_finalize")
def peer_dx_finalize(peer_flags, state, rank, W, dx_slots, x, inv_norm, scale, b, d, out):
    check(lib.pfc_peer_dx_finalize(peer_flags, _p(state, I32), rank, W, _p(dx_slots, F32), _p(x, F32), _p(inv_norm, F32),
                                   scale, b, d, _p(out, F32), _stream()), "pfc_peer_dx_finalize")


@_timed("pfc_peer_dx_scatter")
def peer_dx_scatter(partial, splits, coef, B, b, d, rank, W, peer_dx_slots):
    check(lib.pfc_peer_dx_scatter(_p(partial, F32), splits, _p(coef, F32), B, b, d, rank, W, peer_dx_slots,
                                  _stream()), "pfc_peer_dx_scatter")


# ---- verification scorer
@_timed("fr_pair_score")
def pair_score(e1, e2, labels_u8, scores, dist, hist_g, hist_i):
    N, d = e1.shape
    check(lib.fr_pair_score(_p(e1, F32), _p(e2, F32), _p(labels_u8, U8), N, d, _p(scores, F64), _p(dist, F64),
                            _p(hist_g, I64), _p(hist_i, I64), _stream()), "fr_pair_score")


@_timed("fr_cross_score")
def cross_score(e, labels_i64, scores, label_list, hist_g, hist_i):
    N, d = e.shape
    check(lib.fr_cross_score(_p(e, F32), _p(labels_i64, I64), N, d, _p(scores, F64), _p(label_list, F64),
                             _p(hist_g, I64), _p(hist_i, I64), _stream()), "fr_cross_score")


@_timed("fr_roc")
def roc(hist_g, hist_i, min_level, max_level, out_bytes):
    check(lib.fr_roc(_p(hist_g, I64), _p(hist_i, I64), min_level, max_level, _p(out_bytes, U8), _stream()), "fr_roc")


@_timed("fr_acc_counts")
def acc_counts(scores, labels_u8, threshold, fr_fa):
    check(lib.fr_acc_counts(_p(scores, F64), _p(labels_u8, U8), scores.numel(), threshold, _p(fr_fa, I64), _stream()),
          "fr_acc_counts")


@_timed("fr_kfold_acc")
def kfold_acc(dist, labels_u8, folds, n_thr, step, correct_ws, acc, best_idx):
    check(lib.fr_kfold_acc(_p(dist, F64), _p(labels_u8, U8), dist.numel(), folds, n_thr, step, _p(correct_ws, I32),
                           _p(acc, F64), _p(best_idx, I32), _stream()), "fr_kfold_acc")
