"""Pair-verification scorer with the reference's function signatures (utils/eval.py), running on the GPU.

NumPy arrays in, NumPy arrays / Python scalars out, exactly like the numba / pure-Python originals that
model/FR_PartialFC.py:263-266, :365-368 call on rank 0 -- the arrays are copied to the current CUDA device,
scored by libpfc_b200 (fr_pair_score / fr_roc / fr_acc_counts / fr_kfold_acc) and the small results copied back.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from . import kernels as K


def _dev():
    if not torch.cuda.is_available():
        raise RuntimeError("the verification scorer runs on CUDA only (no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


def _to_dev(a, dtype):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype).to(_dev(), non_blocking=True).contiguous()


def _histograms(bins, dev):
    """Genuine / imposter histograms laid out back to back: the library then zeroes both with one memset."""
    h = torch.empty(2, bins, dtype=torch.int64, device=dev)
    return h[0], h[1]


def pair_score(embedding_1, embedding_2, labels, metric="euclidean", min_level=3, max_level=9, return_dist=False):
    """utils/eval.py:68-99 -> (hist_genuine[100001] f64, hist_imposter[100001] f64, score_list[N] f64)."""
    assert metric in ["euclidean", "cosine"], "Invalid metric !!!"
    e1 = _to_dev(embedding_1, torch.float32)
    e2 = _to_dev(embedding_2, torch.float32)
    lab = _to_dev(np.asarray(labels).astype(bool), torch.uint8)
    N = e1.shape[0]
    bins = K.hist_bins()
    scores = torch.empty(N, dtype=torch.float64, device=e1.device)
    dist = torch.empty(N, dtype=torch.float64, device=e1.device)
    hg, hi = _histograms(bins, e1.device)
    if metric == "euclidean":       # 'cosine' is accepted but has no code path in the reference either (:80-81)
        K.pair_score(e1, e2, lab, scores, dist, hg, hi)
    else:
        scores.zero_(); dist.zero_(); hg.zero_(); hi.zero_()
    out = (hg.cpu().numpy().astype(np.float64), hi.cpu().numpy().astype(np.float64), scores.cpu().numpy())
    if return_dist:
        return out + (dist.cpu().numpy(),)
    return out


def cross_score(embeddings, labels, metric="euclidean"):
    """utils/eval.py:102-137 -> (hist_genuine, hist_imposter, score_list, label_list) over all pairs j < i."""
    assert metric in ["euclidean", "cosine"], "Invalid metric !!!"
    e = _to_dev(embeddings, torch.float32)
    lab = _to_dev(np.asarray(labels).astype(np.int64), torch.int64)
    N = e.shape[0]
    npairs = int((N - 1) * N / 2)
    bins = K.hist_bins()
    scores = torch.zeros(npairs, dtype=torch.float64, device=e.device)
    label_list = torch.zeros(npairs, dtype=torch.float64, device=e.device)
    hg, hi = _histograms(bins, e.device)
    if metric == "euclidean":
        K.cross_score(e, lab, scores, label_list, hg, hi)
    else:
        hg.zero_(); hi.zero_()
    return (hg.cpu().numpy().astype(np.float64), hi.cpu().numpy().astype(np.float64), scores.cpu().numpy(),
            label_list.cpu().numpy())


def roc_sweep(hist_genuine, hist_imposter, min_level=3, max_level=9):
    """The numbers behind performance_roc: EER threshold / value and FRR @ FAR=1e-k."""
    hg = _to_dev(np.asarray(hist_genuine).astype(np.int64), torch.int64)
    hi = _to_dev(np.asarray(hist_imposter).astype(np.int64), torch.int64)
    buf = torch.zeros(ctypes.sizeof(_lib.RocOut), dtype=torch.uint8, device=hg.device)
    K.roc(hg, hi, min_level, max_level, buf)
    raw = _lib.RocOut.from_buffer_copy(buf.cpu().numpy().tobytes())
    levels = max_level - min_level + 1
    frr_at = [None if raw.th_at[i] < 0 else float(raw.frr_at[i]) for i in range(levels)]
    th_at = [None if raw.th_at[i] < 0 else int(raw.th_at[i]) for i in range(levels)]
    return dict(eer_threshold=int(raw.eer_threshold), eer=float(raw.eer), frr_at=frr_at, th_at=th_at,
                total_genuine=int(raw.total_genuine), total_imposter=int(raw.total_imposter))


def performance_roc(hist_genuine, hist_imposter, min_level=3, max_level=9):
    """utils/eval.py:7-51 -> (report string, eer_threshold)."""
    sw = roc_sweep(hist_genuine, hist_imposter, min_level, max_level)
    roc_result = "\n"
    for idx in range(max_level - min_level + 1):
        roc_result += f"- FRR @ FAR{idx + min_level} {100 * sw['frr_at'][idx]:6.3f}%, (Threshold = {sw['th_at'][idx] / 1e5:.5f})  \n"
    roc_result += "- EER {0:6.3f}%, (Threshold = {1:.5f})\n".format(100 * sw["eer"], sw["eer_threshold"] / 1e5)
    roc_result += "- Total count = {:,}\n".format(sw["total_genuine"] + sw["total_imposter"])
    roc_result += "- Total genuine count = {:,}\n".format(sw["total_genuine"])
    roc_result += "- Total imposter count = {:,}\n".format(sw["total_imposter"])
    return roc_result, sw["eer_threshold"]


def performance_acc(score_list, label_list, th):
    """utils/eval.py:54-66 -> accuracy in percent at threshold th / 1e5."""
    scores = _to_dev(np.asarray(score_list, dtype=np.float64), torch.float64)
    lab = _to_dev(np.asarray(label_list).astype(np.uint8), torch.uint8)
    out = torch.zeros(2, dtype=torch.int64, device=scores.device)
    K.acc_counts(scores, lab, th / 1e5, out)
    fr, fa = (int(v) for v in out.cpu().tolist())
    return (1 - (fa + fr) / (len(score_list))) * 100


def kfold_accuracy(embedding_1, embedding_2, labels, folds=10, n_thr=400, step=0.01):
    """Standard LFW 10-fold protocol on squared distances (BASELINE config 5; not part of the reference).
    Returns (per-fold accuracy [folds], per-fold best threshold index [folds])."""
    e1 = _to_dev(embedding_1, torch.float32)
    e2 = _to_dev(embedding_2, torch.float32)
    lab = _to_dev(np.asarray(labels).astype(bool), torch.uint8)
    N = e1.shape[0]
    bins = K.hist_bins()
    scores = torch.empty(N, dtype=torch.float64, device=e1.device)
    dist = torch.empty(N, dtype=torch.float64, device=e1.device)
    hg, hi = _histograms(bins, e1.device)
    K.pair_score(e1, e2, lab, scores, dist, hg, hi)
    ws = torch.zeros(folds * n_thr, dtype=torch.int32, device=e1.device)
    acc = torch.zeros(folds, dtype=torch.float64, device=e1.device)
    best = torch.zeros(folds, dtype=torch.int32, device=e1.device)
    K.kfold_acc(dist, lab, folds, n_thr, step, ws, acc, best)
    return acc.cpu().numpy(), best.cpu().numpy()
