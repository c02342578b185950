"""CPU oracle for the pair-verification scorer (utils/eval.py of the reference).  TEST INFRASTRUCTURE ONLY.

NumPy restatement; pinned by tests/test_oracle_golden.py against fixtures produced by the reference's own numba /
Python functions (tests/golden/make_golden.py).  The 10-fold protocol (`kfold_accuracy`) does not exist in the
reference (discrepancy 1): its parity is UNPINNED by the reference; it is pinned against the public
LFW evaluation (sklearn KFold restatement) in tests/test_oracle_golden.py.
"""
from __future__ import annotations

import numpy as np

HIST_BINS = 100001


def pair_score(e1: np.ndarray, e2: np.ndarray, labels: np.ndarray):
    """utils/eval.py:68-99.  fp32 difference, fp64 square and accumulation, int() truncation into 100001 bins."""
    diff = (e1.astype(np.float32) - e2.astype(np.float32)).astype(np.float64)
    sum_diff = np.einsum("ij,ij->i", diff, diff)
    score = 1.0 - sum_diff / 4.0
    idx = np.trunc((1e5 - 1.0) * score).astype(np.int64)
    idx = np.where(idx < 0, idx + HIST_BINS, idx)            # numpy negative indexing in the reference loop
    lab = labels.astype(bool)
    hist_g = np.bincount(idx[lab], minlength=HIST_BINS).astype(np.float64)
    hist_i = np.bincount(idx[~lab], minlength=HIST_BINS).astype(np.float64)
    return hist_g, hist_i, score


def cross_score(embeddings: np.ndarray, labels: np.ndarray):
    """utils/eval.py:102-137: all pairs j < i, pair index l = i(i-1)/2 + j, sequential fp64 accumulation over k."""
    e = embeddings.astype(np.float32)
    N = e.shape[0]
    ii, jj = np.tril_indices(N, -1)                      # rows i > j, ordered by i then j == the reference's l
    diff = (e[jj] - e[ii]).astype(np.float64)
    acc = np.zeros(len(ii))
    for k in range(e.shape[1]):                          # same summation order as the reference loop
        acc += diff[:, k] * diff[:, k]
    score = 1.0 - acc / 4.0
    same = np.asarray(labels)[jj] == np.asarray(labels)[ii]
    idx = np.trunc((1e5 - 1.0) * score).astype(np.int64)
    idx = np.where(idx < 0, idx + HIST_BINS, idx)
    hist_g = np.bincount(idx[same], minlength=HIST_BINS).astype(np.float64)
    hist_i = np.bincount(idx[~same], minlength=HIST_BINS).astype(np.float64)
    return hist_g, hist_i, score, same.astype(np.float64)


def roc_sweep(hist_g: np.ndarray, hist_i: np.ndarray, min_level: int = 3, max_level: int = 9):
    """The numbers behind performance_roc (utils/eval.py:7-51, :140-144).

    Thresholds run 100000 .. 1; at threshold th the counters hold the bins strictly above th.
    Returns dict(eer_threshold, eer, frr_at[level], th_at[level], totals).
    """
    hg = np.asarray(hist_g, dtype=np.float64)
    hi = np.asarray(hist_i, dtype=np.float64)
    total_g = float(int(hg.sum()))
    total_i = float(int(hi.sum()))
    ths = np.arange(100000, 0, -1)
    # cumulative counts of bins > th, in sweep order
    cg = np.concatenate(([0.0], np.cumsum(hg[ths])[:-1]))
    ci = np.concatenate(([0.0], np.cumsum(hi[ths])[:-1]))
    far = (ci + hi[ths]) / total_i
    frr = (total_g - cg) / total_g
    diff = np.abs(far - frr)
    k = int(np.argmin(diff))                                  # first minimum from the top == strict '<' updates
    if diff[k] < 1:
        eer_threshold, eer = int(ths[k]), float((far[k] + frr[k]) / 2)
    else:
        eer_threshold, eer = 100000, float("nan")
    frr_at, th_at = [], []
    for level in range(min_level, max_level + 1):
        ok = far <= float(f"1e-{level}")
        if ok.any():
            cand = np.where(ok, frr, np.inf)
            j = int(np.argmin(cand))
            frr_at.append(float(frr[j]))
            th_at.append(int(ths[j]))
        else:
            frr_at.append(None)
            th_at.append(None)
    return dict(eer_threshold=eer_threshold, eer=eer, frr_at=frr_at, th_at=th_at, total_genuine=int(total_g),
                total_imposter=int(total_i))


def format_roc(sweep: dict, min_level: int = 3) -> str:
    """The report string of performance_roc (utils/eval.py:42-48)."""
    out = "\n"
    for i, (frr, th) in enumerate(zip(sweep["frr_at"], sweep["th_at"])):
        out += f"- FRR @ FAR{i + min_level} {100 * frr:6.3f}%, (Threshold = {th / 1e5:.5f})  \n"
    out += "- EER {0:6.3f}%, (Threshold = {1:.5f})\n".format(100 * sweep["eer"], sweep["eer_threshold"] / 1e5)
    out += "- Total count = {:,}\n".format(sweep["total_genuine"] + sweep["total_imposter"])
    out += "- Total genuine count = {:,}\n".format(sweep["total_genuine"])
    out += "- Total imposter count = {:,}\n".format(sweep["total_imposter"])
    return out


def performance_roc(hist_g, hist_i, min_level: int = 3, max_level: int = 9):
    sweep = roc_sweep(hist_g, hist_i, min_level, max_level)
    return format_roc(sweep, min_level), sweep["eer_threshold"]


def performance_acc(scores: np.ndarray, labels: np.ndarray, th) -> float:
    """utils/eval.py:54-66."""
    thd = th / 1e5
    lab = np.asarray(labels)
    fr = int(np.count_nonzero((scores <= thd) & (lab == 1)))
    fa = int(np.count_nonzero((scores > thd) & (lab == 0)))
    return (1 - (fa + fr) / len(scores)) * 100


def kfold_accuracy(dist: np.ndarray, labels: np.ndarray, folds: int = 10, n_thr: int = 400, step: float = 0.01):
    """Standard LFW protocol (not in the reference; pinned against an sklearn-KFold restatement of the public LFW
    evaluation in tests/test_oracle_golden.py): contiguous KFold(folds) over the pair list, thresholds
    k*step on the squared distance, the threshold with the best training-fold accuracy (first maximum) is applied
    to the held-out fold.  Returns (per-fold accuracy, per-fold best threshold index)."""
    n = len(dist)
    lab = labels.astype(bool)
    thr = np.arange(n_thr) * step
    correct = (dist[None, :] < thr[:, None]) == lab[None, :]          # [n_thr, n]
    sizes = np.full(folds, n // folds)
    sizes[: n % folds] += 1
    starts = np.concatenate(([0], np.cumsum(sizes)))
    accs, best = [], []
    for f in range(folds):
        test = np.zeros(n, dtype=bool)
        test[starts[f]:starts[f + 1]] = True
        train_acc = correct[:, ~test].sum(axis=1) / float((~test).sum())
        b = int(np.argmax(train_acc))
        best.append(b)
        accs.append(correct[b, test].sum() / float(test.sum()))
    return np.array(accs), np.array(best)
