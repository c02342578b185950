"""CPU restatements of the two scorer-kernel DESIGNS (csrc/pfc_eval.cu), checked against the oracle's threshold-by-threshold
loops -- the CUDA kernels themselves are compared with the same oracle in tests/test_gpu_kernels.py.

  * k-fold (kfold_kernel): a pair's verdict `dist < t*step` flips exactly once along the sweep, at tmin = the first t with
    dist < t*step; a genuine pair is correct for t >= tmin, an imposter pair for t < tmin.  One histogram entry per pair and
    fold + prefix sums replace n_thr passes over the distances.
  * ROC (roc_kernel): for "minimum FRR with FAR <= 1e-k, first from the top" the candidates are compared by the integer
    numerator of the FRR (strictly monotone in it), ties to the larger threshold; only far is divided per threshold.
"""
import numpy as np

from oracle import eval_oracle as eo


def _kfold_by_flip_threshold(dist, lab, folds, n_thr, step):
    n = len(dist)
    t = np.clip(np.floor(dist / step), 0, n_thr).astype(np.int64)
    t = np.where(np.isnan(dist), n_thr, t)
    for _ in range(3):          # the floor is only a first guess: the fp64 comparison itself decides
        t = np.where((t > 0) & (dist < (t - 1) * step), t - 1, t)
    for _ in range(3):
        t = np.where((t < n_thr) & ~(dist < t * step), t + 1, t)
    base, rem = n // folds, n % folds
    cut = rem * (base + 1)
    idx = np.arange(n)
    fold = np.where(idx < cut, idx // (base + 1), rem + (idx - cut) // max(base, 1))
    hg = np.zeros((folds, n_thr + 1), np.int64)
    hi = np.zeros((folds, n_thr + 1), np.int64)
    np.add.at(hg, (fold[lab], t[lab]), 1)
    np.add.at(hi, (fold[~lab], t[~lab]), 1)
    pg, pi = np.cumsum(hg, 1), np.cumsum(hi, 1)
    correct = pg[:, :n_thr] + (pi[:, -1:] - pi[:, :n_thr])          # [folds, n_thr]
    tot = correct.sum(0)
    acc, best = [], []
    for f in range(folds):
        test_n = base + (1 if f < rem else 0)
        a = (tot - correct[f]) / float(n - test_n)
        b = int(np.argmax(a))
        best.append(b)
        acc.append(correct[f, b] / float(test_n))
    return np.array(acc), np.array(best)


def test_kfold_flip_threshold_design_matches_the_sweep():
    rng = np.random.default_rng(5)
    for n, folds, n_thr in [(6000, 10, 400), (1003, 7, 400), (64, 64, 50), (257, 3, 1), (999, 10, 37)]:
        dist = rng.random(n) * 4.3
        dist[::7] = np.round(dist[::7], 2)
        dist[::11] = rng.integers(0, n_thr + 3, len(dist[::11])) * 0.01          # exactly on (and past) the thresholds
        dist[::13] = np.nextafter(dist[::13], 0)
        lab = rng.random(n) < 0.5
        a1, b1 = _kfold_by_flip_threshold(dist, lab, folds, n_thr, 0.01)
        a2, b2 = eo.kfold_accuracy(dist, lab, folds, n_thr, 0.01)
        assert np.array_equal(b1, b2) and np.array_equal(a1, a2), (n, folds, n_thr)


def _roc_levels_by_numerator(hg, hi, min_level, max_level):
    hg, hi = hg.astype(np.int64), hi.astype(np.int64)
    tot_g, tot_i = int(hg.sum()), int(hi.sum())
    ths = np.arange(100000, 0, -1)
    cg = np.concatenate(([0], np.cumsum(hg[ths])[:-1]))
    ci = np.concatenate(([0], np.cumsum(hi[ths])[:-1]))
    far = (ci + hi[ths]).astype(np.float64) / float(tot_i)
    num = tot_g - cg                                                              # integers
    out_frr, out_th = [], []
    for level in range(min_level, max_level + 1):
        ok = far <= float(f"1e-{level}")
        if not ok.any():
            out_frr.append(None); out_th.append(None)
            continue
        cand = np.where(ok, num, np.iinfo(np.int64).max)
        j = int(np.argmin(cand))                                                  # first minimum = largest threshold
        out_frr.append(float(num[j]) / float(tot_g)); out_th.append(int(ths[j]))
    return out_frr, out_th


def test_roc_level_selection_by_integer_numerator_matches_the_float_sweep():
    rng = np.random.default_rng(6)
    for _ in range(4):
        hg = rng.integers(0, 50, 100001) * (rng.random(100001) < 0.3)
        hi = rng.integers(0, 50, 100001) * (rng.random(100001) < 0.3)
        hi[rng.integers(60000, 100001):] = 0
        ref = eo.roc_sweep(hg.astype(np.float64), hi.astype(np.float64), 1, 9)
        frr, th = _roc_levels_by_numerator(hg, hi, 1, 9)
        assert th == ref["th_at"] and frr == ref["frr_at"]
