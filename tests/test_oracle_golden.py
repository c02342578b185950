"""Pins the oracle: every function of oracle/ against fixtures produced by the unmodified reference
(tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from helpers import HEAD_CASES, load_case, case_inputs, case_margin, case_perms
from inputs import eval_inputs_cfg5
from oracle import head_oracle as ho
from oracle import eval_oracle as eo


def _close(got, ref, tol):
    """fp64 oracle vs the reference's fp32 run: max abs error relative to the largest reference entry."""
    got = got.numpy() if hasattr(got, "numpy") else np.asarray(got)
    err = np.abs(got.astype(np.float64) - ref.astype(np.float64)).max()
    assert err <= tol * max(np.abs(ref).max(), 1e-12), (err, np.abs(ref).max())


@pytest.mark.parametrize("name", HEAD_CASES)
def test_head_steps_match_reference(name):
    cfg, z = load_case(name)
    weights, xs, ls = case_inputs(cfg)
    W, b = cfg["W"], cfg["b"]
    orc = ho.PartialFCOracle(weights, cfg["C"], case_margin(cfg), cfg["sample_rate"], cfg["lr"], cfg["momentum"],
                             cfg["wd"])
    for s in range(cfg["steps"]):
        xl = [xs[s][r * b:(r + 1) * b] for r in range(W)]
        ll = [ls[s][r * b:(r + 1) * b] for r in range(W)]
        res = orc.step(xl, ll, case_perms(cfg, z, s))
        for r in range(W):
            ref_loss = float(z[f"r{r}_loss_{s}"])
            assert abs(float(res.loss) - ref_loss) <= 2e-6 * max(1.0, abs(ref_loss))
            _close(res.dx_local[r], z[f"r{r}_dx_{s}"], 5e-5)
            _close(res.dw[r], z[f"r{r}_dw_{s}"], 5e-5)
            if cfg["sample_rate"] < 1:
                assert np.array_equal(res.index[r].numpy(), z[f"r{r}_index_{s}"])   # bit-exact index set
    wf, mf = orc.full_weights()
    for r in range(W):
        _close(wf[r], z[f"r{r}_weight_final"], 5e-5)
        _close(mf[r], z[f"r{r}_mom_final"], 2e-4)


@pytest.mark.parametrize("name", ["head_w1_adamw_sampled", "head_w2_adamw_sampled", "head_w1_adamw_full",
                                  "head_w1_adam_sampled"])
def test_adamw_head_steps_match_reference(name):
    """PartialFCAdamW + torch.optim.AdamW (nets/PartialFC.py:235-432): exp_avg / exp_avg_sq rows gathered and scattered
    back around every step, and the reference's step-count quirk (the sampled path corrects the bias with t + 1)."""
    cfg, z = load_case(name)
    weights, xs, ls = case_inputs(cfg)
    W, b = cfg["W"], cfg["b"]
    orc = ho.PartialFCOracle(weights, cfg["C"], case_margin(cfg), cfg["sample_rate"], cfg["lr"], 0.0, cfg["wd"],
                             optimizer=cfg["optimizer"])
    for s in range(cfg["steps"]):
        xl = [xs[s][r * b:(r + 1) * b] for r in range(W)]
        ll = [ls[s][r * b:(r + 1) * b] for r in range(W)]
        res = orc.step(xl, ll, case_perms(cfg, z, s))
        for r in range(W):
            ref_loss = float(z[f"r{r}_loss_{s}"])
            assert abs(float(res.loss) - ref_loss) <= 2e-6 * max(1.0, abs(ref_loss))
            _close(res.dx_local[r], z[f"r{r}_dx_{s}"], 5e-5)
            _close(res.dw[r], z[f"r{r}_dw_{s}"], 5e-5)
            if cfg["sample_rate"] < 1:
                assert np.array_equal(res.index[r].numpy(), z[f"r{r}_index_{s}"])
    wf, _ = orc.full_weights()
    m, v = orc.full_adam_state()
    for r in range(W):
        # Adam divides by sqrt(v): entries with a near-zero gradient amplify fp32-vs-fp64 rounding, so the weights are
        # compared through the UPDATE they received (lr = 1e-3: three steps move a weight by <= 3e-3)
        # (an oracle run in fp32 reproduces the reference to 1.5e-8; off by one in the step count gives cosine 0.995)
        w0 = weights[r].double().numpy()
        got, ref = wf[r].numpy() - w0, z[f"r{r}_weight_final"].astype(np.float64) - w0
        well = z[f"r{r}_exp_avg_sq_final"] > 1e-12               # sqrt(v) >> eps: the update is well-conditioned
        assert well.mean() > 0.3
        assert np.abs(got - ref)[well].max() <= 1e-3 * np.abs(ref).max()   # off by one in the step: 0.5
        assert float((got * ref).sum() / (np.linalg.norm(got) * np.linalg.norm(ref))) >= 0.9995
        _close(m[r], z[f"r{r}_exp_avg_final"], 2e-4)
        _close(v[r], z[f"r{r}_exp_avg_sq_final"], 2e-4)


def test_cfg1_shape_matches_reference():
    """BASELINE.json configs[0] (batch 128, 10 000 classes, d = 512, one process): the [C, d] arrays of the fixture are
    stored as norm + 16-column random projection (tests/golden/make_golden.py::make_cfg1_case)."""
    from inputs import proj_matrix
    cfg, z = load_case("head_cfg1")
    weights, xs, ls = case_inputs(cfg)
    R = proj_matrix(cfg["d"])
    orc = ho.PartialFCOracle(weights, cfg["C"], case_margin(cfg), 1.0, cfg["lr"], cfg["momentum"], cfg["wd"])
    for s in range(cfg["steps"]):
        res = orc.step([xs[s]], [ls[s]])
        ref_loss = float(z[f"r0_loss_{s}"])
        assert abs(float(res.loss) - ref_loss) <= 2e-6 * abs(ref_loss)
        _close(res.dx_local[0], z[f"r0_dx_{s}"], 5e-5)
        _close(res.dw[0].numpy() @ R, z[f"r0_dw_{s}_proj"], 5e-5)
        assert abs(float(res.dw[0].norm()) / float(z[f"r0_dw_{s}_norm"]) - 1) <= 2e-5
    wf, mf = orc.full_weights()
    _close(wf[0].numpy() @ R, z["r0_weight_final_proj"], 5e-5)
    _close(mf[0].numpy() @ R, z["r0_mom_final_proj"], 2e-4)


def test_shard_arithmetic_covers_all_classes():
    for C in (10, 301, 93431, 360232):
        for W in (1, 2, 3, 8):
            spans = [ho.shard_range(C, r, W) for r in range(W)]
            assert sum(n for n, _ in spans) == C
            pos = 0
            for n, start in spans:
                assert start == pos
                pos += n
    assert ho.shard_range(93431, 0, 8) == (11679, 0)
    assert ho.shard_range(93431, 7, 8)[0] == 11678


@pytest.mark.parametrize("key,kind,s,m,thr", [
    ("arcface", "arcface", 64.0, 0.5, 0.0), ("arcface_30", "arcface", 30.0, 0.35, 0.0),
    ("cosface", "cosface", 64.0, 0.4, 0.0), ("combined_arc", "arcface", 64.0, 0.5, 0.0),
    ("combined_cos", "cosface", 64.0, 0.4, 0.0), ("combined_filter", "arcface", 64.0, 0.5, 0.5)])
def test_margin_modules(key, kind, s, m, thr):
    z = np.load(__import__("os").path.join(__import__("helpers").GOLDEN, "margins.npz"))
    logits = torch.from_numpy(z["logits"]).double()
    labels = torch.from_numpy(z["labels"]).reshape(-1)
    mg = ho.Margin(kind=kind, s=s, m=m, filter_thr=thr)
    # rank_logits works on cosines produced by a GEMM; feed the fixture logits through an identity "GEMM"
    n = logits.shape[1]
    f = ho.rank_logits(logits, torch.eye(n, dtype=torch.float64), labels, mg)
    np.testing.assert_allclose(f.z.numpy(), z[key], rtol=1e-5, atol=1e-5)
    # d sum(z) / d logits == grad_gate (identity weight => wn = I)
    np.testing.assert_allclose(f.grad_gate.numpy(), z[key + "_grad"], rtol=1e-5, atol=1e-5)


def _hist(z, prefix, which):
    h = np.zeros(eo.HIST_BINS)
    h[z[f"{prefix}_{which}_nz"]] = z[f"{prefix}_{which}_val"]
    return h


def test_eval_small_case():
    z = np.load(__import__("os").path.join(__import__("helpers").GOLDEN, "eval.npz"))
    hg, hi, sc = eo.pair_score(z["small_e1"], z["small_e2"], z["small_lab"])
    assert np.array_equal(hg, _hist(z, "small", "hg")) and np.array_equal(hi, _hist(z, "small", "hi"))
    np.testing.assert_allclose(sc, z["small_scores"], rtol=0, atol=4e-16)
    rep, th = eo.performance_roc(hg, hi, 1, 3)
    assert th == int(z["small_th"]) and rep == str(z["small_report"])
    assert eo.performance_acc(sc, z["small_lab"], th) == float(z["small_acc"])


def test_eval_cfg5_bit_exact_accuracy():
    z = np.load(__import__("os").path.join(__import__("helpers").GOLDEN, "eval.npz"))
    a, b, lab = eval_inputs_cfg5()
    chk = z["cfg5_input_checksum"]
    assert float(a.astype(np.float64).sum()) == chk[0] and float(b.astype(np.float64).sum()) == chk[1]
    hg, hi, sc = eo.pair_score(a, b, lab)
    assert np.array_equal(hg, _hist(z, "cfg5", "hg")) and np.array_equal(hi, _hist(z, "cfg5", "hi"))
    rep, th = eo.performance_roc(hg, hi)
    assert th == 63399 == int(z["cfg5_th"])
    assert rep == str(z["cfg5_report"])
    acc = eo.performance_acc(sc, lab, th)
    assert acc == float(z["cfg5_acc"]) and abs(acc - 94.66666666666667) < 1e-12


def test_kfold_protocol_properties():
    a, b, lab = eval_inputs_cfg5()
    _, _, sc = eo.pair_score(a, b, lab)
    dist = 4.0 * (1.0 - sc)
    acc, best = eo.kfold_accuracy(dist, lab)
    assert acc.shape == (10,) and 0.9 < acc.mean() < 0.97
    # perfectly separable data -> 100 %
    d2 = np.where(lab, 0.5, 2.5)
    acc2, _ = eo.kfold_accuracy(d2, lab)
    assert np.all(acc2 == 1.0)


def test_kfold_protocol_matches_the_public_lfw_evaluation():
    """The 10-fold variant is not in the reference (BASELINE config 5 asks for it): the oracle is pinned instead against
    the public LFW evaluation as facenet / insightface implement it -- sklearn KFold(10, shuffle=False) over the pair
    list, thresholds arange(0, 4, 0.01) on the squared L2 distance, per fold the threshold with the best TRAINING
    accuracy (np.argmax: first maximum) applied to the held-out pairs -- restated here with sklearn's own splitter."""
    from sklearn.model_selection import KFold
    a, b, lab = eval_inputs_cfg5()
    diff = a.astype(np.float32) - b.astype(np.float32)
    dist = np.sum(np.square(diff.astype(np.float64)), 1)
    thresholds = np.arange(0, 4, 0.01)

    def calc_acc(thr, d, issame):
        pred = np.less(d, thr)
        tp = np.sum(np.logical_and(pred, issame))
        tn = np.sum(np.logical_and(np.logical_not(pred), np.logical_not(issame)))
        return float(tp + tn) / d.size

    accs, bests = [], []
    for train, test in KFold(n_splits=10, shuffle=False).split(np.arange(len(dist))):
        acc_train = np.array([calc_acc(t, dist[train], lab[train]) for t in thresholds])
        k = int(np.argmax(acc_train))
        bests.append(k)
        accs.append(calc_acc(thresholds[k], dist[test], lab[test]))
    got_acc, got_best = eo.kfold_accuracy(dist, lab, folds=10, n_thr=len(thresholds), step=0.01)
    assert np.array_equal(got_best, np.array(bests))
    assert np.array_equal(got_acc, np.array(accs))
    assert abs(100 * got_acc.mean() - 94.52) < 0.01          # 94.52 % on the cfg-5 inputs
    # uneven folds (sklearn gives the first n % folds folds one extra pair)
    n = 1003
    rng = np.random.default_rng(5)
    d2, l2 = rng.random(n) * 4, rng.random(n) < 0.5
    ref = []
    for train, test in KFold(n_splits=10, shuffle=False).split(np.arange(n)):
        k = int(np.argmax([calc_acc(t, d2[train], l2[train]) for t in thresholds]))
        ref.append(calc_acc(thresholds[k], d2[test], l2[test]))
    assert np.array_equal(eo.kfold_accuracy(d2, l2, 10, len(thresholds), 0.01)[0], np.array(ref))


def test_cross_score_matches_reference():
    z = np.load(__import__("os").path.join(__import__("helpers").GOLDEN, "eval.npz"))
    hg, hi, sc, lb = eo.cross_score(z["cross_e"], z["cross_lab"])
    assert np.array_equal(sc, z["cross_scores"])            # same sequential fp64 accumulation: bit-identical
    assert np.array_equal(lb, z["cross_labels"])
    assert np.array_equal(hg, _hist(z, "cross", "hg")) and np.array_equal(hi, _hist(z, "cross", "hi"))
