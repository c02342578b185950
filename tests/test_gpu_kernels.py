"""Kernel-level parity through the C ABI: each CUDA kernel family against plain torch / the oracle
(tools/gpu_probe.py holds the case bodies so the same code serves bring-up diagnostics and the test-suite)."""
import pytest

from tools import gpu_probe

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", ["rows", "fwd_small", "fwd_mid", "dx", "dw", "sample", "eval", "fp16", "fwd_big", "dx_big",
                                  "dw_big"])
def test_kernel_case(case):
    ok = getattr(gpu_probe, "case_" + case)()
    assert ok or ok is None


def test_l2norm_matches_torch_normalize_bitwise_mostly():
    import torch
    from face_recognition_pytorch_b200 import kernels as K
    g = torch.Generator().manual_seed(0)
    x = torch.randn(4096, 512, generator=g).cuda()
    idx = torch.randperm(4096, generator=g)[:1000].cuda()
    xn = torch.empty(1000, 512, dtype=torch.bfloat16, device="cuda")
    inv = torch.empty(1000, device="cuda")
    K.l2norm_rows(x, idx, 1000, xn, inv)
    ref = torch.nn.functional.normalize(x[idx])
    # bf16(x / ||x||): at most one bf16 ulp away from torch (the row norm may differ in the last fp32 bit)
    assert float((xn.float() - ref.to(torch.bfloat16).float()).abs().max()) <= 2 ** -8 * float(ref.abs().max())
    assert float((xn == ref.to(torch.bfloat16)).float().mean()) > 0.99
    torch.testing.assert_close(inv, 1 / x[idx].norm(dim=1), rtol=1e-6, atol=0)
    # zero rows: denominator clamps at 1e-12 like F.normalize
    z = torch.zeros(8, 512, device="cuda")
    K.l2norm_rows(z, None, 8, xn[:8], inv[:8])
    assert float(xn[:8].float().abs().max()) == 0.0


def test_margin_modules_standalone():
    import numpy as np
    import torch
    import face_recognition_pytorch_b200 as pfc
    from helpers import GOLDEN
    import os
    z = np.load(os.path.join(GOLDEN, "margins.npz"))
    labels = torch.from_numpy(z["labels"]).cuda()
    mods = {"arcface": pfc.ArcFace(64.0, 0.5), "arcface_30": pfc.ArcFace(30.0, 0.35), "cosface": pfc.CosFace(64.0, 0.4),
            "combined_arc": pfc.CombinedMarginLoss(64.0, 1.0, 0.5, 0.0),
            "combined_cos": pfc.CombinedMarginLoss(64.0, 1.0, 0.0, 0.4),
            "combined_filter": pfc.CombinedMarginLoss(64.0, 1.0, 0.5, 0.0, interclass_filtering_threshold=0.5)}
    for k, mod in mods.items():
        lg = torch.from_numpy(z["logits"]).cuda().requires_grad_(True)
        y = mod(lg, labels)
        y.backward(torch.ones_like(y))
        np.testing.assert_allclose(y.detach().cpu().numpy(), z[k], rtol=1e-5, atol=1e-4)
        np.testing.assert_allclose(lg.grad.cpu().numpy(), z[k + "_grad"], rtol=1e-4, atol=1e-3)


def test_eval_small_golden_bit_exact():
    import os
    import numpy as np
    import face_recognition_pytorch_b200 as pfc
    from helpers import GOLDEN
    z = np.load(os.path.join(GOLDEN, "eval.npz"))
    hg, hi, sc = pfc.pair_score(z["small_e1"], z["small_e2"], z["small_lab"])
    ref_hg = np.zeros(100001); ref_hg[z["small_hg_nz"]] = z["small_hg_val"]
    ref_hi = np.zeros(100001); ref_hi[z["small_hi_nz"]] = z["small_hi_val"]
    assert np.array_equal(hg, ref_hg) and np.array_equal(hi, ref_hi)
    np.testing.assert_allclose(sc, z["small_scores"], rtol=0, atol=4e-16)
    rep, th = pfc.performance_roc(hg, hi, 1, 3)
    assert th == int(z["small_th"]) and rep == str(z["small_report"])
    assert pfc.performance_acc(sc, z["small_lab"], th) == float(z["small_acc"])
    # empty input
    hg0, hi0, sc0 = pfc.pair_score(np.zeros((0, 64), np.float32), np.zeros((0, 64), np.float32), np.zeros(0, bool))
    assert hg0.sum() == 0 and sc0.shape == (0,)


def test_cross_score_golden_bit_exact():
    import os
    import numpy as np
    import face_recognition_pytorch_b200 as pfc
    from helpers import GOLDEN
    z = np.load(os.path.join(GOLDEN, "eval.npz"))
    hg, hi, sc, lb = pfc.cross_score(z["cross_e"], z["cross_lab"])
    assert np.array_equal(sc, z["cross_scores"]) and np.array_equal(lb, z["cross_labels"])
    ref_hg = np.zeros(100001); ref_hg[z["cross_hg_nz"]] = z["cross_hg_val"]
    ref_hi = np.zeros(100001); ref_hi[z["cross_hi_nz"]] = z["cross_hi_val"]
    assert np.array_equal(hg, ref_hg) and np.array_equal(hi, ref_hi)
    # odd sizes / d not a multiple of 32
    rng = np.random.default_rng(3)
    e = rng.standard_normal((45, 50)).astype(np.float32)
    e /= np.linalg.norm(e, axis=1, keepdims=True)          # the scorer is defined on unit vectors (score in [0, 1])
    lab = rng.integers(0, 4, 45)
    from oracle import eval_oracle as eo
    a = pfc.cross_score(e, lab)
    b = eo.cross_score(e, lab)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


def test_eval_cfg5_accuracy_report_and_kfold_bit_exact():
    """BASELINE configs[4]: 6 000 pairs x 512.  Histograms, EER threshold 63 399, the report string and ACC = 94.6667 %
    against the fixture the unmodified reference produced (tests/golden/make_golden.py); the 10-fold protocol (not in the
    reference) against the oracle: per-fold accuracies and threshold indices identical."""
    import os
    import sys
    import numpy as np
    import face_recognition_pytorch_b200 as pfc
    from helpers import GOLDEN
    sys.path.insert(0, GOLDEN)
    from inputs import eval_inputs_cfg5
    from oracle import eval_oracle as eo
    z = np.load(os.path.join(GOLDEN, "eval.npz"))
    a, b, lab = eval_inputs_cfg5()
    hg, hi, sc, dist = pfc.pair_score(a, b, lab, return_dist=True)
    ref_hg = np.zeros(100001); ref_hg[z["cfg5_hg_nz"]] = z["cfg5_hg_val"]
    ref_hi = np.zeros(100001); ref_hi[z["cfg5_hi_nz"]] = z["cfg5_hi_val"]
    assert np.array_equal(hg, ref_hg) and np.array_equal(hi, ref_hi)
    np.testing.assert_allclose(sc, z["cfg5_scores"], rtol=0, atol=4e-16)
    rep, th = pfc.performance_roc(hg, hi)
    assert th == 63399 == int(z["cfg5_th"])
    assert rep == str(z["cfg5_report"])
    acc = pfc.performance_acc(sc, lab, th)
    assert acc == float(z["cfg5_acc"]) and abs(acc - 94.66666666666667) < 1e-12
    kacc, kbest = pfc.kfold_accuracy(a, b, lab)
    kacc2, kbest2 = eo.kfold_accuracy(dist, lab)
    assert np.array_equal(kbest, kbest2) and np.array_equal(kacc, kacc2)
    assert abs(kacc.mean() * 100 - 94.52) < 0.01


def test_roc_sweep_and_kfold_match_the_oracle_on_dense_and_degenerate_inputs():
    """The cluster ROC kernel and the one-pass k-fold kernel against the oracle's threshold-by-threshold loops: dense random
    histograms (every bin populated, large counts), all mass in one bin, mass at the ends (bins 0 / 1 / 100000), no
    imposters below the genuines (FAR levels never reached -> None), fold sizes that do not divide N, distances on
    threshold boundaries."""
    import numpy as np
    import face_recognition_pytorch_b200 as pfc
    from face_recognition_pytorch_b200 import eval as E
    from oracle import eval_oracle as eo
    rng = np.random.default_rng(11)
    cases = []
    cases.append((rng.integers(0, 1000, 100001), rng.integers(0, 1000, 100001)))
    g = np.zeros(100001, np.int64); i = np.zeros(100001, np.int64); g[70000] = 5; i[30000] = 7
    cases.append((g, i))
    g = np.zeros(100001, np.int64); i = np.zeros(100001, np.int64); g[[0, 1, 100000]] = [3, 4, 5]; i[[0, 1, 99999, 100000]] = [1, 2, 3, 4]
    cases.append((g, i))
    g = rng.integers(0, 3, 100001); i = rng.integers(0, 3, 100001); i[50000:] = 0; g[:50000] = 0
    cases.append((g, i))
    g = (rng.random(100001) < 0.02).astype(np.int64) * rng.integers(1, 2 ** 33, 100001)     # counts beyond 32 bits
    i = (rng.random(100001) < 0.02).astype(np.int64) * rng.integers(1, 2 ** 33, 100001)
    cases.append((g, i))
    for (hg, hi), (lo, hi_lvl) in zip(cases, [(3, 9), (1, 16), (0, 5), (3, 9), (2, 12)]):
        got = E.roc_sweep(hg.astype(np.float64), hi.astype(np.float64), lo, hi_lvl)
        ref = eo.roc_sweep(hg.astype(np.float64), hi.astype(np.float64), lo, hi_lvl)
        assert got["eer_threshold"] == ref["eer_threshold"]
        assert got["eer"] == ref["eer"]
        assert got["th_at"] == ref["th_at"] and got["frr_at"] == ref["frr_at"]
        assert got["total_genuine"] == ref["total_genuine"] and got["total_imposter"] == ref["total_imposter"]
    # k-fold: N not divisible by folds, distances exactly on thresholds and beyond the sweep
    import torch
    from face_recognition_pytorch_b200 import kernels as K
    for N, folds, n_thr in [(6000, 10, 400), (1003, 7, 400), (64, 64, 50), (257, 3, 1)]:
        dist = rng.random(N) * 4.2
        dist[::7] = np.round(dist[::7], 2)             # t * 0.01 up to rounding: the comparison itself must decide
        dist[::11] = (rng.integers(0, 400, len(dist[::11])) * 0.01)
        lab = rng.random(N) < 0.5
        d_dev = torch.from_numpy(dist).cuda()
        l_dev = torch.from_numpy(lab.astype(np.uint8)).cuda()
        ws = torch.zeros(folds * n_thr, dtype=torch.int32, device="cuda")
        acc = torch.zeros(folds, dtype=torch.float64, device="cuda")
        best = torch.zeros(folds, dtype=torch.int32, device="cuda")
        K.kfold_acc(d_dev, l_dev, folds, n_thr, 0.01, ws, acc, best)
        acc2, best2 = eo.kfold_accuracy(dist, lab, folds, n_thr, 0.01)
        assert np.array_equal(best.cpu().numpy(), best2) and np.array_equal(acc.cpu().numpy(), acc2)
