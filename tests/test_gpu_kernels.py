"""Kernel-level parity through the C ABI: each CUDA kernel family against plain torch / the oracle
(tools/gpu_probe.py holds the case bodies so the same code serves bring-up diagnostics and the test-suite)."""
import pytest

from tools import gpu_probe

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", ["rows", "fwd_small", "fwd_mid", "dx", "dw", "sample", "eval", "fwd_big", "dx_big",
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
