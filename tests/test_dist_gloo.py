"""world_size-2 tests of the head's HOST logic on CPU (gloo): class sharding with an uneven split, all-gather
order, label localisation, the single [B,2] statistics all-reduce, the dX reduce-scatter and its x world_size,
sampling + optimizer patching + scatter-back.  The CUDA kernels are replaced by tests/fake_kernels.py (the C-ABI
contract restated on CPU); results are compared with fixtures the unmodified reference produced with 2 gloo ranks."""
import os
import sys
import types

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


# (name, fused, direct = PartialFC.fused_step instead of autograd, mode): mode "" = defaults, "gather" = conf.inplace_update
# off (sampled + fused: gather / scatter of the active rows like the reference instead of the in-place indexed update),
# "amp" = conf.mixed_precision (fp16 operand storage like the reference's autocast: the 1e-3 loss gate holds at d = 64)
SGD_CASES = [("head_w2_full", False, False, ""), ("head_w2_sampled", False, False, ""),
             ("head_w2_full", True, False, ""), ("head_w2_sampled", True, False, ""),
             ("head_w2_sampled", False, True, ""), ("head_w2_full", True, True, ""),
             # one rank: CombinedMarginLoss with inter-class filtering
             ("head_w1_filter_wide", False, False, ""), ("head_w1_filter_wide", True, True, ""),
             ("head_w2_sampled", True, False, "gather"), ("head_w2_sampled", True, True, "gather"),
             # d = 128, several 256-class tiles per rank
             ("head_w2_d128", True, False, ""), ("head_w1_d128", False, False, ""),
             ("head_w2_d128", True, True, ""),
             ("head_w2_full", True, True, "amp"), ("head_w2_sampled", False, False, "amp"), ("head_w1_d128", True, False, "amp")]
# (name, fused)
ADAM_CASES = [("head_w2_adamw_sampled", False), ("head_w2_adamw_sampled", True), ("head_w1_adamw_full", True),
              ("head_w1_adam_sampled", True), ("head_w1_adam_sampled", False)]


def _world_of(name):
    return 2 if "_w2_" in name else 1


def _group_main(rank, W, port, q):
    """One process per rank and WORLD SIZE: every case of that world size runs in the same process group (spawning a
    pair of interpreters per case made this file the slowest of the CPU suite)."""
    for p in (ROOT, HERE, os.path.join(HERE, "golden")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=W)
    from face_recognition_pytorch_b200 import partial_fc, kernels
    from fake_kernels import FakeKernels
    partial_fc.K = FakeKernels(kernels)            # test-only substitution of the kernel layer
    for case in SGD_CASES:
        if _world_of(case[0]) == W:
            try:
                q.put((rank, ("sgd",) + case, _run_sgd_case(rank, W, *case)))
            except Exception as e:     # report per case: the others still run
                q.put((rank, ("sgd",) + case, {"error": f"{type(e).__name__}: {e}"}))
    for case in ADAM_CASES:
        if _world_of(case[0]) == W:
            try:
                q.put((rank, ("adam",) + case, _run_adam_case(rank, W, *case)))
            except Exception as e:
                q.put((rank, ("adam",) + case, {"error": f"{type(e).__name__}: {e}"}))
    if W == 1:
        try:
            q.put((rank, ("scale",), _run_scale_case()))
        except Exception as e:
            q.put((rank, ("scale",), {"error": f"{type(e).__name__}: {e}"}))
        try:
            q.put((rank, ("bulk_draw",), _run_bulk_draw_case()))
        except Exception as e:
            q.put((rank, ("bulk_draw",), {"error": f"{type(e).__name__}: {e}"}))
    dist.barrier()
    dist.destroy_process_group()


def _run_bulk_draw_case():
    """The head's own sampling draw (no `perm` argument) for a shard big enough for the bulk host generator
    (hostrng.cpu_rand, >= 4096 classes) against the same head drawing with torch.rand: same seed -> same index sets, same
    final weights, same generator state afterwards (nets/PartialFC.py:110 semantics)."""
    import face_recognition_pytorch_b200 as pfc
    from face_recognition_pytorch_b200 import hostrng
    C, d, b, steps = 6000, 64, 48, 3
    g = torch.Generator().manual_seed(11)
    w0 = torch.normal(0, 0.01, (C, d), generator=g)
    data = [(torch.nn.functional.normalize(torch.randn(b, d, generator=g)), torch.randint(0, C, (b,), generator=g))
            for _ in range(steps)]
    out = {"bulk_available": bool(hostrng.enabled())}
    for tag, bulk in (("bulk", True), ("torch", False)):
        saved = hostrng._state["ok"]
        hostrng._state["ok"] = saved if bulk else False
        try:
            conf = types.SimpleNamespace(emd_size=d, sample_rate=0.3, mixed_precision=False, loss_s=64.0, loss_m=0.5)
            head = pfc.PartialFC(conf, C)
            head.load_state_dict({"weight": w0.clone()})
            head.train()
            opt = torch.optim.SGD(head.parameters(), lr=0.1, momentum=0.9, weight_decay=5e-4)
            torch.manual_seed(2024)
            for s, (x, lab) in enumerate(data):
                x = x.clone().requires_grad_(True)
                loss = head(x, lab.clone(), opt)
                loss.backward()
                opt.step()
                opt.zero_grad()
                out[f"{tag}_index_{s}"] = head.weight_index.numpy().copy()
                out[f"{tag}_loss_{s}"] = float(loss)
            head.update()
            out[f"{tag}_weight"] = head.weight.numpy().copy()
            out[f"{tag}_rng"] = torch.get_rng_state().numpy().copy()
        finally:
            hostrng._state["ok"] = saved
    out["steps"] = steps
    return out


def _run_scale_case():
    """A scaled loss (GradScaler: model/FR_PartialFC.py:178-180 calls amp.scale(loss).backward(), then unscale_) must give
    scale x the same gradients: DistCrossEntropyFunc.backward multiplies by loss_gradient.item() (nets/PartialFC.py:474)."""
    from helpers import load_case, case_inputs
    import face_recognition_pytorch_b200 as pfc
    cfg, z = load_case("head_w1_full")
    weights, xs, ls = case_inputs(cfg)
    out = {}
    for tag, scale in (("plain", 1.0), ("scaled", 1024.0)):
        conf = types.SimpleNamespace(emd_size=cfg["d"], sample_rate=1.0, mixed_precision=False, loss_s=cfg["s"],
                                     loss_m=cfg["m"])
        head = pfc.PartialFC(conf, cfg["C"])
        head.load_state_dict({"weight": weights[0].clone()})
        opt = torch.optim.SGD(head.parameters(), lr=cfg["lr"], momentum=cfg["momentum"], weight_decay=cfg["wd"])
        x = xs[0].clone().requires_grad_(True)
        loss = head(x, ls[0].clone(), opt)
        (loss * scale).backward()
        out[tag + "_loss"] = float(loss.detach())
        out[tag + "_dx"] = x.grad.numpy().copy()
        out[tag + "_dw"] = head.weight_activated.grad.numpy().copy()
    return out


def _run_sgd_case(rank, W, name, fused, direct, mode):
    from helpers import load_case, case_inputs, case_perms
    import face_recognition_pytorch_b200 as pfc
    cfg, z = load_case(name)
    weights, xs, ls = case_inputs(cfg)
    b = cfg["b"]
    conf = types.SimpleNamespace(emd_size=cfg["d"], sample_rate=cfg["sample_rate"], mixed_precision=mode == "amp",
                                 loss_s=cfg["s"], loss_m=cfg["m"], fused_optimizer=fused,
                                 inplace_update=mode != "gather")
    if cfg["margin"] == "combined_filter":
        thr = cfg["filter_thr"]
        margin = lambda s_, m_: pfc.CombinedMarginLoss(s_, 1.0, m_, 0.0, interclass_filtering_threshold=thr)  # noqa: E731
    else:
        margin = {"arcface": pfc.ArcFace, "cosface": pfc.CosFace}[cfg["margin"]]
    head = pfc.PartialFC(conf, cfg["C"], margin_loss=margin)
    assert (head.num_local, head.class_start) == pfc.shard_range(cfg["C"], rank, W)
    head.load_state_dict({"weight": weights[rank].clone()})
    dummy = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.SGD([{"params": [dummy]}, {"params": head.parameters()}], lr=cfg["lr"],
                          momentum=cfg["momentum"], weight_decay=cfg["wd"])
    out = {}
    for s in range(cfg["steps"]):
        x = xs[s][rank * b:(rank + 1) * b].clone().requires_grad_(True)
        lab = ls[s][rank * b:(rank + 1) * b].clone()
        perms = case_perms(cfg, z, s)
        opt.zero_grad()
        if direct:       # same step without autograd (PartialFC.fused_step)
            loss, dx = head.fused_step(x.detach(), lab, opt, perm=None if perms is None else perms[rank])
            x.grad = dx
        else:
            loss = head(x, lab, opt, perm=None if perms is None else perms[rank])
            loss.backward()
        out[f"loss_{s}"] = float(loss.detach())
        out[f"dx_{s}"] = x.grad.numpy().copy()
        if not fused:
            out[f"dw_{s}"] = head.weight_activated.grad.numpy().copy()
        if cfg["sample_rate"] < 1:
            out[f"index_{s}"] = head.weight_index.numpy().copy()
        opt.step()
    if cfg["sample_rate"] < 1:
        head.update()
        out["weight_final"] = head.weight.numpy().copy()
    else:
        out["weight_final"] = head.weight_activated.detach().numpy().copy()
    out["state_dict_shape"] = tuple(head.state_dict()["weight"].shape)
    out["state_dict_keys"] = list(head.state_dict().keys())
    out["state_dict_weight"] = head.state_dict()["weight"].detach().numpy().copy()
    return out


def _run_adam_case(rank, W, name, fused):
    """PartialFCAdamW host logic (state rows gathered / scattered, step patched into the optimizer, fused update)."""
    from helpers import load_case, case_inputs, case_perms
    import face_recognition_pytorch_b200 as pfc
    cfg, z = load_case(name)
    weights, xs, ls = case_inputs(cfg)
    b = cfg["b"]
    conf = types.SimpleNamespace(emd_size=cfg["d"], sample_rate=cfg["sample_rate"], mixed_precision=False,
                                 loss_s=cfg["s"], loss_m=cfg["m"], fused_optimizer=fused)
    head = pfc.PartialFCAdamW(conf, cfg["C"])
    head.load_state_dict({"weight": weights[rank].clone()})
    dummy = torch.nn.Parameter(torch.zeros(1))
    opt_cls = torch.optim.AdamW if cfg["optimizer"] == "adamw" else torch.optim.Adam      # Adam: coupled weight decay
    opt = opt_cls([{"params": [dummy]}, {"params": head.parameters()}], lr=cfg["lr"], weight_decay=cfg["wd"])
    out = {}
    for s in range(cfg["steps"]):
        x = xs[s][rank * b:(rank + 1) * b].clone().requires_grad_(True)
        lab = ls[s][rank * b:(rank + 1) * b].clone()
        perms = case_perms(cfg, z, s)
        opt.zero_grad()
        loss = head(x, lab, opt, perm=None if perms is None else perms[rank])
        loss.backward()
        out[f"loss_{s}"] = float(loss.detach())
        if cfg["sample_rate"] < 1:
            out[f"index_{s}"] = head.weight_index.numpy().copy()
        opt.step()
    if cfg["sample_rate"] < 1:
        head.update()
        out["weight_final"] = head.weight.numpy().copy()
        out["exp_avg_final"] = head.weight_exp_avg.numpy().copy()
        out["exp_avg_sq_final"] = head.weight_exp_avg_sq.numpy().copy()
    else:
        out["weight_final"] = head.weight_activated.detach().numpy().copy()
    return out


@pytest.fixture(scope="module")
def group_results():
    """{(kind, name, ...): {rank: out}} for every case, from ONE spawn per world size."""
    results = {}
    ctx = mp.get_context("spawn")
    for W, port in ((2, 29821), (1, 29822)):
        q = ctx.Queue()
        procs = [ctx.Process(target=_group_main, args=(r, W, port, q)) for r in range(W)]
        for p in procs:
            p.start()
        n_cases = (sum(_world_of(c[0]) == W for c in SGD_CASES) + sum(_world_of(c[0]) == W for c in ADAM_CASES)
                   + (2 if W == 1 else 0))            # + the scale and the bulk-draw case
        for _ in range(n_cases * W):
            rank, key, out = q.get(timeout=600)
            results.setdefault(key, {})[rank] = out
        for p in procs:
            p.join(timeout=120)
            assert p.exitcode == 0
    return results


def _case_result(group_results, key):
    res = group_results[key]
    for r, out in res.items():
        assert "error" not in out, f"rank {r}: {out.get('error')}"
    return res


def test_bulk_host_draw_gives_the_sampling_of_torch_rand(group_results):
    res = _case_result(group_results, ("bulk_draw",))[0]
    assert res["bulk_available"]
    for s in range(res["steps"]):
        assert np.array_equal(res[f"bulk_index_{s}"], res[f"torch_index_{s}"])
        assert len(res[f"bulk_index_{s}"]) == 1800                     # int(0.3 * 6000), nets/PartialFC.py:63
        assert res[f"bulk_loss_{s}"] == res[f"torch_loss_{s}"]
    assert np.array_equal(res["bulk_weight"], res["torch_weight"])
    assert np.array_equal(res["bulk_rng"], res["torch_rng"])


def test_scaled_loss_scales_the_gradients(group_results):
    res = _case_result(group_results, ("scale",))[0]
    assert res["plain_loss"] == res["scaled_loss"]
    for k in ("dx", "dw"):
        a, b = res["plain_" + k].astype(np.float64) * 1024.0, res["scaled_" + k].astype(np.float64)
        # bf16 rounding of xs = c_i * xn is scale-invariant for a power-of-two scale: the two runs agree to fp32 rounding
        assert np.abs(a - b).max() <= 1e-5 * np.abs(a).max()


@pytest.mark.parametrize("name,fused", ADAM_CASES)
def test_adamw_host_logic_matches_reference(group_results, name, fused):
    """Against fixtures of the reference's PartialFCAdamW: same sampled rows, the Adam state of re-sampled rows carried
    across steps, and the reference's step count (sampled: bias correction with t + 1) in the un-fused AND fused path."""
    sys.path.insert(0, HERE)
    from helpers import load_case
    cfg, z = load_case(name)
    W = cfg["W"]
    res = _case_result(group_results, ("adam", name, fused))
    from inputs import synth_inputs, shard
    w_full, _, _ = synth_inputs(cfg["C"], cfg["d"], cfg["b"] * W, 1)
    for r in range(W):
        for s in range(cfg["steps"]):
            ref_loss = float(z[f"r{r}_loss_{s}"])
            assert abs(res[r][f"loss_{s}"] - ref_loss) <= 6e-3 * abs(ref_loss)
            if cfg["sample_rate"] < 1:
                assert np.array_equal(res[r][f"index_{s}"], z[f"r{r}_index_{s}"])
        nl, cs = shard(cfg["C"], r, W)
        w0 = w_full[cs:cs + nl].numpy().astype(np.float64)
        got, ref = res[r]["weight_final"] - w0, z[f"r{r}_weight_final"] - w0
        # the operands are bf16 here (tests/fake_kernels.py) and Adam normalises the update: elements whose gradient is
        # noise-sized may flip sign, so the update is compared where the first moment is well above the noise
        m_ref = np.abs(z[f"r{r}_exp_avg_final"])
        well = m_ref > 0.05 * m_ref.max()
        assert well.mean() > 0.05
        err = np.abs(got - ref)[well].mean() / np.abs(ref).max()
        assert err <= 5e-3, err                     # measured 4e-4; bias correction off by one step: 8e-2
        assert _cos(got, ref) >= 0.995
        assert np.array_equal(np.abs(got).max(1) > 0, np.abs(ref).max(1) > 0)       # exactly the same rows were touched
        if cfg["sample_rate"] < 1:
            assert _cos(res[r]["exp_avg_final"], z[f"r{r}_exp_avg_final"]) >= 0.999
            assert _cos(res[r]["exp_avg_sq_final"], z[f"r{r}_exp_avg_sq_final"]) >= 0.999


def _cos(a, b):
    a, b = a.astype(np.float64).ravel(), b.astype(np.float64).ravel()
    return float(a @ b / (np.linalg.norm(a) * np.linalg.norm(b)))


@pytest.mark.parametrize("name,fused,direct,mode", SGD_CASES)
def test_two_rank_host_logic_matches_reference(group_results, name, fused, direct, mode):
    sys.path.insert(0, HERE)
    from helpers import load_case
    cfg, z = load_case(name)
    W = cfg["W"]
    res = _case_result(group_results, ("sgd", name, fused, direct, mode))
    for s in range(cfg["steps"]):
        if W > 1:
            assert res[0][f"loss_{s}"] == res[1][f"loss_{s}"]                # every rank returns the global loss
        for r in range(W):
            ref_loss = float(z[f"r{r}_loss_{s}"])
            rtol = 1e-3 if mode == "amp" else 6e-3                           # bf16 operands at d = 64 (see test_gpu_head.py)
            assert abs(res[r][f"loss_{s}"] - ref_loss) <= rtol * abs(ref_loss)
            assert _cos(res[r][f"dx_{s}"], z[f"r{r}_dx_{s}"]) >= 0.999
            assert abs(np.linalg.norm(res[r][f"dx_{s}"]) / np.linalg.norm(z[f"r{r}_dx_{s}"]) - 1) < 2e-2   # incl. x W
            if not fused:
                assert _cos(res[r][f"dw_{s}"], z[f"r{r}_dw_{s}"]) >= 0.999
            if cfg["sample_rate"] < 1:
                assert np.array_equal(res[r][f"index_{s}"], z[f"r{r}_index_{s}"])
    from inputs import synth_inputs, shard
    w_full, _, _ = synth_inputs(cfg["C"], cfg["d"], cfg["b"] * W, 1)
    for r in range(W):
        nl, cs = shard(cfg["C"], r, W)
        assert res[r]["state_dict_shape"] == (nl, cfg["d"])
        assert res[r]["state_dict_keys"] == ["weight"]                      # nets/PartialFC.py:210-221: no prefix, one key
        w0 = w_full[cs:cs + nl].numpy()
        if f"r{r}_state_dict_weight" in z.files:
            assert _cos(res[r]["state_dict_weight"] - w0, z[f"r{r}_state_dict_weight"] - w0) >= 0.999
        assert _cos(res[r]["weight_final"] - w0, z[f"r{r}_weight_final"] - w0) >= 0.999
