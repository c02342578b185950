"""GPU checks of the head's execution modes against each other and against fixtures of the reference:

  * GraphedHeadStep(autograd=False) / PartialFC.fused_step: same kernels as forward + loss.backward() with d loss = 1,
    so the results must be bit-identical (host logic also covered on CPU by tests/test_dist_gloo.py).
  * PartialFCAdamW with sampling, fused update: bias correction with the reference's step count (t + 1, pinned on the CPU
    by tests/test_oracle_golden.py and tests/test_dist_gloo.py against fixtures of the reference's PartialFCAdamW).
  * CombinedMarginLoss with inter-class filtering INSIDE the head (the kFilter branch of the forward epilogue).
  * conf.dx_side_stream (fork of the dX tail next to the dW GEMM + update) on and off, eager and graph-replayed, at shapes
    with odd tile counts and ragged class tails: identical bits.
  * the L2 residency hints of the bf16 gradient (pfc_debug_l2_grad): cache hints only, bit-identical results.
  * PartialFCAdamW inside GraphedHeadStep (step count in a device scalar).
"""
import os
import types

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pfc():
    import torch.distributed as dist
    if not dist.is_initialized():
        dist.init_process_group("nccl", init_method="tcp://127.0.0.1:29715", rank=0, world_size=1)
    torch.cuda.set_device(0)
    import face_recognition_pytorch_b200 as m
    return m


def _graph_run(pfc, autograd, B=1024, C=20000, d=512, steps=4, **conf_extra):
    g = torch.Generator().manual_seed(33)
    w = torch.normal(0, 0.01, (C, d), generator=g)
    conf = types.SimpleNamespace(emd_size=d, sample_rate=1.0, mixed_precision=False, loss_s=64.0, loss_m=0.5,
                                 fused_optimizer=True, **conf_extra)
    head = pfc.PartialFC(conf, C)
    head.load_state_dict({"weight": w.clone()})
    head = head.train().cuda()
    opt = torch.optim.SGD(head.parameters(), lr=0.1, momentum=0.9, weight_decay=5e-4)
    step = pfc.GraphedHeadStep(head, opt, B, d, autograd=autograd)
    out = []
    for s in range(steps):
        lab = torch.randint(0, C, (B,), generator=g).cuda()
        x = torch.nn.functional.normalize(torch.nn.functional.normalize(w[lab.cpu()]) +
                                          1.5 * torch.randn(B, d, generator=g) / d ** 0.5).cuda()
        loss, dx = step(x, lab)
        out += [loss.detach().clone().reshape(()), dx.clone()]
    out.append(head.state_dict()["weight"].clone())
    torch.cuda.synchronize()
    return out


def test_graph_without_autograd_is_bit_identical(pfc):
    a = _graph_run(pfc, True)
    b = _graph_run(pfc, False)
    for i, (u, v) in enumerate(zip(a, b)):
        assert torch.equal(u, v), f"output {i} differs without autograd"


@pytest.mark.parametrize("B,C,d", [(1024, 20000, 512), (320, 3100, 512), (96, 1500, 64)])
def test_loss_and_coefficients_in_one_launch_is_bit_identical(pfc, B, C, d):
    """conf.fuse_prepare (row statistics + loss + backward coefficients in ONE kernel on the no-autograd path) against
    the two launches of the autograd path, ragged batch / class tails included."""
    a = _graph_run(pfc, False, B=B, C=C, d=d, fuse_prepare=True)
    b = _graph_run(pfc, False, B=B, C=C, d=d, fuse_prepare=False)
    for i, (u, v) in enumerate(zip(a, b)):
        assert torch.equal(u, v), f"output {i} differs with the fused loss + coefficient kernel"


def test_fused_step_eager_matches_autograd(pfc):
    d, B, C = 512, 320, 3100
    g = torch.Generator().manual_seed(5)
    w = torch.normal(0, 0.01, (C, d), generator=g)
    res = []
    for direct in (False, True):
        conf = types.SimpleNamespace(emd_size=d, sample_rate=1.0, mixed_precision=False, loss_s=64.0, loss_m=0.5)
        head = pfc.PartialFC(conf, C)
        head.load_state_dict({"weight": w.clone()})
        head = head.train().cuda()
        opt = torch.optim.SGD(head.parameters(), lr=0.1, momentum=0.9, weight_decay=5e-4)
        gg = torch.Generator().manual_seed(6)
        lab = torch.randint(0, C, (B,), generator=gg).cuda()
        x = torch.nn.functional.normalize(torch.randn(B, d, generator=gg)).cuda()
        if direct:
            loss, dx = head.fused_step(x, lab, opt)
            dw = head.weight_activated.grad
        else:
            xg = x.clone().requires_grad_(True)
            loss = head(xg, lab, opt)
            loss.backward()
            dx, dw = xg.grad, head.weight_activated.grad
        res.append((loss.detach().clone().reshape(()), dx.clone(), dw.clone()))
    for u, v in zip(*res):
        assert torch.equal(u, v)


def test_adamw_sampled_fused_matches_unfused_and_reference(pfc):
    import numpy as np
    from helpers import load_case, case_inputs, case_perms, cosine
    cfg, z = load_case("head_w1_adamw_sampled")
    weights, xs, ls = case_inputs(cfg)
    finals = []
    for fused in (False, True):
        conf = types.SimpleNamespace(emd_size=cfg["d"], sample_rate=cfg["sample_rate"], mixed_precision=False,
                                     loss_s=cfg["s"], loss_m=cfg["m"], fused_optimizer=fused)
        head = pfc.PartialFCAdamW(conf, cfg["C"])
        head.load_state_dict({"weight": weights[0].clone()})
        head = head.train().cuda()
        dummy = torch.nn.Parameter(torch.zeros(1).cuda())
        opt = torch.optim.AdamW([{"params": [dummy]}, {"params": head.parameters()}], lr=cfg["lr"], weight_decay=cfg["wd"])
        for s in range(cfg["steps"]):
            x = xs[s].clone().cuda().requires_grad_(True)
            opt.zero_grad()
            loss = head(x, ls[s].clone().cuda(), opt, perm=case_perms(cfg, z, s)[0].cuda())
            loss.backward()
            assert np.array_equal(head.weight_index.cpu().numpy(), z[f"r0_index_{s}"])
            opt.step()
        head.update()
        finals.append(head.weight.cpu().numpy().astype(np.float64))
    w0 = weights[0].numpy().astype(np.float64)
    ref = z["r0_weight_final"] - w0
    m_ref = np.abs(z["r0_exp_avg_final"])
    well = m_ref > 0.05 * m_ref.max()
    for f in finals:
        err = np.abs((f - w0) - ref)[well].mean() / np.abs(ref).max()
        assert err <= 5e-3, err
    assert cosine(finals[0] - w0, finals[1] - w0) >= 0.9999


@pytest.mark.parametrize("fused", [False, True])
def test_head_with_interclass_filter_matches_reference(pfc, fused):
    """tests/golden/head_w1_filter_wide.npz: every non-target cosine stays >= 0.02 away from the threshold, so bf16
    operand rounding cannot flip a filter decision."""
    import numpy as np
    from helpers import load_case, case_inputs, cosine
    cfg, z = load_case("head_w1_filter_wide")
    weights, xs, ls = case_inputs(cfg)
    thr = cfg["filter_thr"]
    conf = types.SimpleNamespace(emd_size=cfg["d"], sample_rate=1.0, mixed_precision=False, loss_s=cfg["s"],
                                 loss_m=cfg["m"], fused_optimizer=fused)
    head = pfc.PartialFC(conf, cfg["C"], margin_loss=lambda s_, m_: pfc.CombinedMarginLoss(
        s_, 1.0, m_, 0.0, interclass_filtering_threshold=thr))
    head.load_state_dict({"weight": weights[0].clone()})
    head = head.train().cuda()
    opt = torch.optim.SGD(head.parameters(), lr=cfg["lr"], momentum=cfg["momentum"], weight_decay=cfg["wd"])
    for s in range(cfg["steps"]):
        x = xs[s].clone().cuda().requires_grad_(True)
        opt.zero_grad()
        loss = head(x, ls[s].clone().cuda(), opt)
        loss.backward()
        ref_loss = float(z[f"r0_loss_{s}"])
        assert abs(float(loss.detach()) - ref_loss) <= 6e-3 * abs(ref_loss), (s, float(loss.detach()), ref_loss)
        assert cosine(x.grad.cpu(), z[f"r0_dx_{s}"]) >= 0.999
        if not fused:
            assert cosine(head.weight_activated.grad.cpu(), z[f"r0_dw_{s}"]) >= 0.999
        opt.step()
    w0 = weights[0].double()
    assert cosine(head.weight_activated.data.cpu().double() - w0,
                  torch.from_numpy(z["r0_weight_final"]).double() - w0) >= 0.999


@pytest.mark.parametrize("B,C,d,fused", [(1024, 20000, 512, True), (320, 3100, 512, False), (96, 1500, 64, True),
                                         (200, 777, 128, True), (1024, 300, 512, False)])
def test_dx_tail_fork_is_bit_identical(pfc, B, C, d, fused):
    """conf.dx_side_stream: the dX finalize on a side stream next to the dW GEMM / update -- a scheduling change only;
    conf.dx_fork_gemm: the fork in front of the dX GEMM (both gradient GEMMs side by side, the fused update waiting for
    the dX GEMM's last read of the shard) -- likewise."""
    g0 = torch.Generator().manual_seed(44)
    w = torch.normal(0, 0.01, (C, d), generator=g0)
    outs = []
    for fork, gemm in ((False, False), (True, False), (True, True)):
        g = torch.Generator().manual_seed(45)
        conf = types.SimpleNamespace(emd_size=d, sample_rate=1.0, mixed_precision=False, loss_s=64.0, loss_m=0.5,
                                     fused_optimizer=fused, dx_side_stream=fork, dx_fork_gemm=gemm)
        head = pfc.PartialFC(conf, C)
        head.load_state_dict({"weight": w.clone()})
        head = head.train().cuda()
        opt = torch.optim.SGD(head.parameters(), lr=0.1, momentum=0.9, weight_decay=5e-4)
        rec = []
        for s in range(3):
            lab = torch.randint(0, C, (B,), generator=g).cuda()
            x = torch.nn.functional.normalize(torch.nn.functional.normalize(w[lab.cpu()]) +
                                              1.5 * torch.randn(B, d, generator=g) / d ** 0.5).cuda().requires_grad_(True)
            opt.zero_grad()
            loss = head(x, lab, opt)
            loss.backward()
            rec += [loss.detach().clone(), x.grad.clone()]
            if not fused:
                rec.append(head.weight_activated.grad.clone())
                opt.step()
        rec.append(head.weight_activated.data.clone())
        torch.cuda.synchronize()
        outs.append(rec)
    for i, (u, v, t) in enumerate(zip(*outs)):
        assert torch.equal(u, v), f"output {i} differs with the dX tail forked"
        assert torch.equal(u, t), f"output {i} differs with the dX GEMM forked"


def test_forward_only_and_eval_paths(pfc):
    """No gradient wanted for the embeddings: same loss as the training path's forward; a forward whose backward never
    runs does not disturb the next step."""
    d, B, C = 128, 200, 777
    g = torch.Generator().manual_seed(9)
    w = torch.normal(0, 0.01, (C, d), generator=g)
    lab = torch.randint(0, C, (B,), generator=g).cuda()
    x = torch.nn.functional.normalize(torch.randn(B, d, generator=g)).cuda()
    conf = types.SimpleNamespace(emd_size=d, sample_rate=1.0, mixed_precision=False, loss_s=64.0, loss_m=0.5)
    head = pfc.PartialFC(conf, C)
    head.load_state_dict({"weight": w.clone()})
    head = head.train().cuda()
    opt = torch.optim.SGD(head.parameters(), lr=0.1)
    with torch.no_grad():
        l0 = head(x, lab.clone(), opt).clone()
    l1 = head(x.clone().requires_grad_(True), lab.clone(), opt)          # backward never called
    assert torch.equal(l0, l1.detach())
    xg = x.clone().requires_grad_(True)
    l2 = head(xg, lab.clone(), opt)
    l2.backward()
    assert torch.equal(l0, l2.detach()) and bool(torch.isfinite(xg.grad).all())


def test_dx_tail_fork_in_a_graph_is_bit_identical(pfc):
    a = _graph_run(pfc, False)
    b = _graph_run(pfc, False, dx_side_stream=False)
    c = _graph_run(pfc, True, B=320, C=3100)
    d = _graph_run(pfc, True, dx_side_stream=False, B=320, C=3100)
    for (u, v) in list(zip(a, b)) + list(zip(c, d)):
        assert torch.equal(u, v)


def test_adamw_in_a_graph_matches_eager(pfc):
    """PartialFCAdamW (sample_rate 1, fused) inside GraphedHeadStep: the bias-correction step count comes from a device
    scalar, so replays advance it like eager steps do."""
    d, B, C = 512, 256, 3000
    g = torch.Generator().manual_seed(77)
    w = torch.normal(0, 0.01, (C, d), generator=g)
    data = [(torch.nn.functional.normalize(torch.randn(B, d, generator=g)), torch.randint(0, C, (B,), generator=g))
            for _ in range(4)]
    outs = []
    for graphed in (False, True):
        conf = types.SimpleNamespace(emd_size=d, sample_rate=1.0, mixed_precision=False, loss_s=64.0, loss_m=0.5,
                                     fused_optimizer=True)
        head = pfc.PartialFCAdamW(conf, C)
        head.load_state_dict({"weight": w.clone()})
        head = head.train().cuda()
        opt = torch.optim.AdamW(head.parameters(), lr=1e-3, weight_decay=0.05)
        step = pfc.GraphedHeadStep(head, opt, B, d) if graphed else None
        losses = []
        for x, lab in data:
            if graphed:
                loss, _ = step(x.cuda(), lab.cuda())
            else:
                xg = x.clone().cuda().requires_grad_(True)
                loss = head(xg, lab.clone().cuda(), opt)
                loss.backward()
            losses.append(float(loss.detach()))
        outs.append((losses, head.weight_activated.data.clone(), head.step))
    # the replayed kernels compute 1 - beta^t on the device (double), the eager step on the host: same formula, but the two
    # pow() implementations may differ in the last bit of the float the kernels finally use
    for a, b in zip(outs[0][0], outs[1][0]):
        assert abs(a - b) <= 1e-6 * abs(a)
    assert outs[0][2] == outs[1][2] == 4
    torch.testing.assert_close(outs[0][1], outs[1][1], rtol=0, atol=2e-7)


def test_l2_resident_gradient_is_bit_identical(pfc):
    from face_recognition_pytorch_b200 import _lib
    a = _graph_run(pfc, True)
    _lib.lib.pfc_debug_l2_grad(0)
    try:
        b = _graph_run(pfc, True)
        c = _graph_run(pfc, True, B=320, C=3100)          # odd tile counts, ragged last class tile
    finally:
        _lib.lib.pfc_debug_l2_grad(1)
    d = _graph_run(pfc, True, B=320, C=3100)
    for u, v in zip(a, b):
        assert torch.equal(u, v)
    for u, v in zip(c, d):
        assert torch.equal(u, v)


@pytest.mark.parametrize("adamw", [False, True])
@pytest.mark.parametrize("B,C,d", [(256, 3100, 512), (96, 1500, 64)])
def test_amp_update_writes_the_bf16_twin_of_the_shard(pfc, adamw, B, C, d):
    """conf.mixed_precision (fp16 operands): the dX contraction needs the shard as bf16.  The fused update writes that twin
    next to the fp16 rows (specialised SGD kernel at d = 512, generic row kernel at d = 64 and for AdamW), bit for bit what
    pfc_cast_f16_to_bf16 makes of them, so the cast pass runs on the first step only -- and the steps are bit-identical to
    steps that cast every time."""
    from face_recognition_pytorch_b200 import partial_fc as PF
    g = torch.Generator().manual_seed(91)
    w = torch.normal(0, 0.01, (C, d), generator=g)
    data = [(torch.nn.functional.normalize(torch.randn(B, d, generator=g)).cuda(),
             torch.randint(0, C, (B,), generator=g).cuda()) for _ in range(4)]
    runs = []
    for always_cast in (False, True):
        conf = types.SimpleNamespace(emd_size=d, sample_rate=1.0, mixed_precision=True, loss_s=64.0, loss_m=0.5,
                                     fused_optimizer=True)
        head = (pfc.PartialFCAdamW if adamw else pfc.PartialFC)(conf, C)
        head.load_state_dict({"weight": w.clone()})
        head = head.train().cuda()
        opt = (torch.optim.AdamW(head.parameters(), lr=1e-3, weight_decay=0.05) if adamw else
               torch.optim.SGD(head.parameters(), lr=0.1, momentum=0.9, weight_decay=5e-4))
        cast, calls, out = PF.K.cast_f16_to_bf16, [], []

        def counted_cast(*a, **k):
            calls.append(1)
            return cast(*a, **k)
        PF.K.cast_f16_to_bf16 = counted_cast
        try:
            for x, lab in data:
                if always_cast:
                    head._wn_b_valid = False
                loss, dx = head.fused_step(x, lab.clone(), opt)
                ws = head._ws
                assert ws.wn.dtype == torch.float16 and ws.wn_b.dtype == torch.bfloat16
                assert torch.equal(ws.wn_b[:C], ws.wn[:C].to(torch.bfloat16)), "twin differs from the cast of the shard"
                out += [loss.clone(), dx.clone()]
        finally:
            PF.K.cast_f16_to_bf16 = cast
        assert len(calls) == (len(data) if always_cast else 1)
        out.append(head.weight_activated.data.clone())
        runs.append(out)
    for i, (u, v) in enumerate(zip(*runs)):
        assert torch.equal(u, v), f"output {i}: twin written by the update vs. cast every step"
