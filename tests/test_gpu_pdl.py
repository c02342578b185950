"""Programmatic dependent launch (csrc/pfc_launch.cuh) only changes WHEN a step kernel's CTAs are scheduled, never
what they read: every kernel reaches global memory after griddepcontrol.wait, and every reduction of the step has a
fixed order.  So the same steps with the switch on (mode 1, and mode 2 = deferred waits) and off must agree BIT FOR BIT -- eagerly launched and replayed
from a CUDA graph, un-fused and fused update, SGD and AdamW, d = 512 fast paths and the generic d = 64 ones."""
import types

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pfc():
    import torch.distributed as dist
    if not dist.is_initialized():
        dist.init_process_group("nccl", init_method="tcp://127.0.0.1:29714", rank=0, world_size=1)
    torch.cuda.set_device(0)
    import face_recognition_pytorch_b200 as m
    return m


def _run(pfc, pdl, B, C, d, fused, adamw, steps=3, graph=False):
    from face_recognition_pytorch_b200 import kernels as K
    prev = K.set_pdl(pdl)
    try:
        g = torch.Generator().manual_seed(21)
        w = torch.normal(0, 0.01, (C, d), generator=g)
        conf = types.SimpleNamespace(emd_size=d, sample_rate=1.0, mixed_precision=False, loss_s=64.0, loss_m=0.5,
                                     fused_optimizer=fused)
        head = (pfc.PartialFCAdamW if adamw else pfc.PartialFC)(conf, C)
        head.load_state_dict({"weight": w.clone()})
        head = head.train().cuda()
        if adamw:
            opt = torch.optim.AdamW(head.parameters(), lr=1e-3, weight_decay=0.05)
        else:
            opt = torch.optim.SGD(head.parameters(), lr=0.1, momentum=0.9, weight_decay=5e-4)
        step = pfc.GraphedHeadStep(head, opt, B, d) if graph else None
        out = []
        for s in range(steps):
            lab = torch.randint(0, C, (B,), generator=g).cuda()
            x = torch.nn.functional.normalize(torch.nn.functional.normalize(w[lab.cpu()]) +
                                              1.5 * torch.randn(B, d, generator=g) / d ** 0.5).cuda()
            if graph:
                loss, dx = step(x, lab)
                out += [loss.detach().clone(), dx.clone()]
            else:
                xg = x.clone().requires_grad_(True)
                opt.zero_grad()
                loss = head(xg, lab, opt)
                loss.backward()
                out += [loss.detach().clone(), xg.grad.clone()]
                if not fused:
                    out.append(head.weight_activated.grad.clone())
                    opt.step()
        out.append(head.weight_activated.data.clone())
        torch.cuda.synchronize()
        return out
    finally:
        K.set_pdl(prev)


@pytest.mark.parametrize("mode", [1, 2])
@pytest.mark.parametrize("B,C,d,fused,adamw", [
    (1024, 20000, 512, True, False),     # the bench configuration's kernels (pair GEMMs, d512 finalize, fused SGD rows)
    (320, 3100, 512, False, False),      # odd tile counts, un-fused: dw_finalize + torch.optim.SGD between the steps
    (96, 1500, 64, True, True),          # generic row kernels, fused AdamW
])
def test_pdl_on_off_bit_identical_eager(pfc, mode, B, C, d, fused, adamw):
    off = _run(pfc, 0, B, C, d, fused, adamw)
    on = _run(pfc, mode, B, C, d, fused, adamw)
    assert len(on) == len(off)
    for i, (a, b) in enumerate(zip(on, off)):
        assert torch.equal(a, b), f"output {i} differs with PDL on (max abs diff {(a - b).abs().max().item():.3e})"


@pytest.mark.parametrize("mode", [1, 2])
def test_pdl_on_off_bit_identical_graph_replay(pfc, mode):
    """The graph captures the PDL launches as programmatic edges; five replays against five replays without.  Mode 2
    additionally lets the dX GEMM start under the tail of the dW GEMM (deferred wait)."""
    off = _run(pfc, 0, 1024, 20000, 512, True, False, steps=5, graph=True)
    on = _run(pfc, mode, 1024, 20000, 512, True, False, steps=5, graph=True)
    for i, (a, b) in enumerate(zip(on, off)):
        assert torch.equal(a, b), f"output {i} differs with PDL on"


def test_pdl_switch_round_trips(pfc):
    from face_recognition_pytorch_b200 import kernels as K
    prev = K.set_pdl(1)
    assert K.get_pdl() == 1
    assert K.set_pdl(2) == 1
    assert K.set_pdl(7) == 2 and K.get_pdl() == 2      # clamped
    assert K.set_pdl(0) == 2 and K.get_pdl() == 0
    K.set_pdl(prev)
