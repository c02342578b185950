"""BASELINE.json configs[0] on the GPU against the fixture the unmodified reference produced on the CPU.  Added after the
round's GPU budget was spent (its body ran on the CPU through tests/fake_kernels.py, tools/sim_gpu_tests.py); the file
name sorts last so that it runs after every GPU test that has already been seen green on a B200."""
import types

import pytest
import torch

from helpers import cosine

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pfc():
    import torch.distributed as dist
    if not dist.is_initialized():
        dist.init_process_group("nccl", init_method="tcp://127.0.0.1:29716", rank=0, world_size=1)
    torch.cuda.set_device(0)
    import face_recognition_pytorch_b200 as m
    return m


@pytest.mark.parametrize("fused", [False, True])
def test_cfg1_shape_against_the_reference_fixture(pfc, fused):
    """BASELINE.json configs[0]: batch 128, 10 000 classes, d = 512, s = 64, m = 0.5, one process -- against the fixture
    the unmodified reference produced on the CPU (tests/golden/head_cfg1.npz; [C, d] arrays as norm + projection).
    A CPU restatement of the kernels with bf16 operands lands at loss 7e-5 / cosine 0.99994 on this fixture."""
    import numpy as np
    from helpers import load_case, case_inputs
    from inputs import proj_matrix
    cfg, z = load_case("head_cfg1")
    weights, xs, ls = case_inputs(cfg)
    R = proj_matrix(cfg["d"])
    conf = types.SimpleNamespace(emd_size=cfg["d"], sample_rate=1.0, mixed_precision=False, loss_s=cfg["s"],
                                 loss_m=cfg["m"], fused_optimizer=fused)
    head = pfc.PartialFC(conf, cfg["C"])
    head.load_state_dict({"weight": weights[0].clone()})
    head = head.train().cuda()
    opt = torch.optim.SGD(head.parameters(), lr=cfg["lr"], momentum=cfg["momentum"], weight_decay=cfg["wd"])
    for s in range(cfg["steps"]):
        x = xs[s].clone().cuda().requires_grad_(True)
        opt.zero_grad()
        loss = head(x, ls[s].clone().cuda(), opt)
        loss.backward()
        ref_loss = float(z[f"r0_loss_{s}"])
        assert abs(float(loss.detach()) - ref_loss) <= 1e-3 * abs(ref_loss), (s, float(loss.detach()), ref_loss)
        assert cosine(x.grad.cpu(), z[f"r0_dx_{s}"]) >= 0.999
        assert abs(float(x.grad.norm()) / np.linalg.norm(z[f"r0_dx_{s}"]) - 1) < 2e-2
        if not fused:
            dw = head.weight_activated.grad.cpu().double().numpy()
            assert cosine(dw @ R, z[f"r0_dw_{s}_proj"]) >= 0.999
            assert abs(np.linalg.norm(dw) / float(z[f"r0_dw_{s}_norm"]) - 1) < 2e-2
        opt.step()
    w0 = weights[0].double().numpy()
    wf = head.weight_activated.data.cpu().double().numpy()
    assert cosine((wf - w0) @ R, z["r0_weight_final_proj"].astype(np.float64) - w0 @ R) >= 0.999
