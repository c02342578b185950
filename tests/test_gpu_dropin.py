"""Drop-in through the reference's OWN call site: `Model.training_step` of
model/FR_PartialFC.py:162-193 -- constructor `importlib.import_module(f"nets.{conf.loss}").PartialFC(conf=..., num_classes=...)`
(:102-109), `.train().to(local_rank)` (:111), `self.loss.parameters()` as the optimizer's LAST param group (:438-449),
`loss = self.loss(feat, id_, self.opt)` (:175), `loss.backward()` / the GradScaler flow (:178-188) -- is run twice on the same
seeded data: once with the reference's head (nets/PartialFC.py, unmodified, as vendored into oracle/_ref by
oracle/make_ref.py) and once with this package's head substituted purely through sys.path (dropin/ in front).  Same encoder
(a two-layer stub standing in for nets/resnet.py), same optimizer, same CPU sampling draws.  Compared per step: the loss, and
after the steps: the encoder's weights (they receive dX through the head's autograd edge) and the head's state_dict.

Needs a GPU (the reference head hard-codes .cuda(), the Model creates CUDA events and wraps the encoder in DDP) and
oracle/_ref (git-ignored build output; absent -> skipped).  torchmetrics / torchsummary / easydict are not in this image:
three stub modules stand in for them."""
import importlib
import os
import sys
import textwrap

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.path.join(ROOT, "oracle", "_ref")
DROPIN = os.path.join(ROOT, "face-recognition-pytorch_b200", "dropin")

STUBS = {
    "easydict.py": """
        class EasyDict(dict):
            def __getattr__(self, k):
                try:
                    return self[k]
                except KeyError:
                    raise AttributeError(k)
            def __setattr__(self, k, v):
                self[k] = v
    """,
    "torchmetrics.py": """
        class Accuracy:
            def __init__(self, *a, **k): pass
            def __call__(self, *a, **k): return 0.0
    """,
    "torchsummary.py": """
        def summary(*a, **k): pass
    """,
    # stands in for nets/resnet.py (the backbone is out of scope): flatten -> linear -> relu -> linear to emd_size
    "nets/resnet.py": """
        import torch
        class Encoder(torch.nn.Module):
            def __init__(self, conf):
                super().__init__()
                k = 3 * conf.img_size * conf.img_size
                self.f1 = torch.nn.Linear(k, 256)
                self.f2 = torch.nn.Linear(256, conf.emd_size)
            def forward(self, x):
                return self.f2(torch.relu(self.f1(x.flatten(1))))
    """,
}


def _purge():
    for k in list(sys.modules):
        if k.split(".")[0] in ("nets", "utils", "model", "easydict", "torchmetrics", "torchsummary"):
            del sys.modules[k]
    importlib.invalidate_caches()


def _run_model(stub_dir, use_ours, optimizer, sample_rate, amp, steps, C, b, d, img):
    """Builds the reference's Model and runs `steps` training_step calls; returns (losses, encoder weights, head weight)."""
    _purge()
    saved_path = list(sys.path)
    saved_cuda = torch.Tensor.cuda
    # namespace packages `nets` / `utils` span these directories in order: with dropin/ first, nets.PartialFC, nets.ArcFace and
    # utils.eval resolve to this package's shims, everything else (utils.logger, utils.scheduler, model.*) to the reference
    sys.path[:0] = ([DROPIN] if use_ours else []) + [REF, stub_dir]
    try:
        from easydict import EasyDict as edict
        M = importlib.import_module("model.FR_PartialFC")
        head_mod = importlib.import_module("nets.PartialFC")
        assert ("face-recognition-pytorch_b200" in head_mod.__file__) == use_ours, head_mod.__file__
        conf = edict(lr=0.05 if optimizer == "SGD" else 1e-3, security_level=3, max_level=9, min_level=3, val_dataset=[], network="ResNet18", ckpt_path=None,
                     local_rank=0, optimizer=optimizer, loss="PartialFC", n_classes=C, img_size=img, mixed_precision=amp,
                     wd=5e-4, eps=1e-8, betas=(0.9, 0.999), mom=0.9, lr_scheduler="MultiStep",
                     num_epoch=10, warmup_steps=1, min_lr=1e-4, emd_size=d, sample_rate=sample_rate, loss_s=64.0,
                     loss_m=0.5, lr_decay_ratio=0.1, lr_decay_epoch=[100], lr_decay_epoch_size=5)
        torch.manual_seed(1234)
        torch.cuda.manual_seed(1234)
        model = M.Model(conf, logger=os.path.join(stub_dir, "log.txt"), stage="train")
        # identical class centres for both heads (their initialisers consume the generator differently)
        g = torch.Generator().manual_seed(99)
        w0 = torch.normal(0, 0.01, (C, d), generator=g)
        model.loss.load_state_dict({"weight": w0.clone().cuda()})
        data_g = torch.Generator().manual_seed(7)
        enc0 = torch.cat([p.detach().flatten().cpu() for p in model.encoder.parameters()])
        losses = []
        torch.manual_seed(4321)                       # the sampling draws come from the CPU default generator (:110)
        for s in range(steps):
            imgs = torch.randn(b, 3, img, img, generator=data_g)
            ids = torch.randint(0, C, (b,), generator=data_g)
            out = model.training_step((imgs, ids))
            losses.append(float(out["loss"]))
        torch.cuda.synchronize()
        if sample_rate < 1 and hasattr(model.loss, "update"):
            model.loss.update()                       # scatter the last step's rows back (nets/PartialFC.py:133-143)
        enc = torch.cat([p.detach().flatten().cpu() for p in model.encoder.parameters()])
        head_w = model.loss.state_dict()["weight"].detach().cpu().clone()
        return losses, enc - enc0, head_w, w0, enc0
    finally:
        sys.path[:] = saved_path
        torch.Tensor.cuda = saved_cuda
        _purge()


def _cos(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float(a @ b / (a.norm() * b.norm()).clamp_min(1e-300))


@pytest.mark.parametrize("optimizer,sample_rate,amp", [("SGD", 1.0, False), ("SGD", 0.5, False), ("AdamW", 1.0, False),
                                                       ("SGD", 1.0, True)])
def test_training_step_with_the_head_swapped_in(tmp_path, optimizer, sample_rate, amp):
    if not os.path.exists(os.path.join(REF, "model", "FR_PartialFC.py")):
        pytest.skip("oracle/_ref not built (python oracle/make_ref.py where /root/reference exists)")
    import torch.distributed as dist
    if not dist.is_initialized():
        dist.init_process_group("nccl", init_method="tcp://127.0.0.1:29719", rank=0, world_size=1)
    torch.cuda.set_device(0)
    stub_dir = str(tmp_path)
    for name, body in STUBS.items():
        path = os.path.join(stub_dir, name)
        os.makedirs(os.path.dirname(path), exist_ok=True)
        with open(path, "w") as fh:
            fh.write(textwrap.dedent(body))
    C, b, d, img, steps = 2000, 64, 512, 8, 3
    ours = _run_model(stub_dir, True, optimizer, sample_rate, amp, steps, C, b, d, img)
    # The comparator is the reference in fp32.  With conf.mixed_precision the reference rounds its logits to fp16 and its
    # GradScaler may skip steps; mathematically the loss scale is a no-op, and this package's head computes the same numbers
    # with and without it -- so the GradScaler flow through our head must reproduce the reference's fp32 run.  The
    # reference's own mixed-precision run is only compared on the loss (its fp16 rounding: 1e-2).
    ref = _run_model(stub_dir, False, optimizer, sample_rate, False, steps, C, b, d, img)
    for s, (lr, lo) in enumerate(zip(ref[0], ours[0])):
        assert np.isfinite(lo) and abs(lo - lr) <= 1e-3 * abs(lr), (s, lo, lr)
    if amp:
        ref_amp = _run_model(stub_dir, False, optimizer, sample_rate, True, steps, C, b, d, img)
        assert abs(ours[0][0] - ref_amp[0][0]) <= 1e-2 * abs(ref_amp[0][0])
        print("reference fp16-AMP vs reference fp32: losses", ref_amp[0], ref[0], "head-update cosine",
              _cos(ref_amp[2] - ref_amp[3], ref[2] - ref[3]))
    # the head moved like the reference's: direction and size of the total update of the class centres
    w0 = ref[3]
    assert torch.equal(w0, ours[3])
    du_ref, du_ours = ref[2] - w0, ours[2] - w0
    # (Adam divides by sqrt(v): where a gradient entry is ~0 its sign, hence that entry's step, is rounding noise)
    cos_min = 0.97 if optimizer == "AdamW" else 0.999
    assert _cos(du_ours, du_ref) >= cos_min
    assert abs(float(du_ours.norm() / du_ref.norm()) - 1) < 5e-2
    # ... and so did the encoder, which only sees the head through dX (same initial weights, compare the updates)
    assert torch.equal(ours[4], ref[4]) and bool(torch.isfinite(ours[1]).all())
    assert _cos(ours[1], ref[1]) >= cos_min
    assert abs(float(ours[1].norm() / ref[1].norm()) - 1) < 5e-2
