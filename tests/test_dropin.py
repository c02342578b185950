"""The drop-in shims (face-recognition-pytorch_b200/dropin/) resolve exactly like the reference's call sites:
`importlib.import_module(f"nets.{conf.loss}").PartialFC / .PartialFCAdamW` (model/FR_PartialFC.py:102-109) and
`from utils.eval import performance_roc, cross_score, pair_score, performance_acc` (model/FR_PartialFC.py:13).
CPU only: importing and constructing, no kernel call."""
import importlib
import inspect
import os
import sys
import types

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DROPIN = os.path.join(ROOT, "face-recognition-pytorch_b200", "dropin")


@pytest.fixture()
def dropin_path():
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k.split(".")[0] in ("nets", "utils")}
    sys.path.insert(0, DROPIN)
    try:
        yield
    finally:
        sys.path.remove(DROPIN)
        for k in list(sys.modules):
            if k.split(".")[0] in ("nets", "utils"):
                del sys.modules[k]
        sys.modules.update(saved)


def test_head_modules_resolve_by_name(dropin_path):
    import face_recognition_pytorch_b200 as pfc
    for loss in ("PartialFC",):                                    # conf.loss in configs/ms1m_arcface_122.py
        mod = importlib.import_module(f"nets.{loss}")
        assert mod.PartialFC is pfc.PartialFC and mod.PartialFCAdamW is pfc.PartialFCAdamW
        # the reference constructs them with keywords conf=, num_classes= (and margin_loss defaulting to ArcFace)
        for cls in (mod.PartialFC, mod.PartialFCAdamW):
            params = list(inspect.signature(cls.__init__).parameters)
            assert params[:4] == ["self", "conf", "num_classes", "margin_loss"]
            assert inspect.signature(cls.__init__).parameters["margin_loss"].default is pfc.ArcFace
            fwd = list(inspect.signature(cls.forward).parameters)
            assert fwd[:4] == ["self", "local_embeddings", "local_labels", "optimizer"]
    arc = importlib.import_module("nets.ArcFace")
    assert arc.ArcFace is pfc.ArcFace and arc.CosFace is pfc.CosFace and arc.CombinedMarginLoss is pfc.CombinedMarginLoss
    assert list(inspect.signature(arc.ArcFace.__init__).parameters) == ["self", "s", "margin"]
    assert list(inspect.signature(arc.CosFace.__init__).parameters) == ["self", "s", "m"]
    assert list(inspect.signature(arc.CombinedMarginLoss.__init__).parameters) == [
        "self", "s", "m1", "m2", "m3", "interclass_filtering_threshold"]


def test_eval_functions_resolve_by_name(dropin_path):
    from utils.eval import performance_roc, cross_score, pair_score, performance_acc   # model/FR_PartialFC.py:13
    import face_recognition_pytorch_b200 as pfc
    assert pair_score is pfc.pair_score and cross_score is pfc.cross_score
    assert performance_roc is pfc.performance_roc and performance_acc is pfc.performance_acc
    # argument names of utils/eval.py:7, :54, :68, :102
    assert list(inspect.signature(performance_roc).parameters) == ["hist_genuine", "hist_imposter", "min_level", "max_level"]
    assert list(inspect.signature(performance_acc).parameters) == ["score_list", "label_list", "th"]
    assert list(inspect.signature(pair_score).parameters)[:6] == ["embedding_1", "embedding_2", "labels", "metric",
                                                                  "min_level", "max_level"]
    assert list(inspect.signature(cross_score).parameters) == ["embeddings", "labels", "metric"]


def test_constructor_needs_initialised_process_group(dropin_path):
    """nets/PartialFC.py:47-49: assert distributed.is_initialized()."""
    import torch.distributed as dist
    if dist.is_initialized():
        pytest.skip("a process group is already up in this process")
    mod = importlib.import_module("nets.PartialFC")
    conf = types.SimpleNamespace(emd_size=64, sample_rate=1.0, mixed_precision=False, loss_s=64.0, loss_m=0.5)
    with pytest.raises(AssertionError):
        mod.PartialFC(conf=conf, num_classes=100)


def test_amp_twin_host_logic_on_fake_kernels():
    """Host logic of the AMP bf16 twin (which step casts, which step lets the fused update write it, bit-identity of the
    two) with tests/fake_kernels.py standing in for the CUDA kernels: the body of the GPU test, run on the CPU in a
    separate process (tools/sim_gpu_tests.py patches .cuda() to the identity, which must not leak into this one)."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "sim_gpu_tests.py"), "test_amp_update_writes"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("ok ") == 2 and "FAILED" not in r.stdout, r.stdout
