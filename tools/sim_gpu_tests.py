"""DEV TOOL: run the BODIES of selected GPU tests on the CPU, with tests/fake_kernels.py standing in for the CUDA
kernels and .cuda() patched to the identity.  It catches Python-level mistakes in GPU tests (and in the host logic they
drive) when no GPU is at hand; it says nothing about the kernels.  Not part of the test-suite.

    python tools/sim_gpu_tests.py [substring of the test names to run]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HERE = os.path.join(ROOT, "tests")
for p in (ROOT, HERE, os.path.join(HERE, "golden")):
    sys.path.insert(0, p)

import torch                                   # noqa: E402
import torch.distributed as dist               # noqa: E402

torch.Tensor.cuda = lambda self, *a, **k: self
torch.nn.Module.cuda = lambda self, *a, **k: self
torch.cuda.synchronize = lambda *a, **k: None
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:29879", rank=0, world_size=1)
import face_recognition_pytorch_b200 as pfc    # noqa: E402
from face_recognition_pytorch_b200 import partial_fc, kernels   # noqa: E402
from fake_kernels import FakeKernels           # noqa: E402

partial_fc.K = FakeKernels(kernels)
import test_gpu_z_cfg1 as tc                   # noqa: E402
import test_gpu_modes as te                    # noqa: E402
import test_gpu_head as th                     # noqa: E402

RUNS = [
    (tc.test_cfg1_shape_against_the_reference_fixture, [(pfc, False), (pfc, True)]),
    (te.test_fused_step_eager_matches_autograd, [(pfc,)]),
    (te.test_adamw_sampled_fused_matches_unfused_and_reference, [(pfc,)]),
    (te.test_head_with_interclass_filter_matches_reference, [(pfc, False), (pfc, True)]),
    (te.test_dx_tail_fork_is_bit_identical, [(pfc, 320, 3100, 512, False), (pfc, 96, 1500, 64, True)]),
    (te.test_forward_only_and_eval_paths, [(pfc,)]),
    (te.test_amp_update_writes_the_bf16_twin_of_the_shard, [(pfc, False, 256, 3100, 512), (pfc, True, 96, 1500, 64)]),
    (th.test_steps_match_reference_and_oracle, [(pfc, "head_w1_d128", "fused"),
                                                (pfc, "head_w1_sampled", "fused"), (pfc, "head_w1_full", "unfused")]),
    (th.test_scaled_loss_through_the_kernels, [(pfc, "unfused"), (pfc, "fused")]),
]
failed = 0
ONLY = sys.argv[1] if len(sys.argv) > 1 else ""
for fn, arglists in RUNS:
    if ONLY not in fn.__name__:
        continue
    for args in arglists:
        try:
            fn(*args)
            print("ok    ", fn.__name__, args[1:], flush=True)
        except Exception as e:   # noqa: BLE001
            failed += 1
            print("FAILED", fn.__name__, args[1:], type(e).__name__, str(e)[:300], flush=True)
sys.exit(1 if failed else 0)
