"""Sampler with the CTA-wide pick kernel (PFC_SAMPLE_PICK=parallel): the six oracle cases of gpu_probe.case_sample
(forced ties, num_sample = 0, num_sample = num_local, cfg-3 / cfg-4 rank shapes), then device time of both variants."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch                                        # noqa: E402
from face_recognition_pytorch_b200 import _lib, kernels as K   # noqa: E402
from tools import gpu_probe                          # noqa: E402

ok = gpu_probe.case_sample()          # runs the oracle cases with both pick kernels
print("PARALLEL_PICK_PARITY", "OK" if ok else "MISMATCH", flush=True)
for nl, k, B in ((45029, 4502, 1024), (257489, 51497, 4096)):
    g = torch.Generator().manual_seed(1)
    perm = torch.rand(nl, generator=g).cuda()
    lab = torch.randint(-1, nl, (B,), generator=g).to(torch.int32).cuda()
    idx = torch.zeros(max(k, B), dtype=torch.int64, device="cuda")
    n_out = torch.zeros(1, dtype=torch.int32, device="cuda")
    rem = torch.zeros(B, dtype=torch.int32, device="cuda")
    ws = torch.zeros(K.sample_workspace_bytes(nl), dtype=torch.uint8, device="cuda")
    for par in (0, 1, 2):       # 2 = the one-launch sampler (PFC_SAMPLE_FUSED, shards of up to 65 536 classes)
        _lib.lib.pfc_debug_sample_pick(1 if par else 0)
        _lib.lib.pfc_debug_sample_fused(1 if par == 2 else 0)
        ts = []
        for i in range(8):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            K.sample(perm, lab, nl, k, idx, n_out, rem, ws)
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) * 1e3)
        print(f"pfc_sample nl={nl} k={k} pick={('serial', 'parallel', 'fused one-launch')[par]}: median {sorted(ts[2:])[3]:.1f} us", flush=True)
_lib.lib.pfc_debug_sample_fused(0)
