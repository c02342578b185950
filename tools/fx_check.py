"""First-contact check of the forward + dX kernel (pfc_forward_dx) and the ordered lazy update on a GPU: small shapes
first, each in its own process under `timeout`, so that a protocol bug (bounded waits trap after ~2 s) cannot hang
the box.  Usage: python tools/fx_check.py [case]"""
import os
import subprocess
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = {
    "tiny": dict(B=48, C=777, d=128), "d64": dict(B=96, C=1500, d=64), "mid": dict(B=320, C=3100, d=512),
    "wide": dict(B=1024, C=20000, d=512), "full": dict(B=1024, C=93431, d=512),
}


def run(name):
    import torch
    import torch.distributed as dist
    dist.init_process_group("nccl", init_method="tcp://127.0.0.1:29733", rank=0, world_size=1)
    torch.cuda.set_device(0)
    import face_recognition_pytorch_b200 as pfc
    c = CASES[name]
    B, C, d = c["B"], c["C"], c["d"]
    g = torch.Generator().manual_seed(3)
    w = torch.normal(0, 0.01, (C, d), generator=g)
    res = {}
    for mode in ("nofx", "fx", "lazy"):
        if mode == "lazy" and d % 128:
            continue
        conf = types.SimpleNamespace(emd_size=d, sample_rate=1.0, mixed_precision=False, loss_s=64.0, loss_m=0.5,
                                     fused_optimizer=True, fx=mode != "nofx", lazy_update=mode == "lazy")
        head = pfc.PartialFC(conf, C)
        head.load_state_dict({"weight": w.clone()})
        head = head.train().cuda()
        opt = torch.optim.SGD(head.parameters(), lr=0.1, momentum=0.9, weight_decay=5e-4)
        gg = torch.Generator().manual_seed(4)
        out = []
        for s in range(3):
            lab = torch.randint(0, C, (B,), generator=gg)
            x = torch.nn.functional.normalize(torch.nn.functional.normalize(w[lab]) +
                                              1.5 * torch.randn(B, d, generator=gg) / d ** 0.5).cuda().requires_grad_(True)
            loss = head(x, lab.cuda(), opt)
            loss.backward()
            torch.cuda.synchronize()
            out.append((float(loss), x.grad.clone()))
        out.append(head.state_dict()["weight"].clone())
        torch.cuda.synchronize()
        res[mode] = out
        print(name, mode, "losses", [round(o[0], 6) for o in out[:3]], flush=True)
    ref = res["nofx"]
    for mode, out in res.items():
        if mode == "nofx":
            continue
        for s in range(3):
            a, b = ref[s][1].double().flatten(), out[s][1].double().flatten()
            cos = float(a @ b / (a.norm() * b.norm()))
            print(f"  {mode} step {s}: dloss {abs(ref[s][0] - out[s][0]):.2e} cos(dX) {cos:.8f} "
                  f"max|ddX| {float((a - b).abs().max()):.2e} / {float(a.abs().max()):.2e}")
        dw = (ref[3] - out[3]).abs().max()
        print(f"  {mode} weights: max diff {float(dw):.2e} (|w - w0| max {float((ref[3] - w.cuda()).abs().max()):.2e})")


if __name__ == "__main__":
    if len(sys.argv) > 1:
        run(sys.argv[1])
    else:
        for name in CASES:
            r = subprocess.run(["timeout", "120", sys.executable, __file__, name], capture_output=True, text=True)
            print(r.stdout[-3000:])
            if r.returncode != 0:
                print(f"CASE {name} FAILED rc={r.returncode}\n{r.stderr[-3000:]}")
                break
