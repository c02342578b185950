"""EXPERIMENT: loop each hot kernel of the step alone (and cuBLAS on the same GEMM shapes) for ~1.5 s and report
time per launch together with the SM clock and board power nvidia-smi saw meanwhile -- tells a power-capped kernel
from a pipe-/bandwidth-bound one.     python tools/exp_power.py [--gemm-mode M]"""
import argparse
import os
import subprocess
import sys
import threading
import time
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


class Smi:
    def __init__(self):
        self.rows = []
        self.proc = subprocess.Popen(["nvidia-smi", "-i", "0", "--query-gpu=clocks.sm,power.draw",
                                      "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
        threading.Thread(target=self._rd, daemon=True).start()

    def _rd(self):
        for line in self.proc.stdout:
            try:
                a, b = line.split(",")
                self.rows.append((time.perf_counter(), float(a), float(b)))
            except Exception:
                pass

    def window(self, t0, t1):
        r = [x for x in self.rows if t0 + 0.4 <= x[0] <= t1]
        if not r:
            return None, None
        return sum(x[1] for x in r) / len(r), sum(x[2] for x in r) / len(r)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gemm-mode", type=int, default=0)
    ap.add_argument("--secs", type=float, default=1.5)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    dist.init_process_group("nccl", init_method="tcp://127.0.0.1:29547", rank=0, world_size=1, device_id=dev)
    import bench
    import face_recognition_pytorch_b200 as pfc
    from face_recognition_pytorch_b200 import kernels as K
    if args.gemm_mode:
        pfc._lib.lib.pfc_debug_cluster(args.gemm_mode)
    w_shard, xs, ls = bench.synth(0, 1, 2, dev)
    conf = types.SimpleNamespace(emd_size=512, sample_rate=1.0, mixed_precision=False, loss_s=64.0, loss_m=0.5,
                                 fused_optimizer=True)
    head = pfc.PartialFC(conf, bench.C_CLASSES)
    head.load_state_dict({"weight": w_shard})
    head = head.train().cuda()
    dummy = torch.nn.Parameter(torch.zeros(1, device=dev))
    opt = torch.optim.SGD([{"params": [dummy]}, {"params": head.parameters()}], lr=0.1, momentum=0.9, weight_decay=5e-4)
    for i in range(2):
        x = xs[i].to(dev).requires_grad_(True)
        head(x, ls[i].to(dev), opt).backward()
    torch.cuda.synchronize()
    ws = head._ws
    B, n, d = ws.B, head._n, 512
    n_pad = head._n_pad
    kind, s, m2, m3, thr = head.margin_softmax.margin_spec()
    w = head.weight_activated.data
    mom = head._fused_state
    splits = K.dx_splits(B, n, d)
    a_bf = torch.randn(B, d, device=dev, dtype=torch.bfloat16)
    w_bf = torch.randn(n, d, device=dev, dtype=torch.bfloat16)
    e_bf = torch.randn(B, n_pad, device=dev, dtype=torch.bfloat16)
    out_s = torch.empty(B, n, device=dev, dtype=torch.bfloat16)
    out_dx = torch.empty(B, d, device=dev, dtype=torch.bfloat16)
    out_dw = torch.empty(n, d, device=dev, dtype=torch.bfloat16)
    big_a = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16)
    big_b = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16)
    big_c = torch.empty(8192, 8192, device=dev, dtype=torch.bfloat16)
    cases = {
        "pfc_forward": lambda: K.forward(ws.xn_all, ws.wn, ws.labels_act, B, n, d, s, kind, m2, m3, thr, ws.E, n_pad,
                                         ws.part_sum, ws.tgt_raw, ws.tgt_e, ws.tgt_z),
        "pfc_backward_dx": lambda: K.backward_dx(ws.E, n_pad, ws.wn, B, n, d, ws.dx_partial, splits),
        "pfc_backward_dw": lambda: K.backward_dw(ws.E, n_pad, ws.xs, B, n, d, ws.dwn_bf16),
        "pfc_dw_sgd": lambda: K.dw_sgd(ws.dwn_bf16, w, mom, ws.inv_w, n, d, 0.0, 0.9, 5e-4, 1.0, ws.wn, ws.inv_w),
        "cublas_fwd  [B,d]x[d,n]": lambda: torch.matmul(a_bf, w_bf.t(), out=out_s),
        "cublas_dx   [B,n]x[n,d]": lambda: torch.matmul(e_bf[:, :n], w_bf, out=out_dx),
        "cublas_dw   [n,B]x[B,d]": lambda: torch.matmul(e_bf[:, :n].t(), a_bf, out=out_dw),
        "cublas_8k^3": lambda: torch.matmul(big_a, big_b, out=big_c),
    }
    flops = {k: 2.0 * B * n * d for k in cases}
    flops["cublas_8k^3"] = 2.0 * 8192 ** 3
    flops["pfc_dw_sgd"] = 0
    smi = Smi()
    time.sleep(0.5)
    for name, fn in cases.items():
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        count = 0
        e0.record()
        while time.perf_counter() - t0 < args.secs:
            for _ in range(50):
                fn()
            count += 50
            torch.cuda.synchronize()
        e1.record()
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        us = e0.elapsed_time(e1) * 1e3 / count
        clk, pw = smi.window(t0, t1)
        tf = flops[name] / us / 1e6 if flops[name] else 0
        print(f"{name:28s} {us:8.1f} us  {tf:7.0f} TFLOP/s   sm {clk} MHz   {pw} W", flush=True)
        time.sleep(0.5)
    smi.proc.terminate()


if __name__ == "__main__":
    main()
