"""EXPERIMENT (one GPU): do CTAs of a small register-only kernel get resident next to the persistent GEMM CTAs?
Launches a GEMM of the head (dW / dX / forward) on one stream and tools/probe's spinning probe kernel on another, and
reports when and where the probe CTAs ran relative to the GEMM's own start / end (%globaltimer stamps).
    python tools/exp_coresidency.py"""
import ctypes
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    dist.init_process_group("nccl", init_method="tcp://127.0.0.1:29548", rank=0, world_size=1, device_id=dev)
    import bench
    import face_recognition_pytorch_b200 as pfc
    from face_recognition_pytorch_b200 import kernels as K
    probe = ctypes.CDLL(os.path.join(ROOT, "tools", "probe", "libprobe.so"))
    cfg = dict(bench.CONFIGS[2])
    w_shard, xs, ls = bench.synth(cfg, 0, 1, 2, dev)
    conf = types.SimpleNamespace(emd_size=512, sample_rate=1.0, mixed_precision=False, loss_s=64.0, loss_m=0.5,
                                 fused_optimizer=True)
    head = pfc.PartialFC(conf, cfg["C"])
    head.load_state_dict({"weight": w_shard})
    head = head.train().cuda()
    dummy = torch.nn.Parameter(torch.zeros(1, device=dev))
    opt = torch.optim.SGD([{"params": [dummy]}, {"params": head.parameters()}], lr=0.1, momentum=0.9, weight_decay=5e-4)
    for i in range(2):
        x = xs[i].requires_grad_(True)
        head(x, ls[i], opt).backward()
    torch.cuda.synchronize()
    ws = head._ws
    B, n, d = ws.B, head._n, 512
    n_pad = head._n_pad
    kind, s, m2, m3, thr = head.margin_softmax.margin_spec()
    splits = K.dx_splits(B, n, d)
    dwn = ws.grad_buffer(True, d)
    gemms = {
        "dW (96 regs)": lambda: K.backward_dw(ws.E, n_pad, ws.xs, B, n, d, dwn, keep_in_l2=False),
        "dX (96 regs)": lambda: K.backward_dx(ws.E, n_pad, ws.wn, B, n, d, ws.dx_partial, splits),
        "forward (168 regs)": lambda: K.forward(ws.xn_all, ws.wn, ws.labels_act, B, n, d, s, kind, m2, m3, thr, ws.E, n_pad,
                                                ws.part_sum, ws.tgt_raw, ws.tgt_e, ws.tgt_z),
    }
    main_s = torch.cuda.Stream(device=dev, priority=-1)
    side_s = torch.cuda.Stream(device=dev, priority=0)
    stamps = torch.zeros(2, dtype=torch.int64, device=dev)
    sink = torch.zeros(1, device=dev)
    ctas = 148 * 4
    rec = torch.zeros(ctas * 3, dtype=torch.int64, device=dev)
    P = lambda t, off=0: ctypes.c_void_p(t.data_ptr() + off)   # noqa: E731
    for gname, gemm in gemms.items():
        for regs, rname in ((0, "18 regs"), (1, "47 regs"), (3, "63 regs"), (2, "79 regs")):
            for r in range(2):
                rec.zero_()
                torch.cuda.synchronize()
                with torch.cuda.stream(main_s):
                    ev = torch.cuda.Event()
                    ev.record()
                    side_s.wait_event(ev)
                    probe.probe_stamp(P(stamps), ctypes.c_void_p(main_s.cuda_stream))
                    gemm()
                    probe.probe_stamp(P(stamps, 8), ctypes.c_void_p(main_s.cuda_stream))
                    with torch.cuda.stream(side_s):
                        probe.probe_launch(P(rec), ctas, 10000, regs, P(sink), ctypes.c_void_p(side_s.cuda_stream))
                    main_s.wait_stream(side_s)
                torch.cuda.synchronize()
            t0, t1 = (int(v) for v in stamps.cpu())
            rr = rec.cpu().view(ctas, 3)
            st = (rr[:, 1] - t0).double() / 1e3
            during = int(((rr[:, 1] < t1 - 5000) & (rr[:, 1] > t0)).sum())
            sms = len(set(int(v) for v in rr[(rr[:, 1] < t1 - 5000), 0]))
            print(f"{gname:20s} ({(t1 - t0) / 1e3:6.1f} us) + probe {rname}: {during:4d} of {ctas} probe CTAs started while the GEMM "
                  f"ran, on {sms:3d} SMs; probe start min/median/max = {st.min():.1f} / {st.median():.1f} / {st.max():.1f} us "
                  f"after the GEMM's start stamp", flush=True)


if __name__ == "__main__":
    main()
