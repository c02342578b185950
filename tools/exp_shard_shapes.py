"""EXPERIMENT (one GPU, no exchange): the head step at the per-GPU shard shapes of BASELINE configs[1] on 1 / 2 / 4 / 8 GPUs
(global batch 1024 against 93 431 / 46 716 / 23 358 / 11 679 classes), graph-replayed, for each step variant -- what the
strong-scaling curve would be if the exchanges were free, and where the fixed per-launch costs are.
    python tools/exp_shard_shapes.py  [--kernels]"""
import argparse
import os
import statistics
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shards", default="93431,46716,23358,11679")
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--reps", type=int, default=30)
    ap.add_argument("--kernels", action="store_true", help="also print per-kernel times of an instrumented eager pass")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    dist.init_process_group("nccl", init_method="tcp://127.0.0.1:29546", rank=0, world_size=1, device_id=dev)
    import bench
    import face_recognition_pytorch_b200 as pfc
    from face_recognition_pytorch_b200 import kernels as K
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for C in (int(v) for v in args.shards.split(",")):
        cfg = dict(bench.CONFIGS[2], C=C, B=args.batch)
        w_shard, xs, ls = bench.synth(cfg, 0, 1, 4, dev)
        for mode in ("serial", "gemm_fork"):     # conf.dx_fork_gemm: dX GEMM on the side stream next to the dW GEMM
            conf = types.SimpleNamespace(emd_size=512, sample_rate=1.0, mixed_precision=False, loss_s=64.0, loss_m=0.5,
                                         fused_optimizer=True, dx_fork_gemm=mode == "gemm_fork")
            head = pfc.PartialFC(conf, C)
            head.load_state_dict({"weight": w_shard.clone()})
            head = head.train().cuda()
            dummy = torch.nn.Parameter(torch.zeros(1, device=dev))
            opt = torch.optim.SGD([{"params": [dummy]}, {"params": head.parameters()}], lr=0.1, momentum=0.9,
                                  weight_decay=5e-4)
            for i in range(3):
                head.fused_step(xs[i % 4], ls[i % 4], opt)
            torch.cuda.synchronize()
            g = pfc.GraphedHeadStep(head, opt, args.batch, 512)
            for i in range(5):
                g(xs[i % 4], ls[i % 4])
            ts = []
            for i in range(args.reps):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                g(xs[i % 4], ls[i % 4])
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1) * 1e3)
            line = f"n={C:6d} mode={mode:9s} graph step {statistics.median(ts):7.1f} us (min {min(ts):7.1f})"
            if args.kernels:
                K.enable_timing(True)
                for i in range(6):
                    flush.zero_()
                    head.fused_step(xs[i % 4], ls[i % 4], opt)
                kt = K.collect_timing()
                K.enable_timing(False)
                line += "  eager: " + " ".join(f"{k.replace('pfc_', '')}={v['ms_avg'] * 1e3:.1f}" for k, v in kt.items())
            print(line, flush=True)
            del g, head, opt
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
