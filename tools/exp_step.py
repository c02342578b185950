"""EXPERIMENT: launch-structure variants of the step -- programmatic dependent launch (csrc/pfc_launch.cuh), the dX-tail
fork (conf.dx_side_stream), a high-priority side stream for it, and the graph without autograd (head.fused_step).

One process (or one per GPU under torchrun), the bench workload (BASELINE configs[1]); for every (mode, mask) pair the
CUDA graph of the step is re-captured and 40 replays are timed with CUDA events (L2 flushed between replays, max over
ranks).  Mask bits are PdlId of pfc_launch.cuh: 0 normalise, 1 forward GEMM, 2 row stats / loss, 3 backward_prepare,
4 dW GEMM, 5 dX GEMM, 6 dX finalize / scatter, 7 update rows, 8 label localisation / barrier.

    python tools/exp_step.py [--configs mode:mask[:fork[:direct[:prio[:early[:l2]]]]],...] [--steps 40]
        fork   -1 auto / 0 / 1   conf.dx_side_stream
        direct 0 / 1             GraphedHeadStep(autograd=False): head.fused_step instead of forward + loss.backward()
        prio   0 / 1             high-priority side stream for the dX tail
        early  0 / 1             conf.early_dx: dX GEMM on the unpatched spill, next to statistics / loss / prepare
        l2     0 / 1             pfc_debug_l2_grad: bf16 gradient kept in L2 between the dW GEMM and the update
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/exp_step.py
"""
import argparse
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch                         # noqa: E402
import torch.distributed as dist     # noqa: E402

import bench                         # noqa: E402

DEFAULT = "0:0,1:0x1ff,2:0x1ff,1:0x001,1:0x002,1:0x004,1:0x008,1:0x010,1:0x020,1:0x040,1:0x080,0:0"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default=DEFAULT)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--repeat", type=int, default=2)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if "MASTER_ADDR" in os.environ and "RANK" in os.environ:
        dist.init_process_group("nccl", device_id=dev)
    else:
        dist.init_process_group("nccl", init_method="tcp://127.0.0.1:29547", rank=0, world_size=1, device_id=dev)
    import face_recognition_pytorch_b200 as pfc
    from face_recognition_pytorch_b200 import kernels as K

    n_data = 4
    w_shard, xs, ls = bench.synth(rank, world, n_data, dev)
    b = bench.GLOBAL_BATCH // world
    conf = types.SimpleNamespace(emd_size=bench.EMB, sample_rate=1.0, mixed_precision=False, loss_s=bench.S,
                                 loss_m=bench.M, fused_optimizer=True)
    head = pfc.PartialFC(conf, bench.C_CLASSES)
    head.load_state_dict({"weight": w_shard})
    head = head.train().cuda()
    opt = torch.optim.SGD(head.parameters(), lr=bench.LR, momentum=bench.MOMENTUM, weight_decay=bench.WD)
    x_dev = [x.to(dev) for x in xs]
    l_dev = [l.to(dev) for l in ls]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for i in range(4):
        x = x_dev[i % n_data].clone().requires_grad_(True)
        head(x, l_dev[i % n_data], opt).backward()
    torch.cuda.synchronize()
    gsteps = {}

    results = []
    for cfg in args.configs.split(","):
        f = cfg.split(":")
        mode, mask, fork = int(f[0]), int(f[1], 0), (int(f[2]) if len(f) > 2 else -1)
        direct = int(f[3]) if len(f) > 3 else 0
        prio = int(f[4]) if len(f) > 4 else 0
        early = int(f[5]) if len(f) > 5 else 0
        head.early_dx = bool(early)
        l2 = int(f[6]) if len(f) > 6 else 0
        pfc._lib.lib.pfc_debug_l2_grad(l2)
        head.dx_side_stream = "auto" if fork < 0 else bool(fork)
        if bool(prio) != head.dx_side_priority:
            head.dx_side_priority = bool(prio)
            head._side_stream = None                      # re-created with the new priority at the next step
        K.set_pdl(mode)
        pfc._lib.lib.pfc_debug_pdl_mask(mask)
        if direct not in gsteps:
            gsteps[direct] = pfc.GraphedHeadStep(head, opt, b, bench.EMB, autograd=not direct)
        gstep = gsteps[direct]
        gstep.recapture()
        best = []
        for _ in range(args.repeat):
            for i in range(5):
                gstep(x_dev[i % n_data], l_dev[i % n_data])
            ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
            dist.barrier()
            torch.cuda.synchronize()
            for i in range(args.steps):
                flush.zero_()
                ev[i][0].record()
                gstep(x_dev[i % n_data], l_dev[i % n_data])
                ev[i][1].record()
            torch.cuda.synchronize()
            ms = sum(a.elapsed_time(c) for a, c in ev) / args.steps
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, dist.ReduceOp.MAX)
            best.append(float(t))
        results.append((mode, mask, best))
        if rank == 0:
            print(f"mode {mode} mask {mask:#05x} fork {fork} direct {direct} prio {prio} early {early} l2 {l2}: " + " ".join(f"{v:.4f}" for v in best) + " ms/step", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
