"""SURVEY.md section 8(d) comparator (ii): the UNMODIFIED reference head (nets/PartialFC.py as vendored into oracle/_ref)
run eagerly on the same B200 under NCCL world size 1, fp32 and fp16-AMP (autocast inside the module, GradScaler flow of
model/FR_PartialFC.py:175-188 around it), next to this package's head on the same inputs: eager autograd call (the
drop-in's default flow) and GraphedHeadStep.  Informative -- the bench's reference arm is the CPU head -- but it is the
"real bar" a user of the reference sees when switching.

    python tools/bench_ref_gpu.py [--config 2] [--steps 50] [--warmup 10] [--out gpurun_out/ref_gpu.json]

Timing: CUDA events around `--steps` back-to-back steps after the warm-up (every step streams more state than the L2
holds), plus the median of single steps timed alone with the L2 flushed and a host synchronisation after each."""
import argparse
import json
import os
import statistics
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def timed(step, steps, warmup, flush):
    """-> (ms per step of `steps` back-to-back steps, ms of the median synchronised single step, last loss).
    Back to back = as a training loop issues them (no host synchronisation between steps: host work of step i+1 -- Python,
    launches, the CPU sampling draw of nets/PartialFC.py:110 -- overlaps the device work of step i); single = each step
    alone, L2 flushed before it, host-synchronised after it."""
    for i in range(warmup):
        step(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(warmup, warmup + steps):
        loss = step(i)
    e1.record()
    torch.cuda.synchronize()
    stream_ms = e0.elapsed_time(e1) / steps
    ts = []
    for i in range(min(steps, 20)):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        loss = step(i)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return stream_ms, statistics.median(ts), float(loss)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=2, choices=[1, 2, 3, 4])
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "ref_gpu.json"))
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    dist.init_process_group("nccl", init_method="tcp://127.0.0.1:29547", rank=0, world_size=1, device_id=dev)
    import bench
    import face_recognition_pytorch_b200 as pfc
    ref_dir = os.path.join(ROOT, "oracle", "_ref")
    if not os.path.exists(os.path.join(ref_dir, "nets", "PartialFC.py")):
        sys.exit("oracle/_ref is absent: run `python oracle/make_ref.py` where /root/reference exists")
    sys.dont_write_bytecode = True
    sys.path.insert(0, ref_dir)
    from nets.PartialFC import PartialFC as RefPartialFC       # the reference's module, unmodified

    cfg = bench.CONFIGS[args.config]
    C, B, r = cfg["C"], cfg["B"], cfg["r"]
    w_shard, xs, ls = bench.synth(cfg, 0, 1, 4, dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    rows = []

    def sgd(head):
        return torch.optim.SGD([{"params": head.parameters()}], lr=bench.LR, momentum=bench.MOMENTUM,
                               weight_decay=bench.WD)

    # ---- the reference head, eager (fp32; fp16 autocast + GradScaler)
    for amp in (False, True):
        conf = types.SimpleNamespace(emd_size=bench.EMB, sample_rate=r, mixed_precision=amp, loss_s=bench.S, loss_m=bench.M)
        head = RefPartialFC(conf=conf, num_classes=C)
        head.load_state_dict({"weight": w_shard.clone()})
        head = head.train().to(dev)
        opt = sgd(head)
        scaler = torch.amp.GradScaler("cuda", enabled=amp)

        def step(i, head=head, opt=opt, scaler=scaler, amp=amp):
            x = xs[i % 4].detach().clone().requires_grad_(True)
            opt.zero_grad()
            loss = head(x, ls[i % 4].clone(), opt)             # model/FR_PartialFC.py:175
            if amp:                                            # :178-184
                scaler.scale(loss).backward()
                scaler.unscale_(opt)
                scaler.step(opt)
                scaler.update()
            else:
                loss.backward()                                # :184
                opt.step()                                     # :188
            return loss.detach()
        ms, single, loss = timed(step, args.steps, args.warmup, flush)
        rows.append({"impl": "reference head, torch eager" + (" fp16 autocast + GradScaler" if amp else " fp32"),
                     "ms_per_step": ms, "ms_single_step": single, "samples_per_s": B / (ms * 1e-3), "last_loss": loss,
                     "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30})
        print(rows[-1], flush=True)
        del head, opt, step
        torch.cuda.empty_cache()
        torch.cuda.reset_peak_memory_stats()

    # ---- this package's head on the same inputs
    variants = [("eager autograd, torch.optim.SGD (drop-in default)", False, False, False, False),
                ("eager autograd, fused update", False, True, False, False),
                ("GraphedHeadStep" + (", CPU-generator draw like the reference" if r < 1 else " (bench.py's step)"),
                 False, True, True, False),
                ("GraphedHeadStep, fp16 operands (conf.mixed_precision)", True, True, True, False)]
    if r < 1:   # the draw of nets/PartialFC.py:110 on the device instead (opt-in, not the reference's index set)
        variants.append(("GraphedHeadStep, conf.device_sampling", False, True, True, True))
    for name, amp, fused, graphed, dev_draw in variants:
        conf = types.SimpleNamespace(emd_size=bench.EMB, sample_rate=r, mixed_precision=amp, loss_s=bench.S, loss_m=bench.M,
                                     fused_optimizer=fused, device_sampling=dev_draw)
        head = pfc.PartialFC(conf, C)
        head.load_state_dict({"weight": w_shard.clone()})
        head = head.train().cuda()
        opt = sgd(head)
        if graphed:
            for i in range(3):
                head.fused_step(xs[i % 4].detach(), ls[i % 4], opt)
            g = pfc.GraphedHeadStep(head, opt, B, bench.EMB)

            def step(i, g=g):
                return g(xs[i % 4].detach(), ls[i % 4])[0]
        else:
            def step(i, head=head, opt=opt, fused=fused):
                x = xs[i % 4].detach().clone().requires_grad_(True)
                loss = head(x, ls[i % 4].clone(), opt)
                loss.backward()
                if not fused:
                    opt.step()
                    opt.zero_grad(set_to_none=True)
                return loss.detach()
        ms, single, loss = timed(step, args.steps, args.warmup, flush)
        rows.append({"impl": "this head, " + name, "ms_per_step": ms, "ms_single_step": single,
                     "samples_per_s": B / (ms * 1e-3), "last_loss": loss, "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30})
        print(rows[-1], flush=True)
        del head, opt, step
        torch.cuda.empty_cache()
        torch.cuda.reset_peak_memory_stats()

    out = {"workload": cfg["name"], "gpu": torch.cuda.get_device_name(0), "steps": args.steps, "warmup": args.warmup,
           "timing": "ms_per_step: CUDA events around the back-to-back steps / steps; ms_single_step: median of steps timed alone "
                     "(L2 flushed before, host-synchronised after)", "rows": rows}
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(out, f, indent=1)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
