"""Headline benchmark: PartialFC ArcFace fwd + bwd (+ fused SGD update) samples/s at 93,431 classes, d = 512,
global batch 1024 (BASELINE.json configs[1]) on N B200s of one node, with the roofline of the dominant kernel and the
reference head's CPU implementation timed on the host cores beside it.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config 1|2|3|4] [--scaling strong|weak]
                    [--mode fused|unfused] [--amp] [--no-graph] [--no-parity] [--no-cpu-baseline]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \\
        bench.py --gpus N --steps K --warmup W

One JSON line on stdout (rank 0).  A "step" is one pass of the head's hot path over one synthetic global batch:
normalise -> (all-gather) -> [sample + gather rows] -> cosine GEMM + margin/softmax epilogue (+ dX contraction) ->
(all-reduce) -> loss -> backward GEMMs -> (reduce-scatter) -> normalise-backward -> SGD/momentum update of the class
shard [-> scatter rows back].  --config picks the BASELINE.json configuration (default 2, the one `metric` is quoted
on; 3 and 4 are the sampled Glint360K / WebFace42M shapes, meant for --gpus 8); --scaling weak keeps 128 samples per
GPU instead of the global batch.

Before the timed region every run checks the step it is about to time against the CPU oracle on the same inputs
(rank 0 computes the whole world's step; loss <= 1e-3 relative, dX and weight-update cosine >= 0.999 per rank, sampled
index sets identical) and reports it as "parity_check"; a failed check fails the run.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
import types

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

EMB = 512
S, M = 64.0, 0.5
LR, MOMENTUM, WD = 0.1, 0.9, 5e-4
METRIC = "PartialFC ArcFace fwd+bwd samples/sec @93k cls,d=512"
UNIT = "samples/s"
CONFIGS = {
    # noise: spread of the synthetic embeddings around their class centre (x = centre/|centre| + noise * N(0,1)/sqrt(d));
    # configs[0] uses init-like embeddings (cos to the target ~0.3, loss O(10)) so that "loss within 1e-3 RELATIVE" is a
    # statement about the arithmetic and not about a loss that is itself ~0.03 at batch 128
    1: dict(C=10000, B=128, r=1.0, noise=3.0, name="configs[0]: ArcFace s=64 m=0.5, 512-d, batch 128, 10k classes"),
    2: dict(C=93431, B=1024, r=1.0, name="configs[1]: PartialFC C=93431 d=512 global_batch=1024 sample_rate=1.0 s=64 m=0.5"),
    3: dict(C=360232, B=1024, r=0.1, name="configs[2]: PartialFC C=360232 d=512 global_batch=1024 sample_rate=0.1 s=64 m=0.5"),
    4: dict(C=2000000, B=4096, r=0.2, name="configs[3]: PartialFC C=2000000 d=512 global_batch=4096 sample_rate=0.2 s=64 m=0.5"),
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return dict(hbm=j["hbm_gbs"], tf_burst=j["bf16_tflops"], tf_sust=j.get("bf16_tflops_sustained", j["bf16_tflops"]),
                    kind="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, kind="fallback")


def synth(cfg, rank, world, n_data, dev):
    """Synthetic data of the named shape, generated ON the device with one seeded generator so that every rank sees
    the same world: N(0, 0.01) class centres, uniform labels, trained-like unit embeddings (cos to the target ~0.7).
    Returns this rank's shard and, per batch, this rank's local rows (fp32) and labels (int64)."""
    import torch
    from face_recognition_pytorch_b200 import shard_range
    C, B = cfg["C"], cfg["B"]
    nl, cs = shard_range(C, rank, world)
    g = torch.Generator(device=dev).manual_seed(1234)
    w_full = torch.empty(C, EMB, device=dev).normal_(0, 0.01, generator=g)
    b = B // world
    xs, ls = [], []
    for s in range(n_data):
        lab = torch.randint(0, C, (B,), generator=g, device=dev)
        x = torch.nn.functional.normalize(w_full[lab]) + cfg.get("noise", 1.0) * torch.randn(
            B, EMB, generator=g, device=dev) / EMB ** 0.5
        x = torch.nn.functional.normalize(x)
        xs.append(x[rank * b:(rank + 1) * b].contiguous())
        ls.append(lab[rank * b:(rank + 1) * b].contiguous())
    shard = w_full[cs:cs + nl].clone()
    del w_full
    torch.cuda.empty_cache()
    return shard, xs, ls


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, pw, reasons = [], [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1])); pw.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        load = [v for v in sm if v > 0.5 * max(sm)] or sm
        return {"sm_mhz": statistics.median(load), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm),
                "power_w": statistics.median(pw) if pw else None}


# ---------------------------------------------------------------------------------------------- CPU reference arm
def _reference_modules():
    """The reference's own head modules (oracle/_ref, vendored from /root/reference by oracle/make_ref.py -- git-ignored,
    travels to the GPU box), patched for a CPU-only run exactly as SURVEY.md section 8c describes; None if absent."""
    ref = os.path.join(ROOT, "oracle", "_ref")
    if not os.path.exists(os.path.join(ref, "nets", "PartialFC.py")):
        return None
    import torch
    import torch.distributed as dist
    sys.dont_write_bytecode = True
    if ref not in sys.path:
        sys.path.insert(0, ref)
    torch.Tensor.cuda = lambda self, *a, **k: self            # the module hard-codes .cuda() (nets/PartialFC.py:108-113)
    if not dist.is_initialized():
        dist.init_process_group("gloo", init_method="tcp://127.0.0.1:29543", rank=0, world_size=1)
    import warnings
    warnings.filterwarnings("ignore")
    from nets.PartialFC import PartialFC                       # noqa: E402  (the reference's module, unmodified)
    return PartialFC


def _cpu_reference_runner(cfg):
    """Returns (step_fn(i), kind): one fwd + bwd + SGD step of the reference head on the host cores, fp32, one rank
    holding the whole problem.  kind "reference": the reference's own nn.Module; "port": oracle.cpu_reference_step."""
    import torch
    C, B, r = cfg["C"], cfg["B"], cfg["r"]
    torch.set_num_threads(os.cpu_count() or 1)
    g = torch.Generator().manual_seed(1234)
    w0 = torch.normal(0, 0.01, (C, EMB), generator=g)
    data = []
    for s in range(4):
        lab = torch.randint(0, C, (B,), generator=torch.Generator().manual_seed(7 + s))
        x = torch.nn.functional.normalize(torch.nn.functional.normalize(w0[lab]) + cfg.get("noise", 1.0) * torch.randn(
            B, EMB, generator=torch.Generator().manual_seed(42 + s)) / EMB ** 0.5)
        data.append((x, lab))
    PartialFC = _reference_modules()
    if PartialFC is not None:
        conf = types.SimpleNamespace(emd_size=EMB, sample_rate=r, mixed_precision=False, loss_s=S, loss_m=M)
        head = PartialFC(conf=conf, num_classes=C)
        head.load_state_dict({"weight": w0.clone()})
        head.train()
        opt = torch.optim.SGD([{"params": head.parameters()}], lr=LR, momentum=MOMENTUM, weight_decay=WD)

        def step(i):
            x, lab = data[i % len(data)]
            x = x.clone().requires_grad_(True)
            opt.zero_grad()
            loss = head(x, lab.clone(), opt)         # model/FR_PartialFC.py:175
            loss.backward()                          # :184
            opt.step()                               # :188
            return float(loss.detach())
        return step, "reference"
    if r < 1:
        raise RuntimeError("the CPU port only covers sample_rate == 1; run oracle/make_ref.py where /root/reference exists")
    from oracle import head_oracle as ho
    w = torch.nn.Parameter(w0)
    opt = torch.optim.SGD([w], lr=LR, momentum=MOMENTUM, weight_decay=WD)
    margin = ho.Margin("arcface", S, M)

    def step(i):
        x, lab = data[i % len(data)]
        return ho.cpu_reference_step(x, lab, w, opt, margin)
    return step, "port"


def run_reference(args, rank, world, cfg):
    """`--impl reference`: the reference head's CPU implementation of the path, all host threads, rank 0 only, on this
    arm's config / metric / unit; each timed step is one full global batch."""
    if rank != 0:
        return
    step, kind = _cpu_reference_runner(cfg)
    cores = os.cpu_count() or 1
    # bounded: a step costs ~0.5 s of all host cores at configs[1], so the run stays within a few minutes even for the
    # driver's K; the cap is stated in the line
    steps, warm = max(1, min(args.steps, 40)), max(1, min(args.warmup, 3))
    for i in range(warm):
        step(i)
    t0 = time.perf_counter()
    for i in range(warm, warm + steps):
        step(i)
    dt = (time.perf_counter() - t0) / steps
    v = cfg["B"] / dt
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warm, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": cfg["name"] + ", fwd+bwd+SGD step",
                       "note": "reference head (its own nn.Module when oracle/_ref is present) on the host CPU, one rank holding every class, fp32, each step = one full "
                               f"global batch; steps capped at 40 / warm-up at 3 (asked: {args.steps} / {args.warmup})"},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": kind,
                             "sample": f"{steps} full steps (B={cfg['B']}, C={cfg['C']}, d={EMB}) after {warm} warm-up"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


def cpu_baseline_sample(cfg):
    step, kind = _cpu_reference_runner(cfg)
    step(0)
    n, t0 = 0, time.perf_counter()
    while n < 3 or (time.perf_counter() - t0 < 10 and n < 12):
        step(n + 1)
        n += 1
    dt = (time.perf_counter() - t0) / n
    return {"value": cfg["B"] / dt, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": kind,
            "sample": f"{n} full steps of the same workload (B={cfg['B']}, C={cfg['C']}, d={EMB}, fp32) after 1 warm-up, "
                      f"{dt * 1e3:.0f} ms/step"}


_REAL_STDOUT = None


def _reserve_stdout():
    """stdout carries exactly ONE JSON line: libraries that print to fd 1 (NCCL's version banner) go to stderr."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


# ---------------------------------------------------------------------------------------------- parity check
def parity_check(cfg, head, opt, x_local, lab_local, perm_local, w_shard, rank, world, dev, fused):
    """One eager step from the initial weights on every rank, compared on rank 0 with the CPU oracle's step for the whole
    world (oracle/head_oracle.py, fp32): loss, every rank's dX, every rank's weight update (fused) or dW (un-fused), and the
    sampled index sets.  Restores the head afterwards.  Returns the dict for the JSON line (rank 0) or None."""
    import torch
    import torch.distributed as dist
    sampled = cfg["r"] < 1
    x = x_local.clone().requires_grad_(True)
    loss = head(x, lab_local.clone(), opt, perm=perm_local)
    loss.backward()
    if sampled:
        idx = head.weight_index.clone()
        if fused:      # in-place update through the index list (conf.inplace_update), or the gathered rows
            upd = (head.weight[idx] if head._indexed else head.weight_activated.data) - w_shard[idx]
        else:
            upd = head.weight_activated.grad.clone()
    else:
        idx = None
        upd = (head.state_dict()["weight"] - w_shard) if fused else head.weight_activated.grad.clone()
    dx = x.grad.clone()
    torch.cuda.synchronize()
    # ship everything to rank 0
    def gather(t):
        if world == 1:
            return [t.cpu()]
        sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
        dist.all_gather(sizes, torch.tensor([t.shape[0]], dtype=torch.int64, device=dev))
        mx = int(max(int(s) for s in sizes))
        pad = torch.zeros((mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=dev)
        pad[:t.shape[0]] = t
        outs = [torch.zeros_like(pad) for _ in range(world)] if rank == 0 else None
        dist.gather(pad, outs, dst=0)
        return [o[:int(s)].cpu() for o, s in zip(outs, sizes)] if rank == 0 else None
    g_x, g_l, g_w, g_dx, g_upd = gather(x_local), gather(lab_local), gather(w_shard), gather(dx), gather(upd)
    g_idx = gather(idx) if sampled else None
    g_perm = gather(perm_local) if sampled else None
    out = None
    if rank == 0:
        from oracle import head_oracle as ho
        torch.set_num_threads(os.cpu_count() or 1)
        t0 = time.perf_counter()
        res = ho.head_step(g_x, g_l, g_w, cfg["C"], ho.Margin("arcface", S, M), sample_rate=cfg["r"], perms=g_perm,
                           dtype=torch.float32)

        def cos(a, b):
            a, b = a.double().flatten(), b.double().flatten()
            return float(a @ b / (a.norm() * b.norm()).clamp_min(1e-300))
        rel = abs(float(loss) - float(res.loss)) / abs(float(res.loss))
        cdx = min(cos(a, b) for a, b in zip(g_dx, res.dx_local))
        cup = []
        for r in range(world):
            if fused:      # first SGD step from zero momentum: w1 - w0 = -lr (dw + wd w0)
                w_act = g_w[r][res.index[r]] if sampled else g_w[r]
                ref = -LR * (res.dw[r].float() + WD * w_act)
            else:
                ref = res.dw[r].float()
            cup.append(cos(g_upd[r], ref))
        same_idx = all(torch.equal(a, b) for a, b in zip(g_idx, res.index)) if sampled else None
        ok = rel <= 1e-3 and cdx >= 0.999 and min(cup) >= 0.999 and same_idx is not False
        out = {"ok": bool(ok), "loss": float(loss), "oracle_loss": float(res.loss), "loss_rel_err": rel, "dx_cos_min": cdx,
               ("update_cos_min" if fused else "dw_cos_min"): min(cup), "sampled_index_sets_equal": same_idx,
               "ranks": world, "oracle": "oracle/head_oracle.py::head_step fp32 on rank 0's host cores, whole world, "
               f"same inputs ({time.perf_counter() - t0:.1f} s)"}
    # restore: weights, optimizer state, bookkeeping of the head
    head.load_state_dict({"weight": w_shard.clone()})
    for nm in ("weight_mom", "weight_activated_mom"):
        t = getattr(head, nm, None)
        if isinstance(t, torch.Tensor) and t.numel():
            t.zero_()
    if getattr(head, "_fused_state", None) is not None:
        head._fused_state.zero_()
    opt.state.clear()
    head.init_weight_update = True
    opt.zero_grad(set_to_none=True)
    flag = torch.tensor([1 if (out is None or out["ok"]) else 0], device=dev)
    if world > 1:
        dist.broadcast(flag, 0)
    if int(flag.item()) == 0:
        if rank == 0:
            print("# parity_check FAILED: " + json.dumps(out), file=sys.stderr)
        raise SystemExit(3)
    return out


def main():
    _reserve_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=[1, 2, 3, 4], help="BASELINE.json configs[config - 1]")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="strong: the config's global batch on every N; weak: 128 samples per GPU (global batch 128 N)")
    ap.add_argument("--mode", default="fused", choices=["fused", "unfused"],
                    help="fused: SGD update fused into the backward (conf.fused_optimizer); unfused: dW handed to "
                         "torch.optim.SGD")
    ap.add_argument("--no-graph", action="store_true", help="launch eagerly instead of replaying a CUDA graph")
    ap.add_argument("--autograd", action="store_true", help="capture forward + loss.backward() instead of head.fused_step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle check before the timed region")
    ap.add_argument("--gemm-mode", type=int, default=0, help="0 auto, 1 single CTA, 2 multicast pair, 3 cta_group::2")
    ap.add_argument("--no-flush-l2", dest="flush_l2", action="store_false",
                    help="skip the 256 MB write between timed steps (one step streams > 1 GB through the 126 MB L2 anyway)")
    ap.add_argument("--dw-first", default="auto", choices=["auto", "0", "1"],
                    help="order of the gradient GEMMs: 1 = dW, dX, update; 0 = dX, dW, update; auto = 1 on one GPU")
    ap.add_argument("--amp", action="store_true",
                    help="conf.mixed_precision = True: fp16 GEMM operands like the reference's autocast (default: bf16)")
    ap.add_argument("--no-peer", action="store_true",
                    help="N>1: NCCL collectives instead of the peer-memory (NVLink) exchanges fused into the kernels")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cfg = dict(CONFIGS[args.config])
    if args.scaling == "weak":
        cfg["B"] = 128 * world
        cfg["name"] += f" (weak scaling: 128 per GPU, global batch {cfg['B']})"
    if args.impl == "reference":
        run_reference(args, rank, world, cfg)
        return
    args.warmup = max(args.warmup, 3)
    C, GLOBAL_BATCH, rate = cfg["C"], cfg["B"], cfg["r"]
    sampled = rate < 1
    mode = args.mode

    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if not dist.is_initialized():
        if "MASTER_ADDR" in os.environ and "RANK" in os.environ:
            dist.init_process_group("cpu:gloo,cuda:nccl", device_id=dev)
        else:
            # gloo beside nccl: the CPU-baseline leg runs the reference's own module, which all-reduces CPU tensors
            dist.init_process_group("cpu:gloo,cuda:nccl", init_method="tcp://127.0.0.1:29541", rank=0, world_size=1,
                                    device_id=dev)
    import face_recognition_pytorch_b200 as pfc
    from face_recognition_pytorch_b200 import kernels as K

    if args.gemm_mode:
        pfc._lib.lib.pfc_debug_cluster(int(args.gemm_mode))
    n_data = 4
    w_shard, xs, ls = synth(cfg, rank, world, n_data, dev)
    b = GLOBAL_BATCH // world
    fused = mode != "unfused"
    conf = types.SimpleNamespace(emd_size=EMB, sample_rate=rate, mixed_precision=bool(args.amp), loss_s=S, loss_m=M,
                                 fused_optimizer=fused,
                                 dw_first="auto" if args.dw_first == "auto" else bool(int(args.dw_first)),
                                 peer_collectives=False if args.no_peer else "auto")
    head = pfc.PartialFC(conf, C)
    head.load_state_dict({"weight": w_shard.clone()})
    head = head.train().cuda()
    dummy = torch.nn.Parameter(torch.zeros(1, device=dev))
    opt = torch.optim.SGD([{"params": [dummy]}, {"params": head.parameters()}], lr=LR, momentum=MOMENTUM,
                          weight_decay=WD)
    x_dev = [x.requires_grad_(True) for x in xs]
    l_dev = ls
    # sampling draws: one [num_local] uniform draw per step and rank, seeded, resident on the device (the reference draws
    # on the CPU generator and uploads, nets/PartialFC.py:110; the index set is a pure function of the draw)
    perms = None
    if sampled:
        gp = torch.Generator(device=dev).manual_seed(100 + rank)
        perms = [torch.rand(head.num_local, generator=gp, device=dev) for _ in range(n_data)]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)     # > 126 MB L2

    def step_eager(i):
        k = i % n_data
        if fused and not args.autograd:       # the same no-autograd kernel sequence GraphedHeadStep captures
            return head.fused_step(x_dev[k].detach(), l_dev[k], opt, perm=None if perms is None else perms[k])[0]
        x_dev[k].grad = None
        loss = head(x_dev[k], l_dev[k], opt, perm=None if perms is None else perms[k])
        loss.backward()
        if not fused:
            opt.step()
            opt.zero_grad(set_to_none=True)
        return loss

    # ---- parity against the oracle on exactly this configuration, before anything is timed
    parity = None
    if not args.no_parity:
        parity = parity_check(cfg, head, opt, xs[0].detach(), ls[0], None if perms is None else perms[0], w_shard, rank,
                              world, dev, fused)
    del w_shard

    clk = ClockSampler(local_rank)
    clk.__enter__()                     # sampled every 200 ms across warm-up, timed region and e2e (all under load)
    # ---- warm-up (also builds workspaces, NCCL communicators)
    l0 = K.launch_count()
    step_eager(0)
    launches_per_step = K.launch_count() - l0
    for i in range(args.warmup):
        step_eager(i)
    torch.cuda.synchronize()

    # CUDA-graph replay through the package's own public wrapper (face_recognition_pytorch_b200.GraphedHeadStep)
    gstep = None
    if not args.no_graph and fused:
        try:
            gstep = pfc.GraphedHeadStep(head, opt, b, EMB, autograd=args.autograd)
        except Exception as e:   # report, fall back to eager launches (still the CUDA path)
            if rank == 0:
                print(f"# CUDA graph capture failed ({type(e).__name__}: {e}); timing eager launches", file=sys.stderr)
            gstep = None
            torch.cuda.synchronize()

    def run_step(i):
        if gstep is not None:
            if sampled:     # the sampling draw is an input of the step: resident on the device like the batch
                gstep(x_dev[i % n_data].data, l_dev[i % n_data], perm=perms[i % n_data])
            else:
                gstep(x_dev[i % n_data].data, l_dev[i % n_data])
        else:
            step_eager(i)

    for i in range(3):
        run_step(i)
    torch.cuda.synchronize()

    # ---- timed region: K steps, device time by CUDA events per step (L2 flushed, untimed, between steps)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    dist.barrier()
    torch.cuda.synchronize()
    t_wall0 = time.perf_counter()
    for i in range(args.steps):
        if args.flush_l2:
            flush.zero_()
        ev[i][0].record()
        run_step(i)
        ev[i][1].record()
    torch.cuda.synchronize()
    dist.barrier()
    t_wall = time.perf_counter() - t_wall0
    if gstep is not None:
        launches_per_step = gstep.launches_per_replay          # the captured (no-autograd) step
    launches = launches_per_step * args.steps      # kernels of libpfc_b200 per step (replayed from the graph or eager)
    ms_steps = [a.elapsed_time(bb) for a, bb in ev]
    t = torch.tensor([sum(ms_steps)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, dist.ReduceOp.MAX)
    ms_per_step = float(t.item()) / args.steps
    value = GLOBAL_BATCH / (ms_per_step * 1e-3)

    # ---- per-kernel durations: a separate instrumented EAGER pass (events around each C-ABI call, on the stream the
    # call is made on).  One GPU only: with several ranks the eager pass times rank skew inside the exchange barriers.
    kt = {}
    if world == 1:
        K.enable_timing(True)
        for i in range(6):
            flush.zero_()
            step_eager(i)
        torch.cuda.synchronize()
        kt = K.collect_timing()
        K.enable_timing(False)

    # ---- end to end through the public API with HOST buffers (pinned), H2D + D2H inside the timed region
    hx = [x.detach().cpu().pin_memory() for x in xs]
    hl = [l.cpu().pin_memory() for l in ls]
    e2e_steps = max(5, min(args.steps, 20))

    # Software-pipelined like a training loop with a prefetching loader: the H2D copy of step i+1's inputs and the D2H
    # copy of step i's dX run on a copy stream underneath the replay of the step in between; every step still moves its
    # own inputs host -> device and its own dX + loss device -> host inside the timed region.
    copy_stream = torch.cuda.Stream(device=dev)
    st_x = [torch.empty(b, EMB, device=dev) for _ in range(2)]
    st_l = [torch.empty(b, dtype=torch.int64, device=dev) for _ in range(2)]
    st_dx = [torch.empty(b, EMB, device=dev) for _ in range(2)]
    dx_hosts = [torch.empty(b, EMB, dtype=torch.float32).pin_memory() for _ in range(2)]
    st_loss = [torch.zeros(1, device=dev) for _ in range(2)]
    loss_hosts = [torch.zeros(1, dtype=torch.float32).pin_memory() for _ in range(2)]
    losses_seen = []
    ev_in = [torch.cuda.Event() for _ in range(2)]       # inputs of slot k are on the device
    ev_used = [torch.cuda.Event() for _ in range(2)]     # the step has consumed slot k's inputs / produced its dX
    ev_out = [torch.cuda.Event() for _ in range(2)]      # dX of slot k has reached the host

    def e2e_prefetch(i):
        k = i % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(ev_used[k])
            st_x[k].copy_(hx[i % n_data], non_blocking=True)
            st_l[k].copy_(hl[i % n_data], non_blocking=True)
            ev_in[k].record(copy_stream)

    def head_call(x, lab, i):
        if gstep is not None:
            # sampled: perm=None -- the draw is made on the CPU generator and uploaded every step, as in the reference
            return gstep(x, lab)
        xg = x.detach().clone().requires_grad_(True)
        loss = head(xg, lab.clone(), opt, perm=None if perms is None else perms[i % n_data])
        loss.backward()
        if not fused:
            opt.step()
            opt.zero_grad(set_to_none=True)
        return loss, xg.grad

    # the extra stream / event calls cost ~50 us of host time per step: worth it only while the copies are long enough
    pipelined = b * EMB * 4 >= (1 << 20)

    def e2e_step_simple(i):
        if gstep is not None:         # pinned host -> static device buffers -> graph replay -> pinned host
            loss, dx = gstep(hx[i % n_data], hl[i % n_data])
        else:
            loss, dx = head_call(hx[i % n_data].to(dev, non_blocking=True), hl[i % n_data].to(dev, non_blocking=True), i)
        # same lagged read as the pipelined loop, on one stream: step i's dX and loss go to pinned memory asynchronously
        # and the host reads them while step i+1 runs
        k = i % 2
        dx_hosts[k].copy_(dx, non_blocking=True)
        loss_hosts[k].copy_(loss.detach().reshape(1), non_blocking=True)
        ev_out[k].record(torch.cuda.current_stream())
        if i >= 1:
            ev_out[1 - k].synchronize()
            losses_seen.append(float(loss_hosts[1 - k][0]))

    def e2e_step(i):
        if not pipelined:
            return e2e_step_simple(i)
        k = i % 2
        cur = torch.cuda.current_stream()
        e2e_prefetch(i + 1)
        cur.wait_event(ev_in[k])
        loss, dx = head_call(st_x[k], st_l[k], i)
        cur.wait_event(ev_out[k])                        # slot k's previous dX / loss have left the staging buffers
        st_dx[k].copy_(dx, non_blocking=True)
        st_loss[k].copy_(loss.detach().reshape(1), non_blocking=True)
        ev_used[k].record(cur)
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(ev_used[k])
            dx_hosts[k].copy_(st_dx[k], non_blocking=True)
            loss_hosts[k].copy_(st_loss[k], non_blocking=True)
            ev_out[k].record(copy_stream)
        # every step's loss and dX reach the host; the host READS them one step late (while the next step runs), as a
        # training loop that logs the previous step's loss does -- no pipeline bubble for a 4-byte read
        if i >= 1:
            ev_out[1 - k].synchronize()
            losses_seen.append(float(loss_hosts[1 - k][0]))

    for k in range(2):
        ev_used[k].record(torch.cuda.current_stream())
        ev_out[k].record(torch.cuda.current_stream())
    e2e_prefetch(0)
    for i in range(3):
        e2e_step(i)
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(3, 3 + e2e_steps):
        e2e_step(i)
    torch.cuda.synchronize()
    losses_seen.append(float(loss_hosts[(3 + e2e_steps - 1) % 2][0]))    # the last step's loss
    assert all(v == v and v > 0 for v in losses_seen[-e2e_steps:]), "e2e losses must be finite"
    dist.barrier()
    te = torch.tensor([(time.perf_counter() - t0) / e2e_steps], dtype=torch.float64, device=dev)
    dist.all_reduce(te, dist.ReduceOp.MAX)
    e2e_value = GLOBAL_BATCH / float(te.item())

    # keep the GPU under the same load until the sampler has a few readings
    t_hold = time.perf_counter()
    i = 0
    while True:                                   # rank 0 decides, so every rank issues the same collectives
        for _ in range(100):
            run_step(i); i += 1
        torch.cuda.synchronize()
        more = torch.tensor([1 if (len(clk.rows) < 5 and time.perf_counter() - t_hold < 3.0) else 0], device=dev)
        dist.broadcast(more, 0)
        if int(more.item()) == 0:
            break
    clk.__exit__()
    if rank != 0:
        dist.barrier()
        return

    pk = peaks()
    nl = head.num_local
    n_act = head._n                               # active classes per rank and step (num_local, or the sampled count)
    flops_gemm = 2.0 * GLOBAL_BATCH * n_act * EMB          # one of the three contractions, THIS rank
    kern = {k: v for k, v in kt.items()}
    # algorithmic work per launch of each timed kernel (DESIGN.md "Kernels and rooflines"):
    #   GEMMs: 2*B*n*d each (the forward + dX kernel does two of them);
    #   update: SGD state w r/w + momentum r/w (16 B per element) + the next step's bf16 shard (2 B) -- the kernel's own
    #   bf16 gradient spill is traffic, not algorithmic work.
    alg = {
        "pfc_forward": ("tensor", flops_gemm), "pfc_backward_dx": ("tensor", flops_gemm),
        "pfc_backward_dw": ("tensor", flops_gemm),
        "pfc_dw_sgd": ("hbm", n_act * EMB * 18.0),
        "pfc_dw_finalize": ("hbm", n_act * EMB * 4 * 3.0),
    }
    dom = max((k for k in kern if k in alg), key=lambda k: kern[k]["ms_total"], default=None)
    roof = None
    if dom:
        bound, work = alg[dom]
        ms = kern[dom]["ms_avg"]
        if bound == "tensor":
            ach, peak, unit = work / (ms * 1e-3) / 1e12, pk["tf_burst"], "TFLOP/s"
        else:
            ach, peak, unit = work / (ms * 1e-3) / 1e9, pk["hbm"], "GB/s"
        roof = {"kernel": dom, "bound": bound, "achieved": ach, "peak": peak, "unit": unit, "frac": ach / peak,
                "traffic": None, "traffic_note": "dram__bytes per launch: profiles/r02*_ncu_full_summary.txt (ncu --set full)",
                "algorithmic_work": work, "peak_source": pk["kind"] + (" burst bf16" if bound == "tensor" else " copy"),
                "ms_per_launch": ms,
                "note": ("achieved = (16 B fp32 state + 2 B bf16 shard per element) / time of the update kernel"
                         if bound == "hbm" else "achieved = algorithmic flops of this kernel / its time (eager pass, alone)")}
    step_tf = 3 * flops_gemm / (ms_per_step * 1e-3) / 1e12         # per GPU: flops_gemm is this rank's share
    line = {
        "metric": METRIC if args.config == 2 else "PartialFC ArcFace fwd+bwd samples/sec", "value": value, "unit": UNIT,
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "fp16" if args.amp else "bf16", "data": "synthetic",
        "config": {"workload": cfg["name"] + ", fwd+bwd+SGD step",
                   "classes_per_gpu": nl, "active_classes_per_gpu": n_act, "local_batch": b,
                   "parallelism": f"class-sharded x{world}", "mode": mode,
                   "update": "in-step (fused into the backward)" if fused else "torch.optim.SGD",
                   "launch": ("cuda-graph replay (GraphedHeadStep" + (")" if args.autograd else ", no autograd)")
                              if gstep is not None else
                              "eager (head.forward + loss.backward)" if args.autograd or not fused else
                              "eager (head.fused_step, no autograd)"),
                   "dx_side_stream": True,
                   "exchange": ("none (1 GPU)" if world == 1 else
                                "peer-memory stores + flag barriers (NVLink)" if head._peer is not None else
                                "NCCL all-gather / all-reduce / reduce-scatter"),
                   "l2": ("256 MB buffer written between timed steps (untimed); per-step CUDA events summed"
                          if args.flush_l2 else
                          "inputs larger than L2: every step streams this GPU's weights + momentum (fp32), bf16 shard, "
                          f"bf16 spill and gradient = {n_act * EMB * 16 / 1e6:.0f} MB >> 126 MB L2; back-to-back steps, "
                          "per-step CUDA events summed")},
        "roofline": roof,
        "step_roofline": {"bound": "tensor", "achieved": step_tf, "peak": pk["tf_burst"], "unit": "TFLOP/s/GPU",
                          "frac": step_tf / pk["tf_burst"], "frac_of_sustained": step_tf / pk["tf_sust"],
                          "work": "6*B*n*D per step and GPU (3 GEMMs, no recompute credited); the HBM-bound update is inside the step",
                          "peak_source": pk["kind"] + " burst bf16"},
        "kernels_ms": {k: round(v["ms_avg"], 4) for k, v in kern.items()} if kern else None,
        "kernels_ms_note": ("instrumented eager pass, CUDA events around each call on the stream it is made on (calls on "
                            "the side streams overlap the main stream's)" if kern else
                            "per-kernel times are reported on one GPU only"),
        "parity_check": parity,
        "e2e": {"value": e2e_value, "unit": UNIT,
                "h2d_bytes_per_step": b * EMB * 4 + b * 8 + (nl * 4 if sampled and gstep is not None else 0),
                "d2h_bytes_per_step": 4 + b * EMB * 4, "steps": e2e_steps, "timing": "host wall clock, max over ranks" +
                          ("; H2D of step i+1 and D2H of step i's dX + loss on a copy stream, host reads them one step late"
                           if pipelined else "; D2H of step i's dX + loss asynchronous, host reads them one step late"),
                "api": ("GraphedHeadStep(head, opt)(x, labels) fed from pinned host buffers; dX and loss copied to pinned host "
                        "memory every step" if gstep is not None else
                        "head(x, labels, opt); loss.backward(); dX and loss copied to pinned host memory every step")},
        "gpu_launches": launches,
        "clocks": clk.summary(),
        "wall_ms_per_step": t_wall / args.steps * 1e3,
    }
    if world == 1 and not args.no_cpu_baseline:
        try:
            line["cpu_baseline"] = cpu_baseline_sample(cfg)
        except Exception as e:
            line["cpu_baseline"] = {"unavailable": f"{type(e).__name__}: {e}"}
    else:
        line["cpu_baseline"] = None
    emit(line)
    dist.barrier()


if __name__ == "__main__":
    main()
