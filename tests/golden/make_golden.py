"""Generates the golden fixtures in this directory by running the UNMODIFIED reference modules.

Run in the build container only (needs /root/reference; the GPU box does not have it):
    python tests/golden/make_golden.py
The reference head hard-codes .cuda() (nets/PartialFC.py:108-113,176,180); on this CPU-only box Tensor.cuda is
patched to the identity and torch.distributed runs over gloo, which leaves every arithmetic op untouched.
torch.rand is wrapped only to RECORD the draws sample() makes (nets/PartialFC.py:110) so that the same draw
can be replayed into the oracle and the CUDA sampler.
"""
import math
import os
import sys
import types
import warnings

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

sys.dont_write_bytecode = True
REF = os.environ.get("PFC_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
warnings.filterwarnings("ignore")
sys.path.insert(0, HERE)
from inputs import synth_inputs, shard, eval_inputs_cfg5, proj_matrix   # noqa: E402


def _import_reference(adamw=False):
    sys.path.insert(0, REF)
    torch.Tensor.cuda = lambda self, *a, **k: self
    from nets.PartialFC import PartialFC, PartialFCAdamW      # noqa
    from nets.ArcFace import ArcFace, CosFace, CombinedMarginLoss   # noqa
    return (PartialFCAdamW if adamw else PartialFC), ArcFace, CosFace, CombinedMarginLoss


def run_head_rank(rank, W, port, cfg, out_q):
    """One reference rank: builds PartialFC, loads its shard, runs `steps` of forward/backward/SGD."""
    adamw = cfg.get("optimizer", "sgd") in ("adamw", "adam")        # both drive PartialFCAdamW (:320)
    PartialFC, ArcFace, CosFace, CombinedMarginLoss = _import_reference(adamw)
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=W)
    C, d, b, steps, r = cfg["C"], cfg["d"], cfg["b"], cfg["steps"], cfg["sample_rate"]
    B = b * W
    w_full, xs, ls = synth_inputs(C, d, B, steps, sigma=cfg.get("sigma", 1.0))
    conf = types.SimpleNamespace(emd_size=d, sample_rate=r, mixed_precision=False, loss_s=cfg["s"], loss_m=cfg["m"])
    if cfg.get("margin", "arcface") == "combined_filter":
        # CombinedMarginLoss in its ArcFace form with inter-class filtering (nets/ArcFace.py:30-52); the head calls
        # margin_loss(conf.loss_s, conf.loss_m) (nets/PartialFC.py:88), so the extra arguments are bound here
        thr = cfg["filter_thr"]
        margin_cls = lambda s, m: CombinedMarginLoss(s, 1.0, m, 0.0, interclass_filtering_threshold=thr)   # noqa: E731
    else:
        margin_cls = {"arcface": ArcFace, "cosface": CosFace}[cfg.get("margin", "arcface")]
    head = PartialFC(conf, C, margin_loss=margin_cls)
    nl, cs = shard(C, rank, W)
    head.load_state_dict({"weight": w_full[cs:cs + nl].clone()})
    dummy = torch.nn.Parameter(torch.zeros(1))                # stands for the encoder param group
    if adamw:
        opt_cls = torch.optim.AdamW if cfg["optimizer"] == "adamw" else torch.optim.Adam   # Adam: coupled weight decay
        opt = opt_cls([{"params": [dummy]}, {"params": head.parameters()}], lr=cfg["lr"], weight_decay=cfg["wd"])
    else:
        opt = torch.optim.SGD([{"params": [dummy]}, {"params": head.parameters()}], lr=cfg["lr"],
                              momentum=cfg["momentum"], weight_decay=cfg["wd"])
    draws = []
    real_rand = torch.rand

    def recording_rand(*a, **k):
        t = real_rand(*a, **k)
        draws.append(t.clone())
        return t

    torch.manual_seed(100 + rank)
    rec = {}
    for s in range(steps):
        x = xs[s][rank * b:(rank + 1) * b].clone().requires_grad_(True)
        lab = ls[s][rank * b:(rank + 1) * b].clone()
        opt.zero_grad()
        n_before = len(draws)
        torch.rand = recording_rand
        try:
            loss = head(x, lab, opt)
        finally:
            torch.rand = real_rand
        loss.backward()
        rec[f"loss_{s}"] = loss.detach().numpy().copy()
        rec[f"dx_{s}"] = x.grad.numpy().copy()
        rec[f"dw_{s}"] = head.weight_activated.grad.numpy().copy()
        if r < 1:
            rec[f"index_{s}"] = head.weight_index.numpy().copy()
            # n_pos > num_sample: no draw is made (nets/PartialFC.py:109,114-115)
            rec[f"perm_{s}"] = draws[-1].numpy().copy() if len(draws) > n_before else np.zeros(0, np.float32)
        if adamw and r < 1:
            # The ONLY adaptation: PartialFCAdamW.sample stores a Python int in optimizer.state[...]["step"]
            # (nets/PartialFC.py:327); torch >= 2.0 rejects that ("state_steps must contain singleton tensors"), so the
            # unmodified reference cannot take an AdamW step on this torch.  Same number, as a tensor.
            st = opt.state[head.weight_activated]
            if not torch.is_tensor(st["step"]):
                st["step"] = torch.tensor(float(st["step"]))
        opt.step()
    if adamw and r < 1:
        head.update()
        rec["weight_final"] = head.weight.numpy().copy()
        rec["exp_avg_final"] = head.weight_exp_avg.numpy().copy()
        rec["exp_avg_sq_final"] = head.weight_exp_avg_sq.numpy().copy()
    elif adamw:
        rec["weight_final"] = head.weight_activated.detach().numpy().copy()
        st = opt.state[head.weight_activated]
        rec["exp_avg_final"] = st["exp_avg"].numpy().copy()
        rec["exp_avg_sq_final"] = st["exp_avg_sq"].numpy().copy()
    elif r < 1:
        head.update()                                          # flush the last step's rows (reference quirk)
        rec["weight_final"] = head.weight.numpy().copy()
        rec["mom_final"] = head.weight_mom.numpy().copy()
    else:
        rec["weight_final"] = head.weight_activated.detach().numpy().copy()
        st = opt.state[head.weight_activated]
        rec["mom_final"] = st["momentum_buffer"].numpy().copy()
    if cfg["d"] < 512:
        rec["state_dict_weight"] = head.state_dict()["weight"].numpy().copy()
    out_q.put((rank, rec))
    dist.barrier()
    dist.destroy_process_group()


def make_head_case(name, W, cfg, port):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=run_head_rank, args=(r, W, port, cfg, q)) for r in range(W)]
    for p in procs:
        p.start()
    results = dict(q.get() for _ in range(W))
    for p in procs:
        p.join()
    out = {"cfg_" + k: np.array(v) for k, v in cfg.items()}
    out["cfg_W"] = np.array(W)
    for r in range(W):
        for k, v in results[r].items():
            out[f"r{r}_{k}"] = v
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, {k: v.shape for k, v in out.items() if not k.startswith("cfg_")})


def make_cfg1_case():
    """BASELINE.json configs[0]: ArcFace margin loss (s=64, m=0.5), 512-d embeddings, batch 128, 10 000 classes, one
    process.  dW / final weights / momentum are [10000, 512]: stored as their norm and a 16-column random projection."""
    cfg = dict(d=512, s=64.0, m=0.5, lr=0.1, momentum=0.9, wd=5e-4, steps=2, C=10000, b=128, sample_rate=1.0)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    p = ctx.Process(target=run_head_rank, args=(0, 1, 29625, cfg, q))
    p.start()
    _, rec = q.get()
    p.join()
    R = proj_matrix(cfg["d"])
    out = {"cfg_" + k: np.array(v) for k, v in cfg.items()}
    out["cfg_W"] = np.array(1)
    for k, v in rec.items():
        if v.ndim == 2 and v.shape[0] == cfg["C"]:
            out[f"r0_{k}_proj"] = (v.astype(np.float64) @ R).astype(np.float32)
            out[f"r0_{k}_norm"] = np.array(np.linalg.norm(v.astype(np.float64)))
        else:
            out[f"r0_{k}"] = v
    np.savez_compressed(os.path.join(HERE, "head_cfg1.npz"), **out)
    print("head_cfg1", {k: v.shape for k, v in out.items() if not k.startswith("cfg_")})


def make_margin_case():
    _, ArcFace, CosFace, CombinedMarginLoss = _import_reference()
    g = torch.Generator().manual_seed(5)
    B, n = 24, 40
    logits = (torch.rand(B, n, generator=g) * 2 - 1) * 0.98
    labels = torch.randint(0, n, (B, 1), generator=g)
    labels[::5] = -1
    out = {"logits": logits.numpy().copy(), "labels": labels.numpy().copy()}
    mods = {
        "arcface": ArcFace(64.0, 0.5), "arcface_30": ArcFace(30.0, 0.35), "cosface": CosFace(64.0, 0.4),
        "combined_arc": CombinedMarginLoss(64.0, 1.0, 0.5, 0.0), "combined_cos": CombinedMarginLoss(64.0, 1.0, 0.0, 0.4),
        "combined_filter": CombinedMarginLoss(64.0, 1.0, 0.5, 0.0, interclass_filtering_threshold=0.5),
    }
    for k, mod in mods.items():
        lg = logits.clone().requires_grad_(True)
        y = mod(lg.clone(), labels)
        y.backward(torch.ones_like(y))
        out[k] = y.detach().numpy().copy()
        out[k + "_grad"] = lg.grad.numpy().copy()
    np.savez_compressed(os.path.join(HERE, "margins.npz"), **out)
    print("margins", list(out))


def make_eval_case():
    sys.path.insert(0, REF)
    from utils.eval import pair_score, performance_roc, performance_acc, cross_score
    out = {}
    # small case with full inputs
    rng = np.random.default_rng(11)
    N, d = 400, 64
    e1 = rng.standard_normal((N, d)).astype(np.float32)
    e1 /= np.linalg.norm(e1, axis=1, keepdims=True)
    e2 = (e1 + 0.9 * rng.standard_normal((N, d)).astype(np.float32) / np.sqrt(d)).astype(np.float32)
    e2[N // 2:] = rng.standard_normal((N - N // 2, d)).astype(np.float32)
    e2 /= np.linalg.norm(e2, axis=1, keepdims=True)
    lab = np.zeros(N, dtype=bool)
    lab[: N // 2] = True
    hg, hi, sc = pair_score(e1, e2, lab)
    rep, th = performance_roc(hg, hi, 1, 3)
    acc = performance_acc(sc, lab, th)
    out.update(small_e1=e1, small_e2=e2, small_lab=lab, small_scores=sc, small_hg_nz=np.nonzero(hg)[0],
               small_hg_val=hg[np.nonzero(hg)[0]], small_hi_nz=np.nonzero(hi)[0], small_hi_val=hi[np.nonzero(hi)[0]],
               small_th=np.array(th), small_acc=np.array(acc), small_report=np.array(rep))
    # cross_score: all pairs of a small labelled set
    rngc = np.random.default_rng(21)
    ce = rngc.standard_normal((70, 64)).astype(np.float32)
    ce /= np.linalg.norm(ce, axis=1, keepdims=True)
    clab = rngc.integers(0, 9, 70)
    chg, chi, csc, clb = cross_score(ce, clab)
    out.update(cross_e=ce, cross_lab=clab, cross_scores=csc, cross_labels=clb, cross_hg_nz=np.nonzero(chg)[0],
               cross_hg_val=chg[np.nonzero(chg)[0]], cross_hi_nz=np.nonzero(chi)[0], cross_hi_val=chi[np.nonzero(chi)[0]])
    # cfg-5: inputs are regenerated from the seed, only results are stored
    a, b, lab5 = eval_inputs_cfg5()
    hg, hi, sc = pair_score(a, b, lab5)
    rep, th = performance_roc(hg, hi)   # default levels 3..9: FAR 1e-3 is reachable with 3000 imposters
    acc = performance_acc(sc, lab5, th)
    out.update(cfg5_scores=sc, cfg5_hg_nz=np.nonzero(hg)[0], cfg5_hg_val=hg[np.nonzero(hg)[0]],
               cfg5_hi_nz=np.nonzero(hi)[0], cfg5_hi_val=hi[np.nonzero(hi)[0]], cfg5_th=np.array(th),
               cfg5_acc=np.array(acc), cfg5_report=np.array(rep),
               cfg5_input_checksum=np.array([float(a.astype(np.float64).sum()), float(b.astype(np.float64).sum())]))
    np.savez_compressed(os.path.join(HERE, "eval.npz"), **out)
    print("eval: small th", th, "cfg5 th", out["cfg5_th"], "cfg5 acc", out["cfg5_acc"])


def make_adamw_cases():
    """PartialFCAdamW (nets/PartialFC.py:235-432): sampled (exp_avg / exp_avg_sq rows gathered and scattered back, the
    step count patched into the optimizer) and full."""
    base = dict(d=64, s=64.0, m=0.5, lr=1e-3, momentum=0.0, wd=0.05, steps=3, optimizer="adamw")
    make_head_case("head_w1_adamw_sampled", 1, dict(base, C=400, b=32, sample_rate=0.25), 29619)
    make_head_case("head_w2_adamw_sampled", 2, dict(base, C=401, b=16, sample_rate=0.5), 29620)
    make_head_case("head_w1_adamw_full", 1, dict(base, C=300, b=32, sample_rate=1.0), 29621)
    if "--adam-only" in sys.argv or "--adamw-only" not in sys.argv:
        make_head_case("head_w1_adam_sampled", 1, dict(base, C=400, b=32, sample_rate=0.25, optimizer="adam", wd=5e-2),
                       29624)


def make_filter_case():
    base = dict(d=64, s=64.0, m=0.5, lr=0.1, momentum=0.9, wd=5e-4, steps=3)
    # non-target cosines at d = 64 are ~N(0, 0.125): a threshold of 0.25 filters ~2 % of them
    make_head_case("head_w1_filter", 1, dict(base, C=300, b=32, sample_rate=1.0, margin="combined_filter",
                                             filter_thr=0.25), 29622)
    # The filter is a step function of the cosine: with bf16 operands (the CUDA path) a logit within ~2e-3 of the
    # threshold can land on the other side and move the loss by e^(64*thr).  The first case (dozens of such logits)
    # pins the fp64 oracle; this one, whose non-target cosines all stay >= 0.02 away from the threshold over the three
    # steps (two logits filtered), is the one bf16-operand implementations are compared with.
    make_head_case("head_w1_filter_wide", 1, dict(base, C=300, b=32, sample_rate=1.0, margin="combined_filter",
                                                  filter_thr=0.47), 29623)


if __name__ == "__main__":
    if "--adamw-only" in sys.argv:      # added after the other fixtures: leaves them untouched
        make_adamw_cases()
        sys.exit(0)
    if "--adam-only" in sys.argv:
        make_head_case("head_w1_adam_sampled", 1, dict(d=64, s=64.0, m=0.5, lr=1e-3, momentum=0.0, wd=5e-2, steps=3,
                                                       optimizer="adam", C=400, b=32, sample_rate=0.25), 29624)
        sys.exit(0)
    if "--d128-only" in sys.argv:       # round 2: d a multiple of 128 and several 256-class tiles per rank, multi-step
        b128 = dict(d=128, s=64.0, m=0.5, lr=0.1, momentum=0.9, wd=5e-4, steps=3)     # (conf.lazy_update, FX groups)
        make_head_case("head_w1_d128", 1, dict(b128, C=777, b=48, sample_rate=1.0), 29631)
        make_head_case("head_w2_d128", 2, dict(b128, C=1101, b=24, sample_rate=1.0), 29632)
        sys.exit(0)
    if "--cfg1-only" in sys.argv:
        make_cfg1_case()
        sys.exit(0)
    if "--filter-only" in sys.argv:
        make_filter_case()
        sys.exit(0)
    base = dict(d=64, s=64.0, m=0.5, lr=0.1, momentum=0.9, wd=5e-4, steps=3)
    make_head_case("head_w1_full", 1, dict(base, C=300, b=32, sample_rate=1.0), 29611)
    make_head_case("head_w1_s30", 1, dict(base, C=300, b=32, sample_rate=1.0, s=30.0, m=0.35, sigma=0.6), 29612)
    make_head_case("head_w1_cosface", 1, dict(base, C=300, b=32, sample_rate=1.0, m=0.4, margin="cosface"), 29613)
    make_head_case("head_w1_sampled", 1, dict(base, C=400, b=32, sample_rate=0.25), 29614)
    make_head_case("head_w1_manypos", 1, dict(base, C=64, b=32, sample_rate=0.25), 29615)
    make_head_case("head_w2_full", 2, dict(base, C=301, b=16, sample_rate=1.0), 29616)
    make_head_case("head_w2_sampled", 2, dict(base, C=401, b=16, sample_rate=0.5), 29617)
    make_head_case("head_w1_d512", 1, dict(base, C=520, b=64, d=512, sample_rate=1.0, steps=1), 29618)
    make_adamw_cases()
    make_filter_case()
    make_cfg1_case()
    make_margin_case()
    make_eval_case()
