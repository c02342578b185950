"""GPU bring-up probe: runs each CUDA kernel family in its own subprocess (a trapped kernel cannot take the rest
down) and prints error statistics against plain torch.  For the MN-major tcgen05 operands it also tries alternative
shared-memory descriptor geometries, so one GPU call says which encoding is right.

    python tools/gpu_probe.py            # all cases
    python tools/gpu_probe.py --case dx  # one case in-process
"""
import argparse
import ctypes
import math
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = ["rows", "fwd_small", "fwd_mid", "dx", "dw", "sample", "eval", "fp16", "fwd_big", "dx_big", "dw_big"]


def _stats(name, got, ref):
    import torch
    got, ref = got.double().flatten(), ref.double().flatten()
    err = (got - ref).abs()
    den = ref.abs().max().clamp_min(1e-30)
    cos = float((got @ ref) / (got.norm() * ref.norm()).clamp_min(1e-300))
    print(f"  {name:28s} max_abs_err={float(err.max()):.3e} rel_to_max={float(err.max() / den):.3e} "
          f"cos={cos:.8f} ref_max={float(den):.3e} nan={bool(torch.isnan(got).any())}", flush=True)
    return float(err.max() / den), cos


def _mk(B, n, d, dtype=None, seed=0):
    import torch
    g = torch.Generator(device="cpu").manual_seed(seed)
    dtype = dtype or torch.bfloat16
    w = torch.nn.functional.normalize(torch.randn(n, d, generator=g)).cuda()
    lab = torch.randint(0, n, (B,), generator=g)
    x = torch.nn.functional.normalize(w[lab].cpu() + 1.0 * torch.randn(B, d, generator=g) / math.sqrt(d)).cuda()
    return x.to(dtype).contiguous(), w.to(dtype).contiguous(), lab.cuda()


def case_rows():
    import torch
    from face_recognition_pytorch_b200 import kernels as K
    x = torch.randn(1000, 512, device="cuda") * 0.3
    xn = torch.empty(1000, 512, dtype=torch.bfloat16, device="cuda")
    inv = torch.empty(1000, device="cuda")
    K.l2norm_rows(x, None, 1000, xn, inv)
    ref = torch.nn.functional.normalize(x)
    _stats("l2norm xn", xn.float(), ref.to(torch.bfloat16).float())
    _stats("l2norm inv", inv, 1 / x.norm(dim=1))
    print("  bf16 bit-equal fraction:", float((xn == ref.to(torch.bfloat16)).float().mean()))


def _fwd(B, n, d, s=64.0, m=0.5, dtype=None):
    import torch
    from face_recognition_pytorch_b200 import kernels as K
    xn, wn, lab = _mk(B, n, d, dtype)
    n_pad = K.padded_classes(n)
    E = torch.zeros(B, n_pad, dtype=torch.bfloat16, device="cuda")
    nt = K.num_class_tiles(n)
    Bp = K.padded_batch(B)
    part = torch.zeros(nt, Bp, device="cuda")
    traw, te, tz = (torch.zeros(B, device="cuda") for _ in range(3))
    lab32 = lab.to(torch.int32)
    lab32[::7] = -1
    t0 = time.time()
    K.forward(xn, wn, lab32, B, n, d, s, 0, m, 0.0, 0.0, E, n_pad, part, traw, te, tz)
    torch.cuda.synchronize()
    print(f"  forward launched+synced in {time.time() - t0:.3f}s", flush=True)
    raw = xn.float() @ wn.float().t()
    cl = raw.clamp(-1, 1)
    k1 = s * 1.4426950408889634
    e = torch.exp2(cl * k1 - (k1 - K.exp_top()))
    rows = torch.nonzero(lab32 >= 0).flatten()
    cols = lab32[rows].long()
    e_nt = e.clone()
    e_nt[rows, cols] = 0
    L_ref = e_nt.sum(1)
    r1, _ = _stats("row sum (non-target)", part.sum(0)[:B], L_ref)
    Eg = K.spill_to_rowmajor(E, B, n_pad)[:, :n].float()
    Eg_cmp = Eg.clone()
    Eg_cmp[rows, cols] = 0
    r2, c2 = _stats("E spill (non-target)", torch.log2(Eg_cmp.clamp_min(1e-30)), torch.log2(e_nt.clamp_min(1e-30)))
    t = cl[rows, cols]
    fin = torch.where(t > math.cos(math.pi - m), t * math.cos(m) - torch.sqrt(1 - t * t) * math.sin(m),
                      t - math.sin(math.pi - m) * m)
    _stats("target raw", traw[rows], raw[rows, cols])
    _stats("target z", tz[rows], fin * s)
    _stats("target e (log2)", torch.log2(te[rows]), fin * k1 - (k1 - K.exp_top()))
    ok = r1 < 2e-3 and r2 < 1e-2
    print("  FWD", "OK" if ok else "MISMATCH", flush=True)
    return ok


def case_fwd_small():
    return _fwd(128, 256, 64)


def case_fwd_mid():
    return _fwd(300, 1000, 512) and _fwd(1024, 4099, 512, s=30.0, m=0.35)


def case_fwd_big():
    return _fwd(1024, 93431, 512)


def _set_desc(lbo, sbo, kstep):
    from face_recognition_pytorch_b200 import _lib
    fn = _lib.lib.pfc_debug_mn_desc
    fn.argtypes = [ctypes.c_uint, ctypes.c_uint, ctypes.c_uint]
    fn.restype = None
    fn(lbo, sbo, kstep)


ALT_DESCS = [(0, 0, 0), (1024, 8192, 2048), (8192, 1024, 1024), (1024, 8192, 1024), (16, 1024, 2048),
             (8192, 128, 2048), (128, 8192, 2048)]


def _dx(B, n, d, try_alts=True):
    import torch
    from face_recognition_pytorch_b200 import kernels as K
    g = torch.Generator().manual_seed(3)
    n_pad = K.padded_classes(n)
    E = torch.zeros(B, n_pad, dtype=torch.bfloat16, device="cuda")
    E[:, :n] = torch.rand(B, n, generator=g).cuda().to(torch.bfloat16)
    wn = torch.nn.functional.normalize(torch.randn(n, d, generator=g)).cuda().to(torch.bfloat16).contiguous()
    ref = E[:, :n].float() @ wn.float()
    E = K.spill_from_rowmajor(E)                 # the kernels read the class-blocked layout
    splits = K.dx_splits(B, n, d)
    print(f"  dx B={B} n={n} d={d} splits={splits} max={K.dx_max_splits(B, d)}", flush=True)
    for cfg in (ALT_DESCS if try_alts else ALT_DESCS[:1]):
        _set_desc(*cfg)
        part = torch.zeros(splits, B, d, device="cuda")
        K.backward_dx(E, n_pad, wn, B, n, d, part, splits)
        torch.cuda.synchronize()
        r, c = _stats(f"dx desc{cfg}", part.sum(0), ref)
        if r < 1e-3:
            print("  DX OK with", cfg, flush=True)
            _set_desc(0, 0, 0)
            return cfg == (0, 0, 0)
    _set_desc(0, 0, 0)
    print("  DX MISMATCH for every descriptor geometry", flush=True)
    return False


def _dw(B, n, d, try_alts=True):
    import torch
    from face_recognition_pytorch_b200 import kernels as K
    g = torch.Generator().manual_seed(4)
    n_pad = K.padded_classes(n)
    E = torch.zeros(B, n_pad, dtype=torch.bfloat16, device="cuda")
    E[:, :n] = torch.rand(B, n, generator=g).cuda().to(torch.bfloat16)
    xs = (torch.randn(B, d, generator=g) * 0.05).cuda().to(torch.bfloat16).contiguous()
    ref = E[:, :n].float().t() @ xs.float()
    E = K.spill_from_rowmajor(E)                 # the kernels read the class-blocked layout
    for cfg in (ALT_DESCS if try_alts else ALT_DESCS[:1]):
        _set_desc(*cfg)
        dwn = torch.zeros(n, d, device="cuda")
        K.backward_dw(E, n_pad, xs, B, n, d, dwn)
        torch.cuda.synchronize()
        r, c = _stats(f"dw desc{cfg}", dwn, ref)
        if r < 1e-3:
            print("  DW OK with", cfg, flush=True)
            _set_desc(0, 0, 0)
            return cfg == (0, 0, 0)
    _set_desc(0, 0, 0)
    print("  DW MISMATCH for every descriptor geometry", flush=True)
    return False


def case_dx():
    return _dx(128, 256, 256) and _dx(300, 1000, 512, False) and _dx(64, 777, 64, False)


def case_dw():
    return _dw(128, 256, 256) and _dw(300, 1000, 512, False) and _dw(64, 777, 64, False)


def case_dx_big():
    return _dx(1024, 93431, 512, False)


def case_dw_big():
    return _dw(1024, 93431, 512, False)


def case_fp16():
    """fp16 operands (conf.mixed_precision): the forward with fp16 Xn / Wn against an fp32 matmul of the same rounded
    values, the row kernels' fp16 output against torch's rounding, and the fp16 -> bf16 cast of the shard for the dX
    contraction (tcgen05 kind::f16 raises an illegal-instruction error for a mixed bf16 x fp16 operand pair)."""
    import torch
    from face_recognition_pytorch_b200 import kernels as K
    ok = _fwd(300, 1000, 512, dtype=torch.float16) and _fwd(128, 256, 64, dtype=torch.float16)
    x = torch.randn(1000, 512, device="cuda") * 0.3
    xn = torch.empty(1000, 512, dtype=torch.float16, device="cuda")
    inv = torch.empty(1000, device="cuda")
    K.l2norm_rows(x, None, 1000, xn, inv)
    ref = torch.nn.functional.normalize(x).to(torch.float16)
    frac = float((xn == ref).float().mean())
    print("  l2norm fp16 bit-equal fraction:", frac, flush=True)
    xb = torch.empty(1000, 512, dtype=torch.bfloat16, device="cuda")
    K.cast_f16_to_bf16(xn, xb, xn.numel())
    cast_ok = torch.equal(xb, xn.to(torch.bfloat16))
    print("  fp16 -> bf16 cast bit-equal:", cast_ok, flush=True)
    return ok and frac > 0.99 and cast_ok


def case_sample():
    """The sampler against the oracle (index set and remapped labels, with forced ties)."""
    return _case_sample_once()


def _case_sample_once():
    """Every path of pfc_sample: the cluster kernel with 16 and with 8 CTAs, the tiled six-launch path, and the automatic
    choice."""
    import torch
    from face_recognition_pytorch_b200 import kernels as K
    from face_recognition_pytorch_b200._lib import lib
    from oracle import head_oracle as ho
    ok = True
    cases = [(400, 100, 32, 1), (45029, 4502, 1024, 2), (257489, 51497, 4096, 3), (64, 16, 32, 4), (5000, 0, 16, 5),
             (1000, 1000, 8, 6), (2000000, 400000, 4096, 7), (400000, 3, 4096, 8)]
    refs = {}
    try:
        for mode in (0, 16, 8, -1):
            lib.pfc_sample_debug_cluster(mode)
            for nl, ns, B, seed in cases:
                g = torch.Generator().manual_seed(seed)
                perm = torch.rand(nl, generator=g)
                perm = torch.floor(perm * 4096) / 4096                               # force plenty of ties
                lab = torch.randint(-1, nl, (B,), generator=g).to(torch.int32)
                if seed not in refs:
                    refs[seed] = ho.sample_indices(perm, lab.long(), ns)
                idx_ref, lab_ref = refs[seed]
                ws = torch.zeros(K.sample_workspace_bytes(nl), dtype=torch.uint8, device="cuda")
                idx = torch.zeros(max(ns, B), dtype=torch.int64, device="cuda")
                n_out = torch.zeros(1, dtype=torch.int32, device="cuda")
                rem = torch.zeros(B, dtype=torch.int32, device="cuda")
                for _ in range(2):                                                    # the second call reuses the workspace
                    K.sample(perm.cuda(), lab.cuda(), nl, ns, idx, n_out, rem, ws)
                n = int(n_out.item())
                good = (n == idx_ref.numel() and torch.equal(idx[:n].cpu(), idx_ref)
                        and torch.equal(rem.cpu().long(), lab_ref))
                print(f"  sample mode={mode} launches={K.sample_launches(nl)} nl={nl} ns={ns} B={B}: n={n} "
                      f"ref_n={idx_ref.numel()} {'OK' if good else 'MISMATCH'}", flush=True)
                ok &= good
    finally:
        lib.pfc_sample_debug_cluster(0)
    return ok


def case_eval():
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    from inputs import eval_inputs_cfg5
    import face_recognition_pytorch_b200 as pfc
    from oracle import eval_oracle as eo
    a, b, lab = eval_inputs_cfg5()
    hg, hi, sc = pfc.pair_score(a, b, lab)
    hg2, hi2, sc2 = eo.pair_score(a, b, lab)
    print("  hist equal:", np.array_equal(hg, hg2), np.array_equal(hi, hi2), "score max diff", np.abs(sc - sc2).max())
    rep, th = pfc.performance_roc(hg, hi)
    rep2, th2 = eo.performance_roc(hg2, hi2)
    acc = pfc.performance_acc(sc, lab, th)
    print("  th", th, th2, "acc", acc, eo.performance_acc(sc2, lab, th2), "report equal:", rep == rep2)
    kacc, kbest = pfc.kfold_accuracy(a, b, lab)
    kacc2, kbest2 = eo.kfold_accuracy(4 * (1 - sc2), lab)
    print("  kfold", kacc.mean(), kacc2.mean(), np.array_equal(kbest, kbest2))
    return th == th2 == 63399 and rep == rep2 and np.array_equal(hg, hg2)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", default=None)
    ap.add_argument("--skip-big", action="store_true")
    args = ap.parse_args()
    if args.case:
        import torch
        print(f"[{args.case}] device={torch.cuda.get_device_name(0)} sms={torch.cuda.get_device_properties(0).multi_processor_count}",
              flush=True)
        mode = int(os.environ.get("PFC_GEMM_MODE", "0"))   # 0 auto, 1 single CTA, 2 multicast pair, 3 cta_group::2 pair
        if mode:
            from face_recognition_pytorch_b200 import _lib
            _lib.lib.pfc_debug_cluster(mode)
            print(f"  (GEMM mode {mode})", flush=True)
        ok = globals()["case_" + args.case]()
        sys.exit(0 if ok or ok is None else 3)
    results = {}
    for c in CASES:
        if args.skip_big and c.endswith("_big"):
            continue
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--case", c], timeout=240, cwd=ROOT)
            results[c] = r.returncode
        except subprocess.TimeoutExpired:
            results[c] = "timeout"
        print(f"== case {c}: rc={results[c]} ({time.time() - t0:.1f}s)", flush=True)
    print("PROBE SUMMARY", results, flush=True)


if __name__ == "__main__":
    main()
