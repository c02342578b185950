"""EXPERIMENT (one GPU): can the HBM-bound fused SGD update run UNDERNEATH the tensor-bound gradient GEMMs of the same step?

The dW / dX GEMM kernels are persistent (one 320-thread CTA per SM, 96 registers, ~200 KB shared memory), which leaves
~34 K registers and 1 700 thread slots per SM: room for four 128-thread CTAs of the update kernel (64 registers, no
shared memory).  This script times, at BASELINE configs[1] (B = 1024, n = 93 431, d = 512):
  * each kernel alone (dW, dX, forward, update full-grid / persistent);
  * GEMM || update on two streams (different buffers);
  * the pipelined backward: dW in K class chunks on the main stream, the update of chunk k on a side stream as soon as
    dW(k) is done (its bf16 gradient chunk is still in L2), dX last on the main stream while the remaining updates run;
    the update writes next step's bf16 shard into a second buffer, so dX keeps reading this step's.
Prints one line per variant (median of --reps, L2 flushed before each).   python tools/exp_overlap.py
"""
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
    ap.add_argument("--reps", type=int, default=15)
    ap.add_argument("--classes", type=int, default=93431)
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--l2-only", action="store_true", help="only the variants with the L2 hints on")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    dist.init_process_group("nccl", init_method="tcp://127.0.0.1:29547", rank=0, world_size=1, device_id=dev)
    import bench
    import face_recognition_pytorch_b200 as pfc
    from face_recognition_pytorch_b200 import kernels as K
    lib = pfc._lib.lib
    cfg = dict(bench.CONFIGS[2], C=args.classes, B=args.batch)
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
    w = head.weight_activated.data
    mom = head._fused_state
    splits = K.dx_splits(B, n, d)
    dwn = ws.grad_buffer(True, d)
    wn2 = torch.empty_like(ws.wn)
    inv2 = torch.empty_like(ws.inv_w)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    LR = 1e-7      # the update really runs (in place, hundreds of times): keep the weights where they are
    lib.pfc_debug_l2_grad(0)

    def fwd():
        K.forward(ws.xn_all, ws.wn, ws.labels_act, B, n, d, s, kind, m2, m3, thr, ws.E, n_pad, ws.part_sum, ws.tgt_raw,
                  ws.tgt_e, ws.tgt_z)

    def dw(c0=0, c1=None, hint=False):
        c1 = n if c1 is None else c1
        nn = c1 - c0
        K.backward_dw(ws.E[c0 * B:], K.padded_classes(nn), ws.xs, B, nn, d, dwn[c0:c1], keep_in_l2=hint)

    def dx():
        K.backward_dx(ws.E, n_pad, ws.wn, B, n, d, ws.dx_partial, splits)

    def upd(c0=0, c1=None, wn_out=None, inv_out=None):
        c1 = n if c1 is None else c1
        wn_out = wn2 if wn_out is None else wn_out
        inv_out = inv2 if inv_out is None else inv_out
        K.dw_sgd(dwn[c0:c1], w[c0:c1], mom[c0:c1], ws.inv_w[c0:c1], c1 - c0, d, LR, 0.9, 5e-4, None, wn_out[c0:c1],
                 inv_out[c0:c1])

    lo_pri, hi_pri = torch.cuda.Stream.priority_range() if hasattr(torch.cuda.Stream, "priority_range") else (0, -1)
    main_s = torch.cuda.Stream(device=dev, priority=-1)
    side_s = torch.cuda.Stream(device=dev, priority=0)

    def timed(name, fn, note=""):
        ts = []
        for r in range(args.reps + 2):
            flush.zero_()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(main_s):
                e0.record()
                fn()
                e1.record()
            torch.cuda.synchronize()
            if r >= 2:
                ts.append(e0.elapsed_time(e1) * 1e3)
        print(f"{name:58s} {statistics.median(ts):8.1f} us   (min {min(ts):7.1f}) {note}", flush=True)
        return statistics.median(ts)

    def both(gemm, upd_first):
        def run():
            ev = torch.cuda.Event()
            ev.record()
            side_s.wait_event(ev)
            if upd_first:
                with torch.cuda.stream(side_s):
                    upd()
                gemm()
            else:
                gemm()
                with torch.cuda.stream(side_s):
                    upd()
            main_s.wait_stream(side_s)
        return run

    def pipeline(chunks, dx_last=True, hint=False):
        # class chunks on 256-class boundaries
        tiles = (n + 255) // 256
        cuts = [min(n, 256 * ((tiles * k) // chunks)) for k in range(chunks + 1)]
        cuts[-1] = n

        def run():
            if not dx_last:
                dx()
            for k in range(chunks):
                dw(cuts[k], cuts[k + 1], hint)
                ev = torch.cuda.Event()
                ev.record()
                side_s.wait_event(ev)
                with torch.cuda.stream(side_s):
                    upd(cuts[k], cuts[k + 1])
            if dx_last:
                dx()
            main_s.wait_stream(side_s)
        return run

    print(f"B={B} n={n} d={d} splits={splits}")
    for pw in (() if args.l2_only else (0, 8, 16)):
        lib.pfc_debug_sgd_persistent(pw)
        tag = "full grid" if pw == 0 else f"persistent {pw} warps/SM"
        print(f"--- update kernel: {tag}")
        if pw == 0:
            t_f = timed("forward alone", fwd)
            t_dw = timed("dW alone", dw)
            t_dx = timed("dX alone", dx)
        t_u = timed("update alone", upd)
        if pw == 0:
            timed("serial: dW, dX, update", lambda: (dw(), dx(), upd()), f"sum {t_dw + t_dx + t_u:.1f}")
            timed("serial with L2 hints: dW(hint), update(l2)", lambda: (dw(hint=True), upd()))
        for nm, g in (("dW", dw), ("dX", dx), ("forward", fwd)):
            timed(f"{nm} || update (GEMM launched first)", both(g, False))
            timed(f"{nm} || update (update launched first)", both(g, True))
        for ch in (2, 4, 8):
            timed(f"pipeline: {ch} x [dW chunk -> update chunk on side], dX last", pipeline(ch))
        timed("pipeline: 4 chunks, dX first", pipeline(4, dx_last=False))
    lib.pfc_debug_sgd_persistent(0)
    lib.pfc_debug_l2_grad(1)
    print("--- with L2 hints on the chunked gradient (evict_last stores, discard after use), update full grid")
    for ch in (3, 4, 6):
        timed(f"pipeline+L2: {ch} chunks, dX last", pipeline(ch, hint=True))
    # serial chunked with L2 hints (no overlap): what keeping the gradient in L2 alone buys
    def serial_chunks(ch):
        tiles = (n + 255) // 256
        cuts = [min(n, 256 * ((tiles * k) // ch)) for k in range(ch + 1)]
        cuts[-1] = n

        def run():
            for k in range(ch):
                dw(cuts[k], cuts[k + 1], True)
                upd(cuts[k], cuts[k + 1])
            dx()
        return run
    for ch in (3, 4, 6):
        timed(f"serial chunks+L2: {ch} x [dW chunk, update chunk], dX", serial_chunks(ch))
    # the update with evict_first streams next to dX (which lives on L2 reuse of its operand stages)
    timed("L2: dW(hint), update, dX (serial)", lambda: (dw(hint=True), upd(), dx()))
    timed("L2: dX, dW(hint), update (serial, early-dX order)", lambda: (dx(), dw(hint=True), upd()))

    def dw_then_both():
        dw(hint=True)
        ev = torch.cuda.Event()
        ev.record()
        side_s.wait_event(ev)
        with torch.cuda.stream(side_s):
            upd()
        dx()
        main_s.wait_stream(side_s)
    timed("L2: dW(hint), then [update || dX]", dw_then_both)
    for pw in (8, 12, 16):
        lib.pfc_debug_sgd_persistent(pw)
        timed(f"L2: dW(hint), then [update persistent {pw} warps/SM || dX]", dw_then_both)
    lib.pfc_debug_sgd_persistent(0)


if __name__ == "__main__":
    main()
