"""Device timings of the HBM-/latency-bound kernels AROUND the GEMMs -- normalise, negative-class sampling (radix
select), row gather / scatter, verification scorer -- at the BASELINE shapes, as achieved GB/s against the measured
copy peak (MEASURED_PEAKS.json).  CUDA events per launch, L2 flushed (256 MB write) between launches, median of 15.

    python tools/bench_aux.py [--out gpurun_out/bench_aux.json]

Algorithmic bytes: normalise rows*d*(4+2)+4*rows; sampling nl*4 read + n*8 written (+ B labels);
gather / scatter n*d*4 read + written per tensor; scorer N*2*d*4.
"""
import argparse
import json
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np   # noqa: E402
import torch         # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "bench_aux.json"))
    ap.add_argument("--iters", type=int, default=15)
    args = ap.parse_args()
    import torch.distributed as dist
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    if not dist.is_initialized():
        dist.init_process_group("nccl", init_method="tcp://127.0.0.1:29549", rank=0, world_size=1, device_id=dev)
    import face_recognition_pytorch_b200 as pfc
    from face_recognition_pytorch_b200 import kernels as K
    sys.path.insert(0, ROOT)
    import bench
    peak = bench.peaks()["hbm"]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    rows_out = []

    def timed(name, fn, nbytes, launches=1, note=""):
        try:
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            ms = []
            for _ in range(args.iters):
                flush.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                fn()
                b.record()
                torch.cuda.synchronize()
                ms.append(a.elapsed_time(b))
            med = statistics.median(ms)
            r = {"kernel": name, "us": round(med * 1e3, 2), "us_min": round(min(ms) * 1e3, 2), "launches": launches,
                 "algorithmic_bytes": int(nbytes), "GB/s": round(nbytes / (med * 1e-3) / 1e9, 1),
                 "frac_of_hbm_peak": round(nbytes / (med * 1e-3) / 1e9 / peak, 4), "note": note}
        except Exception as e:   # keep going: one failing section must not lose the others
            r = {"kernel": name, "error": f"{type(e).__name__}: {e}"}
        rows_out.append(r)
        print(json.dumps(r), flush=True)

    g = torch.Generator().manual_seed(1)
    d = 512
    # ---- normalise: the class shard (first step / un-fused mode) and one batch
    for nm, rows in (("l2norm_rows W cfg-2 [93431,512]", 93431), ("l2norm_rows X [1024,512]", 1024)):
        x = torch.randn(rows, d, generator=g).to(dev)
        xn = torch.empty(rows, d, dtype=torch.bfloat16, device=dev)
        inv = torch.empty(rows, device=dev)
        timed(nm, lambda: K.l2norm_rows(x, None, rows, xn, inv), rows * d * 6 + 4 * rows)
        del x, xn, inv

    # ---- sampling + gather / scatter at the per-rank shapes of cfg-3 / cfg-4
    for nm, nl, rate, B in (("cfg-3 rank shape", 45029, 0.1, 1024), ("cfg-4 rank shape", 257489, 0.2, 4096)):
        k = int(rate * nl)
        n_max = max(k, min(B, nl))
        perm = torch.rand(nl, generator=g).to(dev)
        lab = torch.randint(0, nl * 8, (B,), generator=g)
        lab = torch.where(lab < nl, lab, torch.full_like(lab, -1)).to(torch.int32).to(dev)
        index = torch.zeros(n_max, dtype=torch.int64, device=dev)
        n_out = torch.zeros(1, dtype=torch.int32, device=dev)
        remap = torch.zeros(B, dtype=torch.int32, device=dev)
        wsb = torch.zeros(K.sample_workspace_bytes(nl), dtype=torch.uint8, device=dev)
        timed(f"pfc_sample {nm} (nl={nl}, k={k}, B={B})",
              lambda: K.sample(perm, lab, nl, k, index, n_out, remap, wsb), nl * 4 + k * 8 + B * 8, launches=1,
              note="radix select in ONE cluster of 8 CTAs: bitmap of the positives + 3 histogram passes in shared memory, "
                   "merged through DSMEM, ordered compaction + label remap")
        torch.cuda.synchronize()
        n = int(n_out.item())
        w = torch.randn(nl, d, generator=g).to(dev)
        m = torch.zeros(nl, d, device=dev)
        wa = torch.empty(n_max, d, device=dev)
        ma = torch.empty(n_max, d, device=dev)
        timed(f"pfc_gather_rows {nm} (n={n}, weight + momentum)",
              lambda: K.gather_rows([w, m], [wa[:n], ma[:n]], index[:n], n), 2 * n * d * 4 * 2)
        timed(f"pfc_scatter_rows {nm} (n={n}, weight + momentum)",
              lambda: K.scatter_rows([wa[:n], ma[:n]], [w, m], index[:n], n), 2 * n * d * 4 * 2)
        # what the fused step does instead of gather / scatter (conf.inplace_update): normalise through the index list, and
        # the SGD step on rows index[r] of the full arrays in place
        wn = torch.empty(n_max, d, dtype=torch.bfloat16, device=dev)
        inv = torch.empty(n_max, device=dev)
        timed(f"pfc_l2norm_rows through the index list {nm} (n={n})",
              lambda: K.l2norm_rows(w, index, n, wn, inv), n * d * 6 + 12 * n)
        gbf = torch.zeros(n_max, d, dtype=torch.bfloat16, device=dev)
        timed(f"pfc_dw_sgd in place through the index list {nm} (n={n})",
              lambda: K.dw_sgd(gbf, w, m, inv, n, d, 1e-7, 0.9, 5e-4, None, None, None, index=index),
              n * d * 18 + 12 * n, note="bf16 gradient + fp32 weight / momentum read and written")
        del w, m, wa, ma, perm, wn, inv, gbf

    # ---- verification scorer, cfg-5 (6000 pairs x 512)
    N = 6000
    rng = np.random.default_rng(2024)
    a = rng.standard_normal((N, d)); a /= np.linalg.norm(a, axis=1, keepdims=True)
    nz = rng.standard_normal((N, d)); nz -= (nz * a).sum(1, keepdims=True) * a; nz /= np.linalg.norm(nz, axis=1, keepdims=True)
    labn = np.zeros(N, bool)
    for f in range(10):
        labn[f * 600: f * 600 + 300] = True
    rho = np.where(labn, rng.normal(0.55, 0.18, N), rng.normal(0.08, 0.12, N)).clip(-0.99, 0.99)
    bb = rho[:, None] * a + np.sqrt(1 - rho ** 2)[:, None] * nz
    e1n, e2n = a.astype(np.float32), bb.astype(np.float32)
    e1, e2 = torch.from_numpy(e1n).to(dev), torch.from_numpy(e2n).to(dev)
    lab8 = torch.from_numpy(labn.astype(np.uint8)).to(dev)
    bins = K.hist_bins()
    scores = torch.empty(N, dtype=torch.float64, device=dev)
    dist_ = torch.empty(N, dtype=torch.float64, device=dev)
    hh = torch.empty(2, bins, dtype=torch.int64, device=dev)      # back to back: one memset zeroes both
    hg, hi = hh[0], hh[1]
    timed("fr_pair_score cfg-5 (6000 x 512)", lambda: K.pair_score(e1, e2, lab8, scores, dist_, hg, hi), N * 2 * d * 4,
          note="includes zeroing + filling the two 100001-bin histograms")
    import ctypes
    from face_recognition_pytorch_b200 import _lib
    buf = torch.zeros(ctypes.sizeof(_lib.RocOut), dtype=torch.uint8, device=dev)
    timed("fr_roc (100000-threshold sweep, FAR 1e-3..1e-9 + EER)", lambda: K.roc(hg, hi, 3, 9, buf), 2 * bins * 8,
          note="one cluster of 8 CTAs, histograms staged in shared memory, DSMEM exchange of the CTA totals / winners")
    out2 = torch.zeros(2, dtype=torch.int64, device=dev)
    timed("fr_acc_counts (6000 scores)", lambda: K.acc_counts(scores, lab8, 0.63399, out2), N * 9)
    ws = torch.zeros(10 * 400, dtype=torch.int32, device=dev)
    acc = torch.zeros(10, dtype=torch.float64, device=dev)
    best = torch.zeros(10, dtype=torch.int32, device=dev)
    timed("fr_kfold_acc (10 folds x 400 thresholds)", lambda: K.kfold_acc(dist_, lab8, 10, 400, 0.01, ws, acc, best),
          N * 9, launches=1, note="one CTA: per-fold flip-threshold histograms in shared memory, prefix sums, argmax")

    # ---- the reference-facing call: NumPy in, report out (H2D + kernels + D2H + host formatting), wall clock
    try:
        t = []
        for _ in range(5):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            hgn, hin, sc = pfc.pair_score(e1n, e2n, labn)
            rep, th = pfc.performance_roc(hgn, hin, 3, 9)
            ac = pfc.performance_acc(sc, labn, th)
            t.append(time.perf_counter() - t0)
        r = {"kernel": "e2e pair_score + performance_roc + performance_acc (NumPy in / out)",
             "ms_wall": round(statistics.median(t) * 1e3, 3), "eer_threshold": int(th), "acc": ac,
             "note": "reference on CPU in this container (SURVEY 8a): numba pair_score 33 ms + Python ROC loop 260 ms + "
                     "acc loop 24 ms"}
    except Exception as e:
        r = {"kernel": "e2e eval", "error": f"{type(e).__name__}: {e}"}
    rows_out.append(r)
    print(json.dumps(r), flush=True)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump({"hbm_peak_GBps": peak, "rows": rows_out}, f, indent=1)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
