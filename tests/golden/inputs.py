"""Seeded synthetic inputs shared by make_golden.py (which feeds them to the reference) and the tests
(which feed the same tensors to the oracle and the CUDA path).  SURVEY.md section 8(d)."""
import math

import numpy as np
import torch


def synth_inputs(C, d, B, steps, seed_w=1234, seed_l=7, seed_x=42, sigma=1.0, n_neg_aligned=2):
    """N(0,0.01) weights, uniform labels, trained-like embeddings plus a few rows aligned
    with -W[y] so that the t <= cos(pi-m) branch of the margin is exercised (|t| < 1 throughout)."""
    g = torch.Generator().manual_seed(seed_w)
    w_full = torch.normal(0, 0.01, (C, d), generator=g)
    xs, ls = [], []
    for s in range(steps):
        gl = torch.Generator().manual_seed(seed_l + s)
        labels = torch.randint(0, C, (B,), generator=gl)
        gx = torch.Generator().manual_seed(seed_x + s)
        wy = torch.nn.functional.normalize(w_full[labels])
        x = wy + sigma * torch.randn(B, d, generator=gx) / math.sqrt(d)
        x[:n_neg_aligned] = -wy[:n_neg_aligned] + 0.15 * torch.randn(n_neg_aligned, d, generator=gx) / math.sqrt(d)
        xs.append(torch.nn.functional.normalize(x))
        ls.append(labels)
    return w_full, xs, ls


def shard(C, rank, W):
    nl = C // W + int(rank < C % W)
    cs = C // W * rank + min(rank, C % W)
    return nl, cs


def eval_inputs_cfg5(N=6000, d=512):
    """SURVEY.md section 8(d) cfg-5 synthetic verification set."""
    rng = np.random.default_rng(2024)
    a = rng.standard_normal((N, d))
    a /= np.linalg.norm(a, axis=1, keepdims=True)
    n = rng.standard_normal((N, d))
    n -= (n * a).sum(1, keepdims=True) * a
    n /= np.linalg.norm(n, axis=1, keepdims=True)
    lab = np.zeros(N, dtype=bool)
    per = N // 10
    for f in range(10):
        lab[f * per: f * per + per // 2] = True
    rho = np.where(lab, rng.normal(0.55, 0.18, N), rng.normal(0.08, 0.12, N)).clip(-0.99, 0.99)
    b = rho[:, None] * a + np.sqrt(1 - rho ** 2)[:, None] * n
    return a.astype(np.float32), b.astype(np.float32), lab


def proj_matrix(d, k=16, seed=99):
    """Fixed random projection that keeps the [C, d] arrays of the cfg-1 fixture small: X -> X @ R, R [d, k]."""
    return torch.randn(d, k, generator=torch.Generator().manual_seed(seed)).numpy().astype(np.float64)
