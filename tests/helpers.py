"""Shared test helpers: golden fixture access and the oracle driven with fixture inputs."""
import os

import numpy as np
import torch

from inputs import synth_inputs, shard   # tests/golden/inputs.py
from oracle import head_oracle as ho

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
HEAD_CASES = ["head_w1_full", "head_w1_s30", "head_w1_cosface", "head_w1_sampled", "head_w1_manypos",
              "head_w2_full", "head_w2_sampled", "head_w1_d512", "head_w1_filter", "head_w1_filter_wide", "head_w1_d128",
              "head_w2_d128"]


def load_case(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    cfg = {k[4:]: z[k].item() for k in z.files if k.startswith("cfg_")}
    cfg.setdefault("margin", "arcface")
    cfg.setdefault("sigma", 1.0)
    return cfg, z


def case_inputs(cfg):
    W, b = cfg["W"], cfg["b"]
    w_full, xs, ls = synth_inputs(cfg["C"], cfg["d"], b * W, cfg["steps"], sigma=cfg["sigma"])
    weights = []
    for r in range(W):
        nl, cs = shard(cfg["C"], r, W)
        weights.append(w_full[cs:cs + nl].clone())
    return weights, xs, ls


def case_margin(cfg):
    if cfg["margin"] == "combined_filter":      # CombinedMarginLoss(s, 1, m, 0, interclass_filtering_threshold)
        return ho.Margin(kind="arcface", s=cfg["s"], m=cfg["m"], filter_thr=cfg["filter_thr"])
    return ho.Margin(kind=cfg["margin"], s=cfg["s"], m=cfg["m"])


def case_perms(cfg, z, step):
    if cfg["sample_rate"] >= 1:
        return None
    return [torch.from_numpy(z[f"r{r}_perm_{step}"]) for r in range(cfg["W"])]


def cosine(a, b):
    a = torch.as_tensor(a, dtype=torch.float64).reshape(-1)
    b = torch.as_tensor(b, dtype=torch.float64).reshape(-1)
    return float((a @ b) / (a.norm() * b.norm()).clamp_min(1e-300))
