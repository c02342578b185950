"""The radix-select ALGORITHM of csrc/pfc_sample.cu restated in NumPy, stage by stage, against the oracle's
reference-pinned sampler (oracle/head_oracle.py::sample_indices, nets/PartialFC.py:108-118): order-preserving keys with
positives forced to 2.0, three MSB-first histogram passes (11 + 11 + 10 bits) whose pick is "the largest bin b >= 1
with an inclusive suffix count >= k_rem" (the CTA-wide scan of pick_parallel_kernel / sample_fused_kernel), and the
ordered compaction rule "position = (#keys > T before) + min(#keys == T before, need_eq)" of the one-launch sampler.
CPU only: it checks the design the CUDA kernels implement, the kernels themselves are checked on the GPU
(tools/gpu_probe.py::case_sample)."""
import numpy as np
import pytest
import torch

from oracle import head_oracle as ho


def _sortable(f32):
    b = f32.view(np.uint32).astype(np.uint64)
    return np.where(b >> 31, b ^ 0xFFFFFFFF, b ^ 0x80000000).astype(np.uint64)


def radix_select(perm, labels, num_sample):
    nl = len(perm)
    flags = np.zeros(nl, bool)
    flags[labels[labels >= 0]] = True
    key = np.where(flags, _sortable(np.array([2.0], np.float32))[0], _sortable(perm.astype(np.float32)))
    k_eff = min(max(int(flags.sum()), num_sample), nl)
    rem, prefix = k_eff, 0
    for p in range(3):
        bits, shift = (10 if p == 2 else 11), (21, 10, 0)[p]
        bins, hi_mask = 1 << bits, (0, 0xFFE00000, 0xFFFFFC00)[p]
        if rem == 0:
            prefix = 0xFFFFFFFF
            break
        sel = (key & hi_mask) == prefix
        hist = np.bincount(((key[sel] >> shift) & (bins - 1)).astype(np.int64), minlength=bins)
        rev = hist[::-1]
        incl = np.cumsum(rev)                                   # inclusive suffix counts, top bin first
        cand = np.nonzero((incl >= rem) & (np.arange(bins) < bins - 1))[0]
        r = int(cand.min()) if len(cand) else bins - 1          # atomicMin over the reversed positions; default: bin 0
        prefix |= (bins - 1 - r) << shift
        rem -= int(incl[r]) - int(rev[r])
    if k_eff == 0:
        return np.zeros(0, np.int64), np.where(labels >= 0, 0, -1)
    gt, eq = key > prefix, key == prefix
    gt_before, eq_before = np.cumsum(gt) - gt, np.cumsum(eq) - eq
    take = gt | (eq & (eq_before < rem))
    pos = gt_before + np.minimum(eq_before, rem)
    index = np.zeros(k_eff, np.int64)
    index[pos[take]] = np.nonzero(take)[0]
    slot = np.full(nl, -1)
    slot[take] = pos[take]
    return index, np.where(labels >= 0, slot[np.maximum(labels, 0)], -1)


CASES = [(400, 100, 32, 1), (45029, 4502, 1024, 2), (257489, 51497, 4096, 3), (64, 16, 32, 4), (5000, 0, 16, 5),
         (1000, 1000, 8, 6)] + [(int(np.random.default_rng(t).integers(50, 3000)),
                                 int(np.random.default_rng(t + 99).integers(0, 60)), 64, 100 + t) for t in range(24)]


@pytest.mark.parametrize("ties", [True, False])
def test_radix_select_design_matches_the_oracle(ties):
    for nl, ns, B, seed in CASES:
        g = torch.Generator().manual_seed(seed)
        perm = torch.rand(nl, generator=g)
        if ties:
            perm = torch.floor(perm * 4096) / 4096              # plenty of equal scores, also at the k-th value
        lab = torch.randint(-1, nl, (B,), generator=g)
        idx_ref, lab_ref = ho.sample_indices(perm, lab.long(), ns)
        idx, lab_new = radix_select(perm.numpy(), lab.numpy(), ns)
        assert np.array_equal(idx, idx_ref.numpy()), (nl, ns, B, seed)
        assert np.array_equal(lab_new, lab_ref.numpy()), (nl, ns, B, seed)
