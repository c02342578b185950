"""The radix-select ALGORITHM of csrc/pfc_sample.cu restated in NumPy, stage by stage, against the oracle's
reference-pinned sampler (oracle/head_oracle.py::sample_indices, nets/PartialFC.py:108-118): order-preserving keys with
positives forced to 2.0, three MSB-first histogram passes (11 + 11 + 10 bits) whose pick is "the largest bin b >= 1
with an inclusive suffix count >= k_rem" (the CTA-wide scan of merge_and_pick / big::pick_digit), and the ordered
compaction rule "position = (#keys > T before) + min(#keys == T before, need_eq)".  A second restatement follows the
one-launch cluster kernel's decomposition: contiguous slices per CTA (multiples of 1024 slots), per-CTA histograms summed
bin by bin, the pick as the ONE reversed bin where the running count crosses k_rem, per-warp sub-ranges of the slice with
their (> T, == T) counts, and positions from the CTA prefix + warp prefix + ballot rank inside each group of 32 slots.
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


def radix_select_cluster(perm, labels, num_sample, cl):
    """sample_cluster_kernel<cl>: same result through the per-CTA / per-warp bookkeeping."""
    nl = len(perm)
    S = -(-(-(-nl // cl)) // 1024) * 1024
    two = int(_sortable(np.array([2.0], np.float32))[0])
    keys, pos_bit = [], []
    for c in range(cl):                                           # stage + bitmap per CTA
        lo, hi = min(c * S, nl), min(c * S + S, nl)
        k = _sortable(perm[lo:hi].astype(np.float32)).astype(np.int64)
        bit = np.zeros(hi - lo, bool)
        own = labels[(labels >= lo) & (labels < hi)] - lo
        bit[own] = True
        k[own] = two
        keys.append(k)
        pos_bit.append(bit)
    n_pos = sum(int(b.sum()) for b in pos_bit)
    k_eff = min(max(n_pos, num_sample), nl)
    rem, prefix = k_eff, 0
    for p in range(3):
        bits, shift = (10 if p == 2 else 11), (21, 10, 0)[p]
        bins, hi_mask = 1 << bits, (0, 0xFFE00000, 0xFFFFFC00)[p]
        merged = np.zeros(bins, np.int64)
        for c in range(cl):                                       # owner-reduce: plain sums per bin
            sel = (keys[c] & hi_mask) == prefix
            merged += np.bincount((keys[c][sel] >> shift) & (bins - 1), minlength=bins)
        if rem == 0:
            prefix = 0xFFFFFFFF
            continue
        rev = merged[::-1]
        run = np.cumsum(rev)
        before = run - rev
        r_idx = np.arange(bins)
        mine = ((r_idx < bins - 1) & (before < rem) & (run >= rem)) | ((r_idx == bins - 1) & (before < rem))
        assert mine.sum() == 1                                    # exactly one thread writes the pick
        r = int(np.nonzero(mine)[0][0])
        prefix |= (bins - 1 - r) << shift
        rem -= int(run[r]) - int(rev[r])
    if k_eff == 0:
        return np.zeros(0, np.int64), np.where(labels >= 0, 0, -1)
    T, need_eq, wlen = prefix, rem, S // 32
    n_gt = [int((k > T).sum()) for k in keys]
    n_eq = [int((k == T).sum()) for k in keys]
    index = np.full(k_eff, -1, np.int64)
    slot = np.full(nl, -1)
    for c in range(cl):
        lo, ln = min(c * S, nl), len(keys[c])
        wg = [int((keys[c][w * wlen:(w + 1) * wlen] > T).sum()) for w in range(32)]
        we = [int((keys[c][w * wlen:(w + 1) * wlen] == T).sum()) for w in range(32)]
        for w in range(32):
            gt_before, eq_before = sum(n_gt[:c]) + sum(wg[:w]), sum(n_eq[:c]) + sum(we[:w])
            eq_run, sel_run = eq_before, gt_before + min(eq_before, need_eq)
            for j in range(w * wlen, min((w + 1) * wlen, ln), 32):        # one ballot group
                k = keys[c][j:j + 32]
                eq = k == T
                eq_rank = eq_run + np.cumsum(eq) - eq
                sel = (k > T) | (eq & (eq_rank < need_eq))
                p_ = sel_run + np.cumsum(sel) - sel
                index[p_[sel]] = lo + j + np.nonzero(sel)[0]
                bits_ = pos_bit[c][j:j + 32] & sel
                slot[lo + j + np.nonzero(bits_)[0]] = p_[bits_]
                eq_run += int(eq.sum())
                sel_run += int(sel.sum())
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
        for cl in (8, 16):
            idx, lab_new = radix_select_cluster(perm.numpy(), lab.numpy(), ns, cl)
            assert np.array_equal(idx, idx_ref.numpy()), (nl, ns, B, seed, cl)
            assert np.array_equal(lab_new, lab_ref.numpy()), (nl, ns, B, seed, cl)
