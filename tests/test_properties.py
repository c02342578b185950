"""Property tests (hypothesis) for the integer / index parts of the path: shard arithmetic (nets/PartialFC.py:57-62),
negative sampling invariants (:108-121), label localisation (:188-193), pair-score binning (utils/eval.py:85-97) and
checkpoint re-sharding.  CPU only: they exercise the oracle and the host helpers the CUDA tests compare against."""
import numpy as np
import torch
from hypothesis import given, settings, strategies as st

from oracle import eval_oracle as eo
from oracle import head_oracle as ho


@settings(max_examples=200, deadline=None)
@given(C=st.integers(1, 5000), W=st.integers(1, 16))
def test_shards_partition_the_classes(C, W):
    import face_recognition_pytorch_b200 as pfc
    pos, sizes = 0, []
    for r in range(W):
        nl, cs = ho.shard_range(C, r, W)
        assert (nl, cs) == pfc.shard_range(C, r, W)
        assert cs == pos and nl >= 0
        pos += nl
        sizes.append(nl)
    assert pos == C and max(sizes) - min(sizes) <= 1
    assert sizes == sorted(sizes, reverse=True)          # the remainder goes to the lowest ranks


@settings(max_examples=150, deadline=None)
@given(nl=st.integers(8, 400), B=st.integers(1, 64), rate=st.floats(0.05, 0.95), seed=st.integers(0, 2 ** 31 - 1),
       foreign=st.floats(0.0, 0.9))
def test_sampling_invariants(nl, B, rate, seed, foreign):
    g = torch.Generator().manual_seed(seed)
    labels = torch.randint(0, nl, (B,), generator=g)
    labels[torch.rand(B, generator=g) < foreign] = -1            # rows whose class lives on another rank
    perm = torch.rand(nl, generator=g)
    k = ho.num_sample(rate, nl)
    index, remapped = ho.sample_indices(perm, labels.clone(), k)
    pos = torch.unique(labels[labels >= 0])
    assert index.numel() == max(k, pos.numel())                  # :114-115 -- n is data-dependent
    assert bool((index[1:] > index[:-1]).all())                  # strictly ascending
    assert set(pos.tolist()) <= set(index.tolist())              # every positive class is kept
    own = labels >= 0
    assert torch.equal(index[remapped[own]], labels[own])        # labels point at positions of the index list
    assert bool((remapped[~own] == -1).all())
    if k > pos.numel():                                          # the rest are the top-scoring negatives
        neg = index[~torch.isin(index, pos)]
        others = torch.ones(nl, dtype=torch.bool)
        others[index] = False
        if others.any() and neg.numel():
            assert float(perm[neg].min()) >= float(perm[others].max())


@settings(max_examples=100, deadline=None)
@given(C=st.integers(4, 3000), W=st.integers(1, 8), B=st.integers(1, 64), seed=st.integers(0, 2 ** 31 - 1))
def test_every_label_is_local_on_exactly_one_rank(C, W, B, seed):
    labels = torch.randint(0, C, (B,), generator=torch.Generator().manual_seed(seed))
    owners = torch.zeros(B, dtype=torch.int64)
    for r in range(W):
        nl, cs = ho.shard_range(C, r, W)
        loc = ho.localize_labels(labels, cs, nl)
        own = loc >= 0
        owners += own.long()
        assert torch.equal(loc[own] + cs, labels[own]) and bool((loc[own] < nl).all())
    assert bool((owners == 1).all())


@settings(max_examples=40, deadline=None)
@given(n=st.integers(1, 200), d=st.sampled_from([8, 64, 512]), seed=st.integers(0, 2 ** 31 - 1))
def test_pair_score_binning(n, d, seed):
    rng = np.random.default_rng(seed)
    e1 = rng.standard_normal((n, d)).astype(np.float32)
    e2 = rng.standard_normal((n, d)).astype(np.float32)
    e1 /= np.linalg.norm(e1, axis=1, keepdims=True)
    e2 /= np.linalg.norm(e2, axis=1, keepdims=True)
    lab = rng.random(n) < 0.5
    hg, hi, sc = eo.pair_score(e1, e2, lab)
    assert hg.shape == hi.shape == (100001,)
    assert hg.sum() == lab.sum() and hi.sum() == (~lab).sum()
    assert np.all(sc > -1e-6) and np.all(sc < 1 + 1e-6)          # 1 - |a-b|^2/4 on unit vectors
    idx = (99999 * sc).astype(np.int64)
    assert np.array_equal(np.bincount(idx[lab], minlength=100001), hg.astype(np.int64))
    assert np.array_equal(np.bincount(idx[~lab], minlength=100001), hi.astype(np.int64))
    cos = np.sum(e1.astype(np.float64) * e2.astype(np.float64), axis=1)
    np.testing.assert_allclose(sc, (1 + cos) / 2, atol=1e-6)


@settings(max_examples=60, deadline=None)
@given(C=st.integers(1, 700), W=st.integers(1, 9), W2=st.integers(1, 9), seed=st.integers(0, 2 ** 31 - 1))
def test_reshard_any_to_any(C, W, W2, seed):
    import face_recognition_pytorch_b200 as pfc
    full = torch.randn(C, 4, generator=torch.Generator().manual_seed(seed))
    shards = []
    for r in range(W):
        nl, cs = pfc.shard_range(C, r, W)
        shards.append({"weight": full[cs:cs + nl].clone(),
                       "meta": {"rank": r, "world_size": W, "num_local": nl, "class_start": cs, "num_classes": C,
                                "step": 0, "optimizer": "sgd"}})
    new = pfc.reshard(shards, W2)
    assert torch.equal(torch.cat([s["weight"] for s in new]), full)
    assert [s["meta"]["class_start"] for s in new] == [pfc.shard_range(C, r, W2)[1] for r in range(W2)]
