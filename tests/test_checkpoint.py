"""Sharded head checkpoint helpers (face_recognition_pytorch_b200.checkpoint): layout, re-sharding arithmetic
(nets/PartialFC.py:57-62) and round trips.  CPU only; the GPU save -> load -> same-loss test is in test_gpu_head.py."""
import types

import pytest
import torch


@pytest.fixture(scope="module")
def pfc():
    import torch.distributed as dist
    if not dist.is_initialized():
        dist.init_process_group("gloo", init_method="tcp://127.0.0.1:29733", rank=0, world_size=1)
    import face_recognition_pytorch_b200 as m
    return m


def _fake_shards(pfc, C, d, W, with_mom=True):
    g = torch.Generator().manual_seed(5)
    full_w, full_m = torch.randn(C, d, generator=g), torch.randn(C, d, generator=g)
    out = []
    for r in range(W):
        nl, cs = pfc.shard_range(C, r, W)
        s = {"weight": full_w[cs:cs + nl].clone(),
             "meta": {"rank": r, "world_size": W, "num_local": nl, "class_start": cs, "num_classes": C, "step": 7,
                      "optimizer": "sgd"}}
        if with_mom:
            s["weight_mom"] = full_m[cs:cs + nl].clone()
        out.append(s)
    return full_w, full_m, out


@pytest.mark.parametrize("C,W,W2", [(1003, 3, 2), (1003, 2, 8), (64, 8, 1), (10, 4, 3)])
def test_reshard_round_trip(pfc, C, W, W2):
    full_w, full_m, shards = _fake_shards(pfc, C, 16, W)
    new = pfc.reshard(list(reversed(shards)), W2)          # any order in
    assert len(new) == W2
    assert torch.equal(torch.cat([s["weight"] for s in new]), full_w)
    assert torch.equal(torch.cat([s["weight_mom"] for s in new]), full_m)
    for r, s in enumerate(new):
        nl, cs = pfc.shard_range(C, r, W2)
        assert s["meta"]["rank"] == r and s["meta"]["world_size"] == W2
        assert s["meta"]["num_local"] == nl == s["weight"].shape[0] and s["meta"]["class_start"] == cs
        assert s["meta"]["step"] == 7
    back = pfc.reshard(new, W)
    for a, b in zip(back, shards):
        assert torch.equal(a["weight"], b["weight"]) and torch.equal(a["weight_mom"], b["weight_mom"])


def test_reshard_rejects_incomplete_or_foreign_partitions(pfc):
    _, _, shards = _fake_shards(pfc, 100, 8, 4)
    with pytest.raises(ValueError):
        pfc.reshard(shards[:3], 2)
    shards[1]["meta"]["class_start"] += 1
    with pytest.raises(ValueError):
        pfc.reshard(shards, 2)


def test_state_of_a_head_keeps_the_reference_layout(pfc):
    conf = types.SimpleNamespace(emd_size=16, sample_rate=1.0, mixed_precision=False, loss_s=64.0, loss_m=0.5)
    head = pfc.PartialFC(conf, 37)
    sd = pfc.head_shard_state(head)
    assert torch.equal(sd["weight"], head.state_dict()["weight"]) and tuple(sd["weight"].shape) == (37, 16)
    assert sd["meta"] == {"rank": 0, "world_size": 1, "num_local": 37, "class_start": 0, "num_classes": 37, "step": 0,
                          "optimizer": "sgd"}
    head2 = pfc.PartialFC(conf, 37)
    pfc.load_head_shard(head2, sd)
    assert torch.equal(head2.state_dict()["weight"], sd["weight"])
    with pytest.raises(ValueError):
        pfc.load_head_shard(pfc.PartialFC(conf, 38), sd)


def test_spill_layout_helpers_round_trip(pfc):
    """kernels.spill_from_rowmajor / spill_to_rowmajor: the class-blocked E'[n_pad/64][B][64] layout of include/pfc.h."""
    from face_recognition_pytorch_b200 import kernels as K
    B, n_pad = 5, 192
    M = torch.arange(B * n_pad, dtype=torch.float32).view(B, n_pad).to(torch.bfloat16)
    flat = K.spill_from_rowmajor(M)
    assert flat.shape == (B * n_pad,)
    # element (row i, class c) sits at ((c // 64) * B + i) * 64 + c % 64
    for i, c in [(0, 0), (3, 63), (4, 64), (2, 130), (4, 191)]:
        assert flat[((c // 64) * B + i) * 64 + c % 64] == M[i, c]
    assert torch.equal(K.spill_to_rowmajor(flat, B, n_pad), M)
