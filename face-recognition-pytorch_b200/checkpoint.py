"""Sharded head checkpoints.

The reference saves only the encoder (utils/trainer.py:107-115); its head exposes the per-rank layout
`PartialFC.state_dict() == {"weight": [num_local, d]}` (nets/PartialFC.py:210-222) but nothing ever writes it.  These
helpers keep exactly that layout -- one file per rank whose "weight" entry is what `load_state_dict` takes -- and add,
under separate keys, the optimizer state rows (momentum, or Adam's two moments) and the shard geometry, so that a run
can resume and a checkpoint written by W ranks can be re-split for W' ranks (class shards are contiguous and in rank
order, nets/PartialFC.py:57-62).
"""
import torch

from .partial_fc import shard_range


def head_shard_state(head):
    """Per-rank checkpoint dict: reference-compatible "weight" + optimizer-state rows + shard geometry."""
    if head.sample_rate < 1:
        head.update()         # scatter the last step's sampled rows back first (the reference leaves them stale, :210-222)
    sd = {"weight": head.state_dict()["weight"].detach().clone().cpu()}
    for nm in head._state_names:
        if head.sample_rate < 1:
            t = getattr(head, "weight_" + nm)
        elif head.fused_optimizer and head._fused_state is not None:
            st = head._fused_state if isinstance(head._fused_state, tuple) else (head._fused_state,)
            t = st[head._state_names.index(nm)]
        else:
            t = None
        if t is not None:
            sd["weight_" + nm] = t.detach().clone().cpu()
    sd["meta"] = {"rank": head.rank, "world_size": head.world_size, "num_local": head.num_local,
                  "class_start": head.class_start, "num_classes": _num_classes(head), "step": int(head.step),
                  "optimizer": head._optimizer_kind}
    return sd


def _num_classes(head):
    return int(head._num_classes)


def load_head_shard(head, sd):
    """Inverse of head_shard_state for the SAME world size: weight through the reference's load_state_dict
    (nets/PartialFC.py:224-232, which zeroes the optimizer state), then the optimizer-state rows if present."""
    if tuple(sd["weight"].shape) != (head.num_local, head.embedding_size):
        raise ValueError(f"shard shape {tuple(sd['weight'].shape)} does not match this rank "
                         f"({head.num_local}, {head.embedding_size}); use reshard() first")
    head.load_state_dict({"weight": sd["weight"]})
    dev = head.weight_activated.device if head.sample_rate == 1 else head.weight.device
    states = [sd.get("weight_" + nm) for nm in head._state_names]
    if all(s is not None for s in states):
        if head.sample_rate < 1:
            for nm, s in zip(head._state_names, states):
                getattr(head, "weight_" + nm).copy_(s.to(dev))
        elif head.fused_optimizer:
            ts = tuple(s.to(dev).clone() for s in states)
            head._fused_state = ts if len(ts) > 1 else ts[0]
            if head._state_names == ["mom"]:
                head.weight_activated_mom = head._fused_state
    head.step = int(sd.get("meta", {}).get("step", head.step))
    # the next forward must not scatter the (zeroed) activated rows over the freshly loaded shard
    head.init_weight_update = True
    return head


def reshard(shards, new_world_size):
    """Re-split the per-rank dicts of one checkpoint (any order; `meta.rank` sorts them) for `new_world_size` ranks.
    Every tensor entry whose first dimension is the class dimension is concatenated in rank order and cut again with
    the reference's shard arithmetic (nets/PartialFC.py:57-62)."""
    shards = sorted(shards, key=lambda s: s["meta"]["rank"])
    old_w = shards[0]["meta"]["world_size"]
    if len(shards) != old_w or [s["meta"]["rank"] for s in shards] != list(range(old_w)):
        raise ValueError("reshard needs exactly one shard per rank of the writing job")
    keys = [k for k, v in shards[0].items() if torch.is_tensor(v)]
    full = {k: torch.cat([s[k] for s in shards], dim=0) for k in keys}
    num_classes = full["weight"].shape[0]
    for s in shards:
        nl, cs = shard_range(num_classes, s["meta"]["rank"], old_w)
        if s["weight"].shape[0] != nl or s["meta"]["class_start"] != cs:
            raise ValueError("shards are not a contiguous class partition in rank order")
    out = []
    for r in range(new_world_size):
        nl, cs = shard_range(num_classes, r, new_world_size)
        d = {k: full[k][cs:cs + nl].clone() for k in keys}
        d["meta"] = dict(shards[0]["meta"], rank=r, world_size=new_world_size, num_local=nl, class_start=cs,
                         num_classes=num_classes)
        out.append(d)
    return out
