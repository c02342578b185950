import sys, os, types
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import numpy as np, torch, torch.distributed as dist
from helpers import load_case, case_inputs, case_margin, case_perms
from oracle import head_oracle as ho
dist.init_process_group("nccl", init_method="tcp://127.0.0.1:29712", rank=0, world_size=1)
import face_recognition_pytorch_b200 as pfc
name = sys.argv[1] if len(sys.argv) > 1 else "head_w1_sampled"
cfg, z = load_case(name)
weights, xs, ls = case_inputs(cfg)
conf = types.SimpleNamespace(emd_size=cfg["d"], sample_rate=cfg["sample_rate"], mixed_precision=False, loss_s=cfg["s"], loss_m=cfg["m"], fused_optimizer=True)
head = pfc.PartialFC(conf, cfg["C"])
head.load_state_dict({"weight": weights[0].clone()})
head = head.train().cuda()
opt = torch.optim.SGD([{"params": [torch.nn.Parameter(torch.zeros(1, device="cuda"))]}, {"params": head.parameters()}], lr=cfg["lr"], momentum=cfg["momentum"], weight_decay=cfg["wd"])
orc = ho.PartialFCOracle(weights, cfg["C"], case_margin(cfg), cfg["sample_rate"], cfg["lr"], cfg["momentum"], cfg["wd"])
for s in range(cfg["steps"]):
    perms = case_perms(cfg, z, s)
    res = orc.step([xs[s]], [ls[s]], perms)
    x = xs[s].clone().cuda().requires_grad_(True)
    perm = perms[0].cuda() if perms is not None and perms[0].numel() else None
    loss = head(x, ls[s].clone().cuda(), opt, perm=perm)
    ws = head._ws
    n = head._n
    idx = head.weight_index.cpu()
    print(f"step {s}: n={n} loss={float(loss):.6f} ref={float(z[f'r0_loss_{s}']):.6f} oracle={float(res.loss):.6f}")
    print("  index equal:", torch.equal(idx, res.index[0]), " labels equal:", torch.equal(ws.labels_act.cpu().long(), res.labels_local[0]))
    w_or = orc.weight[0][res.index[0]]
    print("  w_act max diff vs oracle-gathered:", float((head.weight_activated.data.cpu().double() - w_or).abs().max()))
    wn_or, _ = ho.normalize_rows(w_or)
    print("  wn max diff:", float((ws.wn[:n].float().cpu().double() - wn_or).abs().max()))
    st = ws.stats.cpu().double()
    f = ho.rank_logits(ho.normalize_rows(xs[s].double())[0], w_or, res.labels_local[0], case_margin(cfg))
    k1 = cfg["s"] * 1.4426950408889634
    e = torch.exp2(f.z / cfg["s"] * k1 - (k1 - 64))
    rows = torch.nonzero(res.labels_local[0] >= 0).flatten(); cols = res.labels_local[0][rows]
    te = torch.zeros(len(st)); te = te.double(); te[rows] = e[rows, cols]
    e2 = e.clone(); e2[rows, cols] = 0
    print("  stats others rel err:", float(((st[:, 0] - e2.sum(1)).abs() / e2.sum(1)).max()), " tgt_e rel err:", float(((st[rows, 1] - te[rows]).abs() / te[rows]).max()))
    print("  per-row loss diff max:", float((-(torch.log((st[:,1]/(st[:,0]+st[:,1])).clamp_min(1e-30))) + torch.log((te/(te+e2.sum(1))).clamp_min(1e-30))).abs().max()))
    loss.backward()
    opt.step()
    torch.cuda.synchronize()
    # after fused update compare activated rows with the oracle's pending update
    idx_p, w_new, m_new = orc.pending[0]
    print("  post-step w_act diff:", float((head.weight_activated.data.cpu().double() - w_new).abs().max()), " mom diff:", float((head.weight_activated_mom.cpu().double() - m_new).abs().max()), " |update|max:", float((w_new - w_or).abs().max()))
