"""GPU diagnostic: per-tensor gradient / weight deviation of the CUDA path from the numpy oracle, step by step."""
import sys, os
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import fibinet_numpy as orc, synth
from gpu_common import make_model, to_dev, named_grads
from ctr_recommendation_b200 import FusedAdam, clip_grad_norm_

fused = "--torch-adam" not in sys.argv
B, steps, total = 256, 3, 40
P = synth.make_weights(7)
oopt = orc.Adam(lr=1e-3, weight_decay=1e-5); osched = orc.OneCycle(1e-2, total)
model = make_model(train=True)
opt = FusedAdam(model, lr=1e-3, weight_decay=1e-5) if fused else torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-5)
sched = torch.optim.lr_scheduler.OneCycleLR(opt, max_lr=1e-2, total_steps=total, pct_start=0.3, div_factor=25.0, final_div_factor=1000.0)
hist = {}
for s in range(steps):
    lr, b1 = osched.at(s); oopt.lr, oopt.betas = lr, (b1, 0.999)
    batch, labels = synth.make_batch(seed=100 + s, batch=B, index_dtype=np.float64)
    m1, m2 = synth.make_dropout_masks(50 + s, B)
    if "--golden" in sys.argv:
        from helpers import unpack_mask
        gold = dict(np.load(os.path.join(ROOT, "tests", "golden", "fibinet_golden.npz")))
        m1 = unpack_mask(gold[f"train_u/step{s}/mask1"], (B, 512)); m2 = unpack_mask(gold[f"train_u/step{s}/mask2"], (B, 256))
    model._test_masks = (torch.from_numpy(m1), torch.from_numpy(m2))
    opt.zero_grad()
    y = model(to_dev(batch)); loss = torch.nn.BCELoss()(y, torch.from_numpy(labels).cuda()); loss.backward()
    G = named_grads(model)
    if fused:
        G["item_emb.weight"] = (model._item_grad * (model._row_touched > 0).unsqueeze(1)).cpu().numpy()
    prob, cache = orc.forward(P, batch, train=True, masks=(m1, m2)); oloss, dprob = orc.bce_loss(prob, labels)
    OG = orc.backward(P, cache, dprob)
    if "--golden" in sys.argv:
        gp = gold[f"train_u/step{s}/prob"]
        print(f"   vs golden: gpu {np.abs(y.detach().cpu().numpy()-gp).max():.2e} oracle {np.abs(prob-gp).max():.2e}; worst idx {np.abs(y.detach().cpu().numpy()-gp).argmax()}")
    print(f"step {s}: prob err {np.abs(y.detach().cpu().numpy()-prob).max():.2e} lr {lr:.2e} (gpu lr {opt.param_groups[0]['lr']:.2e}, b1 {opt.param_groups[0]['betas'][0]:.4f} vs {b1:.4f})")
    for k in OG:
        d = np.abs(G[k].astype(np.float64) - OG[k])
        # elementwise relative error where the gradient matters for Adam (|g| > 1e-7)
        big = np.abs(OG[k]) > 1e-7
        rel = (d[big] / np.abs(OG[k][big])).max() if big.any() else 0
        print(f"   grad {k:28s} max|g| {np.abs(OG[k]).max():.2e} abs err {d.max():.2e} max elem-rel(|g|>1e-7) {rel:.2e}")
    if fused: clip_grad_norm_(model, 10.0)
    else: torch.nn.utils.clip_grad_norm_(model.parameters(), 10.0)
    opt.step(); sched.step()
    orc.clip_grad_norm_(OG, 10.0); oopt.step(P, OG)
    sd = {k: v.detach().cpu().numpy() for k, v in model.state_dict().items()}
    for k in OG:
        d = np.abs(sd[k].astype(np.float64) - P[k])
        print(f"   wgt  {k:28s} max err {d.max():.2e} = {d.max()/lr:.3f} lr ; mean {d.mean():.2e} ; frac>0.01lr {(d>0.01*lr).mean():.2e}")
    for k in ("item_emb.weight", "mlp.0.weight", "mm_proj.0.weight"):
        d = np.abs(sd[k].astype(np.float64) - P[k])
        idx = np.unravel_index(d.argmax(), d.shape)
        print(f"   worst {k} idx {idx}: w gpu {sd[k][idx]:.6e} oracle {P[k][idx]:.6e}; this-step grad gpu {G[k][idx]:.4e} oracle(clipped) {OG[k][idx]:.4e}"
              f"; row grad absmax gpu {np.abs(G[k][idx[0]]).max():.3e} oracle {np.abs(OG[k][idx[0]]).max():.3e}")
        if k == "item_emb.weight":
            r = idx[0]
            occ_item = np.flatnonzero(batch["item_id"].astype(np.int64) == r)
            occ_seq = np.argwhere(batch["item_seq"] == r)
            print(f"      row {r}: occurrences as target {occ_item.tolist()} in history {occ_seq.tolist()[:6]}")
            hist.setdefault(r, [])
        for r in list(hist):
            hist[r].append((s, float(G["item_emb.weight"][r, idx[1] if k == "item_emb.weight" else 0]), float(OG["item_emb.weight"][r, idx[1] if k == "item_emb.weight" else 0])))
    for k in ("mlp.1.running_mean", "mlp.1.running_var", "mlp.5.running_mean", "mlp.5.running_var"):
        print(f"   buf  {k:28s} max err {np.abs(sd[k]-P[k]).max():.2e}")
print(hist)
