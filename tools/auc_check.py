"""Validation-AUC parity on a fixed synthetic set (north_star: within 1e-4).

Trains the CUDA path and the numpy oracle from the same weights on the same batches / dropout masks with the reference's
recipe (Adam wd 1e-5, OneCycleLR, clip 10), then scores a fixed validation set with BOTH weight sets and compares AUC
(src/utils.py compute_auc semantics).  Usage: python tools/auc_check.py [steps] [batch] [precision]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from ctr_recommendation_b200 import FusedAdam, build_model, clip_grad_norm_  # noqa: E402
from oracle import fibinet_numpy as orc  # noqa: E402
from oracle import synth  # noqa: E402


def run(steps=40, B=1024, precision="tf32x3", n_valid=20000, verbose=True):
    table = synth.make_item_mm_table(seed=11)
    W = synth.make_weights(seed=7)
    model = build_model({"precision": precision}, {"embedding_dim": 128})
    model.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in W.items()})
    model = model.cuda().train()
    opt = FusedAdam(model, lr=1e-3, weight_decay=1e-5)
    total = 5 * steps
    sched = torch.optim.lr_scheduler.OneCycleLR(opt, max_lr=1e-2, total_steps=total, pct_start=0.3, div_factor=25.0, final_div_factor=1000.0)
    oopt = orc.Adam(lr=1e-3, weight_decay=1e-5)
    osched = orc.OneCycle(1e-2, total)
    P = {k: np.array(v) for k, v in W.items()}
    for s in range(steps):
        b, y = synth.make_batch(seed=9000 + s, batch=B, id_dist="zipf", index_dtype=np.float64, mm_table=table, edge_cases=False)
        m1, m2 = synth.make_dropout_masks(9000 + s, B)
        model._test_masks = (torch.from_numpy(m1), torch.from_numpy(m2))
        opt.zero_grad()
        out = model({k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in b.items()})
        torch.nn.BCELoss()(out, torch.from_numpy(y).cuda()).backward()
        clip_grad_norm_(model, 10.0)
        opt.step()
        sched.step()
        lr, b1 = osched.at(s)
        oopt.lr, oopt.betas = lr, (b1, 0.999)
        orc.train_step(P, oopt, b, y, masks=(m1, m2))
    vb, vy = synth.make_batch(seed=123456, batch=n_valid, id_dist="zipf", index_dtype=np.int64, mm_table=table, edge_cases=False)
    dev = {k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in vb.items()}
    model.eval()
    model._test_masks = None
    with torch.no_grad():
        p_gpu = model(dev).cpu().numpy()
    other = build_model({"precision": precision}, {"embedding_dim": 128})
    other.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in P.items()})
    other = other.cuda().eval()
    with torch.no_grad():
        p_orc = other(dev).cpu().numpy()
    a_gpu, a_orc = orc.auc(vy, p_gpu), orc.auc(vy, p_orc)
    a0 = orc.auc(vy[:2000], orc.forward(W, {k: v[:2000] for k, v in vb.items()}, train=False)[0]) if verbose else None
    if verbose:
        print(f"precision {precision}: {steps} steps of B={B}; AUC cuda-trained {a_gpu:.6f} oracle-trained {a_orc:.6f} "
              f"|diff| {abs(a_gpu - a_orc):.2e}; max |p diff| {np.abs(p_gpu - p_orc).max():.2e}; untrained AUC (2k rows) {a0:.4f}")
    return a_gpu, a_orc


if __name__ == "__main__":
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
    prec = sys.argv[3] if len(sys.argv) > 3 else "tf32x3"
    run(steps, B, prec)
