"""CPU-only: how far apart do *exact* restatements of the reference land in validation AUC after N training steps?
fp32 oracle vs fp64 oracle, same weights / batches / masks / recipe as tools/auc_check.py -- the floor any second implementation
of the reference can be held to -- plus two arms whose every matmul is an exact emulation of a tensor-core operand split
(tools/split_precision_sim.py): tf32x3 and f16x3, the default precision of the CUDA path.

    python tools/auc_sensitivity_cpu.py [steps] [arms, e.g. fp32,fp64,f16x3,tf32x3]

Measured (B = 1024, 20000 validation rows, |AUC - AUC of the fp64 oracle|; profiles/r2c_auc_sensitivity_cpu.txt):
    4 steps : fp32 1.4e-06   f16x3 3.1e-06   tf32x3 6.1e-06
   40 steps : fp32 8.4e-04   f16x3 9.4e-05   tf32x3 2.4e-04
The split-operand arms sit inside the spread of the two plain restatements (multi-step trajectories are chaotic at the 1e-4
level, DESIGN.md section 2, finding 1).
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import split_precision_sim as sim  # noqa: E402
from oracle import fibinet_numpy as orc, synth  # noqa: E402

steps, B, n_valid = int(sys.argv[1]) if len(sys.argv) > 1 else 40, 1024, 20000
arms = (sys.argv[2] if len(sys.argv) > 2 else "fp32,fp64,f16x3,tf32x3").split(",")
table = synth.make_item_mm_table(seed=11)
res = {}
vb, vy = synth.make_batch(seed=123456, batch=n_valid, id_dist="zipf", index_dtype=np.int64, mm_table=table, edge_cases=False)
for arm in arms:
    dt = np.float64 if arm == "fp64" else np.float32
    emulated = arm in ("f16x3", "tf32x3")
    P = synth.make_weights(seed=7)
    if dt == np.float64:
        P = {k: (v.astype(np.float64) if v.dtype == np.float32 else v) for k, v in P.items()}
    if emulated:      # every `@` of the oracle becomes the emulated product (operand representation error only)
        sim.SCHEME = arm
        P = {k: (np.array(v).view(sim.Q) if v.dtype.kind == "f" else np.array(v)) for k, v in P.items()}
    opt = orc.Adam(lr=1e-3, weight_decay=1e-5, dtype=dt)
    sched = orc.OneCycle(1e-2, 5 * steps)
    for s in range(steps):
        b, y = synth.make_batch(seed=9000 + s, batch=B, id_dist="zipf", index_dtype=np.float64, mm_table=table, edge_cases=False)
        if emulated:
            b["item_emb_d128"] = np.array(b["item_emb_d128"]).view(sim.Q)
        m = synth.make_dropout_masks(9000 + s, B)
        opt.lr, b1 = sched.at(s); opt.betas = (b1, 0.999)
        orc.train_step(P, opt, b, y, masks=m, dtype=dt)
        if emulated:      # the optimizer may hand back plain arrays
            P = {k: (np.asarray(v).view(sim.Q) if np.asarray(v).dtype.kind == "f" else v) for k, v in P.items()}
    Pv = {k: np.asarray(v) for k, v in P.items()}          # score every arm with the plain forward: only the weights differ
    p, _ = orc.forward(Pv, vb, train=False, dtype=dt)
    res[arm] = (orc.auc(vy, p), np.asarray(p, np.float64))
    print(f"{arm:7s} {steps} steps of B={B}: AUC {res[arm][0]:.6f}", flush=True)
ref = "fp64" if "fp64" in res else arms[0]
for arm in arms:
    if arm != ref:
        print(f"{arm:7s} vs {ref}: |AUC diff| {abs(res[arm][0] - res[ref][0]):.2e}; max |p diff| {np.abs(res[arm][1] - res[ref][1]).max():.2e}")
