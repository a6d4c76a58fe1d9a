"""CPU-only: how far apart do two *exact* restatements of the reference land in validation AUC after N training steps?
fp32 oracle vs fp64 oracle, same weights / batches / masks / recipe as tools/auc_check.py.  This is the floor any second
implementation of the reference can be held to."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import fibinet_numpy as orc, synth

steps, B, n_valid = int(sys.argv[1]) if len(sys.argv) > 1 else 40, 1024, 20000
table = synth.make_item_mm_table(seed=11)
res = {}
vb, vy = synth.make_batch(seed=123456, batch=n_valid, id_dist="zipf", index_dtype=np.int64, mm_table=table, edge_cases=False)
for dt in (np.float32, np.float64):
    P = synth.make_weights(seed=7)
    if dt == np.float64:
        P = {k: (v.astype(np.float64) if v.dtype == np.float32 else v) for k, v in P.items()}
    opt = orc.Adam(lr=1e-3, weight_decay=1e-5, dtype=dt)
    sched = orc.OneCycle(1e-2, 5 * steps)
    for s in range(steps):
        b, y = synth.make_batch(seed=9000 + s, batch=B, id_dist="zipf", index_dtype=np.float64, mm_table=table, edge_cases=False)
        m = synth.make_dropout_masks(9000 + s, B)
        opt.lr, b1 = sched.at(s); opt.betas = (b1, 0.999)
        orc.train_step(P, opt, b, y, masks=m, dtype=dt)
    p, _ = orc.forward(P, vb, train=False, dtype=dt)
    res[dt.__name__] = (orc.auc(vy, p), p)
a32, a64 = res["float32"][0], res["float64"][0]
print(f"{steps} steps of B={B}: AUC fp32-oracle {a32:.6f} fp64-oracle {a64:.6f} |diff| {abs(a32-a64):.2e}; max |p diff| "
      f"{np.abs(res['float32'][1] - res['float64'][1]).max():.2e}")
