"""CPU: the oracle's optimizer restatements against torch.optim (the de-facto pinned version of the un-vendored dependency the
reference's arithmetic lives in).  Adam is pinned through the golden vectors (tests/test_oracle_golden.py); Adagrad -- an extension
named by BASELINE north_star (2), which the reference never builds -- is pinned here."""
import numpy as np
import pytest
import torch

from oracle import fibinet_numpy as orc


@pytest.mark.parametrize("wd,lr_decay,init", [(0.0, 0.0, 0.0), (1e-5, 0.0, 0.0), (1e-3, 0.05, 0.1)])
def test_oracle_adagrad_matches_torch(wd, lr_decay, init):
    g = torch.Generator().manual_seed(3)
    shapes = {"a": (37, 128), "b": (512,), "c": (6, 3)}
    P = {k: torch.randn(*s, generator=g) for k, s in shapes.items()}
    tp = {k: v.clone().requires_grad_(True) for k, v in P.items()}
    opt = torch.optim.Adagrad(list(tp.values()), lr=1e-2, lr_decay=lr_decay, weight_decay=wd, initial_accumulator_value=init, eps=1e-10)
    o = orc.Adagrad(lr=1e-2, lr_decay=lr_decay, weight_decay=wd, initial_accumulator_value=init, eps=1e-10)
    Pn = {k: v.numpy().copy() for k, v in P.items()}
    for step in range(5):
        G = {k: torch.randn(*s, generator=g) * (10.0 ** -step) for k, s in shapes.items()}
        G["a"][::3] = 0.0                                   # untouched rows: identity when wd == 0
        for k in tp:
            tp[k].grad = G[k].clone()
        opt.step()
        o.step(Pn, {k: v.numpy() for k, v in G.items()})
        for k in tp:
            d = np.abs(tp[k].detach().numpy().astype(np.float64) - Pn[k]).max()
            assert d <= 2e-7 * max(1.0, np.abs(Pn[k]).max()), (step, k, d)
            s = opt.state[tp[k]]["sum"].numpy()
            assert np.abs(s - o.state[k]["sum"]).max() <= 1e-6 * max(np.abs(s).max(), 1e-30)
    if wd == 0.0:
        assert np.array_equal(Pn["a"][::3], P["a"].numpy()[::3])
