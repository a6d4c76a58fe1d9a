"""CPU: the F-field FiBiNET oracle (oracle/fibinet_general.py, the target of BASELINE config 5's 40-field model) against a torch
model assembled from the REFERENCE'S OWN SENetLayer / BilinearInteraction classes + torch autograd, in float64."""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import fibinet_general as gen

REF = "/root/reference/src"


def _ref_classes():
    if not os.path.exists(os.path.join(REF, "model_fibinet.py")):
        pytest.skip("reference not present on this box")
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_model_fibinet", os.path.join(REF, "model_fibinet.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


@pytest.mark.parametrize("F,D,bilinear_type,dropout", [(40, 8, "all", 0.0), (40, 8, "each", 0.0), (6, 16, "all", 0.25), (9, 4, "each", 0.25)])
def test_general_oracle_matches_reference_blocks(F, D, bilinear_type, dropout):
    ref = _ref_classes()
    torch.manual_seed(0)
    B, vocab, hidden = 24, 50, (32, 16)
    P = gen.make_params(F, D, vocab, hidden, bilinear_type, reduction_ratio=2, seed=3)
    rng = np.random.default_rng(5)
    ids = rng.integers(0, vocab, (B, F))
    labels = rng.integers(0, 2, B).astype(np.float64)
    masks = [(rng.random((B, h)) >= dropout).astype(np.float64) for h in hidden] if dropout > 0 else None
    prob, cache = gen.forward(P, ids, masks=masks, dropout_p=dropout)
    dprob = (prob - labels) / np.maximum(prob * (1 - prob), 1e-12) / B          # d mean-BCE / d prob
    G = gen.backward(P, cache, dprob)

    t = lambda a: torch.tensor(np.asarray(a), dtype=torch.float64, requires_grad=True)
    tables = [t(x) for x in P["tables"]]
    senet = ref.SENetLayer(F, reduction_ratio=2).double()
    bil = ref.BilinearInteraction(D, F, bilinear_type=bilinear_type).double()
    with torch.no_grad():
        senet.excitation[0].weight.copy_(torch.from_numpy(P["se_w1"])); senet.excitation[0].bias.copy_(torch.from_numpy(P["se_b1"]))
        senet.excitation[2].weight.copy_(torch.from_numpy(P["se_w2"])); senet.excitation[2].bias.copy_(torch.from_numpy(P["se_b2"]))
        if bilinear_type == "all":
            bil.W.copy_(torch.from_numpy(P["bil_w"][0]))
        else:
            for w, src in zip(bil.W_list, P["bil_w"]):
                w.copy_(torch.from_numpy(src))
    tid = torch.from_numpy(ids)
    X = torch.stack([tables[f][tid[:, f]] for f in range(F)], dim=1)
    V = senet(X)
    Pm = bil(V)
    a = torch.cat([V.flatten(1), Pm.flatten(1)], dim=1)
    lin = []
    for i, h in enumerate(hidden):
        w, b, g, be = t(P[f"w{i}"]), t(P[f"b{i}"]), t(P[f"bn_g{i}"]), t(P[f"bn_b{i}"])
        lin.append((w, b, g, be))
        hpre = a @ w.T + b
        a = torch.relu(torch.nn.functional.batch_norm(hpre, None, None, g, be, training=True, eps=1e-5))
        if dropout > 0:
            a = a * torch.from_numpy(masks[i]) / (1 - dropout)
    wo, bo = t(P["w_out"]), t(P["b_out"])
    p_t = torch.sigmoid((a @ wo.T + bo)[:, 0])
    assert np.abs(p_t.detach().numpy() - prob).max() <= 1e-12
    loss = torch.nn.BCELoss()(p_t, torch.from_numpy(labels))
    loss.backward()
    close = lambda a, b, name: np.testing.assert_allclose(a, b.detach().numpy() if hasattr(b, "detach") else b, rtol=1e-8, atol=1e-11,
                                                          err_msg=name)
    for f in range(F):
        close(G["tables"][f], tables[f].grad, f"table {f}")
    close(G["se_w1"], senet.excitation[0].weight.grad, "se_w1"); close(G["se_b1"], senet.excitation[0].bias.grad, "se_b1")
    close(G["se_w2"], senet.excitation[2].weight.grad, "se_w2"); close(G["se_b2"], senet.excitation[2].bias.grad, "se_b2")
    if bilinear_type == "all":
        close(G["bil_w"][0], bil.W.grad, "bil_w")
    else:
        for k, w in enumerate(bil.W_list):
            if w.grad is not None:
                close(G["bil_w"][k], w.grad, f"bil_w[{k}]")
    for i, (w, b, g, be) in enumerate(lin):
        close(G[f"w{i}"], w.grad, f"w{i}"); close(G[f"bn_g{i}"], g.grad, f"bn_g{i}"); close(G[f"bn_b{i}"], be.grad, f"bn_b{i}")
        assert np.abs(b.grad.numpy()).max() <= 1e-12 and np.abs(G[f"b{i}"]).max() <= 1e-12     # a bias before BatchNorm: zero gradient
    close(G["w_out"], wo.grad, "w_out"); close(G["b_out"], bo.grad, "b_out")
