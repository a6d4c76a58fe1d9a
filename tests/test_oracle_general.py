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


def _bag_spec():
    """user | likes, views (one table) | item_id (padding row 0) | item_seq (bag of 6 into item_id's table) | item_tags (bag of 3)"""
    spec = [dict(table=0, col=0, bag=1, pad=-1), dict(table=1, col=1, bag=1, pad=-1), dict(table=1, col=2, bag=1, pad=-1),
            dict(table=2, col=3, bag=1, pad=0), dict(table=2, col=4, bag=6, pad=0), dict(table=3, col=10, bag=3, pad=0)]
    return spec, [17, 11, 40, 9], 13


def _bag_ids(rng, B, vocabs, spec, cols):
    ids = np.zeros((B, cols), dtype=np.int64)
    for f in spec:
        ids[:, f["col"]:f["col"] + f["bag"]] = rng.integers(0, vocabs[f["table"]], (B, f["bag"]))
    ids[0, 4:10] = 0            # an all-padding history: count clamps to 1 (ref :173)
    ids[1, 3] = 0               # the padding id as a single lookup: zero row, zero gradient (ref :100)
    ids[2, 4:10] = ids[2, 3]    # heavy duplicates
    return ids


def test_general_oracle_bags_and_shared_tables_match_torch():
    """Fields that share a table, a padding row, and bags pooled like the reference pools item_seq (src/model_fibinet.py:165-174:
    embedding -> mask ids == 0 -> sum -> / clamp(count, 1)), against torch autograd + the reference's own SENetLayer /
    BilinearInteraction."""
    ref = _ref_classes()
    spec, vocabs, cols = _bag_spec()
    F, D, B, hidden = len(spec), 8, 20, (32, 16)
    P = gen.make_params(F, D, 5, hidden, "all", reduction_ratio=2, seed=4)
    rng = np.random.default_rng(9)
    P["tables"] = [(rng.standard_normal((v, D)) * 0.3) for v in vocabs]
    P["tables"][2][0] = 0.0                                                   # nn.Embedding(padding_idx=0)
    ids = _bag_ids(rng, B, vocabs, spec, cols)
    labels = rng.integers(0, 2, B).astype(np.float64)
    prob, cache = gen.forward(P, ids, spec=spec)
    dprob = (prob - labels) / np.maximum(prob * (1 - prob), 1e-12) / B
    G = gen.backward(P, cache, dprob)

    t = lambda a: torch.tensor(np.asarray(a), dtype=torch.float64, requires_grad=True)
    tables = [t(x) for x in P["tables"]]
    senet = ref.SENetLayer(F, reduction_ratio=2).double()
    bil = ref.BilinearInteraction(D, F, bilinear_type="all").double()
    with torch.no_grad():
        senet.excitation[0].weight.copy_(torch.from_numpy(P["se_w1"])); senet.excitation[0].bias.copy_(torch.from_numpy(P["se_b1"]))
        senet.excitation[2].weight.copy_(torch.from_numpy(P["se_w2"])); senet.excitation[2].bias.copy_(torch.from_numpy(P["se_b2"]))
        bil.W.copy_(torch.from_numpy(P["bil_w"][0]))
    tid = torch.from_numpy(ids)
    feats = []
    for f in spec:
        idf = tid[:, f["col"]:f["col"] + f["bag"]]
        e = torch.nn.functional.embedding(idf, tables[f["table"]], padding_idx=f["pad"] if f["pad"] >= 0 else None)
        if f["bag"] == 1:
            feats.append(e[:, 0])
        else:                                                                   # the reference's pooling, lines 165-174
            mask = idf == 0
            e = e * (~mask).unsqueeze(-1).to(e.dtype)
            feats.append(e.sum(1) / (~mask).sum(1, keepdim=True).clamp(min=1).to(e.dtype))
    X = torch.stack(feats, dim=1)
    a = torch.cat([senet(X).flatten(1), bil(senet(X)).flatten(1)], dim=1)
    lin = []
    for i, h in enumerate(hidden):
        w, b, g, be = t(P[f"w{i}"]), t(P[f"b{i}"]), t(P[f"bn_g{i}"]), t(P[f"bn_b{i}"])
        lin.append((w, g, be))
        a = torch.relu(torch.nn.functional.batch_norm(a @ w.T + b, None, None, g, be, training=True, eps=1e-5))
    wo, bo = t(P["w_out"]), t(P["b_out"])
    p_t = torch.sigmoid((a @ wo.T + bo)[:, 0])
    assert np.abs(p_t.detach().numpy() - prob).max() <= 1e-12
    torch.nn.BCELoss()(p_t, torch.from_numpy(labels)).backward()
    for k in range(len(tables)):
        np.testing.assert_allclose(G["tables"][k], tables[k].grad.numpy(), rtol=1e-8, atol=1e-11, err_msg=f"table {k}")
    assert np.all(G["tables"][2][0] == 0)
    np.testing.assert_allclose(G["bil_w"][0], bil.W.grad.numpy(), rtol=1e-8, atol=1e-11)
    np.testing.assert_allclose(G["se_w1"], senet.excitation[0].weight.grad.numpy(), rtol=1e-8, atol=1e-11)
    for i, (w, g, be) in enumerate(lin):
        np.testing.assert_allclose(G[f"w{i}"], w.grad.numpy(), rtol=1e-8, atol=1e-11)
