"""GPU parity of the F-field model (ctr_recommendation_b200/general.py, BASELINE config 5's 40 fields) against the numpy oracle
oracle/fibinet_general.py (fp64; pinned on CPU against the reference's own SENetLayer / BilinearInteraction classes)."""
import numpy as np
import pytest
import torch

from oracle import fibinet_general as gen

pytestmark = pytest.mark.gpu
D = 128


def _load(model, P, vocab):
    F = model.num_fields
    with torch.no_grad():
        model.emb.weight.copy_(torch.from_numpy(np.concatenate(P["tables"], 0).astype(np.float32)))
        e0, e2 = model.senet.excitation[0], model.senet.excitation[2]
        e0.weight.copy_(torch.from_numpy(P["se_w1"].astype(np.float32))); e0.bias.copy_(torch.from_numpy(P["se_b1"].astype(np.float32)))
        e2.weight.copy_(torch.from_numpy(P["se_w2"].astype(np.float32))); e2.bias.copy_(torch.from_numpy(P["se_b2"].astype(np.float32)))
        for w, src in zip(model.bilinear.weights(), P["bil_w"]):
            w.copy_(torch.from_numpy(src.astype(np.float32)))
        for i, (lin, bn) in enumerate(((0, 1), (4, 5))):
            model.mlp[lin].weight.copy_(torch.from_numpy(P[f"w{i}"].astype(np.float32)))
            model.mlp[lin].bias.copy_(torch.from_numpy(P[f"b{i}"].astype(np.float32)))
            model.mlp[bn].weight.copy_(torch.from_numpy(P[f"bn_g{i}"].astype(np.float32)))
            model.mlp[bn].bias.copy_(torch.from_numpy(P[f"bn_b{i}"].astype(np.float32)))
        model.mlp[8].weight.copy_(torch.from_numpy(P["w_out"].astype(np.float32)))
        model.mlp[8].bias.copy_(torch.from_numpy(P["b_out"].astype(np.float32)))


def _round32(P):
    """the oracle sees exactly the fp32 weights the GPU holds"""
    out = {}
    for k, v in P.items():
        if isinstance(v, list):
            out[k] = [a.astype(np.float32).astype(np.float64) for a in v]
        elif isinstance(v, np.ndarray):
            out[k] = v.astype(np.float32).astype(np.float64)
        else:
            out[k] = v
    return out


def _oracle_under_gpu_gates(model, P, ids, labels, B, tol, masks=None, dropout=0.0, spec=None):
    """Oracle forward / backward with the ReLU decisions of the GPU run (kept elements only), after checking that they differ from the
    oracle's own decisions only where its pre-activation is within the tolerance band around zero (see test_gpu_parity.py:
    test_full_batch_forward_backward_vs_oracle)."""
    prob0, cache0 = gen.forward(P, ids, masks=masks, dropout_p=dropout, spec=spec)
    gates = []
    for i, (name, h) in enumerate((("A1", 512), ("A2", 256))):
        A = model.tower_view(name, (B, h)).cpu().numpy()
        own, pre = cache0["layers"][i][3], cache0["layers"][i][5]
        g = (A > 0) if masks is None else np.where(masks[i] > 0, A > 0, own)
        mis = g != own
        assert mis.sum() <= 4 and ((not mis.any()) or np.abs(pre[mis]).max() <= tol * max(1.0, np.abs(pre).max())), (name, int(mis.sum()))
        gates.append(g)
    prob, cache = gen.forward(P, ids, masks=masks, dropout_p=dropout, spec=spec, relu_gates=gates)
    dprob = (prob - labels) / np.maximum(prob * (1 - prob), 1e-12) / B
    return prob0, gen.backward(P, cache, dprob)


CASES = [("fp32", 8, 300, "all", 0.0), ("tf32x3", 8, 300, "all", 0.25), ("tf32x3", 8, 257, "each", 0.0), ("tf32x3", 6, 200, "interaction", 0.25),
         ("tf32x3", 40, 256, "all", 0.0), ("bf16", 8, 300, "all", 0.0), ("f16x3", 8, 300, "all", 0.25), ("f16x3", 40, 256, "all", 0.0)]


@pytest.mark.parametrize("precision,F,B,btype,dropout", CASES)
def test_general_model_vs_oracle(precision, F, B, btype, dropout):
    from ctr_recommendation_b200 import GeneralFiBiNET, build_model
    vocab = 57
    tol = {"fp32": 1e-5, "tf32x3": 1e-5, "f16x3": 1e-5, "bf16": 1e-2}[precision]
    P = _round32(gen.make_params(F, D, vocab, (512, 256), btype, reduction_ratio=2, seed=11))
    rng = np.random.default_rng(7)
    ids = rng.integers(0, vocab, (B, F))
    ids[:, 0] = ids[0, 0]                                   # a hot row: one id owns a whole field (chunked segment sum for B > 256)
    labels = rng.integers(0, 2, B).astype(np.float64)
    masks = [(rng.random((B, h)) >= dropout).astype(np.uint8) for h in (512, 256)] if dropout > 0 else None
    fm = {"fields": [(f"f{i}", vocab) for i in range(F)], "bilinear_type": btype, "senet_reduction": 2, "dropout": dropout,
          "precision": precision}
    model = build_model(fm, {"embedding_dim": 128})
    assert isinstance(model, GeneralFiBiNET) and model.k1 == (F + F * (F - 1) // 2) * D
    _load(model, P, vocab)
    model = model.cuda().train()
    if masks is not None:
        model._test_masks = tuple(torch.from_numpy(m) for m in masks)
    batch = {f"f{i}": torch.from_numpy(ids[:, i].copy()).cuda() for i in range(F)}
    y = model(batch)
    loss = torch.nn.BCELoss()(y, torch.from_numpy(labels.astype(np.float32)).cuda())
    loss.backward()
    rel = lambda a, b: float(np.abs(np.asarray(a, np.float64) - b).max() / max(np.abs(b).max(), 1e-30))
    if precision == "bf16":                                 # bf16 gradients: unpinned
        prob, _ = gen.forward(P, ids, masks=masks, dropout_p=dropout)
        assert rel(y.detach().cpu().numpy(), prob) <= tol
        return
    prob, G = _oracle_under_gpu_gates(model, P, ids, labels, B, tol, masks, dropout)
    assert rel(y.detach().cpu().numpy(), prob) <= tol, rel(y.detach().cpu().numpy(), prob)
    got = {
        "tables": model.emb.weight.grad.cpu().numpy(), "se_w1": model.senet.excitation[0].weight.grad.cpu().numpy(),
        "se_b1": model.senet.excitation[0].bias.grad.cpu().numpy(), "se_w2": model.senet.excitation[2].weight.grad.cpu().numpy(),
        "se_b2": model.senet.excitation[2].bias.grad.cpu().numpy(),
        "bil_w": np.stack([w.grad.cpu().numpy() for w in model.bilinear.weights()]),
        "w0": model.mlp[0].weight.grad.cpu().numpy(), "bn_g0": model.mlp[1].weight.grad.cpu().numpy(), "bn_b0": model.mlp[1].bias.grad.cpu().numpy(),
        "w1": model.mlp[4].weight.grad.cpu().numpy(), "bn_g1": model.mlp[5].weight.grad.cpu().numpy(), "bn_b1": model.mlp[5].bias.grad.cpu().numpy(),
        "w_out": model.mlp[8].weight.grad.cpu().numpy(), "b_out": model.mlp[8].bias.grad.cpu().numpy()}
    want = dict(G)
    want["tables"] = np.concatenate(G["tables"], 0)
    want["bil_w"] = np.stack(G["bil_w"])
    bad = []
    for k, g in got.items():
        e = rel(g, np.asarray(want[k]).reshape(g.shape))
        if e > tol:
            bad.append(f"{k}: {e:.3e}")
    assert not bad, "; ".join(bad)
    assert model.mlp[0].bias.grad.abs().max().item() <= 2e-7           # Linear bias in front of BatchNorm: exactly-zero true gradient


def test_general_model_eval_and_bad_ids():
    from ctr_recommendation_b200 import GeneralFiBiNET
    model = GeneralFiBiNET([("a", 10), ("b", 20), ("c", 5)], precision="fp32").cuda().eval()
    ids = torch.tensor([[1, 2, 3], [9, 19, 4]], device="cuda")
    with torch.no_grad():
        p = model({"ids": ids})
    assert p.shape == (2,) and torch.isfinite(p).all() and (p > 0).all() and (p < 1).all()
    # field offsets: (a, 9) and (b, 0) are different rows of the shared table
    X = model._buf[2]["X"]
    assert torch.equal(X[1, 0], model.emb.weight[9]) and torch.equal(X[1, 1], model.emb.weight[10 + 19]) and torch.equal(X[1, 2], model.emb.weight[30 + 4])
    with pytest.raises(IndexError):
        with torch.no_grad():
            model({"ids": torch.tensor([[1, 20, 3], [0, 0, 0]], device="cuda")})


def test_general_model_trains_with_a_torch_optimizer():
    from ctr_recommendation_b200 import GeneralFiBiNET
    torch.manual_seed(0)
    F, B, vocab = 5, 512, 40
    model = GeneralFiBiNET([(f"f{i}", vocab) for i in range(F)], precision="tf32x3", dropout=0.0).cuda().train()
    opt = torch.optim.Adam(model.parameters(), lr=1e-2)
    g = torch.Generator().manual_seed(1)
    ids = torch.randint(0, vocab, (B, F), generator=g).cuda()
    y = ((ids[:, 0] + ids[:, 1]) % 2).float()
    first = None
    for _ in range(30):
        opt.zero_grad()
        loss = torch.nn.BCELoss()(model({"ids": ids}), y)
        loss.backward()
        opt.step()
        first = first if first is not None else loss.item()
    assert loss.item() < 0.5 * first


@pytest.mark.parametrize("precision", ["fp32", "tf32x3"])
def test_bags_shared_tables_and_padding_vs_oracle(precision):
    """The field kinds the reference model itself is made of, plus the ones it leaves unused: a learned user lookup, two
    fields on one table (likes / views -> cate_emb), a table with a padding row, a history bag pooled into the item table exactly like
    src/model_fibinet.py:165-174, and an item_tags bag (src/dataloader.py:100-102) -- forward and every gradient against the oracle."""
    from ctr_recommendation_b200 import build_model
    from test_oracle_general import _bag_ids
    vocabs = [600, 11, 900, 50]
    spec = [dict(table=0, col=0, bag=1, pad=-1), dict(table=1, col=1, bag=1, pad=-1), dict(table=1, col=2, bag=1, pad=-1),
            dict(table=2, col=3, bag=1, pad=0), dict(table=2, col=4, bag=6, pad=0), dict(table=3, col=10, bag=3, pad=0)]
    # same layout as tests/test_oracle_general._bag_spec (history bag of 6 at columns 4..9, tag bag of 3 at 10..12)
    fm = {"fields": [{"name": "user_id", "vocab": vocabs[0]}, {"name": "likes_level", "vocab": vocabs[1]},
                     {"name": "views_level", "table": "likes_level"}, {"name": "item_id", "vocab": vocabs[2], "padding_idx": 0},
                     {"name": "item_seq", "table": "item_id", "bag": 6}, {"name": "item_tags", "vocab": vocabs[3], "bag": 3}],
          "precision": precision, "dropout": 0.0}
    F, B, cols = 6, 700, 13
    P = _round32(gen.make_params(F, D, 5, (512, 256), "all", reduction_ratio=2, seed=21))
    rng = np.random.default_rng(3)
    P["tables"] = [(rng.standard_normal((v, D)) * 0.3).astype(np.float32).astype(np.float64) for v in vocabs]
    P["tables"][2][0] = 0.0
    ids = _bag_ids(rng, B, vocabs, spec, cols)
    ids[:, 1] = 4                                             # a hot row of the shared table: > 256 occurrences -> chunked segment sum
    labels = rng.integers(0, 2, B).astype(np.float64)
    cache = gen.forward(P, ids, spec=spec)[1]
    model = build_model(fm, {"embedding_dim": 128})
    assert model.id_cols == cols and model.k1 == 21 * D and model.emb.weight.shape[0] == sum(vocabs)
    _load(model, P, None)
    model = model.cuda().train()
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    batch = {"user_id": t(ids[:, 0]), "likes_level": t(ids[:, 1]), "views_level": t(ids[:, 2]), "item_id": t(ids[:, 3]),
             "item_seq": t(ids[:, 4:10]), "item_tags": t(ids[:, 10:13])}
    y = model(batch)
    torch.nn.BCELoss()(y, torch.from_numpy(labels.astype(np.float32)).cuda()).backward()
    rel = lambda a, b: float(np.abs(np.asarray(a, np.float64) - b).max() / max(np.abs(b).max(), 1e-30))
    prob, G = _oracle_under_gpu_gates(model, P, ids, labels, B, 1e-5, spec=spec)
    assert rel(y.detach().cpu().numpy(), prob) <= 1e-5
    X = model._buf[B]["X"].cpu().numpy()
    assert rel(X, cache["X"]) <= 1e-6 and np.all(X[1, 3] == 0) and np.all(X[0, 4] == 0)          # padding lookup / all-padding bag
    g_emb = model.emb.weight.grad.cpu().numpy()
    want = np.concatenate(G["tables"], 0)
    errs = {"emb": rel(g_emb, want)}
    o = 0
    for k, v in enumerate(vocabs):
        errs[f"table{k}"] = rel(g_emb[o:o + v], G["tables"][k])
        o += v
    for got, key in ((model.mlp[0].weight.grad, "w0"), (model.mlp[4].weight.grad, "w1"), (model.mlp[8].weight.grad, "w_out"),
                     (model.senet.excitation[0].weight.grad, "se_w1"), (model.senet.excitation[2].weight.grad, "se_w2"),
                     (model.bilinear.W.grad, None)):
        ref = G["bil_w"][0] if key is None else G[key]
        errs[key or "bil_w"] = rel(got.cpu().numpy(), np.asarray(ref).reshape(got.shape))
    bad = {k: v for k, v in errs.items() if v > 1e-5}
    assert not bad, "; ".join(f"{k} {v:.2e}" for k, v in errs.items())
    assert np.all(g_emb[vocabs[0] + vocabs[1]] == 0)                                               # the padding row gets no gradient
    # same batch as one (B, id columns) tensor
    with torch.no_grad():
        model.eval()
        a = model(batch)
        b = model({"ids": t(ids)})
    assert torch.equal(a, b)
