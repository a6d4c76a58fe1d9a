"""GPU: stand-alone SENetLayer / BilinearInteraction modules (reference classes) against the golden vectors
recorded from the reference and, for the 'interaction' extension, the numpy oracle."""
import numpy as np
import pytest
import torch

from oracle import fibinet_numpy as orc
from oracle import synth
from helpers import rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _inputs():
    B, F, D = 37, 6, 128
    x = synth.normal(5, 1, B * F * D).reshape(B, F, D).astype(np.float32)
    gy = synth.normal(5, 2, B * F * D).reshape(B, F, D).astype(np.float32)
    gp = synth.normal(5, 3, B * 15 * D).reshape(B, 15, D).astype(np.float32)
    return x, gy, gp


@pytest.mark.parametrize("ratio", [2, 3])
def test_senet_layer_golden(golden, ratio):
    from ctr_recommendation_b200 import SENetLayer
    t = f"senet_r{ratio}"
    x, gy, _ = _inputs()
    m = SENetLayer(6, reduction_ratio=ratio)
    m.load_state_dict({k[len(t) + 3:]: torch.from_numpy(v) for k, v in golden.items() if k.startswith(t + "/w/")}, strict=True)
    m = m.cuda()
    xt = torch.from_numpy(x).cuda().requires_grad_(True)
    y = m(xt)
    y.backward(torch.from_numpy(gy).cuda())
    assert rel_err(y.detach().cpu().numpy(), golden[f"{t}/y"]) <= TOL
    assert rel_err(xt.grad.cpu().numpy(), golden[f"{t}/dx"]) <= TOL
    for k, p in m.named_parameters():
        assert rel_err(p.grad.cpu().numpy(), golden[f"{t}/g/{k}"]) <= TOL, k


@pytest.mark.parametrize("precision,tol", [("fp32", TOL), ("tf32x3", TOL)])
@pytest.mark.parametrize("bt", ["all", "each"])
def test_bilinear_interaction_golden(golden, bt, precision, tol):
    from ctr_recommendation_b200 import BilinearInteraction
    from ctr_recommendation_b200.functional import bilinear
    t = f"bilinear_{bt}"
    x, _, gp = _inputs()
    m = BilinearInteraction(128, 6, bt)
    m.load_state_dict({k[len(t) + 3:]: torch.from_numpy(v) for k, v in golden.items() if k.startswith(t + "/w/")}, strict=True)
    m = m.cuda()
    xt = torch.from_numpy(x).cuda().requires_grad_(True)
    y = bilinear(xt, m.weights(), bt, precision=precision)
    y.backward(torch.from_numpy(gp).cuda())
    assert rel_err(y.detach().cpu().numpy(), golden[f"{t}/y"]) <= tol
    assert rel_err(xt.grad.cpu().numpy(), golden[f"{t}/dx"]) <= tol
    for k, p in m.named_parameters():
        assert rel_err(p.grad.cpu().numpy(), golden[f"{t}/g/{k}"]) <= tol, k
    if precision == "fp32":
        y2 = m(xt)                                  # module forward == functional
        assert torch.equal(y2, y)


@pytest.mark.parametrize("F,D", [(6, 128), (9, 64), (3, 256)])
def test_bilinear_field_interaction_vs_oracle(F, D):
    """'interaction' (one matrix per pair) does not exist in the reference (ValueError there): extension, unpinned --
    checked against the numpy restatement of the FiBiNET-paper formula."""
    from ctr_recommendation_b200.functional import bilinear
    B, P = 50, F * (F - 1) // 2
    x = synth.normal(8, 1, B * F * D).reshape(B, F, D).astype(np.float32)
    W = [(synth.normal(8, 10 + i, D * D).reshape(D, D) / np.sqrt(D)).astype(np.float32) for i in range(P)]
    gp = synth.normal(8, 3, B * P * D).reshape(B, P, D).astype(np.float32)
    xt = torch.from_numpy(x).cuda().requires_grad_(True)
    Wt = [torch.from_numpy(w).cuda().requires_grad_(True) for w in W]
    y = bilinear(xt, Wt, "field_interaction")
    y.backward(torch.from_numpy(gp).cuda())
    ref = orc.bilinear_forward(x, W, "interaction")
    dx, dW = orc.bilinear_backward(x, W, gp, "interaction")
    assert rel_err(y.detach().cpu().numpy(), ref) <= TOL
    assert rel_err(xt.grad.cpu().numpy(), dx) <= TOL
    for i in range(P):
        assert rel_err(Wt[i].grad.cpu().numpy(), dW[i]) <= TOL, i


@pytest.mark.parametrize("bt", ["each", "interaction"])
def test_full_model_other_bilinear_types(golden, bt):
    from gpu_common import make_model, to_dev, named_grads
    B = 130
    batch, labels = synth.make_batch(seed=321, batch=B, index_dtype=np.int64)
    if bt == "each":
        model = make_model(bilinear_type="each")
        with torch.no_grad():
            prob = model(to_dev(batch)).cpu().numpy()
        assert rel_err(prob, golden["each_model/prob"]) <= TOL
    # train-mode gradients against the oracle
    W = synth.make_weights(seed=7, bilinear_type="each")
    if bt == "interaction":
        for i in range(5, 15):
            W[f"bilinear.W_list.{i}"] = (synth.normal(77, i, 128 * 128).reshape(128, 128) * 0.0884).astype(np.float32)
    from ctr_recommendation_b200 import build_model
    model = build_model({"bilinear_type": bt}, {"embedding_dim": 128})
    model.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in W.items()}, strict=True)
    model = model.cuda().train()
    m1, m2 = synth.make_dropout_masks(5, B)
    model._test_masks = (torch.from_numpy(m1), torch.from_numpy(m2))
    y = model(to_dev(batch))
    torch.nn.BCELoss()(y, torch.from_numpy(labels).cuda()).backward()
    prob, cache = orc.forward(dict(W), batch, train=True, masks=(m1, m2), update_running=False)
    _, dprob = orc.bce_loss(prob, labels)
    G = orc.backward(W, cache, dprob)
    assert rel_err(y.detach().cpu().numpy(), prob) <= TOL
    got = named_grads(model)
    assert set(got) == set(G)
    for k in G:
        scale = max(np.abs(G[k]).max(), 1e-30)
        assert np.abs(got[k].astype(np.float64) - G[k]).max() <= TOL * scale + 2e-7, k


@pytest.mark.parametrize("ratio,hidden", [(1, 6), (3, 2), (6, 1), (7, 1)])
@pytest.mark.parametrize("precision", ["fp32", "tf32x3"])
def test_senet_reduction_in_the_fused_model(ratio, hidden, precision):
    """`senet_reduction` of config/fibinet_config.yaml (the reference hard-codes ratio 2, src/model_fibinet.py:114): the fused six-field
    kernels run SENetLayer(6, ratio) with hidden = max(1, 6 // ratio) and match the oracle (forward and all gradients) at 1e-5."""
    from gpu_common import make_model, to_dev, named_grads
    from oracle import fibinet_numpy as orc, synth
    from helpers import rel_err
    B = 320
    model = make_model(train=True, precision=precision, senet_reduction=ratio)
    assert model.senet.excitation[0].weight.shape == (hidden, 6)
    batch, labels = synth.make_batch(seed=140 + ratio, batch=B, id_dist="zipf", index_dtype=np.float64)
    m1, m2 = synth.make_dropout_masks(3, B)
    model._test_masks = (torch.from_numpy(m1), torch.from_numpy(m2))
    y = model(to_dev(batch))
    assert model._params_struct().se_hidden == hidden
    torch.nn.BCELoss()(y, torch.from_numpy(labels).cuda()).backward()
    P = synth.make_weights(seed=7, senet_reduction=ratio)
    prob, cache = orc.forward(dict(P), batch, train=True, masks=(m1, m2), update_running=False)
    _, dprob = orc.bce_loss(prob, labels)
    G = orc.backward(P, cache, dprob)
    assert rel_err(y.detach().cpu().numpy(), prob) <= 1e-5
    got = named_grads(model)
    assert set(got) == set(G)
    for k in G:
        err = np.abs(got[k].astype(np.float64) - G[k]).max()
        assert err <= 1e-5 * max(np.abs(G[k]).max(), 1e-30) + 2e-7, (k, err)


def test_honor_config_reads_senet_reduction():
    from ctr_recommendation_b200 import build_model
    cfg = {"embedding_dim": 128, "senet_reduction": 3, "bilinear_type": "each", "net_dropout": 0.25}
    assert build_model(None, cfg).senet.reduced_size == 3                                  # ignored like in the reference
    assert build_model(None, dict(cfg, honor_config=True)).senet.reduced_size == 2         # honoured: max(1, 6 // 3)
