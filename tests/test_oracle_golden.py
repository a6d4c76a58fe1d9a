"""CPU: the numpy oracle reproduces the reference's recorded outputs (tests/golden)."""
import numpy as np
import pytest

from oracle import fibinet_numpy as orc
from oracle import synth
from helpers import check_summary, rel_err, unpack_mask

TOL = 1e-5  # fp32 logits / gradients: 1e-5 relative (north_star)
# Linear biases feeding a train-mode BatchNorm have an exactly-zero true gradient; the reference
# stores ~1e-8 rounding noise there, so those tensors get an absolute floor.
GRAD_ATOL = 2e-7
# ... and Adam normalises that noise into full-size steps, so the two biases (which have no effect
# on the function: BatchNorm subtracts them) are chaotic in the reference itself.
# Multi-step state: Adam turns ~1e-7 gradient differences on near-zero entries into lr-sized steps, so
# after 2-3 steps weights/moments agree to ~5e-5 rather than 1e-5 (single-step quantities use TOL).
DRIFT_TOL = 5e-5
NOISE_DRIVEN = ("mlp.0.bias", "mlp.4.bias")


@pytest.mark.parametrize("tag,id_dist,total", [("train_u", "uniform", 40), ("train_z", "zipf", 10)])
def test_train_steps_match_reference(golden, tag, id_dist, total):
    B, steps, total_steps, with_seq = [int(v) for v in golden[f"{tag}/meta"]]
    assert total_steps == total
    P = synth.make_weights(seed=7)
    opt = orc.Adam(lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-5)
    sched = orc.OneCycle(max_lr=1e-2, total_steps=total_steps)
    for s in range(steps):
        lr, b1 = sched.at(s)
        glr, gb1 = golden[f"{tag}/step{s}/lr_beta1"]
        assert abs(lr - glr) <= 1e-12 + 1e-9 * glr and abs(b1 - gb1) <= 1e-12
        opt.lr, opt.betas = lr, (b1, 0.999)
        batch, labels = synth.make_batch(seed=100 + s, batch=B, id_dist=id_dist, index_dtype=np.float64)
        m1 = unpack_mask(golden[f"{tag}/step{s}/mask1"], (B, 512))
        m2 = unpack_mask(golden[f"{tag}/step{s}/mask2"], (B, 256))
        if s == 0:
            prob, cache = orc.forward(dict(P), batch, train=True, masks=(m1, m2), update_running=False)
            _, dprob = orc.bce_loss(prob, labels)
            G = orc.backward(P, cache, dprob)
            assert bool(golden[f"{tag}/grad0/user_emb_is_none"]) and "user_emb.weight" not in G
            for k, g in G.items():
                check_summary(golden, f"{tag}/grad0", k, g, TOL, atol=GRAD_ATOL)
        loss, prob, total_norm, _ = orc.train_step(P, opt, batch, labels, masks=(m1, m2))
        assert rel_err(prob, golden[f"{tag}/step{s}/prob"]) <= TOL
        assert abs(loss - float(golden[f"{tag}/step{s}/loss"])) <= 1e-5
        # torch reduces the 1.4 M-element norm in fp32 (observed 2.3e-5 low vs an fp64 sum of ITS OWN grads)
        assert abs(total_norm - float(golden[f"{tag}/step{s}/total_norm"])) <= 1e-4 * total_norm
    for k, v in P.items():
        if k in NOISE_DRIVEN:
            assert np.abs(v - synth.make_weights(seed=7)[k]).max() <= steps * 1e-2  # bounded by lr per step
            continue
        # running_mean absorbs the chaotic bias above (momentum 0.1 x lr-sized drift per step)
        check_summary(golden, f"{tag}/final", k, v, DRIFT_TOL, atol=2e-4 if k.endswith("running_mean") else 0.0)
    for k, st in opt.state.items():
        if k in NOISE_DRIVEN:
            continue
        check_summary(golden, f"{tag}/final_m", k, st["m"], 4 * DRIFT_TOL, atol=1e-9)
        check_summary(golden, f"{tag}/final_v", k, st["v"], 4 * DRIFT_TOL, atol=1e-12)
    batch, _ = synth.make_batch(seed=900, batch=300, id_dist=id_dist, index_dtype=np.int64)
    prob, _ = orc.forward(P, batch, train=False)
    assert rel_err(prob, golden[f"{tag}/eval_prob"]) <= 5e-5  # after 2-3 Adam steps of drift


CASES = {
    "eval/f64_b300": dict(batch=300, index_dtype=np.float64),
    "eval/i64_b1": dict(batch=1, index_dtype=np.int64, edge_cases=False),
    "eval/i32_b777_zipf": dict(batch=777, index_dtype=np.int32, id_dist="zipf"),
    "eval/noseq_b64": dict(batch=64, index_dtype=np.int64, with_seq=False),
    "eval/short_b50_l7": dict(batch=50, index_dtype=np.int64, max_len=7),
}


@pytest.mark.parametrize("tag", list(CASES))
def test_eval_forward_matches_reference(golden, tag):
    P = synth.make_weights(seed=7)
    batch, _ = synth.make_batch(seed=321, **CASES[tag])
    prob, cache = orc.forward(P, batch, train=False)
    assert prob.dtype == np.float32 and prob.shape == (CASES[tag]["batch"],)
    assert rel_err(prob, golden[f"{tag}/prob"]) <= TOL
    if tag == "eval/f64_b300":
        X = cache["X"]
        assert np.array_equal(X[:, 3], golden[f"{tag}/item_f"])       # gathers: bit exact
        assert np.array_equal(X[:, 1], golden[f"{tag}/like_f"])
        assert np.all(X[:, 0] == 0)
        assert rel_err(X[:, 4], golden[f"{tag}/img_f"]) <= TOL
        assert rel_err(X[:, 5], golden[f"{tag}/hist_f"]) <= TOL


@pytest.mark.parametrize("ratio", [2, 3])
def test_senet_module(golden, ratio):
    t = f"senet_r{ratio}"
    B, F, D = 37, 6, 128
    x = synth.normal(5, 1, B * F * D).reshape(B, F, D).astype(np.float32)
    gy = synth.normal(5, 2, B * F * D).reshape(B, F, D).astype(np.float32)
    w1, b1 = golden[f"{t}/w/excitation.0.weight"], golden[f"{t}/w/excitation.0.bias"]
    w2, b2 = golden[f"{t}/w/excitation.2.weight"], golden[f"{t}/w/excitation.2.bias"]
    assert w1.shape == (max(1, F // ratio), F)
    y, saved = orc.senet_forward(x, w1, b1, w2, b2)
    assert rel_err(y, golden[f"{t}/y"]) <= TOL
    dx, dw1, db1, dw2, db2 = orc.senet_backward(x, w1, w2, saved, gy)
    assert rel_err(dx, golden[f"{t}/dx"]) <= TOL
    assert rel_err(dw1, golden[f"{t}/g/excitation.0.weight"]) <= TOL
    assert rel_err(db1, golden[f"{t}/g/excitation.0.bias"]) <= TOL
    assert rel_err(dw2, golden[f"{t}/g/excitation.2.weight"]) <= TOL
    assert rel_err(db2, golden[f"{t}/g/excitation.2.bias"]) <= TOL


@pytest.mark.parametrize("bt", ["all", "each"])
def test_bilinear_module(golden, bt):
    t = f"bilinear_{bt}"
    B, F, D = 37, 6, 128
    x = synth.normal(5, 1, B * F * D).reshape(B, F, D).astype(np.float32)
    gp = synth.normal(5, 3, B * 15 * D).reshape(B, 15, D).astype(np.float32)
    if bt == "all":
        W = golden[f"{t}/w/W"]
    else:
        W = [golden[f"{t}/w/W_list.{i}"] for i in range(F - 1)]
    y = orc.bilinear_forward(x, W, bt)
    assert rel_err(y, golden[f"{t}/y"]) <= TOL
    dx, dW = orc.bilinear_backward(x, W, gp, bt)
    assert rel_err(dx, golden[f"{t}/dx"]) <= TOL
    if bt == "all":
        assert rel_err(dW, golden[f"{t}/g/W"]) <= TOL
    else:
        for i in range(F - 1):
            assert rel_err(dW[i], golden[f"{t}/g/W_list.{i}"]) <= TOL


def test_bilinear_rejects_unknown_type():
    with pytest.raises(ValueError):
        orc.bilinear_forward(np.zeros((1, 3, 4), np.float32), None, "bogus")


def test_each_model(golden):
    P = synth.make_weights(seed=7, bilinear_type="each")
    batch, _ = synth.make_batch(seed=321, batch=130, index_dtype=np.int64)
    prob, _ = orc.forward(P, batch, train=False)
    assert rel_err(prob, golden["each_model/prob"]) <= TOL


def test_onecycle_schedule(golden):
    traj = golden["schedule/lr_beta1"]
    sched = orc.OneCycle(max_lr=1e-2, total_steps=100)
    for s in range(100):
        lr, b1 = sched.at(s)
        assert abs(lr - traj[s, 0]) <= 1e-9 * traj[s, 0] + 1e-15
        assert abs(b1 - traj[s, 1]) <= 1e-12
    assert abs(traj[0, 0] - 4e-4) < 1e-12 and abs(traj[0, 1] - 0.95) < 1e-12


def test_auc_matches_sklearn_and_single_class():
    from sklearn.metrics import roc_auc_score
    rng = np.random.default_rng(0)
    y = (rng.random(5000) < 0.4).astype(np.float32)
    s = np.round(rng.random(5000), 2)   # many ties
    assert abs(orc.auc(y, s) - roc_auc_score(y, s)) < 1e-12
    assert orc.auc(np.ones(10), rng.random(10)) == 0.5   # src/utils.py:25-27


def test_oracle_adamw_matches_torch():
    """The oracle's decoupled option (used to check FusedAdam(decoupled_weight_decay=True)) against torch.optim.AdamW on CPU."""
    import torch
    rng = np.random.default_rng(0)
    p0 = rng.standard_normal((37, 16)).astype(np.float32)
    P = {"w": p0.copy()}
    opt = orc.Adam(lr=3e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, decoupled=True)
    tp = torch.nn.Parameter(torch.from_numpy(p0.copy()))
    topt = torch.optim.AdamW([tp], lr=3e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2)
    for s in range(4):
        g = rng.standard_normal(p0.shape).astype(np.float32)
        opt.step(P, {"w": g})
        tp.grad = torch.from_numpy(g.copy())
        topt.step()
        assert np.abs(P["w"] - tp.detach().numpy()).max() <= 2e-7
