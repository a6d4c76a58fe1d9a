"""GPU tests of the loop-body engines' host logic: hyper-parameter hand-off without host syncs, the shared dropout
stream, out-of-range ids (IndexError like nn.Embedding, reference src/model_fibinet.py:155-167), loss weighting."""
import numpy as np
import pytest
import torch

from oracle import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    from gpu_common import make_model, to_dev
    import ctr_recommendation_b200  # noqa: F401
    return dict(make_model=make_model, to_dev=to_dev)


def _pinned(b):
    return {k: torch.from_numpy(np.ascontiguousarray(v)).pin_memory() for k, v in b.items() if k != "user_id"}


def test_many_unsynchronised_steps_graph_equals_eager(gpu):
    """The host runs far ahead of the device (no .item(), no synchronize between steps) while OneCycleLR rewrites lr and
    beta1 every step: each queued step must still see ITS hyper-parameters (ring of pinned slots), i.e. the CUDA-graph
    engine ends bit-identical to the same launches issued eagerly, and different from a run with frozen hyper-parameters."""
    from ctr_recommendation_b200 import FusedAdam
    from ctr_recommendation_b200.engine import TrainStep
    B, steps = 2048, 72                       # > 4 * HYPER_SLOTS queued copies
    pool = [synth.make_batch(seed=900 + s, batch=B, index_dtype=np.float64, edge_cases=False) for s in range(3)]
    dev = [({k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in b.items() if k != "user_id"}, torch.from_numpy(y).cuda())
           for b, y in pool]
    finals = []
    for graph in (True, False):
        model = gpu["make_model"](train=True, precision="tf32x3")
        opt = FusedAdam(model, lr=1e-3, weight_decay=1e-5)
        sched = torch.optim.lr_scheduler.OneCycleLR(opt, max_lr=1e-2, total_steps=steps + 8, pct_start=0.3)
        eng = TrainStep(model, opt, B, 20, idx_dtype=torch.float64, graph=graph)
        assert TrainStep.HYPER_SLOTS * 4 <= steps
        for s in range(steps):
            eng(*dev[s % 3])
            sched.step()
        torch.cuda.synchronize()
        finals.append({k: v.detach().cpu().numpy().copy() for k, v in model.state_dict().items()})
        m, v = opt.moments()["mlp.4.weight"]
        finals[-1]["_m"] = m.cpu().numpy().copy()
    for k in finals[0]:
        assert np.array_equal(finals[0][k], finals[1][k]), k


def test_engines_of_one_model_share_the_dropout_stream(gpu):
    from ctr_recommendation_b200 import FusedAdam
    from ctr_recommendation_b200.engine import TrainStep
    model = gpu["make_model"](train=True, precision="fp32")
    opt = FusedAdam(model, lr=1e-3)
    a = TrainStep(model, opt, 256, 20, idx_dtype=torch.float64)
    b = TrainStep(model, opt, 96, 20, idx_dtype=torch.float64)
    assert a.step_counter.data_ptr() == b.step_counter.data_ptr()
    ba, ya = synth.make_batch(seed=1, batch=256, index_dtype=np.float64)
    bb, yb = synth.make_batch(seed=2, batch=96, index_dtype=np.float64)
    for _ in range(3):
        a(_pinned(ba), torch.from_numpy(ya).pin_memory())
    b(_pinned(bb), torch.from_numpy(yb).pin_memory())
    assert int(a.step_counter.item()) == 4
    model.set_dropout_counter(11)
    assert int(b.step_counter.item()) == 11


def test_dropout_seed_follows_torch_seed(gpu):
    from ctr_recommendation_b200 import build_model
    torch.manual_seed(123)
    s1 = build_model(None, {"embedding_dim": 128})._seed
    torch.manual_seed(123)
    s2 = build_model(None, {"embedding_dim": 128})._seed
    torch.manual_seed(124)
    s3 = build_model(None, {"embedding_dim": 128})._seed
    assert s1 == s2 and s1 != s3


@pytest.mark.parametrize("column,value", [("item_id", 91718.0), ("item_id", -1.0), ("likes_level", 11.0), ("views_level", -2.0),
                                          ("item_seq", 91718)])
def test_out_of_range_ids_raise_index_error(gpu, column, value):
    model = gpu["make_model"]()
    batch, _ = synth.make_batch(seed=5, batch=64, index_dtype=np.float64)
    with torch.no_grad():
        model(gpu["to_dev"](batch))                       # a clean batch passes
    if column == "item_seq":
        batch["item_seq"][17, -1] = value
    else:
        batch[column][17] = value
    with pytest.raises(IndexError, match="index out of range"):
        with torch.no_grad():
            model(gpu["to_dev"](batch))
    batch2, _ = synth.make_batch(seed=5, batch=64, index_dtype=np.float64)
    with torch.no_grad():
        model(gpu["to_dev"](batch2))                      # the flag was cleared by the raise


def test_engine_check_ids(gpu):
    from ctr_recommendation_b200.engine import Scorer
    model = gpu["make_model"]()
    sc = Scorer(model, 128, 20, idx_dtype=torch.int64)
    batch, _ = synth.make_batch(seed=6, batch=128, index_dtype=np.int64)
    sc(gpu["to_dev"](batch))
    sc.check_ids()
    batch["item_id"][3] = 10 ** 9
    sc(gpu["to_dev"](batch))
    with pytest.raises(IndexError):
        sc.check_ids()
    sc.check_ids()                                        # sticky flag cleared


def test_global_batch_loss_weight(gpu):
    """DataParallel semantics: a rank's gradient contribution is (rows / global rows) x its local-mean gradient."""
    from ctr_recommendation_b200 import FusedAdam
    from ctr_recommendation_b200.engine import TrainStep
    b, y = synth.make_batch(seed=3, batch=200, index_dtype=np.float64)
    grads = []
    for n in (None, 500):
        model = gpu["make_model"](train=True, precision="fp32")
        model.dropout_p = 0.0
        opt = FusedAdam(model, lr=1e-3)
        eng = TrainStep(model, opt, 200, 20, idx_dtype=torch.float64, graph=False, global_batch=n)
        eng(_pinned(b), torch.from_numpy(y).pin_memory())
        torch.cuda.synchronize()
        grads.append(model._gflat.clone())
    ratio = (grads[1].double().norm() / grads[0].double().norm()).item()
    assert abs(ratio - 200 / 500) <= 1e-5


@pytest.mark.parametrize("wd", [0.0, 1e-5])
@pytest.mark.parametrize("engine", [False, True])
def test_fused_adagrad_vs_oracle(gpu, wd, engine):
    """FusedAdagrad (module path and TrainStep graph) against the oracle's Adagrad (pinned to torch.optim.Adagrad on CPU) fed the
    GPU's own gradients: element-wise to fp32 rounding, including untouched table rows (identity for wd == 0, pure decay else)."""
    from oracle import fibinet_numpy as orc
    from ctr_recommendation_b200 import FusedAdagrad, clip_grad_norm_
    from ctr_recommendation_b200.engine import TrainStep
    from gpu_common import named_grads
    B, lr = 300, 1e-2
    model = gpu["make_model"](train=True, precision="fp32")
    model.dropout_p = 0.0
    opt = FusedAdagrad(model, lr=lr, lr_decay=0.01, weight_decay=wd, initial_accumulator_value=0.0)
    oopt = orc.Adagrad(lr=lr, lr_decay=0.01, weight_decay=wd)
    eng = TrainStep(model, opt, B, 20, idx_dtype=torch.float64, max_norm=10.0) if engine else None
    P = {k: v.detach().cpu().numpy().copy() for k, v in model.state_dict().items()}
    for s in range(4):
        batch, labels = synth.make_batch(seed=340 + s, batch=B, id_dist="zipf", index_dtype=np.float64)
        if eng is not None:
            eng(_pinned(batch), torch.from_numpy(labels).pin_memory())
            torch.cuda.synchronize()
            G = {}
            names = {id(p): n for n, p in model.named_parameters()}
            for (field, plist), (off, _) in zip(model._dense_params(), model._layout):
                o = off
                for p in plist:
                    G[names[id(p)]] = model._gflat[o:o + p.numel()].view(p.shape).cpu().numpy().copy()
                    o += (p.numel() + 3) // 4 * 4
        else:
            y = model(gpu["to_dev"](batch))
            torch.nn.BCELoss()(y, torch.from_numpy(labels).cuda()).backward()
            G = named_grads(model)
            clip_grad_norm_(model, 10.0)
            opt.step()
        G["item_emb.weight"] = (model._item_grad * (model._row_touched > 0).unsqueeze(1)).cpu().numpy()
        orc.clip_grad_norm_(G, 10.0)
        oopt.step(P, G)
        sd = {k: v.detach().cpu().numpy() for k, v in model.state_dict().items()}
        for k in G:
            d = np.abs(sd[k].astype(np.float64) - P[k])
            assert d.max() <= 2e-3 * lr + 1e-7, f"step {s} {k}: {d.max():.3e}"
        acc = opt.accumulators()
        for k in G:                                   # continue from the GPU state so errors cannot accumulate
            P[k] = sd[k].copy()
            oopt.state[k]["sum"] = acc[k].cpu().numpy().copy()
    untouched = (model._row_touched == 0).cpu().numpy()
    untouched[0] = False
    w0 = synth.make_weights(7)["item_emb.weight"]
    moved = np.abs(sd["item_emb.weight"] - w0)[untouched]
    if wd == 0.0:
        assert (acc["item_emb.weight"].cpu().numpy()[untouched & (np.abs(acc["item_emb.weight"].cpu().numpy()).sum(1) == 0)] == 0).all()
    assert np.all(sd["item_emb.weight"][0] == 0)
    assert moved.size > 0


@pytest.mark.parametrize("schedule,precision", [("partial", "tf32x3"), ("full", "tf32x3"), ("wgrad", "tf32x3"), ("full", "f16x3"), ("wgrad", "f16x3")])
def test_phased_backward_equals_single_call(gpu, schedule, precision):
    """fbn_backward_phase (the schedules the data-parallel engine interleaves with its gradient all-reduces: CHAIN [+ LEAF1 beside it],
    then the remaining leaves) produces bit-identical gradients, weights and moments to the single fbn_backward call -- same kernels,
    same per-tensor summation order; run here on one GPU without the collectives.  With SMs reserved for a collective the split-K
    factor of the weight gradients may change, so that variant is held to rounding distance instead."""
    from ctr_recommendation_b200 import FusedAdam
    from ctr_recommendation_b200.engine import TrainStep
    B = 9000                                             # large enough for the CTA-pair GEMMs (and their SM reservation) to engage
    pool = [synth.make_batch(seed=1300 + s, batch=B, id_dist="zipf", index_dtype=np.float64, edge_cases=False) for s in range(3)]
    finals = []
    for phased, reserve in ((False, 0), (True, 0), (True, 8)):
        model = gpu["make_model"](train=True, precision=precision)
        opt = FusedAdam(model, lr=1e-3, weight_decay=1e-5)
        eng = TrainStep(model, opt, B, 20, idx_dtype=torch.float64, overlap=schedule, reserve_sms=reserve, phased_single=phased)
        assert eng._phased_single == phased
        for b, y in pool[:1 if reserve else 3]:
            eng(_pinned(b), torch.from_numpy(y).pin_memory())
        torch.cuda.synchronize()
        st = {k: v.detach().cpu().numpy().copy() for k, v in model.state_dict().items()}
        st["_gflat"] = model._gflat.cpu().numpy().copy()
        st["_item_grad"] = (model._item_grad * (model._row_touched > 0).unsqueeze(1)).cpu().numpy()
        st["_m_item"] = opt._m_item.cpu().numpy().copy()
        finals.append(st)
    for k in finals[0]:
        assert np.array_equal(finals[0][k], finals[1][k]), k
    # one step with 8 SMs reserved: same gradients up to the split-K summation order
    model = gpu["make_model"](train=True, precision=precision)
    opt = FusedAdam(model, lr=1e-3, weight_decay=1e-5)
    eng = TrainStep(model, opt, B, 20, idx_dtype=torch.float64)
    eng(_pinned(pool[0][0]), torch.from_numpy(pool[0][1]).pin_memory())
    torch.cuda.synchronize()
    ref, got = model._gflat.double(), torch.from_numpy(finals[2]["_gflat"]).cuda().double()
    assert ((ref - got).abs().max() / ref.abs().max()).item() <= 1e-6
