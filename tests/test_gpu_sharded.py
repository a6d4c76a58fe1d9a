"""GPU parity of the row-sharded item table (BASELINE config 5): every kernel is driven through the C ABI.

N "virtual ranks" live on the one GPU of the test box (their slices and exchange blocks are ordinary device buffers, which
is exactly what a peer mapping is to the kernels), so the N-way paths are covered without a second GPU:
  * sharded forward gather == replicated gather, bit for bit;
  * local segment sums + owner-side N-way merge == the oracle's dense scatter-add (oracle/shard_numpy.py), exact row sets,
    run-to-run bitwise determinism;
  * ShardedTrainStep (1 rank) == TrainStep on the replicated table, bit for bit, dense-exact Adam;
  * lazy row Adam == the oracle's restatement on the touched rows, untouched rows bit-identical.
The real 2-GPU run (CUDA IPC + NCCL) is tools/shard_check.py."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import shard_numpy as sorc
from oracle import synth
from helpers import rel_err

pytestmark = pytest.mark.gpu

D = 128


@pytest.fixture(scope="module")
def env():
    from gpu_common import make_model, to_dev, load_weights
    from ctr_recommendation_b200 import _lib, sharded
    return dict(make_model=make_model, to_dev=to_dev, load_weights=load_weights, lib=_lib.load(), _lib=_lib, sharded=sharded)


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_sharded_gather_bit_exact(env, world):
    _lib, lib, sharded = env["_lib"], env["lib"], env["sharded"]
    model = env["make_model"]()
    B, L = 257, 20
    batch, _ = synth.make_batch(seed=77, batch=B, id_dist="zipf", index_dtype=np.float64)
    with torch.no_grad():
        model(env["to_dev"](batch))
    ref = model.workspace_view("X5", (B, 5, D)).clone()
    refC = model.workspace_view("C", (B, 2688)).clone()
    full = model.item_emb.weight.data
    slices = [sharded.slice_of_full(full, r, world) for r in range(world)]
    assert torch.equal(sharded.full_from_slices(slices, full.shape[0]), full)
    bs, keep, _, _ = model._batch_struct(env["to_dev"](batch))
    for rank in range(world):
        P = model._params_struct()
        P.item_emb = slices[rank].data_ptr()
        P.n_shards, P.shard_rank, P.shard_rows = world, rank, slices[0].shape[0]
        for r in range(world):
            P.shard[r] = slices[r].data_ptr()
        ws = torch.zeros(lib.fbn_workspace_bytes(B, L, 1), dtype=torch.uint8, device="cuda")
        _lib.check(lib.fbn_embed_forward(C.byref(P), C.byref(bs), _lib.ptr(ws), ws.numel(), 1, _lib.stream_ptr()), "fbn_embed_forward")
        torch.cuda.synchronize()
        off = lib.fbn_workspace_offset(B, L, 1, b"X5")
        X5 = ws[off:off + B * 5 * D * 4].view(torch.float32).view(B, 5, D)
        assert torch.equal(X5, ref), f"rank {rank}/{world}: sharded gather differs from the replicated gather"
        off = lib.fbn_workspace_offset(B, L, 1, b"C")
        Cm = ws[off:off + B * 2688 * 4].view(torch.float32).view(B, 2688)
        assert torch.equal(Cm[:, 128:768], refC[:, 128:768])


def _virtual_exchange(env, world, V, B, L, seed, id_dist, lazy, integer_grads=False):
    """Runs index + local sums for `world` virtual ranks and the merge for every owner; returns per-owner results."""
    _lib, lib = env["_lib"], env["lib"]
    R = sorc.shard_rows(V, world)
    cap = B * (1 + L)
    merge_cap = world * cap
    xchg = [torch.zeros(lib.fbn_shard_xchg_bytes(cap), dtype=torch.uint8, device="cuda") for _ in range(world)]
    sws = [torch.zeros(lib.fbn_shard_ws_bytes(cap, merge_cap, world, R), dtype=torch.uint8, device="cuda") for _ in range(world)]
    plans, inputs = [], []
    rng = np.random.default_rng(seed)
    for r in range(world):
        p = _lib.ShardPlan()
        p.n_shards, p.rank, p.item_rows, p.shard_rows, p.cap, p.merge_cap = world, r, V, R, cap, merge_cap
        for q in range(world):
            p.xchg[q] = xchg[q].data_ptr()
        plans.append(p)
        batch, _ = synth.make_batch(seed=seed + 31 * r, batch=B, max_len=max(L, 1), id_dist=id_dist, index_dtype=np.int64)
        ids = (batch["item_id"].astype(np.int64) % V)
        seq = (batch["item_seq"].astype(np.int64) % V) if L > 0 else None
        if integer_grads:    # small integers: every fp32 sum is exact, so the routing can be checked bit for bit
            dXi = rng.integers(-8, 9, (B, D)).astype(np.float32)
            dXh = rng.integers(-8, 9, (B, D)).astype(np.float32)
        else:
            dXi = rng.standard_normal((B, D)).astype(np.float32)
            dXh = rng.standard_normal((B, D)).astype(np.float32)
        inputs.append((ids, seq, dXi, dXh))
    st = _lib.stream_ptr()
    keep = []
    for r in range(world):
        ids, seq, dXi, dXh = inputs[r]
        t_ids = torch.from_numpy(ids).cuda()
        t_seq = torch.from_numpy(np.ascontiguousarray(seq)).cuda() if seq is not None else None
        t_dXi, t_dXh = torch.from_numpy(dXi).cuda(), torch.from_numpy(dXh).cuda()
        keep += [t_ids, t_seq, t_dXi, t_dXh]
        bs = _lib.Batch()
        bs.batch, bs.seq_len = B, L
        bs.item_id = bs.likes_level = bs.views_level = t_ids.data_ptr()
        bs.item_seq = t_seq.data_ptr() if t_seq is not None else None
        bs.idx_dtype, bs.seq_dtype = _lib.IDX_I64, _lib.IDX_I64
        _lib.check(lib.fbn_shard_index(C.byref(plans[r]), C.byref(bs), _lib.ptr(sws[r]), sws[r].numel(), st), "fbn_shard_index")
        _lib.check(lib.fbn_shard_local_sum(C.byref(plans[r]), C.byref(bs), _lib.ptr(t_dXi), _lib.ptr(t_dXh), _lib.ptr(sws[r]),
                                           sws[r].numel(), st), "fbn_shard_local_sum")
    out = []
    for o in range(world):
        sq = torch.zeros(1, device="cuda")
        if lazy:
            _lib.check(lib.fbn_shard_merge(C.byref(plans[o]), _lib.ptr(sws[o]), sws[o].numel(), None, None, _lib.ptr(sq), st))
            out.append(dict(sq=sq))
        else:
            g = torch.full((R, D), 7.0, device="cuda")          # untouched rows must be left alone (flag = 0)
            touched = torch.full((R,), 5, dtype=torch.int32, device="cuda")
            _lib.check(lib.fbn_shard_merge(C.byref(plans[o]), _lib.ptr(sws[o]), sws[o].numel(), _lib.ptr(g), _lib.ptr(touched), _lib.ptr(sq), st))
            out.append(dict(g=g, touched=touched, sq=sq))
    torch.cuda.synchronize()
    stats = []
    for o in range(world):
        h = (C.c_int32 * 24)()
        _lib.check(lib.fbn_shard_stats(C.byref(plans[o]), _lib.ptr(sws[o]), sws[o].numel(), h, st))
        stats.append(list(h))
    return dict(R=R, inputs=inputs, out=out, stats=stats, plans=plans, sws=sws, xchg=xchg, keep=keep)


@pytest.mark.parametrize("world,V,B,L,id_dist", [(1, 5000, 300, 20, "zipf"), (2, 91718, 1024, 20, "zipf"), (3, 1000, 777, 20, "uniform"),
                                                 (8, 91718, 512, 20, "uniform"), (8, 64, 256, 5, "uniform"), (4, 3001, 100, 0, "zipf"),
                                                 (2, 5000, 4096, 20, "zipf")])     # last: hot rows (> 256 occurrences, chunked sums)
@pytest.mark.parametrize("integer_grads", [True, False])
def test_shard_exchange_vs_oracle(env, world, V, B, L, id_dist, integer_grads):
    res = _virtual_exchange(env, world, V, B, L, seed=500 + world, id_dist=id_dist, lazy=False, integer_grads=integer_grads)
    R = res["R"]
    dense = np.zeros((V, D), dtype=np.float64)
    for ids, seq, dXi, dXh in res["inputs"]:
        dense += sorc.table_grad_dense(ids, seq, dXi, dXh, V)
    for o in range(world):
        want = sorc.slice_of(dense, o, world)
        want_touched = np.zeros(R, dtype=bool)
        for ids, seq, _, _ in res["inputs"]:
            allids = np.concatenate([ids] + ([seq.reshape(-1)] if seq is not None else []))
            mine = allids[(allids % world == o) & (allids != 0)]
            want_touched[mine // world] = True
        got_t = res["out"][o]["touched"].cpu().numpy()
        assert set(np.unique(got_t)) <= {0, 1}
        assert np.array_equal(got_t.astype(bool), want_touched), f"owner {o}: touched row set differs"
        g = res["out"][o]["g"].cpu().numpy()
        assert np.all(g[~want_touched] == 7.0), "untouched rows were written"
        if integer_grads:
            assert np.array_equal(g[want_touched].astype(np.float64), want[want_touched]), f"owner {o}: routing error"
        else:   # fp32 sums in a fixed order vs the float64 scatter-add: north_star's 1e-5 for fp32 gradients
            assert rel_err(g[want_touched], want[want_touched]) <= 1e-5
        sq = float(res["out"][o]["sq"].item())
        assert abs(sq - (want[want_touched] ** 2).sum()) <= 1e-5 * max(1.0, sq)
        h = res["stats"][o]
        assert h[22] == 0, "merge overflow flagged"
        assert h[21] == int(want_touched.sum()), "unique merged rows"
    # run-to-run bitwise determinism
    res2 = _virtual_exchange(env, world, V, B, L, seed=500 + world, id_dist=id_dist, lazy=False, integer_grads=integer_grads)
    for o in range(world):
        assert torch.equal(res["out"][o]["g"], res2["out"][o]["g"])
        assert torch.equal(res["out"][o]["sq"], res2["out"][o]["sq"])


def test_shard_merge_overflow_is_flagged(env):
    _lib, lib = env["_lib"], env["lib"]
    res = _virtual_exchange(env, 2, 5000, 256, 20, seed=9, id_dist="uniform", lazy=True)
    p = res["plans"][0]
    T = res["stats"][0][20]
    assert T > 16
    small = _lib.ShardPlan()
    for f, _t in _lib.ShardPlan._fields_:
        setattr(small, f, getattr(p, f))
    small.merge_cap = T // 2
    sws = torch.zeros(lib.fbn_shard_ws_bytes(small.cap, small.merge_cap, 2, small.shard_rows), dtype=torch.uint8, device="cuda")
    sq = torch.zeros(1, device="cuda")
    # the merge reads the peers' exchange blocks only, so it can run on a fresh scratch block carved for the smaller capacity
    _lib.check(lib.fbn_shard_merge(C.byref(small), _lib.ptr(sws), sws.numel(), None, None, _lib.ptr(sq), _lib.stream_ptr()))
    h = (C.c_int32 * 24)()
    _lib.check(lib.fbn_shard_stats(C.byref(small), _lib.ptr(sws), sws.numel(), h, _lib.stream_ptr()))
    assert h[22] == 1 and h[20] == T // 2


def _sharded_model(env, precision="fp32", dropout=0.0):
    from ctr_recommendation_b200 import build_model
    m = build_model({"precision": precision, "dropout": dropout, "table_sharding": "row", "shard_rank": 0, "shard_world": 1},
                    {"embedding_dim": 128})
    env["load_weights"](m, synth.make_weights(seed=7))
    return m.cuda().train()


def _plain_model(env, precision="fp32", dropout=0.0):
    from ctr_recommendation_b200 import build_model
    m = build_model({"precision": precision, "dropout": dropout}, {"embedding_dim": 128})
    env["load_weights"](m, synth.make_weights(seed=7))
    return m.cuda().train()


@pytest.mark.parametrize("graph,precision", [(False, "fp32"), (True, "fp32"), (True, "f16x3")])
def test_sharded_trainstep_matches_replicated_bitwise(env, graph, precision):
    from ctr_recommendation_b200 import FusedAdam
    from ctr_recommendation_b200.engine import ShardedTrainStep, TrainStep
    B, steps = 512, 3
    a, b = _plain_model(env, precision), _sharded_model(env, precision)
    oa, ob = FusedAdam(a, lr=1e-3, weight_decay=1e-5), FusedAdam(b, lr=1e-3, weight_decay=1e-5)
    ea = TrainStep(a, oa, B, 20, idx_dtype=torch.float64, graph=graph)
    eb = ShardedTrainStep(b, ob, B, 20, idx_dtype=torch.float64, graph=graph)
    for s in range(steps):
        batch, y = synth.make_batch(seed=900 + s, batch=B, id_dist="zipf", index_dtype=np.float64)
        tb = {k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in batch.items() if k != "user_id"}
        ty = torch.from_numpy(y).cuda()
        la = ea(tb, ty).clone()
        lb = eb(tb, ty).clone()
        assert torch.equal(la, lb)
    torch.cuda.synchronize()
    sa, sb = a.state_dict(), b.state_dict()
    for k in sa:
        assert torch.equal(sa[k], sb[k]), f"{k} differs between the replicated and the row-sharded step"
    assert torch.equal(oa._m_item, ob._m_item) and torch.equal(oa._v_item, ob._v_item)
    st = b._shard.stats()
    assert st["overflow"] == 0 and st["T"] == st["U"] == st["Um"] > 0


def test_sharded_lazy_adam_vs_oracle(env):
    from ctr_recommendation_b200 import FusedAdam
    from ctr_recommendation_b200.engine import ShardedTrainStep, TrainStep
    B = 384
    a, b = _plain_model(env), _sharded_model(env)
    oa = FusedAdam(a, lr=2e-3, weight_decay=1e-5)
    ob = FusedAdam(b, lr=2e-3, weight_decay=1e-5)
    ea = TrainStep(a, oa, B, 20, idx_dtype=torch.float64, graph=False)       # supplies the table gradient of step 1
    eb = ShardedTrainStep(b, ob, B, 20, idx_dtype=torch.float64, graph=True, lazy=True)
    w0 = b.item_emb.weight.detach().cpu().numpy().copy()
    batch, y = synth.make_batch(seed=1234, batch=B, id_dist="uniform", index_dtype=np.float64)
    tb = {k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in batch.items() if k != "user_id"}
    ty = torch.from_numpy(y).cuda()
    eb(tb, ty)
    ea(tb, ty)
    torch.cuda.synchronize()
    grad = a._item_grad.cpu().numpy()
    touched = a._row_touched.cpu().numpy() > 0
    coef = float(oa._clip[1].item())
    zeros = np.zeros_like(w0)
    p, m, v = sorc.lazy_adam_rows(w0, zeros, zeros, grad, touched, lr=2e-3, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=1e-5,
                                  step=1, coef=coef)
    got = b.item_emb.weight.detach().cpu().numpy()
    assert np.array_equal(got[~touched], w0[~touched]), "lazy Adam moved an untouched row"
    assert rel_err(got[touched], p[touched]) <= 1e-5
    assert rel_err(ob._m_item.cpu().numpy()[touched], m[touched]) <= 1e-5
    assert rel_err(ob._v_item.cpu().numpy()[touched], v[touched]) <= 1e-5
    assert np.all(ob._m_item.cpu().numpy()[~touched] == 0)


def test_sharded_scorer_matches_replicated(env):
    """Inference (Prediction.py loop body) over a row-sharded table == the replicated model, bit for bit."""
    from ctr_recommendation_b200.engine import Scorer
    a, b = _plain_model(env), _sharded_model(env)
    a.eval(); b.eval()
    B = 1000
    sa, sb = Scorer(a, B, 20, idx_dtype=torch.int64), Scorer(b, B, 20, idx_dtype=torch.int64)
    for s in range(2):
        batch, _ = synth.make_batch(seed=60 + s, batch=B, id_dist="zipf", index_dtype=np.int64)
        dev = env["to_dev"]({k: v for k, v in batch.items() if k != "user_id"})
        assert torch.equal(sa(dev), sb(dev))


@pytest.mark.parametrize("world", [1, 3])
def test_shard_exchange_matches_specified_summation_order_bitwise(env, world):
    """The CUDA path documents its summation order (source order; > 256 occurrences: 256-chunks relative to the row's start, chunk
    sums in chunk order; ranks in rank order).  oracle.table_grad_ordered restates exactly that order in fp32, so the comparison
    is bit for bit -- on a Zipf batch whose head rows take the chunked path."""
    V, B, L = 3000, 1500, 20
    res = _virtual_exchange(env, world, V, B, L, seed=77, id_dist="zipf", lazy=False)
    total = None
    for ids, seq, dXi, dXh in res["inputs"]:
        g = sorc.table_grad_ordered(ids, seq, dXi, dXh, V)
        total = g if total is None else (total + g).astype(np.float32)       # contributions are added in rank order
    occ = np.zeros(V, dtype=np.int64)
    for ids, seq, _, _ in res["inputs"]:
        occ = np.maximum(occ, np.bincount(np.r_[ids, seq.reshape(-1)], minlength=V))
    assert occ[1:].max() > 256, "the batch must contain a hot row"
    for o in range(world):
        want = sorc.slice_of(total, o, world)
        touched = res["out"][o]["touched"].cpu().numpy().astype(bool)
        got = res["out"][o]["g"].cpu().numpy()
        assert np.array_equal(got[touched], want[touched]), f"owner {o}: summation order differs from the specification"


def test_replicated_table_gradient_matches_specified_order_bitwise(env):
    """Same statement for the replicated table (embbwd.cu shares segsum.cuh): item_emb.grad of a Zipf batch == the oracle's
    ordered fp32 sums of the dX rows the kernels themselves produced."""
    model = env["make_model"](train=True)
    B = 2048
    batch, labels = synth.make_batch(seed=88, batch=B, id_dist="zipf", index_dtype=np.float64)
    y = model(env["to_dev"](batch))
    torch.nn.BCELoss()(y, torch.from_numpy(labels).cuda()).backward()
    dXi = model.workspace_view("dXitem", (B, D)).cpu().numpy()
    dXh = model.workspace_view("dXhist", (B, D)).cpu().numpy()
    want = sorc.table_grad_ordered(batch["item_id"], batch["item_seq"], dXi, dXh, model.item_emb.weight.shape[0])
    got = model.item_emb.weight.grad.cpu().numpy()
    occ = np.bincount(np.r_[batch["item_id"].astype(np.int64), batch["item_seq"].reshape(-1)], minlength=want.shape[0])
    assert occ[1:].max() > 256
    assert np.array_equal(got, want)
