"""N-rank check of the row-sharded item table (run: torchrun --nproc-per-node N tools/shard_check.py [--rows V] [--batch B]).

1. Parity: the same global batches are trained (a) with replicated tables + NCCL all-reduce (engine.TrainStep, already
   checked against the single-process DataParallel emulation by tools/dp_check.py) and (b) with the table row-sharded over
   the ranks (engine.ShardedTrainStep: remote gather over NVLink, owner-side gradient merge).  The assembled sharded table,
   its Adam moments and every dense parameter must agree -- bit for bit at 2 ranks (a two-term sum has one rounding), to
   1e-5 beyond (NCCL's ring order and the rank-order merge round differently; Adam then amplifies that a little).
2. Scale: a (--rows, default 25 M per 2 ranks) table that only exists sharded, lazy row Adam; prints samples/s."""
import argparse
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ctr_recommendation_b200 import FusedAdam, build_model  # noqa: E402
from ctr_recommendation_b200 import dist as fdist  # noqa: E402
from ctr_recommendation_b200 import sharded  # noqa: E402
from ctr_recommendation_b200.engine import ShardedTrainStep, TrainStep  # noqa: E402
from oracle import synth  # noqa: E402


def weights():
    return {k: torch.from_numpy(np.array(v)) for k, v in synth.make_weights(7).items()}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=0)
    ap.add_argument("--batch", type=int, default=16384)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--precision", default="f16x3")
    args = ap.parse_args()
    rank, local, world = fdist.init_from_env()
    torch.cuda.set_device(local)
    W = weights()

    # ---- 1. parity against the replicated-table data-parallel step
    rep = build_model({"precision": args.precision, "dropout": 0.0}, {"embedding_dim": 128})
    rep.load_state_dict(W)
    rep = rep.cuda().train()
    fdist.broadcast_parameters(rep)
    shd = build_model({"precision": args.precision, "dropout": 0.0, "table_sharding": "row"}, {"embedding_dim": 128})
    Ws = dict(W)
    Ws["item_emb.weight"] = sharded.slice_of_full(W["item_emb.weight"], rank, world)
    shd.load_state_dict(Ws)
    shd = shd.cuda().train()
    o1, o2 = FusedAdam(rep, lr=1e-3, weight_decay=1e-5), FusedAdam(shd, lr=1e-3, weight_decay=1e-5)
    B, steps = 1024, 3
    per = B // world
    e1 = TrainStep(rep, o1, per, 20, idx_dtype=torch.float64)
    e2 = ShardedTrainStep(shd, o2, per, 20, idx_dtype=torch.float64)
    for s in range(steps):
        b, y = synth.make_batch(seed=700 + s, batch=B, id_dist="zipf", index_dtype=np.float64)
        tb = {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in b.items() if k != "user_id"}
        sh, ys, w = fdist.shard_batch(tb, torch.from_numpy(y), rank, world)
        sh = {k: v.cuda() for k, v in sh.items()}
        l1 = e1(sh, ys.cuda()).clone()
        l2 = e2(sh, ys.cuda()).clone()
        assert torch.equal(l1, l2), (s, l1, l2)
    torch.cuda.synchronize()
    full = sharded.gather_full_table(shd)
    m_full = sharded.full_from_slices(_gather(o2._m_item, world), full.shape[0])
    v_full = sharded.full_from_slices(_gather(o2._v_item, world), full.shape[0])
    worst = 0.0
    for name, a, b_ in [("item_emb.weight", rep.item_emb.weight.data, full), ("exp_avg", o1._m_item, m_full), ("exp_avg_sq", o1._v_item, v_full)]:
        d = (a.double() - b_.double()).abs().max().item() / max(a.abs().max().item(), 1e-30)
        worst = max(worst, d)
        if world == 2:
            assert torch.equal(a, b_), f"{name}: sharded != replicated (max rel {d:.3e})"
        else:
            assert d <= 1e-5, (name, d)
    sd1, sd2 = rep.state_dict(), shd.state_dict()
    for k in sd1:
        if k == "item_emb.weight" or "running" in k or "num_batches" in k:
            continue
        d = (sd1[k].double() - sd2[k].double()).abs().max().item() / max(sd1[k].abs().max().item(), 1e-30)
        if world > 2 and k in ("mlp.0.bias", "mlp.4.bias"):
            # a Linear bias in front of BatchNorm has an identically zero gradient in exact arithmetic: what Adam normalises is
            # rounding noise, so beyond the bit-identical 2-rank case these two move by rounding-determined +-lr steps
            continue
        worst = max(worst, d)
        assert d <= (0.0 if world == 2 else 1e-5), (k, d)
    st = shd._shard.stats()
    assert st["overflow"] == 0
    if rank == 0:
        print(f"shard_check parity OK: {world} ranks, sharded == replicated (worst rel diff {worst:.2e}); last step: U={st['U']} "
              f"owner_start={st['owner_start']} T={st['T']} Um={st['Um']}", flush=True)

    # ---- 2. a table that only exists sharded
    del e1, e2, rep, shd, o1, o2
    torch.cuda.empty_cache()
    V = args.rows or 12_500_000 * world
    big = build_model({"precision": args.precision, "table_sharding": "row", "item_rows": V}, {"embedding_dim": 128}).cuda().train()
    fdist.broadcast_parameters(big)
    opt = FusedAdam(big, lr=1e-3, weight_decay=1e-5)
    Bp = args.batch
    eng = ShardedTrainStep(big, opt, Bp, 20, idx_dtype=torch.int64, lazy=True, merge_cap=4 * Bp * 21)
    g = torch.Generator().manual_seed(100 + rank)
    pool = []
    for i in range(4):
        b, y = synth.make_batch(seed=2025 + 1000 * rank + i, batch=Bp, id_dist="uniform", index_dtype=np.int64, edge_cases=False)
        b.pop("user_id")
        tb = {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in b.items()}
        tb["item_id"] = torch.randint(1, V, (Bp,), generator=g)
        seq = torch.randint(1, V, (Bp, 20), generator=g)
        tb["item_seq"] = torch.where(tb["item_seq"] != 0, seq, torch.zeros_like(seq))
        pool.append(({k: v.cuda() for k, v in tb.items()}, torch.from_numpy(y).cuda()))
    for k in range(5):
        eng(*pool[k % 4])
    torch.cuda.synchronize()
    dist.barrier() if world > 1 else None
    e0, e1_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(args.steps):
        eng(*pool[k % 4])
    e1_.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1_)], device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    st = big._shard.stats()
    loss = float(eng.loss.item())
    assert st["overflow"] == 0 and np.isfinite(loss)
    if rank == 0:
        sps = Bp * world * args.steps / (ms.item() / 1e3)
        print(f"shard_check scale: V={V} rows over {world} ranks ({V / world * 512 * 3 / 2**30:.1f} GiB p/m/v per GPU), per-GPU batch {Bp}, "
              f"lazy row Adam: {ms.item() / args.steps:.3f} ms/step, {sps / 1e6:.2f} M samples/s; U={st['U']} T={st['T']} Um={st['Um']} "
              f"loss {loss:.4f}", flush=True)
    # ---- 3. where the time goes: the stages of the sharded step, eager, CUDA events (all ranks run them at the same time,
    #         so the remote loads of every rank share the NVLink fabric as they do in the real step)
    import ctypes as C
    from ctr_recommendation_b200 import _lib
    lib = _lib.load()
    eng._use_graph = False
    flush = torch.empty(160 << 20, dtype=torch.uint8, device="cuda")
    b0, y0 = pool[0]
    eng.inp.load(b0, y0)
    eng._write_hyper()

    def timed(fn, reps=5):
        tot = 0.0
        for _ in range(reps):
            flush.fill_(1)
            if world > 1:
                dist.barrier()
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record()
            fn()
            t1.record()
            torch.cuda.synchronize()
            tot += t0.elapsed_time(t1)
        return tot / reps
    P = big._params_struct()
    msf = C.c_float(0)
    if world > 1:
        dist.barrier()
    _lib.check(lib.fbn_time_stage(C.byref(P), C.byref(eng._bs), _lib.ptr(eng.ws), eng.ws.numel(), b"embed", _lib.ptr(flush), flush.numel(),
                                  5, C.byref(msf), _lib.stream_ptr()), "fbn_time_stage")
    t_gather = float(msf.value)
    eng._fwd_bwd()
    if world > 1:
        dist.barrier()
    t_merge = timed(eng._merge)
    t_update = timed(eng._update)
    nvalid = int((b0["item_seq"] != 0).sum().item()) + Bp
    remote = nvalid * 512 * (world - 1) / world
    st = big._shard.stats()
    pulled = st["T"] * 512 * (world - 1) / world
    res = torch.tensor([t_gather, t_merge, t_update], device="cuda")
    if world > 1:
        dist.all_reduce(res, op=dist.ReduceOp.MAX)
    if rank == 0:
        tg, tm, tu = (float(x) for x in res)
        print(f"shard_check stages (max over ranks, B={Bp}/GPU): gather+SENET fwd {tg * 1e3:.0f} us ({remote / 1e6:.0f} MB of remote rows/GPU"
              f" = {remote / tg / 1e6:.0f} GB/s over NVLink, {(nvalid * 512 + Bp * 12000) / tg / 1e6:.0f} GB/s total); owner-side merge "
              f"{tm * 1e3:.0f} us ({pulled / 1e6:.0f} MB of partial rows pulled = {pulled / tm / 1e6:.0f} GB/s); clip + lazy row Adam "
              f"({st['Um']} rows, {st['Um'] * 3072 / 1e6:.0f} MB) + dense Adam {tu * 1e3:.0f} us", flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _gather(t, world):
    if world == 1:
        return [t]
    parts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(parts, t.contiguous())
    return parts


if __name__ == "__main__":
    main()
