"""2-rank data-parallel equivalence check (run: torchrun --nproc-per-node 2 tools/dp_check.py).

Rank r trains on its shard through engine.TrainStep (NCCL all-reduce of the gradients).  Rank 0 then replays the same
global batches on ONE model the way the DataParallel-equivalence oracle prescribes: per-shard forward /
backward with per-shard BatchNorm statistics, shard losses weighted by B_r/B, gradients summed, one clip + Adam -- and the
weights must agree; all ranks must hold identical replicas."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from ctr_recommendation_b200 import FusedAdam, build_model, clip_grad_norm_  # noqa: E402
from ctr_recommendation_b200 import dist as fdist  # noqa: E402
from ctr_recommendation_b200.engine import TrainStep  # noqa: E402
from oracle import synth  # noqa: E402


def make(precision):
    m = build_model({"precision": precision, "dropout": 0.0}, {"embedding_dim": 128})
    m.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in synth.make_weights(7).items()})
    return m.cuda().train()


def main():
    rank, local, world = fdist.init_from_env()
    torch.cuda.set_device(local)
    precision = os.environ.get("FBN_PRECISION", "f16x3")
    # global batch sizes: even split, uneven tail (rank shards of different size: loss weights B_r / B), and a tail so small that
    # torch's scatter chunking leaves the last rank WITHOUT rows (it joins the collectives with zero gradients)
    sizes = [1024, 1023, world - 1] if world > 1 else [1024, 1023]
    steps = len(sizes)
    model = make(precision)
    fdist.broadcast_parameters(model)
    opt = FusedAdam(model, lr=1e-3, weight_decay=1e-5)
    sched = torch.optim.lr_scheduler.OneCycleLR(opt, max_lr=1e-2, total_steps=20)
    engines = {}

    def engine(rows, n):
        if (rows, n) not in engines:
            engines[(rows, n)] = TrainStep(model, opt, rows, 20, idx_dtype=torch.float64, global_batch=n,
                                           overlap={"0": False, "1": "partial", "full": "full", "wgrad": "wgrad"}[os.environ.get("DP_OVERLAP", "1")])
        return engines[(rows, n)]
    batches = [synth.make_batch(seed=700 + s, batch=B, id_dist="zipf", index_dtype=np.float64, edge_cases=B >= 8)
               for s, B in enumerate(sizes)]
    grads_dp = []
    for b, y in batches:
        tb = {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in b.items() if k != "user_id"}
        shard, ys, w = fdist.shard_batch(tb, torch.from_numpy(y), rank, world)
        if ys.shape[0] == 0:
            next(iter(engines.values())).step_empty()
        else:
            engine(ys.shape[0], y.shape[0])({k: v.pin_memory() for k, v in shard.items()}, ys.pin_memory())
        sched.step()
        if len(grads_dp) == 0:      # the all-reduced gradients of the first step (identical weights on both sides)
            torch.cuda.synchronize()
            grads_dp.append((model._gflat.clone(), model._item_grad.clone()))
    torch.cuda.synchronize()
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    # replicas identical?
    for k, v in sd.items():
        ref = v.clone()
        dist.broadcast(ref, 0)
        if "running" in k or "num_batches" in k:
            continue        # BatchNorm running statistics are per replica (DataParallel keeps replica 0's; rank 0 saves)
        assert torch.equal(ref, v), f"rank {rank}: replica differs in {k}"
    if rank == 0:
        single = make(precision)
        sopt = FusedAdam(single, lr=1e-3, weight_decay=1e-5)
        ssched = torch.optim.lr_scheduler.OneCycleLR(sopt, max_lr=1e-2, total_steps=20)
        single._dense_table_grad = True
        first, grad_worst = True, 0.0
        for b, y in batches:
            tb = {k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in b.items() if k != "user_id"}
            ty = torch.from_numpy(y).cuda()
            gsum, isum = None, None
            for r in range(world):
                shard, ys, w = fdist.shard_batch(tb, ty, r, world)
                if ys.shape[0] == 0:
                    continue
                sopt.zero_grad()
                out = single(shard)
                (torch.nn.BCELoss()(out, ys) * w).backward()
                g, it = single._gflat.clone(), single._item_grad.clone()
                gsum = g if gsum is None else gsum + g
                isum = it if isum is None else isum + it
            if first:
                first = False
                for name, a, b_ in (("dense", grads_dp[0][0], gsum), ("item_emb", grads_dp[0][1], isum)):
                    rel = (a.double() - b_.double()).abs().max().item() / max(b_.abs().max().item(), 1e-30)
                    assert rel <= 1e-5, ("first-step gradient", name, rel)
                    grad_worst = max(grad_worst, rel)
            single._gflat.copy_(gsum)
            single._item_grad.copy_(isum)
            single._grad_sumsq[0] = (gsum.double() ** 2).sum().float()
            single._grad_sumsq[1] = (isum.double() ** 2).sum().float()
            clip_grad_norm_(single, 10.0)
            sopt.step()
            ssched.step()
        worst = 0.0
        for k, v in single.state_dict().items():
            # running statistics follow replica 0's shard in DataParallel (and here on rank 0); the single-process emulation
            # above updates them once per shard, so they are not comparable
            if "num_batches" in k or "running" in k or k in ("mlp.0.bias", "mlp.4.bias"):
                continue
            # After 3 Adam steps two exact implementations differ by whole lr-sized steps on the elements whose gradient is at
            # rounding-noise level (Adam normalises; DESIGN section 2, finding 1): the weights are held to a MEAN difference of a
            # small fraction of one step (max lr over the run = 1e-2 * cycle ~ 5e-4 here), the first-step gradients to 1e-5.
            d = (v.double() - sd[k].double()).abs()
            rel = d.mean().item()
            worst = max(worst, rel)
            assert rel <= 0.05 * 5e-4, (k, rel)
        print(f"dp_check OK (collectives {os.environ.get('DP_OVERLAP', '1')}): {world} ranks, global batches {sizes} (uneven + empty-shard tails), replicas identical, first-step all-reduced gradients within {grad_worst:.2e} (rel) of the "
              f"single-process DataParallel emulation, worst mean |weight diff| after {steps} steps {worst:.2e}")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
