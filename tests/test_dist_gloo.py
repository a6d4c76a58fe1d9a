"""CPU, world_size 2, gloo: the data-parallel host logic (sharding, gradient all-reduce + weights, prediction gather)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _FakeModel:
    def __init__(self, rank):
        g = torch.Generator().manual_seed(10 + rank)
        self._gflat = torch.randn(1000, generator=g)
        self._item_grad = torch.randn(50, 128, generator=g)
        self._grad_sumsq = torch.zeros(2)
        self._dense_table_grad = False

    def named_parameters(self):
        return []

    def named_buffers(self):
        return []


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from ctr_recommendation_b200 import dist as fdist
    r, l, w = fdist.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    m = _FakeModel(rank)
    ref = [_FakeModel(i) for i in range(world)]
    fdist.broadcast_parameters(m)
    assert m._dense_table_grad is True
    # unequal shards: weights B_r / B as in the DataParallel-equivalence oracle
    n = 10
    lo, hi = fdist.shard_bounds(n, rank, world)
    weight = (hi - lo) / n
    fdist.sync_gradients(m, weight=weight)
    exp_flat = sum(((fdist.shard_bounds(n, i, world)[1] - fdist.shard_bounds(n, i, world)[0]) / n) * ref[i]._gflat for i in range(world))
    assert torch.allclose(m._gflat, exp_flat, atol=1e-6)
    batch = {"item_id": torch.arange(n), "item_seq": torch.arange(n * 3).reshape(n, 3)}
    shard, lab, wgt = fdist.shard_batch(batch, torch.arange(n).float(), rank, world)
    assert abs(wgt - weight) < 1e-12 and shard["item_seq"].shape == (hi - lo, 3)
    pred = fdist.gather_predictions(lab * 2)
    assert torch.equal(pred, torch.arange(n).float() * 2)            # rank order == original row order
    # a tail batch so small that scatter chunking leaves the last rank without rows: weights still sum to 1, the empty rank still
    # takes part in the prediction gather (src/train_fibinet.py drives TrainStep.step_empty / an empty tensor for it)
    n1 = world - 1
    b1 = {"item_id": torch.arange(n1)}
    sh1, lab1, w1 = fdist.shard_batch(b1, torch.arange(n1).float(), rank, world)
    wsum = torch.tensor([w1])
    dist.all_reduce(wsum)
    assert abs(wsum.item() - 1.0) < 1e-12 and (lab1.numel() == 0) == (rank == world - 1)
    assert torch.equal(fdist.gather_predictions(lab1 + 5), torch.arange(n1).float() + 5)
    # row-sharded table host logic: slices -> all_gather -> the full table again; a model built with table_sharding="row"
    # holds exactly its slice of the replicated initialisation
    from ctr_recommendation_b200 import build_model, sharded
    V = 1001
    full = torch.arange(V * 4, dtype=torch.float32).reshape(V, 4)
    mine = sharded.slice_of_full(full, rank, world)
    assert mine.shape[0] == sharded.shard_rows(V, world)
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine)
    assert torch.equal(sharded.full_from_slices(parts, V), full)
    torch.manual_seed(5)
    rep = build_model({"item_rows": V}, {"embedding_dim": 128})
    torch.manual_seed(5)
    shd = build_model({"table_sharding": "row", "item_rows": V}, {"embedding_dim": 128})     # rank / world from the process group
    assert (shd._shard.rank, shd._shard.world) == (rank, world)
    assert torch.equal(shd.item_emb.weight.data, sharded.slice_of_full(rep.item_emb.weight.data, rank, world))
    assert torch.equal(shd.mlp[0].weight.data, rep.mlp[0].weight.data)      # same RNG consumption as the replicated build

    class _M:
        pass
    holder = _M()
    holder._shard, holder.item_emb = shd._shard, shd.item_emb
    assert torch.equal(sharded.gather_full_table(holder), rep.item_emb.weight.data)
    if rank == 0:
        out.put("ok")
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29610 + (os.getpid() % 200)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert q.get(timeout=5) == "ok"


def test_shard_bounds_match_torch_chunking():
    from ctr_recommendation_b200.dist import shard_bounds
    for n in (1, 7, 8, 4096, 4097):
        for world in (1, 2, 3, 8):
            chunks = torch.arange(n).chunk(world)       # what DataParallel's scatter does on dim 0
            for r in range(world):
                lo, hi = shard_bounds(n, r, world)
                exp = chunks[r] if r < len(chunks) else torch.arange(0)
                assert hi - lo == exp.numel()
                if exp.numel():
                    assert lo == int(exp[0])
