"""Process-per-GPU data parallelism for the FiBiNET hot path (replaces the reference's single-process
nn.DataParallel, src/train_fibinet.py:69-70).

Semantics kept from DataParallel: the batch is split on dim 0, BatchNorm statistics
are per replica, the loss is the mean over the global batch (each rank's local mean weighted by
B_r / B), parameters stay replicated.  What changes: one process per GPU, no per-forward parameter
broadcast, gradients are summed with NCCL all-reduce over NVLink instead of reduce-to-GPU-0.

Collectives per step (replicated tables):
  * one all-reduce of the flat dense gradient buffer (1.54 M floats),
  * one all-reduce of the dense (rows,128) embedding-table gradient,
after which every rank recomputes the global gradient norm locally (identical on all ranks) and runs
the same deterministic fused Adam, so replicas stay bit-identical without any parameter traffic.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist

from . import _lib


def init_from_env(backend: str | None = None, nccl_max_ctas: int | None = None):
    """torchrun-style initialisation (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT).

    ``nccl_max_ctas``: hold NCCL to that many CTAs per collective (NCCL_MAX_CTAS, unless the user already set it).  Pass the
    ``reserve_sms`` of engine.TrainStep when its overlapped schedule is used: the gradient all-reduces then run BESIDE the
    weight-gradient GEMMs, a collective CTA cannot share an SM with a 225 KB-smem GEMM CTA, and the GEMMs leave exactly that many SMs
    free."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if nccl_max_ctas:
            os.environ.setdefault("NCCL_MAX_CTAS", str(int(nccl_max_ctas)))
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, rank=rank, world_size=world, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, local, world


def shard_bounds(n: int, rank: int, world: int):
    """Rows [lo, hi) of a global batch owned by `rank`: the same chunking as torch's scatter
    (ceil(n / world) per rank, last ranks may be short or empty)."""
    per = -(-n // world)
    lo = min(n, rank * per)
    return lo, min(n, lo + per)


def shard_batch(batch: dict, labels, rank: int, world: int):
    n = next(iter(batch.values())).shape[0]
    lo, hi = shard_bounds(n, rank, world)
    return {k: v[lo:hi] for k, v in batch.items()}, (None if labels is None else labels[lo:hi]), (hi - lo) / max(n, 1)


def broadcast_parameters(model, src: int = 0):
    """One-time replication at start-up (the reference re-broadcasts every forward).  Also switches the
    model to dense table gradients (every row written) so that the table all-reduce is well defined."""
    sharded = getattr(model, "_shard", None) is not None
    if not sharded:
        model._dense_table_grad = True
    if not (dist.is_available() and dist.is_initialized()):
        return
    for name, t in list(model.named_parameters()) + list(model.named_buffers()):
        if sharded and name == "item_emb.weight":
            continue        # every rank owns a different slice of a row-sharded table
        dist.broadcast(t.data, src)


def sync_gradients(model, weight: float | None = None):
    """Sum the gradients of all ranks (after scaling the local ones by `weight` = B_r / B, default
    1 / world) and refresh the gradient sum-of-squares used by the global clip.  No-op for world 1."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    world = dist.get_world_size()
    w = (1.0 / world) if weight is None else float(weight)
    flat, table = model._gflat, model._item_grad
    if w != 1.0:
        flat.mul_(w)
        table.mul_(w)
    h1 = dist.all_reduce(flat, op=dist.ReduceOp.SUM, async_op=True)
    h2 = dist.all_reduce(table, op=dist.ReduceOp.SUM, async_op=True)
    h1.wait()
    h2.wait()
    if flat.is_cuda:
        lib = _lib.load()
        n1 = lib.fbn_sumsq_partial_floats(flat.numel())
        n2 = lib.fbn_sumsq_partial_floats(table.numel())
        scratch = torch.empty(max(n1, n2), dtype=torch.float32, device=flat.device)
        st = _lib.stream_ptr()
        ss = model._grad_sumsq
        _lib.check(lib.fbn_sumsq(_lib.ptr(flat), flat.numel(), _lib.ptr(scratch), _lib.ptr(ss), st), "fbn_sumsq")
        _lib.check(lib.fbn_sumsq(_lib.ptr(table), table.numel(), _lib.ptr(scratch), C_ptr_offset(ss, 1), st), "fbn_sumsq")


def C_ptr_offset(t: torch.Tensor, idx: int):
    import ctypes
    return ctypes.c_void_p(t.data_ptr() + idx * t.element_size())


def gather_predictions(local: torch.Tensor) -> torch.Tensor:
    """Inference is embarrassingly parallel: concatenate per-rank predictions in rank order."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    sizes = [torch.zeros(1, dtype=torch.int64, device=local.device) for _ in range(dist.get_world_size())]
    dist.all_gather(sizes, torch.tensor([local.numel()], dtype=torch.int64, device=local.device))
    mx = int(max(s.item() for s in sizes))
    pad = torch.zeros(mx, dtype=local.dtype, device=local.device)
    pad[: local.numel()] = local
    outs = [torch.empty_like(pad) for _ in sizes]
    dist.all_gather(outs, pad)
    return torch.cat([o[: int(s.item())] for o, s in zip(outs, sizes)])
