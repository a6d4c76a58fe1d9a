"""Row-sharded item table (BASELINE config 5, SURVEY 8e "row-sharded tables").

The reference replicates ``item_emb`` on every GPU (nn.DataParallel, src/train_fibinet.py:69-70).  Here global row g
lives on rank ``g % N`` at local row ``g // N``; every rank maps every other rank's slice and gradient-exchange block
into its address space with CUDA IPC, so

  * the forward gather kernel reads remote rows straight out of the owner's HBM over NVLink (no all-to-all, no staging),
  * the backward lets every owner PULL the partial gradient rows of its slice from all peers and add them in rank order
    (fbn_shard_merge), followed by dense-exact Adam on the slice or lazy row Adam on the touched rows only.

This module holds the host logic: the partition functions, the IPC exchange of device pointers (torch.distributed is only
the courier for the 64-byte handles) and the assembly of the full table for checkpoints.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional

import torch
import torch.distributed as dist

from . import _lib

D = 128


# ---------------------------------------------------------------------------------------------- partition (pure functions)
def shard_rows(item_rows: int, world: int) -> int:
    """Rows of every rank's slice: ceil(V / N)."""
    return -(-int(item_rows) // int(world))


def owner_of(row, world: int):
    return row % world


def local_row(row, world: int):
    return row // world


def global_row(rank: int, local, world: int):
    return local * world + rank


def slice_of_full(full: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    """The (shard_rows, D) slice rank ``rank`` owns of a full (V, D) table (zero rows pad the tail)."""
    R = shard_rows(full.shape[0], world)
    out = torch.zeros(R, full.shape[1], dtype=full.dtype, device=full.device)
    part = full[rank::world]
    out[: part.shape[0]] = part
    return out


def full_from_slices(slices: List[torch.Tensor], item_rows: int) -> torch.Tensor:
    """Inverse of slice_of_full: interleave the N slices back into the (V, D) table."""
    world = len(slices)
    full = torch.empty(item_rows, slices[0].shape[1], dtype=slices[0].dtype, device=slices[0].device)
    for r, s in enumerate(slices):
        n = len(range(r, item_rows, world))
        full[r::world] = s[:n]
    return full


# ---------------------------------------------------------------------------------------------- IPC exchange
class PeerMap:
    """Exchanges device pointers of same-purpose buffers between the ranks of one node.

    ``share(t)`` is a collective: every rank passes its own tensor and gets the list of N device pointers (its own entry is
    the local pointer, the others are peer mappings opened through fbn_ipc_open).  Allocation handles are opened once and
    cached, because one cudaMalloc block of torch's caching allocator may hold several shared tensors."""

    def __init__(self, rank: int, world: int, group=None):
        self.rank, self.world, self.group = rank, world, group
        self._opened: Dict[bytes, int] = {}
        self._keep = []

    def share(self, t: torch.Tensor) -> List[int]:
        if self.world == 1:
            return [t.data_ptr()]
        lib = _lib.load()
        handle = (C.c_ubyte * 64)()
        off = C.c_int64(0)
        _lib.check(lib.fbn_ipc_export(C.c_void_p(t.data_ptr()), C.cast(handle, C.c_void_p), C.byref(off)), "fbn_ipc_export")
        mine = (bytes(handle), int(off.value), self.rank)
        everyone = [None] * self.world
        dist.all_gather_object(everyone, mine, group=self.group)
        ptrs = []
        for h, o, r in everyone:
            if r == self.rank:
                ptrs.append(t.data_ptr())
                continue
            base = self._opened.get(h)
            if base is None:
                out = C.c_void_p()
                buf = (C.c_ubyte * 64).from_buffer_copy(h)
                _lib.check(lib.fbn_ipc_open(C.cast(buf, C.c_void_p), C.byref(out)), "fbn_ipc_open")
                base = int(out.value)
                self._opened[h] = base
            ptrs.append(base + o)
        self._keep.append(t)      # exported memory must outlive the peers' mappings
        return ptrs

    def close(self):
        lib = _lib.load()
        for base in self._opened.values():
            lib.fbn_ipc_close(C.c_void_p(base))
        self._opened.clear()
        self._keep.clear()


class ShardState:
    """Per-model state of a row-sharded item table: the plan handed to fbn_shard_* and the peer pointers of the slices."""

    def __init__(self, item_rows: int, rank: int, world: int):
        if not (1 <= world <= 16):
            raise ValueError("row sharding supports 1..16 ranks (one NVLink domain)")
        self.item_rows, self.rank, self.world = int(item_rows), int(rank), int(world)
        self.shard_rows = shard_rows(item_rows, world)
        self.peers = PeerMap(rank, world)
        self.table_ptrs: Optional[List[int]] = None
        self._table_key = None
        self.plan: Optional[_lib.ShardPlan] = None      # the most recently created exchange (what stats() reports on)
        self.xchg: Optional[torch.Tensor] = None
        self.sws: Optional[torch.Tensor] = None
        self.cap = 0
        self._exchanges = {}      # (cap, merge_cap) -> (plan, xchg, sws): one per engine size, never freed while peers map them

    def ensure_table(self, weight: torch.Tensor):
        """(Re-)exchange the slice pointers when the parameter storage moved (collective on first use / after .to())."""
        key = (weight.data_ptr(), weight.device)
        if self._table_key != key:
            if weight.shape[0] != self.shard_rows:
                raise RuntimeError(f"item_emb slice has {weight.shape[0]} rows, expected {self.shard_rows}")
            self.table_ptrs = self.peers.share(weight.data)
            self._table_key = key
        return self.table_ptrs

    def ensure_exchange(self, occurrences: int, device, merge_cap: Optional[int] = None):
        """Allocate the gradient exchange block and the private scratch for up to ``occurrences`` = B*(1+L) ids per rank
        (the maximum over ranks is used) and exchange the block pointers (collective)."""
        lib = _lib.load()
        cap = int(occurrences)
        if self.world > 1:
            t = torch.tensor([cap], dtype=torch.int64, device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            cap = int(t.item())
        if merge_cap is None:
            merge_cap = self.world * cap
        merge_cap = max(1, min(int(merge_cap), self.world * cap))
        hit = self._exchanges.get((cap, merge_cap))
        if hit is None:
            xchg = torch.zeros(lib.fbn_shard_xchg_bytes(cap), dtype=torch.uint8, device=device)
            sws = torch.zeros(lib.fbn_shard_ws_bytes(cap, merge_cap, self.world, self.shard_rows), dtype=torch.uint8, device=device)
            ptrs = self.peers.share(xchg)
            plan = _lib.ShardPlan()
            plan.n_shards, plan.rank = self.world, self.rank
            plan.item_rows, plan.shard_rows, plan.cap, plan.merge_cap = self.item_rows, self.shard_rows, cap, merge_cap
            for r, p in enumerate(ptrs):
                plan.xchg[r] = p
            hit = self._exchanges[(cap, merge_cap)] = (plan, xchg, sws)
        self.plan, self.xchg, self.sws = hit
        self.cap = cap
        return self.plan

    def stats(self, plan=None, sws=None) -> dict:
        """{U, owner_start, T, Um, overflow} of the last step of an exchange (default: the most recently created one).
        Synchronises the stream; tests / diagnostics / the periodic overflow check of ShardedTrainStep.check()."""
        plan = self.plan if plan is None else plan
        sws = self.sws if sws is None else sws
        out = (C.c_int32 * 24)()
        _lib.check(_lib.load().fbn_shard_stats(C.byref(plan), _lib.ptr(sws), sws.numel(), out, _lib.stream_ptr()), "fbn_shard_stats")
        return {"U": out[0], "owner_start": [out[1 + o] for o in range(self.world + 1)], "T": out[20], "Um": out[21],
                "overflow": out[22]}


def gather_full_table(model) -> torch.Tensor:
    """Assemble the full (V,128) item table from the slices (every rank gets it) -- what a 28-key reference checkpoint
    stores under ``item_emb.weight``."""
    st: ShardState = model._shard
    w = model.item_emb.weight.data
    if st.world == 1:
        return full_from_slices([w], st.item_rows)
    parts = [torch.empty_like(w) for _ in range(st.world)]
    dist.all_gather(parts, w.contiguous())
    return full_from_slices(parts, st.item_rows)
