"""TrainStep / Scorer: the reference loop bodies as CUDA-graph replays.

``TrainStep`` is the body of the reference's training loop (src/train_fibinet.py:113-122)

    zero_grad -> forward -> BCELoss -> backward -> clip_grad_norm_(10) -> Adam.step -> (scheduler.step)

captured once into CUDA graphs over static device buffers: a step is then one host->device copy of the batch,
one 32-byte copy of the optimizer hyper-parameters (so any torch LR scheduler, e.g. the reference's
OneCycleLR which rewrites lr AND betas[0], keeps working unchanged) and a graph launch -- no per-kernel launch
cost, no host synchronisation (the loss stays on the device until the caller reads it).

With more than one rank (process per GPU) the step is split into two graphs around the NCCL all-reduce of the
gradients:  [forward+loss+backward]  ->  all-reduce  ->  [clip+Adam].

``Scorer`` is the inference loop body (src/Prediction.py:108-113): eval-mode forward as one graph replay.
"""
from __future__ import annotations

import ctypes as C
import math

import torch

from . import _lib
from . import dist as fdist
from .model import MM_FiBiNET, D, _IDX_DTYPES
from .optim import FusedAdagrad, FusedAdam


class _StaticBatch:
    """Device-resident input buffers with fixed addresses (what the captured kernels read)."""

    def __init__(self, B, L, idx_dtype, seq_dtype, device, with_mm=True, with_labels=True):
        self.B, self.L = B, L
        z = dict(device=device)
        self.t = {"item_id": torch.zeros(B, dtype=idx_dtype, **z), "likes_level": torch.zeros(B, dtype=idx_dtype, **z),
                  "views_level": torch.zeros(B, dtype=idx_dtype, **z)}
        if L > 0:
            self.t["item_seq"] = torch.zeros(B, L, dtype=seq_dtype, **z)
        if with_mm:
            self.t["item_emb_d128"] = torch.zeros(B, D, dtype=torch.float32, **z)
        self.labels = torch.zeros(B, dtype=torch.float32, **z) if with_labels else None

    def load(self, batch: dict, labels=None):
        for k, dst in self.t.items():
            src = batch[k]
            if src.shape != dst.shape:
                src = src.reshape(dst.shape)
            dst.copy_(src, non_blocking=True)     # dtype conversion (if any) happens in the copy
        if labels is not None and self.labels is not None:
            self.labels.copy_(labels.reshape(-1), non_blocking=True)

    def nbytes(self):
        n = sum(t.numel() * t.element_size() for t in self.t.values())
        return n + (self.labels.numel() * 4 if self.labels is not None else 0)


class TrainStep:
    HYPER_SLOTS = 16

    def __init__(self, model: MM_FiBiNET, optimizer: FusedAdam, batch_size: int, seq_len: int = 20, idx_dtype=torch.float64,
                 seq_dtype=torch.int64, max_norm: float | None = 10.0, use_mm_table: bool = False, graph: bool = True,
                 global_batch: int | None = None, overlap=None, reserve_sms: int = 16, phased_single: bool = False):
        if not isinstance(model, MM_FiBiNET) or not isinstance(optimizer, (FusedAdam, FusedAdagrad)):
            raise TypeError("TrainStep needs a ctr_recommendation_b200 MM_FiBiNET and its FusedAdam / FusedAdagrad")
        self.model, self.opt, self.max_norm = model, optimizer, max_norm
        self._want_phased = bool(phased_single)     # one rank: run the phased schedule anyway (tests / A-B runs), without collectives
        self.lib = _lib.load()
        model._ensure_flat()
        optimizer._ensure_state()
        dev = model._flat.device
        self.dev = dev
        self.world = torch.distributed.get_world_size() if (torch.distributed.is_available() and torch.distributed.is_initialized()) else 1
        if self.world > 1 and model._shard is None:
            model._dense_table_grad = True
        # overlap the gradient all-reduces with the weight-gradient GEMMs (4 graphs + 3 async collectives per step instead of 2 + 2
        # blocking ones): worth it once the leaf phase is long enough to hide a 47 MB all-reduce
        if overlap is None:
            # measured at 65536 rows per GPU: from 4 ranks on the 47 MB table all-reduce is worth hiding behind
            # the MLP-1 weight gradient; at 2 ranks the blocking schedule (which keeps every leaf beside the data-gradient chain) wins
            overlap = "wgrad" if (batch_size >= 8192 and self.world >= 4) else False
        elif overlap is True:
            overlap = "wgrad"
        if overlap not in (False, "partial", "full", "wgrad"):
            raise ValueError("overlap must be None, False, True, 'partial', 'full' or 'wgrad'")
        self.overlap = overlap
        self.reserve_sms = int(reserve_sms)
        self._phased_single = bool(overlap) and self.world == 1 and batch_size > 0 and getattr(self, "_want_phased", False)
        self._pending = []
        self.inp = _StaticBatch(batch_size, seq_len, idx_dtype, seq_dtype, dev, with_mm=not use_mm_table)
        if use_mm_table and model._mm_table is None:
            raise ValueError("use_mm_table=True needs model.attach_mm_table(...)")
        self.B, self.L = batch_size, seq_len
        self.ws = model._workspace(batch_size, seq_len)
        self._alloc_table_grad()
        self.prob = torch.zeros(batch_size, dtype=torch.float32, device=dev)
        self.dprob = torch.zeros(batch_size, dtype=torch.float32, device=dev)
        self.loss = torch.zeros(1, dtype=torch.float32, device=dev)
        self._bce_scratch = torch.zeros(self.lib.fbn_bce_scratch_bytes(), dtype=torch.uint8, device=dev)   # partials + arrival counter
        # dropout stream position: ONE counter per model, shared by every engine that drives it (a tail-batch engine must not
        # replay the masks the main engine used at the same local step); model.set_dropout_counter(n) resumes a stream
        self.step_counter = model._dropout_counter(dev)
        self.hyper_dev = torch.zeros(12, dtype=torch.float32, device=dev)
        # the host never waits for the device between steps, so the pinned source of the hyper-parameter copy of step t must
        # not be rewritten while that copy is still queued: a ring of pinned slots, each guarded by the event of its last copy
        self._hyper_ring = [torch.zeros(12, dtype=torch.float32).pin_memory() for _ in range(self.HYPER_SLOTS)]
        self._hyper_ev = [None] * self.HYPER_SLOTS
        self._hyper_pos = 0
        self.sumsq_scratch = torch.zeros(max(self.lib.fbn_sumsq_partial_floats(max(self._table_grad_numel(), model._flat.numel())), 16),
                                         dtype=torch.float32, device=dev)
        self._side = torch.cuda.Stream(device=dev)
        # input prefetch: the next batch is copied host->device on a copy stream into a staging set while the current step
        # computes (what the reference's DataLoader workers + pin_memory would give); the step then takes it with a D2D copy
        self._copy = torch.cuda.Stream(device=dev)
        self._stage = None
        self._h2d_done = torch.cuda.Event()
        self._stage_free = torch.cuda.Event()
        self._prefetched = False
        # DataParallel semantics (reference src/train_fibinet.py:69-70,115): the loss is the mean over the GLOBAL batch, so this
        # rank's local mean is weighted by B_r / B (scatter chunking gives the last ranks fewer rows on a tail batch)
        self.loss_weight = (batch_size / float(global_batch)) if global_batch else 1.0 / self.world
        self._bs = self._batch_struct()
        self._graphs = None
        self._use_graph = graph
        self._steps = 0
        self.kernels_per_step = 0

    # ------------------------------------------------------------------
    def _alloc_table_grad(self):
        model, dev = self.model, self.dev
        if model._shard is not None:
            raise TypeError("row-sharded item table: use ShardedTrainStep")
        rows = model.item_emb.weight.shape[0]
        if model._item_grad is None or model._item_grad.device != dev:
            model._item_grad = torch.zeros(rows, D, dtype=torch.float32, device=dev)
            model._row_touched = torch.zeros(rows, dtype=torch.int32, device=dev)
            model._grad_sumsq = torch.zeros(2, dtype=torch.float32, device=dev)

    def _table_grad_numel(self):
        return self.model._item_grad.numel()

    def _stages(self):
        """The step as a list of (device work, collective that follows it or None); consecutive stages without a collective
        between them are captured into one CUDA graph."""
        if self.world == 1:
            if self.overlap and self._phased_single:      # test hook: the phased backward without collectives
                if self.overlap == "full":
                    return [(self._fwd_chain, None), (self._leaf1, None), (self._leaf2, None), (self._update, None)]
                if self.overlap == "wgrad":
                    return [(self._fwd_chain, None), (self._leaf1, None), (self._update, None)]
                return [(self._fwd_chain, None), (self._leaf2, None), (self._update, None)]
            return [(self._fwd_bwd, None), (self._update, None)]
        if not self.overlap:
            return [(self._fwd_bwd, self._allreduce), (self._update, None)]
        if self.overlap == "full":
            # [forward + loss + data-gradient chain + table rows] -> all-reduce(table gradient) || [MLP-1 weight gradient]
            #   -> all-reduce(MLP-1 bucket) || [all other leaf gradients] -> all-reduce(rest), wait for the three -> [norms + clip + Adam]
            return [(self._fwd_chain, self._ar_table), (self._leaf1, self._ar_bucket1), (self._leaf2, self._ar_rest_wait),
                    (self._update, None)]
        if self.overlap == "wgrad":
            # every leaf except MLP-1's stays interleaved with the data-gradient chain; the table all-reduce (47 MB) and the small
            # dense bucket then run beside the MLP-1 weight gradient (0.5 ms of tensor work), only the MLP-1 bucket is exposed
            #   [forward + loss + chain + small leaves + table rows] -> all-reduce(table), all-reduce(rest) || [MLP-1 leaf]
            #   -> all-reduce(MLP-1 bucket), wait for the three -> [update]
            return [(self._fwd_chain, self._ar_table_rest), (self._leaf1, self._ar_bucket1_wait), (self._update, None)]
        # "partial" (default): the MLP-1 weight gradient keeps running beside the data-gradient chain on the side stream (that
        # concurrency is worth more than hiding its bucket); the 47 MB table all-reduce overlaps the remaining leaf gradients
        #   [forward + loss + chain + MLP-1 leaf + table rows] -> all-reduce(table) || [other leaves] -> all-reduce(dense), wait -> [update]
        return [(self._fwd_chain, self._ar_table), (self._leaf2, self._ar_dense_wait), (self._update, None)]

    def _batch_struct(self):
        t = self.inp.t
        bs = _lib.Batch()
        bs.batch, bs.seq_len = self.B, self.L
        bs.item_id, bs.likes_level, bs.views_level = t["item_id"].data_ptr(), t["likes_level"].data_ptr(), t["views_level"].data_ptr()
        bs.idx_dtype = _IDX_DTYPES[t["item_id"].dtype]
        bs.seq_dtype = _lib.IDX_I64
        if self.L > 0:
            bs.item_seq = t["item_seq"].data_ptr()
            bs.seq_dtype = _IDX_DTYPES[t["item_seq"].dtype]
        if "item_emb_d128" in t:
            bs.item_mm = t["item_emb_d128"].data_ptr()
        else:
            bs.mm_table = self.model._mm_table.data_ptr()
        return bs

    def _fwd_bwd(self):
        m, lib = self.model, self.lib
        P, G = m._params_struct(), m._grads_struct()
        ws = self.ws
        # the occurrence index of the embedding backward only needs the batch ids: build it on a side stream while the
        # forward pass runs (fork / join, also inside graph capture)
        cur = torch.cuda.current_stream()
        self._side.wait_stream(cur)
        with torch.cuda.stream(self._side):
            _lib.check(lib.fbn_embed_index(C.byref(P), C.byref(self._bs), _lib.ptr(ws), ws.numel(), _lib.ptr(m._row_touched),
                                           _lib.stream_ptr()), "fbn_embed_index")
        st = _lib.stream_ptr()
        _lib.check(lib.fbn_forward(C.byref(P), C.byref(self._bs), _lib.ptr(ws), ws.numel(), 1, m.dropout_p, None, None, m._seed, 0,
                                   _lib.ptr(self.step_counter), _lib.ptr(self.prob), st), "fbn_forward")
        _lib.check(lib.fbn_bce_loss_ws(_lib.ptr(self.prob), _lib.ptr(self.inp.labels), self.B, self.loss_weight, _lib.ptr(self.loss),
                                       _lib.ptr(self.dprob), _lib.ptr(self._bce_scratch), self._bce_scratch.numel(), st), "fbn_bce_loss_ws")
        dense_table = m._dense_table_grad
        cur.wait_stream(self._side)            # join: the index is ready before the table gradient is summed
        _lib.check(lib.fbn_backward(C.byref(P), C.byref(self._bs), _lib.ptr(ws), ws.numel(), 1, m.dropout_p, _lib.ptr(self.dprob),
                                    C.byref(G), _lib.ptr(m._gflat), m._gflat.numel(), _lib.ptr(m._item_grad),
                                    _lib.ptr(m._row_touched), 1 if dense_table else 0, 1, _lib.ptr(m._grad_sumsq), st), "fbn_backward")

    def _fwd_chain(self):
        """Forward, loss and the data-gradient chain of the backward pass down to the table gradient (no leaf gradients)."""
        m, lib = self.model, self.lib
        P, G = m._params_struct(), m._grads_struct()
        ws = self.ws
        cur = torch.cuda.current_stream()
        self._side.wait_stream(cur)
        with torch.cuda.stream(self._side):
            _lib.check(lib.fbn_embed_index(C.byref(P), C.byref(self._bs), _lib.ptr(ws), ws.numel(), _lib.ptr(m._row_touched),
                                           _lib.stream_ptr()), "fbn_embed_index")
        st = _lib.stream_ptr()
        _lib.check(lib.fbn_forward(C.byref(P), C.byref(self._bs), _lib.ptr(ws), ws.numel(), 1, m.dropout_p, None, None, m._seed, 0,
                                   _lib.ptr(self.step_counter), _lib.ptr(self.prob), st), "fbn_forward")
        _lib.check(lib.fbn_bce_loss_ws(_lib.ptr(self.prob), _lib.ptr(self.inp.labels), self.B, self.loss_weight, _lib.ptr(self.loss),
                                       _lib.ptr(self.dprob), _lib.ptr(self._bce_scratch), self._bce_scratch.numel(), st), "fbn_bce_loss_ws")
        cur.wait_stream(self._side)
        self._backward_phase({"full": _lib.BWD_CHAIN, "wgrad": _lib.BWD_CHAIN | _lib.BWD_LEAF2}.get(self.overlap, _lib.BWD_CHAIN | _lib.BWD_LEAF1))

    def _backward_phase(self, phases, reserve=0):
        m, lib = self.model, self.lib
        P, G = m._params_struct(), m._grads_struct()
        if reserve:      # leave SMs to the collective running beside the weight-gradient GEMMs (read at launch / capture time)
            _lib.check(lib.fbn_set_option(b"tc_reserve_sms", int(reserve)), "fbn_set_option")
        try:
            _lib.check(lib.fbn_backward_phase(C.byref(P), C.byref(self._bs), _lib.ptr(self.ws), self.ws.numel(), 1, m.dropout_p,
                                              _lib.ptr(self.dprob), C.byref(G), _lib.ptr(m._item_grad), _lib.ptr(m._row_touched),
                                              1 if m._dense_table_grad else 0, 1, _lib.ptr(m._grad_sumsq), int(phases),
                                              _lib.stream_ptr()), "fbn_backward_phase")
        finally:
            if reserve:
                lib.fbn_set_option(b"tc_reserve_sms", 0)

    def _leaf1(self):
        self._backward_phase(_lib.BWD_LEAF1, self.reserve_sms)

    def _leaf2(self):
        self._backward_phase(_lib.BWD_LEAF2, self.reserve_sms)

    def _ar_table(self):
        import torch.distributed as dist
        self._pending = [dist.all_reduce(self.model._item_grad, op=dist.ReduceOp.SUM, async_op=True)]

    def _ar_bucket1(self):
        import torch.distributed as dist
        m = self.model
        self._pending.append(dist.all_reduce(m._gflat[m._bucket1_offset():], op=dist.ReduceOp.SUM, async_op=True))

    def _ar_table_rest(self):
        import torch.distributed as dist
        m = self.model
        self._pending = [dist.all_reduce(m._item_grad, op=dist.ReduceOp.SUM, async_op=True),
                         dist.all_reduce(m._gflat[:m._bucket1_offset()], op=dist.ReduceOp.SUM, async_op=True)]

    def _ar_bucket1_wait(self):
        import torch.distributed as dist
        m = self.model
        self._pending.append(dist.all_reduce(m._gflat[m._bucket1_offset():], op=dist.ReduceOp.SUM, async_op=True))
        for w in self._pending:
            w.wait()
        self._pending = []

    def _ar_dense_wait(self):
        import torch.distributed as dist
        self._pending.append(dist.all_reduce(self.model._gflat, op=dist.ReduceOp.SUM, async_op=True))
        for w in self._pending:
            w.wait()
        self._pending = []

    def _ar_rest_wait(self):
        import torch.distributed as dist
        m = self.model
        self._pending.append(dist.all_reduce(m._gflat[:m._bucket1_offset()], op=dist.ReduceOp.SUM, async_op=True))
        for w in self._pending:
            w.wait()                 # the compute stream waits for the collectives; the host does not
        self._pending = []

    def _update(self):
        m, o, lib, st = self.model, self.opt, self.lib, _lib.stream_ptr()
        if self.world > 1 or self._phased_single:   # norm of the (all-reduced) dense gradients: the phased backward leaves it to the host
            _lib.check(lib.fbn_sumsq(_lib.ptr(m._gflat), m._gflat.numel(), _lib.ptr(self.sumsq_scratch), _lib.ptr(m._grad_sumsq), st))
        if self.world > 1:                          # and of the all-reduced table gradient
            _lib.check(lib.fbn_sumsq(_lib.ptr(m._item_grad), m._item_grad.numel(), _lib.ptr(self.sumsq_scratch),
                                     fdist.C_ptr_offset(m._grad_sumsq, 1), st))
        clip = None
        if self.max_norm is not None:
            _lib.check(lib.fbn_clip_coef(_lib.ptr(m._grad_sumsq), 2, float(self.max_norm), _lib.ptr(o._clip), st), "fbn_clip_coef")
            clip = _lib.ptr(o._clip)
        o._launch_update(lib, clip, _lib.ptr(self.hyper_dev), st, m._dense_table_grad)      # Adam / AdamW / Adagrad row + dense update
        self.step_counter += 1

    def check_ids(self):
        """IndexError if any batch since the last call carried an id outside its table (synchronises; see MM_FiBiNET.check_ids)."""
        self.model.check_ids(self.ws, self.B, self.L)

    def _capture(self):
        m = self.model
        keep = [b.clone() for b in (m.mlp[1].running_mean, m.mlp[1].running_var, m.mlp[5].running_mean, m.mlp[5].running_var)]
        side = torch.cuda.Stream(device=self.dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):          # warm-up outside capture (lazy inits: func attributes, workspaces)
            self._fwd_bwd()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        for b, k in zip((m.mlp[1].running_mean, m.mlp[1].running_var, m.mlp[5].running_mean, m.mlp[5].running_var), keep):
            b.copy_(k)                         # the warm-up forward must not count as a training step
        # group the stages into graphs: a new graph starts after every collective
        groups, cur_fns = [], []
        for fn, coll in self._stages():
            cur_fns.append(fn)
            if coll is not None:
                groups.append((cur_fns, coll))
                cur_fns = []
        if cur_fns:
            groups.append((cur_fns, None))
        n0 = self.lib.fbn_launch_count()
        graphs = []
        for fns, coll in groups:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for fn in fns:
                    fn()
            graphs.append((g, coll))
        self._graphs = graphs
        self.kernels_per_step = int(self.lib.fbn_launch_count() - n0)   # library kernels inside the captured step

    def _write_hyper(self):
        slot = self._hyper_pos % self.HYPER_SLOTS
        self._hyper_pos += 1
        if self._hyper_ev[slot] is not None:
            self._hyper_ev[slot].synchronize()      # the copy that last read this slot has executed
        h = self._hyper_ring[slot]
        self.opt._fill_hyper(h)                     # advances the optimizer's step count
        self.hyper_dev.copy_(h, non_blocking=True)
        ev = self._hyper_ev[slot] or torch.cuda.Event()
        ev.record()
        self._hyper_ev[slot] = ev

    # ------------------------------------------------------------------
    def prefetch(self, batch: dict, labels: torch.Tensor):
        """Start the host->device copy of the NEXT batch on the copy stream; the following ``step()`` (no arguments)
        consumes it.  Overlaps the input transfer with the current step's kernels."""
        if self._stage is None:
            t = self.inp.t
            self._stage = _StaticBatch(self.B, self.L, t["item_id"].dtype, t["item_seq"].dtype if "item_seq" in t else torch.int64,
                                       self.dev, with_mm="item_emb_d128" in t)
        self._copy.wait_event(self._stage_free)          # the previous staged batch has been taken
        with torch.cuda.stream(self._copy):
            self._stage.load(batch, labels)
            self._h2d_done.record(self._copy)
        self._prefetched = True

    def __call__(self, batch: dict | None = None, labels: torch.Tensor | None = None) -> torch.Tensor:
        """One optimizer step on ``batch`` (host-pinned or device tensors), or on the batch handed to ``prefetch()`` when
        called without arguments.  Returns the mean BCE loss of this rank's shard as a 1-element device tensor (read it
        with .item() only when you need it)."""
        m = self.model
        if not m.training:
            raise RuntimeError("TrainStep needs model.train()")
        if batch is None:
            if not self._prefetched:
                raise RuntimeError("TrainStep() without a batch needs a preceding prefetch()")
            cur = torch.cuda.current_stream()
            cur.wait_event(self._h2d_done)
            self.inp.load(self._stage.t, self._stage.labels)      # device-to-device, ~15 us
            self._stage_free.record(cur)
            self._prefetched = False
        else:
            self.inp.load(batch, labels)
        self._write_hyper()
        if self._use_graph and self._graphs is None:
            self._capture()
        if self._use_graph:
            for g, coll in self._graphs:
                g.replay()
                if coll is not None:
                    coll()
        else:
            for fn, coll in self._stages():
                fn()
                if coll is not None:
                    coll()
        m.mlp[1].num_batches_tracked += 1
        m.mlp[5].num_batches_tracked += 1
        self._steps += 1
        return self.loss

    def _allreduce(self):
        # the same collective sequence as this engine's overlapped schedule (a rank without rows must match it call for call)
        import torch.distributed as dist
        m = self.model
        dist.all_reduce(m._item_grad, op=dist.ReduceOp.SUM)
        if self.overlap == "full":
            off = m._bucket1_offset()
            dist.all_reduce(m._gflat[off:], op=dist.ReduceOp.SUM)
            dist.all_reduce(m._gflat[:off], op=dist.ReduceOp.SUM)
        elif self.overlap == "wgrad":
            off = m._bucket1_offset()
            dist.all_reduce(m._gflat[:off], op=dist.ReduceOp.SUM)
            dist.all_reduce(m._gflat[off:], op=dist.ReduceOp.SUM)
        else:
            dist.all_reduce(m._gflat, op=dist.ReduceOp.SUM)

    def step_empty(self) -> torch.Tensor:
        """This rank received no rows of the global batch (torch's scatter chunking leaves the last ranks empty when
        n < world * ceil(n / world), e.g. n = 17 on 8 ranks): contribute zero gradients, join the collectives and apply the
        same update as every other rank.  Eager launches (a tail batch happens once per epoch)."""
        m = self.model
        if self.world == 1:
            raise RuntimeError("step_empty() only makes sense with more than one rank")
        m._gflat.zero_()
        m._item_grad.zero_()
        self._write_hyper()
        self._allreduce()
        self._update()
        self._steps += 1
        self.loss.zero_()
        return self.loss


class ShardedTrainStep(TrainStep):
    """TrainStep for a row-sharded item table (``feature_map={"table_sharding": "row", ...}``, sharded.py).

    Per step and rank:  [index || forward, loss, backward, local segment sums -> exchange block]
                        -> all-reduce of the dense gradients (also the barrier that publishes the exchange blocks)
                        -> [owner-side merge of the N partial lists over NVLink, slice sum of squares]
                        -> all-reduce of that scalar (clip; also releases the exchange blocks)
                        -> [clip coefficient, Adam on the slice (dense-exact, or ``lazy=True``: touched rows only), dense Adam]
                        -> barrier (the next forward reads the updated rows remotely).
    With one rank the same kernels run back to back in one graph (the single-GPU tests cover the whole path)."""

    def __init__(self, model: MM_FiBiNET, optimizer: FusedAdam, batch_size: int, seq_len: int = 20, idx_dtype=torch.float64,
                 seq_dtype=torch.int64, max_norm: float | None = 10.0, use_mm_table: bool = False, graph: bool = True,
                 lazy: bool = False, merge_cap: int | None = None, global_batch: int | None = None):
        if model._shard is None:
            raise TypeError("ShardedTrainStep needs a model built with table_sharding='row'")
        if not isinstance(optimizer, FusedAdam):
            raise TypeError("the row-sharded table is updated by (lazy or dense-exact) Adam: pass a FusedAdam")
        self.lazy = bool(lazy)
        self._merge_cap = merge_cap
        super().__init__(model, optimizer, batch_size, seq_len, idx_dtype, seq_dtype, max_norm, use_mm_table, graph, global_batch)
        st = model._shard
        if st.world != self.world:
            raise ValueError(f"the table is sharded over {st.world} ranks but the process group has {self.world}")
        self.plan = st.ensure_exchange(batch_size * (1 + seq_len), self.dev, merge_cap)
        self._sws = st.sws          # this engine's private scratch (a model may drive several engines, e.g. a tail batch)
        off = lambda name: self.lib.fbn_workspace_offset(self.B, self.L, 1, name.encode())
        self._dX = (C.c_void_p(self.ws.data_ptr() + off("dXitem")), C.c_void_p(self.ws.data_ptr() + off("dXhist")))
        self._sq_local = torch.zeros(1, dtype=torch.float32, device=self.dev)
        self._bar = torch.zeros(1, dtype=torch.float32, device=self.dev)

    def _alloc_table_grad(self):
        m, dev = self.model, self.dev
        R = m._shard.shard_rows
        if m._grad_sumsq is None or m._grad_sumsq.device != dev:
            m._grad_sumsq = torch.zeros(2, dtype=torch.float32, device=dev)
        # shared by every engine of the model (a captured graph keeps raw pointers: never re-allocate what exists)
        if not self.lazy and (m._item_grad is None or m._item_grad.device != dev or m._item_grad.shape[0] != R):
            m._item_grad = torch.zeros(R, D, dtype=torch.float32, device=dev)
            m._row_touched = torch.zeros(R, dtype=torch.int32, device=dev)

    def _table_grad_numel(self):
        return 0

    def _stages(self):
        if self.world == 1:
            return [(self._fwd_bwd, None), (self._merge, None), (self._update, None)]
        return [(self._fwd_bwd, self._allreduce), (self._merge, self._allreduce_sumsq), (self._update, self._barrier)]

    def _fwd_bwd(self):
        m, lib, st = self.model, self.lib, self.model._shard
        P, G = m._params_struct(), m._grads_struct()
        ws, sws = self.ws, self._sws
        cur = torch.cuda.current_stream()
        self._side.wait_stream(cur)
        with torch.cuda.stream(self._side):     # ids only: runs beside the forward pass
            _lib.check(lib.fbn_shard_index(C.byref(self.plan), C.byref(self._bs), _lib.ptr(sws), sws.numel(), _lib.stream_ptr()),
                       "fbn_shard_index")
        s = _lib.stream_ptr()
        _lib.check(lib.fbn_forward(C.byref(P), C.byref(self._bs), _lib.ptr(ws), ws.numel(), 1, m.dropout_p, None, None, m._seed, 0,
                                   _lib.ptr(self.step_counter), _lib.ptr(self.prob), s), "fbn_forward")
        _lib.check(lib.fbn_bce_loss_ws(_lib.ptr(self.prob), _lib.ptr(self.inp.labels), self.B, self.loss_weight, _lib.ptr(self.loss),
                                       _lib.ptr(self.dprob), _lib.ptr(self._bce_scratch), self._bce_scratch.numel(), s), "fbn_bce_loss_ws")
        _lib.check(lib.fbn_backward(C.byref(P), C.byref(self._bs), _lib.ptr(ws), ws.numel(), 1, m.dropout_p, _lib.ptr(self.dprob),
                                    C.byref(G), _lib.ptr(m._gflat), m._gflat.numel(), None, None, 0, 1, _lib.ptr(m._grad_sumsq), s),
                   "fbn_backward")
        cur.wait_stream(self._side)
        _lib.check(lib.fbn_shard_local_sum(C.byref(self.plan), C.byref(self._bs), self._dX[0], self._dX[1], _lib.ptr(sws), sws.numel(), s),
                   "fbn_shard_local_sum")

    def _merge(self):
        m, lib, st, s = self.model, self.lib, self.model._shard, _lib.stream_ptr()
        if self.world > 1:     # norm of the all-reduced dense gradients
            _lib.check(lib.fbn_sumsq(_lib.ptr(m._gflat), m._gflat.numel(), _lib.ptr(self.sumsq_scratch), _lib.ptr(m._grad_sumsq), s))
        dense = None if self.lazy else m._item_grad
        touched = None if self.lazy else m._row_touched
        _lib.check(lib.fbn_shard_merge(C.byref(self.plan), _lib.ptr(self._sws), self._sws.numel(), _lib.ptr(dense), _lib.ptr(touched),
                                       _lib.ptr(self._sq_local), s), "fbn_shard_merge")

    def _update(self):
        m, o, lib, st, s = self.model, self.opt, self.lib, self.model._shard, _lib.stream_ptr()
        m._grad_sumsq[1:2].copy_(self._sq_local)
        clip = None
        if self.max_norm is not None:
            _lib.check(lib.fbn_clip_coef(_lib.ptr(m._grad_sumsq), 2, float(self.max_norm), _lib.ptr(o._clip), s), "fbn_clip_coef")
            clip = _lib.ptr(o._clip)
        w = m.item_emb.weight.data
        if self.lazy:
            _lib.check(lib.fbn_shard_adam_rows(C.byref(self.plan), _lib.ptr(self._sws), self._sws.numel(), _lib.ptr(w), _lib.ptr(o._m_item),
                                               _lib.ptr(o._v_item), clip, None, _lib.ptr(self.hyper_dev), s), "fbn_shard_adam_rows")
        else:
            _lib.check(lib.fbn_adam_table(_lib.ptr(w), _lib.ptr(o._m_item), _lib.ptr(o._v_item), _lib.ptr(m._item_grad),
                                          _lib.ptr(m._row_touched), w.shape[0], clip, None, _lib.ptr(self.hyper_dev), s), "fbn_adam_table")
        _lib.check(lib.fbn_adam_dense(_lib.ptr(m._flat), _lib.ptr(o._m_flat), _lib.ptr(o._v_flat), _lib.ptr(m._gflat), m._flat.numel(),
                                      clip, None, _lib.ptr(self.hyper_dev), s), "fbn_adam_dense")
        self.step_counter += 1

    def _allreduce(self):
        import torch.distributed as dist
        dist.all_reduce(self.model._gflat, op=dist.ReduceOp.SUM)

    def check(self) -> dict:
        """Synchronising health check: raises if the owner-side merge lists overflowed ``merge_cap`` in the last step (partial
        gradient rows would have been dropped); returns the exchange statistics.  Call it every few hundred steps."""
        st = self.model._shard.stats(self.plan, self._sws)
        if st["overflow"]:
            raise RuntimeError(f"row-sharded table: {st['T']} partial rows addressed to rank {self.plan.rank} exceed merge_cap="
                               f"{self.plan.merge_cap}; build ShardedTrainStep with a larger merge_cap")
        return st

    def _allreduce_sumsq(self):
        import torch.distributed as dist
        dist.all_reduce(self._sq_local, op=dist.ReduceOp.SUM)

    def _barrier(self):
        import torch.distributed as dist
        dist.all_reduce(self._bar, op=dist.ReduceOp.SUM)


class Scorer:
    """Eval-mode forward (the Prediction.py loop body) over static buffers, one CUDA-graph replay per batch."""

    def __init__(self, model: MM_FiBiNET, batch_size: int, seq_len: int = 20, idx_dtype=torch.int64, seq_dtype=torch.int64,
                 use_mm_table: bool = False, graph: bool = True):
        self.model, self.lib = model, _lib.load()
        model._ensure_flat()
        dev = model._flat.device
        self.inp = _StaticBatch(batch_size, seq_len, idx_dtype, seq_dtype, dev, with_mm=not use_mm_table, with_labels=False)
        self.B, self.L = batch_size, seq_len
        self.ws = model._workspace(batch_size, seq_len)
        self.prob = torch.zeros(batch_size, dtype=torch.float32, device=dev)
        self._bs = TrainStep._batch_struct(self)
        self._graph, self._use_graph = None, graph
        self.kernels_per_step = 0
        self.dev = dev
        # input prefetch (same scheme as TrainStep.prefetch): the next batch crosses PCIe on a copy stream while this one is scored
        self._copy = torch.cuda.Stream(device=dev)
        self._stage = None
        self._h2d_done = torch.cuda.Event()
        self._stage_free = torch.cuda.Event()
        self._prefetched = False

    def prefetch(self, batch: dict):
        """Start the host->device copy of the NEXT batch on the copy stream; the following call without arguments scores it."""
        if self._stage is None:
            t = self.inp.t
            self._stage = _StaticBatch(self.B, self.L, t["item_id"].dtype, t["item_seq"].dtype if "item_seq" in t else torch.int64,
                                       self.dev, with_mm="item_emb_d128" in t, with_labels=False)
        self._copy.wait_event(self._stage_free)
        with torch.cuda.stream(self._copy):
            self._stage.load(batch)
            self._h2d_done.record(self._copy)
        self._prefetched = True

    def check_ids(self):
        """IndexError if any batch scored since the last call carried an id outside its table (synchronises)."""
        self.model.check_ids(self.ws, self.B, self.L)

    def _fwd(self):
        m = self.model
        P = m._params_struct()
        _lib.check(self.lib.fbn_forward(C.byref(P), C.byref(self._bs), _lib.ptr(self.ws), self.ws.numel(), 0, 0.0, None, None, 0, 0, None,
                                        _lib.ptr(self.prob), _lib.stream_ptr()), "fbn_forward")

    def __call__(self, batch: dict | None = None) -> torch.Tensor:
        if self.model.training:
            raise RuntimeError("Scorer needs model.eval()")
        if batch is None:
            if not self._prefetched:
                raise RuntimeError("Scorer() without a batch needs a preceding prefetch()")
            cur = torch.cuda.current_stream()
            cur.wait_event(self._h2d_done)
            self.inp.load(self._stage.t)              # device-to-device
            self._stage_free.record(cur)
            self._prefetched = False
        else:
            self.inp.load(batch)
        if not self._use_graph:
            self._fwd()
            return self.prob
        if self._graph is None:
            self._fwd()
            torch.cuda.synchronize()
            self._graph = torch.cuda.CUDAGraph()
            n0 = self.lib.fbn_launch_count()
            with torch.cuda.graph(self._graph):
                self._fwd()
            self.kernels_per_step = int(self.lib.fbn_launch_count() - n0)
        self._graph.replay()
        return self.prob
