"""Host-side mirror of the reference model interface (src/model_fibinet.py of the reference):

    build_model(feature_map, model_cfg) -> MM_FiBiNET ;  MM_FiBiNET.forward(batch_dict) -> (B,) probabilities
    SENetLayer(num_fields, reduction_ratio=3) ;  BilinearInteraction(input_dim, num_fields, bilinear_type="all")

Same constructor arguments, same 28-key state_dict, same forward contract -- but every FLOP and
byte of the forward/backward runs in libfibinet_b200.so (hand-written sm_100a CUDA) through the
C ABI of include/fibinet_b200.h.  torch supplies parameter storage, the autograd hook-up and the
CUDA stream; it computes nothing.  The sub-modules below only *own* parameters (so that
state_dict / load_state_dict / .to() behave exactly like the reference); their torch forward
methods are never called by MM_FiBiNET.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import torch
import torch.nn as nn

from . import _lib

D = 128
NUM_FIELDS = 6
ITEM_ROWS, USER_ROWS, CATE_ROWS = 91718, 20000, 11   # literals of reference src/model_fibinet.py:100-102

_IDX_DTYPES = {torch.int32: _lib.IDX_I32, torch.int64: _lib.IDX_I64, torch.float64: _lib.IDX_F64,
               torch.float32: _lib.IDX_F32}

# flat layout of the dense parameters: (C-ABI field, state_dict key)
_DENSE = [("cate_emb", "cate_emb.weight"), ("mm_w", "mm_proj.0.weight"), ("mm_b", "mm_proj.0.bias"),
          ("ln_g", "mm_proj.1.weight"), ("ln_b", "mm_proj.1.bias"),
          ("se_w1", "senet.excitation.0.weight"), ("se_b1", "senet.excitation.0.bias"),
          ("se_w2", "senet.excitation.2.weight"), ("se_b2", "senet.excitation.2.bias"),
          ("bil_w", None),
          ("w2", "mlp.4.weight"), ("b2", "mlp.4.bias"), ("bn2_g", "mlp.5.weight"), ("bn2_b", "mlp.5.bias"),
          ("w3", "mlp.8.weight"), ("b3", "mlp.8.bias"),
          # last: the MLP-1 bucket (89 % of the dense bytes).  Its gradients are complete as soon as the first leaf phase of the
          # backward pass is, and form one contiguous tail of the flat buffer -> one all-reduce that overlaps the other leaves
          ("w1", "mlp.0.weight"), ("b1", "mlp.0.bias"), ("bn1_g", "mlp.1.weight"), ("bn1_b", "mlp.1.bias")]
BUCKET1_FIRST = "w1"


def _require_cuda(t: torch.Tensor, what: str):
    if not t.is_cuda:
        raise RuntimeError(f"{what} is on {t.device}: ctr_recommendation_b200 has no CPU path "
                           "(hand-written sm_100a CUDA only). Move the model and the batch to a B200.")


class SENetLayer(nn.Module):
    """Squeeze-excitation over fields (reference src/model_fibinet.py:5-35); CUDA: fbn_senet_fwd/bwd."""

    def __init__(self, num_fields, reduction_ratio=3):
        super().__init__()
        reduced_size = max(1, num_fields // reduction_ratio)
        self.num_fields, self.reduced_size = num_fields, reduced_size
        self.excitation = nn.Sequential(nn.Linear(num_fields, reduced_size), nn.ReLU(),
                                        nn.Linear(reduced_size, num_fields), nn.Sigmoid())

    def forward(self, x):
        from .functional import senet
        return senet(x, self.excitation[0].weight, self.excitation[0].bias,
                     self.excitation[2].weight, self.excitation[2].bias)


class BilinearInteraction(nn.Module):
    """Pairwise bilinear interaction (reference src/model_fibinet.py:37-89); CUDA: fbn_bilinear_fwd/bwd.

    "all" / "each" as in the reference; "interaction" (one matrix per pair, FiBiNET paper) is an
    extension the reference rejects with ValueError.  FuxiCTR names field_all / field_each /
    field_interaction are accepted as aliases.
    """

    def __init__(self, input_dim, num_fields, bilinear_type="all"):
        super().__init__()
        if bilinear_type not in _lib.BILINEAR_TYPES:
            raise ValueError("bilinear_type must be 'all' or 'each'")
        self.bilinear_type = {0: "all", 1: "each", 2: "interaction"}[_lib.BILINEAR_TYPES[bilinear_type]]
        self.input_dim, self.num_fields = input_dim, num_fields
        if self.bilinear_type == "all":
            self.W = nn.Parameter(torch.empty(input_dim, input_dim))
            nn.init.xavier_normal_(self.W)
        else:
            n = num_fields - 1 if self.bilinear_type == "each" else num_fields * (num_fields - 1) // 2
            self.W_list = nn.ParameterList([nn.Parameter(torch.empty(input_dim, input_dim)) for _ in range(n)])
            for w in self.W_list:
                nn.init.xavier_normal_(w)

    def weights(self):
        return [self.W] if self.bilinear_type == "all" else list(self.W_list)

    def forward(self, x):
        from .functional import bilinear
        return bilinear(x, self.weights(), self.bilinear_type)


class _FibinetFn(torch.autograd.Function):
    """One autograd node for the whole network: forward = fbn_forward, backward = fbn_backward.

    Parameter gradients are written by the kernels straight into the model's flat gradient buffer
    and attached to ``param.grad`` (accumulating if a gradient is already present), so no autograd
    copy of the 47 MB table gradient is ever made.
    """

    @staticmethod
    def forward(ctx, model, batch, masks, anchor):
        prob = model._run_forward(batch, masks)
        ctx.model = model
        ctx.token = model._fwd_token
        return prob

    @staticmethod
    def backward(ctx, dprob):
        model = ctx.model
        if ctx.token != model._fwd_token:
            raise RuntimeError("MM_FiBiNET: backward() after another forward() of the same module is not supported "
                               "(activations live in a per-module workspace)")
        model._run_backward(dprob.contiguous())
        return None, None, None, None


class MM_FiBiNET(nn.Module):
    """Drop-in for the reference MM_FiBiNET (src/model_fibinet.py:91-199).

    ``feature_map`` is accepted and ignored like in the reference; when it is a dict it may carry
    B200-side options (the reference's callers pass None):  {"precision": "fp32"|"tf32x3"|"f16x3"|"bf16",
    "bilinear_type": "all"|"each"|"interaction", "dropout": float, "table_sharding": "row", "item_rows": V,
    "shard_rank": r, "shard_world": N}.  With ``table_sharding="row"`` the item table is partitioned by ``id % N`` over the
    N ranks of the process group (see sharded.py): ``item_emb`` then holds this rank's (ceil(V/N),128) slice and training
    goes through engine.ShardedTrainStep.  ``model_cfg`` keys other than
    ``embedding_dim`` are ignored exactly as the reference ignores them unless
    ``model_cfg["honor_config"]`` is true, in which case bilinear_type / net_dropout are honoured.
    """

    def __init__(self, feature_map, model_cfg):
        super().__init__()
        self.emb_dim = model_cfg.get("embedding_dim", 64)
        if self.emb_dim != D:
            raise ValueError(f"embedding_dim={self.emb_dim}: the sm_100a kernels are specialised for "
                             f"embedding_dim == {D} (the value in config/fibinet_config.yaml)")
        opts = dict(feature_map) if isinstance(feature_map, dict) else {}
        bilinear_type, dropout, senet_reduction = "all", 0.2, 2      # hard-coded in the reference (:114,:118,:129,:133)
        if model_cfg.get("honor_config", False):
            bilinear_type = model_cfg.get("bilinear_type", bilinear_type)
            dropout = float(model_cfg.get("net_dropout", dropout))
            senet_reduction = int(model_cfg.get("senet_reduction", senet_reduction))
        senet_reduction = int(opts.get("senet_reduction", senet_reduction))
        if senet_reduction < 1:
            raise ValueError("senet_reduction must be >= 1")
        bilinear_type = opts.get("bilinear_type", bilinear_type)
        self.dropout_p = float(opts.get("dropout", dropout))
        self.precision = opts.get("precision", model_cfg.get("precision", "fp32"))
        if self.precision not in _lib.PRECISIONS:
            raise ValueError(f"precision must be one of {list(_lib.PRECISIONS)}")
        mm_input_dim = 128
        self._shard = None
        sharding = opts.get("table_sharding", None)
        if sharding not in (None, "none", "replicated", "row"):
            raise ValueError("table_sharding must be 'row' or None")
        item_rows = int(opts.get("item_rows", ITEM_ROWS))
        # creation order == the reference's (same RNG consumption, same state_dict order)
        if sharding == "row":
            from . import sharded
            import torch.distributed as dist
            on = dist.is_available() and dist.is_initialized()
            rank = int(opts.get("shard_rank", dist.get_rank() if on else 0))
            world = int(opts.get("shard_world", dist.get_world_size() if on else 1))
            self._shard = sharded.ShardState(item_rows, rank, world)
            R = self._shard.shard_rows
            if item_rows <= (1 << 22):
                # same RNG consumption and the same values as the replicated table: initialise all rows, keep my slice
                full = nn.Embedding(item_rows, self.emb_dim, padding_idx=0)
                self.item_emb = nn.Embedding(R, self.emb_dim, _weight=sharded.slice_of_full(full.weight.data, rank, world))
                del full
            else:
                g = torch.Generator().manual_seed(0x5EED + 7919 * rank)
                self.item_emb = nn.Embedding(R, self.emb_dim, _weight=torch.randn(R, self.emb_dim, generator=g))
                if rank == 0:
                    with torch.no_grad():
                        self.item_emb.weight[0].zero_()      # global row 0 = padding
        else:
            self.item_emb = nn.Embedding(item_rows, self.emb_dim, padding_idx=0)
        self.user_emb = nn.Embedding(USER_ROWS, self.emb_dim)   # allocated, never used
        self.cate_emb = nn.Embedding(CATE_ROWS, self.emb_dim)
        self.mm_proj = nn.Sequential(nn.Linear(mm_input_dim, self.emb_dim), nn.LayerNorm(self.emb_dim), nn.ReLU())
        self.num_fields = NUM_FIELDS
        self.senet = SENetLayer(self.num_fields, reduction_ratio=senet_reduction)     # hidden = max(1, 6 // ratio) in {6, 3, 2, 1}
        self.bilinear = BilinearInteraction(self.emb_dim, self.num_fields, bilinear_type=bilinear_type)
        num_pairs = (self.num_fields * (self.num_fields - 1)) // 2
        total_input_dim = (self.num_fields + num_pairs) * self.emb_dim
        self.mlp = nn.Sequential(
            nn.Linear(total_input_dim, 512), nn.BatchNorm1d(512), nn.ReLU(), nn.Dropout(self.dropout_p),
            nn.Linear(512, 256), nn.BatchNorm1d(256), nn.ReLU(), nn.Dropout(self.dropout_p),
            nn.Linear(256, 1))
        self.sigmoid = nn.Sigmoid()
        # ---- B200-side state (never in the state_dict) ----
        self._flat: Optional[torch.Tensor] = None       # all dense parameters, one buffer
        self._gflat: Optional[torch.Tensor] = None      # their gradients, same layout
        self._layout = None
        self._item_grad: Optional[torch.Tensor] = None  # (rows,128) dense table gradient
        self._row_touched: Optional[torch.Tensor] = None
        self._grad_sumsq: Optional[torch.Tensor] = None
        self._ws = {}
        self._cur = None
        self._fwd_token = 0
        self._mm_table: Optional[torch.Tensor] = None   # optional resident item_emb_d128 table
        self.check_ids_every_forward = True             # eager module path: IndexError for out-of-range ids, like torch (one sync)
        self._fused_optimizer = None                    # set by FusedAdam
        # Philox key of the dropout masks: from torch's seed (set_seed(cfg seed) in the scripts) mixed with the rank, so that
        # runs with different seeds draw different masks and every rank draws its own (like DataParallel's replicas)
        self._seed, self._offset = self._derive_seed(), 0
        self._drop_ctr: Optional[torch.Tensor] = None   # device step counter shared by all TrainStep engines of this model
        self._test_masks = None
        self._dense_table_grad = False                  # data parallel: every table-gradient row is written

    @staticmethod
    def _derive_seed() -> int:
        import torch.distributed as dist
        rank = dist.get_rank() if (dist.is_available() and dist.is_initialized()) else 0
        z = (int(torch.initial_seed()) * 0x9E3779B97F4A7C15 + (rank + 1) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
        z ^= z >> 31
        return z & 0x7FFFFFFFFFFFFFFF

    def _dropout_counter(self, device) -> torch.Tensor:
        if self._drop_ctr is None or self._drop_ctr.device != torch.device(device):
            self._drop_ctr = torch.zeros(1, dtype=torch.int32, device=device)
        return self._drop_ctr

    def set_dropout_counter(self, n: int):
        """Continue the dropout stream of an earlier run at optimizer step ``n`` (checkpoint resume)."""
        self._dropout_counter(self.item_emb.weight.device).fill_(int(n))

    # ------------------------------------------------------------------ parameter plumbing
    def _dense_params(self):
        sd = dict(self.named_parameters())
        out = []
        for field, key in _DENSE:
            if field == "bil_w":
                out.append((field, self.bilinear.weights()))
            else:
                out.append((field, [sd[key]]))
        return out

    def _ensure_flat(self):
        """Re-home every dense parameter into one flat fp32 buffer (16-byte aligned slices)."""
        dev = self.item_emb.weight.device
        groups = self._dense_params()
        ok = self._flat is not None and self._flat.device == dev
        if ok:
            for (field, plist), (off, _) in zip(groups, self._layout):
                o = off
                for p in plist:
                    if p.data_ptr() != self._flat.data_ptr() + 4 * o or p.dtype != torch.float32:
                        ok = False
                    o += (p.numel() + 3) // 4 * 4
        if ok:
            return
        layout, total = [], 0
        for field, plist in groups:
            layout.append((total, sum((p.numel() + 3) // 4 * 4 for p in plist)))
            total += layout[-1][1]
        flat = torch.zeros(total, dtype=torch.float32, device=dev)
        with torch.no_grad():
            for (field, plist), (off, _) in zip(groups, layout):
                o = off
                for p in plist:
                    n = p.numel()
                    flat[o:o + n].copy_(p.data.reshape(-1).to(torch.float32))
                    p.data = flat[o:o + n].view(p.shape)
                    o += (n + 3) // 4 * 4
        self._flat, self._layout = flat, layout
        self._gflat = torch.zeros_like(flat)
        self._ws = {}
        self._item_grad = None

    def attach_mm_table(self, table: torch.Tensor):
        """Keep the frozen (item_rows,128) item_emb_d128 matrix resident on the GPU; batches that do
        not carry ``item_emb_d128`` are then gathered from it inside the fused kernel."""
        rows = self._shard.item_rows if self._shard is not None else self.item_emb.weight.shape[0]
        table = torch.as_tensor(table)
        if table.dim() != 2 or table.shape[1] != D:
            raise ValueError(f"item_emb_d128 table must be (rows, {D})")
        if table.shape[0] < rows:       # the kernel indexes it with the clamped item id: pad so that every id is in range
            table = torch.cat([table, torch.zeros(rows - table.shape[0], D, dtype=table.dtype)], 0)
        self._mm_table = table.to(self.item_emb.weight.device, torch.float32).contiguous()

    def _params_struct(self) -> _lib.Params:
        P = _lib.Params()
        P.item_emb = self.item_emb.weight.data_ptr()
        P.item_rows = self.item_emb.weight.shape[0]
        if self._shard is not None:
            st = self._shard
            P.item_rows = st.item_rows
            P.n_shards, P.shard_rank, P.shard_rows = st.world, st.rank, st.shard_rows
            for r, q in enumerate(st.ensure_table(self.item_emb.weight)):
                P.shard[r] = q
        P.cate_rows = self.cate_emb.weight.shape[0]
        for (field, plist), (off, _) in zip(self._dense_params(), self._layout):
            setattr(P, field, self._flat.data_ptr() + 4 * off)
        P.bn1_mean = self.mlp[1].running_mean.data_ptr()
        P.bn1_var = self.mlp[1].running_var.data_ptr()
        P.bn2_mean = self.mlp[5].running_mean.data_ptr()
        P.bn2_var = self.mlp[5].running_var.data_ptr()
        P.bilinear_type = _lib.BILINEAR_TYPES[self.bilinear.bilinear_type]
        P.precision = _lib.PRECISIONS[self.precision]
        P.se_hidden = self.senet.reduced_size
        return P

    def _bucket1_offset(self) -> int:
        """Element offset of the MLP-1 gradient bucket (the tail of the flat gradient buffer)."""
        for (field, _), (off, _n) in zip(self._dense_params(), self._layout):
            if field == BUCKET1_FIRST:
                return off
        raise AssertionError

    def _grads_struct(self) -> _lib.Grads:
        G = _lib.Grads()
        for (field, plist), (off, _) in zip(self._dense_params(), self._layout):
            setattr(G, field, self._gflat.data_ptr() + 4 * off)
        return G

    def _ws_rows(self) -> int:
        return 1 if self._shard is not None else self.item_emb.weight.shape[0]

    def _workspace(self, B: int, L: int):
        key = (B, L)
        ws = self._ws.get(key)
        if ws is None:
            lib = _lib.load()
            nbytes = lib.fbn_workspace_bytes(B, L, self._ws_rows())
            if len(self._ws) >= 4:           # bound the cache (distinct tail-batch sizes)
                self._ws.pop(next(iter(self._ws)))
            ws = torch.zeros(nbytes, dtype=torch.uint8, device=self.item_emb.weight.device)
            self._ws[key] = ws
        return ws

    def workspace_view(self, name: str, shape, dtype=torch.float32) -> torch.Tensor:
        """Debug/test accessor: a named activation of the most recent forward (see fbn_workspace_offset)."""
        B, L, ws = self._cur["B"], self._cur["L"], self._cur["ws"]
        off = _lib.load().fbn_workspace_offset(B, L, self._ws_rows(), name.encode())
        if off == C.c_size_t(-1).value:
            raise KeyError(name)
        n = 1
        for s in shape:
            n *= s
        itemsize = torch.empty((), dtype=dtype).element_size()
        return ws[off:off + n * itemsize].view(dtype).view(*shape)

    def check_ids(self, ws: Optional[torch.Tensor] = None, B: Optional[int] = None, L: Optional[int] = None):
        """Raise IndexError if the gather kernel met an id outside its table since the last check (torch's nn.Embedding raises
        for these, reference src/model_fibinet.py:155-167; the kernel clamps so that it cannot fault and sets a sticky device
        flag).  Synchronises the stream: the eager module path calls it after every forward (``check_ids_every_forward``),
        TrainStep / Scorer users call ``engine.check_ids()`` whenever they read results back."""
        if ws is None:
            ws, B, L = self._cur["ws"], self._cur["B"], self._cur["L"]
        off = _lib.load().fbn_workspace_offset(B, L, self._ws_rows(), b"idflag")
        flag = ws[off:off + 16].view(torch.int32)
        f = flag.tolist()
        if any(f[:3]):
            flag.zero_()
            bad = [n for n, v in zip(("item_id", "likes_level / views_level", "item_seq"), f) if v]
            rows = self._shard.item_rows if self._shard is not None else self.item_emb.weight.shape[0]
            raise IndexError(f"index out of range in self: {', '.join(bad)} (item table has {rows} rows, "
                             f"cate table {self.cate_emb.weight.shape[0]})")

    # ------------------------------------------------------------------ batch marshalling
    def _batch_struct(self, batch_dict: Dict[str, torch.Tensor]):
        item_id = batch_dict["item_id"]
        likes, views = batch_dict["likes_level"], batch_dict["views_level"]
        _require_cuda(item_id, "batch_dict['item_id']")
        if item_id.dtype not in _IDX_DTYPES:
            raise TypeError(f"item_id dtype {item_id.dtype} not supported (int32/int64/float32/float64)")
        keep = []
        idt = item_id.dtype

        def col(t, name):
            _require_cuda(t, f"batch_dict['{name}']")
            if t.dtype != idt:
                t = t.to(idt)
            t = t.reshape(-1).contiguous()
            keep.append(t)
            return t
        item_id, likes, views = col(item_id, "item_id"), col(likes, "likes_level"), col(views, "views_level")
        B = item_id.shape[0]
        bs = _lib.Batch()
        bs.batch = B
        bs.item_id, bs.likes_level, bs.views_level = item_id.data_ptr(), likes.data_ptr(), views.data_ptr()
        bs.idx_dtype = _IDX_DTYPES[idt]
        seq = batch_dict.get("item_seq", None)
        L = 0
        bs.seq_dtype = _lib.IDX_I64
        if seq is not None:
            _require_cuda(seq, "batch_dict['item_seq']")
            if seq.dtype not in (torch.int32, torch.int64):
                raise TypeError("item_seq must be an integer tensor (the reference indexes nn.Embedding with it as-is)")
            seq = seq.reshape(B, -1).contiguous()
            keep.append(seq)
            L = seq.shape[1]
            if L > 64:
                raise ValueError("item_seq longer than 64 ids is not supported (the loader crops to max_len=20)")
            bs.item_seq = seq.data_ptr()
            bs.seq_dtype = _IDX_DTYPES[seq.dtype]
        bs.seq_len = L
        mm = batch_dict.get("item_emb_d128", None)
        if mm is not None:
            _require_cuda(mm, "batch_dict['item_emb_d128']")
            mm = mm.to(torch.float32).reshape(B, D).contiguous()       # .float(), ref :141
            keep.append(mm)
            bs.item_mm = mm.data_ptr()
        elif self._mm_table is not None:
            bs.mm_table = self._mm_table.data_ptr()
        else:
            raise KeyError("item_emb_d128")
        return bs, keep, B, L

    # ------------------------------------------------------------------ kernels
    def _run_forward(self, batch_dict, masks=None) -> torch.Tensor:
        lib = _lib.load()
        _require_cuda(self.item_emb.weight, "MM_FiBiNET parameters")
        self._ensure_flat()
        bs, keep, B, L = self._batch_struct(batch_dict)
        ws = self._workspace(B, L)
        P = self._params_struct()
        prob = torch.empty(B, dtype=torch.float32, device=ws.device)
        train = bool(self.training)
        masks = masks if masks is not None else self._test_masks
        m1 = m2 = None
        if masks is not None:
            m1, m2 = (m.to(device=ws.device, dtype=torch.uint8).contiguous() for m in masks)
        p_drop = self.dropout_p if train else 0.0
        self._offset += 1
        rc = lib.fbn_forward(C.byref(P), C.byref(bs), _lib.ptr(ws), ws.numel(), int(train), p_drop, _lib.ptr(m1), _lib.ptr(m2),
                             self._seed, self._offset << 32, None, _lib.ptr(prob), _lib.stream_ptr())
        _lib.check(rc, "fbn_forward")
        if train:
            with torch.no_grad():
                self.mlp[1].num_batches_tracked += 1
                self.mlp[5].num_batches_tracked += 1
        self._fwd_token += 1
        self._cur = dict(B=B, L=L, ws=ws, bs=bs, keep=keep, train=train, p_drop=p_drop, masks=(m1, m2))
        if self.check_ids_every_forward and not torch.cuda.is_current_stream_capturing():
            self.check_ids()
        return prob

    def _run_backward(self, dprob: torch.Tensor):
        lib = _lib.load()
        if self._shard is not None:
            raise RuntimeError("row-sharded item table: loss.backward() through autograd is not supported -- the table gradient "
                               "is exchanged between ranks; train with engine.ShardedTrainStep")
        cur = self._cur
        dev = cur["ws"].device
        rows = self.item_emb.weight.shape[0]
        if self._item_grad is None or self._item_grad.device != dev:
            self._item_grad = torch.zeros(rows, D, dtype=torch.float32, device=dev)
            self._row_touched = torch.zeros(rows, dtype=torch.int32, device=dev)
            self._grad_sumsq = torch.zeros(2, dtype=torch.float32, device=dev)
        fused = self._fused_optimizer is not None
        accumulate = (not fused) and any(p.grad is not None for _, pl in self._dense_params() for p in pl)
        if accumulate:
            prev_dense = self._gflat.clone()
            prev_item = self.item_emb.weight.grad.clone() if self.item_emb.weight.grad is not None else None
        P, G = self._params_struct(), self._grads_struct()
        rc = lib.fbn_backward(C.byref(P), C.byref(cur["bs"]), _lib.ptr(cur["ws"]), cur["ws"].numel(), int(cur["train"]),
                              cur["p_drop"], _lib.ptr(dprob), C.byref(G), _lib.ptr(self._gflat), self._gflat.numel(),
                              _lib.ptr(self._item_grad), _lib.ptr(self._row_touched), 0 if (fused and not self._dense_table_grad) else 1, 0,
                              _lib.ptr(self._grad_sumsq), _lib.stream_ptr())
        _lib.check(rc, "fbn_backward")
        if accumulate:
            self._gflat += prev_dense
            if prev_item is not None:
                self._item_grad += prev_item
        # attach: .grad aliases the flat gradient buffer
        for (field, plist), (off, _) in zip(self._dense_params(), self._layout):
            o = off
            for p in plist:
                n = p.numel()
                p.grad = self._gflat[o:o + n].view(p.shape)
                o += (n + 3) // 4 * 4
        if not fused:
            self.item_emb.weight.grad = self._item_grad
        # user_emb: no gradient at all, like the reference (grad stays None)

    def forward(self, batch_dict):
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            return _FibinetFn.apply(self, batch_dict, None, self.mlp[8].bias)
        return self._run_forward(batch_dict)


def build_model(feature_map, model_cfg):
    """The reference's factory (src/model_fibinet.py:201-202).  ``feature_map`` is None / ignored for the reference's six-field
    model; a dict with a "fields" list selects the F-field model of general.py (BASELINE config 5)."""
    from .general import GeneralFiBiNET, fields_from_feature_map
    fields = fields_from_feature_map(feature_map)
    if fields is not None:
        fm = feature_map
        return GeneralFiBiNET(fields, model_cfg, bilinear_type=fm.get("bilinear_type", "all"),
                              senet_reduction=int(fm.get("senet_reduction", 2)), dropout=float(fm.get("dropout", 0.2)),
                              precision=fm.get("precision", "f16x3"))
    return MM_FiBiNET(feature_map, model_cfg)
