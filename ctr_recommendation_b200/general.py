"""F-field FiBiNET on the CUDA path (BASELINE config 5: "scaled synthetic FiBiNET: 40 fields").

The reference model is hard-wired to six fields (src/model_fibinet.py:112,179-182) -- its building blocks are not: F embedding
lookups (:155-159), SENetLayer(num_fields, reduction_ratio) (:5-35), BilinearInteraction(input_dim, num_fields, bilinear_type)
(:37-89), the concat (:191-194) and the MLP tower (:125-136).  ``GeneralFiBiNET`` assembles exactly those blocks for an arbitrary
field list, every FLOP in libfibinet_b200.so:

    fbn_fields_gather -> fbn_senet_fwd -> fbn_bilinear_fwd_ld -> fbn_tower_forward          (forward)
    fbn_tower_backward -> fbn_bilinear_bwd_ld -> fbn_senet_bwd -> fbn_fields_scatter        (backward)

``build_model(feature_map, model_cfg)`` returns it when ``feature_map`` carries a field list:

    feature_map = {"fields": [("user_id", 20000), ("item_id", 91718), ("likes_level", 11), ...],      # (name, vocabulary)
                   "bilinear_type": "all" | "each" | "interaction", "senet_reduction": 2, "dropout": 0.2, "precision": "f16x3"}

A field may also be a dict -- what the reference's commented-out tag lookup (src/dataloader.py:100-102), its unused ``user_emb``
(src/model_fibinet.py:101,152) and its shared tables (:155-156,159,167) need:

    {"name": "item_seq", "table": "item_id", "bag": 20}      # 20 ids per sample into item_id's table, id 0 = padding, masked MEAN
    {"name": "item_tags", "vocab": 3000, "bag": 5}            # pooled exactly like the reference pools item_seq (:165-174)
    {"name": "item_id", "vocab": 91718, "padding_idx": 0}     # nn.Embedding(padding_idx=0): zero row, zero gradient (:100)
    {"name": "views_level", "table": "likes_level"}           # two fields, one table (cate_emb)

``forward(batch_dict)`` takes one integer id tensor per field name -- (B,) or, for a bag, (B, bag) -- or a ready (B, id columns)
tensor under "ids", and returns (B,) probabilities.  The tables are stored back to back in ONE (sum of vocabularies, 128) parameter
``emb.weight``, so the deterministic sorted-segment scatter-add and the table optimizers are shared with the six-field model.
Parameters are ordinary nn.Parameters with ordinary ``.grad`` tensors: any torch optimizer drives it.

Oracle: oracle/fibinet_general.py (pinned on CPU against a torch model built from the reference's own SENetLayer /
BilinearInteraction classes, tests/test_oracle_general.py); GPU parity: tests/test_gpu_general.py.  Unfused by design -- the
six-field model keeps its fused kernels; this path exists so that the 40-field configuration runs on the same tcgen05 GEMMs.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Sequence, Tuple

import torch
import torch.nn as nn

from . import _lib
from .model import BilinearInteraction, SENetLayer, _require_cuda

D = 128


class _GeneralFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, ids, *params):
        prob = model._run_forward(ids)
        ctx.model, ctx.token = model, model._fwd_token
        return prob

    @staticmethod
    def backward(ctx, dprob):
        model = ctx.model
        if ctx.token != model._fwd_token:
            raise RuntimeError("GeneralFiBiNET: backward() after another forward() of the same module is not supported")
        grads = model._run_backward(dprob.contiguous())
        return (None, None, *grads)


class GeneralFiBiNET(nn.Module):
    def __init__(self, fields: Sequence[Tuple[str, int]], model_cfg: dict | None = None, bilinear_type: str = "all",
                 senet_reduction: int = 2, dropout: float = 0.2, precision: str = "f16x3"):
        super().__init__()
        model_cfg = model_cfg or {}
        if int(model_cfg.get("embedding_dim", D)) != D:
            raise ValueError(f"embedding_dim must be {D} (the sm_100a kernels move one 512-byte row per warp instruction)")
        specs = [f if isinstance(f, dict) else {"name": f[0], "vocab": f[1]} for f in fields]
        self.field_names: List[str] = [str(f["name"]) for f in specs]
        F = len(specs)
        if not 2 <= F <= 64:
            raise ValueError("GeneralFiBiNET supports 2..64 fields")
        if len(set(self.field_names)) != F:
            raise ValueError("field names must be unique")
        if precision not in _lib.PRECISIONS:
            raise ValueError(f"precision must be one of {list(_lib.PRECISIONS)}")
        self.num_fields, self.precision, self.dropout_p = F, precision, float(dropout)
        # tables (a field either brings its own or names an earlier field's), stored back to back
        table_of, row0, vocab_of, pad_of = {}, [], [], []
        self.bags: List[int] = []
        desc, col_field, col = [], [], 0
        total = 0
        for fi, f in enumerate(specs):
            name = str(f["name"])
            if f.get("table") is not None:
                t = table_of.get(str(f["table"]))
                if t is None:
                    raise ValueError(f"field {name!r}: table {f['table']!r} must name an earlier field that owns a table")
            else:
                v = int(f.get("vocab", f.get("vocab_size", 0)))
                if v < 1:
                    raise ValueError(f"field {name!r} needs a vocabulary of at least one row")
                t = len(row0)
                row0.append(total)
                vocab_of.append(v)
                pad_of.append(int(f["padding_idx"]) if f.get("padding_idx") is not None else -1)
                total += v
            table_of[name] = t
            bag = int(f.get("bag", 1))
            if bag < 1:
                raise ValueError(f"field {name!r}: bag must be >= 1")
            # a bag masks its padding id (0 unless the table says otherwise); a single lookup only if the table has a padding row
            pad = pad_of[t] if pad_of[t] >= 0 else (0 if bag > 1 else -1)
            desc.append([row0[t], vocab_of[t], col, bag, pad])
            col_field += [fi] * bag
            col += bag
            self.bags.append(bag)
        self.id_cols = col
        self.vocabs: List[int] = vocab_of
        self._pad_rows = [r + p for r, p in zip(row0, pad_of) if p >= 0]
        self.register_buffer("field_desc", torch.tensor(desc, dtype=torch.int64), persistent=False)
        self.register_buffer("col_field", torch.tensor(col_field, dtype=torch.int32), persistent=False)
        offs = [0, total]
        # creation order follows the reference's __init__: tables, SENET, bilinear, MLP (src/model_fibinet.py:100-135)
        self.emb = nn.Embedding(offs[-1], D)
        with torch.no_grad():
            for r in self._pad_rows:
                self.emb.weight[r].zero_()                        # nn.Embedding(padding_idx=...) initialises that row to zero
        self.senet = SENetLayer(F, reduction_ratio=senet_reduction)
        self.bilinear = BilinearInteraction(D, F, bilinear_type=bilinear_type)
        self.num_pairs = F * (F - 1) // 2
        self.k1 = (F + self.num_pairs) * D
        self.mlp = nn.Sequential(
            nn.Linear(self.k1, 512), nn.BatchNorm1d(512), nn.ReLU(), nn.Dropout(self.dropout_p),
            nn.Linear(512, 256), nn.BatchNorm1d(256), nn.ReLU(), nn.Dropout(self.dropout_p),
            nn.Linear(256, 1))
        self.sigmoid = nn.Sigmoid()
        self._seed, self._offset = 0x5EED ^ int(torch.initial_seed() & 0x7FFFFFFF), 0
        self._test_masks = None
        self._fwd_token = 0
        self._cur = None
        self._buf: Dict[tuple, dict] = {}
        self.check_ids_every_forward = True

    # ------------------------------------------------------------------ plumbing
    def _param_list(self):
        e0, e2 = self.senet.excitation[0], self.senet.excitation[2]
        return [self.emb.weight, e0.weight, e0.bias, e2.weight, e2.bias, *self.bilinear.weights(),
                self.mlp[0].weight, self.mlp[0].bias, self.mlp[1].weight, self.mlp[1].bias,
                self.mlp[4].weight, self.mlp[4].bias, self.mlp[5].weight, self.mlp[5].bias, self.mlp[8].weight, self.mlp[8].bias]

    def _params_struct(self) -> _lib.Params:
        P = _lib.Params()
        m = self.mlp
        P.w1, P.b1, P.bn1_g, P.bn1_b = (t.data_ptr() for t in (m[0].weight, m[0].bias, m[1].weight, m[1].bias))
        P.bn1_mean, P.bn1_var = m[1].running_mean.data_ptr(), m[1].running_var.data_ptr()
        P.w2, P.b2, P.bn2_g, P.bn2_b = (t.data_ptr() for t in (m[4].weight, m[4].bias, m[5].weight, m[5].bias))
        P.bn2_mean, P.bn2_var = m[5].running_mean.data_ptr(), m[5].running_var.data_ptr()
        P.w3, P.b3 = m[8].weight.data_ptr(), m[8].bias.data_ptr()
        P.precision = _lib.PRECISIONS[self.precision]
        P.bilinear_type = _lib.BILINEAR_TYPES[self.bilinear.bilinear_type]
        return P

    def _scratch(self, B: int) -> dict:
        buf = self._buf.get(B)
        if buf is None:
            lib, dev, F = _lib.load(), self.emb.weight.device, self.num_fields
            R = self.senet.reduced_size
            btype = _lib.BILINEAR_TYPES[self.bilinear.bilinear_type]
            z = dict(dtype=torch.float32, device=dev)
            if len(self._buf) >= 2:
                self._buf.pop(next(iter(self._buf)))
            buf = dict(
                X=torch.empty(B, F, D, **z), V=torch.empty(B, F, D, **z), gate=torch.empty(B, F, **z), cnt=torch.empty(B, F, **z),
                C=torch.zeros(B, self.k1, **z), dC=torch.empty(B, self.k1, **z), dV=torch.empty(B, F, D, **z), dX=torch.empty(B, F, D, **z),
                tower=torch.zeros(lib.fbn_tower_workspace_bytes(B, self.k1), dtype=torch.uint8, device=dev),
                bil=torch.empty(lib.fbn_bilinear_scratch_bytes(B, F, D, btype), dtype=torch.uint8, device=dev),
                se=torch.empty(lib.fbn_senet_scratch_bytes(B, F, R), dtype=torch.uint8, device=dev),
                scatter=torch.empty(lib.fbn_fields_scatter_bytes(B, self.id_cols, self.emb.weight.shape[0]), dtype=torch.uint8, device=dev),
                flag=torch.zeros(4, dtype=torch.int32, device=dev), sumsq=torch.zeros(2, **z))
            self._buf[B] = buf
        return buf

    def _ids(self, batch) -> torch.Tensor:
        if torch.is_tensor(batch):
            ids = batch
        elif "ids" in batch:
            ids = batch["ids"]
        else:
            cols = []
            for n, bag in zip(self.field_names, self.bags):
                t = batch[n]
                _require_cuda(t, f"batch_dict['{n}']")
                t = t.long()                                   # tensor.long(), like src/model_fibinet.py:140-143
                t = t.reshape(-1, 1) if bag == 1 else t.reshape(t.shape[0], -1)
                if t.shape[1] != bag:
                    raise ValueError(f"batch_dict['{n}'] must carry {bag} id(s) per sample, got {tuple(t.shape)}")
                cols.append(t)
            ids = torch.cat(cols, 1)
        _require_cuda(ids, "ids")
        if ids.dtype not in (torch.int32, torch.int64):
            ids = ids.long()
        if ids.dim() != 2 or ids.shape[1] != self.id_cols:
            raise ValueError(f"ids must be (B, {self.id_cols})")
        return ids.contiguous()

    def tower_view(self, name: str, shape) -> torch.Tensor:
        """Debug/test accessor: a named activation of the most recent forward's MLP tower ("A1", "A2", "logit", ...)."""
        B = self._cur["B"]
        off = _lib.load().fbn_tower_workspace_offset(B, self.k1, name.encode())
        if off == C.c_size_t(-1).value:
            raise KeyError(name)
        n = 1
        for d in shape:
            n *= d
        return self._cur["buf"]["tower"][off:off + 4 * n].view(torch.float32).view(*shape)

    def check_ids(self):
        """IndexError if a lookup since the last check was outside its field's vocabulary (nn.Embedding raises; the kernel clamps
        and sets a sticky device flag).  Synchronises."""
        for buf in self._buf.values():
            if int(buf["flag"][0].item()):
                buf["flag"].zero_()
                raise IndexError("index out of range in self: a field id is outside its vocabulary")

    # ------------------------------------------------------------------ kernels
    def _run_forward(self, ids: torch.Tensor) -> torch.Tensor:
        lib, st = _lib.load(), _lib.stream_ptr()
        _require_cuda(self.emb.weight, "GeneralFiBiNET parameters")
        B, F = ids.shape[0], self.num_fields
        buf = self._scratch(B)
        idt = _lib.IDX_I32 if ids.dtype == torch.int32 else _lib.IDX_I64
        e0, e2 = self.senet.excitation[0], self.senet.excitation[2]
        R = self.senet.reduced_size
        btype = _lib.BILINEAR_TYPES[self.bilinear.bilinear_type]
        prec = _lib.PRECISIONS["tf32x3" if self.precision == "f16x3" else self.precision]   # short-K bilinear GEMMs: see _lib.PRECISIONS
        _lib.check(lib.fbn_fields_gather(_lib.ptr(self.emb.weight), _lib.ptr(self.field_desc), _lib.ptr(ids), idt, B, F, self.id_cols,
                                         _lib.ptr(buf["X"]), _lib.ptr(buf["cnt"]), _lib.ptr(buf["flag"]), st), "fbn_fields_gather")
        _lib.check(lib.fbn_senet_fwd(_lib.ptr(buf["X"]), _lib.ptr(e0.weight), _lib.ptr(e0.bias), _lib.ptr(e2.weight), _lib.ptr(e2.bias),
                                     B, F, D, R, _lib.ptr(buf["V"]), _lib.ptr(buf["gate"]), st), "fbn_senet_fwd")
        Cm = buf["C"]
        Cm[:, :F * D].copy_(buf["V"].view(B, F * D))                      # the concat (ref :191-194): a strided device copy
        W = self._bil_weight()
        pairs = C.c_void_p(Cm.data_ptr() + 4 * F * D)
        _lib.check(lib.fbn_bilinear_fwd_ld(_lib.ptr(buf["V"]), _lib.ptr(W), btype, B, F, D, pairs, self.k1, _lib.ptr(buf["bil"]),
                                           buf["bil"].numel(), prec, st), "fbn_bilinear_fwd_ld")
        train = bool(self.training)
        masks = self._test_masks
        m1 = m2 = None
        if masks is not None:
            m1, m2 = (m.to(device=Cm.device, dtype=torch.uint8).contiguous() for m in masks)
        self._offset += 1
        prob = torch.empty(B, dtype=torch.float32, device=Cm.device)
        logit = torch.empty(B, dtype=torch.float32, device=Cm.device)
        P = self._params_struct()
        _lib.check(lib.fbn_tower_forward(C.byref(P), _lib.ptr(Cm), B, self.k1, _lib.ptr(buf["tower"]), buf["tower"].numel(), int(train),
                                         self.dropout_p if train else 0.0, _lib.ptr(m1), _lib.ptr(m2), self._seed, self._offset << 32,
                                         None, _lib.ptr(prob), _lib.ptr(logit), st), "fbn_tower_forward")
        if train:
            with torch.no_grad():
                self.mlp[1].num_batches_tracked += 1
                self.mlp[5].num_batches_tracked += 1
        self._fwd_token += 1
        self._cur = dict(B=B, ids=ids, idt=idt, buf=buf, train=train, W=W, logit=logit, keep=(m1, m2))
        if self.check_ids_every_forward and not torch.cuda.is_current_stream_capturing():
            self.check_ids()
        return prob

    def _bil_weight(self) -> torch.Tensor:
        ws = self.bilinear.weights()
        return ws[0] if len(ws) == 1 else torch.stack([w.detach() for w in ws], 0).contiguous()      # (nW, D, D) operand for the kernels

    def _run_backward(self, dprob: torch.Tensor):
        lib, st = _lib.load(), _lib.stream_ptr()
        cur = self._cur
        B, F, buf = cur["B"], self.num_fields, cur["buf"]
        dev = dprob.device
        e0, e2 = self.senet.excitation[0], self.senet.excitation[2]
        R = self.senet.reduced_size
        btype = _lib.BILINEAR_TYPES[self.bilinear.bilinear_type]
        prec = _lib.PRECISIONS["tf32x3" if self.precision == "f16x3" else self.precision]
        m = self.mlp
        new = lambda t: torch.empty_like(t, memory_format=torch.contiguous_format)
        g_w1, g_b1, g_g1, g_be1 = new(m[0].weight), new(m[0].bias), new(m[1].weight), new(m[1].bias)
        g_w2, g_b2, g_g2, g_be2 = new(m[4].weight), new(m[4].bias), new(m[5].weight), new(m[5].bias)
        g_w3, g_b3 = new(m[8].weight), new(m[8].bias)
        G = _lib.Grads()
        G.w1, G.b1, G.bn1_g, G.bn1_b = (t.data_ptr() for t in (g_w1, g_b1, g_g1, g_be1))
        G.w2, G.b2, G.bn2_g, G.bn2_b = (t.data_ptr() for t in (g_w2, g_b2, g_g2, g_be2))
        G.w3, G.b3 = g_w3.data_ptr(), g_b3.data_ptr()
        P = self._params_struct()
        Cm, dC = buf["C"], buf["dC"]
        _lib.check(lib.fbn_tower_backward(C.byref(P), _lib.ptr(Cm), B, self.k1, _lib.ptr(buf["tower"]), buf["tower"].numel(),
                                          int(cur["train"]), self.dropout_p if cur["train"] else 0.0, _lib.ptr(dprob), C.byref(G),
                                          _lib.ptr(dC), st), "fbn_tower_backward")
        W = cur["W"]
        dW = torch.empty_like(W)
        dpairs = C.c_void_p(dC.data_ptr() + 4 * F * D)
        _lib.check(lib.fbn_bilinear_bwd_ld(_lib.ptr(buf["V"]), _lib.ptr(W), dpairs, self.k1, _lib.ptr(dC), self.k1, btype, B, F, D,
                                           _lib.ptr(buf["dV"]), _lib.ptr(dW), _lib.ptr(buf["bil"]), buf["bil"].numel(), prec, st),
                   "fbn_bilinear_bwd_ld")
        g_sw1, g_sb1, g_sw2, g_sb2 = new(e0.weight), new(e0.bias), new(e2.weight), new(e2.bias)
        _lib.check(lib.fbn_senet_bwd(_lib.ptr(buf["X"]), _lib.ptr(buf["gate"]), _lib.ptr(e0.weight), _lib.ptr(e0.bias), _lib.ptr(e2.weight),
                                     _lib.ptr(buf["dV"]), B, F, D, R, _lib.ptr(buf["dX"]), _lib.ptr(g_sw1), _lib.ptr(g_sb1), _lib.ptr(g_sw2),
                                     _lib.ptr(g_sb2), _lib.ptr(buf["se"]), buf["se"].numel(), st), "fbn_senet_bwd")
        rows = self.emb.weight.shape[0]
        g_emb = torch.empty(rows, D, dtype=torch.float32, device=dev)
        _lib.check(lib.fbn_fields_scatter(_lib.ptr(buf["dX"]), _lib.ptr(buf["cnt"]), _lib.ptr(self.field_desc), _lib.ptr(self.col_field),
                                          _lib.ptr(cur["ids"]), cur["idt"], B, F, self.id_cols, rows, _lib.ptr(g_emb), None, 1,
                                          _lib.ptr(buf["sumsq"]), _lib.ptr(buf["scatter"]), buf["scatter"].numel(), st), "fbn_fields_scatter")
        bil_grads = [dW] if len(self.bilinear.weights()) == 1 else list(dW.unbind(0))
        return [g_emb, g_sw1, g_sb1, g_sw2, g_sb2, *bil_grads, g_w1, g_b1, g_g1, g_be1, g_w2, g_b2, g_g2, g_be2, g_w3, g_b3]

    def forward(self, batch):
        ids = self._ids(batch)
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            return _GeneralFn.apply(self, ids, *self._param_list())
        return self._run_forward(ids)


def fields_from_feature_map(feature_map):
    """The field list of feature_map["fields"]: (name, vocabulary) pairs or dicts (name / vocab / table / bag / padding_idx)."""
    if not isinstance(feature_map, dict) or not feature_map.get("fields"):
        return None
    return [dict(f) if isinstance(f, dict) else (str(f[0]), int(f[1])) for f in feature_map["fields"]]
