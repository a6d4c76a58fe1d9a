"""FusedAdam: the reference's optimizer step (src/train_fibinet.py:78,119,121) on the CUDA path.

torch.optim.Adam(lr, weight_decay) -- L2-coupled, NOT AdamW -- preceded by clip_grad_norm_(10.0), with
OneCycleLR rewriting ``lr`` and ``betas[0]`` in ``param_groups`` before every step (SURVEY facts 6,7,9).
The embedding table is updated dense-exactly: every row gets g = (segment-sum or 0)*coef + wd*p.

Usage mirrors the reference loop:

    optimizer = FusedAdam(model, lr=lr, weight_decay=wd)        # instead of torch.optim.Adam(model.parameters(), ...)
    scheduler = torch.optim.lr_scheduler.OneCycleLR(optimizer, ...)   # unchanged
    loss.backward(); clip_grad_norm_(model, 10.0); optimizer.step(); scheduler.step()
"""
from __future__ import annotations

import ctypes as C
import math

import torch

from . import _lib


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, model, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, decoupled_weight_decay=False):
        from .model import MM_FiBiNET
        inner = model.module if hasattr(model, "module") else model
        if not isinstance(inner, MM_FiBiNET):
            raise TypeError("FusedAdam drives a ctr_recommendation_b200 MM_FiBiNET (pass the model, not parameters())")
        params = [p for n, p in inner.named_parameters() if not n.startswith("user_emb.")]
        super().__init__(params, dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay))
        self.model = inner
        # False: torch.optim.Adam (L2 folded into the gradient) = what the reference builds; True: torch.optim.AdamW
        # (`optimizer: adamw` of config/fibinet_config.yaml, which the reference never honours)
        self.decoupled = bool(decoupled_weight_decay)
        inner._fused_optimizer = self
        self._step = 0
        self._max_norm = None
        self._clip = None
        self._m_flat = self._v_flat = self._m_item = self._v_item = None

    # ---- state -------------------------------------------------------------------------------
    def _ensure_state(self):
        m = self.model
        m._ensure_flat()
        dev = m._flat.device
        if self._m_flat is None or self._m_flat.device != dev or self._m_flat.numel() != m._flat.numel():
            self._m_flat = torch.zeros_like(m._flat)
            self._v_flat = torch.zeros_like(m._flat)
            self._m_item = torch.zeros_like(m.item_emb.weight.data)
            self._v_item = torch.zeros_like(m.item_emb.weight.data)
            self._clip = torch.ones(2, dtype=torch.float32, device=dev)

    def moments(self):
        """{state_dict key: (exp_avg, exp_avg_sq)} views, for parity tests and checkpoints."""
        m = self.model
        out = {"item_emb.weight": (self._m_item, self._v_item)}
        names = {id(p): n for n, p in m.named_parameters()}
        for (field, plist), (off, _) in zip(m._dense_params(), m._layout):
            o = off
            for p in plist:
                n = p.numel()
                out[names[id(p)]] = (self._m_flat[o:o + n].view(p.shape), self._v_flat[o:o + n].view(p.shape))
                o += (n + 3) // 4 * 4
        return out

    def state_dict(self):
        self._ensure_state()
        return {"step": self._step, "param_groups": [{k: v for k, v in g.items() if k != "params"} for g in self.param_groups],
                "m_flat": self._m_flat, "v_flat": self._v_flat, "m_item": self._m_item, "v_item": self._v_item}

    def load_state_dict(self, sd):
        self._ensure_state()
        self._step = int(sd["step"])
        for g, s in zip(self.param_groups, sd["param_groups"]):
            g.update(s)
        for k in ("m_flat", "v_flat", "m_item", "v_item"):
            getattr(self, "_" + k).copy_(sd[k])

    # ---- clip + step -------------------------------------------------------------------------
    def clip_grad_norm_(self, max_norm: float):
        """Global L2-norm clip over all gradients (embedding rows included); the scaling itself is
        folded into the Adam kernels.  Returns the total norm as a 0-d device tensor (no host sync)."""
        lib = _lib.load()
        self._ensure_state()
        m = self.model
        if m._grad_sumsq is None:
            raise RuntimeError("clip_grad_norm_ called before backward()")
        _lib.check(lib.fbn_clip_coef(_lib.ptr(m._grad_sumsq), 2, float(max_norm), _lib.ptr(self._clip), _lib.stream_ptr()),
                   "fbn_clip_coef")
        self._max_norm = max_norm
        return self._clip[0]

    @torch.no_grad()
    def step(self, closure=None):
        if closure is not None:
            raise NotImplementedError("FusedAdam does not support closures")
        lib = _lib.load()
        self._ensure_state()
        m = self.model
        if m._item_grad is None:
            raise RuntimeError("FusedAdam.step() called before backward()")
        g = self.param_groups[0]
        self._step += 1
        h = _lib.AdamHyper(float(g["lr"]), float(g["betas"][0]), float(g["betas"][1]), float(g["eps"]),
                           float(g["weight_decay"]), self._step, 1.0 - float(g["betas"][0]), 1.0 - float(g["betas"][1]),
                           int(self.decoupled))
        clip = _lib.ptr(self._clip) if self._max_norm is not None else None
        st = _lib.stream_ptr()
        w = m.item_emb.weight.data
        _lib.check(lib.fbn_adam_table(_lib.ptr(w), _lib.ptr(self._m_item), _lib.ptr(self._v_item), _lib.ptr(m._item_grad),
                                      None if m._dense_table_grad else _lib.ptr(m._row_touched), w.shape[0], clip, C.byref(h), None, st), "fbn_adam_table")
        _lib.check(lib.fbn_adam_dense(_lib.ptr(m._flat), _lib.ptr(self._m_flat), _lib.ptr(self._v_flat), _lib.ptr(m._gflat),
                                      m._flat.numel(), clip, C.byref(h), None, st), "fbn_adam_dense")
        self._max_norm = None
        return None

    def zero_grad(self, set_to_none: bool = True):
        # gradients are overwritten (not accumulated) by every backward in fused mode
        for p in self.param_groups[0]["params"]:
            p.grad = None

    # ---- hooks of engine.TrainStep (CUDA-graph replay: hyper-parameters travel through a 12-float device array) ----------
    def _fill_hyper(self, h):
        """Advance the step count and write this step's values into the pinned staging tensor ``h`` (12 floats)."""
        g = self.param_groups[0]
        self._step += 1
        t = self._step
        lr, b1, b2 = float(g["lr"]), float(g["betas"][0]), float(g["betas"][1])
        h[0], h[1], h[2], h[3], h[4] = lr, b1, b2, float(g["eps"]), float(g["weight_decay"])
        h[5] = lr / (1.0 - b1 ** t)
        h[6] = math.sqrt(1.0 - b2 ** t)
        h[7] = float(t)
        h[8], h[9] = 1.0 - b1, 1.0 - b2          # evaluated in double like torch, rounded to fp32 by the store
        h[10] = (1.0 - lr * float(g["weight_decay"])) if self.decoupled else 0.0     # AdamW decay multiplier

    def _launch_update(self, lib, clip, hyper_dev, st, dense_table: bool):
        m = self.model
        w = m.item_emb.weight.data
        _lib.check(lib.fbn_adam_table(_lib.ptr(w), _lib.ptr(self._m_item), _lib.ptr(self._v_item), _lib.ptr(m._item_grad),
                                      None if dense_table else _lib.ptr(m._row_touched), w.shape[0], clip, None, hyper_dev, st),
                   "fbn_adam_table")
        _lib.check(lib.fbn_adam_dense(_lib.ptr(m._flat), _lib.ptr(self._m_flat), _lib.ptr(self._v_flat), _lib.ptr(m._gflat), m._flat.numel(),
                                      clip, None, hyper_dev, st), "fbn_adam_dense")


class FusedAdagrad(torch.optim.Optimizer):
    """torch.optim.Adagrad(lr, lr_decay, weight_decay, initial_accumulator_value, eps) on the CUDA path: the sorted-segment row
    gradients of the embedding backward fused into the Adagrad row update (BASELINE north_star (2)).  The reference builds Adam
    (src/train_fibinet.py:78), so this is an extension; `optimizer: adagrad` under ``honor_config`` selects it in
    src/train_fibinet.py.  With weight_decay == 0 only touched table rows are read and written (the update of an untouched row is
    the identity); the global clip coefficient is applied inside the kernels like in FusedAdam."""

    def __init__(self, model, lr=1e-2, lr_decay=0.0, weight_decay=0.0, initial_accumulator_value=0.0, eps=1e-10):
        from .model import MM_FiBiNET
        inner = model.module if hasattr(model, "module") else model
        if not isinstance(inner, MM_FiBiNET):
            raise TypeError("FusedAdagrad drives a ctr_recommendation_b200 MM_FiBiNET (pass the model, not parameters())")
        params = [p for n, p in inner.named_parameters() if not n.startswith("user_emb.")]
        super().__init__(params, dict(lr=lr, lr_decay=lr_decay, eps=eps, weight_decay=weight_decay,
                                      initial_accumulator_value=initial_accumulator_value))
        self.model = inner
        inner._fused_optimizer = self
        self._step = 0
        self._max_norm = None
        self._clip = None
        self._sum_flat = self._sum_item = None

    def _ensure_state(self):
        m = self.model
        m._ensure_flat()
        dev = m._flat.device
        if self._sum_flat is None or self._sum_flat.device != dev or self._sum_flat.numel() != m._flat.numel():
            init = float(self.param_groups[0]["initial_accumulator_value"])
            self._sum_flat = torch.full_like(m._flat, init)
            self._sum_item = torch.full_like(m.item_emb.weight.data, init)
            self._clip = torch.ones(2, dtype=torch.float32, device=dev)

    def accumulators(self):
        """{state_dict key: state_sum} views, for parity tests and checkpoints."""
        m = self.model
        out = {"item_emb.weight": self._sum_item}
        names = {id(p): n for n, p in m.named_parameters()}
        for (field, plist), (off, _) in zip(m._dense_params(), m._layout):
            o = off
            for p in plist:
                n = p.numel()
                out[names[id(p)]] = self._sum_flat[o:o + n].view(p.shape)
                o += (n + 3) // 4 * 4
        return out

    def state_dict(self):
        self._ensure_state()
        return {"step": self._step, "param_groups": [{k: v for k, v in g.items() if k != "params"} for g in self.param_groups],
                "sum_flat": self._sum_flat, "sum_item": self._sum_item}

    def load_state_dict(self, sd):
        self._ensure_state()
        self._step = int(sd["step"])
        for g, s in zip(self.param_groups, sd["param_groups"]):
            g.update(s)
        self._sum_flat.copy_(sd["sum_flat"])
        self._sum_item.copy_(sd["sum_item"])

    clip_grad_norm_ = FusedAdam.clip_grad_norm_
    zero_grad = FusedAdam.zero_grad

    def _hyper_struct(self):
        g = self.param_groups[0]
        return _lib.AdagradHyper(float(g["lr"]), float(g["lr_decay"]), float(g["eps"]), float(g["weight_decay"]), self._step)

    @torch.no_grad()
    def step(self, closure=None):
        if closure is not None:
            raise NotImplementedError("FusedAdagrad does not support closures")
        lib = _lib.load()
        self._ensure_state()
        m = self.model
        if m._item_grad is None:
            raise RuntimeError("FusedAdagrad.step() called before backward()")
        self._step += 1
        h = self._hyper_struct()
        clip = _lib.ptr(self._clip) if self._max_norm is not None else None
        st = _lib.stream_ptr()
        w = m.item_emb.weight.data
        _lib.check(lib.fbn_adagrad_table(_lib.ptr(w), _lib.ptr(self._sum_item), _lib.ptr(m._item_grad),
                                         None if m._dense_table_grad else _lib.ptr(m._row_touched), w.shape[0], clip, C.byref(h), None, st),
                   "fbn_adagrad_table")
        _lib.check(lib.fbn_adagrad_dense(_lib.ptr(m._flat), _lib.ptr(self._sum_flat), _lib.ptr(m._gflat), m._flat.numel(), clip,
                                         C.byref(h), None, st), "fbn_adagrad_dense")
        self._max_norm = None
        return None

    def _fill_hyper(self, h):
        g = self.param_groups[0]
        self._step += 1
        h.zero_()
        h[0] = float(g["lr"]) / (1.0 + (self._step - 1) * float(g["lr_decay"]))      # clr
        h[3], h[4], h[7] = float(g["eps"]), float(g["weight_decay"]), float(self._step)

    def _launch_update(self, lib, clip, hyper_dev, st, dense_table: bool):
        m = self.model
        w = m.item_emb.weight.data
        _lib.check(lib.fbn_adagrad_table(_lib.ptr(w), _lib.ptr(self._sum_item), _lib.ptr(m._item_grad),
                                         None if dense_table else _lib.ptr(m._row_touched), w.shape[0], clip, None, hyper_dev, st),
                   "fbn_adagrad_table")
        _lib.check(lib.fbn_adagrad_dense(_lib.ptr(m._flat), _lib.ptr(self._sum_flat), _lib.ptr(m._gflat), m._flat.numel(), clip,
                                         None, hyper_dev, st), "fbn_adagrad_dense")


def clip_grad_norm_(model_or_params, max_norm: float):
    """Drop-in for torch.nn.utils.clip_grad_norm_ in the training script: with a FusedAdam-driven
    model the clip is recorded on the device and applied inside the fused Adam kernels; anything
    else is forwarded to torch."""
    from .model import MM_FiBiNET
    inner = getattr(model_or_params, "module", model_or_params)
    if isinstance(inner, MM_FiBiNET) and inner._fused_optimizer is not None:
        return inner._fused_optimizer.clip_grad_norm_(max_norm)      # FusedAdam or FusedAdagrad
    params = inner.parameters() if isinstance(inner, torch.nn.Module) else model_or_params
    return torch.nn.utils.clip_grad_norm_(params, max_norm=max_norm)
