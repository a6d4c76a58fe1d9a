"""Differentiable wrappers of the stand-alone CUDA ops (module-level API of the reference:
SENetLayer / BilinearInteraction, src/model_fibinet.py:5-89).  No torch math: each op is one or two
calls into libfibinet_b200.so."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib


def _chk(t: torch.Tensor, what: str) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError(f"{what} is on {t.device}: ctr_recommendation_b200 has no CPU path")
    return t.to(torch.float32).contiguous()


class _SENetFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2):
        lib = _lib.load()
        x, w1, b1, w2, b2 = (_chk(t, "senet input") for t in (x, w1, b1, w2, b2))
        B, F, Dm = x.shape
        R = w1.shape[0]
        y = torch.empty_like(x)
        gate = torch.empty(B, F, dtype=torch.float32, device=x.device)
        _lib.check(lib.fbn_senet_fwd(_lib.ptr(x), _lib.ptr(w1), _lib.ptr(b1), _lib.ptr(w2), _lib.ptr(b2), B, F, Dm, R,
                                     _lib.ptr(y), _lib.ptr(gate), _lib.stream_ptr()), "fbn_senet_fwd")
        ctx.save_for_backward(x, gate, w1, b1, w2)
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = _lib.load()
        x, gate, w1, b1, w2 = ctx.saved_tensors
        B, F, Dm = x.shape
        R = w1.shape[0]
        dy = _chk(dy, "senet grad")
        dx = torch.empty_like(x)
        dw1, db1, dw2, db2 = (torch.empty_like(t) for t in (w1, b1, w2, torch.empty(F, device=x.device)))
        nbytes = lib.fbn_senet_scratch_bytes(B, F, R)
        scratch = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
        _lib.check(lib.fbn_senet_bwd(_lib.ptr(x), _lib.ptr(gate), _lib.ptr(w1), _lib.ptr(b1), _lib.ptr(w2), _lib.ptr(dy), B, F, Dm, R,
                                     _lib.ptr(dx), _lib.ptr(dw1), _lib.ptr(db1), _lib.ptr(dw2), _lib.ptr(db2), _lib.ptr(scratch),
                                     nbytes, _lib.stream_ptr()), "fbn_senet_bwd")
        return dx, dw1, db1, dw2, db2


def senet(x, w1, b1, w2, b2):
    """y = x * sigmoid(W2 relu(W1 mean_d(x) + b1) + b2)[..., None]"""
    return _SENetFn.apply(x, w1, b1, w2, b2)


class _BilinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, btype, precision, *ws):
        lib = _lib.load()
        x = _chk(x, "bilinear input")
        B, F, Dm = x.shape
        w = torch.stack([_chk(t, "bilinear weight") for t in ws]).contiguous()
        P = F * (F - 1) // 2
        out = torch.empty(B, P, Dm, dtype=torch.float32, device=x.device)
        nbytes = lib.fbn_bilinear_scratch_bytes(B, F, Dm, btype)
        scratch = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
        _lib.check(lib.fbn_bilinear_fwd(_lib.ptr(x), _lib.ptr(w), btype, B, F, Dm, _lib.ptr(out), _lib.ptr(scratch), nbytes,
                                        precision, _lib.stream_ptr()), "fbn_bilinear_fwd")
        ctx.save_for_backward(x, w)
        ctx.btype, ctx.precision, ctx.nw = btype, precision, len(ws)
        return out

    @staticmethod
    def backward(ctx, dp):
        lib = _lib.load()
        x, w = ctx.saved_tensors
        B, F, Dm = x.shape
        dp = _chk(dp, "bilinear grad")
        dx = torch.empty_like(x)
        dw = torch.empty_like(w)
        nbytes = lib.fbn_bilinear_scratch_bytes(B, F, Dm, ctx.btype)
        scratch = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
        _lib.check(lib.fbn_bilinear_bwd(_lib.ptr(x), _lib.ptr(w), _lib.ptr(dp), ctx.btype, B, F, Dm, _lib.ptr(dx), _lib.ptr(dw),
                                        _lib.ptr(scratch), nbytes, ctx.precision, _lib.stream_ptr()), "fbn_bilinear_bwd")
        return (dx, None, None) + tuple(dw[i] for i in range(ctx.nw))


def bilinear(x, weights, bilinear_type="all", precision="fp32"):
    """(B,F,D) -> (B,F(F-1)/2,D) pairwise bilinear products in the reference's pair order."""
    if bilinear_type not in _lib.BILINEAR_TYPES:
        raise ValueError("bilinear_type must be 'all' or 'each'")
    return _BilinearFn.apply(x, _lib.BILINEAR_TYPES[bilinear_type], _lib.PRECISIONS[precision], *weights)


def gemm(a: torch.Tensor, b: torch.Tensor, bias=None, a_t=False, b_t=False, precision="fp32") -> torch.Tensor:
    """C = op(A) op(B) (+bias) through fbn_gemm (tests / benches)."""
    lib = _lib.load()
    a, b = _chk(a, "A"), _chk(b, "B")
    M = a.shape[1] if a_t else a.shape[0]
    K = a.shape[0] if a_t else a.shape[1]
    N = b.shape[0] if b_t else b.shape[1]
    c = torch.empty(M, N, dtype=torch.float32, device=a.device)
    prec = _lib.GEMM_PRECISIONS[precision]
    nbytes = lib.fbn_gemm_scratch_bytes(M, N, K, prec)
    scratch = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=a.device)
    _lib.check(lib.fbn_gemm(_lib.ptr(a), _lib.ptr(b), _lib.ptr(bias), _lib.ptr(c), M, N, K, a.shape[1], b.shape[1], N, int(a_t),
                            int(b_t), prec, _lib.ptr(scratch), nbytes, _lib.stream_ptr()), "fbn_gemm")
    return c
