"""CPU study: how accurate is a split-operand tensor-core GEMM scheme *inside the model*?

Every matmul of the numpy oracle (forward and backward) is replaced by an emulation of one scheme

    tf32x3 : Ah*Bh + Ah*Bl + Al*Bh, hi = x & 0xffffe000 (tf32), lo = x - hi (truncated to tf32 by the tensor core)
    tf32x2 : Ah*Bh (tf32) + bf16(Ah)*bf16(Bl) + bf16(Al)*bf16(Bh)
    f16x3  : one power-of-two scale per tensor (amax -> [2^14, 2^15)), h = fp16_rn(s x), l = fp16_rn(s x - h) ; (hh + hl + lh) / (sa sb):
             the same 22 operand bits as tf32x3 at the bf16 MMA rate and half the operand bytes
    bf16x3 : h = bf16_rn(x), l = bf16_rn(x - h) ; hh + hl + lh   (three kind::f16 passes at the bf16 rate)
    bf16   : one bf16 pass

with exact (fp64) accumulation of the products, rounded to fp32 once -- i.e. the operand-representation error only,
which is what separates the schemes (accumulation order noise is the same ~1e-7 for all of them).  The result is compared
with the fp64 oracle on probabilities, logits and all 21 gradients (max|a-b| / max|b| per tensor).

    python tools/split_precision_sim.py [B]
    python tools/split_precision_sim.py [B] headroom     # f16x3 only, with every tensor's scale deliberately too small by 2^-h

Scale sensitivity of f16x3 (B = 1024; h = 0 is the kernels' choice, h = 12 puts amax at 4..8):  h = 0, 4, 8, 12, 14 all give
logit 5.2e-7 .. 6.5e-7 and a worst gradient of 4.2e-6 .. 5.7e-6 (senet.excitation.*): the result does not depend on the scale
within 2^12.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import f16x3_numpy as F16  # noqa: E402
from oracle import fibinet_numpy as O  # noqa: E402
from oracle import synth  # noqa: E402


def bf16_rn(x):
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(np.float32)


def tf32_trunc(x):
    return (np.ascontiguousarray(x, dtype=np.float32).view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


SHIFT = 0


def f16_split(x, shift=None):
    """fp16 hi|lo split under one power-of-two scale per tensor: the kernels' rule (oracle/f16x3_numpy.py: amax * s in [2^14, 2^15)),
    optionally mis-scaled by 2^-shift to probe how much the result depends on the scale"""
    shift = SHIFT if shift is None else shift
    amax = float(np.abs(x).max())
    s = F16.scale_from_amax(amax) * 2.0 ** -shift
    xs = x.astype(np.float32) * np.float32(s)
    hi = xs.astype(np.float16).astype(np.float32)
    lo = (xs - hi).astype(np.float16).astype(np.float32)
    return hi, lo, s


SCHEME = "fp32"


def split_mm(a, b):
    a = np.asarray(a, dtype=np.float32)
    b = np.asarray(b, dtype=np.float32)
    d = np.float64
    if SCHEME == "fp32":
        return (a.astype(d) @ b.astype(d)).astype(np.float32)
    if SCHEME == "bf16":
        return (bf16_rn(a).astype(d) @ bf16_rn(b).astype(d)).astype(np.float32)
    if SCHEME == "f16x3":
        (ah, al, sa), (bh, bl, sb) = f16_split(a), f16_split(b)
        return ((ah.astype(d) @ bh.astype(d) + ah.astype(d) @ bl.astype(d) + al.astype(d) @ bh.astype(d)) / (sa * sb)).astype(np.float32)
    if SCHEME == "bf16x3":
        ah, bh = bf16_rn(a), bf16_rn(b)
        al, bl = bf16_rn(a - ah), bf16_rn(b - bh)
        return (ah.astype(d) @ bh.astype(d) + ah.astype(d) @ bl.astype(d) + al.astype(d) @ bh.astype(d)).astype(np.float32)
    ah, bh = tf32_trunc(a), tf32_trunc(b)
    al, bl = a - ah, b - bh
    if SCHEME == "tf32x3":
        al, bl = tf32_trunc(al), tf32_trunc(bl)
        return (ah.astype(d) @ bh.astype(d) + ah.astype(d) @ bl.astype(d) + al.astype(d) @ bh.astype(d)).astype(np.float32)
    if SCHEME == "tf32x2":
        return (ah.astype(d) @ bh.astype(d) + bf16_rn(ah).astype(d) @ bf16_rn(bl).astype(d)
                + bf16_rn(al).astype(d) @ bf16_rn(bh).astype(d)).astype(np.float32)
    raise ValueError(SCHEME)


class Q(np.ndarray):
    """ndarray whose @ is the emulated tensor-core product; propagates through ufuncs / views."""
    __array_priority__ = 100

    def __matmul__(self, other):
        return split_mm(self, other).view(Q)

    def __rmatmul__(self, other):
        return split_mm(other, self).view(Q)


def run(P, batch, labels, masks, scheme, gates=None):
    global SCHEME
    SCHEME = scheme
    Pq = {k: (np.array(v).view(Q) if v.dtype.kind == "f" else np.array(v)) for k, v in P.items()}
    bq = dict(batch)
    bq["item_emb_d128"] = np.array(batch["item_emb_d128"]).view(Q)
    prob, cache = O.forward(Pq, bq, train=True, masks=masks, dtype=np.float32, update_running=False, relu_gates=gates)
    _, dprob = O.bce_loss(np.asarray(prob), labels)
    G = O.backward(Pq, cache, dprob.view(Q))
    return np.asarray(prob), np.asarray(cache["logit"]), {k: np.asarray(v) for k, v in G.items()}, cache


def rel(a, b):
    b = np.asarray(b, np.float64)
    return float(np.abs(np.asarray(a, np.float64) - b).max() / max(np.abs(b).max(), 1e-30))


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
    P = synth.make_weights(7)
    batch, labels = synth.make_batch(101, B, id_dist="zipf")
    masks = synth.make_dropout_masks(5, B)
    P64 = {k: v.copy() for k, v in P.items()}
    p64, c64 = O.forward(P64, batch, train=True, masks=masks, dtype=np.float64, update_running=False)
    _, dp64 = O.bce_loss(p64, labels, np.float64)
    G64 = O.backward(P64, c64, dp64)
    global SHIFT
    sweep = len(sys.argv) > 2 and sys.argv[2] == "headroom"
    for scheme in ([("f16x3", h) for h in (0, 4, 8, 12, 14)] if sweep else ("fp32", "tf32x3", "f16x3", "tf32x2", "bf16x3", "bf16")):
        if sweep:
            scheme, SHIFT = scheme
            print(f"scale x 2^-{SHIFT}: ", end="")
        p, logit, G, c = run(P, batch, labels, masks, scheme)
        flips1 = int(((c["Y1"] > 0) != (c64["Y1"] > 0)).sum())
        flips2 = int(((c["Y2"] > 0) != (c64["Y2"] > 0)).sum())
        if flips1 + flips2:    # compare the gradients under the fp64 run's ReLU decisions (a flipped decision is not an operand error)
            _, _, G, _ = run(P, batch, labels, masks, scheme, gates=(c64["Y1"] > 0, c64["Y2"] > 0))
        errs = {k: rel(G[k], G64[k]) for k in G if k not in ("mlp.0.bias", "mlp.4.bias")}   # pre-BatchNorm biases: exact gradient 0
        worst = sorted(errs.items(), key=lambda kv: -kv[1])[:4]
        print(f"{scheme:7s} prob {rel(p, p64):.2e} logit {rel(logit, c64['logit']):.2e} relu flips {flips1}+{flips2}  "
              f"grad max {max(errs.values()):.2e}  worst: " + ", ".join(f"{k} {v:.1e}" for k, v in worst), flush=True)


if __name__ == "__main__":
    main()
