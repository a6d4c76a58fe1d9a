"""CPU: the numpy restatement of the f16x3 operand format (oracle/f16x3_numpy.py) -- scale window, representation error, product
error against float64.  (The CUDA kernels are compared with it bit for bit in tests/test_gpu_gemm.py.)"""
import numpy as np
import pytest

from oracle import f16x3_numpy as F


@pytest.mark.parametrize("amax", [1.0, 0.75, 65504.0, 3e-9, 7e4, 1e-30, 2.0 ** -20, 1e20, np.float32(1.1754944e-38), 1e-42])
def test_scale_window(amax):
    s = F.scale_from_amax(amax)
    m, e = np.frexp(s)
    assert m == 0.5 and -120 <= e - 1 <= 120            # an exact power of two whose inverse is a normal fp32 number too
    if 2.0 ** -106 <= amax <= 2.0 ** 120:
        assert 2.0 ** 14 <= float(np.float32(amax)) * s < 2.0 ** 15     # never overflows fp16 (65504), top of its range used


def test_scale_of_degenerate_tensors():
    assert F.scale_from_amax(0.0) == 1.0 and F.scale_from_amax(np.inf) == 1.0 and F.scale_from_amax(np.nan) == 1.0
    hi, lo, s = F.split(np.zeros((3, 8), np.float32))
    assert s == 1.0 and not hi.any() and not lo.any()


@pytest.mark.parametrize("scale", [1.0, 3e-9, 7e4, 1e-30, 1e20])
def test_split_error_bound(scale):
    """|x - (hi + lo) / s| <= max(2^-22 |x|, 2^-25 / s): full 22-bit relative precision down to 2^-28 of the tensor's maximum (fp16
    keeps 11 significand bits twice), an absolute floor of 2^-39 amax below that."""
    rng = np.random.default_rng(0)
    x = (rng.standard_normal((64, 256)) * scale).astype(np.float32)
    x[:, :32] *= 1e-6                                     # a block far below the tensor's maximum
    hi, lo, s = F.split(x)
    assert np.isfinite(hi.astype(np.float32)).all() and np.abs(hi.astype(np.float32)).max() < 2.0 ** 15
    rec = (hi.astype(np.float64) + lo.astype(np.float64)) / s
    err = np.abs(rec - x.astype(np.float64))
    bound = np.maximum(2.0 ** -22 * np.abs(x.astype(np.float64)), 2.0 ** -25 / s)
    assert (err <= bound).all()


@pytest.mark.parametrize("a_scale,b_scale", [(1.0, 1.0), (3e-9, 0.04), (7e4, 2e-6)])
def test_matmul_is_fp32_grade(a_scale, b_scale):
    rng = np.random.default_rng(1)
    a = (rng.standard_normal((96, 2688)) * a_scale).astype(np.float32)
    b = (rng.standard_normal((2688, 64)) * b_scale).astype(np.float32)
    ref = a.astype(np.float64) @ b.astype(np.float64)
    got = F.matmul(a, b)
    assert np.abs(got - ref).max() / np.abs(ref).max() <= 1e-6
    plain = (a @ b)                                       # numpy's own fp32 product for scale
    assert np.abs(got - ref).max() <= 4 * max(np.abs(plain - ref).max(), 1e-7 * np.abs(ref).max())


def test_model_level_emulation_is_fp32_grade():
    """The whole FiBiNET train step of the numpy oracle with EVERY matmul (forward and backward) replaced by the exact emulation of
    the f16x3 operand split (tools/split_precision_sim.py, the kernels' scale rule): logits and all gradients stay within the
    north_star tolerance (1e-5 relative) of the fp64 oracle -- and so does the same step with every scale deliberately 2^-12 too
    small.  bf16x3 (bf16 hi|lo, the other 3-pass candidate at the bf16 rate) does not."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import split_precision_sim as sim
    from oracle import fibinet_numpy as O, synth
    B = 512
    P = synth.make_weights(7)
    batch, labels = synth.make_batch(101, B, id_dist="zipf")
    masks = synth.make_dropout_masks(5, B)
    p64, c64 = O.forward({k: v.copy() for k, v in P.items()}, batch, train=True, masks=masks, dtype=np.float64, update_running=False)
    _, dp64 = O.bce_loss(p64, labels, np.float64)
    G64 = O.backward(P, c64, dp64)
    gates = (c64["Y1"] > 0, c64["Y2"] > 0)      # a flipped ReLU decision is not an operand error
    skip = ("mlp.0.bias", "mlp.4.bias")         # pre-BatchNorm biases: the exact gradient is 0
    worst = {}
    for scheme, shift in (("f16x3", 0), ("f16x3", 12), ("bf16x3", 0)):
        sim.SHIFT = shift
        try:
            _, logit, G, _ = sim.run(P, batch, labels, masks, scheme, gates=gates)
        finally:
            sim.SHIFT = 0
        worst[(scheme, shift)] = (sim.rel(logit, c64["logit"]), max(sim.rel(G[k], G64[k]) for k in G if k not in skip))
    for key in (("f16x3", 0), ("f16x3", 12)):
        assert worst[key][0] <= 2e-6 and worst[key][1] <= 1e-5, (key, worst[key])
    assert worst[("bf16x3", 0)][1] > 1e-5, worst
