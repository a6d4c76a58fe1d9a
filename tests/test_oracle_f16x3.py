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
