"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the FBN_PREC_F16X3 operand format and product (csrc/gemm_tc.cu, common.cuh).

Only tests/, tools/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this; the product path never does.

Parity status: *extension, unpinned by the reference* (the reference multiplies in fp32 through torch, src/model_fibinet.py:72,126,130;
f16x3 is this repository's way of reproducing those fp32 products on the fp16 tensor-core path).  What is restated here, bit for
bit, is what the CUDA kernels define:

  * scale_from_amax : s = 2^(14 - floor(log2 amax)), i.e. amax * s in [2^14, 2^15) (fp16 overflows at 65504); exponent clamped to
                      [-120, 120]; 1 for an all-zero or non-finite tensor                    (common.cuh f16x3_scale)
  * split           : hi = fp16_rn(s x), lo = fp16_rn(s x - hi); s x is exact (power of two), s x - hi is exact in fp32
                                                                                              (common.cuh store_f16x3_4)
  * matmul          : (Ah Bh + Ah Bl + Al Bh) / (sa sb); each fp16 x fp16 product is exact in fp32 (11 + 11 significand bits); the
                      tensor core's accumulation order is not restated -- products are summed in float64 and rounded once

tests/test_oracle_f16x3.py pins the restatement (scale window, representation error bound, product error against float64);
tests/test_gpu_gemm.py::test_f16x3_pack_bits_match_restatement compares the device's packed bytes and scale record with it.
"""
from __future__ import annotations

import numpy as np


def scale_from_amax(amax: float) -> float:
    a = np.float32(amax)
    bits = int(a.view(np.uint32))
    e = ((bits >> 23) & 0xFF) - 127
    if bits == 0 or e == 128:
        return 1.0
    k = max(-120, min(120, 14 - e))
    return float(np.ldexp(1.0, k))


def split(x: np.ndarray):
    """-> (hi, lo, s): hi, lo float16 arrays, s the tensor's power-of-two scale"""
    x = np.ascontiguousarray(x, dtype=np.float32)
    amax = float(np.abs(x).max()) if x.size else 0.0
    s = scale_from_amax(amax)
    xs = x * np.float32(s)
    hi = xs.astype(np.float16)
    lo = (xs - hi.astype(np.float32)).astype(np.float16)
    return hi, lo, s


def matmul(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """C = A @ B in the f16x3 scheme (A: (M,K), B: (K,N)); exact products, float64 accumulation, one rounding to fp32"""
    ah, al, sa = split(a)
    bh, bl, sb = split(b)
    d = np.float64
    acc = ah.astype(d) @ bh.astype(d) + ah.astype(d) @ bl.astype(d) + al.astype(d) @ bh.astype(d)
    return ((acc / sa) / sb).astype(np.float32)
