"""GPU: fbn_gemm (every layout) against torch's fp32 matmul on the same device."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("a_t,b_t", [(False, True), (False, False), (True, False), (True, True)])
@pytest.mark.parametrize("M,N,K", [(300, 512, 2688), (4096, 256, 512), (128, 128, 1000), (512, 2688, 300), (77 * 4, 128, 128)])
def test_sgemm_layouts(a_t, b_t, M, N, K):
    from ctr_recommendation_b200.functional import gemm
    torch.backends.cuda.matmul.allow_tf32 = False
    g = torch.Generator(device="cuda").manual_seed(1)
    A = torch.randn(M, K, device="cuda", generator=g)
    Bm = torch.randn(K, N, device="cuda", generator=g)
    bias = torch.randn(N, device="cuda", generator=g)
    ref = (A.double() @ Bm.double() + bias.double()).float()
    a_in = A.t().contiguous() if a_t else A
    b_in = Bm.t().contiguous() if b_t else Bm
    out = gemm(a_in, b_in, bias, a_t=a_t, b_t=b_t, precision="fp32")
    err = (out - ref).abs().max().item() / ref.abs().max().item()
    assert err <= 2e-6, err


@pytest.mark.parametrize("precision,tol", [("tf32x3", 3e-6), ("bf16", 6e-3)])
@pytest.mark.parametrize("a_t,b_t", [(False, True), (False, False), (True, False)])
@pytest.mark.parametrize("M,N,K", [(128, 128, 32), (300, 512, 2688), (4096, 256, 512), (128, 128, 1000), (512, 2688, 300)])
def test_tcgen05_gemm_layouts(precision, tol, a_t, b_t, M, N, K):
    """tcgen05 path (TMA -> smem -> tcgen05.mma -> TMEM -> tcgen05.ld) against an fp64 matmul of the same fp32 inputs.
    tf32x3 must be fp32-grade; bf16 carries 8-bit mantissa operands."""
    from ctr_recommendation_b200.functional import gemm
    g = torch.Generator(device="cuda").manual_seed(2)
    A = torch.randn(M, K, device="cuda", generator=g)
    Bm = torch.randn(K, N, device="cuda", generator=g)
    bias = torch.randn(N, device="cuda", generator=g)
    ref = (A.double() @ Bm.double() + bias.double()).float()
    a_in = A.t().contiguous() if a_t else A
    b_in = Bm.t().contiguous() if b_t else Bm
    out = gemm(a_in, b_in, bias, a_t=a_t, b_t=b_t, precision=precision)
    torch.cuda.synchronize()
    err = (out - ref).abs().max().item() / ref.abs().max().item()
    assert err <= tol, err
