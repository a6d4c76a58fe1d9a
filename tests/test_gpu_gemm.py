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


@pytest.mark.parametrize("precision,tol", [("tf32x3", 3e-6), ("f16x3", 3e-6), ("bf16", 6e-3)])
@pytest.mark.parametrize("a_t,b_t", [(False, True), (False, False), (True, False)])
@pytest.mark.parametrize("M,N,K", [(128, 128, 32), (300, 512, 2688), (4096, 256, 512), (128, 128, 1000), (512, 2688, 300)])
def test_tcgen05_gemm_layouts(precision, tol, a_t, b_t, M, N, K):
    """tcgen05 path (TMA -> smem -> tcgen05.mma -> TMEM -> tcgen05.ld) against an fp64 matmul of the same fp32 inputs.
    tf32x3 and f16x3 (three fp16 passes under one power-of-two scale per operand) must be fp32-grade; bf16 carries 8-bit
    mantissa operands."""
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


@pytest.fixture
def knobs():
    from ctr_recommendation_b200 import _lib
    lib = _lib.load()
    yield lib
    lib.fbn_set_option(b"tc_pair", 1)
    lib.fbn_set_option(b"tc_persistent", 0)
    lib.fbn_set_option(b"tc_pair_persistent", 1)


@pytest.mark.parametrize("pair,persistent,pair_persistent", [(1, 0, 1), (1, 0, 0), (0, -1, 1), (0, 1, 1), (1, 1, 1)])
@pytest.mark.parametrize("precision,tol", [("tf32x3", 3e-6), ("f16x3", 3e-6), ("bf16", 6e-3)])
@pytest.mark.parametrize("a_t,b_t,M,N,K", [(False, True, 40000, 256, 512), (False, False, 40000, 512, 256), (False, False, 39000, 128, 128),
                                           (True, False, 512, 384, 5000), (False, True, 300, 512, 2688),
                                           (False, False, 20001, 2688, 512),      # 21 column blocks: the last pair tile is half dead
                                           (True, False, 512, 2688, 20000)])
def test_tcgen05_kernel_variants(knobs, pair, persistent, pair_persistent, precision, tol, a_t, b_t, M, N, K):
    """Every tcgen05 kernel variant the dispatcher can pick -- CTA pairs (persistent tile loop over 74 clusters, or one tile per
    cluster), one tile per CTA, the single-CTA persistent tile loop (forced on / off / chosen by the heuristic at these sizes)
    -- against an fp64 matmul."""
    from ctr_recommendation_b200.functional import gemm
    knobs.fbn_set_option(b"tc_pair", pair)
    knobs.fbn_set_option(b"tc_persistent", persistent)
    knobs.fbn_set_option(b"tc_pair_persistent", pair_persistent)
    g = torch.Generator(device="cuda").manual_seed(3)
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


@pytest.mark.parametrize("precision", ["tf32x3", "f16x3", "bf16"])
def test_model_step_identical_under_kernel_variants(knobs, precision):
    """The train forward/backward of a batch large enough for the persistent-kernel heuristic (B = 40000: 4 x 313 bilinear tiles)
    gives the same probabilities and gradients whichever GEMM kernel variant runs (same MMA order per tile -> bitwise)."""
    from gpu_common import make_model, to_dev, named_grads
    from oracle import synth
    B = 40000
    batch, labels = synth.make_batch(seed=31, batch=B, id_dist="uniform", index_dtype=np.float64, edge_cases=False)
    outs = []
    for pair, persistent, pp in ((1, 0, 1), (1, 0, 0), (1, -1, 1), (0, 1, 1)):
        knobs.fbn_set_option(b"tc_pair", pair)
        knobs.fbn_set_option(b"tc_persistent", persistent)
        knobs.fbn_set_option(b"tc_pair_persistent", pp)
        model = make_model(train=True, precision=precision)
        model.dropout_p = 0.0
        y = model(to_dev(batch))
        torch.nn.BCELoss()(y, torch.from_numpy(labels).cuda()).backward()
        outs.append((y.detach().clone(), {k: torch.from_numpy(v) for k, v in named_grads(model).items()}))
        del model
    for y, g in outs[1:]:
        assert torch.allclose(y, outs[0][0], rtol=0, atol=2e-6 if precision != "bf16" else 2e-3)
        for k in g:
            ref = outs[0][1][k]
            err = (g[k] - ref).abs().max().item() / max(ref.abs().max().item(), 1e-30)
            assert err <= (1e-5 if precision != "bf16" else 5e-2), (k, err)


@pytest.mark.parametrize("a_scale,b_scale", [(1.0, 1.0), (3e-9, 0.04), (7e4, 2e-6), (1e-30, 1e20)])
@pytest.mark.parametrize("a_t,b_t,M,N,K", [(False, True, 1000, 512, 2688), (True, False, 512, 640, 9000), (False, False, 20000, 512, 256)])
def test_f16x3_scales(a_t, b_t, M, N, K, a_scale, b_scale):
    """FBN_PREC_F16X3 on operands far outside the fp16 range (gradient-sized, weight-sized, huge): the per-tensor power-of-two
    scale found by the amax pass keeps the result fp32-grade, and a block of columns 10^4 times smaller than the rest of its
    tensor (the pair blocks of the MLP input next to the LayerNorm field) is still resolved to 1e-5 of ITS OWN magnitude in a
    weight-gradient-shaped product."""
    from ctr_recommendation_b200.functional import gemm
    g = torch.Generator(device="cuda").manual_seed(5)
    A = torch.randn(M, K, device="cuda", generator=g) * a_scale
    Bm = torch.randn(K, N, device="cuda", generator=g) * b_scale
    Bm[:, : N // 2] *= 1e-4                                   # half of the output columns come from a much smaller block
    ref = A.double() @ Bm.double()
    a_in = A.t().contiguous() if a_t else A
    b_in = Bm.t().contiguous() if b_t else Bm
    out = gemm(a_in, b_in, None, a_t=a_t, b_t=b_t, precision="f16x3").double()
    torch.cuda.synchronize()
    for cols in (slice(0, N // 2), slice(N // 2, N)):
        err = (out[:, cols] - ref[:, cols]).abs().max().item() / ref[:, cols].abs().max().item()
        assert err <= (1e-5 if cols.start == 0 else 3e-6), (cols, err)


def test_f16x3_pack_bits_match_restatement():
    """The operand format itself, byte for byte: after an fbn_gemm call in f16x3 the scratch holds both packed operands --
    [hi (rows, pitch) fp16 | lo | record {s, 1/s, amax, ...}] at the next 1024-byte boundary -- and they must equal the numpy
    restatement (oracle/f16x3_numpy.py: amax -> power-of-two scale, hi = rn(s x), lo = rn(s x - hi)) bit for bit, zero padding
    included; the product agrees with the restatement's exact-product sum to accumulation rounding (3e-6)."""
    import ctypes as C
    from ctr_recommendation_b200 import _lib
    from oracle import f16x3_numpy as F
    lib = _lib.load()
    M, N, K = 300, 128, 204                                   # K % 8 != 0: 4 zero padding columns per row
    pitch = (K + 7) // 8 * 8
    g = torch.Generator(device="cuda").manual_seed(11)
    A = torch.randn(M, K, device="cuda", generator=g) * 3e-5  # gradient-sized
    A[:, :16] *= 1e-5                                         # a block near the fp16 subnormal floor after scaling
    W = torch.randn(N, K, device="cuda", generator=g) * 0.05  # nn.Linear weight (N, K)
    Cm = torch.empty(M, N, device="cuda")
    prec = _lib.PRECISIONS["f16x3"]
    nbytes = lib.fbn_gemm_scratch_bytes(M, N, K, prec)
    scratch = torch.zeros(nbytes + 1024, dtype=torch.uint8, device="cuda")
    _lib.check(lib.fbn_gemm(_lib.ptr(A), _lib.ptr(W), None, _lib.ptr(Cm), M, N, K, K, K, N, 0, 1, prec, _lib.ptr(scratch), nbytes,
                            _lib.stream_ptr()), "fbn_gemm")
    torch.cuda.synchronize()
    raw = scratch.cpu().numpy()
    base = scratch.data_ptr()

    def region(start, rows):
        off = start + (-(base + start)) % 1024
        hi = raw[off: off + rows * pitch * 2].view(np.float16).reshape(rows, pitch)
        lo = raw[off + rows * pitch * 2: off + rows * pitch * 4].view(np.float16).reshape(rows, pitch)
        rec = raw[off + rows * pitch * 4: off + rows * pitch * 4 + 16].view(np.float32)
        return hi, lo, rec

    region_bytes = lambda rows: rows * pitch * 4 + 1024 + (16 + 1024) * 4      # packed_bytes(rows, K, f16x3)
    for name, x, start, rows in (("A", A, 0, M), ("W", W, region_bytes(M), N)):
        xn = x.cpu().numpy()
        hi, lo, rec = region(start, rows)
        ohi, olo, s = F.split(xn)
        assert rec[0] == np.float32(s) and rec[1] == np.float32(1.0 / s) and rec[2] == np.abs(xn).max(), (name, rec[:3], s)
        assert np.array_equal(hi[:, :K].view(np.uint16), ohi.view(np.uint16)), name
        assert np.array_equal(lo[:, :K].view(np.uint16), olo.view(np.uint16)), name
        assert not hi[:, K:].view(np.uint16).any() and not lo[:, K:].view(np.uint16).any(), name
    ref = F.matmul(A.cpu().numpy(), W.cpu().numpy().T)
    got = Cm.cpu().numpy()
    assert np.abs(got - ref).max() / np.abs(ref).max() <= 3e-6      # TMEM accumulation (measured 1.1e-6), same bar as the layout tests


def test_f16x3_zero_operand():
    """an all-zero operand (amax = 0) is left unscaled and gives an exact zero product + bias"""
    from ctr_recommendation_b200.functional import gemm
    A = torch.zeros(256, 128, device="cuda")
    W = torch.randn(128, 128, device="cuda")
    bias = torch.randn(128, device="cuda")
    out = gemm(A, W, bias, a_t=False, b_t=True, precision="f16x3")
    assert torch.equal(out, bias.expand(256, 128))


@pytest.mark.parametrize("persistent", [-1, 1])
@pytest.mark.parametrize("M,N,K", [(128, 128, 32), (300, 512, 2688), (4096, 256, 512), (40000, 128, 128), (1000, 384, 1000), (256, 512, 4096),
                                   (20000, 512, 1056)])        # last: large enough for the CTA-pair kernel
def test_tf32x2_gemm(knobs, persistent, M, N, K):
    """FBN_PREC_TF32X2 (hi*hi as one tf32 MMA + the two correction terms as bf16 MMAs = 2 tensor-pass equivalents instead of 3)
    against an fp64 matmul: K-major x K-major operands, fp32-grade within 4e-6 (numpy simulation of the scheme: 1.3-1.5e-6)."""
    from ctr_recommendation_b200.functional import gemm
    knobs.fbn_set_option(b"tc_persistent", persistent)
    g = torch.Generator(device="cuda").manual_seed(4)
    A = torch.relu(torch.randn(M, K, device="cuda", generator=g))            # ReLU-like activations
    W = torch.randn(N, K, device="cuda", generator=g) * 0.05                 # nn.Linear weight (N, K)
    bias = torch.randn(N, device="cuda", generator=g)
    ref = (A.double() @ W.double().t() + bias.double()).float()
    out = gemm(A, W, bias, a_t=False, b_t=True, precision="tf32x2")
    torch.cuda.synchronize()
    err = (out - ref).abs().max().item() / ref.abs().max().item()
    assert err <= 4e-6, err
    x3 = gemm(A, W, bias, a_t=False, b_t=True, precision="tf32x3")
    assert (x3 - ref).abs().max().item() / ref.abs().max().item() <= 3e-6


def test_tf32x2_rejects_other_layouts(knobs):
    from ctr_recommendation_b200.functional import gemm
    from ctr_recommendation_b200._lib import FibinetCudaError
    A = torch.randn(256, 128, device="cuda")
    B = torch.randn(128, 128, device="cuda")
    with pytest.raises(FibinetCudaError):
        gemm(A, B, None, a_t=False, b_t=False, precision="tf32x2")          # B stored (K,N) would be MN-major
