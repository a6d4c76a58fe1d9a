"""GPU probe for the tcgen05 GEMM: small shapes first, prints errors (run under `timeout`)."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ctr_recommendation_b200.functional import gemm
torch.manual_seed(0)
from ctr_recommendation_b200 import _lib
if "--no-pair" in sys.argv:
    _lib.load().fbn_set_option(b"tc_pair", 0)
import time
for prec in ("tf32x3", "bf16"):
    for (M, N, K, a_t, b_t) in [(128, 128, 32, False, True), (128, 128, 64, False, True), (128, 128, 256, False, True),
                                (256, 256, 512, False, True), (300, 512, 2688, False, True), (4096, 512, 2688, False, True),
                                (300, 512, 2688, False, False), (512, 2688, 300, True, False), (256, 128, 64, True, True), (4096, 2688, 512, False, False), (128, 128, 16384, True, False), (16384, 512, 2688, False, True), (512, 2688, 16384, True, False), (16384, 2688, 512, False, False)]:
        A = torch.randn(M, K, device="cuda"); B = torch.randn(K, N, device="cuda")
        ref = (A.double() @ B.double()).float()
        a_in = A.t().contiguous() if a_t else A
        b_in = B.t().contiguous() if b_t else B
        out = gemm(a_in, b_in, None, a_t=a_t, b_t=b_t, precision=prec)
        torch.cuda.synchronize()
        err = (out - ref).abs().max().item() / ref.abs().max().item()
        e32 = ((A @ B) - ref).abs().max().item() / ref.abs().max().item()
        t0 = time.time()
        for _ in range(5):
            gemm(a_in, b_in, None, a_t=a_t, b_t=b_t, precision=prec)
        torch.cuda.synchronize()
        ms = (time.time() - t0) / 5 * 1e3
        print(f"[{ms:.3f} ms incl. packing] ", end="")
        print(f"{prec} M{M} N{N} K{K} a_t{int(a_t)} b_t{int(b_t)}: rel err {err:.3e} (torch fp32 {e32:.1e}) nan={torch.isnan(out).any().item()}", flush=True)
