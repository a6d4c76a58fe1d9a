"""GPU probe: time of the K=128 / N=128 tcgen05 GEMM (the bilinear-transform shape) for contiguous operands."""
import ctypes as C, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ctr_recommendation_b200 import _lib
lib = _lib.load()
st = _lib.stream_ptr()
flush = torch.empty(160 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")
ms = C.c_float(0)
full = (1 << 64) - 1
for pers in (0, 1):
    lib.fbn_set_option(b"tc_persistent", pers)
    for (M, N, K, a_t, b_t) in [(262144, 128, 128, 0, 1), (65536, 128, 128, 0, 1), (262144, 128, 128, 0, 0), (128, 128, 262144, 1, 0),
                                (65536, 256, 512, 0, 1), (65536, 512, 256, 0, 0)]:
        A = torch.randn((K, M) if a_t else (M, K), device="cuda")
        B = torch.randn((N, K) if b_t else (K, N), device="cuda")
        Cc = torch.empty(M, N, device="cuda")
        for prec in (1, 2):
            n = lib.fbn_gemm_scratch_bytes(M, N, K, prec)
            scr = torch.empty(n, dtype=torch.uint8, device="cuda")
            _lib.check(lib.fbn_time_gemm(_lib.ptr(A), _lib.ptr(B), _lib.ptr(Cc), M, N, K, a_t, b_t, full, prec, _lib.ptr(scr), n,
                                         _lib.ptr(flush), flush.numel() * 4, 5, C.byref(ms), st))
            fl = 2.0 * M * N * K
            byt = (M * K + N * K) * (8 if prec == 1 else 2) + M * N * 4
            print(f"persistent={pers} prec={'tf32x3' if prec == 1 else 'bf16'} M{M} N{N} K{K} a_t{a_t} b_t{b_t}: {ms.value*1e3:8.1f} us  "
                  f"{fl/ms.value/1e9:7.1f} TFLOP/s  {byt/ms.value/1e6:7.0f} GB/s", flush=True)
