"""GPU probe: short-K GEMMs of the step under (tc_pair, tc_persistent) -- is the overlapped epilogue of the persistent single-CTA
kernel worth more than the halved operand traffic of the CTA-pair kernel when K <= 512?"""
import ctypes as C, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ctr_recommendation_b200 import _lib
lib = _lib.load()
st = _lib.stream_ptr()
flush = torch.empty(160 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")
ms = C.c_float(0)
full = (1 << 64) - 1
live = 0
for blk in list(range(1, 6)) + list(range(11, 21)):
    live |= 1 << blk
Bs = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
shapes = [("dgrad1 (B x 2688 x 512, W (K,N))", Bs, 2688, 512, 0, 0, full, live),
          ("mlp2 fwd (B x 256 x 512, W (N,K))", Bs, 256, 512, 0, 1, full, full),
          ("dgrad2 (B x 512 x 256, W (K,N))", Bs, 512, 256, 0, 0, full, full),
          ("mlp1 fwd (B x 512 x 2688 live K)", Bs, 512, 2688, 0, 1, live, full)]
for name, M, N, K, a_t, b_t, kmask, nmask in shapes:
    A = torch.randn((K, M) if a_t else (M, K), device="cuda")
    B = torch.randn((N, K) if b_t else (K, N), device="cuda")
    Cc = torch.empty(M, N, device="cuda")
    for prec in (1, 2, 3):
        if prec == 3 and not (a_t == 0 and b_t == 1):
            continue
        n = lib.fbn_gemm_scratch_bytes(M, N, K, prec)
        scr = torch.empty(n, dtype=torch.uint8, device="cuda")
        for pair, pers in ((1, 0), (0, 0), (0, 1)):
            lib.fbn_set_option(b"tc_pair", pair)
            lib.fbn_set_option(b"tc_persistent", pers)
            _lib.check(lib.fbn_time_gemm(_lib.ptr(A), _lib.ptr(B), _lib.ptr(Cc), M, N, K, a_t, b_t, kmask, prec, _lib.ptr(scr), n,
                                         _lib.ptr(flush), flush.numel() * 4, 6, C.byref(ms), st))
            print(f"{name:36s} { {1: 'tf32x3', 2: 'bf16  ', 3: 'tf32x2'}[prec] } pair={pair} persistent={pers}: {ms.value * 1e3:8.1f} us "
                  f"{2.0 * M * N * K / ms.value / 1e9:7.1f} TFLOP/s(full shape)", flush=True)
