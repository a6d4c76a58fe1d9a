"""GPU probe: in-model stage timings of the forward pass (fbn_time_stage) under the runtime knobs of the tcgen05 path."""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ctr_recommendation_b200 import _lib, build_model
from oracle import synth
lib = _lib.load()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
prec = sys.argv[2] if len(sys.argv) > 2 else "tf32x3"
model = build_model({"precision": prec}, {"embedding_dim": 128}).cuda().train()
b, y = synth.make_batch(seed=1, batch=B, index_dtype=np.float64, edge_cases=False)
dev = {k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in b.items() if k != "user_id"}
flush = torch.empty(160 << 20, dtype=torch.uint8, device="cuda")
ms = C.c_float(0)
opts = [("tc_persistent", 0), ("tc_persistent", 1)]
for name, val in opts:
    lib.fbn_set_option(name.encode(), val)
    with torch.no_grad():
        model(dev)
    torch.cuda.synchronize()
    cur = model._cur
    P = model._params_struct()
    for stage in ("embed", "bil_gemm", "bil_pairs", "mlp1"):
        _lib.check(lib.fbn_time_stage(C.byref(P), C.byref(cur["bs"]), _lib.ptr(cur["ws"]), cur["ws"].numel(), stage.encode(), _lib.ptr(flush),
                                      flush.numel(), 10, C.byref(ms), _lib.stream_ptr()), "fbn_time_stage")
        print(f"B={B} {prec} {name}={val} {stage:10s} {ms.value * 1e3:8.1f} us", flush=True)
