"""Live per-stage device times of one eager train step (fbn_set_option("stage_events") + fbn_stage_report): the main stream's
critical path, with the leaf gradients running on the library's side stream.   python tools/stage_report.py [B] [precision] [id_dist]"""
import collections, ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ctr_recommendation_b200 import _lib, build_model, FusedAdam, clip_grad_norm_
from oracle import synth
lib = _lib.load()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
prec = sys.argv[2] if len(sys.argv) > 2 else "tf32x3"
dist = sys.argv[3] if len(sys.argv) > 3 else "uniform"
model = build_model({"precision": prec}, {"embedding_dim": 128}).cuda().train()
opt = FusedAdam(model, lr=1e-3, weight_decay=1e-5)
pool = []
for i in range(3):
    b, y = synth.make_batch(seed=1 + i, batch=B, index_dtype=np.float64, id_dist=dist, edge_cases=False)
    pool.append(({k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in b.items() if k != "user_id"}, torch.from_numpy(y).cuda()))
loss_fn = torch.nn.BCELoss()
acc = collections.OrderedDict()
opt_ms = []
STEPS = 6
for it in range(3 + STEPS):
    if it == 3:
        lib.fbn_set_option(b"stage_events", 1)
    b, y = pool[it % 3]
    opt.zero_grad()
    out = model(b)
    loss = loss_fn(out, y)
    loss.backward()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    clip_grad_norm_(model, 10.0)
    opt.step()
    e1.record()
    if it >= 3:
        buf = C.create_string_buffer(8192)
        _lib.check(lib.fbn_stage_report(buf, 8192))
        for line in buf.value.decode().strip().split("\n"):
            name, ms = line.split("\t")
            acc[name] = acc.get(name, 0.0) + float(ms)
        opt_ms.append(e0.elapsed_time(e1))
tot = 0.0
print(f"B={B} precision={prec} ids={dist}: mean over {STEPS} eager steps")
for name, ms in acc.items():
    print(f"  {name:45s} {ms / STEPS * 1e3:9.1f} us")
    tot += ms / STEPS
print(f"  {'optimizer: clip + adam table + adam dense':45s} {sum(opt_ms) / STEPS * 1e3:9.1f} us")
print(f"  {'sum':45s} {(tot + sum(opt_ms) / STEPS) * 1e3:9.1f} us")
