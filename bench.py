#!/usr/bin/env python
"""bench.py -- FiBiNET train-step throughput on synthetic MicroLens_1M_x1-shaped batches.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (sm_100a)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU path (oracle/_ref, else the port), rank 0 only

One "step" = zero_grad -> forward -> BCELoss -> backward -> clip_grad_norm_(10) -> Adam(wd 1e-5) ->
OneCycleLR (reference src/train_fibinet.py:113-122) over one batch.  Prints ONE JSON line (rank 0).
  value : whole-job samples/s with the batch already resident in HBM (CUDA-event timed, max over ranks)
  e2e   : same through the public API from pinned HOST buffers, H2D copy and loss read-back inside the
          timed region
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "train samples/sec FiBiNET MicroLens-shape"
UNIT = "samples/s"
L_HIST = 20


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=int(os.environ.get("FBN_BENCH_BATCH", "65536")),
                    help="per-GPU batch (BASELINE config 2 sweeps 1K-64K; 65536 is its largest point)")
    ap.add_argument("--precision", default=os.environ.get("FBN_BENCH_PRECISION", "f16x3"), choices=["fp32", "tf32x3", "f16x3", "bf16"])
    ap.add_argument("--id-dist", default="uniform", choices=["uniform", "zipf"])
    ap.add_argument("--cpu-sample", type=int, default=0, help="rows per CPU-baseline step (0 = the per-GPU batch itself)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-config1", action="store_true", help="skip the second block at the reference's batch_size 4096")
    ap.add_argument("--pool", type=int, default=4, help="distinct synthetic batches cycled through")
    ap.add_argument("--bilinear", default="all", choices=["all", "each", "interaction"],
                    help="bilinear_type (BASELINE config 2 sweep); the reference hard-codes 'all'")
    ap.add_argument("--mode", default="train", choices=["train", "infer"],
                    help="train = the headline train step; infer = eval forward over the batch (Prediction.py loop body, BASELINE config 3)")
    ap.add_argument("--sharding", default="replicated", choices=["replicated", "row"],
                    help="row = item table partitioned by id %% N over the ranks (BASELINE config 5): remote gather over NVLink, "
                         "owner-side gradient merge (engine.ShardedTrainStep)")
    ap.add_argument("--item-rows", type=int, default=0, help="rows of the item table (default: the reference's 91718)")
    ap.add_argument("--lazy", action="store_true", help="row sharding: lazy row Adam (touched rows only) instead of dense-exact Adam")
    ap.add_argument("--resident-mm", action="store_true",
                    help="keep the frozen (91718,128) item_emb_d128 matrix on the GPU and gather it by item_id inside the fused kernel "
                         "(SURVEY 8f-1): batches then carry ids only (188 instead of 700 bytes / sample over PCIe)")
    ap.add_argument("--int32-ids", action="store_true", help="loader delivers int32 ids / history instead of float64 / int64 (100 B / sample)")
    ap.add_argument("--opt", action="append", default=[], metavar="NAME=INT", help="fbn_set_option knob, e.g. tc_persistent=-1 (A/B runs)")
    ap.add_argument("--no-overlap", action="store_true", help="data parallel: blocking gradient all-reduces between two graphs (A/B against "
                    "the default schedule that overlaps them with the weight-gradient GEMMs)")
    ap.add_argument("--overlap", default=None, choices=["partial", "full", "wgrad"], help="data-parallel schedule (default: wgrad from 4 ranks and 8192 rows per rank, else blocking)")
    ap.add_argument("--reserve-sms", type=int, default=16, help="data parallel, overlapped: SMs the weight-gradient GEMMs leave to the collectives")
    ap.add_argument("--phased", action="store_true", help="1 GPU: run the phased backward (chain / leaf 1 / leaf 2 graphs) without collectives")
    ap.add_argument("--fields", type=int, default=0, help="F > 0: benchmark the F-field model of ctr_recommendation_b200/general.py "
                    "(BASELINE config 5's 40 fields; one table per field, --field-vocab rows each) instead of the six-field model")
    ap.add_argument("--field-vocab", type=int, default=100000)
    ap.add_argument("--shard-rows-per-gpu", type=int, default=int(os.environ.get("FBN_BENCH_SHARD_ROWS", "12500000")),
                    help="N > 1: after the replicated-table measurement the same bench line gets a `sharded` block -- the item table with "
                         "this many rows PER GPU (12.5 M x 8 = BASELINE config 5's 100 M rows) row-sharded over the ranks; 0 disables it")
    ap.add_argument("--eager", action="store_true", help="per-kernel launches through autograd instead of the CUDA-graph TrainStep")
    return ap.parse_args()


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            p = json.load(fh)
        return dict(hbm=float(p["hbm_gbs"]), tf=float(p["bf16_tflops"]), tf_sus=float(p.get("bf16_tflops_sustained", p["bf16_tflops"])),
                    src="measured")
    return dict(hbm=6650.0, tf=1590.0, tf_sus=1400.0, src="fallback")


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self._stop, self._th = index, [], threading.Event(), None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.05)

    def __enter__(self):
        self._th = threading.Thread(target=self._run, daemon=True)
        self._th.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._th.join(timeout=6)

    def summary(self):
        sm = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 8:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
def make_pool(args, rank, n):
    """n synthetic batches as pinned host tensors in the reference loader's dtypes
    (float64 scalars, int64 item_seq, fp32 item_emb_d128, fp32 labels)."""
    from oracle import synth
    table = synth.make_item_mm_table(seed=11)
    pool = []
    big = args.item_rows > synth.V_ITEM
    g = torch.Generator().manual_seed(4242 + rank)
    for i in range(n):
        b, y = synth.make_batch(seed=2025 + 1000 * rank + i, batch=args.batch, max_len=L_HIST, id_dist=args.id_dist,
                                index_dtype=np.float64, mm_table=table, edge_cases=False)
        b.pop("user_id")
        if args.resident_mm:
            b.pop("item_emb_d128")
        if args.int32_ids:
            for k in ("item_id", "likes_level", "views_level", "item_seq"):
                b[k] = b[k].astype(np.int32)
        if big:    # scaled synthetic table: ids uniform over [1, V); the padding pattern of the history is kept
            V = args.item_rows
            b["item_id"] = torch.randint(1, V, (args.batch,), generator=g).numpy().astype(np.float64)
            seq = torch.randint(1, V, (args.batch, L_HIST), generator=g).numpy()
            b["item_seq"] = np.where(b["item_seq"] != 0, seq, 0).astype(np.int64)
        host = {k: torch.from_numpy(np.ascontiguousarray(v)).pin_memory() for k, v in b.items()}
        pool.append((host, torch.from_numpy(y).pin_memory()))
    return pool


def run_ours(args):
    from ctr_recommendation_b200 import build_model, FusedAdam, clip_grad_norm_, _lib
    from ctr_recommendation_b200 import dist as fdist
    import torch.distributed as dist
    world_env = int(os.environ.get("WORLD_SIZE", "1"))
    overlapped = (not args.no_overlap and args.sharding != "row" and args.mode == "train" and not args.eager and args.batch >= 8192
                  and (args.overlap is not None or world_env >= 4))
    rank, local, world = fdist.init_from_env(nccl_max_ctas=args.reserve_sms if overlapped else None)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    lib = _lib.load()
    _lib.check(lib.fbn_check_device(local), "fbn_check_device")
    for kv in args.opt:
        name, val = kv.split("=")
        _lib.check(lib.fbn_set_option(name.encode(), int(val)), f"fbn_set_option({kv})")
    torch.manual_seed(2025)
    sharded = args.sharding == "row"
    fm = {"precision": args.precision, "bilinear_type": args.bilinear}
    if args.item_rows:
        fm["item_rows"] = args.item_rows
    if sharded:
        fm["table_sharding"] = "row"
        if args.eager or args.mode == "infer":
            raise SystemExit("--sharding row is benchmarked through ShardedTrainStep (train mode, no --eager)")
    model = build_model(fm, {"embedding_dim": 128}).to(dev).train()
    if world > 1:
        fdist.broadcast_parameters(model)
    if args.resident_mm:
        from oracle import synth as _synth
        model.attach_mm_table(torch.from_numpy(_synth.make_item_mm_table(seed=11)))
    opt = FusedAdam(model, lr=1e-3, weight_decay=1e-5)
    total_steps = max(10, 40 * (args.steps + args.warmup) * 2 + 512)
    sched = torch.optim.lr_scheduler.OneCycleLR(opt, max_lr=1e-2, total_steps=total_steps, pct_start=0.3, div_factor=25.0,
                                                final_div_factor=1000.0)
    loss_fn = torch.nn.BCELoss()
    pool = make_pool(args, rank, args.pool)
    dev_pool = [({k: v.to(dev) for k, v in b.items()}, y.to(dev)) for b, y in pool]
    stage = ({k: torch.empty_like(v, device=dev) for k, v in pool[0][0].items()}, torch.empty_like(pool[0][1], device=dev))
    h2d_bytes = sum(v.numel() * v.element_size() for v in pool[0][0].values()) + pool[0][1].numel() * 4

    from ctr_recommendation_b200.engine import Scorer, ShardedTrainStep, TrainStep
    infer = args.mode == "infer"
    if infer:
        model.eval()
    if sharded:
        engine = ShardedTrainStep(model, opt, args.batch, L_HIST, idx_dtype=torch.float64, max_norm=10.0, lazy=args.lazy,
                                  merge_cap=4 * args.batch * (1 + L_HIST))
    else:
        idt = torch.int32 if args.int32_ids else torch.float64
        sdt = torch.int32 if args.int32_ids else torch.int64
        engine = None if args.eager else (
            Scorer(model, args.batch, L_HIST, idx_dtype=idt, seq_dtype=sdt, use_mm_table=args.resident_mm) if infer else
            TrainStep(model, opt, args.batch, L_HIST, idx_dtype=idt, seq_dtype=sdt, max_norm=10.0, use_mm_table=args.resident_mm,
                      overlap=False if args.no_overlap else args.overlap, reserve_sms=args.reserve_sms, phased_single=args.phased))

    def step(batch, labels):
        if infer:                    # scoring: forward only, predictions read back by the caller
            if engine is not None:
                return engine(batch)
            with torch.no_grad():
                return model(batch)
        if engine is not None:       # CUDA-graph replay of the same loop body
            loss = engine(batch, labels)
            sched.step()
            return loss
        opt.zero_grad()
        y = model(batch)
        loss = loss_fn(y, labels)
        loss.backward()
        if world > 1:
            fdist.sync_gradients(model)
        clip_grad_norm_(model, 10.0)
        opt.step()
        sched.step()
        return loss

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(steps):
            fn(k)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    def resident(k):
        b, y = dev_pool[k % len(dev_pool)]
        step(b, y)

    def e2e(k):
        hb, hy = pool[k % len(pool)]
        if engine is not None and not infer:
            # every step's inputs cross PCIe inside the timed region; the copy of batch k+1 is issued (copy stream) right
            # after step k is launched, as a prefetching loader would
            if not engine._prefetched:
                engine.prefetch(hb, hy)
            loss = engine()
            sched.step()
            nb, ny = pool[(k + 1) % len(pool)]
            engine.prefetch(nb, ny)
            return loss.item()
        if engine is not None:      # Scorer: the copy of batch k+1 (copy stream) overlaps the scoring of batch k, as a prefetching
            if not engine._prefetched:   # loader would; the predictions are read back every batch like Prediction.py:113
                engine.prefetch(hb)
            loss = engine()
            engine.prefetch(pool[(k + 1) % len(pool)][0])
        else:
            for name, t in hb.items():
                stage[0][name].copy_(t, non_blocking=True)
            stage[1].copy_(hy, non_blocking=True)
            loss = step(stage[0], stage[1])
        if infer:
            return loss.cpu()   # predictions to the host every batch, like Prediction.py:113
        return loss.item()      # D2H read of the step's result, like the reference loop (:124)

    for k in range(args.warmup):
        resident(k)
    launches0 = lib.fbn_launch_count()
    with ClockSampler(local) as clk:
        ms = timed(resident, args.steps)
    launches = lib.fbn_launch_count() - launches0
    for k in range(max(1, args.warmup // 2)):
        e2e(k)
    with ClockSampler(local) as clk2:
        ms_e2e = timed(e2e, args.steps)

    global_batch = args.batch * world
    value = global_batch * args.steps / (ms / 1e3)
    e2e_value = global_batch * args.steps / (ms_e2e / 1e3)
    peaks = load_peaks()
    kernels = kernel_rooflines(args, model, dev_pool[0], peaks, lib) if (rank == 0 and not infer and not sharded) else {}
    if sharded:
        st_ = model._shard.stats()
        if st_["overflow"]:
            raise SystemExit("row sharding: merge capacity overflow")
    out = {
        "metric": METRIC if not infer else "inference samples/sec FiBiNET MicroLens-shape (Prediction.py path)", "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": DTYPE_LABEL[args.precision], "data": "synthetic",
        "config": workload_config(args, world, infer, sharded, model._shard.item_rows if sharded else None),
        "precision": args.precision, "launch": "eager" if args.eager else "cuda-graph",
        "dp_collectives": (None if world == 1 or sharded else
                           {False: "blocking all-reduces between the backward and update graphs",
                            "partial": "table-gradient all-reduce overlapped with the leaf gradients (all but MLP-1's), then the dense all-reduce",
                            "wgrad": "table + small-bucket all-reduces overlapped with the MLP-1 weight gradient, then the MLP-1 bucket",
                            "full": "3 async all-reduces (table gradient, MLP-1 bucket, rest) overlapped with the weight-gradient GEMMs"}[
                               getattr(engine, "overlap", False)]),
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e / args.steps, "h2d_bytes_per_step": h2d_bytes,
                "d2h_bytes_per_step": 4 if not infer else 4 * args.batch},
        "gpu_launches": int(getattr(engine, "kernels_per_step", 0) * args.steps) if engine is not None else int(launches),
        "clocks": clk.summary(),
        "clocks_e2e": clk2.summary(),
    }
    if world == 1 and not infer and not sharded and engine is not None and args.batch != 4096 and not args.no_config1:
        # second block: the same step at the reference's own batch_size (config/fibinet_config.yaml: 4096 = BASELINE config 1)
        a1 = argparse.Namespace(**vars(args))
        a1.batch = 4096
        pool1 = make_pool(a1, rank, args.pool)
        dev1 = [({k: v.to(dev) for k, v in b.items()}, y.to(dev)) for b, y in pool1]
        eng1 = TrainStep(model, opt, 4096, L_HIST, idx_dtype=idt, seq_dtype=sdt, max_norm=10.0, use_mm_table=args.resident_mm)
        n1 = max(4 * args.steps, 40)

        def res1(k):
            eng1(*dev1[k % len(dev1)])
            sched.step()

        def e2e1(k):
            if not eng1._prefetched:
                eng1.prefetch(*pool1[k % len(pool1)])
            loss = eng1()
            sched.step()
            eng1.prefetch(*pool1[(k + 1) % len(pool1)])
            return loss.item()
        for k in range(max(args.warmup, 3)):
            res1(k)
        ms1 = timed(res1, n1)
        for k in range(3):
            e2e1(k)
        ms1e = timed(e2e1, n1)
        out["config1"] = {"config": workload_config(a1, world), "value": 4096 * n1 / (ms1 / 1e3), "unit": UNIT, "steps": n1,
                          "ms_per_step": ms1 / n1, "gpu_launches_per_step": int(eng1.kernels_per_step),
                          "e2e": {"value": 4096 * n1 / (ms1e / 1e3), "unit": UNIT, "ms_per_step": ms1e / n1,
                                  "h2d_bytes_per_step": sum(v.numel() * v.element_size() for v in pool1[0][0].values()) + 4096 * 4,
                                  "d2h_bytes_per_step": 4}}
    if world > 1 and not infer and not sharded and engine is not None and args.shard_rows_per_gpu > 0:
        out["sharded"] = sharded_block(args, rank, world, dev, out, timed)
    if world == 1 and not infer and not sharded and engine is not None and not args.no_config1 and not args.resident_mm:
        out.update(extra_blocks(args, model, opt, sched, dev, timed, idt, sdt))
    if rank == 0:
        out.update(kernels)
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline(args, steps=4, warmup=1)
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def sharded_block(args, rank, world, dev, main_line, timed):
    """BASELINE config 5's table on the same ranks: the item table (--shard-rows-per-gpu x N rows) row-sharded by id % N, remote
    gather over NVLink inside the forward kernel, owner-side gradient merge, lazy row Adam (engine.ShardedTrainStep).  Reported as a
    block of the main line so that every multi-GPU run of the bench carries it.  A watchdog prints the main line without the block
    and ends every rank if this secondary measurement ever stalls -- it must not be able to take the headline number with it."""
    from ctr_recommendation_b200 import build_model, FusedAdam
    from ctr_recommendation_b200 import dist as fdist
    from ctr_recommendation_b200.engine import ShardedTrainStep

    def bail():
        if rank == 0:
            line = dict(main_line)
            line["sharded"] = {"error": "the row-sharded block did not finish within 420 s"}
            print(json.dumps(line), flush=True)
        os._exit(0)
    dog = threading.Timer(420.0, bail)
    dog.daemon = True
    dog.start()
    try:
        rows = int(args.shard_rows_per_gpu) * world
        a2 = argparse.Namespace(**vars(args))
        a2.item_rows, a2.sharding, a2.lazy = rows, "row", True
        t0 = time.perf_counter()
        model = build_model({"precision": args.precision, "table_sharding": "row", "item_rows": rows}, {"embedding_dim": 128}).to(dev).train()
        fdist.broadcast_parameters(model)
        opt = FusedAdam(model, lr=1e-3, weight_decay=1e-5)
        eng = ShardedTrainStep(model, opt, args.batch, L_HIST, idx_dtype=torch.float64, max_norm=10.0, lazy=True,
                               merge_cap=4 * args.batch * (1 + L_HIST))
        pool = make_pool(a2, rank, 2)
        dev_pool = [({k: v.to(dev) for k, v in b.items()}, y.to(dev)) for b, y in pool]
        setup_s = time.perf_counter() - t0
        steps = max(3, min(args.steps, 10))

        def res(k):
            eng(*dev_pool[k % len(dev_pool)])
        for k in range(3):
            res(k)
        ms = timed(res, steps)
        st = model._shard.stats()
        blk = {"config": workload_config(a2, world, False, True, rows), "value": args.batch * world * steps / (ms / 1e3), "unit": UNIT,
               "steps": steps, "ms_per_step": ms / steps, "item_rows": rows, "rows_per_gpu": int(model._shard.shard_rows),
               "table_and_moments_gb_per_gpu": round(model._shard.shard_rows * 128 * 4 * 3 / 1e9, 2), "optimizer": "lazy row Adam",
               "merge_overflow": bool(st["overflow"]), "setup_s": round(setup_s, 1)}
        return blk       # (the slices stay allocated: peers hold CUDA IPC mappings of them until the process ends)
    except Exception as e:      # never lose the headline line to the secondary measurement
        return {"error": f"{type(e).__name__}: {e}"[:300]}
    finally:
        dog.cancel()


def kernel_rooflines(args, model, dev_batch, peaks, lib):
    """Stand-alone CUDA-event timings of the kernels the step is made of, against their rooflines.
    Algorithmic bytes / FLOPs per unit are the figures of SURVEY 8(d) (restated in DESIGN.md)."""
    import ctypes as C
    from ctr_recommendation_b200 import _lib
    B = args.batch
    st = _lib.stream_ptr()
    flush = torch.empty(160 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")

    def time_kernel(fn, iters=10):
        fn()
        torch.cuda.synchronize()
        tot = 0.0
        for _ in range(iters):
            flush.fill_(1.0)               # evict L2 between timed launches
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        return tot / iters

    out = {}
    # (1) dense-exact table Adam: 6 streams (p,m,v in/out) + gradient rows read
    opt = model._fused_optimizer
    rows = model.item_emb.weight.shape[0]
    h = _lib.AdamHyper(1e-3, 0.9, 0.999, 1e-8, 1e-5, 10)
    w = model.item_emb.weight.data

    def adam():
        _lib.check(lib.fbn_adam_table(_lib.ptr(w), _lib.ptr(opt._m_item), _lib.ptr(opt._v_item), _lib.ptr(model._item_grad),
                                      _lib.ptr(model._row_touched), rows, None, C.byref(h), None, st))
    ms = time_kernel(adam)
    touched = int((model._row_touched > 0).sum().item())
    adam_bytes = 6 * rows * 512 + touched * 512 + rows * 4
    out["adam_table"] = {"ms": ms, "bytes": adam_bytes, "GBps": adam_bytes / ms / 1e6, "frac": adam_bytes / ms / 1e6 / peaks["hbm"]}
    # (2) fused gather + pooling + projection + SENET
    model.eval()
    bs, keep, Bq, L = model._batch_struct(dev_batch[0])
    ws = model._workspace(Bq, L)
    P = model._params_struct()

    def gather():
        _lib.check(lib.fbn_embed_forward(C.byref(P), C.byref(bs), _lib.ptr(ws), ws.numel(), 1, st))
    ms = time_kernel(gather)
    nvalid = int((dev_batch[0]["item_seq"] != 0).sum().item())
    # idx 184 B + 2 cate rows + item row + mm vector + history rows + V out (2560) + saved X5/xhat (3072 + 32)
    g_bytes = B * (184 + 2 * 512 + 512 + 512 + 2560 + 2560 + 512 + 48) + nvalid * 512
    out["gather_senet_fwd"] = {"ms": ms, "bytes": g_bytes, "GBps": g_bytes / ms / 1e6, "frac": g_bytes / ms / 1e6 / peaks["hbm"]}
    model.train()
    # (3) the three MLP-1 GEMMs (forward, data gradient, weight gradient) exactly as the step launches them -- same operands,
    #     strides, structural-zero masks and split-K -- on the workspace the training steps above left behind; L2 flushed before
    #     every timed launch, CUDA events on the launching stream (fbn_time_stage)
    live_cols = 15 * 128                      # 15 live 128-column blocks of the 21 (user field and its pairs are zero)
    msf = C.c_float(0.0)

    def tstage(name, stage, useful_flops):
        _lib.check(lib.fbn_time_stage(C.byref(P), C.byref(bs), _lib.ptr(ws), ws.numel(), stage.encode(), _lib.ptr(flush), flush.numel() * 4,
                                      8, C.byref(msf), st), "fbn_time_stage")
        ms = float(msf.value)
        out[name] = {"ms": ms, "flops": useful_flops, "TFLOPs": useful_flops / ms / 1e9,
                     "frac_of_bf16_peak": useful_flops / ms / 1e9 / peaks["tf"]}
    gemm_flops = 2.0 * B * live_cols * 512
    tstage("mlp1_fwd_gemm", "mlp1", gemm_flops)
    tstage("mlp1_dgrad_gemm", "mlp1_dgrad", gemm_flops)
    tstage("mlp1_wgrad_gemm", "mlp1_wgrad", gemm_flops)       # includes the fixed-order split-K reduction
    out["mlp1_gemm"] = out["mlp1_fwd_gemm"]
    if args.precision == "f16x3":      # the stage that writes the MLP input as fp16 hi|lo under one scale (amax pass + split pass)
        _lib.check(lib.fbn_time_stage(C.byref(P), C.byref(bs), _lib.ptr(ws), ws.numel(), b"bil_pairs", _lib.ptr(flush), flush.numel() * 4,
                                      8, C.byref(msf), st), "fbn_time_stage")
        pk_bytes = B * (2 * 9 * 512 + 15 * 512)             # 5 fields + 4 transforms read twice, 15 blocks of hi|lo pairs written
        out["pairs_split_mlp_input"] = {"ms": float(msf.value), "bytes": pk_bytes, "GBps": pk_bytes / float(msf.value) / 1e6,
                                        "frac": pk_bytes / float(msf.value) / 1e6 / peaks["hbm"]}
    # the kernel with the largest share of the step (ncu launch list, profiles/): the MLP-1 data-gradient GEMM
    dom = max(("mlp1_dgrad_gemm", "mlp1_fwd_gemm", "mlp1_wgrad_gemm"), key=lambda k: out[k]["ms"])
    g = out[dom]
    traffic = None
    for name in ("r2_traffic.json", "r1_traffic.json"):
        tpath = os.path.join(ROOT, "profiles", name)
        if traffic is None and os.path.exists(tpath):
            with open(tpath) as fh:
                traffic = json.load(fh).get(f"{dom}/{args.precision}/b{B}")
    shapes = {"mlp1_fwd_gemm": "forward, H1[B,512] = C[B,1920 live of 2688] x W1^T", "mlp1_dgrad_gemm": "data gradient, dC[B,1920 live] = dH1[B,512] x W1",
              "mlp1_wgrad_gemm": "weight gradient, dW1[512,1920 live] = dH1^T x C, split-K + fixed-order reduce"}
    # gather: SURVEY 8(d) counts 5,304 + 512 * n_valid bytes per sample; the wider count adds what the kernel also writes for
    # backward (X5 / xhat / ids / gates: 3,120 B per sample)
    g8d = B * 5304 + nvalid * 512
    gk = out["gather_senet_fwd"]
    gk["bytes_8d"] = g8d
    gk["GBps_8d"] = g8d / gk["ms"] / 1e6
    gk["frac_8d"] = gk["GBps_8d"] / peaks["hbm"]
    roof = {"bound": "tensor", "kernel": f"gemm_tc2p_kernel (MLP-1 {shapes[dom]})", "achieved": g["TFLOPs"], "peak": peaks["tf"],
            "unit": "TFLOP/s", "frac": g["frac_of_bf16_peak"], "traffic": traffic, "peak_source": peaks["src"] + " bf16 dense (burst)",
            "note": "dominant kernel of the step; algorithmic FLOPs 2*B*1920*512 (structural-zero blocks not counted); " + PREC_NOTE[args.precision],
            "all_mlp1_gemms": {k: {"ms": out[k]["ms"], "TFLOPs": out[k]["TFLOPs"], "frac": out[k]["frac_of_bf16_peak"]}
                               for k in ("mlp1_fwd_gemm", "mlp1_dgrad_gemm", "mlp1_wgrad_gemm")},
            "hbm_kernels": {"adam_table": {"achieved_GBps": out["adam_table"]["GBps"], "frac": out["adam_table"]["frac"]},
                            "gather_senet_fwd": {"achieved_GBps_8d_bytes": gk["GBps_8d"], "frac_8d_bytes": gk["frac_8d"],
                                                 "achieved_GBps_incl_saved": gk["GBps"], "frac_incl_saved": gk["frac"]}}}
    return {"roofline": roof, "kernels": out}


DTYPE_LABEL = {"fp32": "f32", "tf32x3": "tf32x3(f32-grade)", "f16x3": "f16x3(f32-grade)", "bf16": "bf16"}
PREC_NOTE = {
    "tf32x3": "precision tf32x3 issues 3x these on the tensor pipe and kind::tf32 runs at half the bf16 rate used as denominator, so "
              "1/6 = 0.167 is this scheme's ceiling",
    "f16x3": "precision f16x3 (fp16 hi|lo operands under one power-of-two scale per tensor, fp32-grade like tf32x3: "
             "tools/split_precision_sim.py) issues 3x these as kind::f16 MMAs at the bf16 rate used as denominator, so 1/3 = 0.333 is "
             "this scheme's ceiling",
    "bf16": "one bf16 pass", "fp32": "SIMT fp32 FMA (no tensor cores)"}


# ------------------------------------------------------------------------------------------------
def workload_config(args, world, infer=False, sharded=False, item_rows=None):
    """The `config` object of the JSON line: names the workload only, so the reference arm carries the identical object."""
    bil = "bilinear all" if args.bilinear == "all" else f"bilinear {args.bilinear} [not the reference's hard-coded 'all']"
    where = {65536: "largest point of BASELINE config 2's 1K-64K batch sweep",
             4096: "the reference's batch_size in config/fibinet_config.yaml, BASELINE config 1"}.get(
                 args.batch, "a point of BASELINE config 2's 1K-64K batch sweep")
    workload = (f"FiBiNET {'train step' if not infer else 'eval forward'} (config/fibinet_config.yaml model: D=128, 6 fields, {bil}, "
                f"MLP 2688-512-256-1), per-GPU batch {args.batch} ({where}), history L={L_HIST}, "
                f"item ids {args.id_dist}, " +
                (f"item table of {item_rows} rows row-sharded over {world} ranks (remote gather over NVLink, owner-side "
                 f"gradient merge, {'lazy row' if args.lazy else 'dense-exact'} Adam)" if sharded else "replicated tables"))
    return {"workload": workload, "global_batch": args.batch * world, "per_gpu_batch": args.batch,
            "parallelism": f"dp{world}" + ("+row-sharded item table" if sharded else ""),
            "l2": "working set per step (table p/m/v/grad 188 MB + activations) exceeds the 126 MB L2; inputs cycle over "
                  f"{args.pool} distinct batches"}


class ReferenceStep:
    """The reference's own train step on the host CPU.  kind = "reference": the UNMODIFIED src/model_fibinet.py vendored into
    oracle/_ref/ by oracle/make_ref.py (run by build() in the dev container), driven by the loop body of
    src/train_fibinet.py:113-122 (zero_grad, forward, BCELoss, backward, clip_grad_norm_(10), Adam.step, OneCycleLR.step,
    loss.item()).  kind = "port" (oracle/_ref absent): oracle/fibinet_torch_port.py, the same ATen op stream."""

    def __init__(self, lr=1e-3, weight_decay=1e-5, total_steps=1000):
        import importlib.util
        from oracle import make_ref, synth
        W = synth.make_weights(seed=7)
        rd = make_ref.ref_dir()
        if rd is not None:
            spec = importlib.util.spec_from_file_location("_reference_model_fibinet", os.path.join(rd, "model_fibinet.py"))
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            self.kind = "reference"
            self.model = mod.build_model(None, {"embedding_dim": 128})
            self.model.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in W.items()}, strict=True)
            self.model.train()
            self.opt = torch.optim.Adam(self.model.parameters(), lr=lr, weight_decay=weight_decay)          # ref :78
            self.loss_fn = torch.nn.BCELoss()                                                              # ref :79
            self.sched = torch.optim.lr_scheduler.OneCycleLR(self.opt, max_lr=lr * 10, total_steps=total_steps, pct_start=0.3,
                                                             div_factor=25.0, final_div_factor=1000.0)      # ref :84-92
        else:
            from oracle import fibinet_torch_port as port
            self.kind = "port"
            self.tr = port.Trainer(port.tensors_from_numpy(W), lr=lr, weight_decay=weight_decay)

    def step(self, batch, labels):
        if self.kind == "port":
            return self.tr.step(batch, labels)
        self.opt.zero_grad()                                                            # ref :113
        y = self.model(dict(batch))
        loss = self.loss_fn(y, labels)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(self.model.parameters(), max_norm=10.0)          # ref :119
        self.opt.step()
        self.sched.step()
        return loss.item()                                                              # ref :124


def cpu_baseline(args, steps, warmup):
    """The reference's CPU train step on all host cores, on a bounded number of steps of the SAME per-GPU batch."""
    from oracle import synth
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    B = args.batch if args.cpu_sample <= 0 else min(args.cpu_sample, args.batch)
    ref = ReferenceStep(total_steps=steps + warmup + 8)
    table = synth.make_item_mm_table(seed=11)
    batches = []
    for i in range(2):
        b, y = synth.make_batch(seed=2025 + i, batch=B, id_dist=args.id_dist, index_dtype=np.float64, mm_table=table, edge_cases=False)
        b.pop("user_id")
        batches.append(({k: torch.from_numpy(v) for k, v in b.items()}, torch.from_numpy(y)))
    for k in range(warmup):
        ref.step(*batches[k % 2])
    t0 = time.perf_counter()
    for k in range(steps):
        ref.step(*batches[k % 2])
    dt = time.perf_counter() - t0
    what = ("the unmodified reference src/model_fibinet.py (oracle/_ref) + the loop body of src/train_fibinet.py:113-122"
            if ref.kind == "reference" else "oracle/fibinet_torch_port.py (oracle/_ref absent)")
    return {"value": B * steps / dt, "unit": UNIT, "cores": torch.get_num_threads(), "kind": ref.kind,
            "sample": f"{steps} train steps of {B} rows (the per-GPU batch is {args.batch}), {what}, torch {torch.__version__} CPU fp32, "
                      f"{cores} host cores", "ms_per_step": dt / steps * 1e3}


def run_reference(args):
    """`--impl reference`: the reference's CPU path on the same config / metric, rank 0 only (the other ranks exit 0)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cb = cpu_baseline(args, steps=args.steps, warmup=args.warmup)
    out = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": cb["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f32", "data": "synthetic",
           "config": workload_config(args, args.gpus),
           "cpu_baseline": cb,
           "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


def general_measure(precision, bilinear, F, V, B, steps, warmup, pool_n=2):
    """BASELINE config 5's field count on one GPU: GeneralFiBiNET (F lookups -> SENET -> bilinear -> (F + F(F-1)/2) * 128 wide MLP)
    forward + BCELoss + backward + clip_grad_norm_(10) + torch.optim.Adam(fused) per step, ids resident on the device."""
    from ctr_recommendation_b200 import GeneralFiBiNET, _lib
    lib = _lib.load()
    torch.manual_seed(2025)
    model = GeneralFiBiNET([(f"f{i}", V) for i in range(F)], precision=precision, bilinear_type=bilinear).cuda().train()
    model.check_ids_every_forward = False
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-5, fused=True)
    loss_fn = torch.nn.BCELoss()
    g = torch.Generator().manual_seed(7)
    pool = [(torch.randint(0, V, (B, F), generator=g).cuda(), (torch.rand(B, generator=g) < 0.5).float().cuda()) for _ in range(pool_n)]

    def step(k):
        ids, y = pool[k % len(pool)]
        opt.zero_grad(set_to_none=True)
        loss = loss_fn(model({"ids": ids}), y)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), 10.0)
        opt.step()
        return loss
    for k in range(warmup):
        step(k)
    torch.cuda.synchronize()
    n0 = lib.fbn_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(steps):
        step(k)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    model.check_ids()
    P = F * (F - 1) // 2
    flops = 3 * 2.0 * B * ((F + P) * 128 * 512 + 512 * 256 + 256)          # forward + two backward GEMMs per layer
    return {"value": B * steps / (ms / 1e3), "unit": UNIT, "steps": steps, "ms_per_step": ms / steps,
            "config": {"workload": f"F-field FiBiNET train step (ctr_recommendation_b200/general.py), F={F} fields x {V}-row tables (one "
                                   f"(F*V,128) parameter), {P} bilinear pairs ({bilinear}), MLP {(F + P) * 128}-512-256-1, batch {B}, uniform "
                                   "ids; unfused building blocks + torch.optim.Adam(fused)", "global_batch": B, "per_gpu_batch": B,
                       "parallelism": "dp1"},
            "gpu_launches": int(lib.fbn_launch_count() - n0), "mlp_tflops_useful": flops * steps / (ms / 1e3) / 1e12}


def run_general(args):
    torch.cuda.set_device(0)
    r = general_measure(args.precision, args.bilinear, args.fields, args.field_vocab, args.batch, args.steps, args.warmup, args.pool)
    out = {"metric": "train samples/sec F-field FiBiNET (general.py)", "value": r["value"], "unit": UNIT, "n_gpus": 1,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": DTYPE_LABEL[args.precision], "data": "synthetic",
           "config": r["config"], "gpu_launches": r["gpu_launches"], "mlp_tflops_useful": r["mlp_tflops_useful"]}
    print(json.dumps(out), flush=True)


def extra_blocks(args, model, opt, sched, dev, timed, idt, sdt):
    """More of BASELINE's configurations inside the single-GPU line, so that every run of the bench carries them:
    `sweep` (config 2: batch 1K-64K, bilinear all / each / interaction, bf16), `infer` (config 3: the Prediction.py loop body) and
    `fields40` (config 5's field count on the F-field model).  Resident inputs, CUDA-event timed, a few seconds each."""
    from ctr_recommendation_b200 import build_model, FusedAdam
    from ctr_recommendation_b200.engine import Scorer, TrainStep
    blocks = {}
    n = max(args.steps, 10)

    def train_rate(m, o, sc, B, precision_note=None):
        a = argparse.Namespace(**vars(args))
        a.batch = B
        pool = make_pool(a, 0, 2)
        dp = [({k: v.to(dev) for k, v in b.items()}, y.to(dev)) for b, y in pool]
        eng = TrainStep(m, o, B, L_HIST, idx_dtype=idt, seq_dtype=sdt, max_norm=10.0)

        def res(k):
            eng(*dp[k % 2])
            sc.step()
        steps = n if B >= 16384 else 4 * n
        for k in range(3):
            res(k)
        ms = timed(res, steps)
        return {"per_gpu_batch": B, "value": B * steps / (ms / 1e3), "ms_per_step": ms / steps, "launches_per_step": int(eng.kernels_per_step)}
    try:
        sweep = []
        for B in (1024, 4096, 16384):
            if B != args.batch:
                r = train_rate(model, opt, sched, B)
                r.update(bilinear=args.bilinear, precision=args.precision)
                sweep.append(r)
        for bil, prec in (("each", args.precision), ("interaction", args.precision), ("all", "bf16")):
            m2 = build_model({"precision": prec, "bilinear_type": bil}, {"embedding_dim": 128}).to(dev).train()
            o2 = FusedAdam(m2, lr=1e-3, weight_decay=1e-5)
            s2 = torch.optim.lr_scheduler.OneCycleLR(o2, max_lr=1e-2, total_steps=100000)
            r = train_rate(m2, o2, s2, 16384)
            r.update(bilinear=bil, precision=prec)
            sweep.append(r)
            del m2, o2
        blocks["sweep"] = {"unit": UNIT, "what": "BASELINE config 2: train step, resident inputs; the headline line is the 65536 / all point", "points": sweep}
    except Exception as e:
        blocks["sweep"] = {"error": f"{type(e).__name__}: {e}"[:300]}
    try:
        model.eval()
        inf = []
        for B in (8192, 65536):
            a = argparse.Namespace(**vars(args))
            a.batch = B
            pool = make_pool(a, 0, 2)
            hp = [b for b, _ in pool]
            dp = [{k: v.to(dev) for k, v in b.items()} for b in hp]
            sc = Scorer(model, B, L_HIST, idx_dtype=idt, seq_dtype=sdt)
            steps = 4 * n

            def res(k):
                sc(dp[k % 2])

            def e2e(k):
                if not sc._prefetched:
                    sc.prefetch(hp[k % 2])
                p = sc()
                sc.prefetch(hp[(k + 1) % 2])
                return p.cpu()
            for k in range(3):
                res(k)
            ms = timed(res, steps)
            for k in range(2):
                e2e(k)
            ms2 = timed(e2e, steps)
            inf.append({"batch": B, "value": B * steps / (ms / 1e3), "e2e_value": B * steps / (ms2 / 1e3), "ms_per_batch": ms / steps,
                        "h2d_bytes_per_batch": sum(v.numel() * v.element_size() for v in hp[0].values()), "d2h_bytes_per_batch": 4 * B})
        blocks["infer"] = {"unit": UNIT, "what": "BASELINE config 3: eval forward (Prediction.py loop body; 8192 is the script's batch), one "
                                                  "CUDA-graph replay per batch; e2e = pinned host batches in the loader's format + predictions read back",
                           "points": inf}
    except Exception as e:
        blocks["infer"] = {"error": f"{type(e).__name__}: {e}"[:300]}
    finally:
        model.train()
    try:
        blocks["fields40"] = general_measure(args.precision, "all", 40, 100000, 4096, 5, 2)
    except Exception as e:
        blocks["fields40"] = {"error": f"{type(e).__name__}: {e}"[:300]}
    return blocks


if __name__ == "__main__":
    a = parse()
    if a.fields > 0:
        run_general(a)
    elif a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
