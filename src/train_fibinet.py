"""Training entry point with the reference's surface (src/train_fibinet.py): run it from src/ (or the repo
root), it reads config/fibinet_config.yaml, trains MM_FiBiNET with Adam(lr, weight_decay) + OneCycleLR + grad-norm
clip 10, validates every epoch with AUC and saves the best state_dict to ../checkpoints/FiBiNET_best.pth.

What differs from the reference is only *how* the loop body runs: the model, the backward pass and the optimizer
step are hand-written sm_100a CUDA (ctr_recommendation_b200), the step is replayed from a CUDA graph, and more
than one GPU means one process per GPU under torchrun (NCCL all-reduce) instead of nn.DataParallel.  There is no
CPU path: without a CUDA device the script stops with an error.
"""
import os
import sys

import numpy as np
import torch
import yaml

sys.path.append(os.path.dirname(os.path.abspath(__file__)))

from dataloader import MMCTRDataLoader  # noqa: E402
from model_fibinet import build_model  # noqa: E402
from utils import compute_auc, set_seed  # noqa: E402

from ctr_recommendation_b200 import FusedAdagrad, FusedAdam  # noqa: E402
from ctr_recommendation_b200 import dist as fdist  # noqa: E402
from ctr_recommendation_b200 import sharded  # noqa: E402
from ctr_recommendation_b200.engine import Scorer, ShardedTrainStep, TrainStep  # noqa: E402


class DeviceBatches:
    """The training split resident on the GPU, in the loader's own dtypes (float64 scalars, int64 history cropped to max_len,
    float32 labels: 188 bytes / row, 0.7 GB for MicroLens_1M_x1's 3.6 M rows).  A batch is a slice / index_select on the device,
    so an epoch involves no host work at all -- the reference-style loader delivers 0.3-2 M samples/s, the step consumes 7-15 M.
    Opt-in (FBN_DEVICE_DATASET=1); same batches as the loader when shuffling is off."""

    def __init__(self, loader, device, shuffle):
        ds, coll = loader.dataset, loader.collator
        mat, ci = ds.darray, ds.column_index
        coll.check_known(mat[:, ci["item_id"]])                       # the loader's strict check, once for the whole split

        def dev(a, dtype=None):
            a = np.ascontiguousarray(a if dtype is None else a.astype(dtype))
            return torch.from_numpy(a).to(device)
        self.cols = {k: dev(mat[:, ci[k]]) for k in ("item_id", "likes_level", "views_level")}
        if "item_seq" in ci:
            idx = ci["item_seq"]
            lo, hi = idx[0], idx[0] + len(idx)
            self.cols["item_seq"] = dev(mat[:, max(lo, hi - coll.max_len):hi], np.int64)
        self.labels = dev(mat[:, ci["label"]], np.float32)
        self.n, self.batch, self.shuffle, self.device = mat.shape[0], loader.batch_size, shuffle, device

    def __len__(self):
        return -(-self.n // self.batch)

    def __iter__(self):
        order = torch.randperm(self.n).to(self.device) if self.shuffle else None      # CPU generator: covered by the resume state
        for i in range(0, self.n, self.batch):
            if order is None:
                sl = slice(i, min(self.n, i + self.batch))
                yield {k: v[sl] for k, v in self.cols.items()}, self.labels[sl]
            else:
                idx = order[i:i + self.batch]
                yield {k: v.index_select(0, idx) for k, v in self.cols.items()}, self.labels.index_select(0, idx)


def load_config():
    path = "../config/fibinet_config.yaml"
    if not os.path.exists(path):
        path = "config/fibinet_config.yaml"
    print(f"[config] {path}")
    with open(path, "r") as fh:
        cfg = yaml.safe_load(fh)
    return cfg["dataset_config"][cfg["dataset_id"]], cfg[cfg["base_expid"]]


def main():
    dataset_cfg, model_cfg = load_config()
    set_seed(model_cfg.get("seed", 2025))
    if not torch.cuda.is_available():
        raise SystemExit("train_fibinet.py: no CUDA device -- this implementation has no CPU path (sm_100a kernels only)")
    # `dp_overlap: false` in the run config: blocking gradient all-reduces (default: from 4 ranks and 8192 rows per rank the table
    # all-reduce runs beside the MLP-1 weight gradient; the collectives then get 16 SMs and the GEMMs leave those free)
    dp_overlap = bool(model_cfg.get("dp_overlap", True))
    world_env = int(os.environ.get("WORLD_SIZE", "1"))
    overlapping = dp_overlap and world_env >= 4 and int(model_cfg.get("batch_size", 4096)) // max(world_env, 1) >= 8192
    rank, local, world = fdist.init_from_env(nccl_max_ctas=16 if overlapping else None)
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    log = print if rank == 0 else (lambda *a, **k: None)
    log(f"[setup] FiBiNET on {torch.cuda.get_device_name(local)} x{world}")

    batch_size = int(model_cfg.get("batch_size", 4096))
    max_len = int(model_cfg.get("max_len", 20))
    # single-process loader by default: with the batched fetch (ParquetDataset.__getitems__) it outruns 4 workers + IPC.
    # with_mm=False: batches carry ids only, the frozen item_emb_d128 matrix lives on the GPU
    workers = int(os.environ.get("FBN_NUM_WORKERS", "0"))
    shuffle = os.environ.get("FBN_SHUFFLE", "1") != "0"
    train_loader = MMCTRDataLoader(None, dataset_cfg["train_data"], dataset_cfg["item_info"], batch_size=batch_size, shuffle=shuffle,
                                   num_workers=workers, max_len=max_len, with_mm=False, pin_memory=True)
    valid_loader = MMCTRDataLoader(None, dataset_cfg["valid_data"], dataset_cfg["item_info"], batch_size=batch_size, shuffle=False,
                                   num_workers=workers, max_len=max_len, with_mm=False, pin_memory=True)

    fm = {"precision": model_cfg.get("precision", "f16x3")}
    row_sharded = str(model_cfg.get("table_sharding", "")).lower() == "row"   # B200 extra: item table partitioned by id % world
    if row_sharded:
        fm["table_sharding"] = "row"
    model = build_model(fm, model_cfg).to(device)
    model.attach_mm_table(torch.from_numpy(train_loader.collator.item_embedding_matrix))
    if world > 1:
        fdist.broadcast_parameters(model)
    lr = float(model_cfg.get("learning_rate", 1e-3))
    weight_decay = float(model_cfg.get("weight_decay", 1e-5))
    epochs = int(model_cfg.get("epochs", 30))
    # torch.optim.Adam semantics (L2-coupled) like the reference, which ignores `optimizer: adamw`; with honor_config: true
    # the key is honoured (decoupled weight decay)
    adamw = bool(model_cfg.get("honor_config", False)) and str(model_cfg.get("optimizer", "adam")).lower() == "adamw"
    if bool(model_cfg.get("honor_config", False)) and str(model_cfg.get("optimizer", "adam")).lower() == "adagrad":
        optimizer = FusedAdagrad(model, lr=lr, weight_decay=weight_decay)     # extension: Adagrad row update (north_star (2))
    else:
        optimizer = FusedAdam(model, lr=lr, weight_decay=weight_decay, decoupled_weight_decay=adamw)
    steps_per_epoch = len(train_loader)
    scheduler = torch.optim.lr_scheduler.OneCycleLR(optimizer, max_lr=lr * 10, epochs=epochs, steps_per_epoch=steps_per_epoch,
                                                    pct_start=0.3, div_factor=25.0, final_div_factor=1000.0,
                                                    cycle_momentum="betas" in optimizer.defaults)    # Adagrad has no beta1 to cycle
    engines = {}

    def train_engine(rows, L, dtype, n_global):
        # n_global is part of the key: this rank's loss weight is rows / n_global (DataParallel: mean over the global batch)
        key = ("t", rows, L, dtype, n_global)
        if key not in engines:
            cls = ShardedTrainStep if row_sharded else TrainStep
            # ONE collective schedule per run (a rank left without rows replays it through another engine: every engine must issue
            # the same sequence of all-reduces)
            kw = {} if row_sharded else {"overlap": "wgrad" if overlapping else False}
            engines[key] = cls(model, optimizer, rows, L, idx_dtype=dtype, max_norm=10.0, use_mm_table=True, global_batch=n_global, **kw)
        return engines[key]

    def score_engine(rows, L, dtype):
        key = ("s", rows, L, dtype)
        if key not in engines:
            engines[key] = Scorer(model, rows, L, idx_dtype=dtype, use_mm_table=True)
        return engines[key]

    best_auc = 0.0
    os.makedirs("../checkpoints", exist_ok=True)
    best_path = "../checkpoints/FiBiNET_best.pth"
    # Resume (the reference only ever saves the best weights and cannot continue a run): after every epoch rank 0
    # writes the full training state -- weights, Adam moments + step, scheduler, dropout-stream counters, loader RNG -- and
    # FBN_RESUME=1 continues from it bit for bit.  With a row-sharded table every rank writes its own file (its slice of the table
    # and of the Adam moments live only there; the dense state is replicated) and a run resumes on the same number of ranks.
    last_path = f"../checkpoints/FiBiNET_last.rank{rank}of{world}.pth" if row_sharded else "../checkpoints/FiBiNET_last.pth"
    start_epoch = 0
    have_last = os.path.exists(last_path)
    if row_sharded and world > 1:      # all ranks or none
        flag = torch.tensor([1 if have_last else 0], device=device)
        torch.distributed.all_reduce(flag, op=torch.distributed.ReduceOp.MIN)
        have_last = bool(flag.item())
    if os.environ.get("FBN_RESUME") == "1" and have_last:
        state = torch.load(last_path, map_location="cpu", weights_only=True)     # tensors + plain Python values only
        model.load_state_dict(state["model"])
        optimizer.load_state_dict(state["optimizer"])
        scheduler.load_state_dict(state["scheduler"])
        start_epoch, best_auc = int(state["epoch"]), float(state["best_auc"])
        model.set_dropout_counter(int(state["dropout_counter"]))                 # continue the dropout stream
        torch.set_rng_state(state["torch_rng"])
        log(f"[resume] {last_path}: continuing at epoch {start_epoch + 1}")
    stop_after = int(os.environ.get("FBN_STOP_AFTER_EPOCH", "0"))
    train_batches = train_loader
    if os.environ.get("FBN_DEVICE_DATASET") == "1":
        train_batches = DeviceBatches(train_loader, device, shuffle)
        log(f"[data] training split resident on the GPU: {train_batches.n} rows")
    log("[train] start")
    for epoch in range(start_epoch, epochs):
        model.train()
        total_loss = torch.zeros(1, device=device)
        steps = 0
        for batch_dict, labels in train_batches:
            shard, ylab, _ = fdist.shard_batch(batch_dict, labels, rank, world)     # DataParallel-style split on dim 0
            rows, n_global = ylab.shape[0], labels.shape[0]
            seq = shard.get("item_seq")
            if rows == 0:      # scatter chunking left this rank without rows (tiny tail batch): zero gradients, same collectives
                loss = next(e for k, e in engines.items() if k[0] == "t").step_empty()
            else:
                step = train_engine(rows, 0 if seq is None else seq.shape[1], shard["item_id"].dtype, n_global)
                loss = step(shard, ylab)                   # fwd + BCE + bwd + clip(10) + Adam, one graph replay
            scheduler.step()
            total_loss += loss                             # stays on the device: no per-step host sync
            steps += 1
            if steps % 200 == 0:
                log(f"Epoch {epoch + 1} | Step {steps} | Loss: {loss.item():.4f} | LR: {scheduler.get_last_lr()[0]:.6f}")
        avg_loss = (total_loss.item() / steps) if steps else 0.0
        for k, e in engines.items():
            e.check_ids()                                  # IndexError for ids outside the tables, as nn.Embedding would raise
            if row_sharded and k[0] == "t":
                e.check()                                  # RuntimeError if the owner-side merge lists overflowed (rows would be dropped)

        model.eval()
        y_trues, y_preds = [], []
        for batch_dict, labels in valid_loader:
            shard, ylab, _ = fdist.shard_batch(batch_dict, labels, rank, world)
            seq = shard.get("item_seq")
            if ylab.shape[0] == 0:
                local = torch.empty(0, dtype=torch.float32, device=device)
            else:
                local = score_engine(ylab.shape[0], 0 if seq is None else seq.shape[1], shard["item_id"].dtype)(shard).clone()
            pred = fdist.gather_predictions(local)
            y_trues.append(labels.numpy())
            y_preds.append(pred.cpu().numpy())
        if y_trues:
            auc = compute_auc(np.concatenate(y_trues), np.concatenate(y_preds))
            log(f"Epoch {epoch + 1} | Train Loss: {avg_loss:.4f} | Valid AUC: {auc:.4f}")
            if auc > best_auc:
                best_auc = auc
                sd = model.state_dict()
                if row_sharded:      # the checkpoint keeps the reference's format: the full (91718,128) table (collective)
                    sd["item_emb.weight"] = sharded.gather_full_table(model)
                if rank == 0:
                    torch.save(sd, best_path)
                    log(f"[ckpt] new best -> {best_path}")
        if rank == 0 or row_sharded:
            torch.cuda.synchronize()
            # tensors and plain values only (loads with weights_only=True); written beside the old file and renamed over it, so a
            # crash during the save never leaves a truncated resume file
            sched_state = {k: v for k, v in scheduler.state_dict().items() if isinstance(v, (int, float, bool, str, list, tuple, dict))
                           and k != "_scale_fn_ref"}
            tmp = last_path + ".tmp"
            torch.save({"model": model.state_dict(),
                        "optimizer": {k: (v.cpu() if torch.is_tensor(v) else v) for k, v in optimizer.state_dict().items()},
                        "scheduler": sched_state, "epoch": epoch + 1, "best_auc": float(best_auc),
                        "dropout_counter": int(model._dropout_counter(device).item()),
                        "torch_rng": torch.get_rng_state()}, tmp)
            os.replace(tmp, last_path)
        if stop_after and epoch + 1 >= stop_after:
            log(f"[train] stopping after epoch {epoch + 1} (FBN_STOP_AFTER_EPOCH)")
            return
    log(f"Done. Best AUC: {best_auc:.4f}")


if __name__ == "__main__":
    main()
