"""Inference entry point with the reference's surface (src/Prediction.py): rebuild the model, load
../checkpoints/FiBiNET_best.pth (stripping a DataParallel "module." prefix), score test.parquet in batches of 8192
and write prediction_fibinet.csv (ID, Task2) + submission_fibinet.zip.  The eval forward is one CUDA-graph replay
of the sm_100a kernels per batch; there is no CPU path.  Under torchrun (one process per GPU) every rank scores a contiguous
slice of the test split with its own replica -- inference shards with no data-path collective -- and rank 0 gathers
the predictions in rank order = row order and writes the files."""
import os
import sys
import zipfile

import numpy as np
import pandas as pd
import torch
import yaml
from torch.utils.data import DataLoader

sys.path.append(os.path.dirname(os.path.abspath(__file__)))

from dataloader import BatchCollator, ParquetDataset  # noqa: E402
from model_fibinet import build_model  # noqa: E402

from ctr_recommendation_b200 import dist as fdist  # noqa: E402
from ctr_recommendation_b200.engine import Scorer  # noqa: E402


class InferenceCollator(BatchCollator):
    """Label-less collation; unknown item ids get an all-zero item_emb_d128 instead of raising."""

    def __init__(self, max_len, column_index, item_info_path):
        # with_mm=False: item_emb_d128 stays on the GPU (resident matrix gathered by item_id inside the fused kernel)
        super().__init__(None, max_len, column_index, item_info_path, strict=False, with_mm=False)


def main():
    path = "../config/fibinet_config.yaml"
    if not os.path.exists(path):
        path = "config/fibinet_config.yaml"
    with open(path, "r") as fh:
        cfg = yaml.safe_load(fh)
    dataset_cfg = cfg["dataset_config"][cfg["dataset_id"]]
    model_cfg = cfg[cfg["base_expid"]]
    if not torch.cuda.is_available():
        raise SystemExit("Prediction.py: no CUDA device -- this implementation has no CPU path (sm_100a kernels only)")
    rank, local, world = fdist.init_from_env()
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)

    model = build_model({"precision": model_cfg.get("precision", "f16x3")}, model_cfg)
    ckpt = "../checkpoints/FiBiNET_best.pth"
    if not os.path.exists(ckpt):
        ckpt = "checkpoints/FiBiNET_best.pth"
    if rank == 0:
        print(f"[ckpt] {ckpt}")
    state = torch.load(ckpt, map_location="cpu")
    model.load_state_dict({k.replace("module.", ""): v for k, v in state.items()})
    model.to(device)
    model.eval()

    test_dataset = ParquetDataset(dataset_cfg["test_data"])
    collator = InferenceCollator(int(model_cfg.get("max_len", 20)), test_dataset.column_index, dataset_cfg["item_info"])
    model.attach_mm_table(torch.from_numpy(collator.item_embedding_matrix))
    # the batched fetch (ParquetDataset.__getitems__) makes the single-process loader faster than 4 workers + IPC
    rows = test_dataset
    if world > 1:      # this rank's contiguous slice (torch scatter chunking, like every other split in this repo)
        lo, hi = fdist.shard_bounds(len(test_dataset), rank, world)
        rows = torch.utils.data.Subset(test_dataset, range(lo, hi))
    loader = DataLoader(rows, batch_size=8192, shuffle=False, num_workers=int(os.environ.get("FBN_NUM_WORKERS", "0")),
                        collate_fn=collator, pin_memory=True)
    scorers, preds = {}, []
    for batch in loader:
        rows = batch["item_id"].shape[0]
        seq = batch.get("item_seq")
        key = (rows, 0 if seq is None else seq.shape[1], batch["item_id"].dtype)
        if key not in scorers:
            scorers[key] = Scorer(model, key[0], key[1], idx_dtype=key[2], use_mm_table=True)
        preds.append(scorers[key](batch).clone())      # stays on the device: no per-batch host sync (reference: .cpu() per batch)
    local_pred = torch.cat(preds) if preds else torch.empty(0, dtype=torch.float32, device=device)
    predictions = fdist.gather_predictions(local_pred).cpu().numpy()          # rank order == row order
    for sc in scorers.values():
        sc.check_ids()                                  # IndexError for ids outside the tables, as nn.Embedding would raise
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()
    if rank != 0:
        return

    sub = pd.DataFrame({"ID": range(len(predictions)), "Task2": predictions})
    sub.to_csv("prediction_fibinet.csv", index=False)
    with zipfile.ZipFile("submission_fibinet.zip", "w", zipfile.ZIP_DEFLATED) as zf:
        zf.write("prediction_fibinet.csv")
    print("wrote submission_fibinet.zip")


if __name__ == "__main__":
    main()
