"""Write a small MicroLens_1M_x1-shaped parquet set (train/valid/test/item_info) for the entry-point tests.

    python tools/make_synth_dataset.py OUT_DIR [--train N] [--valid N] [--test N] [--items V]
Columns follow SURVEY 8(d): user_id, item_seq list<int64>[100] left-padded, likes_level, views_level, item_id, label;
item_info: item_id, item_tags list<int64>[5], item_emb_d128 list<float32>[128].
"""
import argparse
import os
import sys

import numpy as np
import pandas as pd

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("out")
    ap.add_argument("--train", type=int, default=4096)
    ap.add_argument("--valid", type=int, default=1024)
    ap.add_argument("--test", type=int, default=1500)
    ap.add_argument("--items", type=int, default=synth.V_ITEM)
    a = ap.parse_args()
    os.makedirs(a.out, exist_ok=True)
    table = synth.make_item_mm_table(seed=11, rows=a.items)
    info = pd.DataFrame({"item_id": np.arange(a.items, dtype=np.int64),
                         "item_tags": list(synth.randint(3, 1, a.items * 5, 0, 100).reshape(a.items, 5)),
                         "item_emb_d128": list(table)})
    info.to_parquet(os.path.join(a.out, "item_info.parquet"))
    for name, n, seed in (("train", a.train, 1), ("valid", a.valid, 2), ("test", a.test, 3)):
        b, y = synth.make_batch(seed=seed, batch=n, max_len=100, index_dtype=np.int64, mm_table=table, edge_cases=False, vocab=a.items)
        cols = {"user_id": b["user_id"], "item_seq": list(b["item_seq"]), "likes_level": b["likes_level"],
                "views_level": b["views_level"], "item_id": b["item_id"]}
        if name != "test":
            cols["label"] = y.astype(np.float64)
        pd.DataFrame(cols).to_parquet(os.path.join(a.out, f"{name}.parquet"))
    print("wrote", a.out)


if __name__ == "__main__":
    main()
