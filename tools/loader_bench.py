"""CPU-only: how fast does the entry points' input pipeline (src/dataloader.py: parquet -> float64 matrix -> per-batch collate
with the item_emb_d128 lookup) deliver batches?  BASELINE config 3 asks for the inference number with and without the loader in
the loop; the GPU side scores 46-49 M samples/s, so this is the number that bounds Prediction.py / train_fibinet.py end to end.

    python tools/loader_bench.py [rows] [batch] [workers ...]"""
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "src"))


def main():
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
    batch = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
    workers = [int(w) for w in sys.argv[3:]] or [0, 4]
    with tempfile.TemporaryDirectory() as tmp:
        subprocess.run([sys.executable, os.path.join(ROOT, "tools", "make_synth_dataset.py"), tmp, "--train", "64", "--valid", "64", "--test",
                        str(rows)], check=True, capture_output=True)
        for w in workers:
            from torch.utils.data import DataLoader
            from dataloader import BatchCollator, ParquetDataset
            ds = ParquetDataset(os.path.join(tmp, "test.parquet"))
            coll = BatchCollator(None, 20, ds.column_index, os.path.join(tmp, "item_info.parquet"), strict=False)
            dl = DataLoader(ds, batch_size=batch, shuffle=False, num_workers=w, collate_fn=coll)
            n, t0 = 0, time.perf_counter()
            for b in dl:
                n += b["item_id"].shape[0]
            dt = time.perf_counter() - t0
            print(f"loader: {rows} rows, batch {batch}, num_workers={w}: {n / dt / 1e3:.1f} K samples/s ({dt:.2f} s), "
                  f"{os.cpu_count()} host cores", flush=True)


if __name__ == "__main__":
    main()
