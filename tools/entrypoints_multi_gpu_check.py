"""Entry points under torchrun (run on a box with >= 2 GPUs:  python tools/entrypoints_multi_gpu_check.py [N]).

1. `torchrun --nproc-per-node N src/train_fibinet.py` on a tiny MicroLens-shaped parquet set whose last batch is uneven (and, with
   N = 8, leaves ranks without rows): trains, validates, writes the reference's 28-key checkpoint.
2. `python src/Prediction.py` (one GPU) and `torchrun --nproc-per-node N src/Prediction.py` (N replicas, one contiguous slice of the
   test split each, predictions gathered in rank order) must write the same prediction_fibinet.csv.
"""
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np
import pandas as pd
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    tmp = tempfile.mkdtemp(prefix="fbn_entry_")
    data = os.path.join(tmp, "data", "MicroLens_1M_x1")
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "make_synth_dataset.py"), data, "--train", "4099", "--valid", "701", "--test",
                    "1903"], check=True, capture_output=True)
    cfg = yaml.safe_load(open(os.path.join(ROOT, "config", "fibinet_config.yaml")))
    cfg[cfg["base_expid"]].update(epochs=2, batch_size=1024)          # 4099 rows: 4 full batches + a 3-row tail (empty ranks for N > 3)
    os.makedirs(os.path.join(tmp, "config"))
    yaml.safe_dump(cfg, open(os.path.join(tmp, "config", "fibinet_config.yaml"), "w"))
    cwd = os.path.join(tmp, "src")
    os.makedirs(cwd)
    env = dict(os.environ, FBN_NUM_WORKERS="0", PYTHONPATH=ROOT)
    tr = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1", "--master-port",
          "29541"]
    r = subprocess.run(tr + [os.path.join(ROOT, "src", "train_fibinet.py")], cwd=cwd, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "Valid AUC" in r.stdout and "Best AUC" in r.stdout, r.stdout[-2000:]
    auc_lines = [l for l in r.stdout.splitlines() if "Valid AUC" in l]
    r = subprocess.run([sys.executable, os.path.join(ROOT, "src", "Prediction.py")], cwd=cwd, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    one = pd.read_csv(os.path.join(cwd, "prediction_fibinet.csv"))
    os.remove(os.path.join(cwd, "prediction_fibinet.csv"))
    r = subprocess.run(tr + [os.path.join(ROOT, "src", "Prediction.py")], cwd=cwd, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    many = pd.read_csv(os.path.join(cwd, "prediction_fibinet.csv"))
    assert len(one) == len(many) == 1903 and list(many.columns) == ["ID", "Task2"]
    d = float(np.abs(one["Task2"].to_numpy() - many["Task2"].to_numpy()).max())
    assert d <= 1e-6, d          # different batch shapes per rank -> different GEMM tiles; same values to fp32 rounding
    print(f"entrypoints OK on {n} GPUs: train_fibinet.py under torchrun (uneven + tiny tail batches) -> {auc_lines[-1].strip()}; "
          f"Prediction.py on 1 vs {n} ranks: {len(many)} predictions, max |diff| {d:.2e}")
    shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
