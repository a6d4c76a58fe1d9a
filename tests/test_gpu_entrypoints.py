"""GPU: the reference's entry points (train_fibinet.py / Prediction.py surface) run end to end on a tiny
MicroLens-shaped parquet set, and the checkpoint they write is the reference's 28-key state_dict."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch
import yaml

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("table_sharding", [None, "row"])
def test_train_and_predict_scripts(tmp_path, table_sharding):
    from oracle import synth
    data = tmp_path / "data" / "MicroLens_1M_x1"
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "make_synth_dataset.py"), str(data), "--train", "3000", "--valid", "700",
                    "--test", "900"], check=True, capture_output=True)
    cfg = yaml.safe_load(open(os.path.join(ROOT, "config", "fibinet_config.yaml")))
    run = cfg[cfg["base_expid"]]
    run.update(epochs=2, batch_size=1024)
    if table_sharding:
        run["table_sharding"] = table_sharding      # one rank here: the same kernels, the exchange degenerates to a local copy
    (tmp_path / "config").mkdir()
    yaml.safe_dump(cfg, open(tmp_path / "config" / "fibinet_config.yaml", "w"))
    cwd = tmp_path / "src"
    cwd.mkdir()
    env = dict(os.environ, FBN_NUM_WORKERS="0", PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "src", "train_fibinet.py")], cwd=cwd, env=env, capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "Valid AUC" in r.stdout and "Best AUC" in r.stdout
    ckpt = tmp_path / "checkpoints" / "FiBiNET_best.pth"
    assert ckpt.exists()
    sd = torch.load(ckpt, map_location="cpu")
    shapes = synth.state_dict_shapes()
    assert list(sd) == list(shapes) and all(tuple(sd[k].shape) == tuple(shapes[k]) for k in sd)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "src", "Prediction.py")], cwd=cwd, env=env, capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    import pandas as pd
    sub = pd.read_csv(cwd / "prediction_fibinet.csv")
    assert list(sub.columns) == ["ID", "Task2"] and len(sub) == 900
    assert sub["Task2"].between(0, 1).all() and (cwd / "submission_fibinet.zip").exists()
    # the written checkpoint scores identically through the module API
    from ctr_recommendation_b200 import build_model
    model = build_model({"precision": "tf32x3"}, {"embedding_dim": 128})
    model.load_state_dict(sd)
    model = model.cuda().eval()
    sys.path.insert(0, os.path.join(ROOT, "src"))
    from dataloader import BatchCollator, ParquetDataset
    ds = ParquetDataset(str(data / "test.parquet"))
    coll = BatchCollator(None, 20, ds.column_index, str(data / "item_info.parquet"), strict=False)
    batch = coll([ds[i] for i in range(256)])
    with torch.no_grad():
        y = model({k: v.cuda() for k, v in batch.items()}).cpu().numpy()
    assert np.allclose(y, sub["Task2"].to_numpy()[:256], atol=1e-6)


def _run_train(cwd, env_extra):
    env = dict(os.environ, FBN_NUM_WORKERS="0", PYTHONPATH=ROOT, **env_extra)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "src", "train_fibinet.py")], cwd=cwd, env=env, capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return r.stdout


@pytest.mark.parametrize("table_sharding", [None, "row"])
def test_resume_continues_bit_for_bit(tmp_path, table_sharding):
    """FiBiNET_last.pth carries weights, Adam moments + step, scheduler, dropout-stream counters and the loader RNG;
    1 epoch + FBN_RESUME=1 for the remaining 2 ends in exactly the state of an uninterrupted 3-epoch run (dropout is ON).  With a
    row-sharded table every rank keeps its own resume file (one rank here: the table slice is the whole table)."""
    last = "FiBiNET_last.rank0of1.pth" if table_sharding else "FiBiNET_last.pth"
    states = {}
    for tag in ("straight", "resumed"):
        root = tmp_path / tag
        data = root / "data" / "MicroLens_1M_x1"
        subprocess.run([sys.executable, os.path.join(ROOT, "tools", "make_synth_dataset.py"), str(data), "--train", "1500", "--valid", "300",
                        "--test", "10"], check=True, capture_output=True)
        cfg = yaml.safe_load(open(os.path.join(ROOT, "config", "fibinet_config.yaml")))
        cfg[cfg["base_expid"]].update(epochs=3, batch_size=512)
        if table_sharding:
            cfg[cfg["base_expid"]]["table_sharding"] = table_sharding
        (root / "config").mkdir()
        yaml.safe_dump(cfg, open(root / "config" / "fibinet_config.yaml", "w"))
        cwd = root / "src"
        cwd.mkdir()
        if tag == "straight":
            _run_train(cwd, {})
        else:
            out = _run_train(cwd, {"FBN_STOP_AFTER_EPOCH": "1"})
            assert "stopping after epoch 1" in out
            out = _run_train(cwd, {"FBN_RESUME": "1"})
            assert "continuing at epoch 2" in out
        assert not os.path.exists(root / "checkpoints" / (last + ".tmp"))          # written beside, renamed over
        states[tag] = torch.load(root / "checkpoints" / last, map_location="cpu", weights_only=True)   # plain values + tensors only
    a, b = states["straight"], states["resumed"]
    # 1500 rows at batch 512 = 2 full steps + a 476-row tail per epoch: ONE dropout stream shared by both engines -> 9 steps
    assert a["epoch"] == b["epoch"] == 3 and a["dropout_counter"] == b["dropout_counter"] == 9
    assert a["optimizer"]["step"] == b["optimizer"]["step"] == 9
    for k in a["model"]:
        assert torch.equal(a["model"][k], b["model"][k]), k
    for k in ("m_flat", "v_flat", "m_item", "v_item"):
        assert torch.equal(a["optimizer"][k], b["optimizer"][k]), k


def test_device_resident_dataset_matches_loader(tmp_path):
    """FBN_DEVICE_DATASET=1 (training split resident on the GPU, batches cut on the device) trains to exactly the same state as
    the DataLoader path when both see the rows in file order."""
    states = {}
    for tag, extra in (("loader", {"FBN_SHUFFLE": "0"}), ("device", {"FBN_SHUFFLE": "0", "FBN_DEVICE_DATASET": "1"})):
        root = tmp_path / tag
        data = root / "data" / "MicroLens_1M_x1"
        subprocess.run([sys.executable, os.path.join(ROOT, "tools", "make_synth_dataset.py"), str(data), "--train", "1500", "--valid", "300",
                        "--test", "10"], check=True, capture_output=True)
        cfg = yaml.safe_load(open(os.path.join(ROOT, "config", "fibinet_config.yaml")))
        cfg[cfg["base_expid"]].update(epochs=2, batch_size=512)
        (root / "config").mkdir()
        yaml.safe_dump(cfg, open(root / "config" / "fibinet_config.yaml", "w"))
        cwd = root / "src"
        cwd.mkdir()
        out = _run_train(cwd, extra)
        assert ("resident on the GPU" in out) == (tag == "device")
        states[tag] = torch.load(root / "checkpoints" / "FiBiNET_last.pth", map_location="cpu", weights_only=True)
    a, b = states["loader"], states["device"]
    for k in a["model"]:
        assert torch.equal(a["model"][k], b["model"][k]), k
    assert torch.equal(a["optimizer"]["m_item"], b["optimizer"]["m_item"])


def test_reference_own_scripts_run_on_the_swapped_module(tmp_path):
    """SURVEY section 4, "Entry-point" row: the REFERENCE'S OWN train_fibinet.py / Prediction.py / dataloader.py / utils.py (byte copies
    vendored into oracle/_ref by oracle/make_ref.py) run unmodified with only src/model_fibinet.py swapped for this repo's shim --
    stock torch.optim.Adam, torch clip_grad_norm_, OneCycleLR, the pandas loader with 4 workers, the 28-key checkpoint."""
    ref = os.path.join(ROOT, "oracle", "_ref")
    if not os.path.exists(os.path.join(ref, "train_fibinet.py")):
        pytest.skip("oracle/_ref does not hold the reference scripts (run python oracle/make_ref.py where /root/reference exists)")
    import shutil
    from oracle import synth
    data = tmp_path / "data" / "MicroLens_1M_x1"
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "make_synth_dataset.py"), str(data), "--train", "2500", "--valid", "600",
                    "--test", "700"], check=True, capture_output=True)
    cfg = yaml.safe_load(open(os.path.join(ROOT, "config", "fibinet_config.yaml")))
    cfg[cfg["base_expid"]].update(epochs=2, batch_size=512)
    (tmp_path / "config").mkdir()
    yaml.safe_dump(cfg, open(tmp_path / "config" / "fibinet_config.yaml", "w"))
    cwd = tmp_path / "src"
    cwd.mkdir()
    for f in ("train_fibinet.py", "Prediction.py", "dataloader.py", "utils.py"):
        shutil.copyfile(os.path.join(ref, f), cwd / f)
    shutil.copyfile(os.path.join(ROOT, "src", "model_fibinet.py"), cwd / "model_fibinet.py")         # the only swapped file
    env = dict(os.environ, PYTHONPATH=ROOT, PYTHONIOENCODING="utf-8", CUDA_VISIBLE_DEVICES="0")
    r = subprocess.run([sys.executable, "train_fibinet.py"], cwd=cwd, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "Valid AUC" in r.stdout
    sd = torch.load(tmp_path / "checkpoints" / "FiBiNET_best.pth", map_location="cpu")
    shapes = synth.state_dict_shapes()
    assert list(sd) == list(shapes) and all(tuple(sd[k].shape) == tuple(shapes[k]) for k in sd)
    r = subprocess.run([sys.executable, "Prediction.py"], cwd=cwd, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    import pandas as pd
    sub = pd.read_csv(cwd / "prediction_fibinet.csv")
    assert list(sub.columns) == ["ID", "Task2"] and len(sub) == 700 and sub["Task2"].between(0, 1).all()
