"""CPU: host-side logic -- C-ABI surface, module/state_dict contract, optimizer plumbing, loader parity."""
import os
import re
import sys

import numpy as np
import pytest
import torch

from oracle import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_c_abi_exports_every_declared_symbol():
    from ctr_recommendation_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "fibinet_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(fbn_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 20
    lib = _lib.load()                      # builds with nvcc if the .so is missing; no compute calls here
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/fibinet_b200.h but not exported"
    assert declared == set(_lib.EXPORTED_SYMBOLS), declared ^ set(_lib.EXPORTED_SYMBOLS)
    assert b"sm_100a" in lib.fbn_version()
    assert lib.fbn_workspace_bytes(4096, 20, 91718) > 4096 * 2688 * 4
    assert lib.fbn_workspace_offset(64, 20, 91718, b"C") % 256 == 0
    assert lib.fbn_workspace_offset(64, 20, 91718, b"nope") == 2 ** 64 - 1


def test_state_dict_contract_and_no_cpu_path():
    from ctr_recommendation_b200 import build_model
    model = build_model(None, {"embedding_dim": 128})
    shapes = synth.state_dict_shapes()
    sd = model.state_dict()
    assert list(sd.keys()) == list(shapes.keys())                    # same 28 keys, same order
    for k, v in sd.items():
        assert tuple(v.shape) == tuple(shapes[k]), k
    assert sd["mlp.1.num_batches_tracked"].dtype == torch.int64
    assert torch.all(sd["item_emb.weight"][0] == 0)                  # padding_idx=0
    W = synth.make_weights(seed=7)
    model.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in W.items()}, strict=True)
    model._ensure_flat()                                             # dense params re-homed into one buffer
    for k, v in model.state_dict().items():
        assert np.array_equal(v.numpy(), W[k]), k
    p = model.mlp[0].weight
    assert model._flat.data_ptr() <= p.data_ptr() < model._flat.data_ptr() + model._flat.numel() * 4
    model.load_state_dict({k: torch.from_numpy(np.asarray(np.array(v) * (2 if v.dtype == np.float32 else 1))) for k, v in W.items()})
    assert model._flat.data_ptr() <= model.mlp[0].weight.data_ptr() < model._flat.data_ptr() + model._flat.numel() * 4
    batch, _ = synth.make_batch(seed=1, batch=8)
    with pytest.raises(RuntimeError, match="no CPU path"):
        model({k: torch.from_numpy(v) for k, v in batch.items()})
    with pytest.raises(ValueError):
        build_model(None, {"embedding_dim": 64})
    with pytest.raises(ValueError):
        build_model(None, {})                                        # reference default embedding_dim=64
    from ctr_recommendation_b200 import BilinearInteraction
    with pytest.raises(ValueError, match="bilinear_type must be 'all' or 'each'"):
        BilinearInteraction(128, 6, "bogus")
    assert [tuple(w.shape) for w in BilinearInteraction(128, 6, "each").W_list] == [(128, 128)] * 5


def test_same_seed_same_init_as_reference_order():
    """Parameters are created in the reference's order with torch's initialisers, so a seed gives the same weights
    as the reference (checked against it when /root/reference is present)."""
    from ctr_recommendation_b200 import build_model
    torch.manual_seed(2025)
    a = build_model(None, {"embedding_dim": 128}).state_dict()
    if not os.path.isdir("/root/reference/src"):
        pytest.skip("reference not present on this box")
    sys.path.insert(0, "/root/reference/src")
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_model_fibinet", "/root/reference/src/model_fibinet.py")
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    sys.path.pop(0)
    torch.manual_seed(2025)
    b = ref.build_model(None, {"embedding_dim": 128}).state_dict()
    assert list(a) == list(b)
    for k in a:
        assert torch.equal(a[k], b[k]), k


def test_fused_adam_exposes_param_groups_for_onecycle():
    from ctr_recommendation_b200 import FusedAdam, build_model
    model = build_model(None, {"embedding_dim": 128})
    opt = FusedAdam(model, lr=1e-3, weight_decay=1e-5)
    sched = torch.optim.lr_scheduler.OneCycleLR(opt, max_lr=1e-2, epochs=2, steps_per_epoch=10, pct_start=0.3, div_factor=25.0,
                                                final_div_factor=1000.0)
    g = opt.param_groups[0]
    assert abs(g["lr"] - 4e-4) < 1e-12 and abs(g["betas"][0] - 0.95) < 1e-12
    assert all(id(p) != id(model.user_emb.weight) for p in g["params"])            # user_emb never gets a gradient
    with pytest.raises(TypeError):
        FusedAdam(torch.nn.Linear(2, 2))
    assert sched.get_last_lr()[0] == g["lr"]


def test_loader_matches_reference_loader(tmp_path):
    if not os.path.isdir("/root/reference/src"):
        pytest.skip("reference not present on this box")
    import subprocess
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "make_synth_dataset.py"), str(tmp_path), "--train", "300", "--valid", "10",
                    "--test", "50", "--items", "2000"], check=True, capture_output=True)
    import importlib.util

    def load(name, path):
        spec = importlib.util.spec_from_file_location(name, path)
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        return m
    ours = load("our_dl", os.path.join(ROOT, "src", "dataloader.py"))
    ref = load("ref_dl", "/root/reference/src/dataloader.py")
    kw = dict(batch_size=64, shuffle=False, num_workers=0, max_len=20)
    a = ours.MMCTRDataLoader(None, str(tmp_path / "train.parquet"), str(tmp_path / "item_info.parquet"), **kw)
    b = ref.MMCTRDataLoader(None, str(tmp_path / "train.parquet"), str(tmp_path / "item_info.parquet"), **kw)
    assert a.dataset.darray.dtype == b.dataset.darray.dtype == np.float64 and a.column_index == b.column_index
    n = 0
    for (ba, ya), (bb, yb) in zip(a, b):
        assert list(ba) == list(bb)
        for k in bb:
            assert ba[k].dtype == bb[k].dtype and torch.equal(ba[k], bb[k]), k
        assert torch.equal(ya, yb)
        n += 1
    assert n == 5
    coll = ours.BatchCollator(None, 20, a.column_index, str(tmp_path / "item_info.parquet"))
    with pytest.raises(KeyError):
        coll.lookup(np.array([1, 999999]))


def test_row_partition_roundtrip():
    """Row-sharded table partition (sharded.py): owner = g % N, local = g // N, slices interleave back to the table."""
    import torch
    from ctr_recommendation_b200 import sharded
    from oracle import shard_numpy as sorc
    for V in (2, 7, 64, 91718):
        for N in (1, 2, 3, 8, 16):
            R = sharded.shard_rows(V, N)
            assert R == sorc.shard_rows(V, N) and R * N >= V > (R - 1) * N
            g = np.arange(V)
            o, l = sharded.owner_of(g, N), sharded.local_row(g, N)
            assert np.array_equal(sharded.global_row(o, l, N), g) and l.max() < R
            oo, ll = sorc.owner_local(g, N)
            assert np.array_equal(o, oo) and np.array_equal(l, ll)
            full = torch.arange(V * 2, dtype=torch.float32).reshape(V, 2)
            slices = [sharded.slice_of_full(full, r, N) for r in range(N)]
            assert all(s.shape == (R, 2) for s in slices)
            assert np.array_equal(slices[N - 1].numpy(), sorc.slice_of(full.numpy(), N - 1, N))
            assert torch.equal(sharded.full_from_slices(slices, V), full)


def test_sharded_model_has_no_autograd_backward_and_needs_its_engine():
    import torch
    from ctr_recommendation_b200 import build_model, FusedAdam
    from ctr_recommendation_b200.engine import TrainStep, ShardedTrainStep
    m = build_model({"table_sharding": "row", "shard_rank": 2, "shard_world": 4, "item_rows": 1000}, {"embedding_dim": 128})
    assert m.item_emb.weight.shape == (250, 128) and m._shard.shard_rows == 250
    with pytest.raises(ValueError):
        build_model({"table_sharding": "column"}, {"embedding_dim": 128})
    plain = build_model(None, {"embedding_dim": 128})
    with pytest.raises(TypeError):
        ShardedTrainStep(plain, FusedAdam(plain), 8)


def test_lazy_adam_oracle_matches_dense_adam_on_touched_rows():
    """oracle/shard_numpy.lazy_adam_rows == oracle Adam (pinned against torch.optim.Adam by the golden vectors) on the rows
    it touches, and is the identity elsewhere."""
    from oracle import fibinet_numpy as orc
    from oracle import shard_numpy as sorc
    rng = np.random.default_rng(3)
    p0 = rng.standard_normal((40, 8)).astype(np.float32)
    g = rng.standard_normal((40, 8)).astype(np.float32)
    touched = rng.random(40) < 0.4
    P = {"w": p0.copy()}
    opt = orc.Adam(lr=3e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-5)
    opt.step(P, {"w": g})
    z = np.zeros_like(p0)
    p, m, v = sorc.lazy_adam_rows(p0, z, z, g, touched, 3e-3, 0.9, 0.999, 1e-8, 1e-5, step=1)
    assert np.array_equal(p[touched], P["w"][touched]) and np.array_equal(m[touched], opt.state["w"]["m"][touched])
    assert np.array_equal(p[~touched], p0[~touched]) and np.all(m[~touched] == 0) and np.all(v[~touched] == 0)


def test_loader_batched_fetch_and_resident_mm_mode(tmp_path):
    """ParquetDataset.__getitems__ (one fancy-index copy, or a view for sequential batches) == stacking the rows one by one;
    with_mm=False drops the per-batch item_emb_d128 block (the GPU gathers it from the resident matrix) but keeps the strict check."""
    import subprocess
    import importlib.util
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "make_synth_dataset.py"), str(tmp_path), "--train", "200", "--valid", "10",
                    "--test", "50", "--items", "2000"], check=True, capture_output=True)
    spec = importlib.util.spec_from_file_location("our_dl2", os.path.join(ROOT, "src", "dataloader.py"))
    dl = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(dl)
    ds = dl.ParquetDataset(str(tmp_path / "train.parquet"))
    for idx in ([3, 4, 5, 6], [7], [9, 2, 150, 2, 31]):
        assert np.array_equal(ds.__getitems__(idx), np.stack([ds[i] for i in idx]))
    full = dl.BatchCollator(None, 20, ds.column_index, str(tmp_path / "item_info.parquet"))
    lean = dl.BatchCollator(None, 20, ds.column_index, str(tmp_path / "item_info.parquet"), with_mm=False)
    rows = [ds[i] for i in (5, 1, 77)]
    (bf, yf), (bl, yl) = full(rows), lean(ds.__getitems__([5, 1, 77]))
    assert "item_emb_d128" in bf and "item_emb_d128" not in bl and torch.equal(yf, yl)
    for k in bl:
        assert torch.equal(bf[k], bl[k]), k
    assert np.array_equal(full.item_embedding_matrix[bf["item_id"].long().numpy()], bf["item_emb_d128"].numpy())
    bad = ds.__getitems__([0, 1]).copy()
    bad[1, ds.column_index["item_id"]] = 999999
    with pytest.raises(KeyError):
        lean(bad)
    torch.manual_seed(0)
    a = dl.MMCTRDataLoader(None, str(tmp_path / "train.parquet"), str(tmp_path / "item_info.parquet"), batch_size=64, shuffle=True,
                           num_workers=0, max_len=20, with_mm=False)
    seen = torch.cat([b["user_id"] for b, _ in a])
    assert seen.numel() == 200 and sorted(seen.tolist()) == sorted(ds.darray[:, ds.column_index["user_id"]].tolist())


def test_general_model_field_specs_and_dispatch():
    """Host logic of the F-field model (no GPU needed to construct it): field descriptors for plain lookups, shared tables, bags and
    padding rows; build_model dispatch on feature_map["fields"]; errors for malformed specs."""
    from ctr_recommendation_b200 import GeneralFiBiNET, MM_FiBiNET, build_model
    fm = {"fields": [{"name": "user_id", "vocab": 50}, ("likes_level", 11), {"name": "views_level", "table": "likes_level"},
                     {"name": "item_id", "vocab": 100, "padding_idx": 0}, {"name": "item_seq", "table": "item_id", "bag": 20},
                     {"name": "item_tags", "vocab": 30, "bag": 5}], "senet_reduction": 3, "bilinear_type": "each"}
    m = build_model(fm, {"embedding_dim": 128})
    assert isinstance(m, GeneralFiBiNET) and isinstance(build_model(None, {"embedding_dim": 128}), MM_FiBiNET)
    # {first row of the table, vocabulary, first id column, bag length, padding id}
    assert m.field_desc.tolist() == [[0, 50, 0, 1, -1], [50, 11, 1, 1, -1], [50, 11, 2, 1, -1], [61, 100, 3, 1, 0], [61, 100, 4, 20, 0],
                                     [161, 30, 24, 5, 0]]
    assert m.col_field.tolist() == [0, 1, 2, 3] + [4] * 20 + [5] * 5 and m.id_cols == 29
    assert m.emb.weight.shape == (191, 128) and torch.all(m.emb.weight[61] == 0) and m._pad_rows == [61]
    assert m.k1 == (6 + 15) * 128 and m.senet.reduced_size == 2 and len(m.bilinear.weights()) == 5
    for bad in ([("a", 5)], [("a", 5), ("a", 6)], [("a", 5), {"name": "b", "table": "zzz"}], [("a", 0), ("b", 3)],
                [("a", 5), {"name": "b", "vocab": 3, "bag": 0}]):
        with pytest.raises(ValueError):
            GeneralFiBiNET(bad)
    with pytest.raises(RuntimeError):                       # no CPU path
        m({"ids": torch.zeros(4, 29, dtype=torch.int64)})


def test_fused_adagrad_is_a_torch_optimizer_and_flat_layout_ends_with_the_mlp1_bucket():
    from ctr_recommendation_b200 import FusedAdagrad, FusedAdam, build_model
    model = build_model(None, {"embedding_dim": 128})
    opt = FusedAdagrad(model, lr=1e-2, weight_decay=1e-5)
    assert set(opt.param_groups[0]) >= {"lr", "lr_decay", "eps", "weight_decay", "initial_accumulator_value"}
    sched = torch.optim.lr_scheduler.OneCycleLR(opt, max_lr=0.1, total_steps=10, cycle_momentum="betas" in opt.defaults)
    sched.step()
    assert opt.param_groups[0]["lr"] != 1e-2
    assert all(id(p) != id(model.user_emb.weight) for p in opt.param_groups[0]["params"])      # user_emb: no state, like grad None
    model._ensure_flat()
    off = model._bucket1_offset()
    names = {id(p): n for n, p in model.named_parameters()}
    tail = [names[id(p)] for (f, pl), (o, _) in zip(model._dense_params(), model._layout) if o >= off for p in pl]
    assert tail == ["mlp.0.weight", "mlp.0.bias", "mlp.1.weight", "mlp.1.bias"]
    assert model._flat.numel() - off == 512 * 2688 + 3 * 512
    with pytest.raises(TypeError):
        FusedAdam(torch.nn.Linear(2, 2))
