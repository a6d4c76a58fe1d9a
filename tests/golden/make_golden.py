"""Generate tests/golden/*.npz from the UNMODIFIED reference (dev container only).

Usage:  python tests/golden/make_golden.py          (needs /root/reference)

The reference has no golden vectors of its own, so these files are the
pin for oracle/fibinet_numpy.py and, through it, for the CUDA path.  Inputs and weights are
regenerated from oracle/synth.py seeds at test time; only the reference's *outputs* are
stored (full tensors when small, strided samples + norms when large).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import synth  # noqa: E402

REF_SRC = "/root/reference/src"
OUT = os.path.dirname(os.path.abspath(__file__))

SAMPLE_STRIDE = 997  # prime stride for sampling big tensors


def ref_modules():
    sys.path.insert(0, REF_SRC)
    import model_fibinet as ref  # the reference's own file
    sys.path.pop(0)
    return ref


def load_ref_model(ref, weights, bilinear_type="all"):
    model = ref.build_model(None, {"embedding_dim": 128})
    if bilinear_type != "all":
        model.bilinear = ref.BilinearInteraction(128, 6, bilinear_type)
    sd = {k: torch.from_numpy(np.array(v)) for k, v in weights.items()}
    model.load_state_dict(sd, strict=True)
    return model


def tbatch(b):
    return {k: torch.from_numpy(v) for k, v in b.items()}


def sample(x: np.ndarray) -> np.ndarray:
    return np.ascontiguousarray(x.reshape(-1)[::SAMPLE_STRIDE])


def summarize(name, x: np.ndarray, out: dict, prefix: str):
    x = np.asarray(x)
    if x.size <= 70000:
        out[f"{prefix}/{name}/full"] = x
    else:
        out[f"{prefix}/{name}/sample"] = sample(x)
    out[f"{prefix}/{name}/norm"] = np.array(np.sqrt((x.astype(np.float64) ** 2).sum()))
    out[f"{prefix}/{name}/sum"] = np.array(x.astype(np.float64).sum())


class MaskCapture:
    """Recover the keep-mask nn.Dropout drew (out = in * mask / (1-p)) via forward hooks."""

    def __init__(self, model):
        self.masks = {}
        for idx in (3, 7):
            model.mlp[idx].register_forward_hook(self._hook(idx))

    def _hook(self, idx):
        def fn(mod, inp, out):
            x = inp[0].detach()
            m = (out.detach() != 0) | (x == 0)   # where the input is 0 the mask is irrelevant
            self.masks[idx] = m.to(torch.uint8).numpy()
        return fn


def case_train(ref, out, tag, B, steps, id_dist, index_dtype, total_steps=40, with_seq=True):
    weights = synth.make_weights(seed=7)
    model = load_ref_model(ref, weights)
    model.train()
    cap = MaskCapture(model)
    torch.manual_seed(2025)
    # exactly src/train_fibinet.py:74-92
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-5)
    loss_fn = torch.nn.BCELoss()
    sched = torch.optim.lr_scheduler.OneCycleLR(opt, max_lr=1e-2, total_steps=total_steps, pct_start=0.3,
                                                div_factor=25.0, final_div_factor=1000.0)
    out[f"{tag}/meta"] = np.array([B, steps, total_steps, int(with_seq)], dtype=np.int64)
    for s in range(steps):
        batch, labels = synth.make_batch(seed=100 + s, batch=B, id_dist=id_dist, index_dtype=index_dtype,
                                         with_seq=with_seq)
        out[f"{tag}/step{s}/lr_beta1"] = np.array([opt.param_groups[0]["lr"], opt.param_groups[0]["betas"][0]])
        opt.zero_grad()
        y = model(tbatch(batch))
        loss = loss_fn(y, torch.from_numpy(labels))
        loss.backward()
        if s == 0:
            for k, p in model.named_parameters():
                if p.grad is not None:
                    summarize(k, p.grad.numpy(), out, f"{tag}/grad0")
            out[f"{tag}/grad0/user_emb_is_none"] = np.array(model.user_emb.weight.grad is None)
        total = torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm=10.0)
        opt.step()
        sched.step()
        out[f"{tag}/step{s}/prob"] = y.detach().numpy()
        out[f"{tag}/step{s}/loss"] = np.array(loss.item())
        out[f"{tag}/step{s}/total_norm"] = np.array(float(total))
        out[f"{tag}/step{s}/mask1"] = np.packbits(cap.masks[3], axis=None)
        out[f"{tag}/step{s}/mask2"] = np.packbits(cap.masks[7], axis=None)
    sd = model.state_dict()
    for k, v in sd.items():
        summarize(k, v.numpy(), out, f"{tag}/final")
    for k, p in model.named_parameters():
        st = opt.state.get(p, None)
        if st:
            summarize(k, st["exp_avg"].numpy(), out, f"{tag}/final_m")
            summarize(k, st["exp_avg_sq"].numpy(), out, f"{tag}/final_v")
    # eval-mode scoring with the trained weights (Prediction.py path)
    model.eval()
    batch, _ = synth.make_batch(seed=900, batch=300, id_dist=id_dist, index_dtype=np.int64)
    with torch.no_grad():
        out[f"{tag}/eval_prob"] = model(tbatch(batch)).numpy()


def case_eval(ref, out):
    weights = synth.make_weights(seed=7)
    model = load_ref_model(ref, weights).eval()
    for tag, kw in {
        "eval/f64_b300": dict(batch=300, index_dtype=np.float64),
        "eval/i64_b1": dict(batch=1, index_dtype=np.int64, edge_cases=False),
        "eval/i32_b777_zipf": dict(batch=777, index_dtype=np.int32, id_dist="zipf"),
        "eval/noseq_b64": dict(batch=64, index_dtype=np.int64, with_seq=False),
        "eval/short_b50_l7": dict(batch=50, index_dtype=np.int64, max_len=7),
    }.items():
        batch, _ = synth.make_batch(seed=321, **kw)
        with torch.no_grad():
            y = model(tbatch(batch))
        out[f"{tag}/prob"] = y.numpy()
    # intermediate: stacked fields must be bit-exact for the pure gathers
    batch, _ = synth.make_batch(seed=321, batch=300, index_dtype=np.float64)
    tb = tbatch(batch)
    with torch.no_grad():
        item = model.item_emb(tb["item_id"].long())
        like = model.cate_emb(tb["likes_level"].long())
        img = model.mm_proj(tb["item_emb_d128"].float())
        seq_emb = model.item_emb(tb["item_seq"])
        mask = tb["item_seq"] == 0
        hist = (seq_emb * (~mask.unsqueeze(-1)).float()).sum(1) / (~mask).float().sum(1, keepdim=True).clamp(min=1)
    out["eval/f64_b300/item_f"] = item.numpy()
    out["eval/f64_b300/like_f"] = like.numpy()
    out["eval/f64_b300/img_f"] = img.numpy()
    out["eval/f64_b300/hist_f"] = hist.numpy()


def case_modules(ref, out):
    """Stand-alone SENetLayer / BilinearInteraction("all"/"each") forward+backward."""
    B, F, D = 37, 6, 128
    x = synth.normal(5, 1, B * F * D).reshape(B, F, D).astype(np.float32)
    gy = synth.normal(5, 2, B * F * D).reshape(B, F, D).astype(np.float32)
    for ratio in (2, 3):
        torch.manual_seed(0)
        m = ref.SENetLayer(F, reduction_ratio=ratio)
        xt = torch.from_numpy(x).requires_grad_(True)
        y = m(xt)
        y.backward(torch.from_numpy(gy))
        t = f"senet_r{ratio}"
        for k, v in m.state_dict().items():
            out[f"{t}/w/{k}"] = v.numpy()
        out[f"{t}/y"] = y.detach().numpy()
        out[f"{t}/dx"] = xt.grad.numpy()
        for k, p in m.named_parameters():
            out[f"{t}/g/{k}"] = p.grad.numpy()
    gp = synth.normal(5, 3, B * 15 * D).reshape(B, 15, D).astype(np.float32)
    for bt in ("all", "each"):
        torch.manual_seed(1)
        m = ref.BilinearInteraction(D, F, bt)
        xt = torch.from_numpy(x).requires_grad_(True)
        y = m(xt)
        y.backward(torch.from_numpy(gp))
        t = f"bilinear_{bt}"
        for k, v in m.state_dict().items():
            out[f"{t}/w/{k}"] = v.numpy()
        out[f"{t}/y"] = y.detach().numpy()
        out[f"{t}/dx"] = xt.grad.numpy()
        for k, p in m.named_parameters():
            out[f"{t}/g/{k}"] = p.grad.numpy()


def case_each_model(ref, out):
    """Full model with BilinearInteraction("each") swapped in (reachable only by constructing the
    class directly, SURVEY a8')."""
    weights = synth.make_weights(seed=7, bilinear_type="each")
    model = load_ref_model(ref, weights, "each").eval()
    batch, _ = synth.make_batch(seed=321, batch=130, index_dtype=np.int64)
    with torch.no_grad():
        out["each_model/prob"] = model(tbatch(batch)).numpy()


def case_schedule(out):
    """OneCycleLR lr / beta1 trajectory as built at src/train_fibinet.py:84-92."""
    p = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.Adam([p], lr=1e-3, weight_decay=1e-5)
    sched = torch.optim.lr_scheduler.OneCycleLR(opt, max_lr=1e-2, epochs=4, steps_per_epoch=25, pct_start=0.3,
                                                div_factor=25.0, final_div_factor=1000.0)
    traj = []
    for _ in range(100):
        traj.append((opt.param_groups[0]["lr"], opt.param_groups[0]["betas"][0]))
        opt.step()
        if len(traj) < 100:
            sched.step()
    out["schedule/lr_beta1"] = np.array(traj)


def main():
    torch.set_num_threads(8)
    ref = ref_modules()
    out = {}
    case_train(ref, out, "train_u", B=256, steps=3, id_dist="uniform", index_dtype=np.float64)
    case_train(ref, out, "train_z", B=192, steps=2, id_dist="zipf", index_dtype=np.float64, total_steps=10)
    case_eval(ref, out)
    case_modules(ref, out)
    case_each_model(ref, out)
    case_schedule(out)
    path = os.path.join(OUT, "fibinet_golden.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {len(out)} arrays, {os.path.getsize(path) / 1e6:.2f} MB")


if __name__ == "__main__":
    main()
