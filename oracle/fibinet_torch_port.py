"""CPU baseline port: the reference's train / eval step restated with torch.nn.functional on CPU.

TEST / BENCH INFRASTRUCTURE (bench.py's cpu_baseline and --impl reference legs only).  The reference
itself (/root/reference) does not exist on the GPU box, so this file restates its op stream --
the same ATen CPU kernels in the same order as src/model_fibinet.py:138-199 and the loop body
src/train_fibinet.py:113-121 -- from a plain dict of tensors.  It is validated against the numpy
oracle (and therefore the golden vectors) in tests/test_cpu_port.py.  kind = "port".
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def tensors_from_numpy(weights: dict, requires_grad=True) -> dict:
    out = {}
    for k, v in weights.items():
        t = torch.from_numpy(v.copy()) if hasattr(v, "shape") and v.shape != () else torch.tensor(int(v))
        if t.dtype == torch.float32 and requires_grad and "running" not in k and k != "user_emb.weight":
            t.requires_grad_(True)
        out[k] = t
    return out


def forward(P: dict, batch: dict, train: bool, dropout_p: float = 0.2) -> torch.Tensor:
    item_id = batch["item_id"].long()                                     # ref :140
    mm = batch["item_emb_d128"].float()                                   # ref :141
    likes, views = batch["likes_level"].long(), batch["views_level"].long()
    seq = batch.get("item_seq")
    like_f = F.embedding(likes, P["cate_emb.weight"])                     # ref :155-156
    view_f = F.embedding(views, P["cate_emb.weight"])
    item_f = F.embedding(item_id, P["item_emb.weight"], padding_idx=0)    # ref :159
    img_f = F.relu(F.layer_norm(F.linear(mm, P["mm_proj.0.weight"], P["mm_proj.0.bias"]), (128,),
                                P["mm_proj.1.weight"], P["mm_proj.1.bias"]))   # ref :162
    if seq is not None:
        e = F.embedding(seq, P["item_emb.weight"], padding_idx=0)         # ref :167 (B,L,128) materialised
        keep = (seq != 0)
        e = e * keep.unsqueeze(-1).float()
        hist_f = e.sum(1) / keep.float().sum(1, keepdim=True).clamp(min=1)  # ref :172-174
    else:
        hist_f = torch.zeros_like(item_f)
    X = torch.stack([torch.zeros_like(item_f), like_f, view_f, item_f, img_f, hist_f], 1)   # ref :180-182
    z = X.mean(-1)                                                         # ref :28
    s = torch.sigmoid(F.linear(F.relu(F.linear(z, P["senet.excitation.0.weight"], P["senet.excitation.0.bias"])),
                               P["senet.excitation.2.weight"], P["senet.excitation.2.bias"]))
    V = X * s.unsqueeze(-1)                                                # ref :35
    vid = torch.matmul(V, P["bilinear.W"])                                 # ref :72
    pairs = [V[:, i] * vid[:, j] for i in range(6) for j in range(i + 1, 6)]   # ref :75-79
    Cm = torch.cat([V.reshape(V.shape[0], -1), torch.stack(pairs, 1).reshape(V.shape[0], -1)], 1)   # ref :191-194
    h = F.linear(Cm, P["mlp.0.weight"], P["mlp.0.bias"])
    h = F.batch_norm(h, P["mlp.1.running_mean"], P["mlp.1.running_var"], P["mlp.1.weight"], P["mlp.1.bias"], train, 0.1, 1e-5)
    h = F.dropout(F.relu(h), dropout_p, train)
    h = F.linear(h, P["mlp.4.weight"], P["mlp.4.bias"])
    h = F.batch_norm(h, P["mlp.5.running_mean"], P["mlp.5.running_var"], P["mlp.5.weight"], P["mlp.5.bias"], train, 0.1, 1e-5)
    h = F.dropout(F.relu(h), dropout_p, train)
    return torch.sigmoid(F.linear(h, P["mlp.8.weight"], P["mlp.8.bias"])).squeeze(-1)   # ref :197-199


class Trainer:
    """optimizer.zero_grad -> forward -> BCELoss -> backward -> clip_grad_norm_(10) -> Adam.step (ref train_fibinet.py:113-121)."""

    def __init__(self, P: dict, lr=1e-3, weight_decay=1e-5):
        self.P = P
        self.params = [t for t in P.values() if t.requires_grad]
        self.opt = torch.optim.Adam(self.params, lr=lr, weight_decay=weight_decay)
        self.loss_fn = torch.nn.BCELoss()

    def step(self, batch: dict, labels: torch.Tensor) -> float:
        self.opt.zero_grad()
        y = forward(self.P, batch, True)
        loss = self.loss_fn(y, labels)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(self.params, max_norm=10.0)
        self.opt.step()
        return loss.item()
