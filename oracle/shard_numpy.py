"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the row-sharded item table (BASELINE config 5).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this; the product path never does.

Parity status: *extension, unpinned*.  The reference has no sharded table (it replicates item_emb under
nn.DataParallel, src/train_fibinet.py:69-70); what is restated here is
  * the partition rule the CUDA path uses (row g -> rank g % N, local row g // N),
  * the table gradient the reference's autograd produces for the item_emb lookups (src/model_fibinet.py:159 target item,
    :167-174 history rows divided by the clamped count; nn.Embedding(padding_idx=0) at :100 zeroes row 0), written as a
    scatter-add in float64 -- the sharded exchange must reproduce exactly this dense gradient, slice by slice,
  * torch.optim.Adam's single-tensor update (src/train_fibinet.py:78) applied to the touched rows only ("lazy" Adam, the
    update rule BASELINE config 5 names).
"""
from __future__ import annotations

import math

import numpy as np


def shard_rows(item_rows: int, world: int) -> int:
    return -(-item_rows // world)


def owner_local(rows: np.ndarray, world: int):
    rows = np.asarray(rows, dtype=np.int64)
    return rows % world, rows // world


def table_grad_dense(item_id, item_seq, dXitem, dXhist, item_rows: int, dtype=np.float64) -> np.ndarray:
    """Dense (V, D) gradient of item_emb for one rank's batch: row item_id[b] += dXitem[b]; row item_seq[b, l] += dXhist[b]
    for every non-padding history id (dXhist is already divided by the history count); row 0 stays zero."""
    item_id = np.clip(np.asarray(item_id).astype(np.int64), 0, item_rows - 1)
    g = np.zeros((item_rows, dXitem.shape[1]), dtype=dtype)
    np.add.at(g, item_id, dXitem.astype(dtype))
    if item_seq is not None:
        seq = np.clip(np.asarray(item_seq).astype(np.int64), 0, item_rows - 1)
        B, L = seq.shape
        np.add.at(g, seq.reshape(-1), np.repeat(dXhist.astype(dtype), L, axis=0))
    g[0] = 0
    return g


def slice_of(full: np.ndarray, rank: int, world: int) -> np.ndarray:
    R = shard_rows(full.shape[0], world)
    out = np.zeros((R,) + full.shape[1:], dtype=full.dtype)
    part = full[rank::world]
    out[: part.shape[0]] = part
    return out


def lazy_adam_rows(p, m, v, grad, touched, lr, beta1, beta2, eps, weight_decay, step, coef=1.0, dtype=np.float32):
    """Adam (torch/optim/adam.py single-tensor op order, L2 decay folded into the gradient) on the rows where ``touched``;
    every other row of p / m / v is left exactly as it was.  ``step`` is the 1-based global step (bias corrections)."""
    dt = dtype
    p, m, v = p.copy(), m.copy(), v.copy()
    rows = np.nonzero(touched)[0]
    pp, mm, vv = p[rows].astype(dt), m[rows].astype(dt), v[rows].astype(dt)
    g = (grad[rows].astype(dt) * dt(coef)) + dt(weight_decay) * pp
    mm = mm + dt(1 - beta1) * (g - mm)
    vv = vv * dt(beta2) + dt(1 - beta2) * g * g
    step_size = lr / (1 - beta1 ** step)
    denom = np.sqrt(vv) / dt(math.sqrt(1 - beta2 ** step)) + dt(eps)
    pp = pp - dt(step_size) * (mm / denom)
    p[rows], m[rows], v[rows] = pp, mm, vv
    return p, m, v


def table_grad_ordered(item_id, item_seq, dXitem, dXhist, item_rows: int, hot: int = 256) -> np.ndarray:
    """The same gradient as table_grad_dense but in fp32 with the summation ORDER the CUDA path specifies (segsum.cuh), so that
    the result can be compared bit for bit: a row's occurrences are taken in source order (the B target ids first, then the
    history ids in (sample, position) order); up to `hot` occurrences are added sequentially; longer lists are cut into
    `hot`-sized chunks relative to the row's own start, every chunk is summed sequentially from zero, and the chunk sums are then
    added sequentially in chunk order."""
    item_id = np.clip(np.asarray(item_id).astype(np.int64), 0, item_rows - 1)
    B = item_id.shape[0]
    keys = [item_id]
    src = [np.arange(B)]
    if item_seq is not None:
        seq = np.clip(np.asarray(item_seq).astype(np.int64), 0, item_rows - 1)
        keys.append(seq.reshape(-1))
        src.append(B + np.arange(seq.size))
        L = seq.shape[1]
    keys, src = np.concatenate(keys), np.concatenate(src)
    order = np.argsort(keys, kind="stable")
    keys, src = keys[order], src[order]
    g = np.zeros((item_rows, dXitem.shape[1]), dtype=np.float32)
    start = 0
    n = keys.shape[0]

    def seq_sum(rows):      # sequential fp32 accumulation starting from 0 (cumsum is strictly left to right)
        return np.cumsum(rows, axis=0, dtype=np.float32)[-1]
    while start < n:
        end = start
        while end < n and keys[end] == keys[start]:
            end += 1
        row = keys[start]
        if row != 0:
            s = src[start:end]
            vals = np.where((s < B)[:, None], dXitem[np.minimum(s, B - 1)],
                            dXhist[(np.maximum(s - B, 0) // L) if item_seq is not None else 0]).astype(np.float32)
            if len(s) <= hot:
                g[row] = seq_sum(vals)
            else:
                chunks = np.stack([seq_sum(vals[c:c + hot]) for c in range(0, len(s), hot)])
                g[row] = seq_sum(chunks)
        start = end
    return g
