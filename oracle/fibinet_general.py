"""TEST INFRASTRUCTURE ONLY -- numpy restatement of FiBiNET for an arbitrary number of categorical fields.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this; the product path never does.

This is the oracle for BASELINE config 5 ("scaled synthetic FiBiNET: 40 fields, 100M-row tables"), written ahead of the CUDA
kernels that will implement it.  The reference's model is hard-wired to six fields (src/model_fibinet.py:112,179-182), but its
building blocks are generic in the number of fields and are what is restated here:

  * field stack           F embedding lookups, one table per field (the reference: nn.Embedding at :100-102, gathers at :155-159)
  * SENetLayer            src/model_fibinet.py:5-35  (reduction ratio r -> hidden = max(1, F // r), :13)
  * BilinearInteraction   src/model_fibinet.py:37-89 ("all" / "each"; "interaction" = FiBiNET paper, one matrix per pair)
  * concat + MLP          src/model_fibinet.py:122-135,191-199 (Linear -> BatchNorm1d -> ReLU -> Dropout, twice, Linear, Sigmoid)

Parity status: pinned on CPU against a torch model assembled from the REFERENCE'S OWN SENetLayer / BilinearInteraction classes
and torch autograd (tests/test_oracle_general.py, run where /root/reference exists); "interaction" is an extension, unpinned.
"""
from __future__ import annotations

import numpy as np

from . import fibinet_numpy as orc

EPS_BN = 1e-5


def make_params(num_fields: int, dim: int, vocab: int, hidden=(512, 256), bilinear_type="all", reduction_ratio=2, seed=0,
                dtype=np.float64) -> dict:
    rng = np.random.default_rng(seed)
    F, D = num_fields, dim
    R = max(1, F // reduction_ratio)
    pairs = F * (F - 1) // 2
    nW = {"all": 1, "each": F - 1, "interaction": pairs}[bilinear_type]
    K = (F + pairs) * D
    P = {"tables": [(rng.standard_normal((vocab, D)) * 0.1).astype(dtype) for _ in range(F)],
         "se_w1": (rng.standard_normal((R, F)) * 0.3).astype(dtype), "se_b1": (rng.standard_normal(R) * 0.1).astype(dtype),
         "se_w2": (rng.standard_normal((F, R)) * 0.3).astype(dtype), "se_b2": (rng.standard_normal(F) * 0.1).astype(dtype),
         "bil_w": [(rng.standard_normal((D, D)) / np.sqrt(D)).astype(dtype) for _ in range(nW)],
         "bilinear_type": bilinear_type}
    dims = [K, *hidden]
    for i in range(len(hidden)):
        P[f"w{i}"] = (rng.standard_normal((dims[i + 1], dims[i])) / np.sqrt(dims[i])).astype(dtype)
        P[f"b{i}"] = (rng.standard_normal(dims[i + 1]) * 0.1).astype(dtype)
        P[f"bn_g{i}"] = (1 + 0.1 * rng.standard_normal(dims[i + 1])).astype(dtype)
        P[f"bn_b{i}"] = (0.1 * rng.standard_normal(dims[i + 1])).astype(dtype)
    P["w_out"] = (rng.standard_normal((1, hidden[-1])) / np.sqrt(hidden[-1])).astype(dtype)
    P["b_out"] = np.zeros(1, dtype=dtype)
    P["n_hidden"] = len(hidden)
    return P


def default_spec(num_fields: int):
    """One table and one id column per field: the plain F-lookup model."""
    return [dict(table=f, col=f, bag=1, pad=-1) for f in range(num_fields)]


def field_stack(P: dict, ids: np.ndarray, spec):
    """X (B,F,D) and the per-field counts.  A field is one lookup into its table (nn.Embedding, src/model_fibinet.py:155-159; an id
    equal to the table's padding id gives the zero row, :100) or a bag of ids pooled the way the reference pools item_seq
    (:165-174): rows of the non-padding ids summed and divided by clamp(count, min=1)."""
    B = ids.shape[0]
    X, cnt = [], []
    for f in spec:
        T = P["tables"][f["table"]]
        idf = ids[:, f["col"]:f["col"] + f["bag"]]
        valid = idf != f["pad"]
        rows = T[idf] * valid[..., None]
        if f["bag"] == 1:
            X.append(rows[:, 0]); cnt.append(np.ones(B))
        else:
            c = np.maximum(valid.sum(1), 1).astype(T.dtype)
            X.append(rows.sum(1) / c[:, None]); cnt.append(c)
    return np.stack(X, axis=1), np.stack(cnt, axis=1)


def forward(P: dict, ids: np.ndarray, masks=None, dropout_p: float = 0.0, spec=None, relu_gates=None):
    """Train-mode forward (batch statistics).  ids: (B, id columns) integer (one column per field unless ``spec`` says otherwise).
    Returns (prob (B,), cache).  ``relu_gates``: optional list of boolean arrays replacing the MLP ReLU decisions (same test hook as
    oracle/fibinet_numpy.forward: a pre-activation within rounding distance of zero may legitimately be decided differently by
    another fp32 implementation, and one flipped decision moves a whole BatchNorm column's gradients)."""
    spec = default_spec(ids.shape[1]) if spec is None else spec
    B, F = ids.shape[0], len(spec)
    X, cnt = field_stack(P, ids, spec)                                           # (B, F, D) field stack
    V, se_saved = orc.senet_forward(X, P["se_w1"], P["se_b1"], P["se_w2"], P["se_b2"])
    W = P["bil_w"][0] if P["bilinear_type"] == "all" else P["bil_w"]
    Pm = orc.bilinear_forward(V, W, P["bilinear_type"])
    C = np.concatenate([V.reshape(B, -1), Pm.reshape(B, -1)], axis=1)            # src/model_fibinet.py:191-194
    a, layers = C, []
    for i in range(P["n_hidden"]):
        h = a @ P[f"w{i}"].T + P[f"b{i}"]
        mean, var = h.mean(0), h.var(0)                                          # biased variance normalises (BatchNorm1d, train)
        rstd = 1.0 / np.sqrt(var + EPS_BN)
        xhat = (h - mean) * rstd
        pre = xhat * P[f"bn_g{i}"] + P[f"bn_b{i}"]
        gate = (pre > 0) if relu_gates is None else np.asarray(relu_gates[i], dtype=bool)
        r = np.where(gate, pre, 0.0)
        m = None
        y = r
        if dropout_p > 0:
            m = masks[i].astype(h.dtype) / (1.0 - dropout_p)                      # nn.Dropout: keep-mask scaled by 1/(1-p)
            y = r * m
        layers.append((a, xhat, rstd, gate, m, pre))
        a = y
    logit = (a @ P["w_out"].T + P["b_out"])[:, 0]
    prob = 1.0 / (1.0 + np.exp(-logit))
    return prob, dict(ids=ids, X=X, V=V, se_saved=se_saved, C=C, layers=layers, last=a, prob=prob, spec=spec, cnt=cnt)


def backward(P: dict, cache: dict, dprob: np.ndarray) -> dict:
    """Gradients of everything, embedding tables as dense (vocab, D) arrays (the nn.Embedding(sparse=False) contract)."""
    ids, X, V = cache["ids"], cache["X"], cache["V"]
    B, F, D = X.shape
    G = {}
    p = cache["prob"]
    dlogit = dprob * p * (1 - p)
    G["w_out"] = dlogit[None, :] @ cache["last"]
    G["b_out"] = dlogit.sum(keepdims=True)
    da = dlogit[:, None] * P["w_out"]
    for i in reversed(range(P["n_hidden"])):
        a_in, xhat, rstd, gate, m, _ = cache["layers"][i]
        if m is not None:
            da = da * m
        dy = da * gate
        G[f"bn_g{i}"] = (dy * xhat).sum(0)
        G[f"bn_b{i}"] = dy.sum(0)
        dxhat = dy * P[f"bn_g{i}"]
        dh = rstd * (dxhat - dxhat.mean(0) - xhat * (dxhat * xhat).mean(0))      # BatchNorm backward through the batch statistics
        G[f"w{i}"] = dh.T @ a_in
        G[f"b{i}"] = dh.sum(0)
        da = dh @ P[f"w{i}"]
    dC = da
    dV = dC[:, :F * D].reshape(B, F, D).copy()
    dPm = dC[:, F * D:].reshape(B, -1, D)
    W = P["bil_w"][0] if P["bilinear_type"] == "all" else P["bil_w"]
    dVb, dW = orc.bilinear_backward(V, W, dPm, P["bilinear_type"])
    dV += dVb
    G["bil_w"] = [dW] if P["bilinear_type"] == "all" else dW
    dX, G["se_w1"], G["se_b1"], G["se_w2"], G["se_b2"] = orc.senet_backward(X, P["se_w1"], P["se_w2"], cache["se_saved"], dV)
    G["tables"] = [np.zeros_like(t) for t in P["tables"]]
    for fi, f in enumerate(cache["spec"]):                                       # embedding_dense_backward of every lookup
        idf = ids[:, f["col"]:f["col"] + f["bag"]]
        valid = idf != f["pad"]
        contrib = (dX[:, fi] / cache["cnt"][:, fi, None])[:, None, :] * valid[..., None]
        np.add.at(G["tables"][f["table"]], idf.reshape(-1), contrib.reshape(-1, D))
    return G
