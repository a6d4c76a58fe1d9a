"""Deterministic synthetic MicroLens_1M_x1-shaped data and weights.

TEST / BENCH INFRASTRUCTURE. Shared by tests/, bench.py and the golden-vector
generator so that the reference (in the dev container), the numpy oracle and
the CUDA path all see bit-identical inputs without shipping 63 MB of weights.

Everything is produced by a counter-based splitmix64 hash evaluated with numpy
integer arithmetic, so the streams do not depend on numpy's Generator
implementation.  Shapes follow SURVEY.md section 8(d):

* item_id   in [1, 91718)            (reference: src/model_fibinet.py:100)
* likes_level / views_level in [0, 11)   (src/model_fibinet.py:102)
* item_seq  (B, L) int64, left padded with 0 (src/dataloader.py:111-116)
* item_emb_d128 L2-normalised fp32 rows, row 0 zero (Notebooks/task-1.ipynb:237-239)
* the loader hands scalars to the model as float64 (src/dataloader.py:21-48)
"""
from __future__ import annotations

import numpy as np

V_ITEM = 91718
V_USER = 20000
V_CATE = 11
D_MM = 128
NUM_FIELDS = 6
HIDDEN = (512, 256)

_GOLD = np.uint64(0x9E3779B97F4A7C15)
_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)


def _mix(z: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        z = (z ^ (z >> np.uint64(30))) * _M1
        z = (z ^ (z >> np.uint64(27))) * _M2
        return z ^ (z >> np.uint64(31))


def _bits(seed: int, stream: int, n: int, offset: int = 0) -> np.ndarray:
    with np.errstate(over="ignore"):
        base = _mix(np.uint64(seed) * _GOLD + np.uint64(stream) * _M1 + np.uint64(0x1234567))
        idx = np.arange(offset, offset + n, dtype=np.uint64)
        return _mix((idx + np.uint64(1)) * _GOLD + base)


def uniform(seed: int, stream: int, n: int) -> np.ndarray:
    """float64 uniforms in [0, 1)."""
    return (_bits(seed, stream, n) >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def randint(seed: int, stream: int, n: int, lo: int, hi: int) -> np.ndarray:
    """int64 integers in [lo, hi)."""
    return (lo + np.floor(uniform(seed, stream, n) * (hi - lo))).astype(np.int64)


def normal(seed: int, stream: int, n: int) -> np.ndarray:
    """float64 standard normals (Box-Muller on two hashed uniform streams)."""
    u1 = uniform(seed, 2 * stream + 1000, n)
    u2 = uniform(seed, 2 * stream + 1001, n)
    return np.sqrt(-2.0 * np.log(1.0 - u1)) * np.cos(2.0 * np.pi * u2)


def state_dict_shapes(emb_dim: int = 128, bilinear_type: str = "all", num_fields: int = NUM_FIELDS, senet_reduction: int = 2):
    """The 28 state_dict keys of MM_FiBiNET (src/model_fibinet.py:92-136)."""
    D = emb_dim
    pairs = num_fields * (num_fields - 1) // 2
    k1 = (num_fields + pairs) * D
    red = max(1, num_fields // senet_reduction)      # SENetLayer, src/model_fibinet.py:13 (the reference passes ratio 2, :114)
    shapes = {
        "item_emb.weight": (V_ITEM, D),
        "user_emb.weight": (V_USER, D),
        "cate_emb.weight": (V_CATE, D),
        "mm_proj.0.weight": (D, D_MM),
        "mm_proj.0.bias": (D,),
        "mm_proj.1.weight": (D,),
        "mm_proj.1.bias": (D,),
        "senet.excitation.0.weight": (red, num_fields),
        "senet.excitation.0.bias": (red,),
        "senet.excitation.2.weight": (num_fields, red),
        "senet.excitation.2.bias": (num_fields,),
    }
    if bilinear_type == "all":
        shapes["bilinear.W"] = (D, D)
    else:
        for i in range(num_fields - 1):
            shapes[f"bilinear.W_list.{i}"] = (D, D)
    shapes.update({
        "mlp.0.weight": (HIDDEN[0], k1),
        "mlp.0.bias": (HIDDEN[0],),
        "mlp.1.weight": (HIDDEN[0],),
        "mlp.1.bias": (HIDDEN[0],),
        "mlp.1.running_mean": (HIDDEN[0],),
        "mlp.1.running_var": (HIDDEN[0],),
        "mlp.1.num_batches_tracked": (),
        "mlp.4.weight": (HIDDEN[1], HIDDEN[0]),
        "mlp.4.bias": (HIDDEN[1],),
        "mlp.5.weight": (HIDDEN[1],),
        "mlp.5.bias": (HIDDEN[1],),
        "mlp.5.running_mean": (HIDDEN[1],),
        "mlp.5.running_var": (HIDDEN[1],),
        "mlp.5.num_batches_tracked": (),
        "mlp.8.weight": (1, HIDDEN[1]),
        "mlp.8.bias": (1,),
    })
    return shapes


def make_weights(seed: int = 7, emb_dim: int = 128, bilinear_type: str = "all", senet_reduction: int = 2) -> dict:
    """A full, non-trivial MM_FiBiNET state_dict as numpy arrays (fp32; counters int64).

    Scales mimic torch's default initialisers (src/model_fibinet.py:100-135) but every
    affine/normalisation parameter and running statistic is perturbed away from its
    trivial value so that parity tests exercise them.
    """
    out = {}
    for s, (name, shape) in enumerate(state_dict_shapes(emb_dim, bilinear_type, senet_reduction=senet_reduction).items()):
        n = int(np.prod(shape)) if shape else 1
        if name.endswith("num_batches_tracked"):
            out[name] = np.array(3, dtype=np.int64)
            continue
        g = normal(seed, s, n)
        if name in ("item_emb.weight", "user_emb.weight", "cate_emb.weight"):
            w = g
        elif name.endswith("running_var"):
            w = 0.5 + uniform(seed, 500 + s, n)
        elif name.endswith("running_mean"):
            w = 0.1 * g
        elif name in ("mm_proj.1.weight", "mlp.1.weight", "mlp.5.weight"):
            w = 1.0 + 0.1 * g
        elif name in ("mm_proj.1.bias", "mlp.1.bias", "mlp.5.bias"):
            w = 0.1 * g
        elif name == "senet.excitation.0.weight":
            w = 3.0 * (2.0 * uniform(seed, 500 + s, n) - 1.0) / np.sqrt(shape[-1])   # keep the ReLU alive
        elif name == "senet.excitation.0.bias":
            w = 0.1 * (2.0 * uniform(seed, 500 + s, n) - 1.0)
        elif name.startswith("bilinear."):
            w = g * np.sqrt(2.0 / (2 * emb_dim))  # xavier_normal_, src/model_fibinet.py:49
        elif name.endswith(".weight"):
            w = (2.0 * uniform(seed, 500 + s, n) - 1.0) / np.sqrt(shape[-1])
        else:  # Linear biases
            fan_in = {"mm_proj.0.bias": D_MM, "mlp.0.bias": state_dict_shapes(emb_dim)["mlp.0.weight"][1],
                      "mlp.4.bias": HIDDEN[0], "mlp.8.bias": HIDDEN[1]}.get(name, NUM_FIELDS)
            w = (2.0 * uniform(seed, 500 + s, n) - 1.0) / np.sqrt(fan_in)
        w = w.astype(np.float32).reshape(shape)
        if name == "item_emb.weight":
            w[0] = 0.0  # padding_idx=0, src/model_fibinet.py:100
        out[name] = w
    return out


def make_item_mm_table(seed: int = 11, rows: int = V_ITEM) -> np.ndarray:
    """Frozen (rows, 128) fp32 item_emb_d128 matrix, L2-normalised, row 0 zero."""
    g = normal(seed, 77, rows * D_MM).reshape(rows, D_MM)
    g /= np.maximum(np.linalg.norm(g, axis=1, keepdims=True), 1e-12)
    g[0] = 0.0
    return g.astype(np.float32)


def _zipf_ids(u: np.ndarray, vocab: int, a: float = 1.05) -> np.ndarray:
    # inverse-CDF sampling of a truncated zeta law over ids 1..vocab-1
    ranks = np.arange(1, vocab, dtype=np.float64)
    cdf = np.cumsum(ranks ** (-a))
    cdf /= cdf[-1]
    return (np.searchsorted(cdf, u, side="left") + 1).astype(np.int64)


def make_batch(seed: int, batch: int, max_len: int = 20, id_dist: str = "uniform",
               index_dtype=np.float64, with_seq: bool = True, mm_table: np.ndarray | None = None,
               edge_cases: bool = True, vocab: int = V_ITEM) -> tuple[dict, np.ndarray]:
    """One collated batch exactly as the reference loader delivers it (numpy arrays).

    Returns (batch_dict, labels).  Scalars use ``index_dtype`` (the reference train loader
    yields float64, the test loader int64 -- SURVEY fact 8); item_seq is int64.
    """
    B, L = batch, max_len
    u = uniform(seed, 1, B)
    if id_dist == "zipf":
        item_id = _zipf_ids(u, vocab)
    else:
        item_id = (1 + np.floor(u * (vocab - 1))).astype(np.int64)
    likes = randint(seed, 2, B, 0, V_CATE)
    views = randint(seed, 3, B, 0, V_CATE)
    us = uniform(seed, 4, B * L).reshape(B, L)
    if id_dist == "zipf":
        seq = _zipf_ids(us.reshape(-1), vocab).reshape(B, L)
    else:
        seq = (1 + np.floor(us * (vocab - 1))).astype(np.int64)
    # valid history length: 70 % full, rest U[0, L]; left padded with zeros
    full = uniform(seed, 5, B) < 0.7
    vlen = np.where(full, L, randint(seed, 6, B, 0, L + 1))
    pos = np.arange(L)[None, :]
    seq = np.where(pos >= (L - vlen)[:, None], seq, 0)
    if edge_cases and B >= 8:
        item_id[1] = 0            # padding id as the target item -> zero row, zero grad
        seq[2, :] = 0             # all-padding history -> count clamp 1
        seq[3, :] = item_id[3]    # heavy duplicate inside one sample
        item_id[5] = item_id[4]   # duplicate ids across samples
        seq[6, -1] = item_id[4]
        item_id[7] = vocab - 1    # last row of the table
    if mm_table is None:
        g = normal(seed, 9, B * D_MM).reshape(B, D_MM)
        g /= np.maximum(np.linalg.norm(g, axis=1, keepdims=True), 1e-12)
        mm = g.astype(np.float32)
    else:
        mm = mm_table[item_id % mm_table.shape[0]]
    # planted logistic model so that AUC is learnable
    wv = normal(seed + 1000, 10, D_MM)
    item_bias = normal(4242, 11, vocab) * 0.8
    logit = (0.25 * (likes - 5) - 0.15 * (views - 5) + 2.0 * (mm.astype(np.float64) @ wv)
             + item_bias[item_id] + 0.3 * item_bias[seq].sum(1) / np.maximum((seq != 0).sum(1), 1))
    p = 1.0 / (1.0 + np.exp(-logit))
    labels = (uniform(seed, 12, B) < p).astype(np.float32)
    batch_dict = {
        "user_id": randint(seed, 13, B, 1, 1_000_000).astype(index_dtype),
        "item_id": item_id.astype(index_dtype),
        "likes_level": likes.astype(index_dtype),
        "views_level": views.astype(index_dtype),
        "item_emb_d128": mm,
    }
    if with_seq:
        batch_dict["item_seq"] = seq.astype(np.int64)
    return batch_dict, labels


def make_dropout_masks(seed: int, batch: int, p: float = 0.2):
    """Two uint8 keep-masks (B,512), (B,256) with keep probability 1-p."""
    m1 = (uniform(seed, 21, batch * HIDDEN[0]) >= p).astype(np.uint8).reshape(batch, HIDDEN[0])
    m2 = (uniform(seed, 22, batch * HIDDEN[1]) >= p).astype(np.uint8).reshape(batch, HIDDEN[1])
    return m1, m2
