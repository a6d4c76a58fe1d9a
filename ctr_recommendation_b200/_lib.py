"""ctypes binding of libfibinet_b200.so (the C ABI declared in include/fibinet_b200.h).

There is no CPU or PyTorch fallback: if the shared library is missing and cannot be built with
nvcc, importing an op raises.  PyTorch is used by the host only for device memory and streams.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libfibinet_b200.so")

IDX_I32, IDX_I64, IDX_F64, IDX_F32 = 0, 1, 2, 3
PREC_FP32, PREC_TF32X3, PREC_BF16 = 0, 1, 2
BILINEAR_ALL, BILINEAR_EACH, BILINEAR_INTERACTION = 0, 1, 2
PREC_TF32X2 = 3
PREC_F16X3 = 4
BWD_CHAIN, BWD_LEAF1, BWD_LEAF2 = 1, 2, 4
# what the model path accepts.  "f16x3": the long-K MLP GEMMs as three fp16 passes under one power-of-two scale per operand tensor
# (fp32-grade like tf32x3, at the bf16 MMA rate and half the operand bytes); its short-K GEMMs stay tf32x3
PRECISIONS = {"fp32": PREC_FP32, "tf32x3": PREC_TF32X3, "bf16": PREC_BF16, "f16x3": PREC_F16X3}
GEMM_PRECISIONS = dict(PRECISIONS, tf32x2=PREC_TF32X2)                               # fbn_gemm only (K-major x K-major)
BILINEAR_TYPES = {"all": BILINEAR_ALL, "field_all": BILINEAR_ALL, "each": BILINEAR_EACH, "field_each": BILINEAR_EACH,
                  "interaction": BILINEAR_INTERACTION, "field_interaction": BILINEAR_INTERACTION}

_vp, _i64, _i32, _f, _sz, _u64 = C.c_void_p, C.c_int64, C.c_int32, C.c_float, C.c_size_t, C.c_uint64

PARAM_FIELDS = ["item_emb", "cate_emb", "mm_w", "mm_b", "ln_g", "ln_b", "se_w1", "se_b1", "se_w2", "se_b2", "bil_w",
                "w1", "b1", "bn1_g", "bn1_b", "bn1_mean", "bn1_var", "w2", "b2", "bn2_g", "bn2_b", "bn2_mean", "bn2_var",
                "w3", "b3"]
GRAD_FIELDS = ["cate_emb", "mm_w", "mm_b", "ln_g", "ln_b", "se_w1", "se_b1", "se_w2", "se_b2", "bil_w",
               "w1", "b1", "bn1_g", "bn1_b", "w2", "b2", "bn2_g", "bn2_b", "w3", "b3"]


class Params(C.Structure):
    _fields_ = [("item_emb", _vp), ("item_rows", _i64), ("cate_emb", _vp), ("cate_rows", _i64)] + \
               [(n, _vp) for n in PARAM_FIELDS[2:]] + [("bilinear_type", _i32), ("precision", _i32),
                                                       ("n_shards", _i32), ("shard_rank", _i32), ("shard_rows", _i64),
                                                       ("shard", _vp * 16), ("se_hidden", _i32)]


class Grads(C.Structure):
    _fields_ = [(n, _vp) for n in GRAD_FIELDS]


class Batch(C.Structure):
    _fields_ = [("batch", _i64), ("seq_len", _i64), ("item_id", _vp), ("likes_level", _vp), ("views_level", _vp),
                ("item_seq", _vp), ("item_mm", _vp), ("mm_table", _vp), ("idx_dtype", _i32), ("seq_dtype", _i32)]


class ShardPlan(C.Structure):
    _fields_ = [("n_shards", _i32), ("rank", _i32), ("item_rows", _i64), ("shard_rows", _i64), ("cap", _i64),
                ("merge_cap", _i64), ("xchg", _vp * 16)]


class AdamHyper(C.Structure):
    _fields_ = [("lr", _f), ("beta1", _f), ("beta2", _f), ("eps", _f), ("weight_decay", _f), ("step", _i32),
                ("one_minus_beta1", _f), ("one_minus_beta2", _f), ("decoupled", _i32)]


class AdagradHyper(C.Structure):
    _fields_ = [("lr", _f), ("lr_decay", _f), ("eps", _f), ("weight_decay", _f), ("step", _i32)]


_SIGNATURES = {
    "fbn_workspace_bytes": (_sz, [_i64, _i64, _i64]),
    "fbn_workspace_offset": (_sz, [_i64, _i64, _i64, C.c_char_p]),
    "fbn_forward": (C.c_int, [C.POINTER(Params), C.POINTER(Batch), _vp, _sz, C.c_int, _f, _vp, _vp, _u64, _u64, _vp, _vp, _vp]),
    "fbn_embed_forward": (C.c_int, [C.POINTER(Params), C.POINTER(Batch), _vp, _sz, C.c_int, _vp]),
    "fbn_backward": (C.c_int, [C.POINTER(Params), C.POINTER(Batch), _vp, _sz, C.c_int, _f, _vp, C.POINTER(Grads), _vp, _i64,
                               _vp, _vp, C.c_int, C.c_int, _vp, _vp]),
    "fbn_backward_phase": (C.c_int, [C.POINTER(Params), C.POINTER(Batch), _vp, _sz, C.c_int, _f, _vp, C.POINTER(Grads), _vp, _vp, C.c_int,
                                     C.c_int, _vp, C.c_int, _vp]),
    "fbn_embed_index": (C.c_int, [C.POINTER(Params), C.POINTER(Batch), _vp, _sz, _vp, _vp]),
    "fbn_bce_loss": (C.c_int, [_vp, _vp, _i64, _f, _vp, _vp, _vp]),
    "fbn_bce_scratch_bytes": (_sz, []),
    "fbn_bce_loss_ws": (C.c_int, [_vp, _vp, _i64, _f, _vp, _vp, _vp, _sz, _vp]),
    "fbn_clip_coef": (C.c_int, [_vp, C.c_int, _f, _vp, _vp]),
    "fbn_adam_table": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i64, _vp, C.POINTER(AdamHyper), _vp, _vp]),
    "fbn_adam_dense": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _vp, C.POINTER(AdamHyper), _vp, _vp]),
    "fbn_adagrad_table": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _vp, C.POINTER(AdagradHyper), _vp, _vp]),
    "fbn_adagrad_dense": (C.c_int, [_vp, _vp, _vp, _i64, _vp, C.POINTER(AdagradHyper), _vp, _vp]),
    "fbn_onecycle_hyper": (C.c_int, [_vp, C.c_int, _f, _f, _f, _f, _f, _f, _f, _f, _f, _vp, _vp]),
    "fbn_sumsq": (C.c_int, [_vp, _i64, _vp, _vp, _vp]),
    "fbn_sumsq_partial_floats": (_sz, [_i64]),
    "fbn_fields_gather": (C.c_int, [_vp, _vp, _vp, C.c_int, _i64, C.c_int, C.c_int, _vp, _vp, _vp, _vp]),
    "fbn_fields_scatter_bytes": (_sz, [_i64, C.c_int, _i64]),
    "fbn_fields_scatter": (C.c_int, [_vp, _vp, _vp, _vp, _vp, C.c_int, _i64, C.c_int, C.c_int, _i64, _vp, _vp, C.c_int, _vp, _vp, _sz, _vp]),
    "fbn_tower_workspace_bytes": (_sz, [_i64, _i64]),
    "fbn_tower_workspace_offset": (_sz, [_i64, _i64, C.c_char_p]),
    "fbn_tower_forward": (C.c_int, [C.POINTER(Params), _vp, _i64, _i64, _vp, _sz, C.c_int, _f, _vp, _vp, _u64, _u64, _vp, _vp, _vp, _vp]),
    "fbn_tower_backward": (C.c_int, [C.POINTER(Params), _vp, _i64, _i64, _vp, _sz, C.c_int, _f, _vp, C.POINTER(Grads), _vp, _vp]),
    "fbn_bilinear_fwd_ld": (C.c_int, [_vp, _vp, C.c_int, _i64, C.c_int, C.c_int, _vp, _i64, _vp, _sz, C.c_int, _vp]),
    "fbn_bilinear_bwd_ld": (C.c_int, [_vp, _vp, _vp, _i64, _vp, _i64, C.c_int, _i64, C.c_int, C.c_int, _vp, _vp, _vp, _sz, C.c_int, _vp]),
    "fbn_senet_fwd": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i64, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp]),
    "fbn_senet_bwd": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, _vp, _vp, _vp,
                                _sz, _vp]),
    "fbn_senet_scratch_bytes": (_sz, [_i64, C.c_int, C.c_int]),
    "fbn_bilinear_fwd": (C.c_int, [_vp, _vp, C.c_int, _i64, C.c_int, C.c_int, _vp, _vp, _sz, C.c_int, _vp]),
    "fbn_bilinear_bwd": (C.c_int, [_vp, _vp, _vp, C.c_int, _i64, C.c_int, C.c_int, _vp, _vp, _vp, _sz, C.c_int, _vp]),
    "fbn_bilinear_scratch_bytes": (_sz, [_i64, C.c_int, C.c_int, C.c_int]),
    "fbn_gemm": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _i64, _i64, _i64, _i64, _i64, C.c_int, C.c_int, C.c_int, _vp, _sz, _vp]),
    "fbn_gemm_scratch_bytes": (_sz, [_i64, _i64, _i64, C.c_int]),
    "fbn_time_gemm": (C.c_int, [_vp, _vp, _vp, _i64, _i64, _i64, C.c_int, C.c_int, _u64, C.c_int, _vp, _sz, _vp, _sz, C.c_int,
                                C.POINTER(C.c_float), _vp]),
    "fbn_shard_xchg_bytes": (_sz, [_i64]),
    "fbn_shard_ws_bytes": (_sz, [_i64, _i64, C.c_int, _i64]),
    "fbn_shard_index": (C.c_int, [C.POINTER(ShardPlan), C.POINTER(Batch), _vp, _sz, _vp]),
    "fbn_shard_local_sum": (C.c_int, [C.POINTER(ShardPlan), C.POINTER(Batch), _vp, _vp, _vp, _sz, _vp]),
    "fbn_shard_merge": (C.c_int, [C.POINTER(ShardPlan), _vp, _sz, _vp, _vp, _vp, _vp]),
    "fbn_shard_adam_rows": (C.c_int, [C.POINTER(ShardPlan), _vp, _sz, _vp, _vp, _vp, _vp, C.POINTER(AdamHyper), _vp, _vp]),
    "fbn_shard_stats": (C.c_int, [C.POINTER(ShardPlan), _vp, _sz, C.POINTER(C.c_int32), _vp]),
    "fbn_ipc_export": (C.c_int, [_vp, _vp, C.POINTER(C.c_int64)]),
    "fbn_ipc_open": (C.c_int, [_vp, C.POINTER(C.c_void_p)]),
    "fbn_ipc_close": (C.c_int, [_vp]),
    "fbn_time_stage": (C.c_int, [C.POINTER(Params), C.POINTER(Batch), _vp, _sz, C.c_char_p, _vp, _sz, C.c_int, C.POINTER(C.c_float), _vp]),
    "fbn_stage_report": (C.c_int, [C.c_char_p, _sz]),
    "fbn_set_option": (C.c_int, [C.c_char_p, C.c_int]),
    "fbn_launch_count": (_u64, []),
    "fbn_last_error": (C.c_char_p, []),
    "fbn_version": (C.c_char_p, []),
    "fbn_check_device": (C.c_int, [C.c_int]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None


def load(build_if_missing: bool = True) -> C.CDLL:
    """Load (building first if needed) the CUDA library; raises if that is impossible."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        if not build_if_missing:
            raise RuntimeError(f"{LIB_PATH} is missing; run `python -m ctr_recommendation_b200.build`")
        from . import build as _build
        _build.build()
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header / library mismatch: fail loudly
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class FibinetCudaError(RuntimeError):
    pass


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().fbn_last_error().decode(errors="replace")
        raise FibinetCudaError(f"{what or 'libfibinet_b200'} failed with code {rc}: {msg}")


def ptr(t):
    """Raw device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
