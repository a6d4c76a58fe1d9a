"""ctr_recommendation_b200 -- B200-native (sm_100a) FiBiNET hot path behind the reference's Python API.

Public surface (mirrors the reference's src/model_fibinet.py + the optimizer step of train_fibinet.py):
    build_model, MM_FiBiNET, SENetLayer, BilinearInteraction, FusedAdam, FusedAdagrad, clip_grad_norm_
"""
from .model import MM_FiBiNET, SENetLayer, BilinearInteraction, build_model  # noqa: F401
from .general import GeneralFiBiNET  # noqa: F401
from .optim import FusedAdagrad, FusedAdam, clip_grad_norm_  # noqa: F401

__all__ = ["MM_FiBiNET", "SENetLayer", "BilinearInteraction", "build_model", "GeneralFiBiNET", "FusedAdam", "FusedAdagrad", "clip_grad_norm_"]
