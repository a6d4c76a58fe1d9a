"""Drop-in for the reference's src/model_fibinet.py: same public names, same constructor and forward
contract, same state_dict -- implemented by ctr_recommendation_b200 (hand-written sm_100a CUDA behind a C ABI).

    from model_fibinet import build_model
    model = build_model(None, model_cfg)          # feature_map is accepted and ignored, like the reference
    y = model(batch_dict)                         # (B,) float32 probabilities, differentiable
"""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from ctr_recommendation_b200.model import (  # noqa: E402,F401
    BilinearInteraction, MM_FiBiNET, SENetLayer, build_model)

__all__ = ["SENetLayer", "BilinearInteraction", "MM_FiBiNET", "build_model"]
