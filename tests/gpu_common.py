"""Helpers shared by the -m gpu parity tests (CUDA path through the C ABI vs the numpy oracle)."""
import numpy as np
import torch

from oracle import synth


def to_dev(batch: dict, device="cuda"):
    return {k: torch.from_numpy(np.ascontiguousarray(v)).to(device) for k, v in batch.items()}


def load_weights(model, weights: dict):
    sd = {k: torch.from_numpy(np.array(v)) for k, v in weights.items()}
    model.load_state_dict(sd, strict=True)
    return model


def make_model(precision="fp32", bilinear_type="all", seed=7, train=False, fused=False, senet_reduction=2):
    from ctr_recommendation_b200 import build_model
    fm = {"precision": precision, "bilinear_type": bilinear_type, "senet_reduction": senet_reduction}
    model = build_model(fm, {"embedding_dim": 128})
    load_weights(model, synth.make_weights(seed=seed, bilinear_type=bilinear_type, senet_reduction=senet_reduction))
    model = model.cuda()
    model.train(train)
    return model


def named_grads(model):
    return {k: p.grad.detach().cpu().numpy() for k, p in model.named_parameters() if p.grad is not None}
