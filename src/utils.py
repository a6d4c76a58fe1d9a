"""Seeding and metrics (same function names and behaviour as the reference's src/utils.py)."""
import random

import numpy as np
import torch


def set_seed(seed=2025):
    """Seed python / numpy / torch (all devices) and ask cuDNN for deterministic algorithms."""
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)
        torch.backends.cudnn.deterministic = True
        torch.backends.cudnn.benchmark = False


def compute_auc(y_true, y_pred):
    """ROC AUC; 0.5 when y_true holds a single class (the reference catches sklearn's ValueError)."""
    y_true = np.asarray(y_true)
    if np.unique(y_true).size < 2:
        return 0.5
    from sklearn.metrics import roc_auc_score
    return roc_auc_score(y_true, y_pred)


def compute_logloss(y_true, y_pred):
    from sklearn.metrics import log_loss
    return log_loss(y_true, y_pred, labels=[0, 1])
