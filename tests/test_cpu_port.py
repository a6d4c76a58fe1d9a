"""CPU: the torch-functional baseline port agrees with the numpy oracle (and so with the golden vectors)."""
import numpy as np
import torch

from oracle import fibinet_numpy as orc
from oracle import fibinet_torch_port as port
from oracle import synth
from helpers import rel_err


def test_port_eval_and_grads_match_oracle():
    W = synth.make_weights(seed=7)
    batch, labels = synth.make_batch(seed=100, batch=200, index_dtype=np.float64)
    P = port.tensors_from_numpy(W)
    tb = {k: torch.from_numpy(v) for k, v in batch.items()}
    with torch.no_grad():
        y = port.forward(P, tb, train=False).numpy()
    ref, _ = orc.forward(W, batch, train=False)
    assert rel_err(y, ref) <= 1e-5
    yt = port.forward(P, tb, train=True, dropout_p=0.0)
    torch.nn.BCELoss()(yt, torch.from_numpy(labels)).backward()
    prob, cache = orc.forward(dict(W), batch, train=True, masks=None, update_running=False)
    _, dprob = orc.bce_loss(prob, labels)
    G = orc.backward(W, cache, dprob)
    for k, g in G.items():
        got = P[k].grad.numpy()
        assert np.abs(got - g).max() <= 1e-5 * max(np.abs(g).max(), 1e-30) + 2e-7, k
    assert P["user_emb.weight"].grad is None
