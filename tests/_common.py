"""Shared helpers for the parity tests (imports the oracle: test infrastructure only)."""
import copy
import os

import numpy as np
import torch

from diffdock_pocket_b200 import inputs, so3, torus, utils
from diffdock_pocket_b200.hetero import Batch
from oracle import diffusion_ref as D, factory, sampling_ref as S

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
_CACHE = {}


def rel_err(a, b):
    """max |a - b| / max |b|  -- the 'relative' of the north-star parity gates."""
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    if b.numel() == 0:
        assert a.numel() == 0
        return 0.0
    return float((a - b).abs().max() / b.abs().max().clamp(min=1e-30))


def graph(name='3dpf_holo'):
    return inputs.load_graph_npz(os.path.join(GOLD, name + '.npz'), name=name)


def models(device, small=False):
    """(product score, product confidence, oracle score, oracle confidence, score args, conf args)."""
    key = (str(device), small)
    if key not in _CACHE:
        if small:
            sa = utils.score_model_args(ns=16, nv=4, num_conv_layers=4, sigma_embed_dim=32, distance_embed_dim=32,
                                        cross_distance_embed_dim=32)
            ca = utils.confidence_model_args(ns=8, nv=2, num_conv_layers=3)
        else:
            sa, ca = utils.score_model_args(), utils.confidence_model_args()
        m, c, sa, ca = utils.build_models(device, score_args=sa, conf_args=ca, seed=0)
        om = factory.oracle_model(sa, m.state_dict(), so3.score_norm_np, torus.score_norm)
        oc = factory.oracle_model(ca, c.state_dict(), so3.score_norm_np, torus.score_norm, confidence_mode=True)
        _CACHE[key] = (m, c, om, oc, sa, ca)
    return _CACHE[key]


def randomized_list(g, n, sa, seed=0):
    np.random.seed(seed)
    torch.manual_seed(seed)
    dl = [copy.deepcopy(g) for _ in range(n)]
    S.randomize_position(dl, False, False, sa.tr_sigma_max, flexible_sidechains='flexResidues' in g)
    return dl


def batch_at(dl, t):
    b = Batch.from_data_list(copy.deepcopy(dl))
    D.set_time(b, t, t, t, t, len(dl))
    return b
