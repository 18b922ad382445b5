"""Shared helpers for the parity tests (imports the oracle: test infrastructure only)."""
import copy
import os

import numpy as np
import torch

from diffdock_pocket_b200 import inputs, so3, torus, utils
from diffdock_pocket_b200.hetero import Batch
from oracle import diffusion_ref as D, factory, pyg_mini, sampling_ref as S

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
_CACHE = {}


def rel_err(a, b):
    """max |a - b| / max |b|  -- the 'relative' of the north-star parity gates."""
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    if b.numel() == 0:
        assert a.numel() == 0
        return 0.0
    return float((a - b).abs().max() / b.abs().max().clamp(min=1e-30))


def rel_err_cols(a, b, floor=1e-3):
    """Per-channel version of the gate: for every feature column, max |a - b| over rows / max |b| over rows of THAT column
    (columns whose magnitude is below ``floor`` x the global maximum are normalised by that floor instead, so that an
    all-zero channel does not divide by zero).  A global-max normalisation hides errors on small channels; this does not."""
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    if b.numel() == 0:
        assert a.numel() == 0
        return 0.0
    a, b = a.reshape(-1, a.shape[-1]) if a.dim() > 1 else a.reshape(-1, 1), b.reshape(-1, b.shape[-1]) if b.dim() > 1 else b.reshape(-1, 1)
    scale = b.abs().max(0).values.clamp(min=floor * float(b.abs().max().clamp(min=1e-30)))
    return float(((a - b).abs().max(0).values / scale).max())


def rel_err_elem(a, b, floor=1e-2):
    """Elementwise relative error |a - b| / (|b| + floor * max |b|)."""
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    if b.numel() == 0:
        assert a.numel() == 0
        return 0.0
    return float(((a - b).abs() / (b.abs() + floor * b.abs().max().clamp(min=1e-30))).max())


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
    S.randomize_position(dl, False, False, sa.tr_sigma_max, flexible_sidechains='flexResidues' in g and 'edge_idx' in g['flexResidues'])
    return dl


def oracle_list(dl):
    """Deep copy of a list of graphs as ``oracle.pyg_mini`` objects (the oracle never sees the product's ``hetero`` classes)."""
    return [pyg_mini.from_any(g) for g in dl]


def oracle_batch_at(dl, t):
    """The oracle's own collate + set_time (independent of the product's ``hetero.Batch``)."""
    b = pyg_mini.Batch.from_data_list(oracle_list(dl))
    D.set_time(b, t, t, t, t, len(dl))
    return b


def batch_at(dl, t):
    b = Batch.from_data_list(copy.deepcopy(dl))
    D.set_time(b, t, t, t, t, len(dl))
    return b
