"""IGSO(3) expected score norm table (host precompute; mirrors utils/so3.py of the reference).

The hot path only needs ``score_norm(eps)`` = nearest-index lookup into ``_exp_score_norms[1000]``
(utils/so3.py:85-89, used at models/all_atom_score_model.py:384).  The reference builds the table
with a 2000-term Python loop per (eps, omega) and caches ``.so3_*.npy`` in the CWD
(utils/so3.py:41-60); here the truncated series is evaluated as two [1000x2000]x[2000x2000]
float64 matmuls (same terms, same truncation L=2000) and cached next to this file.
"""
import os

import numpy as np
import torch

MIN_EPS, MAX_EPS, N_EPS = 0.01, 2, 1000
X_N = 2000
_CACHE = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'data', 'so3_exp_score_norms.npy')


def _compute_exp_score_norms(L=2000):
    eps = 10 ** np.linspace(np.log10(MIN_EPS), np.log10(MAX_EPS), N_EPS)
    om = np.linspace(0, np.pi, X_N + 1)[1:]
    l = np.arange(L, dtype=np.float64)
    A = (2 * l + 1)[None, :] * np.exp(-(l * (l + 1))[None, :] * eps[:, None] ** 2)        # [eps, l]
    S = np.sin((l + 0.5)[:, None] * om[None, :])                                        # [l, om]
    C = (l + 0.5)[:, None] * np.cos((l + 0.5)[:, None] * om[None, :])
    lo, dlo = np.sin(om / 2), 0.5 * np.cos(om / 2)
    AS, AC = A @ S, A @ C
    expansion = AS / lo
    dsigma = (lo * AC - dlo * AS) / lo ** 2
    pdf = expansion * (1 - np.cos(om)) / np.pi
    # score**2 * pdf == dsigma**2 * (1 - cos) / (pi * expansion).  Where the alternating series has
    # cancelled to rounding noise (density < 1e-10 of its peak) the reference divides noise by noise;
    # that region carries ~1e-12 of the integral, so it is dropped instead of risking 0 * inf.
    ok = expansion > 1e-10 * expansion.max(axis=1, keepdims=True)
    contrib = np.where(ok, dsigma ** 2 * (1 - np.cos(om)) / np.pi / np.where(ok, expansion, 1.0), 0.0)
    return np.sqrt(np.sum(contrib, axis=1) / np.sum(pdf, axis=1) / np.pi)


def exp_score_norms():
    global _TABLE
    if _TABLE is None:
        if os.path.exists(_CACHE):
            _TABLE = np.load(_CACHE)
        else:
            _TABLE = _compute_exp_score_norms()
            try:
                np.save(_CACHE, _TABLE)
            except OSError:
                pass
    return _TABLE


_TABLE = None


def eps_index(eps):
    idx = (np.log10(eps) - np.log10(MIN_EPS)) / (np.log10(MAX_EPS) - np.log10(MIN_EPS)) * N_EPS
    return np.clip(np.around(idx).astype(int), a_min=0, a_max=N_EPS - 1)


def score_norm_np(eps):
    return exp_score_norms()[eps_index(np.asarray(eps, dtype=np.float64))]


def score_norm(eps):
    """Same signature as the reference: CPU tensor in, float32 CPU tensor out."""
    return torch.from_numpy(np.asarray(score_norm_np(eps.numpy()))).float()
