"""torch_cluster / torch_scatter semantics for the oracle (TEST INFRASTRUCTURE).

Python surface of pytorch-cluster 1.6.1 ``radius`` / ``radius_graph`` / ``knn_graph`` and
pytorch-scatter 2.1.0 ``scatter(reduce='mean')`` as the reference calls them
(models/all_atom_score_model.py:457,524,545-564,607,627; models/score_model.py:117).
The distance scans run in ``cluster_c.c`` (gcc, ``fmaf``) so the fp32 FMA-contracted
arithmetic of the CUDA kernels is reproduced bit for bit.
"""
import ctypes
import os
import subprocess

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, '_build', 'libcluster_oracle.so')
_LIB = None


def build(force=False):
    src = os.path.join(_HERE, 'cluster_c.c')
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        subprocess.check_call(['gcc', '-O2', '-shared', '-fPIC', '-ffp-contract=off', src, '-o', _SO, '-lm'])
    return _SO


def _lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        p = ctypes.c_void_p
        _LIB.oracle_radius.restype = ctypes.c_int64
        _LIB.oracle_radius.argtypes = [p, p, p, p, ctypes.c_int64, ctypes.c_float, ctypes.c_int64, p, p]
        _LIB.oracle_knn.restype = None
        _LIB.oracle_knn.argtypes = [p, p, p, p, ctypes.c_int64, ctypes.c_int64, p, p]
    return _LIB


def _ptr(batch, n, num_examples):
    if batch is None:
        return np.array([0, n], dtype=np.int64)
    deg = np.bincount(batch.cpu().numpy().astype(np.int64), minlength=num_examples)
    return np.concatenate([[0], np.cumsum(deg)]).astype(np.int64)


def _f32(t):
    return np.ascontiguousarray(t.detach().cpu().numpy().astype(np.float32))


def radius(x, y, r, batch_x=None, batch_y=None, max_num_neighbors=32):
    """-> [2, E] int64; row 0 indexes ``y`` (ascending), row 1 indexes ``x`` (first-K by index)."""
    nb = 1
    if batch_x is not None:
        nb = int(max(batch_x.max().item() if batch_x.numel() else 0, batch_y.max().item() if batch_y.numel() else 0)) + 1
    xs, ys = _f32(x), _f32(y)
    px, py = _ptr(batch_x, len(xs), nb), _ptr(batch_y, len(ys), nb)
    cap = int(sum(min(int(px[b + 1] - px[b]), max_num_neighbors) * int(py[b + 1] - py[b]) for b in range(nb)))
    row = np.empty(max(cap, 1), dtype=np.int64)
    col = np.empty(max(cap, 1), dtype=np.int64)
    r2 = np.float32(float(r) * float(r))
    e = _lib().oracle_radius(xs.ctypes.data, ys.ctypes.data, px.ctypes.data, py.ctypes.data, nb,
                             ctypes.c_float(float(r2)), max_num_neighbors, row.ctypes.data, col.ctypes.data)
    return torch.from_numpy(np.stack([row[:e], col[:e]]))


def radius_graph(x, r, batch=None, loop=False, max_num_neighbors=32):
    """edge_index[0] = neighbour, edge_index[1] = centre (flow source_to_target)."""
    e = radius(x, x, r, batch, batch, max_num_neighbors if loop else max_num_neighbors + 1)
    row, col = e[1], e[0]
    if not loop:
        m = row != col
        row, col = row[m], col[m]
    return torch.stack([row, col], 0)


def knn(x, y, k, batch_x=None, batch_y=None):
    nb = 1
    if batch_x is not None:
        nb = int(max(batch_x.max().item(), batch_y.max().item())) + 1
    xs, ys = _f32(x), _f32(y)
    px, py = _ptr(batch_x, len(xs), nb), _ptr(batch_y, len(ys), nb)
    row = np.empty(len(ys) * k, dtype=np.int64)
    col = np.empty(len(ys) * k, dtype=np.int64)
    _lib().oracle_knn(xs.ctypes.data, ys.ctypes.data, px.ctypes.data, py.ctypes.data, nb, k,
                      row.ctypes.data, col.ctypes.data)
    m = col != -1
    return torch.from_numpy(np.stack([row[m], col[m]]))


def knn_graph(x, k, batch=None, loop=False):
    e = knn(x, x, k if loop else k + 1, batch, batch)
    row, col = e[1], e[0]
    if not loop:
        m = row != col
        row, col = row[m], col[m]
    return torch.stack([row, col], 0)


def scatter(src, index, dim=0, dim_size=None, reduce='mean'):
    """torch_scatter.scatter for dim=0: sum / count.clamp(min=1) (empty segments -> 0)."""
    assert dim == 0
    n = int(dim_size) if dim_size is not None else (int(index.max().item()) + 1 if index.numel() else 0)
    out = torch.zeros((n,) + tuple(src.shape[1:]), dtype=src.dtype)
    out.index_add_(0, index.long(), src)
    if reduce in ('sum', 'add'):
        return out
    assert reduce == 'mean'
    cnt = torch.zeros(n, dtype=src.dtype)
    cnt.index_add_(0, index.long(), torch.ones(index.shape[0], dtype=src.dtype))
    cnt.clamp_(min=1)
    return out / cnt.reshape((n,) + (1,) * (src.dim() - 1))


def scatter_mean(src, index, dim=0, dim_size=None):
    return scatter(src, index, dim, dim_size, 'mean')
