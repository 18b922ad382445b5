"""Operator-level drop-ins for the third-party graph ops the reference calls
(``torch_cluster.radius / radius_graph / knn_graph``, ``torch_scatter.scatter``), on the ddp_b200 kernels.

Same argument meaning and return layout as pytorch-cluster 1.6.1: ``radius`` returns
``[2, E]`` int64 with row 0 indexing ``y`` and row 1 indexing ``x``; the graph builders return
``edge_index[0] = neighbour, edge_index[1] = centre``.  These wrappers allocate their result, so they do one
host synchronisation to size it (exactly like torch_cluster's masking); the resident fast path in
``all_atom_score_model`` calls the same kernels with pre-allocated fixed-capacity buffers instead.
"""
import torch

from . import _lib
from ._lib import ptr


def _ptr_from_batch(batch, n, num_examples, device):
    if batch is None:
        return torch.tensor([0, n], dtype=torch.int32, device=device)
    cnt = torch.bincount(batch.to(device).long(), minlength=num_examples)
    return torch.cat([torch.zeros(1, dtype=torch.long, device=device), cnt.cumsum(0)]).to(torch.int32)


def _num_examples(bx, by):
    if bx is None:
        return 1
    m = 0
    for b in (bx, by):
        if b is not None and b.numel():
            m = max(m, int(b.max().item()))
    return m + 1


def _run_radius(x, y, r, batch_x, batch_y, max_num_neighbors, mode, inv_scale=None):
    dev = x.device
    if dev.type != 'cuda':
        raise RuntimeError('ddp_b200 graph ops run on CUDA only (no CPU fallback)')
    x, y = x.float().contiguous(), y.float().contiguous()
    nb = _num_examples(batch_x, batch_y)
    px, py = _ptr_from_batch(batch_x, x.shape[0], nb, dev), _ptr_from_batch(batch_y, y.shape[0], nb, dev)
    n_y = y.shape[0]
    if n_y == 0 or x.shape[0] == 0:
        return torch.zeros(2, 0, dtype=torch.long, device=dev)
    seg = int((px[1:] - px[:-1]).max().item()) if x.shape[0] else 0
    slab_w = max(min(int(max_num_neighbors), seg), 1)
    slab = torch.empty(max(n_y, 1) * slab_w, dtype=torch.int32, device=dev)
    counts = torch.zeros(n_y + 1, dtype=torch.int32, device=dev)
    cap = max(n_y * slab_w, 1)
    edge = torch.empty(2 * cap, dtype=torch.int32, device=dev)
    n_dev = torch.zeros(1, dtype=torch.int32, device=dev)
    _lib.check(_lib.lib().ddp_radius(ptr(x), ptr(y), ptr(px), ptr(py), nb, n_y, ptr(inv_scale), float(r), int(max_num_neighbors),
                                     mode, 0, ptr(slab), slab_w, ptr(counts), ptr(edge), cap, ptr(n_dev), _lib.stream_ptr()),
               'ddp_radius')
    n = int(n_dev.item())
    return torch.stack([edge[:n], edge[cap:cap + n]]).long()


def radius(x, y, r, batch_x=None, batch_y=None, max_num_neighbors=32, inv_scale=None):
    """``inv_scale`` (optional, per-example divisor) fuses the reference's ``pos / cutoff[batch]`` pre-scaling."""
    return _run_radius(x, y, r, batch_x, batch_y, max_num_neighbors, 0, inv_scale)


def radius_graph(x, r, batch=None, loop=False, max_num_neighbors=32, flow='source_to_target'):
    assert not loop and flow == 'source_to_target'
    return _run_radius(x, x, r, batch, batch, max_num_neighbors + 1, 1)


def knn_graph(x, k, batch=None, loop=False, flow='source_to_target'):
    assert not loop and flow == 'source_to_target'
    dev = x.device
    if dev.type != 'cuda':
        raise RuntimeError('ddp_b200 graph ops run on CUDA only (no CPU fallback)')
    x = x.float().contiguous()
    n = x.shape[0]
    if n == 0:
        return torch.zeros(2, 0, dtype=torch.long, device=dev)
    nb = _num_examples(batch, batch)
    p = _ptr_from_batch(batch, n, nb, dev)
    slab = torch.empty(max(n, 1) * (k + 1), dtype=torch.int32, device=dev)
    counts = torch.zeros(n + 1, dtype=torch.int32, device=dev)
    cap = max(n * (k + 1), 1)
    edge = torch.empty(2 * cap, dtype=torch.int32, device=dev)
    n_dev = torch.zeros(1, dtype=torch.int32, device=dev)
    _lib.check(_lib.lib().ddp_knn_graph(ptr(x), ptr(p), nb, n, int(k), ptr(slab), k + 1, ptr(counts), ptr(edge), cap, ptr(n_dev),
                                        _lib.stream_ptr()), 'ddp_knn_graph')
    m = int(n_dev.item())
    return torch.stack([edge[:m], edge[cap:cap + m]]).long()


def calpha_graph(pos, receptor_radius=15.0, max_neighbors=24, batch=None):
    """Residue contact graph of the reference's preprocessing (datasets/process_mols.py:661-677) on the device:
    ``edge_index[0]`` = residue (repeated), ``edge_index[1]`` = its neighbours -- every residue within
    ``receptor_radius`` in index order, the ``max_neighbors`` nearest (ascending distance) when there are more, the
    single nearest when there are none.  ``batch`` separates complexes as in radius_graph."""
    dev = pos.device
    if dev.type != 'cuda':
        raise RuntimeError('ddp_b200 graph ops run on CUDA only (no CPU fallback)')
    x = pos.float().contiguous()
    n = x.shape[0]
    if n == 0:
        return torch.zeros(2, 0, dtype=torch.long, device=dev)
    nb = _num_examples(batch, batch)
    p = _ptr_from_batch(batch, n, nb, dev)
    k = int(max_neighbors)
    slab = torch.empty(n * k, dtype=torch.int32, device=dev)
    counts = torch.zeros(n + 1, dtype=torch.int32, device=dev)
    cap = n * k
    edge = torch.empty(2 * cap, dtype=torch.int32, device=dev)
    n_dev = torch.zeros(1, dtype=torch.int32, device=dev)
    _lib.check(_lib.lib().ddp_calpha_graph(ptr(x), ptr(p), nb, n, float(receptor_radius), k, ptr(slab), k, ptr(counts), ptr(edge), cap,
                                           ptr(n_dev), _lib.stream_ptr()), 'ddp_calpha_graph')
    m = int(n_dev.item())
    return torch.stack([edge[:m], edge[cap:cap + m]]).long()
