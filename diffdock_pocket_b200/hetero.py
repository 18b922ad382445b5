"""Duck-typed stand-in for ``torch_geometric.data.HeteroData`` / ``Batch`` / ``DataLoader``.

The reference's data surface is PyG 2.4.0 ``HeteroData`` produced by ``PDBBind`` and collated by
``torch_geometric.loader.DataLoader`` (utils/sampling.py:100, inference.py:135).  PyG is not a
dependency of this package; this module keeps the field names and collate rules the hot path
relies on so ``model.forward(data)`` and ``sampling()`` can be
called the same way.  Real PyG objects are accepted wherever these are (same ``data[key].attr``
access pattern).
"""
import copy
from typing import Any, Dict, List

import numpy as np
import torch


class Store:
    """Attribute bag for one node type or edge type."""

    def __init__(self):
        object.__setattr__(self, '_d', {})

    def __getattr__(self, k):
        d = object.__getattribute__(self, '_d')
        if k == 'num_nodes' and 'num_nodes' not in d:
            for key in ('x', 'pos', 'batch'):
                if key in d and torch.is_tensor(d[key]):
                    return d[key].shape[0]
            raise AttributeError(k)
        if k in d:
            return d[k]
        raise AttributeError(k)

    def __setattr__(self, k, v):
        self._d[k] = v

    def __delattr__(self, k):
        del self._d[k]

    def __contains__(self, k):
        return k in self._d

    def __len__(self):
        return len(self._d)

    def keys(self):
        return self._d.keys()

    def items(self):
        return self._d.items()

    def __deepcopy__(self, memo):
        s = Store()
        for k, v in self._d.items():
            s._d[k] = copy.deepcopy(v, memo)
        return s


class HeteroData:
    def __init__(self):
        object.__setattr__(self, '_nodes', {})
        object.__setattr__(self, '_edges', {})
        object.__setattr__(self, '_glob', {})

    # -- store access -----------------------------------------------------------------------
    def _edge_key(self, key):
        if len(key) == 3:
            return key
        src, dst = key
        hits = [k for k in self._edges if k[0] == src and k[2] == dst]
        if len(hits) == 1:
            return hits[0]
        if len(hits) == 0:
            return (src, 'to', dst)
        raise KeyError(f'ambiguous edge type {key}')

    def __getitem__(self, key):
        if isinstance(key, tuple):
            k = self._edge_key(key)
            if k not in self._edges:
                self._edges[k] = Store()
            return self._edges[k]
        if key in self._glob:
            return self._glob[key]
        if key not in self._nodes:
            self._nodes[key] = Store()
        return self._nodes[key]

    def __setitem__(self, key, value):
        if isinstance(value, Store):
            if isinstance(key, tuple):
                self._edges[self._edge_key(key)] = value
            else:
                self._nodes[key] = value
        else:
            self._glob[key] = value

    def __delitem__(self, key):
        if isinstance(key, tuple):
            del self._edges[self._edge_key(key)]
        elif key in self._nodes:
            del self._nodes[key]
        else:
            del self._glob[key]

    def __contains__(self, key):
        return key in self._nodes or key in self._glob or (isinstance(key, tuple) and self._edge_key(key) in self._edges)

    def __getattr__(self, k):
        g = object.__getattribute__(self, '_glob')
        if k in g:
            return g[k]
        raise AttributeError(k)

    def __setattr__(self, k, v):
        self._glob[k] = v

    @property
    def node_types(self):
        return list(self._nodes.keys())

    @property
    def edge_types(self):
        return list(self._edges.keys())

    def __deepcopy__(self, memo):
        d = HeteroData()
        for k, v in self._nodes.items():
            d._nodes[k] = copy.deepcopy(v, memo)
        for k, v in self._edges.items():
            d._edges[k] = copy.deepcopy(v, memo)
        for k, v in self._glob.items():
            d._glob[k] = copy.deepcopy(v, memo)
        return d

    def to(self, device):
        def mv(v):
            if torch.is_tensor(v):
                return v.to(device)
            if isinstance(v, dict):
                return {a: mv(b) for a, b in v.items()}
            return v
        for st in list(self._nodes.values()) + list(self._edges.values()):
            for k in list(st.keys()):
                st._d[k] = mv(st._d[k])
        for k in list(self._glob.keys()):
            self._glob[k] = mv(self._glob[k])
        return self


class Batch(HeteroData):
    """PyG collate: cat along dim 0; keys containing 'index'/'face' cat along
    the last dim and are incremented by the cumulative node count; per node type ``batch``/``ptr``."""

    @staticmethod
    def from_data_list(data_list: List[HeteroData], skip=()) -> 'Batch':
        """``skip``: (node type, attribute) pairs that are NOT concatenated (the sampler passes the static feature
        matrices -- 1281 floats per residue -- to the model per complex instead of per sample)."""
        b = Batch()
        n = len(data_list)
        b._glob['num_graphs'] = n
        node_types = []
        for d in data_list:
            for t in d.node_types:
                if t not in node_types:
                    node_types.append(t)
        offsets: Dict[str, List[int]] = {}
        for t in node_types:
            has = [t in d._nodes and len(d._nodes[t]) > 0 for d in data_list]
            if not any(has):
                continue                                   # absent, or an empty store left behind by data[t] look-ups
            if not all(has):
                # PyG cannot collate such a list.  The only optional node type of the path is 'flexResidues' (complexes
                # without flexible residues in a cross-complex batch): those graphs contribute zero entries.  Anything
                # else is a malformed input and is reported instead of being dropped silently.
                if t != 'flexResidues':
                    missing = [i for i, h in enumerate(has) if not h]
                    raise ValueError(f"node type {t!r} is missing or empty in graphs {missing} of the batch")
            st = Store()
            present = [d for d, h in zip(data_list, has) if h]
            counts = [d._nodes[t].num_nodes if h else 0 for d, h in zip(data_list, has)]
            off = np.concatenate([[0], np.cumsum(counts)])
            offsets[t] = off
            for k in present[0]._nodes[t].keys():
                if (t, k) in skip:
                    continue
                vals = [d._nodes[t]._d[k] for d in present]
                if k == 'num_nodes':
                    st._d[k] = int(sum(vals))
                elif torch.is_tensor(vals[0]):
                    st._d[k] = torch.cat(vals, 0) if vals[0].dim() > 0 else torch.stack(vals)
                else:
                    st._d[k] = vals
            st._d['batch'] = torch.repeat_interleave(torch.arange(n), torch.as_tensor(counts))
            st._d['ptr'] = torch.as_tensor(off, dtype=torch.long)
            b._nodes[t] = st
        for et in data_list[0].edge_types:
            st = Store()
            for k in data_list[0]._edges[et].keys():
                vals = [d._edges[et]._d[k] for d in data_list]
                if torch.is_tensor(vals[0]) and ('index' in k or 'face' in k):
                    inc = torch.stack([torch.as_tensor(offsets[et[0]][:-1]), torch.as_tensor(offsets[et[2]][:-1])], 1)
                    st._d[k] = torch.cat([v + inc[i].reshape(2, 1).to(v.dtype) for i, v in enumerate(vals)], -1)
                elif torch.is_tensor(vals[0]):
                    st._d[k] = torch.cat(vals, 0)
                else:
                    st._d[k] = vals
            b._edges[et] = st
        for k in data_list[0]._glob.keys():
            vals = [d._glob[k] for d in data_list]
            if torch.is_tensor(vals[0]):
                b._glob[k] = torch.cat(vals, 0) if vals[0].dim() > 0 else torch.stack(vals)
            else:
                b._glob[k] = vals
        return b


class DataLoader:
    """``torch_geometric.loader.DataLoader(data_list, batch_size)`` without shuffling."""

    def __init__(self, data_list, batch_size=1, shuffle=False):
        assert not shuffle
        self.data_list, self.batch_size = data_list, batch_size

    def __iter__(self):
        for i in range(0, len(self.data_list), self.batch_size):
            yield Batch.from_data_list(self.data_list[i:i + self.batch_size])

    def __len__(self):
        return (len(self.data_list) + self.batch_size - 1) // self.batch_size


def sample_copies(graph, n):
    """``n`` graphs for ``n`` samples of one complex: everything is shared with ``graph`` (same tensor objects) except the
    coordinates the sampler moves (``ligand.pos``, ``atom.pos``), which are cloned.  The reference deep-copies the whole
    graph per sample (inference.py:135), ESM features included; sharing is what lets ``sampling()`` recognise samples of
    one complex and upload / embed its static part once."""
    out = []
    for _ in range(n):
        g = HeteroData()
        for k, st in graph._nodes.items():
            ns = Store()
            ns._d.update(st._d)
            if k in ('ligand', 'atom') and 'pos' in st._d:
                ns._d['pos'] = st._d['pos'].clone()
            g._nodes[k] = ns
        for k, st in graph._edges.items():
            es = Store()
            es._d.update(st._d)
            g._edges[k] = es
        g._glob.update(graph._glob)
        out.append(g)
    return out
