"""PyG 2.4.0 data surface for the oracle: ``HeteroData`` / ``Batch`` / ``DataLoader`` (TEST INFRASTRUCTURE).

torch_geometric is an un-vendored dependency of the reference (environment.yml:15, ``pyg=2.4.0``) and is
not installed here; this file restates the part of its published behaviour the hot path relies on
 and is deliberately independent of the product's ``hetero.py`` so that a
collate bug in either shows up as a parity failure:

* ``HeteroData``: ``data['ligand']`` / ``data['ligand', 'ligand']`` create-on-access stores; a 2-tuple
  resolves to the unique ``(src, rel, dst)`` edge type or to rel ``'to'``; ``len(store)`` = number of
  attributes; ``store.num_nodes`` explicit or inferred from ``x`` / ``pos`` / ``batch``.
* ``Batch.from_data_list`` (torch_geometric/data/collate.py): tensors are concatenated along
  ``__cat_dim__`` (= -1 for keys containing ``index`` or ``face``, else 0) after adding ``__inc__``
  (= cumulative ``num_nodes`` of the source / destination node type for ``*index*`` keys of an edge
  store, else 0); 0-dim tensors and python numbers are stacked; everything else becomes a list; every
  node store gets ``batch`` and ``ptr``; ``num_graphs`` is set on the batch.
* ``DataLoader(data_list, batch_size)``: torch ``DataLoader`` with PyG's ``Collater``, ``shuffle=False``.
  Creating its iterator draws one int64 from the default CPU generator (the ``_base_seed`` of
  ``torch.utils.data.dataloader._BaseDataLoaderIter.__init__``), which is part of the random stream a
  seeded ``sampling()`` run consumes (utils/sampling.py:100 builds a loader every step).
"""
import copy

import torch


class _Store(dict):
    """One node- or edge-type storage: attribute access over a dict."""

    def __getattr__(self, name):
        if name.startswith('__'):
            raise AttributeError(name)
        if name in self:
            return self[name]
        if name == 'num_nodes':
            for k in ('x', 'pos', 'batch'):
                if k in self and torch.is_tensor(self[k]):
                    return int(self[k].shape[0])
        raise AttributeError(name)

    def __setattr__(self, name, value):
        self[name] = value

    def __delattr__(self, name):
        del self[name]

    def __deepcopy__(self, memo):
        out = type(self)()
        for k, v in self.items():
            dict.__setitem__(out, k, copy.deepcopy(v, memo))
        return out


def _is_edge_key(key):
    return isinstance(key, tuple)


class HeteroData:
    def __init__(self):
        self.__dict__['_stores'] = {}          # insertion ordered: node types (str) and edge types (3-tuples)
        self.__dict__['_attrs'] = {}           # graph-level attributes

    # ------------------------------------------------------------------ keys
    def _resolve(self, key):
        if not _is_edge_key(key):
            return key
        if len(key) == 3:
            return tuple(key)
        src, dst = key
        cands = [k for k in self._stores if _is_edge_key(k) and k[0] == src and k[2] == dst]
        if len(cands) > 1:
            raise KeyError(f'edge type {key} is ambiguous: {cands}')
        return cands[0] if cands else (src, 'to', dst)

    def __getitem__(self, key):
        if not _is_edge_key(key) and key in self._attrs:
            return self._attrs[key]
        key = self._resolve(key)
        if key not in self._stores:
            self._stores[key] = _Store()
        return self._stores[key]

    def __setitem__(self, key, value):
        if isinstance(value, _Store):
            self._stores[self._resolve(key)] = value
        else:
            self._attrs[key] = value

    def __delitem__(self, key):
        key = self._resolve(key)
        if key in self._stores:
            del self._stores[key]
        else:
            del self._attrs[key]

    def __getattr__(self, name):
        if name.startswith('__'):
            raise AttributeError(name)
        attrs = self.__dict__['_attrs']
        if name in attrs:
            return attrs[name]
        raise AttributeError(name)

    def __setattr__(self, name, value):
        self._attrs[name] = value

    def __delattr__(self, name):
        del self._attrs[name]

    @property
    def node_types(self):
        return [k for k in self._stores if not _is_edge_key(k)]

    @property
    def edge_types(self):
        return [k for k in self._stores if _is_edge_key(k)]

    def __deepcopy__(self, memo):
        out = type(self)()
        for k, v in self._stores.items():
            out._stores[k] = copy.deepcopy(v, memo)
        for k, v in self._attrs.items():
            out._attrs[k] = copy.deepcopy(v, memo)
        return out

    def to(self, device, *a, **kw):
        def move(v):
            if torch.is_tensor(v):
                return v.to(device)
            if isinstance(v, dict) and not isinstance(v, _Store):
                return {k: move(x) for k, x in v.items()}
            return v
        for st in self._stores.values():
            for k in list(st.keys()):
                dict.__setitem__(st, k, move(st[k]))
        for k in list(self._attrs.keys()):
            self._attrs[k] = move(self._attrs[k])
        return self

    def cpu(self):
        return self.to('cpu')

    def clone(self):
        return copy.deepcopy(self)


def _cat_dim(key):
    return -1 if ('index' in key or 'face' in key) else 0


def _collate_values(key, vals, incs):
    v0 = vals[0]
    if torch.is_tensor(v0):
        if v0.dim() == 0:
            return torch.stack(vals)
        if incs is not None:
            vals = [v + inc.to(v.dtype) for v, inc in zip(vals, incs)]
        return torch.cat(vals, dim=_cat_dim(key))
    if isinstance(v0, (int, float)) and not isinstance(v0, bool):
        return torch.tensor(vals)
    return list(vals)


class Batch(HeteroData):
    @classmethod
    def from_data_list(cls, data_list):
        out = cls()
        data_list = [d if isinstance(d, HeteroData) else from_any(d) for d in data_list]
        n = len(data_list)
        first = data_list[0]
        # cumulative node counts per node type (the __inc__ of edge-level *index* keys)
        cum = {}
        for nt in first.node_types:
            if any(nt not in d._stores for d in data_list):
                continue
            counts = []
            for d in data_list:
                try:
                    counts.append(int(d._stores[nt].num_nodes))
                except AttributeError:
                    counts = None
                    break
            if counts is not None:
                c = [0]
                for k in counts:
                    c.append(c[-1] + k)
                cum[nt] = c
        for key, st0 in first._stores.items():
            if any(key not in d._stores for d in data_list):
                continue
            st = _Store()
            for attr in st0.keys():
                vals = [d._stores[key][attr] for d in data_list]
                if attr == 'num_nodes':
                    dict.__setitem__(st, attr, int(sum(vals)))
                    continue
                incs = None
                if _is_edge_key(key) and torch.is_tensor(vals[0]) and 'index' in attr:
                    src, _, dst = key
                    incs = [torch.tensor([[cum[src][i]], [cum[dst][i]]]) for i in range(n)]
                dict.__setitem__(st, attr, _collate_values(attr, vals, incs))
            if not _is_edge_key(key) and key in cum and len(st0) > 0:
                c = cum[key]
                dict.__setitem__(st, 'batch', torch.repeat_interleave(
                    torch.arange(n), torch.tensor([c[i + 1] - c[i] for i in range(n)])))
                dict.__setitem__(st, 'ptr', torch.tensor(c))
            out._stores[key] = st
        for attr in first._attrs.keys():
            out._attrs[attr] = _collate_values(attr, [d._attrs[attr] for d in data_list], None)
        out._attrs['num_graphs'] = n
        return out


class DataLoader:
    """``torch_geometric.loader.DataLoader(dataset=list, batch_size=B)`` (shuffle=False, num_workers=0)."""

    def __init__(self, dataset, batch_size=1, shuffle=False, **kwargs):
        assert not shuffle, 'the hot path never shuffles (utils/sampling.py:100,266)'
        self.dataset, self.batch_size = dataset, batch_size

    def __len__(self):
        return (len(self.dataset) + self.batch_size - 1) // self.batch_size if self.dataset is not None else 0

    def __iter__(self):
        torch.empty((), dtype=torch.int64).random_()        # _BaseDataLoaderIter._base_seed, drawn when iter() is called
        return self._batches()

    def _batches(self):
        for i in range(0, len(self.dataset), self.batch_size):
            yield Batch.from_data_list(self.dataset[i:i + self.batch_size])


def from_any(g):
    """Rebuild any duck-typed hetero graph (product ``hetero.HeteroData``, real PyG) as a ``pyg_mini.HeteroData``."""
    out = HeteroData()
    for nt in g.node_types:
        for k, v in g[nt].items():
            dict.__setitem__(out[nt], k, copy.deepcopy(v))
    for et in g.edge_types:
        for k, v in g[et].items():
            dict.__setitem__(out[tuple(et)], k, copy.deepcopy(v))
    glob = getattr(g, '_glob', None)
    if glob is None:
        glob = getattr(g, '_attrs', {})
    for k, v in glob.items():
        out._attrs[k] = copy.deepcopy(v)
    return out
