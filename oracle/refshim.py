"""Execute the UNMODIFIED reference modules of the hot path in this container (TEST INFRASTRUCTURE).

The reference cannot be imported as it stands here: its third-party graph stack (e3nn 0.5.1,
torch_cluster 1.6.1, torch_scatter 2.1.0, torch_geometric 2.4.0) and its preprocessing dependencies
(rdkit, Biopython, spyrmsd ...) are not installed.  ``load()`` injects

* the oracle's restatements of the four numerical third-party packages as ``e3nn`` / ``torch_cluster`` /
  ``torch_scatter`` / ``torch_geometric`` (``oracle.e3nn_mini``, ``oracle.cluster``, ``oracle.pyg_mini``,
  wrapped in adapters with the third-party call signatures), and
* inert placeholders for the packages that only the reference's preprocessing / file-writing code touches
  (rdkit, Bio, spyrmsd, esm, prody, openmm ...; nothing on the hot path calls into them)

into ``sys.modules`` and then imports the reference's own files from ``/root/reference`` unchanged:
``models/layers.py``, ``models/score_model.py``, ``models/all_atom_score_model.py``, ``utils/geometry.py``,
``utils/torsion.py``, ``utils/diffusion_utils.py``, ``utils/so3.py``, ``utils/torus.py``, ``utils/sampling.py``,
``utils/utils.py`` (``get_model``).  What this pins: every line of the REFERENCE'S OWN code on the path
(model assembly, conv layer, FasterTensorProduct, heads, sampler arithmetic, pose updates, Kabsch, the
so3 / torus tables) against ``oracle/*_ref.py``.  What it cannot pin: the third-party arithmetic itself,
which stays a restatement (``oracle/e3nn_mini.py``, ``oracle/cluster_c.c``) checked by algebraic identities.

``utils/so3.py`` and ``utils/torus.py`` compute their tables at import time (6 + 2.5 minutes here) and
cache them as ``.npy`` in the CWD; ``load()`` imports them with the CWD set to ``CACHE`` (``.git/ddp_ref_cache``: outside
history and outside the gpurun snapshot) so the reference's own caching applies.  numpy's global RNG is seeded (``torus_seed``) around the
import of ``utils/torus.py`` because its ``score_norm_`` table is a Monte-Carlo estimate from the unseeded global
RNG.

Only ``tests/`` and ``scripts/make_ref_fixtures.py`` use this module, and only where ``/root/reference`` exists.
"""
import collections
import contextlib
import importlib
import io
import os
import sys
import types
from unittest import mock

import numpy as np
import torch

from . import cluster, e3nn_mini, pyg_mini

REF_ROOT = os.environ.get('DDP_REFERENCE_ROOT', '/root/reference')
_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# 430 MB of .npy tables the reference caches in its CWD: kept under .git/ (never part of a commit, and outside the snapshot
# that travels to the GPU box) when the checkout has one, else under oracle/_ref/cache (git-ignored)
CACHE = os.environ.get('DDP_REF_CACHE') or (os.path.join(_ROOT, '.git', 'ddp_ref_cache') if os.path.isdir(os.path.join(_ROOT, '.git'))
                                            else os.path.join(_ROOT, 'oracle', '_ref', 'cache'))
_LOADED = None


def available():
    return os.path.isdir(os.path.join(REF_ROOT, 'models'))


# ----------------------------------------------------------------------------------- e3nn adapter
class Irrep(collections.namedtuple('Irrep', ['l', 'p'])):
    """e3nn.o3.Irrep: tuple (l, p), ``str`` -> '1o', ``dim`` -> 2l+1."""

    def __str__(self):
        return f"{self.l}{'e' if self.p == 1 else 'o'}"

    __repr__ = __str__

    @property
    def dim(self):
        return 2 * self.l + 1


class Irreps:
    """e3nn.o3.Irreps surface the reference uses: iteration over (mul, Irrep), ``slices``, ``dim``, ``==``."""

    def __init__(self, spec=None):
        if isinstance(spec, Irreps):
            self._m = e3nn_mini.Irreps(spec._m)
        elif isinstance(spec, e3nn_mini.Irreps):
            self._m = e3nn_mini.Irreps(spec)
        else:
            self._m = e3nn_mini.Irreps(spec if spec is not None else '')

    @staticmethod
    def spherical_harmonics(lmax, p=-1):
        return Irreps(e3nn_mini.Irreps.spherical_harmonics(lmax))

    def __iter__(self):
        return iter([(mul, Irrep(l, p)) for mul, l, p in self._m])

    def __len__(self):
        return len(self._m)

    def __getitem__(self, i):
        mul, l, p = self._m[i]
        return (mul, Irrep(l, p))

    def __eq__(self, other):
        return self._m == Irreps(other)._m

    def __hash__(self):
        return hash(repr(self._m))

    @property
    def dim(self):
        return self._m.dim

    def slices(self):
        return self._m.slices()

    def __repr__(self):
        return repr(self._m)


def _mini(ir):
    return ir._m if isinstance(ir, Irreps) else e3nn_mini.Irreps(ir)


class _FCTP(e3nn_mini.FullyConnectedTensorProduct):
    def __init__(self, irreps_in1, irreps_in2, irreps_out, shared_weights=True, internal_weights=None, **kw):
        assert shared_weights is False, 'the reference only builds per-edge-weight products (models/score_model.py:98)'
        super().__init__(_mini(irreps_in1), _mini(irreps_in2), _mini(irreps_out))


class _FullTP(e3nn_mini.FullTensorProduct):
    def __init__(self, irreps_in1, irreps_in2, **kw):
        super().__init__(_mini(irreps_in1), _mini(irreps_in2))
        self.irreps_out = Irreps(self.irreps_out)


class _BatchNorm(e3nn_mini.BatchNorm):
    def __init__(self, irreps, eps=1e-5, **kw):
        super().__init__(_mini(irreps), eps=eps)


def _spherical_harmonics(l, x, normalize, normalization='integral'):
    return e3nn_mini.spherical_harmonics(_mini(l) if isinstance(l, Irreps) else l, x, normalize=normalize,
                                         normalization=normalization)


def _e3nn_modules():
    e3nn = types.ModuleType('e3nn')
    o3 = types.ModuleType('e3nn.o3')
    nn = types.ModuleType('e3nn.nn')
    o3.Irreps, o3.Irrep = Irreps, Irrep
    o3.spherical_harmonics = _spherical_harmonics
    o3.FullyConnectedTensorProduct = _FCTP
    o3.FullTensorProduct = _FullTP
    o3.wigner_3j = e3nn_mini.wigner_3j
    nn.BatchNorm = _BatchNorm
    e3nn.o3, e3nn.nn = o3, nn
    return {'e3nn': e3nn, 'e3nn.o3': o3, 'e3nn.nn': nn}


# ----------------------------------------------------------------------------------- torch_cluster / torch_scatter
def _cluster_modules():
    tc = types.ModuleType('torch_cluster')
    tc.radius, tc.radius_graph, tc.knn_graph, tc.knn = cluster.radius, cluster.radius_graph, cluster.knn_graph, cluster.knn
    ts = types.ModuleType('torch_scatter')

    def scatter(src, index, dim=-1, out=None, dim_size=None, reduce='sum'):
        assert out is None
        return cluster.scatter(src, index, dim if dim >= 0 else src.dim() + dim, dim_size, reduce)

    def scatter_mean(src, index, dim=-1, out=None, dim_size=None):
        return scatter(src, index, dim, out, dim_size, 'mean')

    ts.scatter, ts.scatter_mean = scatter, scatter_mean
    return {'torch_cluster': tc, 'torch_scatter': ts}


# ----------------------------------------------------------------------------------- torch_geometric
def _pyg_modules():
    names = ['torch_geometric', 'torch_geometric.data', 'torch_geometric.loader', 'torch_geometric.utils',
             'torch_geometric.nn', 'torch_geometric.nn.data_parallel', 'torch_geometric.transforms',
             'torch_geometric.loader.dataloader', 'torch_geometric.data.dataset']
    mods = {n: types.ModuleType(n) for n in names}
    d = mods['torch_geometric.data']
    d.HeteroData, d.Batch, d.Data, d.Dataset = pyg_mini.HeteroData, pyg_mini.Batch, pyg_mini.HeteroData, object
    mods['torch_geometric.data.dataset'].Dataset = object
    ld = mods['torch_geometric.loader']
    ld.DataLoader, ld.DataListLoader = pyg_mini.DataLoader, pyg_mini.DataLoader
    mods['torch_geometric.loader.dataloader'].Collater = object
    u = mods['torch_geometric.utils']
    u.to_networkx = u.subgraph = u.to_dense_adj = u.dense_to_sparse = u.unbatch = _never_called('torch_geometric.utils')
    u.degree = lambda index, num_nodes=None, dtype=None: torch.bincount(index, minlength=num_nodes or 0)
    mods['torch_geometric.nn.data_parallel'].DataParallel = _never_called('DataParallel')
    mods['torch_geometric.transforms'].BaseTransform = object
    tg = mods['torch_geometric']
    tg.data, tg.loader, tg.utils, tg.nn, tg.transforms = d, ld, u, mods['torch_geometric.nn'], mods['torch_geometric.transforms']
    mods['torch_geometric.nn'].data_parallel = mods['torch_geometric.nn.data_parallel']
    return mods


def _never_called(what):
    def f(*a, **k):
        raise RuntimeError(f'{what} is outside the hot path and is not shimmed')
    return f


# ----------------------------------------------------------------------------------- inert placeholders
_INERT = ['rdkit', 'rdkit.Chem', 'rdkit.Chem.rdchem', 'rdkit.Chem.AllChem', 'rdkit.Geometry', 'rdkit.Chem.rdMolTransforms',
          'rdkit.Chem.rdMolAlign', 'rdkit.Chem.rdmolops', 'rdkit.RDLogger',
          'Bio', 'Bio.PDB', 'Bio.PDB.PDBExceptions', 'Bio.PDB.Selection', 'Bio.PDB.Polypeptide', 'Bio.SeqRecord', 'Bio.Seq',
          'Bio.PDB.PDBIO', 'spyrmsd', 'spyrmsd.rmsd', 'spyrmsd.molecule', 'esm', 'prody', 'openmm', 'openmm.app',
          'openmm.unit', 'pdbfixer', 'wandb', 'lightning', 'posebusters']


def _inert_modules():
    out = {}
    for n in _INERT:
        try:
            importlib.import_module(n)            # present in this environment: leave it alone
        except Exception:
            m = mock.MagicMock(name=n)
            m.__path__ = []                       # "is a package" for the import machinery
            m.__spec__ = None
            out[n] = m
    return out


# ----------------------------------------------------------------------------------- load
Ref = collections.namedtuple('Ref', ['layers', 'score_model', 'all_atom', 'geometry', 'torsion', 'diffusion_utils',
                                     'so3', 'torus', 'sampling', 'utils'])


def load(torus_seed=0, quiet=True):
    """Import the reference modules (once per process) and return them."""
    global _LOADED
    if _LOADED is not None:
        return _LOADED
    if not available():
        raise FileNotFoundError(f'{REF_ROOT} is not present: reference-pinning needs the reference tree')
    for shim in (_e3nn_modules(), _cluster_modules(), _pyg_modules(), _inert_modules()):
        for k, v in shim.items():
            sys.modules[k] = v
    os.makedirs(CACHE, exist_ok=True)
    for k in list(sys.modules):                   # the reference's top-level package names are generic
        if k.split('.')[0] in ('models', 'utils', 'datasets') and not getattr(sys.modules[k], '__file__', REF_ROOT).startswith(REF_ROOT):
            raise RuntimeError(f'module {k} would shadow the reference package of that name')
    for pkg in ('models', 'utils', 'datasets'):   # namespace packages of the reference (an installed 'datasets' must not win)
        m = types.ModuleType(pkg)
        m.__path__ = [os.path.join(REF_ROOT, pkg)]
        sys.modules[pkg] = m
    cwd, rng_state = os.getcwd(), np.random.get_state()
    sys.path.insert(0, REF_ROOT)
    sink = io.StringIO()
    try:
        os.chdir(CACHE)
        with (contextlib.redirect_stderr(sink) if quiet else contextlib.nullcontext()), np.errstate(all='ignore'):
            so3 = importlib.import_module('utils.so3')
            np.random.seed(torus_seed)
            torus = importlib.import_module('utils.torus')
            torus.score_norm_seed0_ = torus.score_norm_.copy()        # kept aside: parity runs overwrite score_norm_ (F8)
            mods = [importlib.import_module(n) for n in
                    ('models.layers', 'models.score_model', 'models.all_atom_score_model', 'utils.geometry', 'utils.torsion',
                     'utils.diffusion_utils')]
            sampling = importlib.import_module('utils.sampling')
            utils = importlib.import_module('utils.utils')
    finally:
        os.chdir(cwd)
        np.random.set_state(rng_state)
        sys.path.remove(REF_ROOT)
    for m in mods + [so3, torus, sampling, utils]:
        assert m.__file__.startswith(REF_ROOT), m.__file__
    _LOADED = Ref(*mods, so3, torus, sampling, utils)
    return _LOADED
