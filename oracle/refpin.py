"""Shared definitions of the reference-pinning cases (TEST INFRASTRUCTURE).

``scripts/make_ref_fixtures.py`` runs every case below through the UNMODIFIED reference modules
(``oracle.refshim.load()``) and stores the outputs under ``tests/golden/ref_*.npz``;
``tests/test_reference_pin.py`` re-creates the same seeded inputs and checks the oracle (CPU, everywhere) and
``tests/test_gpu_refpin.py`` the CUDA path (GPU box) against those stored reference outputs.  Inputs are
drawn from ``numpy.random.RandomState`` (bit-stable across platforms), never from torch's generator, wherever the
values themselves are not stored in the fixture.
"""
import copy
from argparse import Namespace

import numpy as np
import torch

SEQ = ['60x0e', '60x0e + 10x1o', '60x0e + 10x1o + 10x1e', '60x0e + 10x1o + 10x1e + 60x0o']
FTP_CASES = [(SEQ[0], SEQ[1]), (SEQ[1], SEQ[2]), (SEQ[2], SEQ[3]), (SEQ[3], SEQ[3]), (SEQ[3], '2x1o + 2x1e'),
             ('24x0e + 6x1o + 6x1e + 24x0o', '24x0e + 6x1o + 6x1e + 24x0o')]
# name -> (in_irreps, out_irreps, n_edge_features, faster, sh_irreps)
CONV_CASES = {
    'l0': (SEQ[0], SEQ[1], 180, True, '1x0e+1x1o'), 'l1': (SEQ[1], SEQ[2], 180, True, '1x0e+1x1o'),
    'l2': (SEQ[2], SEQ[3], 180, True, '1x0e+1x1o'), 'l3': (SEQ[3], SEQ[3], 180, True, '1x0e+1x1o'),
    'final': (SEQ[3], '2x1o + 2x1e', 120, True, '1x0e+1x1o'),
    'lmax2': (SEQ[3], SEQ[3], 180, False, '1x0e+1x1o+1x2e'),
}
TEMPS = dict(temp_sampling=[0.9766350103728372, 6.077432837220868, 6.761568162335063, 1.4487910576602347],
             temp_psi=[1.5102572175711826, 0.8141168207563049, 0.7661845361370018, 1.339614553802453],
             temp_sigma_data=0.48884149503636976)                     # inference.py:93-101 defaults


def np_fill(module, seed):
    """Deterministic, platform-independent parameters for a torch module (reference, oracle or product class)."""
    rng = np.random.RandomState(seed)
    with torch.no_grad():
        for name, t in sorted(module.state_dict().items()):
            if not t.dtype.is_floating_point:
                continue
            shape = tuple(t.shape)
            if name.endswith('running_var'):
                v = rng.uniform(0.5, 1.5, shape)
            elif name.endswith('running_mean'):
                v = rng.standard_normal(shape) * 0.2
            elif 'batch_norm' in name and name.endswith('weight'):
                v = rng.uniform(0.7, 1.3, shape)
            elif t.dim() >= 2:
                v = rng.standard_normal(shape) / np.sqrt(shape[-1])
            else:
                v = rng.standard_normal(shape) * 0.1
            t.copy_(torch.from_numpy(np.asarray(v, dtype=np.float32)))
    return module


def weight_checksum(state_dict):
    s = np.zeros(3)
    for k in sorted(state_dict):
        v = state_dict[k].detach().double().cpu().numpy().ravel()
        s += [v.sum(), np.abs(v).sum(), (v * v).sum()]
    return s


def conv_inputs(case, n=50, e=333, seed=1):
    """Seeded operator inputs of one CONV_CASES entry: x, edge_index, edge_attr, edge vectors."""
    from . import e3nn_mini as E
    in_ir, out_ir, nf, faster, sh_ir = CONV_CASES[case]
    rng = np.random.RandomState(seed)
    x = torch.from_numpy(rng.standard_normal((n, E.Irreps(in_ir).dim)).astype(np.float32))
    ei = torch.from_numpy(rng.randint(0, n, (2, e)).astype(np.int64))
    ei[0, :40] = 7                                                    # a hub node; nodes without edges exist too
    ea = torch.from_numpy(rng.standard_normal((e, nf)).astype(np.float32))
    vec = torch.from_numpy(rng.standard_normal((e, 3)).astype(np.float32))
    sh = E.spherical_harmonics(sh_ir, vec)
    return x, ei, ea, sh


def ftp_inputs(i, e=3):
    from . import e3nn_mini as E
    in_ir, out_ir = FTP_CASES[i]
    rng = np.random.RandomState(100 + i)
    x = torch.from_numpy(rng.standard_normal((e, E.Irreps(in_ir).dim)).astype(np.float32))
    sh = E.spherical_harmonics('1x0e+1x1o', torch.from_numpy(rng.standard_normal((e, 3)).astype(np.float32)))
    return x, sh, rng


def pose_inputs(g, n=3, seed=0):
    """Per-sample perturbations for the pose-update pin (tr, rot axis-angle, ligand torsions, side-chain torsions)."""
    rng = np.random.RandomState(seed)
    n_tor = int(g['ligand'].edge_mask.sum())
    n_sc = int(g['flexResidues'].edge_idx.shape[0])
    tr = rng.standard_normal((n, 3)).astype(np.float32)
    rot = (rng.standard_normal((n, 3)) * 0.4).astype(np.float32)
    tor = (rng.standard_normal((n, n_tor)) * 0.5).astype(np.float32)
    sc = (rng.standard_normal((n, n_sc)) * 0.5).astype(np.float32)
    tor[1, 2] = 0.0                                                   # utils/torsion.py:76 skips exact zeros
    return tr, rot, tor, sc


class Capture:
    """Forward pre-hooks on a score model with the reference's module attribute names: records the edge lists the
    first-layer convs receive and the node features entering every interaction layer / head, without touching
    ``forward`` (models/all_atom_score_model.py:238-436)."""

    def __init__(self, model):
        self.model, self.h, self.rec = model, [], {}
        L = model.num_conv_layers
        cl = model.conv_layers

        def grab(key, what):
            def hook(mod, args, kwargs=None):
                self.rec[key] = args[what].detach().clone()
            return hook
        for nm, k in (('ll', 0), ('lr', 1), ('la', 2), ('aa', 3)):
            self.h.append(cl[k].register_forward_pre_hook(grab(nm, 1)))
        for l in range(1, L):
            self.h.append(cl[9 * l].register_forward_pre_hook(grab(('lig', l - 1), 0)))
            self.h.append(cl[9 * l + 3].register_forward_pre_hook(grab(('atom', l - 1), 0)))
            self.h.append(cl[9 * l + 1].register_forward_pre_hook(grab(('rec', l - 1), 0)))
        if hasattr(model, 'final_conv') and not model.confidence_mode:
            self.h.append(model.final_conv.register_forward_pre_hook(grab(('lig', L - 1), 0)))
            if getattr(model, 'flexible_sidechains', False) and hasattr(model, 'sc_tor_bond_conv'):
                self.h.append(model.sc_tor_bond_conv.register_forward_pre_hook(grab(('atom', L - 1), 0)))

    def layers(self):
        L = self.model.num_conv_layers
        out = []
        for l in range(L):
            out.append(tuple(self.rec.get((nt, l)) for nt in ('lig', 'atom', 'rec')))
        return out

    def close(self):
        for h in self.h:
            h.remove()


class Recorder(torch.nn.Module):
    """Wraps a model for ``sampling()``: records every call's outputs (per-step scores) without changing them."""

    def __init__(self, model):
        super().__init__()
        self.model, self.calls = model, []

    def forward(self, data):
        out = self.model(data)
        self.calls.append(tuple(o.detach().clone() for o in out) if isinstance(out, tuple) else out.detach().clone())
        return out


def to_namespace(ns):
    return Namespace(**copy.deepcopy(vars(ns)))
