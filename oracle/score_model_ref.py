"""CPU (PyTorch FP32) restatement of the reference all-atom score / confidence model.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Follows, with the same module attribute names so
that state dicts are interchangeable with the reference's checkpoints:

* ``AtomEncoder`` / ``OldAtomEncoder``   <- models/score_model.py:54-82 / 17-52
* ``GaussianSmearing``                  <- models/score_model.py:661-671
* ``FasterTensorProduct``               <- models/layers.py:8-85
* ``TensorProductConvLayer``            <- models/score_model.py:84-125
* ``TensorProductScoreModel``           <- models/all_atom_score_model.py:21-652
  (``parallel == 1``; affinity / misc-atom branches are outside the hot path)

Third-party pieces come from oracle.e3nn_mini (e3nn 0.5.1) and oracle.cluster
(pytorch-cluster 1.6.1 CUDA semantics, pytorch-scatter 2.1.0).
"""
import math

import numpy as np
import torch
from torch import nn
from torch.nn import functional as F

from . import cluster
from .e3nn_mini import (BatchNorm, FullTensorProduct, FullyConnectedTensorProduct, Irreps, ir_name,
                        spherical_harmonics)

# datasets/process_mols.py:69-97
LIG_FEATURE_DIMS = ([119, 4, 12, 12, 8, 10, 6, 6, 2, 8, 2, 2, 2, 2, 2, 2], 0)
REC_ATOM_FEATURE_DIMS = ([38, 119, 23, 38], 0)
REC_RESIDUE_FEATURE_DIMS = ([38], 0)


class AtomEncoder(nn.Module):
    """models/score_model.py:54-82."""

    def __init__(self, emb_dim, feature_dims, sigma_embed_dim, lm_embedding_type=None):
        super().__init__()
        self.atom_embedding_list = nn.ModuleList()
        self.num_categorical_features = len(feature_dims[0])
        lm_dim = 1280 if lm_embedding_type == 'esm' else 0
        self.additional_features_dim = feature_dims[1] + sigma_embed_dim + lm_dim
        for dim in feature_dims[0]:
            emb = nn.Embedding(dim, emb_dim)
            nn.init.xavier_uniform_(emb.weight.data)
            self.atom_embedding_list.append(emb)
        if self.additional_features_dim > 0:
            self.additional_features_embedder = nn.Linear(self.additional_features_dim + emb_dim, emb_dim)

    def forward(self, x):
        assert x.shape[1] == self.num_categorical_features + self.additional_features_dim
        h = 0
        for i in range(self.num_categorical_features):
            h = h + self.atom_embedding_list[i](x[:, i].long())
        if self.additional_features_dim > 0:
            h = self.additional_features_embedder(torch.cat([h, x[:, self.num_categorical_features:]], 1))
        return h


class OldAtomEncoder(nn.Module):
    """models/score_model.py:17-52."""

    def __init__(self, emb_dim, feature_dims, sigma_embed_dim, lm_embedding_type=None):
        super().__init__()
        self.atom_embedding_list = nn.ModuleList()
        self.num_categorical_features = len(feature_dims[0])
        self.num_scalar_features = feature_dims[1] + sigma_embed_dim
        self.lm_embedding_type = lm_embedding_type
        for dim in feature_dims[0]:
            emb = nn.Embedding(dim, emb_dim)
            nn.init.xavier_uniform_(emb.weight.data)
            self.atom_embedding_list.append(emb)
        if self.num_scalar_features > 0:
            self.linear = nn.Linear(self.num_scalar_features, emb_dim)
        if lm_embedding_type is not None:
            self.lm_embedding_dim = 1280
            self.lm_embedding_layer = nn.Linear(self.lm_embedding_dim + emb_dim, emb_dim)

    def forward(self, x):
        h = 0
        nc, nsf = self.num_categorical_features, self.num_scalar_features
        for i in range(nc):
            h = h + self.atom_embedding_list[i](x[:, i].long())
        if nsf > 0:
            h = h + self.linear(x[:, nc:nc + nsf])
        if self.lm_embedding_type is not None:
            h = self.lm_embedding_layer(torch.cat([h, x[:, -self.lm_embedding_dim:]], 1))
        return h


class GaussianSmearing(nn.Module):
    """models/score_model.py:661-671."""

    def __init__(self, start=0.0, stop=5.0, num_gaussians=50):
        super().__init__()
        offset = torch.linspace(start, stop, num_gaussians)
        self.coeff = -0.5 / (offset[1] - offset[0]).item() ** 2
        self.register_buffer('offset', offset)

    def forward(self, dist):
        d = dist.view(-1, 1) - self.offset.view(1, -1)
        return torch.exp(self.coeff * torch.pow(d, 2))


class FasterTensorProduct(nn.Module):
    """models/layers.py:8-85 -- l<=1 fully connected TP with per-edge [in_k, out_k] weight blocks."""

    def __init__(self, in_irreps, sh_irreps, out_irreps):
        super().__init__()
        assert Irreps(sh_irreps) == Irreps('1x0e+1x1o')
        self.in_irreps, self.out_irreps = Irreps(in_irreps), Irreps(out_irreps)
        im = {'0e': 0, '1o': 0, '1e': 0, '0o': 0}
        om = dict(im)
        for mul, l, p in self.in_irreps:
            im[ir_name(l, p)] = mul
        for mul, l, p in self.out_irreps:
            om[ir_name(l, p)] = mul
        self.weight_shapes = {                                     # layers.py:26-31
            '0e': (im['0e'] + im['1o'], om['0e']),
            '1o': (im['0e'] + im['1o'] + im['1e'], om['1o']),
            '1e': (im['1o'] + im['1e'] + im['0o'], om['1e']),
            '0o': (im['1e'] + im['0o'], om['0o']),
        }
        self.weight_numel = sum(a * b for a, b in self.weight_shapes.values())

    def forward(self, x, sh, weight):
        E = x.shape[0]
        xin = {}
        for (mul, l, p), sl in zip(self.in_irreps, self.in_irreps.slices()):
            v = x[:, sl]
            xin[ir_name(l, p)] = v.reshape(E, mul, 3) if l == 1 else v
        s0, s1 = sh[:, 0], sh[:, 1:]
        basis = {'0e': [], '1o': [], '1e': [], '0o': []}             # layers.py:40-53
        if '0e' in xin:
            basis['0e'].append(xin['0e'] * s0[:, None])
            basis['1o'].append(xin['0e'][:, :, None] * s1[:, None, :])
        if '1o' in xin:
            basis['0e'].append((xin['1o'] * s1[:, None, :]).sum(-1) / np.sqrt(3))
            basis['1o'].append(xin['1o'] * s0[:, None, None])
            basis['1e'].append(torch.linalg.cross(xin['1o'], s1[:, None, :].expand_as(xin['1o']), dim=-1) / np.sqrt(2))
        if '1e' in xin:
            basis['1o'].append(torch.linalg.cross(xin['1e'], s1[:, None, :].expand_as(xin['1e']), dim=-1) / np.sqrt(2))
            basis['1e'].append(xin['1e'] * s0[:, None, None])
            basis['0o'].append((xin['1e'] * s1[:, None, :]).sum(-1) / np.sqrt(3))
        if '0o' in xin:
            basis['1e'].append(xin['0o'][:, :, None] * s1[:, None, :])
            basis['0o'].append(xin['0o'] * s0[:, None])
        wd, start = {}, 0                                            # layers.py:55-61
        for key, (ni, no) in self.weight_shapes.items():
            wd[key] = weight[:, start:start + ni * no].reshape(E, ni, no) / np.sqrt(ni) if ni * no > 0 else None
            start += ni * no
        res = {}
        for key in ('0e', '0o'):
            if basis[key] and wd[key] is not None:
                b = torch.cat(basis[key], -1)
                res[key] = torch.matmul(b.unsqueeze(-2), wd[key]).squeeze(-2)
        for key in ('1o', '1e'):
            if basis[key] and wd[key] is not None:
                b = torch.cat(basis[key], -2)                                       # [E, in_k, 3]
                res[key] = (b.unsqueeze(-2) * wd[key].unsqueeze(-1)).sum(-3).reshape(E, -1)
        return torch.cat([res[ir_name(l, p)] for _, l, p in self.out_irreps], -1)


class TensorProductConvLayer(nn.Module):
    """models/score_model.py:84-125."""

    def __init__(self, in_irreps, sh_irreps, out_irreps, n_edge_features, residual=True, batch_norm=True,
                 dropout=0.0, hidden_features=None, faster=False):
        super().__init__()
        self.in_irreps, self.out_irreps, self.sh_irreps = Irreps(in_irreps), Irreps(out_irreps), Irreps(sh_irreps)
        self.residual = residual
        if hidden_features is None:
            hidden_features = n_edge_features
        if faster:
            self.tp = FasterTensorProduct(in_irreps, sh_irreps, out_irreps)
        else:
            self.tp = FullyConnectedTensorProduct(in_irreps, sh_irreps, out_irreps)
        self.fc = nn.Sequential(nn.Linear(n_edge_features, hidden_features), nn.ReLU(), nn.Dropout(dropout),
                                nn.Linear(hidden_features, self.tp.weight_numel))
        self.batch_norm = BatchNorm(out_irreps) if batch_norm else None

    def forward(self, node_attr, edge_index, edge_attr, edge_sh, out_nodes=None, reduce='mean', edge_weight=1.0):
        if edge_index.numel() == 0:
            return torch.tensor(0, dtype=node_attr.dtype)
        edge_src, edge_dst = edge_index
        tp = self.tp(node_attr[edge_dst], edge_sh, self.fc(edge_attr) * edge_weight)
        out_nodes = out_nodes or node_attr.shape[0]
        out = cluster.scatter(tp, edge_src, dim=0, dim_size=out_nodes, reduce=reduce)
        if self.residual:
            out = out + F.pad(node_attr, (0, out.shape[-1] - node_attr.shape[-1]))
        if self.batch_norm:
            out = self.batch_norm(out)
        return out


def _mlp(i, h, o, dropout):
    return nn.Sequential(nn.Linear(i, h), nn.ReLU(), nn.Dropout(dropout), nn.Linear(h, o))


class TensorProductScoreModel(nn.Module):
    """models/all_atom_score_model.py:21-652 (all-atom, ``parallel == 1``).

    ``so3_score_norm`` / ``torus_score_norm`` are injected callables (numpy in, numpy out) so that
    oracle and product share the same tables."""

    def __init__(self, t_to_sigma, timestep_emb_func, so3_score_norm, torus_score_norm,
                 in_lig_edge_features=4, sigma_embed_dim=32, sh_lmax=2, ns=16, nv=4, num_conv_layers=2,
                 lig_max_radius=5, rec_max_radius=30, cross_max_distance=250, center_max_distance=30,
                 distance_embed_dim=32, cross_distance_embed_dim=32, no_torsion=False, scale_by_sigma=True,
                 use_second_order_repr=False, batch_norm=True, dynamic_max_cross=False, dropout=0.0,
                 smooth_edges=False, lm_embedding_type=None, confidence_mode=False, confidence_dropout=0,
                 confidence_no_batchnorm=False, num_confidence_outputs=1, fixed_center_conv=False,
                 atom_max_neighbors=None, no_aminoacid_identities=False, flexible_sidechains=False,
                 use_old_atom_encoder=False, asyncronous_noise_schedule=False):
        super().__init__()
        self.t_to_sigma, self.timestep_emb_func = t_to_sigma, timestep_emb_func
        # all_atom_score_model.py:370,450,492,517: the sigma embedding reads time 't' instead of 'tr' under the asynchronous schedule
        self.asyncronous_noise_schedule, self._t_key = asyncronous_noise_schedule, 't' if asyncronous_noise_schedule else 'tr'
        self.so3_score_norm, self.torus_score_norm = so3_score_norm, torus_score_norm
        self.in_lig_edge_features = in_lig_edge_features
        self.lig_max_radius, self.rec_max_radius = lig_max_radius, rec_max_radius
        self.cross_max_distance, self.dynamic_max_cross = cross_max_distance, dynamic_max_cross
        self.sh_irreps = Irreps.spherical_harmonics(sh_lmax)
        self.ns, self.nv = ns, nv
        self.scale_by_sigma, self.no_torsion, self.smooth_edges = scale_by_sigma, no_torsion, smooth_edges
        self.num_conv_layers, self.confidence_mode = num_conv_layers, confidence_mode
        self.fixed_center_conv, self.atom_max_neighbors = fixed_center_conv, atom_max_neighbors
        self.no_aminoacid_identities, self.flexible_sidechains = no_aminoacid_identities, flexible_sidechains

        enc = OldAtomEncoder if use_old_atom_encoder else AtomEncoder
        self.lig_node_embedding = enc(ns, LIG_FEATURE_DIMS, sigma_embed_dim)
        self.lig_edge_embedding = _mlp(in_lig_edge_features + sigma_embed_dim + distance_embed_dim, ns, ns, dropout)
        self.rec_node_embedding = enc(ns, REC_RESIDUE_FEATURE_DIMS, sigma_embed_dim, lm_embedding_type)
        self.rec_edge_embedding = _mlp(sigma_embed_dim + distance_embed_dim, ns, ns, dropout)
        self.atom_node_embedding = enc(ns, REC_ATOM_FEATURE_DIMS, sigma_embed_dim)
        self.atom_edge_embedding = _mlp(sigma_embed_dim + distance_embed_dim, ns, ns, dropout)
        self.lr_edge_embedding = _mlp(sigma_embed_dim + cross_distance_embed_dim, ns, ns, dropout)
        self.ar_edge_embedding = _mlp(sigma_embed_dim + distance_embed_dim, ns, ns, dropout)
        self.la_edge_embedding = _mlp(sigma_embed_dim + cross_distance_embed_dim, ns, ns, dropout)
        self.lig_distance_expansion = GaussianSmearing(0.0, lig_max_radius, distance_embed_dim)
        self.rec_distance_expansion = GaussianSmearing(0.0, rec_max_radius, distance_embed_dim)
        self.cross_distance_expansion = GaussianSmearing(0.0, cross_max_distance, cross_distance_embed_dim)

        if use_second_order_repr:                                       # all_atom_score_model.py:87-100
            seq = [f'{ns}x0e', f'{ns}x0e + {nv}x1o + {nv}x2e', f'{ns}x0e + {nv}x1o + {nv}x2e + {nv}x1e + {nv}x2o',
                   f'{ns}x0e + {nv}x1o + {nv}x2e + {nv}x1e + {nv}x2o + {ns}x0o']
        else:
            seq = [f'{ns}x0e', f'{ns}x0e + {nv}x1o', f'{ns}x0e + {nv}x1o + {nv}x1e',
                   f'{ns}x0e + {nv}x1o + {nv}x1e + {ns}x0o']
        faster = sh_lmax == 1 and not use_second_order_repr
        convs = []
        for i in range(num_conv_layers):
            for _ in range(9):
                convs.append(TensorProductConvLayer(seq[min(i, 3)], self.sh_irreps, seq[min(i + 1, 3)], 3 * ns,
                                                    residual=False, batch_norm=batch_norm, dropout=dropout,
                                                    faster=faster))
        self.conv_layers = nn.ModuleList(convs)
        last_irreps = convs[-1].out_irreps

        if confidence_mode:                                             # :124-146
            cin = (2 * ns if num_conv_layers >= 3 else ns) * (2 if flexible_sidechains else 1)
            bn = (lambda: nn.Identity()) if confidence_no_batchnorm else (lambda: nn.BatchNorm1d(ns))
            self.confidence_predictor = nn.Sequential(
                nn.Linear(cin, ns), bn(), nn.ReLU(), nn.Dropout(confidence_dropout),
                nn.Linear(ns, ns), bn(), nn.ReLU(), nn.Dropout(confidence_dropout),
                nn.Linear(ns, num_confidence_outputs))
        else:                                                           # :161-234
            self.center_distance_expansion = GaussianSmearing(0.0, center_max_distance, distance_embed_dim)
            self.center_edge_embedding = _mlp(distance_embed_dim + sigma_embed_dim, ns, ns, dropout)
            self.final_conv = TensorProductConvLayer(last_irreps, self.sh_irreps, '2x1o + 2x1e', 2 * ns, residual=False,
                                                     dropout=dropout, batch_norm=batch_norm, faster=faster)
            self.tr_final_layer = nn.Sequential(nn.Linear(1 + sigma_embed_dim, ns), nn.Dropout(dropout), nn.ReLU(), nn.Linear(ns, 1))
            self.rot_final_layer = nn.Sequential(nn.Linear(1 + sigma_embed_dim, ns), nn.Dropout(dropout), nn.ReLU(), nn.Linear(ns, 1))
            if not no_torsion:
                self.final_edge_embedding = _mlp(distance_embed_dim, ns, ns, dropout)
                self.final_tp_tor = FullTensorProduct(self.sh_irreps, '2e')
                self.tor_bond_conv = TensorProductConvLayer(last_irreps, self.final_tp_tor.irreps_out, f'{ns}x0o + {ns}x0e',
                                                            3 * ns, residual=False, dropout=dropout, batch_norm=batch_norm)
                self.tor_final_layer = nn.Sequential(nn.Linear(2 * ns, ns, bias=False), nn.Tanh(), nn.Dropout(dropout),
                                                     nn.Linear(ns, 1, bias=False))
            if flexible_sidechains:
                self.sidechain_final_edge_embedding = _mlp(distance_embed_dim, ns, ns, dropout)
                self.final_tp_sc_tor = FullTensorProduct(self.sh_irreps, '2e')
                self.sc_tor_bond_conv = TensorProductConvLayer(last_irreps, self.final_tp_sc_tor.irreps_out, f'{ns}x0o + {ns}x0e',
                                                               3 * ns, residual=False, dropout=dropout, batch_norm=batch_norm)
                self.sc_tor_final_layer = nn.Sequential(nn.Linear(2 * ns, ns, bias=False), nn.Tanh(), nn.Dropout(dropout),
                                                        nn.Linear(ns, 1, bias=False))

    # ------------------------------------------------------------------ helpers
    def _sh(self, vec):
        return spherical_harmonics(self.sh_irreps, vec, normalize=True, normalization='component')

    def _edge_weight(self, vec, max_norm):                               # :438-442
        if self.smooth_edges:
            nn_ = torch.clip(vec.norm(dim=-1) * np.pi / max_norm, max=np.pi)
            return 0.5 * (torch.cos(nn_) + 1.0).unsqueeze(-1)
        return 1.0

    # ------------------------------------------------------------------ graph builders
    def build_lig_conv_graph(self, data):                                # :444-484
        lig = data['ligand']
        lig.node_sigma_emb = self.timestep_emb_func(lig.node_t[self._t_key])
        radius_edges = cluster.radius_graph(lig.pos, self.lig_max_radius, lig.batch)
        edge_index = torch.cat([data['ligand', 'ligand'].edge_index, radius_edges], 1).long()
        edge_attr = torch.cat([data['ligand', 'ligand'].edge_attr,
                               torch.zeros(radius_edges.shape[-1], self.in_lig_edge_features)], 0)
        edge_attr = torch.cat([edge_attr, lig.node_sigma_emb[edge_index[0]]], 1)
        node_attr = torch.cat([lig.x, lig.node_sigma_emb], 1)
        src, dst = edge_index
        vec = lig.pos[dst] - lig.pos[src]
        edge_attr = torch.cat([edge_attr, self.lig_distance_expansion(vec.norm(dim=-1))], 1)
        return node_attr, edge_index, edge_attr, self._sh(vec), self._edge_weight(vec, self.lig_max_radius)

    def build_rec_conv_graph(self, data):                                # :486-511
        rec = data['receptor']
        rec.node_sigma_emb = self.timestep_emb_func(rec.node_t[self._t_key])
        node_attr = torch.cat([rec.x, rec.node_sigma_emb], 1)
        edge_index = data['receptor', 'receptor'].edge_index
        src, dst = edge_index
        vec = rec.pos[dst.long()] - rec.pos[src.long()]
        edge_attr = torch.cat([rec.node_sigma_emb[src.long()], self.rec_distance_expansion(vec.norm(dim=-1))], 1)
        return node_attr, edge_index, edge_attr, self._sh(vec), self._edge_weight(vec, self.rec_max_radius)

    def build_atom_conv_graph(self, data):                               # :513-537
        atom = data['atom']
        atom.node_sigma_emb = self.timestep_emb_func(atom.node_t[self._t_key])
        node_attr = torch.cat([atom.x, atom.node_sigma_emb], 1)
        edge_index = cluster.knn_graph(atom.pos, k=self.atom_max_neighbors if self.atom_max_neighbors else 32,
                                       batch=atom.batch)
        vec = atom.pos[edge_index[1]] - atom.pos[edge_index[0]]
        data['atom', 'atom'].edge_index = edge_index
        edge_attr = torch.cat([atom.node_sigma_emb[edge_index[0]], self.lig_distance_expansion(vec.norm(dim=-1))], 1)
        return node_attr, edge_index, edge_attr, self._sh(vec), self._edge_weight(vec, self.lig_max_radius)

    def build_cross_conv_graph(self, data, cutoff):                      # :539-583
        lig, rec, atom = data['ligand'], data['receptor'], data['atom']
        if torch.is_tensor(cutoff):
            lr = cluster.radius(rec.pos / cutoff[rec.batch], lig.pos / cutoff[lig.batch], 1, rec.batch, lig.batch,
                                max_num_neighbors=10000)
        else:
            lr = cluster.radius(rec.pos, lig.pos, cutoff, rec.batch, lig.batch, max_num_neighbors=10000)
        lr_vec = rec.pos[lr[1]] - lig.pos[lr[0]]
        lr_attr = torch.cat([lig.node_sigma_emb[lr[0]], self.cross_distance_expansion(lr_vec.norm(dim=-1))], 1)
        cutoff_d = cutoff[lig.batch[lr[0]]].squeeze() if torch.is_tensor(cutoff) else cutoff
        lr_w = self._edge_weight(lr_vec, cutoff_d)
        la = cluster.radius(atom.pos, lig.pos, self.lig_max_radius, atom.batch, lig.batch, max_num_neighbors=10000)
        la_vec = atom.pos[la[1]] - lig.pos[la[0]]
        la_attr = torch.cat([lig.node_sigma_emb[la[0]], self.cross_distance_expansion(la_vec.norm(dim=-1))], 1)
        la_w = self._edge_weight(la_vec, self.lig_max_radius)
        ar = data['atom', 'receptor'].edge_index
        ar_vec = rec.pos[ar[1].long()] - atom.pos[ar[0].long()]
        ar_attr = torch.cat([atom.node_sigma_emb[ar[0].long()], self.rec_distance_expansion(ar_vec.norm(dim=-1))], 1)
        return lr, lr_attr, self._sh(lr_vec), lr_w, la, la_attr, self._sh(la_vec), la_w, ar, ar_attr, self._sh(ar_vec), 1

    def build_center_conv_graph(self, data):                             # :585-599
        lig = data['ligand']
        edge_index = torch.cat([lig.batch.unsqueeze(0), torch.arange(len(lig.batch)).unsqueeze(0)], 0)
        center = torch.zeros((data.num_graphs, 3))
        center.index_add_(0, lig.batch, lig.pos)
        center = center / torch.bincount(lig.batch).unsqueeze(1)
        vec = lig.pos[edge_index[1]] - center[edge_index[0]]
        attr = torch.cat([self.center_distance_expansion(vec.norm(dim=-1)), lig.node_sigma_emb[edge_index[1]]], 1)
        return edge_index, attr, self._sh(vec)

    def build_bond_conv_graph(self, data):                               # :601-616
        lig = data['ligand']
        bonds = data['ligand', 'ligand'].edge_index[:, lig.edge_mask].long()
        bond_pos = (lig.pos[bonds[0]] + lig.pos[bonds[1]]) / 2
        edge_index = cluster.radius(lig.pos, bond_pos, self.lig_max_radius, lig.batch, lig.batch[bonds[0]])
        vec = lig.pos[edge_index[1]] - bond_pos[edge_index[0]]
        attr = self.final_edge_embedding(self.lig_distance_expansion(vec.norm(dim=-1)))
        return bonds, edge_index, attr, self._sh(vec), self._edge_weight(vec, self.lig_max_radius)

    @staticmethod
    def get_sc_tor_bonds(data):                                          # :638-652
        _, counts = data['atom'].batch.unique(sorted=True, return_counts=True)
        off = torch.cat((torch.zeros(1), counts.cumsum(0)))[:-1].long()
        return off[data['flexResidues'].batch] + data['flexResidues'].edge_idx.T.long()

    def build_sidechain_conv_graph(self, data):                          # :618-636
        atom = data['atom']
        bonds = self.get_sc_tor_bonds(data)
        bond_pos = (atom.pos[bonds[0]] + atom.pos[bonds[1]]) / 2
        edge_index = cluster.radius(atom.pos, bond_pos, self.lig_max_radius, atom.batch, data['flexResidues'].batch)
        vec = atom.pos[edge_index[1]] - bond_pos[edge_index[0]]
        attr = self.sidechain_final_edge_embedding(self.lig_distance_expansion(vec.norm(dim=-1)))
        return bonds, edge_index, attr, self._sh(vec), self._edge_weight(vec, self.lig_max_radius)

    # ------------------------------------------------------------------ forward
    def forward(self, data, return_layers=False):                        # :238-436
        ns = self.ns
        if self.no_aminoacid_identities:
            data['receptor'].x = data['receptor'].x * 0
        ct = [data.complex_t[k] for k in ('tr', 'rot', 'tor', 'sc_tor')]
        tr_sigma, rot_sigma, tor_sigma, sc_sigma = ct if self.confidence_mode else self.t_to_sigma(*ct)

        lig_x, ll, ll_attr, ll_sh, ll_w = self.build_lig_conv_graph(data)
        lig_x, ll_attr = self.lig_node_embedding(lig_x), self.lig_edge_embedding(ll_attr)
        rec_x, rr, rr_attr, rr_sh, rr_w = self.build_rec_conv_graph(data)
        rec_x, rr_attr = self.rec_node_embedding(rec_x), self.rec_edge_embedding(rr_attr)
        atom_x, aa, aa_attr, aa_sh, aa_w = self.build_atom_conv_graph(data)
        atom_x, aa_attr = self.atom_node_embedding(atom_x), self.atom_edge_embedding(aa_attr)
        cutoff = (tr_sigma * 3 + 20).unsqueeze(1) if self.dynamic_max_cross else self.cross_max_distance
        lr, lr_attr, lr_sh, lr_w, la, la_attr, la_sh, la_w, ar, ar_attr, ar_sh, ar_w = self.build_cross_conv_graph(data, cutoff)
        lr_attr, la_attr, ar_attr = self.lr_edge_embedding(lr_attr), self.la_edge_embedding(la_attr), self.ar_edge_embedding(ar_attr)
        layers = []
        flip = lambda e: torch.flip(e, dims=[0])
        nl, na, nr = lig_x.shape[0], atom_x.shape[0], rec_x.shape[0]
        for l in range(self.num_conv_layers):                            # :271-324
            C = self.conv_layers
            last = l == self.num_conv_layers - 1
            ea = torch.cat([ll_attr, lig_x[ll[0], :ns], lig_x[ll[1], :ns]], -1)
            lig_up = C[9 * l](lig_x, ll, ea, ll_sh, edge_weight=ll_w)
            ea = torch.cat([lr_attr, lig_x[lr[0], :ns], rec_x[lr[1], :ns]], -1)
            lr_up = C[9 * l + 1](rec_x, lr, ea, lr_sh, out_nodes=nl, edge_weight=lr_w)
            ea = torch.cat([la_attr, lig_x[la[0], :ns], atom_x[la[1], :ns]], -1)
            la_up = C[9 * l + 2](atom_x, la, ea, la_sh, out_nodes=nl, edge_weight=la_w)
            if self.flexible_sidechains or not last:
                ea = torch.cat([aa_attr, atom_x[aa[0], :ns], atom_x[aa[1], :ns]], -1)
                atom_up = C[9 * l + 3](atom_x, aa, ea, aa_sh, edge_weight=aa_w)
                ea = torch.cat([la_attr, atom_x[la[1], :ns], lig_x[la[0], :ns]], -1)
                al_up = C[9 * l + 4](lig_x, flip(la), ea, la_sh, out_nodes=na, edge_weight=la_w)
                ea = torch.cat([ar_attr, atom_x[ar[0], :ns], rec_x[ar[1], :ns]], -1)
                ar_up = C[9 * l + 5](rec_x, ar, ea, ar_sh, out_nodes=na, edge_weight=ar_w)
                if not last:
                    ea = torch.cat([rr_attr, rec_x[rr[0], :ns], rec_x[rr[1], :ns]], -1)
                    rec_up = C[9 * l + 6](rec_x, rr, ea, rr_sh, edge_weight=rr_w)
                    ea = torch.cat([lr_attr, rec_x[lr[1], :ns], lig_x[lr[0], :ns]], -1)
                    rl_up = C[9 * l + 7](lig_x, flip(lr), ea, lr_sh, out_nodes=nr, edge_weight=lr_w)
                    ea = torch.cat([ar_attr, rec_x[ar[1], :ns], atom_x[ar[0], :ns]], -1)
                    ra_up = C[9 * l + 8](atom_x, flip(ar), ea, ar_sh, out_nodes=nr, edge_weight=ar_w)
            lig_x = F.pad(lig_x, (0, lig_up.shape[-1] - lig_x.shape[-1])) + lig_up + la_up + lr_up
            if self.flexible_sidechains or not last:
                atom_x = F.pad(atom_x, (0, atom_up.shape[-1] - atom_x.shape[-1])) + atom_up + al_up + ar_up
                if not last:
                    rec_x = F.pad(rec_x, (0, rec_up.shape[-1] - rec_x.shape[-1])) + rec_up + ra_up + rl_up
            layers.append((lig_x, atom_x, rec_x))
        self._debug = dict(ll=ll, aa=aa, lr=lr, la=la, layers=layers)

        n_flex = 0 if (not self.flexible_sidechains or len(data['flexResidues']) == 0) else data['flexResidues'].edge_idx.shape[0]
        if self.confidence_mode:                                          # :329-353
            sl = torch.cat([lig_x[:, :ns], lig_x[:, -ns:]], 1) if self.num_conv_layers >= 3 else lig_x[:, :ns]
            sl = cluster.scatter_mean(sl, data['ligand'].batch, dim=0, dim_size=data.num_graphs)
            cin = sl
            if self.flexible_sidechains:
                if n_flex > 0:
                    fa = self.get_sc_tor_bonds(data).unique()
                    sa = torch.cat([atom_x[fa, :ns], atom_x[fa, -ns:]], 1) if self.num_conv_layers >= 3 else atom_x[fa, :ns]
                    sa = cluster.scatter_mean(sa, data['atom'].batch[fa], dim=0, dim_size=sl.shape[0])
                else:
                    sa = torch.zeros_like(sl)
                cin = torch.cat([cin, sa], 1)
            return self.confidence_predictor(cin).squeeze(dim=-1)

        cei, cattr, csh = self.build_center_conv_graph(data)              # :357-384
        cattr = self.center_edge_embedding(cattr)
        cattr = torch.cat([cattr, lig_x[cei[1] if self.fixed_center_conv else cei[0], :ns]], -1)
        g = self.final_conv(lig_x, cei, cattr, csh, out_nodes=data.num_graphs)
        tr_pred = g[:, :3] + g[:, 6:9]
        rot_pred = g[:, 3:6] + g[:, 9:]
        data.graph_sigma_emb = self.timestep_emb_func(data.complex_t[self._t_key])
        tr_norm = torch.linalg.vector_norm(tr_pred, dim=1).unsqueeze(1)
        tr_pred = tr_pred / tr_norm * self.tr_final_layer(torch.cat([tr_norm, data.graph_sigma_emb], 1))
        rot_norm = torch.linalg.vector_norm(rot_pred, dim=1).unsqueeze(1)
        rot_pred = rot_pred / rot_norm * self.rot_final_layer(torch.cat([rot_norm, data.graph_sigma_emb], 1))
        if self.scale_by_sigma:
            tr_pred = tr_pred / tr_sigma.unsqueeze(1)
            rot_pred = rot_pred * torch.from_numpy(np.asarray(self.so3_score_norm(rot_sigma.numpy()))).float().unsqueeze(1)

        lig = data['ligand']
        if self.no_torsion or lig.edge_mask.sum() == 0:                    # :386-408
            tor_pred = torch.empty(0)
        else:
            bonds, ei, eattr, esh, ew = self.build_bond_conv_graph(data)
            bvec = lig.pos[bonds[1]] - lig.pos[bonds[0]]
            battr = lig_x[bonds[0]] + lig_x[bonds[1]]
            bsh = spherical_harmonics('2e', bvec, normalize=True, normalization='component')
            esh = self.final_tp_tor(esh, bsh[ei[0]])
            eattr = torch.cat([eattr, lig_x[ei[1], :ns], battr[ei[0], :ns]], -1)
            tor_pred = self.tor_bond_conv(lig_x, ei, eattr, esh, out_nodes=int(lig.edge_mask.sum()), reduce='mean', edge_weight=ew)
            tor_pred = self.tor_final_layer(tor_pred).squeeze(1)
            if self.scale_by_sigma:
                es = tor_sigma[lig.batch][data['ligand', 'ligand'].edge_index[0]][lig.edge_mask]
                tor_pred = tor_pred * torch.sqrt(torch.tensor(self.torus_score_norm(es.numpy())).float())
        if n_flex == 0:                                                   # :410-434
            sc_pred = torch.empty(0)
        else:
            bonds, ei, eattr, esh, ew = self.build_sidechain_conv_graph(data)
            atom = data['atom']
            bvec = atom.pos[bonds[1]] - atom.pos[bonds[0]]
            battr = atom_x[bonds[0]] + atom_x[bonds[1]]
            bsh = spherical_harmonics('2e', bvec, normalize=True, normalization='component')
            esh = self.final_tp_sc_tor(esh, bsh[ei[0]])
            eattr = torch.cat([eattr, atom_x[ei[1], :ns], battr[ei[0], :ns]], -1)
            sc_pred = self.sc_tor_bond_conv(atom_x, ei, eattr, esh, out_nodes=n_flex, reduce='mean', edge_weight=ew)
            sc_pred = self.sc_tor_final_layer(sc_pred).squeeze(1)
            if self.scale_by_sigma:
                es = sc_sigma[data['flexResidues'].batch]
                sc_pred = sc_pred * torch.sqrt(torch.tensor(self.torus_score_norm(es.numpy())).float())
        return tr_pred, rot_pred, tor_pred, sc_pred
