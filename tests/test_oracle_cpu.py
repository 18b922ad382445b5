"""CPU suite: the oracle against its golden vectors and algebraic self-checks (no reference tests exist,
SURVEY.md F2 -- 'parity unpinned'), host logic, and the C-ABI library's exported symbols."""
import copy
import ctypes
import math
import os
import re

import numpy as np
import pytest
import torch
from scipy.spatial.transform import Rotation as R

from diffdock_pocket_b200 import _lib, inputs, so3, torus, tp, utils
from diffdock_pocket_b200.hetero import Batch, DataLoader
from oracle import cluster, diffusion_ref as D, e3nn_mini as E, factory, sampling_ref as S
from oracle.score_model_ref import FasterTensorProduct, TensorProductConvLayer
import _common as T

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _rot(seed):
    return torch.from_numpy(R.random(random_state=seed).as_matrix())


# ------------------------------------------------------------------------------------------- e3nn pieces
def test_cg_identities():
    eps = torch.zeros(3, 3, 3, dtype=torch.float64)
    for i, j, k in ((0, 1, 2), (1, 2, 0), (2, 0, 1)):
        eps[i, j, k], eps[j, i, k] = 1, -1
    assert torch.allclose(E.wigner_3j(1, 1, 1), eps / math.sqrt(6), atol=1e-12)
    assert torch.allclose(E.wigner_3j(1, 1, 0)[:, :, 0], torch.eye(3, dtype=torch.float64) / math.sqrt(3), atol=1e-12)
    u = torch.nn.functional.normalize(torch.randn(9, 3, dtype=torch.float64), dim=-1)
    y2 = E.spherical_harmonics(2, u) / math.sqrt(5)
    assert torch.allclose(torch.einsum('ijk,zi,zj->zk', E.wigner_3j(1, 1, 2), u, u), math.sqrt(2 / 15) * y2, atol=1e-12)
    for l in (0, 1, 2):
        y = E.spherical_harmonics(l, u)
        assert torch.allclose((y ** 2).sum(-1), torch.full((9,), 2.0 * l + 1, dtype=torch.float64), atol=1e-12)


def test_product_cg_matches_oracle():
    for l1 in range(3):
        for l2 in range(3):
            for l3 in range(abs(l1 - l2), min(l1 + l2, 3) + 1):
                assert np.allclose(tp.wigner_3j(l1, l2, l3), E.wigner_3j(l1, l2, l3).numpy(), atol=1e-12)


def _eval_spec(spec, x, sh, w):
    n = x.shape[0]
    out = np.zeros((n, spec.f_out))
    ct = np.array(spec.ctab)
    for g in spec.groups:
        Cg = ct[g['c_off']:g['c_off'] + g['d1'] * g['d2'] * g['d_out']].reshape(g['d1'], g['d2'], g['d_out'])
        xs = x[:, g['x_off']:g['x_off'] + g['mul_in'] * g['d1']].reshape(n, g['mul_in'], g['d1'])
        basis = np.einsum('ijk,eui,ej->euk', Cg, xs, sh[:, g['sh_off']:g['sh_off'] + g['d2']])
        W = w[:, g['w_off']:g['w_off'] + g['mul_in'] * g['mul_out']].reshape(n, g['mul_in'], g['mul_out'])
        out[:, g['out_off']:g['out_off'] + g['mul_out'] * g['d_out']] += np.einsum('euo,euk->eok', W, basis).reshape(n, -1)
    return out


SEQ = ['60x0e', '60x0e + 10x1o', '60x0e + 10x1o + 10x1e', '60x0e + 10x1o + 10x1e + 60x0o']


@pytest.mark.parametrize('a,b,w', [(0, 1, 4200), (1, 2, 5000), (2, 3, 5800), (3, 3, 10000)])
def test_row_groups_equal_faster_tp(a, b, w):
    spec, ref = tp.faster_tp_spec(SEQ[a], SEQ[b]), FasterTensorProduct(SEQ[a], '1x0e+1x1o', SEQ[b])
    assert spec.weight_numel == ref.weight_numel == w
    x = torch.randn(5, spec.f_in, dtype=torch.float64)
    sh = E.spherical_harmonics('1x0e+1x1o', torch.randn(5, 3, dtype=torch.float64))
    wt = torch.randn(5, w, dtype=torch.float64)
    assert np.abs(ref(x, sh, wt).numpy() - _eval_spec(spec, x.numpy(), sh.numpy(), wt.numpy())).max() < 1e-12


def test_row_groups_equal_fctp_and_torsion_paths():
    ft = E.FullTensorProduct('1x0e+1x1o', '1x2e')
    assert repr(ft.irreps_out) == '1x1o+1x2o+1x2e+1x3o'
    ref = E.FullyConnectedTensorProduct(SEQ[3], ft.irreps_out, '60x0o+60x0e')
    assert ref.weight_numel == 1200 and ref.instructions == [(1, 0, 1), (2, 0, 0)]
    spec = tp.fctp_spec(SEQ[3], tp.full_tp_out_irreps('1x0e+1x1o', '1x2e'), '60x0o + 60x0e', sh_keep=[0])
    x = torch.randn(4, 180, dtype=torch.float64)
    sh = E.spherical_harmonics('1x0e+1x1o', torch.randn(4, 3, dtype=torch.float64))
    sht = ft(sh, torch.randn(4, 5, dtype=torch.float64))
    w = torch.randn(4, 1200, dtype=torch.float64)
    assert np.abs(ref(x, sht, w).numpy() - _eval_spec(spec, x.numpy(), sht[:, :3].numpy(), w.numpy())).max() < 1e-12
    ref2 = E.FullyConnectedTensorProduct(SEQ[3], '1x0e+1x1o+1x2e', SEQ[3])
    spec2 = tp.fctp_spec(SEQ[3], '1x0e+1x1o+1x2e', SEQ[3])
    assert ref2.weight_numel == spec2.weight_numel == 10200
    sh9 = torch.randn(4, 9, dtype=torch.float64)
    w2 = torch.randn(4, 10200, dtype=torch.float64)
    assert np.abs(ref2(x, sh9, w2).numpy() - _eval_spec(spec2, x.numpy(), sh9.numpy(), w2.numpy())).max() < 1e-12


def test_conv_layer_equivariance():
    """SE(3) equivariance of the oracle conv (FasterTP path): rotating inputs rotates vector outputs."""
    torch.manual_seed(0)
    layer = TensorProductConvLayer(SEQ[3], '1x0e+1x1o', SEQ[3], 24, residual=False, batch_norm=False, faster=True).double()
    n, e = 6, 20
    x = torch.randn(n, 180, dtype=torch.float64)
    ei = torch.randint(0, n, (2, e))
    vec = torch.randn(e, 3, dtype=torch.float64)
    ea = torch.randn(e, 24, dtype=torch.float64)
    Rm = _rot(1)

    def rot_feat(f):
        f = f.clone()
        for off in (60, 90):
            f[:, off:off + 30] = (f[:, off:off + 30].reshape(-1, 10, 3) @ Rm.T).reshape(-1, 30)
        return f
    out = layer(x, ei, ea, E.spherical_harmonics('1x0e+1x1o', vec))
    out_r = layer(rot_feat(x), ei, ea, E.spherical_harmonics('1x0e+1x1o', vec @ Rm.T))
    assert torch.allclose(rot_feat(out), out_r, atol=1e-10)


# ------------------------------------------------------------------------------------------- cluster
def _brute_radius(x, y, r, bx, by, cap):
    rows, cols = [], []
    for j in range(len(y)):
        c = 0
        for i in range(len(x)):
            if bx[i] != by[j]:
                continue
            d = np.float32(0)
            for k in range(3):
                t = np.float32(x[i, k]) - np.float32(y[j, k])
                d = np.float32(np.float64(t) * np.float64(t) + np.float64(d))      # fma: exact product, one rounding
            if d < np.float32(r * r):
                rows.append(j)
                cols.append(i)
                c += 1
            if c >= cap:
                break
    return np.array([rows, cols])


def test_radius_and_knn_semantics():
    rng = np.random.RandomState(0)
    x = torch.from_numpy(rng.randn(60, 3).astype(np.float32) * 3)
    y = torch.from_numpy(rng.randn(25, 3).astype(np.float32) * 3)
    bx = torch.from_numpy(np.sort(rng.randint(0, 3, 60)))
    by = torch.from_numpy(np.sort(rng.randint(0, 3, 25)))
    for cap in (3, 32):
        e = cluster.radius(x, y, 4.0, bx, by, cap).numpy()
        assert np.array_equal(e, _brute_radius(x.numpy(), y.numpy(), 4.0, bx.numpy(), by.numpy(), cap))
    rg = cluster.radius_graph(x, 4.0, bx, max_num_neighbors=4)
    assert (rg[0] != rg[1]).all() and (bx[rg[0]] == bx[rg[1]]).all()
    assert (torch.bincount(rg[1], minlength=60) <= 5).all()
    kg = cluster.knn_graph(x, 5, bx)
    assert kg.shape[1] == 5 * 60
    d = (x[kg[0]] - x[kg[1]]).norm(dim=-1).reshape(60, 5)
    assert (d[:, 1:] >= d[:, :-1]).all()                                  # ascending distance per centre
    full = torch.cdist(x.double(), x.double())
    full[bx[:, None] != bx[None, :]] = 1e9
    full.fill_diagonal_(1e9)
    assert torch.equal(kg[0].reshape(60, 5).sort(dim=1)[0], full.topk(5, largest=False)[1].sort(dim=1)[0])
    # duplicates: ties keep the lower index first; a centre that is not among its own k+1 best keeps all k+1
    xd = torch.zeros(4, 3)
    assert cluster.knn_graph(xd, 2).tolist() == [[1, 2, 0, 2, 0, 1, 0, 1, 2], [0, 0, 1, 1, 2, 2, 3, 3, 3]]
    # empty inputs
    assert cluster.radius(x[:0], y, 1.0, bx[:0], by, 32).shape == (2, 0)


def test_scatter_mean_matches_dense():
    src = torch.randn(30, 7)
    idx = torch.randint(0, 9, (30,))
    dense = torch.zeros(11, 30)
    dense[idx, torch.arange(30)] = 1
    ref = (dense @ src) / dense.sum(1, keepdim=True).clamp(min=1)
    assert torch.allclose(cluster.scatter_mean(src, idx, dim_size=11), ref, atol=1e-6)


# ------------------------------------------------------------------------------------------- geometry / tables
def test_rodrigues_and_kabsch():
    v = torch.randn(3) * 0.7
    assert np.allclose(D.axis_angle_to_matrix(v).numpy(), R.from_rotvec(v.numpy()).as_matrix(), atol=1e-6)
    A = torch.randn(3, 17)
    Rm, t = _rot(3).float(), torch.randn(3, 1)
    Rk, tk = D.kabsch(A, Rm @ A + t)
    assert torch.allclose(Rk, Rm, atol=1e-4) and torch.allclose(tk, t, atol=1e-4)


def test_so3_and_torus_tables_match_literal_restatement():
    idx = [0, 250, 500, 750, 999]
    assert np.abs(D.so3_exp_score_norms(idx) / so3.exp_score_norms()[idx] - 1).max() < 1e-6
    sig = torch.tensor([0.03, 0.4, 1.55])
    assert np.array_equal(D.so3_eps_index(sig.numpy()), so3.eps_index(sig.numpy()))
    tab = torus.score_norm_table()
    assert tab.shape == (5001,) and np.isfinite(tab).all()
    assert np.array_equal(D.torus_sigma_index(np.array([0.03, 1.0, 3.14])), torus.sigma_index(np.array([0.03, 1.0, 3.14])))
    # the Monte-Carlo estimate tracks the analytic small-sigma limit E[score^2] = 1/sigma^2 (within MC noise)
    s = 10 ** np.linspace(np.log10(3e-3), np.log10(2), 5001)[[0, 500, 1000]] * np.pi
    assert np.abs(tab[[0, 500, 1000]] * s ** 2 - 1).max() < 0.06
    sel, lit = D.torus_score_norm_table(7, n_mc=20000, sigma_stride=1250)
    assert np.abs(lit / tab[sel] - 1).max() < 0.06                           # same estimator, different draws


def test_schedule_and_step_coefficients():
    from diffdock_pocket_b200 import diffusion_utils as du, sampling as ps
    sch = du.get_t_schedule('expbeta', 20)
    assert np.allclose(sch, np.linspace(1, 0, 21)[:-1])
    sa = utils.score_model_args()
    from functools import partial
    t2s = partial(du.t_to_sigma, args=sa)
    t, coef = ps.step_coefficients(3, 20, (sch,) * 4, t2s, sa, False, [0.9766, 6.0774, 6.7616, 1.4488],
                                   [1.5103, 0.8141, 0.7662, 1.3396], 0.48884, True)
    tr_sigma = sa.tr_sigma_min ** (1 - sch[3]) * sa.tr_sigma_max ** sch[3]
    g = tr_sigma * np.sqrt(2 * np.log(sa.tr_sigma_max / sa.tr_sigma_min))
    sd = np.exp(0.48884 * np.log(sa.tr_sigma_max) + (1 - 0.48884) * np.log(sa.tr_sigma_min))
    lam = (sd + tr_sigma) / (sd + tr_sigma / 0.9766)
    assert np.isclose(coef[0], g ** 2 * 0.05 * (lam + 0.9766 * 1.5103 / 2))
    assert np.isclose(coef[1], g * np.sqrt(0.05 * (1 + 1.5103)))


# ------------------------------------------------------------------------------------------- data + golden
def test_fixture_graph_shapes_and_collate():
    g = T.graph('3dpf_holo')
    assert g['ligand'].x.shape == (37, 16) and g['receptor'].x.shape == (139, 1281) and g['atom'].x.shape == (1111, 4)
    assert g['receptor', 'receptor'].edge_index.shape == (2, 3318) and int(g['ligand'].edge_mask.sum()) == 5
    assert g['flexResidues'].edge_idx.shape == (17, 2)
    assert np.allclose(g.original_center.numpy(), [[9.7742, 27.2863, 14.6573]], atol=1e-3)     # README.md:47
    a = T.graph('3dpf_apo')
    assert a['receptor'].x.shape[0] == 137 and a['atom'].x.shape[0] == 1098 and a['flexResidues'].edge_idx.shape == (17, 2)
    b = Batch.from_data_list([copy.deepcopy(g) for _ in range(3)])
    assert b.num_graphs == 3 and b['ligand'].batch.tolist() == [0] * 37 + [1] * 37 + [2] * 37
    assert int(b['ligand', 'ligand'].edge_index.max()) == 110 and int(b['atom', 'receptor'].edge_index[1].max()) == 3 * 139 - 1
    assert int(b['flexResidues'].edge_idx.max()) < 1111 and b['flexResidues'].batch.tolist() == [0] * 17 + [1] * 17 + [2] * 17
    assert len(list(DataLoader([g] * 5, batch_size=2))) == 3
    # a graph without flexible residues: look-ups like len(data['flexResidues']) leave an empty store behind (as in PyG),
    # which must not break collation; the rigid case also goes through the oracle sampler
    from diffdock_pocket_b200 import inputs
    r = inputs.synthetic_complex(21, n_lig=3, n_res=12, flexible_residues=0)
    assert 'flexResidues' not in r and len(r['flexResidues']) == 0 and 'flexResidues' in r
    rb = Batch.from_data_list([copy.deepcopy(r), copy.deepcopy(r)])
    assert rb.num_graphs == 2 and 'flexResidues' not in rb.node_types and rb['ligand'].pos.shape == (6, 3)
    # cross-complex batch where only some complexes have flexible residues: the others contribute zero entries
    f = inputs.synthetic_complex(22, n_lig=5, n_res=14, flexible_residues=2)
    mb = Batch.from_data_list([copy.deepcopy(r), copy.deepcopy(f), copy.deepcopy(r), copy.deepcopy(f)])
    nf = f['flexResidues'].edge_idx.shape[0]
    assert mb['flexResidues'].edge_idx.shape[0] == 2 * nf and mb['flexResidues'].batch.tolist() == [1] * nf + [3] * nf
    bad = copy.deepcopy(f)
    del bad['atom']
    with pytest.raises(ValueError, match="node type 'atom' is missing"):
        Batch.from_data_list([copy.deepcopy(f), bad])


def test_oracle_sampler_two_steps_runs_and_moves_poses():
    m, c, om, oc, sa, ca = T.models(torch.device('cpu'), small=True)
    g = inputs.synthetic_complex(3, n_lig=12, n_res=20, flexible_residues=2)
    dl = T.randomized_list(g, 2, sa, seed=1)
    before = [d['ligand'].pos.clone() for d in dl]
    sch = D.get_t_schedule(4)
    from functools import partial
    torch.manual_seed(5)
    out, conf = S.sampling(dl, om, 4, sch, sch, sch, sch, partial(D.t_to_sigma, args=sa), sa, confidence_model=oc, batch_size=2,
                           temp_sampling=[0.9766, 6.0774, 6.7616, 1.4488], temp_psi=[1.5103, 0.8141, 0.7662, 1.3396],
                           temp_sigma_data=0.48884, max_steps=2)
    assert conf.shape == (2,) and all((o['ligand'].pos - b).abs().max() > 1e-3 for o, b in zip(out, before))
    # bond lengths are preserved by rigid + torsion updates
    ei = g['ligand', 'ligand'].edge_index
    d0 = (g['ligand'].pos[ei[0]] - g['ligand'].pos[ei[1]]).norm(dim=-1)
    d1 = (out[0]['ligand'].pos[ei[0]] - out[0]['ligand'].pos[ei[1]]).norm(dim=-1)
    assert torch.allclose(d0, d1, atol=1e-3)


# ------------------------------------------------------------------------------------------- C ABI
def test_c_abi_library_exports_every_declared_symbol():
    """The shared library must load without a GPU and export exactly what include/ddp_b200.h declares."""
    import __graft_entry__ as ge
    ge.build()
    header = open(os.path.join(ROOT, 'include', 'ddp_b200.h')).read()
    declared = sorted(set(re.findall(r'\b(ddp_[a-z0-9_]+)\s*\(', header)))
    assert declared == _lib.EXPORTS, (set(declared) ^ set(_lib.EXPORTS))
    lib = ctypes.CDLL(_lib.SO_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert b'sm_100a' in _lib.lib().ddp_version()


def test_no_oracle_import_in_product():
    """The product package must never import / execute the oracle (no CPU fallback)."""
    pkg = os.path.join(ROOT, 'diffdock_pocket_b200')
    for f in os.listdir(pkg):
        if f.endswith('.py'):
            src = open(os.path.join(pkg, f)).read()
            assert not re.search(r'^\s*(from|import)\s+\.*oracle', src, flags=re.M), f


# ------------------------------------------------------------------------------------------- checkpoint loading
def test_strict_load_of_a_reference_style_checkpoint():
    """inference.py:434-435 loads with strict=True.  A reference checkpoint also carries the constant buffers of its e3nn
    modules (output_mask, _w3j_* of the compiled sub-modules, an empty ``weight``): those -- and only those -- are dropped."""
    sa = utils.score_model_args(ns=16, nv=4, num_conv_layers=2, sigma_embed_dim=32, distance_embed_dim=32, cross_distance_embed_dim=32)
    m, _, sa, _ = utils.build_models(torch.device('cpu'), score_args=sa, with_confidence=False, seed=1)
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    extra = {}
    for owner in ('tor_bond_conv.tp', 'sc_tor_bond_conv.tp', 'final_tp_tor', 'final_tp_sc_tor'):
        extra[f'{owner}.output_mask'] = torch.ones(7)
        extra[f'{owner}.weight'] = torch.Tensor()
        extra[f'{owner}._compiled_main_left_right._w3j_1_1_0'] = torch.randn(3, 3, 1)
        extra[f'{owner}._compiled_main_right._w3j_1_1_1'] = torch.randn(3, 3, 3)
    ckpt = dict(sd, **extra)
    m2, _, _, _ = utils.build_models(torch.device('cpu'), score_args=sa, with_confidence=False, seed=2)
    res = m2.load_state_dict(ckpt, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    for k, v in m2.state_dict().items():
        assert torch.equal(v, sd[k]), k
    m2.load_state_dict({'module.' + k: v for k, v in ckpt.items()}, strict=True)          # saved from DataParallel
    with pytest.raises(RuntimeError):                                                     # anything else stays strict
        m2.load_state_dict(dict(ckpt, **{'conv_layers.0.fc.0.extra': torch.zeros(1)}), strict=True)
    with pytest.raises(RuntimeError):
        m2.load_state_dict(dict(ckpt, **{'lig_node_embedding.output_mask': torch.zeros(1)}), strict=True)
    missing = dict(ckpt)
    del missing['conv_layers.0.fc.0.weight']
    with pytest.raises(RuntimeError):
        m2.load_state_dict(missing, strict=True)
    with pytest.raises(RuntimeError):                                                     # a non-empty tp.weight is learned state
        m2.load_state_dict(dict(ckpt, **{'tor_bond_conv.tp.weight': torch.zeros(5)}), strict=True)
