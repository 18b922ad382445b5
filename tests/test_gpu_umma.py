"""GPU parity of the tensor-core (tcgen05 / TMEM) TP-conv kernel: bf16 mode within 1e-2, bf16x3 within 1e-4
(relative, max |a-b| / max |b|), against the CPU oracle; plus the full forward and the sampler in those modes."""
import copy
import os
from functools import partial

import numpy as np
import pytest
import torch

import _common as T
from diffdock_pocket_b200 import diffusion_utils as du, sampling as ps
from diffdock_pocket_b200.score_model import TensorProductConvLayer
from oracle import diffusion_ref as D, e3nn_mini as E, sampling_ref as S
from oracle.score_model_ref import TensorProductConvLayer as RefConv

pytestmark = pytest.mark.gpu
DEV = torch.device('cuda:0')
TOL = {'bf16': 1e-2, 'bf16x3': 1e-4}


def _seq(ns, nv):
    return [f'{ns}x0e', f'{ns}x0e + {nv}x1o', f'{ns}x0e + {nv}x1o + {nv}x1e', f'{ns}x0e + {nv}x1o + {nv}x1e + {ns}x0o']


@pytest.mark.parametrize('mode', ['bf16x3', 'bf16'])
@pytest.mark.parametrize('ns,nv,layer,n_edges', [(60, 10, 3, 1000), (60, 10, 0, 130), (60, 10, 1, 128), (60, 10, 2, 5),
                                                 (24, 6, 3, 700), (16, 4, 2, 300), (60, 10, 3, 40000),
                                                 (60, 10, 1, 45000), (24, 6, 3, 50000), (16, 4, 2, 30000)])
def test_conv_operator_tensor_core(mode, ns, nv, layer, n_edges):
    seq = _seq(ns, nv)
    in_ir, out_ir = seq[min(layer, 3)], seq[min(layer + 1, 3)]
    torch.manual_seed(0)
    prod = TensorProductConvLayer(in_ir, '1x0e+1x1o', out_ir, 3 * ns, residual=False, batch_norm=True, faster=True)
    bn = prod.batch_norm
    bn.running_mean.normal_(0, 0.2); bn.running_var.uniform_(0.5, 1.5); bn.weight.data.uniform_(0.7, 1.3); bn.bias.data.normal_(0, 0.2)
    ref = RefConv(in_ir, '1x0e+1x1o', out_ir, 3 * ns, residual=False, batch_norm=True, faster=True)
    ref.load_state_dict(prod.state_dict())
    prod, ref = prod.to(DEV).eval(), ref.eval()
    n = max(n_edges // 8, 4)
    x = torch.randn(n, E.Irreps(in_ir).dim)
    ei = torch.randint(0, n, (2, n_edges))
    ea = torch.randn(n_edges, 3 * ns)
    sh = E.spherical_harmonics('1x0e+1x1o', torch.randn(n_edges, 3))
    with torch.no_grad():
        want = ref(x, ei, ea, sh, out_nodes=n + 2)
        prod.conv_mode = 'fp32'
        base = prod(x.to(DEV), ei.to(DEV), ea.to(DEV), sh.to(DEV), out_nodes=n + 2)
        prod.conv_mode = mode
        got = prod(x.to(DEV), ei.to(DEV), ea.to(DEV), sh.to(DEV), out_nodes=n + 2)
    torch.cuda.synchronize()
    assert T.rel_err(base, want) < 1e-4
    assert T.rel_err(got, want) < TOL[mode], (mode, T.rel_err(got, want))


@pytest.mark.parametrize('mode', ['bf16x3', 'bf16'])
@pytest.mark.parametrize('run', [3, 8, 24, 100])
def test_conv_operator_sorted_aggregation(mode, run):
    """Edge lists sorted by aggregation node (receptor graph: runs of 24; ligand side of the cross edges: ~100): the epilogue
    pre-reduces each run inside the warp before the atomics.  Runs of random length around `run`, plus a ragged tail."""
    ns, nv = 60, 10
    seq = _seq(ns, nv)
    torch.manual_seed(1)
    prod = TensorProductConvLayer(seq[3], '1x0e+1x1o', seq[3], 3 * ns, residual=False, batch_norm=True, faster=True)
    ref = RefConv(seq[3], '1x0e+1x1o', seq[3], 3 * ns, residual=False, batch_norm=True, faster=True)
    ref.load_state_dict(prod.state_dict())
    prod, ref = prod.to(DEV).eval(), ref.eval()
    n = 200
    lens = torch.randint(max(run // 2, 1), run * 3 // 2 + 2, (n,))
    agg = torch.repeat_interleave(torch.arange(n), lens)[:4001]
    n_edges = agg.numel()
    ei = torch.stack([agg, torch.randint(0, n, (n_edges,))])
    x = torch.randn(n, E.Irreps(seq[3]).dim)
    ea = torch.randn(n_edges, 3 * ns)
    sh = E.spherical_harmonics('1x0e+1x1o', torch.randn(n_edges, 3))
    with torch.no_grad():
        want = ref(x, ei, ea, sh, out_nodes=n)
        prod.conv_mode = mode
        got = prod(x.to(DEV), ei.to(DEV), ea.to(DEV), sh.to(DEV), out_nodes=n)
    assert T.rel_err(got, want) < TOL[mode], (mode, run, T.rel_err(got, want))


@pytest.mark.parametrize('mode', ['bf16x3', 'bf16'])
def test_score_model_forward_tensor_core(mode):
    m, c, om, oc, sa, ca = T.models(DEV)
    dl = T.randomized_list(T.graph(), 3, sa, seed=0)
    b = T.batch_at(dl, 0.3)
    m.conv_mode = mode
    try:
        with torch.no_grad():
            pl = m.make_plan(copy.deepcopy(b))
            got = m.run_plan(pl, b.complex_t, return_layers=True)
            want = om(T.oracle_batch_at(dl, 0.3))
    finally:
        m.conv_mode = 'fp32'
    for l, ((gl, ga, gr), (wl, wa, wr)) in enumerate(zip(pl.last_layers, om._debug['layers'])):
        assert T.rel_err(gl, wl) < TOL[mode] and T.rel_err(ga[:, :wa.shape[1]], wa) < TOL[mode], (l, T.rel_err(gl, wl))
    for g_, w_ in zip(got, want):
        assert T.rel_err(g_, w_) < TOL[mode], T.rel_err(g_, w_)


@pytest.mark.parametrize('mode', ['bf16x3', 'bf16'])
def test_sampling_tensor_core_pose_rmsd(mode):
    """Final ligand / side-chain poses within 0.1 A RMSD of the CPU oracle with the tensor-core conv."""
    m, c, om, oc, sa, ca = T.models(DEV, small=True)
    from diffdock_pocket_b200 import inputs
    g = inputs.synthetic_complex(5, n_lig=20, n_res=40, flexible_residues=3)
    dl = T.randomized_list(g, 4, sa, seed=2)
    steps = 6
    sch = D.get_t_schedule(steps)
    kw = dict(temp_sampling=[0.9766, 6.0774, 6.7616, 1.4488], temp_psi=[1.5103, 0.8141, 0.7662, 1.3396], temp_sigma_data=0.48884)
    torch.manual_seed(11)
    ref, _ = S.sampling(T.oracle_list(dl), om, steps, sch, sch, sch, sch, partial(D.t_to_sigma, args=sa), sa, batch_size=4, **kw)
    torch.manual_seed(11)
    m.conv_mode = mode
    try:
        got, _ = ps.sampling(copy.deepcopy(dl), m, steps, sch, sch, sch, sch, DEV, partial(du.t_to_sigma, args=sa), sa, batch_size=4, **kw)
    finally:
        m.conv_mode = 'fp32'
    for a, r in zip(got, ref):
        rmsd = float(((a['ligand'].pos.cpu() - r['ligand'].pos) ** 2).sum(-1).mean().sqrt())
        assert rmsd < 0.1, (mode, rmsd)


def test_grouped_launch_equals_single_launches():
    """ddp_tpconv_umma_group over three convs that share irreps (different weights, edge sets of 1000 / 0 live / 130
    edges, one with spare capacity) == the same three convs launched one by one."""
    import ctypes as C
    from diffdock_pocket_b200 import _lib
    from diffdock_pocket_b200._lib import ptr
    ns, nv = 60, 10
    seq = _seq(ns, nv)
    L = _lib.lib()
    n = 90
    torch.manual_seed(3)
    x = torch.randn(n, E.Irreps(seq[3]).dim, device=DEV)
    jobs = []
    for j, (n_live, cap) in enumerate([(1000, 1000), (0, 256), (130, 400)]):
        conv = TensorProductConvLayer(seq[3], '1x0e+1x1o', seq[3], 3 * ns, residual=False, batch_norm=False, faster=True).to(DEV)
        pk = conv.packed(DEV, ns, ns)
        img = pk.umma_image(conv, 0, DEV)
        ei = torch.randint(0, n, (2, cap), dtype=torch.int32, device=DEV)
        emb = torch.randn(cap, ns, device=DEV)
        sh = E.spherical_harmonics('1x0e+1x1o', torch.randn(cap, 3)).to(DEV).contiguous()
        n_dev = torch.tensor([n_live], dtype=torch.int32, device=DEV)
        ed = _lib.TpEdges(emb=ptr(emb), p1=ptr(x), i1=ei[0].data_ptr(), ld1=x.shape[1], p2=ptr(x), i2=ei[1].data_ptr(), ld2=x.shape[1],
                          x=ptr(x), gather=ei[1].data_ptr(), ldx=x.shape[1], sh=ptr(sh), agg=ei[0].data_ptr(), ew=None,
                          n_edges_dev=ptr(n_dev), edge_cap=cap)
        jobs.append((conv, pk, img, ed, (ei, emb, sh, n_dev)))
    st = _lib.stream_ptr()
    single = []
    for conv, pk, img, ed, _ in jobs:
        s = torch.zeros(n, pk.spec.f_out, device=DEV)
        _lib.check(L.ddp_tpconv_umma(C.byref(pk.cdesc), ptr(img), 0, C.byref(ed), ptr(s), st), 'single')
        single.append(s)
    grouped = [torch.zeros(n, jobs[0][1].spec.f_out, device=DEV) for _ in jobs]
    k = len(jobs)
    convs = (C.c_void_p * k)(*[C.addressof(j[1].cdesc) for j in jobs])
    imgs = (C.c_void_p * k)(*[ptr(j[2]) for j in jobs])
    eds = (C.c_void_p * k)(*[C.addressof(j[3]) for j in jobs])
    sums = (C.c_void_p * k)(*[ptr(g) for g in grouped])
    _lib.check(L.ddp_tpconv_umma_group(convs, imgs, 0, eds, sums, k, st), 'group')
    torch.cuda.synchronize()
    assert float(single[0].abs().max()) > 0 and float(single[1].abs().max()) == 0
    for a, b in zip(grouped, single):
        assert T.rel_err(a, b) < 1e-5 if float(b.abs().max()) > 0 else float(a.abs().max()) == 0      # atomics order only


@pytest.mark.parametrize('mode,n,tol', [('bf16', 20, 3e-2), ('fp32', 3, 2e-3)])
def test_forward_is_se3_equivariant_at_full_size(mode, n, tol):
    """Size-independent property of the whole path on the full 3dpf apo batch (no oracle needed): a global rotation +
    translation of every coordinate rotates tr / rot scores and leaves the torsion scores unchanged."""
    from scipy.spatial.transform import Rotation
    m, c, om, oc, sa, ca = T.models(DEV)
    dl = T.randomized_list(T.graph('3dpf_apo'), n, sa, seed=4)
    Rm = torch.from_numpy(Rotation.from_rotvec([0.3, -1.1, 0.7]).as_matrix()).float()
    shift = torch.tensor([1.5, -2.0, 0.5])
    dl2 = copy.deepcopy(dl)
    for g in dl2:
        for k in ('ligand', 'atom', 'receptor'):
            g[k].pos = g[k].pos @ Rm.T + shift
    m.conv_mode = mode
    try:
        with torch.no_grad():
            a = [t.cpu() for t in m(T.batch_at(dl, 0.4))]
            b = [t.cpu() for t in m(T.batch_at(dl2, 0.4))]
    finally:
        m.conv_mode = 'fp32'
    assert T.rel_err(b[0], a[0] @ Rm.T) < tol and T.rel_err(b[1], a[1] @ Rm.T) < tol
    assert T.rel_err(b[2], a[2]) < tol and T.rel_err(b[3], a[3]) < tol
    assert a[2].numel() > 0 and a[3].numel() > 0


def test_full_model_final_poses_within_tenth_angstrom_between_modes():
    """North-star pose gate at FULL model size (README big score model, 20 steps, 3dpf apo with 7 flexible residues):
    the tensor-core modes end within 0.1 A RMSD of the fp32 CUDA-core mode (itself within 1e-4 of the oracle per
    layer) for ligand and flexible side-chain atoms, from identical initial poses and noise."""
    m, c, om, oc, sa, ca = T.models(DEV)
    g = T.graph('3dpf_apo')
    dl = T.randomized_list(g, 4, sa, seed=0)
    steps = 20
    sch = du.get_t_schedule('expbeta', steps)
    kw = dict(temp_sampling=[0.9766, 6.0774, 6.7616, 1.4488], temp_psi=[1.5103, 0.8141, 0.7662, 1.3396], temp_sigma_data=0.48884)
    out = {}
    try:
        for mode in ('fp32', 'bf16x3', 'bf16'):
            m.conv_mode = mode
            torch.manual_seed(5)
            res, _ = ps.sampling(copy.deepcopy(dl), m, steps, sch, sch, sch, sch, DEV, partial(du.t_to_sigma, args=sa), sa, batch_size=4, **kw)
            out[mode] = (torch.stack([r['ligand'].pos for r in res]), torch.stack([r['atom'].pos for r in res]))
    finally:
        m.conv_mode = 'fp32'
    flex = g['flexResidues'].subcomponents.unique()
    for mode, tol in (('bf16x3', 1e-2), ('bf16', 0.1)):
        lig = ((out[mode][0] - out['fp32'][0]) ** 2).sum(-1).mean(-1).sqrt().max()
        sc = ((out[mode][1][:, flex] - out['fp32'][1][:, flex]) ** 2).sum(-1).mean(-1).sqrt().max()
        assert float(lig) < tol and float(sc) < tol, (mode, float(lig), float(sc))
