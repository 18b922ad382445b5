"""Host-side description of the tensor products as row groups for the CUDA kernels.

A tensor-product convolution is handed to ``ddp_tpconv_*`` as a list of row groups plus a table of
coupling coefficients (``include/ddp_b200.h``, ``ddp_tp_group_t``).  This module builds that
description for

* the reference's ``FasterTensorProduct`` (models/layers.py:8-85): four weight blocks 0e, 1o, 1e, 0o
  whose rows are the concatenated bases of layers.py:40-53 and whose weights are scaled by
  ``1/sqrt(in_k)`` (layers.py:59-60);
* e3nn 0.5.1 ``FullyConnectedTensorProduct(shared_weights=False)`` (models/score_model.py:98):
  instructions enumerated ``for i1 for i2 for io``, weights ``[mul1, mul2, mul_out]``, coefficient
  ``sqrt((2 l_out + 1) / fan_in) * wigner_3j``.

Irreps are lists of ``(mul, l, p)``.  The real Clebsch-Gordan tensors follow e3nn's convention
(real basis change with the (-i)^l phase, Frobenius norm 1).
"""
import math
import re

import numpy as np


def parse_irreps(spec):
    if not isinstance(spec, str):
        return [tuple(x) for x in spec]
    out = []
    for tok in spec.split('+'):
        m = re.fullmatch(r'\s*(?:(\d+)x)?(\d+)([eo])\s*', tok)
        out.append((int(m.group(1)) if m.group(1) else 1, int(m.group(2)), 1 if m.group(3) == 'e' else -1))
    return out


def irreps_dim(irreps):
    return sum(m * (2 * l + 1) for m, l, _ in irreps)


def irreps_offsets(irreps):
    off, s = [], 0
    for m, l, _ in irreps:
        off.append(s)
        s += m * (2 * l + 1)
    return off


# ------------------------------------------------------------------------------------------- CG
def _su2_cg(j1, j2, j3):
    f = math.factorial
    C = np.zeros((2 * j1 + 1, 2 * j2 + 1, 2 * j3 + 1))
    for m1 in range(-j1, j1 + 1):
        for m2 in range(-j2, j2 + 1):
            m3 = m1 + m2
            if abs(m3) > j3:
                continue
            vmin = max(-j1 + j2 + m3, -j1 + m1, 0)
            vmax = min(j2 + j3 + m1, j3 - j1 + j2, j3 + m3)
            pref = math.sqrt((2 * j3 + 1) * f(j3 + j1 - j2) * f(j3 - j1 + j2) * f(j1 + j2 - j3) / f(j1 + j2 + j3 + 1)
                             * f(j3 + m3) * f(j3 - m3) / (f(j1 + m1) * f(j1 - m1) * f(j2 + m2) * f(j2 - m2)))
            s = sum((-1.0) ** (v + j2 + m2) / f(v) * f(j2 + j3 + m1 - v) * f(j1 - m1 + v)
                    / f(j3 - j1 + j2 - v) / f(j3 + m3 - v) / f(v + j1 - j2 - m3) for v in range(vmin, vmax + 1))
            C[j1 + m1, j2 + m2, j3 + m3] = pref * s
    return C


def _real_to_complex(l):
    q = np.zeros((2 * l + 1, 2 * l + 1), dtype=np.complex128)
    r = 1 / math.sqrt(2)
    for m in range(-l, 0):
        q[l + m, l - m] = r
        q[l + m, l + m] = -1j * r
    q[l, l] = 1
    for m in range(1, l + 1):
        q[l + m, l + m] = (-1) ** m * r
        q[l + m, l - m] = 1j * (-1) ** m * r
    return (-1j) ** l * q


def wigner_3j(l1, l2, l3):
    if not abs(l1 - l2) <= l3 <= l1 + l2:
        return np.zeros((2 * l1 + 1, 2 * l2 + 1, 2 * l3 + 1))
    C = np.einsum('ij,kl,mn,ikn->jlm', _real_to_complex(l1), _real_to_complex(l2),
                  np.conj(_real_to_complex(l3).T), _su2_cg(l1, l2, l3).astype(np.complex128))
    assert np.abs(C.imag).max() < 1e-9
    C = C.real
    return C / np.linalg.norm(C)


# widths the tensor-core kernel is instantiated for (csrc/tpconv_umma.cu pick_cfg): ns -> nv
UMMA_WIDTHS = {60: 10, 24: 6, 16: 4}


def umma_supported(spec, ns):
    """True if ``ddp_tpconv_pack`` accepts this product: an l <= 1 row-group spec whose scalar / vector output
    multiplicities are one of the instantiated (ns, nv) pairs.  Anything else runs on the fp32 CUDA-core kernel."""
    if not spec.tc_eligible or ns not in UMMA_WIDTHS:
        return False
    for m, l, _ in spec.out_irreps:
        if m != (ns if l == 0 else UMMA_WIDTHS[ns]):
            return False
    return True


# ------------------------------------------------------------------------------------------- groups
class TpSpec:
    """Row-group description of one tensor product."""

    def __init__(self, in_irreps, sh_dim, out_irreps):
        self.in_irreps, self.out_irreps = in_irreps, out_irreps
        self.f_in, self.f_out, self.sh_dim = irreps_dim(in_irreps), irreps_dim(out_irreps), sh_dim
        self.groups = []          # dicts with the ddp_tp_group_t fields
        self.ctab = []            # flat float list
        self.weight_numel = 0
        self.faster = False       # FasterTensorProduct-shaped
        self.tc_eligible = False  # candidate for the tensor-core kernel (l <= 1 scalar / vector row groups)

    def add(self, C, x_off, mul_in, sh_off, w_off, out_off, mul_out):
        d1, d2, do = C.shape
        self.groups.append(dict(d1=d1, d2=d2, d_out=do, x_off=x_off, mul_in=mul_in, sh_off=sh_off, w_off=w_off,
                                out_off=out_off, mul_out=mul_out, c_off=len(self.ctab)))
        self.ctab += [float(v) for v in C.reshape(-1)]

    def col_group(self):
        cg = np.zeros(self.weight_numel, dtype=np.uint8)
        for gi, g in enumerate(self.groups):
            cg[g['w_off']:g['w_off'] + g['mul_in'] * g['mul_out']] = gi
        return cg


_EPS = np.zeros((3, 3, 3))
for _i, _j, _k in ((0, 1, 2), (1, 2, 0), (2, 0, 1)):
    _EPS[_i, _j, _k] = 1.0
    _EPS[_j, _i, _k] = -1.0


def faster_tp_spec(in_irreps, out_irreps):
    """models/layers.py:8-85 as row groups (sh = [s0 | s1(3)], sh_dim 4)."""
    in_irreps, out_irreps = parse_irreps(in_irreps), parse_irreps(out_irreps)
    name = lambda l, p: f"{l}{'e' if p == 1 else 'o'}"
    im = {'0e': 0, '1o': 0, '1e': 0, '0o': 0}
    om = dict(im)
    ioff, ooff = {}, {}
    for (m, l, p), o in zip(in_irreps, irreps_offsets(in_irreps)):
        assert l <= 1, 'FasterTensorProduct handles l <= 1 only'
        im[name(l, p)], ioff[name(l, p)] = m, o
    for (m, l, p), o in zip(out_irreps, irreps_offsets(out_irreps)):
        assert l <= 1
        om[name(l, p)], ooff[name(l, p)] = m, o
    one = np.ones((1, 1, 1))
    dot = (np.eye(3) / math.sqrt(3)).reshape(3, 3, 1)
    s_x_v = np.eye(3).reshape(1, 3, 3)            # scalar (x) s1 -> vector
    v_x_s = np.eye(3).reshape(3, 1, 3)            # vector * s0   -> vector
    cross = _EPS / math.sqrt(2)
    # (input irrep, coupling tensor, sh offset) in the concatenation order of layers.py:40-53
    rows = {'0e': [('0e', one, 0), ('1o', dot, 1)],
            '1o': [('0e', s_x_v, 1), ('1o', v_x_s, 0), ('1e', cross, 1)],
            '1e': [('1o', cross, 1), ('1e', v_x_s, 0), ('0o', s_x_v, 1)],
            '0o': [('1e', dot, 1), ('0o', one, 0)]}
    spec = TpSpec(in_irreps, 4, out_irreps)
    spec.faster = True
    spec.tc_eligible = True
    w = 0
    for key in ('0e', '1o', '1e', '0o'):                         # layers.py:26-31
        in_k = sum(im[r[0]] for r in rows[key])
        if in_k * om[key] == 0:
            continue
        scale = 1.0 / math.sqrt(in_k)
        for (src, Cg, sh_off) in rows[key]:
            if im[src] == 0:
                continue
            spec.add(Cg * scale, ioff[src], im[src], sh_off, w, ooff[key], om[key])
            w += im[src] * om[key]
    spec.weight_numel = w
    spec.blocks = {k: (sum(im[r[0]] for r in rows[k]), om[k]) for k in rows}
    return spec


def fctp_spec(in_irreps, sh_irreps, out_irreps, sh_keep=None, sh_base=0):
    """e3nn FullyConnectedTensorProduct (mode uvw, mul2 == 1).  ``sh_keep``: optional list of sh irrep
    indices whose components are actually supplied (others must be unused by every instruction)."""
    in_irreps, sh_irreps, out_irreps = parse_irreps(in_irreps), parse_irreps(sh_irreps), parse_irreps(out_irreps)
    assert all(m == 1 for m, _, _ in sh_irreps)
    ins = [(i1, i2, io) for i1, (_, l1, p1) in enumerate(in_irreps) for i2, (_, l2, p2) in enumerate(sh_irreps)
           for io, (_, lo, po) in enumerate(out_irreps) if abs(l1 - l2) <= lo <= l1 + l2 and po == p1 * p2]
    fan = {}
    for i1, i2, io in ins:
        fan[io] = fan.get(io, 0) + in_irreps[i1][0]
    sh_off_full = irreps_offsets(sh_irreps)
    if sh_keep is None:
        sh_off, sh_dim = sh_off_full, irreps_dim(sh_irreps)
    else:
        sh_off, s = {}, sh_base
        for k in sh_keep:
            sh_off[k] = s
            s += 2 * sh_irreps[k][1] + 1
        sh_dim = s
    xo, oo = irreps_offsets(in_irreps), irreps_offsets(out_irreps)
    spec = TpSpec(in_irreps, sh_dim, out_irreps)
    w = 0
    for i1, i2, io in ins:
        m1, l1, _ = in_irreps[i1]
        _, l2, _ = sh_irreps[i2]
        mo, lo, _ = out_irreps[io]
        if sh_keep is not None and i2 not in sh_off:
            raise NotImplementedError('instruction uses an sh irrep that is not supplied')
        coeff = math.sqrt((2 * lo + 1) / fan[io])
        spec.add(wigner_3j(l1, l2, lo) * coeff, xo[i1], m1, sh_off[i2], w, oo[io], mo)
        w += m1 * mo
    spec.weight_numel = w
    # tensor-core candidates: node irreps with l <= 1 and the plain spherical harmonics up to l = 1 or 2 as second operand --
    # every instruction is then one of the six row kinds of csrc/tpconv_umma.cu (x s0, x1.s1, x (x) s1, x1 s0, x1 x s1, and,
    # for sh_lmax = 2, C(1,2,1)(x1, s2))
    plain_sh = [tuple(t) for t in sh_irreps] in ([(1, 0, 1), (1, 1, -1)], [(1, 0, 1), (1, 1, -1), (1, 2, 1)])
    spec.tc_eligible = bool(sh_keep is None and plain_sh and all(l <= 1 for _, l, _ in in_irreps) and all(l <= 1 for _, l, _ in out_irreps))
    return spec


def full_tp_out_irreps(ir1, ir2):
    """Output irreps of e3nn FullTensorProduct, sorted by (l, p) with odd parity first."""
    ir1, ir2 = parse_irreps(ir1), parse_irreps(ir2)
    outs = [(m1 * m2, lo, p1 * p2) for (m1, l1, p1) in ir1 for (m2, l2, p2) in ir2
            for lo in range(abs(l1 - l2), l1 + l2 + 1)]
    return sorted(outs, key=lambda t: (t[1], t[2]))


def full_tp_paths(ir1, ir2):
    """FullTensorProduct outputs in e3nn's sorted order with their provenance: [(l_out, p_out, i1, i2)]."""
    ir1, ir2 = parse_irreps(ir1), parse_irreps(ir2)
    outs = [(lo, p1 * p2, a, b) for a, (m1, l1, p1) in enumerate(ir1) for b, (m2, l2, p2) in enumerate(ir2)
            for lo in range(abs(l1 - l2), l1 + l2 + 1)]
    return sorted(outs, key=lambda t: (t[0], t[1]))


def fctp_used_sh(in_irreps, sh_irreps, out_irreps):
    """Indices of the sh irreps that some FullyConnectedTensorProduct instruction reads."""
    in_irreps, sh_irreps, out_irreps = parse_irreps(in_irreps), parse_irreps(sh_irreps), parse_irreps(out_irreps)
    return sorted({i2 for (_, l1, p1) in in_irreps for i2, (_, l2, p2) in enumerate(sh_irreps)
                   for (_, lo, po) in out_irreps if abs(l1 - l2) <= lo <= l1 + l2 and po == p1 * p2})


def batch_norm_fold(irreps, running_mean, running_var, weight, bias, eps=1e-5):
    """e3nn.nn.BatchNorm (eval) as per-channel scale / shift over the flattened feature vector:
    y = x * scale + shift with shift != 0 only on 0e channels."""
    scale, shift = [], []
    irm = irv = ib = 0
    for m, l, p in irreps:
        d = 2 * l + 1
        sc = weight[irv:irv + m] / np.sqrt(running_var[irv:irv + m] + eps)
        irv += m
        if l == 0 and p == 1:
            sh = bias[ib:ib + m] - running_mean[irm:irm + m] * sc
            ib += m
            irm += m
        else:
            sh = np.zeros(m)
        scale.append(np.repeat(sc, d))
        shift.append(np.repeat(sh, d))
    return np.concatenate(scale).astype(np.float32), np.concatenate(shift).astype(np.float32)
