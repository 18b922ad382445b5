"""Minimal restatement of the e3nn 0.5.1 pieces the reference hot path uses.

TEST INFRASTRUCTURE (see oracle/__init__.py).  e3nn is an un-vendored
dependency of the reference (environment.yml:31, ``e3nn==0.5.1``); it is not
installed here, so its published algorithms are restated:

* ``spherical_harmonics``  <- ``o3.spherical_harmonics(..., normalize=True,
  normalization='component')`` (call sites models/all_atom_score_model.py:394,
  418,481,508,534,556,570,579,598,613,633)
* ``wigner_3j``            <- ``o3.wigner_3j`` (real basis, Frobenius norm 1)
* ``FullyConnectedTensorProduct`` <- models/score_model.py:98 (``shared_weights=False``)
* ``FullTensorProduct``    <- models/all_atom_score_model.py:193,219
* ``BatchNorm``            <- ``e3nn.nn.BatchNorm`` (models/score_model.py:106), eval mode
"""
import math
import re
from typing import List, Tuple

import torch


# --------------------------------------------------------------------------- irreps
class Irreps:
    """List of (mul, l, p) with p in {+1, -1}; parses '60x0e + 10x1o'."""

    def __init__(self, spec):
        if isinstance(spec, Irreps):
            self.items = list(spec.items)
            return
        if isinstance(spec, (list, tuple)):
            self.items = [tuple(x) for x in spec]
            return
        items = []
        for tok in str(spec).split('+'):
            tok = tok.strip()
            if not tok:
                continue
            m = re.fullmatch(r'(?:(\d+)x)?(\d+)([eo])', tok)
            if m is None:
                raise ValueError(f'bad irrep {tok!r}')
            mul = int(m.group(1)) if m.group(1) else 1
            items.append((mul, int(m.group(2)), 1 if m.group(3) == 'e' else -1))
        self.items = items

    @staticmethod
    def spherical_harmonics(lmax):
        return Irreps([(1, l, (-1) ** l) for l in range(lmax + 1)])

    def __iter__(self):
        return iter(self.items)

    def __len__(self):
        return len(self.items)

    def __getitem__(self, i):
        return self.items[i]

    def __eq__(self, other):
        return self.items == Irreps(other).items

    @property
    def dim(self):
        return sum(mul * (2 * l + 1) for mul, l, _ in self.items)

    def slices(self):
        out, s = [], 0
        for mul, l, _ in self.items:
            out.append(slice(s, s + mul * (2 * l + 1)))
            s += mul * (2 * l + 1)
        return out

    def __repr__(self):
        return '+'.join(f"{mul}x{l}{'e' if p == 1 else 'o'}" for mul, l, p in self.items)


def ir_name(l, p):
    return f"{l}{'e' if p == 1 else 'o'}"


# --------------------------------------------------------------------------- spherical harmonics
def spherical_harmonics(irreps, vec, normalize=True, normalization='component'):
    """e3nn convention: l=1 basis is (x, y, z); y is the polar axis for l=2."""
    assert normalize and normalization == 'component'
    if isinstance(irreps, int):
        ls = [irreps]
    elif isinstance(irreps, str):
        ls = [l for _, l, _ in Irreps(irreps)]
    else:
        ls = [l for _, l, _ in Irreps(irreps)]
    v = torch.nn.functional.normalize(vec, dim=-1)  # zero vector -> 0
    x, y, z = v[..., 0], v[..., 1], v[..., 2]
    out = []
    for l in ls:
        if l == 0:
            out.append(torch.ones_like(x).unsqueeze(-1))
        elif l == 1:
            out.append(math.sqrt(3.0) * torch.stack([x, y, z], -1))
        elif l == 2:
            s3 = math.sqrt(3.0)
            raw = torch.stack([s3 * x * z, s3 * x * y, y * y - 0.5 * (x * x + z * z),
                               s3 * y * z, (s3 / 2.0) * (z * z - x * x)], -1)
            out.append(math.sqrt(5.0) * raw)
        else:
            raise NotImplementedError('l > 2 not used by the reference path')
    return torch.cat(out, -1)


# --------------------------------------------------------------------------- Clebsch-Gordan
def _su2_cg_coeff(j1, m1, j2, m2, j3, m3):
    if m3 != m1 + m2:
        return 0.0
    f = math.factorial
    vmin = int(max(-j1 + j2 + m3, -j1 + m1, 0))
    vmax = int(min(j2 + j3 + m1, j3 - j1 + j2, j3 + m3))
    C = math.sqrt((2.0 * j3 + 1.0) * f(j3 + j1 - j2) * f(j3 - j1 + j2) * f(j1 + j2 - j3) / f(j1 + j2 + j3 + 1)
                  * f(j3 + m3) * f(j3 - m3) / (f(j1 + m1) * f(j1 - m1) * f(j2 + m2) * f(j2 - m2)))
    S = 0.0
    for v in range(vmin, vmax + 1):
        S += (-1.0) ** (v + j2 + m2) / f(v) * f(j2 + j3 + m1 - v) * f(j1 - m1 + v) \
            / f(j3 - j1 + j2 - v) / f(j3 + m3 - v) / f(v + j1 - j2 - m3)
    return C * S


def _su2_cg(j1, j2, j3):
    mat = torch.zeros((2 * j1 + 1, 2 * j2 + 1, 2 * j3 + 1), dtype=torch.float64)
    if abs(j1 - j2) <= j3 <= j1 + j2:
        for m1 in range(-j1, j1 + 1):
            for m2 in range(-j2, j2 + 1):
                if abs(m1 + m2) <= j3:
                    mat[j1 + m1, j2 + m2, j3 + m1 + m2] = _su2_cg_coeff(j1, m1, j2, m2, j3, m1 + m2)
    return mat


def _real_to_complex(l):
    q = torch.zeros((2 * l + 1, 2 * l + 1), dtype=torch.complex128)
    for m in range(-l, 0):
        q[l + m, l + abs(m)] = 1 / math.sqrt(2)
        q[l + m, l - abs(m)] = -1j / math.sqrt(2)
    q[l, l] = 1
    for m in range(1, l + 1):
        q[l + m, l + abs(m)] = (-1) ** m / math.sqrt(2)
        q[l + m, l - abs(m)] = 1j * (-1) ** m / math.sqrt(2)
    return (-1j) ** l * q


_W3J_CACHE = {}


def wigner_3j(l1, l2, l3):
    """Real-basis Clebsch-Gordan tensor [2l1+1, 2l2+1, 2l3+1], Frobenius norm 1."""
    key = (l1, l2, l3)
    if key not in _W3J_CACHE:
        Q1, Q2, Q3 = _real_to_complex(l1), _real_to_complex(l2), _real_to_complex(l3)
        C = _su2_cg(l1, l2, l3).to(torch.complex128)
        C = torch.einsum('ij,kl,mn,ikn->jlm', Q1, Q2, torch.conj(Q3.T), C)
        assert torch.all(torch.abs(C.imag) < 1e-9)
        C = C.real
        _W3J_CACHE[key] = (C / C.norm()).contiguous()
    return _W3J_CACHE[key]


# --------------------------------------------------------------------------- tensor products
class FullyConnectedTensorProduct(torch.nn.Module):
    """mode 'uvw', per-edge weights (shared_weights=False), component / element normalisation."""

    def __init__(self, irreps_in1, irreps_in2, irreps_out):
        super().__init__()
        self.irreps_in1, self.irreps_in2, self.irreps_out = Irreps(irreps_in1), Irreps(irreps_in2), Irreps(irreps_out)
        ins = []
        for i1, (m1, l1, p1) in enumerate(self.irreps_in1):
            for i2, (m2, l2, p2) in enumerate(self.irreps_in2):
                for io, (mo, lo, po) in enumerate(self.irreps_out):
                    if abs(l1 - l2) <= lo <= l1 + l2 and po == p1 * p2:
                        ins.append((i1, i2, io))
        self.instructions = ins
        fan = {}
        for (i1, i2, io) in ins:
            fan[io] = fan.get(io, 0) + self.irreps_in1[i1][0] * self.irreps_in2[i2][0]
        self.path_coeff = [math.sqrt((2 * self.irreps_out[io][1] + 1) / fan[io]) for (_, _, io) in ins]
        self.weight_numel = sum(self.irreps_in1[i1][0] * self.irreps_in2[i2][0] * self.irreps_out[io][0]
                                for (i1, i2, io) in ins)

    def forward(self, x1, x2, weight):
        E = x1.shape[0]
        s1, s2, so = self.irreps_in1.slices(), self.irreps_in2.slices(), self.irreps_out.slices()
        outs = [torch.zeros(E, mo, 2 * lo + 1, dtype=x1.dtype) for (mo, lo, _) in self.irreps_out]
        off = 0
        for (i1, i2, io), c in zip(self.instructions, self.path_coeff):
            m1, l1, _ = self.irreps_in1[i1]
            m2, l2, _ = self.irreps_in2[i2]
            mo, lo, _ = self.irreps_out[io]
            n = m1 * m2 * mo
            w = weight[:, off:off + n].reshape(E, m1, m2, mo)
            off += n
            a = x1[:, s1[i1]].reshape(E, m1, 2 * l1 + 1)
            b = x2[:, s2[i2]].reshape(E, m2, 2 * l2 + 1)
            C = wigner_3j(l1, l2, lo).to(x1.dtype)
            outs[io] = outs[io] + c * torch.einsum('zuvw,ijk,zui,zvj->zwk', w, C, a, b)
        return torch.cat([o.reshape(E, -1) for o in outs], -1)


class FullTensorProduct(torch.nn.Module):
    """mode 'uvuv', no weights, outputs sorted by (l, p) with odd parity first (stable)."""

    def __init__(self, irreps_in1, irreps_in2):
        super().__init__()
        self.irreps_in1, self.irreps_in2 = Irreps(irreps_in1), Irreps(irreps_in2)
        outs = []
        for i1, (m1, l1, p1) in enumerate(self.irreps_in1):
            for i2, (m2, l2, p2) in enumerate(self.irreps_in2):
                for lo in range(abs(l1 - l2), l1 + l2 + 1):
                    outs.append((m1 * m2, lo, p1 * p2, i1, i2))
        order = sorted(range(len(outs)), key=lambda k: (outs[k][1], outs[k][2]))
        self.paths = [outs[k] for k in order]
        self.irreps_out = Irreps([(m, l, p) for (m, l, p, _, _) in self.paths])

    def forward(self, x1, x2):
        E = x1.shape[0]
        s1, s2 = self.irreps_in1.slices(), self.irreps_in2.slices()
        res = []
        for (m, lo, po, i1, i2) in self.paths:
            m1, l1, _ = self.irreps_in1[i1]
            m2, l2, _ = self.irreps_in2[i2]
            a = x1[:, s1[i1]].reshape(E, m1, 2 * l1 + 1)
            b = x2[:, s2[i2]].reshape(E, m2, 2 * l2 + 1)
            C = wigner_3j(l1, l2, lo).to(x1.dtype) * math.sqrt(2 * lo + 1)
            res.append(torch.einsum('ijk,zui,zvj->zuvk', C, a, b).reshape(E, -1))
        return torch.cat(res, -1)


class BatchNorm(torch.nn.Module):
    """e3nn.nn.BatchNorm, eval mode, affine, normalization='component', eps=1e-5."""

    def __init__(self, irreps, eps=1e-5):
        super().__init__()
        self.irreps = Irreps(irreps)
        self.eps = eps
        n_scalar = sum(mul for mul, l, p in self.irreps if l == 0 and p == 1)
        n_feat = sum(mul for mul, _, _ in self.irreps)
        self.register_buffer('running_mean', torch.zeros(n_scalar))
        self.register_buffer('running_var', torch.ones(n_feat))
        self.weight = torch.nn.Parameter(torch.ones(n_feat))
        self.bias = torch.nn.Parameter(torch.zeros(n_scalar))

    def forward(self, x):
        n = x.shape[0]
        ix = irm = irv = iw = ib = 0
        fields = []
        for mul, l, p in self.irreps:
            d = 2 * l + 1
            f = x[:, ix:ix + mul * d].reshape(n, mul, d)
            ix += mul * d
            scalar = (l == 0 and p == 1)
            if scalar:
                f = f - self.running_mean[irm:irm + mul].reshape(1, mul, 1)
                irm += mul
            fn = (self.running_var[irv:irv + mul] + self.eps).pow(-0.5) * self.weight[iw:iw + mul]
            irv += mul
            iw += mul
            f = f * fn.reshape(1, mul, 1)
            if scalar:
                f = f + self.bias[ib:ib + mul].reshape(1, mul, 1)
                ib += mul
            fields.append(f.reshape(n, mul * d))
        return torch.cat(fields, -1)
