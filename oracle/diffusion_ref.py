"""CPU restatement of the sampler-side math of the reference (TEST INFRASTRUCTURE).

* ``t_to_sigma`` / ``sinusoidal_embedding`` / ``get_t_schedule`` / ``set_time``
                                  <- utils/diffusion_utils.py:22-34, 73-84, 112-117, 124-165
* ``axis_angle_to_matrix`` / ``kabsch`` <- utils/geometry.py:7-86, 209-243
* ``modify_conformer_torsion_angles`` / ``modify_sidechain_torsion_angle``
                                  <- utils/torsion.py:68-94, 251-278
* ``modify_conformer`` / ``modify_sidechains`` <- utils/diffusion_utils.py:37-70
* ``so3_exp_score_norms`` / ``so3_score_norm``     <- utils/so3.py:15-60, 85-89
* ``torus_tables`` / ``torus_score_norm_table``    <- utils/torus.py:11-82
"""
import copy
import math

import numpy as np
import torch
from scipy.spatial.transform import Rotation as R
from scipy.stats import beta


# ----------------------------------------------------------------------------- schedules
def t_to_sigma_individual(t, sigma_min, sigma_max):
    return sigma_min ** (1 - t) * sigma_max ** t


def t_to_sigma(t_tr, t_rot, t_tor, t_sc_tor, args):
    return (t_to_sigma_individual(t_tr, args.tr_sigma_min, args.tr_sigma_max),
            t_to_sigma_individual(t_rot, args.rot_sigma_min, args.rot_sigma_max),
            t_to_sigma_individual(t_tor, args.tor_sigma_min, args.tor_sigma_max),
            t_to_sigma_individual(t_sc_tor, args.sidechain_tor_sigma_min, args.sidechain_tor_sigma_max))


def sinusoidal_embedding(timesteps, dim, scale=1.0, max_positions=10000):
    half = dim // 2
    emb = math.log(max_positions) / (half - 1)
    emb = torch.exp(torch.arange(half, dtype=torch.float32) * -emb)
    emb = scale * timesteps.float()[:, None] * emb[None, :]
    emb = torch.cat([torch.sin(emb), torch.cos(emb)], dim=1)
    if dim % 2 == 1:
        emb = torch.nn.functional.pad(emb, (0, 1))
    return emb


def get_t_schedule(inference_steps, alpha=1, beta_=1, t_max=1):
    lin_max = beta.cdf(t_max, a=alpha, b=beta_)
    c = np.linspace(lin_max, 0, inference_steps + 1)[:-1]
    return beta.ppf(c, a=alpha, b=beta_)


def set_time(g, t_tr, t_rot, t_tor, t_sc, batchsize, t=None):
    """utils/diffusion_utils.py:124-165; ``t`` is given under the asynchronous noise schedule (:158-165)."""
    for key in ('ligand', 'receptor', 'atom'):
        n = g[key].num_nodes
        g[key].node_t = {'tr': t_tr * torch.ones(n), 'rot': t_rot * torch.ones(n),
                         'tor': t_tor * torch.ones(n), 'sc_tor': t_sc * torch.ones(n)}
        if t is not None:
            g[key].node_t['t'] = t * torch.ones(n)
    g.complex_t = {'tr': t_tr * torch.ones(batchsize), 'rot': t_rot * torch.ones(batchsize),
                   'tor': t_tor * torch.ones(batchsize), 'sc_tor': t_sc * torch.ones(batchsize)}
    if t is not None:
        g.complex_t['t'] = t * torch.ones(batchsize)


# ----------------------------------------------------------------------------- geometry
def axis_angle_to_matrix(axis_angle):
    angles = torch.norm(axis_angle, p=2, dim=-1, keepdim=True)
    half = 0.5 * angles
    small = angles.abs() < 1e-6
    s = torch.empty_like(angles)
    s[~small] = torch.sin(half[~small]) / angles[~small]
    s[small] = 0.5 - (angles[small] * angles[small]) / 48
    q = torch.cat([torch.cos(half), axis_angle * s], dim=-1)
    r, i, j, k = torch.unbind(q, -1)
    two_s = 2.0 / (q * q).sum(-1)
    o = torch.stack((1 - two_s * (j * j + k * k), two_s * (i * j - k * r), two_s * (i * k + j * r),
                     two_s * (i * j + k * r), 1 - two_s * (i * i + k * k), two_s * (j * k - i * r),
                     two_s * (i * k - j * r), two_s * (j * k + i * r), 1 - two_s * (i * i + j * j)), -1)
    return o.reshape(q.shape[:-1] + (3, 3))


def kabsch(A, B):
    """R, t with R @ A + t ~ B for 3xN inputs (utils/geometry.py:209-243)."""
    ca, cb = A.mean(1, keepdim=True), B.mean(1, keepdim=True)
    H = (A - ca) @ (B - cb).T
    U, S, Vt = torch.linalg.svd(H)
    Rm = Vt.T @ U.T
    if torch.linalg.det(Rm) < 0:
        Rm = (Vt.T @ torch.diag(torch.tensor([1., 1., -1.]))) @ U.T
    assert math.fabs(torch.linalg.det(Rm) - 1) < 3e-3
    return Rm, -Rm @ ca + cb


# ----------------------------------------------------------------------------- pose updates
def modify_conformer_torsion_angles(pos, edge_index, mask_rotate, torsion_updates):
    pos = copy.deepcopy(pos).numpy() if torch.is_tensor(pos) else copy.deepcopy(pos)
    for k, e in enumerate(np.asarray(edge_index)):
        if torsion_updates[k] == 0:
            continue
        u, v = e[0], e[1]
        assert not mask_rotate[k, u] and mask_rotate[k, v]
        rot_vec = pos[u] - pos[v]
        rot_vec = rot_vec * torsion_updates[k] / np.linalg.norm(rot_vec)
        rot_mat = R.from_rotvec(rot_vec).as_matrix()
        pos[mask_rotate[k]] = (pos[mask_rotate[k]] - pos[v]) @ rot_mat.T + pos[v]
    return torch.from_numpy(pos.astype(np.float32))


def modify_sidechain_torsion_angle(pos, edge_index, mask_subcomponent, subcomponents, torsion_update):
    pos = copy.deepcopy(pos).numpy()
    if torsion_update != 0:
        u, v = int(edge_index[0]), int(edge_index[1])
        mask = np.asarray(subcomponents[int(mask_subcomponent[0]):int(mask_subcomponent[1])])
        rot_vec = pos[u] - pos[v]
        rot_vec = rot_vec * torsion_update / np.linalg.norm(rot_vec)
        rot_mat = R.from_rotvec(rot_vec).as_matrix()
        pos[mask] = (pos[mask] - pos[v]) @ rot_mat.T + pos[v]
    return torch.from_numpy(pos.astype(np.float32))


def modify_sidechains(data, torsion_updates):
    fr = data['flexResidues']
    for i, upd in enumerate(torsion_updates):
        data['atom'].pos = modify_sidechain_torsion_angle(data['atom'].pos, fr.edge_idx[i], fr.subcomponentsMapping[i],
                                                          fr.subcomponents, upd)


def modify_conformer(data, tr_update, rot_update, torsion_updates):
    lig = data['ligand']
    center = torch.mean(lig.pos, dim=0, keepdim=True)
    rot_mat = axis_angle_to_matrix(rot_update.squeeze())
    rigid = (lig.pos - center) @ rot_mat.T + tr_update + center
    if torsion_updates is not None:
        mr = lig.mask_rotate if isinstance(lig.mask_rotate, np.ndarray) else lig.mask_rotate[0]
        flex = modify_conformer_torsion_angles(rigid, data['ligand', 'ligand'].edge_index.T[lig.edge_mask], mr, torsion_updates)
        Rm, t = kabsch(flex.T, rigid.T)
        lig.pos = flex @ Rm.T + t.T
    else:
        lig.pos = rigid
    return data


# ----------------------------------------------------------------------------- SO(3) table
SO3_MIN_EPS, SO3_MAX_EPS, SO3_N_EPS, SO3_X_N = 0.01, 2, 1000, 2000


def so3_exp_score_norms(eps_indices=None, L=2000):
    """utils/so3.py:15-60 restated literally (loop over l); optionally only for some eps indices."""
    eps_array = 10 ** np.linspace(np.log10(SO3_MIN_EPS), np.log10(SO3_MAX_EPS), SO3_N_EPS)
    if eps_indices is not None:
        eps_array = eps_array[np.asarray(eps_indices)]
    omegas = np.linspace(0, np.pi, SO3_X_N + 1)[1:]
    out = []
    for eps in eps_array:
        p = 0
        for l in range(L):
            p += (2 * l + 1) * np.exp(-l * (l + 1) * eps ** 2) * np.sin(omegas * (l + 1 / 2)) / np.sin(omegas / 2)
        pdf = p * (1 - np.cos(omegas)) / np.pi
        d = 0
        for l in range(L):
            hi = np.sin(omegas * (l + 1 / 2))
            dhi = (l + 1 / 2) * np.cos(omegas * (l + 1 / 2))
            lo = np.sin(omegas / 2)
            dlo = 1 / 2 * np.cos(omegas / 2)
            d += (2 * l + 1) * np.exp(-l * (l + 1) * eps ** 2) * (lo * dhi - hi * dlo) / lo ** 2
        score = d / p
        out.append(np.sqrt(np.sum(score ** 2 * pdf) / np.sum(pdf) / np.pi))
    return np.asarray(out)


def so3_eps_index(eps):
    idx = (np.log10(eps) - np.log10(SO3_MIN_EPS)) / (np.log10(SO3_MAX_EPS) - np.log10(SO3_MIN_EPS)) * SO3_N_EPS
    return np.clip(np.around(idx).astype(int), a_min=0, a_max=SO3_N_EPS - 1)


# ----------------------------------------------------------------------------- torus table
TORUS_X_MIN, TORUS_X_N = 1e-5, 5000
TORUS_SIGMA_MIN, TORUS_SIGMA_MAX, TORUS_SIGMA_N = 3e-3, 2, 5000


def torus_sigma_index(sigma):
    s = np.log(sigma / np.pi)
    s = (s - np.log(TORUS_SIGMA_MIN)) / (np.log(TORUS_SIGMA_MAX) - np.log(TORUS_SIGMA_MIN)) * TORUS_SIGMA_N
    return np.round(np.clip(s, 0, TORUS_SIGMA_N)).astype(int)


def torus_score_norm_table(seed, n_mc=10000, sigma_stride=1):
    """utils/torus.py:25-75: tabulate p and grad/p on the (sigma, x) grids with N=100 images, then the
    Monte-Carlo mean of score**2 over ``n_mc`` wrapped-normal samples per sigma.  The reference
    draws from the unseeded global numpy RNG; here the generator is seeded."""
    x = 10 ** np.linspace(np.log10(TORUS_X_MIN), 0, TORUS_X_N + 1) * np.pi
    sigma = (10 ** np.linspace(np.log10(TORUS_SIGMA_MIN), np.log10(TORUS_SIGMA_MAX), TORUS_SIGMA_N + 1) * np.pi)
    sel = np.arange(0, TORUS_SIGMA_N + 1, sigma_stride)
    sg = sigma[sel][:, None]
    p_ = np.zeros((len(sel), TORUS_X_N + 1))
    g_ = np.zeros_like(p_)
    for i in range(-100, 101):
        e = np.exp(-(x + 2 * np.pi * i) ** 2 / 2 / sg ** 2)
        p_ += e
        g_ += (x + 2 * np.pi * i) / sg ** 2 * e
    score_ = g_ / p_
    rng = np.random.RandomState(seed)
    out = np.zeros(len(sel))
    for k in range(len(sel)):
        smp = sigma[sel[k]] * rng.randn(n_mc)
        smp = (smp + np.pi) % (2 * np.pi) - np.pi
        sign = np.sign(smp)
        xi = np.log(np.abs(smp) / np.pi)
        xi = (xi - np.log(TORUS_X_MIN)) / (0 - np.log(TORUS_X_MIN)) * TORUS_X_N
        xi = np.round(np.clip(xi, 0, TORUS_X_N)).astype(int)
        sc = -sign * score_[k, xi]
        out[k] = (sc ** 2).mean()
    return sel, out
