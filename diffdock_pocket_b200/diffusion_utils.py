"""Host mirror of ``utils/diffusion_utils.py`` (schedules, time embedding, set_time, pose updates).

The schedule / embedding helpers are plain host math (they produce a handful of scalars per step).
``modify_conformer`` / ``modify_sidechains`` keep the reference signatures
(utils/diffusion_utils.py:37-70) but run the fused ``ddp_pose_update`` kernel; ``PoseState`` is the
resident multi-sample form used by ``sampling()``.
"""
import ctypes as C
import functools
import math

import numpy as np
import torch
from scipy.stats import beta

from . import _lib
from ._lib import ptr


def t_to_sigma_individual(t, schedule_type, sigma_min, sigma_max, schedule_k=10, schedule_m=0.4):
    if schedule_type == 'exponential':
        return sigma_min ** (1 - t) * sigma_max ** t
    raise NotImplementedError(schedule_type)


def t_to_sigma(t_tr, t_rot, t_tor, t_sc_tor, args):
    return (t_to_sigma_individual(t_tr, 'exponential', args.tr_sigma_min, args.tr_sigma_max),
            t_to_sigma_individual(t_rot, 'exponential', args.rot_sigma_min, args.rot_sigma_max),
            t_to_sigma_individual(t_tor, 'exponential', args.tor_sigma_min, args.tor_sigma_max),
            t_to_sigma_individual(t_sc_tor, 'exponential', args.sidechain_tor_sigma_min, args.sidechain_tor_sigma_max))


def sinusoidal_embedding(timesteps, dim, scale=1.0, max_positions=10000):
    """Host/torch form (utils/diffusion_utils.py:73-84); the device form is ``ddp_graph_sigma_proj``."""
    half = dim // 2
    emb = math.log(max_positions) / (half - 1)
    emb = torch.exp(torch.arange(half, dtype=torch.float32, device=timesteps.device) * -emb)
    emb = scale * timesteps.float()[:, None] * emb[None, :]
    emb = torch.cat([torch.sin(emb), torch.cos(emb)], dim=1)
    if dim % 2 == 1:
        emb = torch.nn.functional.pad(emb, (0, 1))
    return emb


def get_timestep_embedding(embedding_type, dim, scale=10000):
    if embedding_type != 'sinusoidal':
        raise NotImplementedError('only the sinusoidal time embedding is accelerated')
    return functools.partial(sinusoidal_embedding, dim=dim, scale=scale)


def get_t_schedule(sigma_schedule, inference_steps, inf_sched_alpha=1, inf_sched_beta=1, t_max=1):
    if sigma_schedule == 'expbeta':
        lin_max = beta.cdf(t_max, a=inf_sched_alpha, b=inf_sched_beta)
        c = np.linspace(lin_max, 0, inference_steps + 1)[:-1]
        return beta.ppf(c, a=inf_sched_alpha, b=inf_sched_beta)
    raise Exception()


def set_time(complex_graphs, t, t_tr, t_rot, t_tor, t_sidechain_tor, batchsize, all_atoms, asyncronous_noise_schedule, device,
             include_miscellaneous_atoms=False):
    """utils/diffusion_utils.py:124-165 (node_t / complex_t dictionaries on the batch)."""
    keys = ['ligand', 'receptor'] + (['atom'] if all_atoms else [])
    for k in keys:
        n = complex_graphs[k].num_nodes
        complex_graphs[k].node_t = {'tr': t_tr * torch.ones(n), 'rot': t_rot * torch.ones(n), 'tor': t_tor * torch.ones(n),
                                    'sc_tor': t_sidechain_tor * torch.ones(n)}
    complex_graphs.complex_t = {'tr': t_tr * torch.ones(batchsize), 'rot': t_rot * torch.ones(batchsize),
                                'tor': t_tor * torch.ones(batchsize), 'sc_tor': t_sidechain_tor * torch.ones(batchsize)}
    if asyncronous_noise_schedule:                                   # :158-165
        for k in keys:
            complex_graphs[k].node_t['t'] = t * torch.ones(complex_graphs[k].num_nodes)
        complex_graphs.complex_t['t'] = t * torch.ones(batchsize)


class PoseState:
    """Device-resident pose arrays of a list of samples + the ``ddp_pose_t`` that updates them in place."""

    def __init__(self, data_list, device, lig_pos=None, atom_pos=None, flexible_sidechains=True, no_torsion=False):
        i32 = dict(dtype=torch.int32, device=device)
        n = len(data_list)
        self.n, self.device = n, device
        nl = [g['ligand'].pos.shape[0] for g in data_list]
        cap = int(_lib.lib().ddp_pose_max_ligand_atoms())
        if max(nl) > cap:
            raise ValueError(f'ligand with {max(nl)} atoms: the fused pose update handles at most {cap} atoms per sample')
        na = [g['atom'].pos.shape[0] for g in data_list]
        lo = np.concatenate([[0], np.cumsum(nl)])
        ao = np.concatenate([[0], np.cumsum(na)])
        self.lig_off, self.atom_off = lo, ao
        self.lig_pos = lig_pos if lig_pos is not None else torch.cat([g['ligand'].pos for g in data_list]).float().to(device).contiguous()
        self.atom_pos = atom_pos if atom_pos is not None else torch.cat([g['atom'].pos for g in data_list]).float().to(device).contiguous()
        self.lig_ptr = torch.as_tensor(lo, **i32)
        tb, tptr, masks, mptr = [], [0], [], [0]
        if not no_torsion:
            for s, g in enumerate(data_list):
                em = g['ligand'].edge_mask.bool()
                b = g['ligand', 'ligand'].edge_index.T[em].cpu().numpy() + lo[s]
                mr = g['ligand'].mask_rotate
                mr = mr if isinstance(mr, np.ndarray) else mr[0]
                tb.append(b.reshape(-1, 2))
                tptr.append(tptr[-1] + b.shape[0])
                for k in range(b.shape[0]):
                    masks.append(np.asarray(mr[k], dtype=np.uint8))
                    mptr.append(mptr[-1] + nl[s])
        self.T = tptr[-1]
        self.has_tor = self.T > 0
        if self.has_tor:
            self.tor_ptr = torch.as_tensor(np.asarray(tptr), **i32)
            self.tor_bonds = torch.as_tensor(np.concatenate(tb), **i32).contiguous()
            self.mask_rotate = torch.as_tensor(np.concatenate(masks), dtype=torch.uint8, device=device)
            self.mask_ptr = torch.as_tensor(np.asarray(mptr), **i32)
        sb, sptr, sub, subptr = [], [0], [], [0]
        if flexible_sidechains:
            for s, g in enumerate(data_list):
                if 'flexResidues' not in g or 'edge_idx' not in g['flexResidues']:
                    sptr.append(sptr[-1])
                    continue
                fr = g['flexResidues']
                e = fr.edge_idx.cpu().numpy().reshape(-1, 2) + ao[s]
                sc = fr.subcomponents.cpu().numpy() + ao[s]
                mp = fr.subcomponentsMapping.cpu().numpy().reshape(-1, 2)
                sb.append(e)
                sptr.append(sptr[-1] + e.shape[0])
                for k in range(e.shape[0]):
                    sub.append(sc[mp[k, 0]:mp[k, 1]])
                    subptr.append(subptr[-1] + (mp[k, 1] - mp[k, 0]))
        self.S = sptr[-1]
        self.has_sc = self.S > 0
        if self.has_sc:
            self.sc_ptr = torch.as_tensor(np.asarray(sptr), **i32)
            self.sc_bonds = torch.as_tensor(np.concatenate(sb), **i32).contiguous()
            self.sc_sub_ptr = torch.as_tensor(np.asarray(subptr), **i32)
            self.sc_sub = torch.as_tensor(np.concatenate(sub), **i32)

    def update(self, coef, tr_score, rot_score, tor_score, sc_score, tr_z=None, rot_z=None, tor_z=None, sc_z=None):
        """perturb = a * score + b * z per component, then the fused pose update (one launch).
        ``coef``: 8 host floats, or a device float32[8] tensor (graph-replayable form)."""
        ck = tuple(None if t is None else t.data_ptr() for t in (tr_score, rot_score, tor_score, sc_score, tr_z, rot_z, tor_z, sc_z)) \
            + ((coef.data_ptr(),) if torch.is_tensor(coef) else ())
        cached = getattr(self, '_pose_call', None)
        if cached is not None and cached[0] == ck and torch.is_tensor(coef):     # same buffers as last step: reuse the descriptor
            _lib.check(_lib.lib().ddp_pose_update_dev(cached[1], ptr(coef), _lib.stream_ptr()), 'ddp_pose_update_dev')
            return
        P = _lib.Pose(n_samples=self.n, lig_pos=ptr(self.lig_pos), lig_ptr=ptr(self.lig_ptr),
                      tor_ptr=ptr(self.tor_ptr) if self.has_tor and tor_score is not None else None,
                      tor_bonds=ptr(self.tor_bonds) if self.has_tor else None,
                      mask_rotate=ptr(self.mask_rotate) if self.has_tor else None,
                      mask_ptr=ptr(self.mask_ptr) if self.has_tor else None,
                      atom_pos=ptr(self.atom_pos),
                      sc_ptr=ptr(self.sc_ptr) if self.has_sc and sc_score is not None else None,
                      sc_bonds=ptr(self.sc_bonds) if self.has_sc else None,
                      sc_sub_ptr=ptr(self.sc_sub_ptr) if self.has_sc else None,
                      sc_sub=ptr(self.sc_sub) if self.has_sc else None,
                      tr_score=ptr(tr_score), rot_score=ptr(rot_score),
                      tor_score=ptr(tor_score) if tor_score is not None else None,
                      sc_score=ptr(sc_score) if sc_score is not None else None,
                      tr_z=ptr(tr_z), rot_z=ptr(rot_z), tor_z=ptr(tor_z), sc_z=ptr(sc_z))
        if torch.is_tensor(coef):
            ref = C.byref(P)
            self._pose_call = (ck, ref, P)
            _lib.check(_lib.lib().ddp_pose_update_dev(ref, ptr(coef), _lib.stream_ptr()), 'ddp_pose_update_dev')
            return
        cf = _lib.StepCoef(*[float(v) for v in coef])
        _lib.check(_lib.lib().ddp_pose_update(C.byref(P), C.byref(cf), _lib.stream_ptr()), 'ddp_pose_update')

    def stage_to_host(self):
        """Enqueue asynchronous copies of the poses into pinned host buffers (``write_back`` then reads those; the caller
        synchronises with an event recorded after this call)."""
        self._host = (torch.empty(self.lig_pos.shape, dtype=self.lig_pos.dtype, pin_memory=True),
                      torch.empty(self.atom_pos.shape, dtype=self.atom_pos.dtype, pin_memory=True))
        self._host[0].copy_(self.lig_pos, non_blocking=True)
        self._host[1].copy_(self.atom_pos, non_blocking=True)

    def write_back(self, data_list):
        lp, apos = getattr(self, '_host', None) or (self.lig_pos.cpu(), self.atom_pos.cpu())
        for s, g in enumerate(data_list):
            g['ligand'].pos = lp[self.lig_off[s]:self.lig_off[s + 1]].clone()
            g['atom'].pos = apos[self.atom_off[s]:self.atom_off[s + 1]].clone()


def _cuda_device(data):
    d = data['ligand'].pos.device
    return d if d.type == 'cuda' else torch.device('cuda', torch.cuda.current_device())


def modify_conformer(data, tr_update, rot_update, torsion_updates, pivot=None):
    """utils/diffusion_utils.py:37-60 on one graph (rigid move + torsions + Kabsch re-alignment)."""
    if pivot is not None:
        raise NotImplementedError('pivot alignment is outside the accelerated path')
    dev = _cuda_device(data)
    orig = data['ligand'].pos.device
    st = PoseState([data], dev, flexible_sidechains=False, no_torsion=torsion_updates is None)
    f = lambda v: torch.as_tensor(np.asarray(v, dtype=np.float32) if not torch.is_tensor(v) else v.detach().float().cpu().numpy()).reshape(-1).to(dev).contiguous()
    tor = f(torsion_updates) if torsion_updates is not None and st.has_tor else None
    st.update((1, 0, 1, 0, 1, 0, 0, 0), f(tr_update), f(rot_update), tor, None)
    data['ligand'].pos = st.lig_pos.to(orig)
    return data


def modify_sidechains(data, torsion_updates):
    """utils/diffusion_utils.py:63-70 on one graph (sequential side-chain bond rotations)."""
    dev = _cuda_device(data)
    orig = data['atom'].pos.device
    st = PoseState([data], dev, flexible_sidechains=True, no_torsion=True)
    if not st.has_sc:
        return
    sc = torch.as_tensor(np.asarray(torsion_updates, dtype=np.float32)).reshape(-1).to(dev).contiguous()
    keep = st.lig_pos.clone()
    z3 = torch.zeros(3, device=dev)
    st.update((0, 0, 0, 0, 0, 0, 1, 0), z3, z3, None, sc)
    st.lig_pos.copy_(keep)
    data['atom'].pos = st.atom_pos.to(orig)
