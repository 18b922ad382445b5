"""CPU restatement of the reference sampler loop (TEST INFRASTRUCTURE).

* ``randomize_position`` <- utils/sampling.py:16-60 (no pocket_knowledge branch: inference.py:140
  calls it with the pocket centre already at the origin)
* ``sampling``           <- utils/sampling.py:70-286 (SVGD branch omitted: svgd_weight is 0 on the
  inference path and the reference raises NotImplementedError for it with flexible side chains)

Random draws use the same generators in the same order as the reference (numpy global RNG for the
initial torsions, scipy ``Rotation.random``, and on the torch CPU default generator per step: the
``_base_seed`` of the step's DataLoader iterator (utils/sampling.py:100,112), then ``torch.normal`` for
tr_z, rot_z, tor_z, sidechain_tor_z) so a seeded run consumes identical streams.
Graphs are collated by ``oracle.pyg_mini`` (not by the product's ``hetero`` module).
"""
import copy

import numpy as np
import torch
from scipy.spatial.transform import Rotation as R

from .pyg_mini import DataLoader
from .diffusion_ref import (modify_conformer, modify_conformer_torsion_angles, modify_sidechains, set_time)


def randomize_position(data_list, no_torsion, no_random, tr_sigma_max, flexible_sidechains=False):
    if not no_torsion:
        for g in data_list:
            upd = np.random.uniform(low=-np.pi, high=np.pi, size=int(g['ligand'].edge_mask.sum()))
            mr = g['ligand'].mask_rotate
            mr = mr if isinstance(mr, np.ndarray) else mr[0]
            g['ligand'].pos = modify_conformer_torsion_angles(
                g['ligand'].pos, g['ligand', 'ligand'].edge_index.T[g['ligand'].edge_mask], mr, upd)
    if flexible_sidechains:
        for g in data_list:
            upd = np.random.uniform(low=-np.pi, high=np.pi, size=len(g['flexResidues'].edge_idx))
            modify_sidechains(g, upd)
    for g in data_list:
        center = torch.mean(g['ligand'].pos, dim=0, keepdim=True)
        rot = torch.from_numpy(R.random().as_matrix()).float()
        g['ligand'].pos = (g['ligand'].pos - center) @ rot.T
        if not no_random:
            g['ligand'].pos += torch.normal(mean=0, std=tr_sigma_max, size=(1, 3))


def sampling(data_list, model, inference_steps, tr_schedule, rot_schedule, tor_schedule, sidechain_tor_schedule,
             t_to_sigma, model_args, no_random=False, ode=False, confidence_model=None, batch_size=32,
             no_final_step_noise=False, temp_sampling=1.0, temp_psi=0.0, temp_sigma_data=0.5,
             flexible_sidechains=None, max_steps=None, trace=None, asyncronous_noise_schedule=False, t_schedule=None):
    """Returns (data_list, confidence).  ``max_steps`` truncates the loop (bounded CPU-baseline
    samples); ``trace`` (a list) receives the per-step scores for parity tests.  Under the asynchronous noise schedule
    ``set_time`` also receives ``t_schedule[t_idx]`` (utils/sampling.py:116-117; the confidence pass: 0, :273-278)."""
    flexible_sidechains = model_args.flexible_sidechains if flexible_sidechains is None else flexible_sidechains
    N = len(data_list)
    ma = model_args
    for t_idx in range(inference_steps if max_steps is None else min(max_steps, inference_steps)):
        t_tr, t_rot, t_tor, t_sc = tr_schedule[t_idx], rot_schedule[t_idx], tor_schedule[t_idx], sidechain_tor_schedule[t_idx]
        last = t_idx == inference_steps - 1
        dt_tr = t_tr - tr_schedule[t_idx + 1] if not last else t_tr
        dt_rot = t_rot - rot_schedule[t_idx + 1] if not last else t_rot
        dt_tor = t_tor - tor_schedule[t_idx + 1] if not last else t_tor
        dt_sc = t_sc - sidechain_tor_schedule[t_idx + 1] if not last else t_sc
        tr_sigma, rot_sigma, tor_sigma, sc_sigma = t_to_sigma(t_tr, t_rot, t_tor, t_sc)
        scores = [[], [], [], []]
        for batch in DataLoader(data_list, batch_size=batch_size):
            set_time(batch, t_tr, t_rot, t_tor, t_sc, batch.num_graphs,
                     t=(t_schedule[t_idx] if t_schedule is not None else None) if asyncronous_noise_schedule else None)
            with torch.no_grad():
                out = model(batch)
            for lst, o in zip(scores, out):
                lst.append(o)
        tr_score, rot_score, tor_score, sc_score = [torch.cat(s, 0) for s in scores]
        if trace is not None:
            trace.append((tr_score.clone(), rot_score.clone(), tor_score.clone(), sc_score.clone()))
        tr_g = tr_sigma * torch.sqrt(torch.tensor(2 * np.log(ma.tr_sigma_max / ma.tr_sigma_min)))
        rot_g = 2 * rot_sigma * torch.sqrt(torch.tensor(np.log(ma.rot_sigma_max / ma.rot_sigma_min)))
        zero_noise = no_random or (no_final_step_noise and last)
        draw = lambda shape: torch.zeros(shape) if zero_noise else torch.normal(mean=0, std=1, size=shape)
        if ode:
            tr_perturb = 0.5 * tr_g ** 2 * dt_tr * tr_score
            rot_perturb = 0.5 * rot_score * dt_rot * rot_g ** 2
        else:
            tr_z = draw((N, 3))
            tr_perturb = tr_g ** 2 * dt_tr * tr_score + tr_g * np.sqrt(dt_tr) * tr_z
            rot_z = draw((N, 3))
            rot_perturb = rot_score * dt_rot * rot_g ** 2 + rot_g * np.sqrt(dt_rot) * rot_z
        tor_perturb = None
        if not ma.no_torsion:
            tor_g = tor_sigma * torch.sqrt(torch.tensor(2 * np.log(ma.tor_sigma_max / ma.tor_sigma_min)))
            if ode:
                tor_perturb = (0.5 * tor_g ** 2 * dt_tor * tor_score).numpy()
            else:
                tor_z = draw(tuple(tor_score.shape))
                tor_perturb = (tor_g ** 2 * dt_tor * tor_score + tor_g * np.sqrt(dt_tor) * tor_z).numpy()
            tpm = tor_perturb.shape[0] // N
        sc_perturb = None
        if flexible_sidechains:
            sc_g = sc_sigma * torch.sqrt(torch.tensor(2 * np.log(ma.sidechain_tor_sigma_max / ma.sidechain_tor_sigma_min)))
            if ode:
                sc_perturb = (0.5 * sc_g ** 2 * dt_sc * sc_score).numpy()
            else:
                sc_z = draw(tuple(sc_score.shape))
                sc_perturb = (sc_g ** 2 * dt_sc * sc_score + sc_g * np.sqrt(dt_sc) * sc_z).numpy()
            spm = sc_perturb.shape[0] // N
        ts = list(temp_sampling) if hasattr(temp_sampling, '__iter__') else [temp_sampling] * 4
        tp = list(temp_psi) if hasattr(temp_psi, '__iter__') else [temp_psi] * 4

        def sigma_data(smax, smin):
            return np.exp(temp_sigma_data * np.log(smax) + (1 - temp_sigma_data) * np.log(smin))
        if ts[0] != 1.0:                                              # utils/sampling.py:177-180
            sd = sigma_data(ma.tr_sigma_max, ma.tr_sigma_min)
            lam = (sd + tr_sigma) / (sd + tr_sigma / ts[0])
            tr_perturb = tr_g ** 2 * dt_tr * (lam + ts[0] * tp[0] / 2) * tr_score + tr_g * np.sqrt(dt_tr * (1 + tp[0])) * tr_z
        if ts[1] != 1.0:
            sd = sigma_data(ma.rot_sigma_max, ma.rot_sigma_min)
            lam = (sd + rot_sigma) / (sd + rot_sigma / ts[1])
            rot_perturb = rot_g ** 2 * dt_rot * (lam + ts[1] * tp[1] / 2) * rot_score + rot_g * np.sqrt(dt_rot * (1 + tp[1])) * rot_z
        if ts[2] != 1.0 and not ma.no_torsion:
            sd = sigma_data(ma.tor_sigma_max, ma.tor_sigma_min)
            lam = (sd + tor_sigma) / (sd + tor_sigma / ts[2])
            tor_perturb = (tor_g ** 2 * dt_tor * (lam + ts[2] * tp[2] / 2) * tor_score + tor_g * np.sqrt(dt_tor * (1 + tp[2])) * tor_z).numpy()
        if flexible_sidechains and ts[3] != 1.0:
            sd = sigma_data(ma.sidechain_tor_sigma_max, ma.sidechain_tor_sigma_min)
            lam = (sd + sc_sigma) / (sd + sc_sigma / ts[3])
            sc_perturb = (sc_g ** 2 * dt_sc * (lam + ts[3] * tp[3] / 2) * sc_score + sc_g * np.sqrt(dt_sc * (1 + tp[3])) * sc_z).numpy()

        if flexible_sidechains:                                        # utils/sampling.py:245-247
            for i, g in enumerate(data_list):
                modify_sidechains(g, sc_perturb[i * spm:(i + 1) * spm])
        data_list = [modify_conformer(g, tr_perturb[i:i + 1].float(), rot_perturb[i:i + 1].squeeze(0).float(),
                                      tor_perturb[i * tpm:(i + 1) * tpm] if not ma.no_torsion else None)
                     for i, g in enumerate(data_list)]
    confidence = None
    if confidence_model is not None:                                   # utils/sampling.py:263-281
        conf = []
        with torch.no_grad():
            loader = DataLoader(data_list, batch_size=batch_size)
            iter(DataLoader(None, batch_size=batch_size))             # filtering_loader (:266): its iterator is created even without filtering data
            for batch in loader:
                set_time(batch, 0, 0, 0, 0, N, t=0 if asyncronous_noise_schedule else None)
                conf.append(confidence_model(batch))
        confidence = torch.cat(conf, 0)
    return data_list, confidence
