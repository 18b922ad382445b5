"""Host-side torsion helpers used by the initial-state generator (``randomize_position``).

Mirrors ``modify_conformer_torsion_angles`` (utils/torsion.py:68-94) and the per-bond side-chain
rotation (utils/torsion.py:251-278) for ONE graph on the CPU: this runs once per complex before
the diffusion loop; inside the loop the same rotations are done by the ``ddp_pose_update`` kernel.
"""
import numpy as np
import torch
from scipy.spatial.transform import Rotation as R


def _rotate(pos, u, v, idx, angle):
    axis = pos[u] - pos[v]
    rot_vec = axis * angle / np.linalg.norm(axis)
    pos[idx] = (pos[idx] - pos[v]) @ R.from_rotvec(rot_vec).as_matrix().T + pos[v]


def modify_conformer_torsion_angles(pos, edge_index, mask_rotate, torsion_updates, as_numpy=False):
    dev = pos.device if torch.is_tensor(pos) else None
    p = pos.detach().cpu().numpy().copy() if torch.is_tensor(pos) else np.array(pos, copy=True)
    edges = edge_index.cpu().numpy() if torch.is_tensor(edge_index) else np.asarray(edge_index)
    for k, (u, v) in enumerate(edges):
        if torsion_updates[k] == 0:
            continue
        assert not mask_rotate[k, u] and mask_rotate[k, v]
        _rotate(p, u, v, mask_rotate[k], torsion_updates[k])
    if as_numpy:
        return p
    return torch.from_numpy(p.astype(np.float32)).to(dev)


def modify_sidechains_host(data, torsion_updates):
    fr = data['flexResidues']
    p = data['atom'].pos.detach().cpu().numpy().copy()
    sub = fr.subcomponents.cpu().numpy()
    for k, upd in enumerate(torsion_updates):
        if upd == 0:
            continue
        u, v = int(fr.edge_idx[k][0]), int(fr.edge_idx[k][1])
        m0, m1 = int(fr.subcomponentsMapping[k][0]), int(fr.subcomponentsMapping[k][1])
        _rotate(p, u, v, sub[m0:m1], upd)
        p = p.astype(np.float32)
    data['atom'].pos = torch.from_numpy(p.astype(np.float32)).to(data['atom'].pos.device)
