"""Host-side torsion helpers used by the initial-state generator (``randomize_position``).

Mirrors ``modify_conformer_torsion_angles`` (utils/torsion.py:68-94) and the per-bond side-chain
rotation (utils/torsion.py:251-278) for ONE graph on the CPU: this runs once per complex before
the diffusion loop; inside the loop the same rotations are done by the ``ddp_pose_update`` kernel.
"""
import numpy as np
import torch


def _rotvec_matrix(rot_vec):
    """float64 rotation matrix of a rotation vector (what ``scipy Rotation.from_rotvec(v).as_matrix()`` returns, to
    rounding): Rodrigues' formula, R = I + sin(t) K + (1 - cos(t)) K^2 with K the cross-product matrix of the unit axis."""
    t = float(np.sqrt(rot_vec[0] * rot_vec[0] + rot_vec[1] * rot_vec[1] + rot_vec[2] * rot_vec[2]))
    if t < 1e-12:
        return np.eye(3)
    x, y, z = rot_vec[0] / t, rot_vec[1] / t, rot_vec[2] / t
    c, s_ = np.cos(t), np.sin(t)
    C = 1.0 - c
    return np.array([[c + x * x * C, x * y * C - z * s_, x * z * C + y * s_],
                     [y * x * C + z * s_, c + y * y * C, y * z * C - x * s_],
                     [z * x * C - y * s_, z * y * C + x * s_, c + z * z * C]])


def _rotate(pos, u, v, idx, angle):
    axis = (pos[u] - pos[v]).astype(np.float64)
    rot_vec = axis * angle / np.linalg.norm(axis)
    pos[idx] = (pos[idx] - pos[v]) @ _rotvec_matrix(rot_vec).T + pos[v]


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
    p = data['atom'].pos.detach().cpu().numpy().astype(np.float32, copy=True)
    sub = fr.subcomponents.cpu().numpy()
    bonds = fr.edge_idx.cpu().numpy().reshape(-1, 2)
    mapping = fr.subcomponentsMapping.cpu().numpy().reshape(-1, 2)
    for k, upd in enumerate(torsion_updates):
        if upd == 0:
            continue
        _rotate(p, int(bonds[k, 0]), int(bonds[k, 1]), sub[int(mapping[k, 0]):int(mapping[k, 1])], upd)     # fp64 rotation, fp32 positions
    data['atom'].pos = torch.from_numpy(p).to(data['atom'].pos.device)
