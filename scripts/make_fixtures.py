"""Generate the committed fixtures under tests/golden/ (run HERE, where /root/reference exists).

1. ``3dpf_holo.npz`` / ``3dpf_apo.npz``: the reference's only real input fixture
   (example_data/3dpf_*), converted to the HeteroData field layout by diffdock_pocket_b200.inputs
   (ESM features are NOT stored: they are regenerated from a seed at load time).
2. ``golden_forward.npz``: outputs of the CPU oracle (oracle/score_model_ref.py) on a seeded batch of
   that graph with seeded random weights -- the reference itself cannot be imported in this container
   (SURVEY.md F4), so these vectors pin the oracle against regressions and give the GPU tests a
   machine-independent target.
"""
import copy
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from diffdock_pocket_b200 import inputs, so3, torus, utils  # noqa: E402
from diffdock_pocket_b200.hetero import Batch  # noqa: E402
from oracle import diffusion_ref as D, factory, sampling_ref as S  # noqa: E402

GOLD = os.path.join(ROOT, 'tests', 'golden')
EX = '/root/reference/example_data'


def main():
    os.makedirs(GOLD, exist_ok=True)
    holo = inputs.load_example_3dpf(EX, apo=False, flexible='auto')
    inputs.save_graph_npz(holo, os.path.join(GOLD, '3dpf_holo.npz'))
    # README.md:47 flexible list is in holo numbering (80..242); the ESMFold file is numbered from 1
    flex = [f'A:{int(r) - 79}' for r in (160, 193, 197, 198, 222, 224, 227)]
    apo = inputs.load_example_3dpf(EX, apo=True, flexible=flex, pocket_center=[9.7742, 27.2863, 14.6573])
    inputs.save_graph_npz(apo, os.path.join(GOLD, '3dpf_apo.npz'))
    print('holo', holo['ligand'].pos.shape, holo['receptor'].pos.shape, holo['atom'].pos.shape, holo['flexResidues'].edge_idx.shape)
    print('apo ', apo['ligand'].pos.shape, apo['receptor'].pos.shape, apo['atom'].pos.shape, apo['flexResidues'].edge_idx.shape)

    torch.set_num_threads(os.cpu_count())
    g = inputs.load_graph_npz(os.path.join(GOLD, '3dpf_holo.npz'))
    model, conf, sa, ca = utils.build_models(torch.device('cpu'), seed=0)
    om = factory.oracle_model(sa, model.state_dict(), so3.score_norm_np, torus.score_norm)
    oc = factory.oracle_model(ca, conf.state_dict(), so3.score_norm_np, torus.score_norm, confidence_mode=True)
    np.random.seed(0)
    torch.manual_seed(0)
    dl = [copy.deepcopy(g) for _ in range(3)]
    S.randomize_position(dl, False, False, sa.tr_sigma_max, flexible_sidechains=True)
    out = {'lig_pos': torch.stack([d['ligand'].pos for d in dl]).numpy(), 'atom_pos': torch.stack([d['atom'].pos for d in dl]).numpy()}
    for tag, t in (('t70', 0.7), ('t05', 0.05)):
        b = Batch.from_data_list(copy.deepcopy(dl))
        D.set_time(b, t, t, t, t, 3)
        with torch.no_grad():
            tr, rot, tor, sc = om(b)
        dbg = om._debug
        out.update({f'{tag}_tr': tr.numpy(), f'{tag}_rot': rot.numpy(), f'{tag}_tor': tor.numpy(), f'{tag}_sc': sc.numpy(),
                    f'{tag}_n_ll': dbg['ll'].shape[1], f'{tag}_n_aa': dbg['aa'].shape[1], f'{tag}_n_lr': dbg['lr'].shape[1],
                    f'{tag}_n_la': dbg['la'].shape[1],
                    f'{tag}_lig_layers': torch.stack([torch.nn.functional.pad(l[0], (0, 180 - l[0].shape[1])) for l in dbg['layers']]).numpy()})
    b = Batch.from_data_list(copy.deepcopy(dl))
    D.set_time(b, 0, 0, 0, 0, 3)
    with torch.no_grad():
        out['confidence'] = oc(b).numpy()
    np.savez_compressed(os.path.join(GOLD, 'golden_forward.npz'), **out)
    print({k: (v.shape if hasattr(v, 'shape') else v) for k, v in out.items()})


if __name__ == '__main__':
    main()
