"""Generate the committed fixtures under tests/golden/ (run HERE, where /root/reference exists).

1. ``3dpf_holo.npz`` / ``3dpf_apo.npz``: the reference's only real input fixture
   (example_data/3dpf_*), converted to the HeteroData field layout by diffdock_pocket_b200.inputs
   (ESM features are NOT stored: they are regenerated from a seed at load time).
Reference OUTPUT fixtures (``ref_*.npz``) come from ``scripts/make_ref_fixtures.py``, which executes the reference's own modules.
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



if __name__ == '__main__':
    main()
