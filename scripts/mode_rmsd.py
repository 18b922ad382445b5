"""Final-pose agreement between conv modes on the FULL model (README big score model, random init): one seeded
20-step sampling run of n samples of 3dpf apo per mode, RMSD of ligand / flexible side-chain atoms vs the fp32 mode."""
import copy
import os
import sys
from functools import partial

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from diffdock_pocket_b200 import diffusion_utils as du, inputs, sampling as ps, utils  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
dev = torch.device('cuda:0')
model, conf, sa, ca = utils.build_models(dev, seed=0, with_confidence=False)
g = inputs.load_graph_npz(os.path.join(ROOT, 'tests', 'golden', '3dpf_apo.npz'))
np.random.seed(0)
torch.manual_seed(0)
dl = [copy.deepcopy(g) for _ in range(n)]
ps.randomize_position(dl, False, False, sa.tr_sigma_max, flexible_sidechains=True)
sch = du.get_t_schedule('expbeta', steps)
t2s = partial(du.t_to_sigma, args=sa)
out = {}
for mode in ('fp32', 'bf16x3', 'bf16'):
    model.conv_mode = mode
    torch.manual_seed(5)
    res, _ = ps.sampling(copy.deepcopy(dl), model, steps, sch, sch, sch, sch, dev, t2s, sa, batch_size=n, **bench.TEMP)
    out[mode] = (torch.stack([r['ligand'].pos for r in res]), torch.stack([r['atom'].pos for r in res]))
flex = g['flexResidues'].subcomponents.unique()
for mode in ('bf16x3', 'bf16'):
    dl_ = ((out[mode][0] - out['fp32'][0]) ** 2).sum(-1).mean(-1).sqrt()
    da_ = ((out[mode][1][:, flex] - out['fp32'][1][:, flex]) ** 2).sum(-1).mean(-1).sqrt()
    print(f'{mode}: ligand RMSD vs fp32 per sample {[round(float(v), 4) for v in dl_]}  side-chain RMSD {[round(float(v), 4) for v in da_]}')
