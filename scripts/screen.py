"""Sharded docking of a synthetic complex set (BASELINE.json configs[3] / [4] shape): ligand sizes follow the PDBBind
test-set law (min 7, median 29, mean 35.9, max 147), complexes are split over the ranks like
np.array_split, no communication inside the loop, one all-gather of the best confidences at the end.

  python scripts/screen.py --complexes 16 --samples 10                              # 1 GPU
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 scripts/screen.py --complexes 16
"""
import argparse
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from diffdock_pocket_b200 import inference, inputs, utils  # noqa: E402

p = argparse.ArgumentParser()
p.add_argument('--complexes', type=int, default=16)
p.add_argument('--samples', type=int, default=10)
p.add_argument('--steps', type=int, default=20)
p.add_argument('--mode', default='bf16')
p.add_argument('--batch-complexes', action='store_true', help='pack several complexes into one sampler call')
p.add_argument('--one-pocket', action='store_true', help='configs[4]: one pocket, many ligands')
a = p.parse_args()
world, rank, local = (int(os.environ.get(k, d)) for k, d in (('WORLD_SIZE', 1), ('RANK', 0), ('LOCAL_RANK', 0)))
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
if world > 1:
    dist.init_process_group('nccl', device_id=dev)
model, conf, sa, ca = utils.build_models(dev, seed=0)
model.conv_mode = conf.conv_mode = a.mode
rng = np.random.RandomState(0)
sizes = np.clip(np.round(np.exp(rng.normal(np.log(29.0), 0.55, a.complexes))), 7, 147).astype(int)   # log-normal fit of the size law
rows = []
for i, n_lig in enumerate(sizes):
    g = inputs.synthetic_complex(0 if a.one_pocket else i, n_lig=int(n_lig), n_res=100 if a.one_pocket else int(rng.randint(60, 160)),
                                 flexible_residues=5, name=f'cplx{i}')
    rows.append(dict(complex_name=g.name, complex_graph=g))
args = inference.default_args(samples_per_complex=a.samples, batch_size=20, inference_steps=a.steps)
np.random.seed(1 + rank)
torch.manual_seed(1 + rank)
torch.cuda.synchronize()
t0 = time.time()
res, best, ok = inference.infer_sharded(rows, model, args, sa, dev, filtering_model=conf, filtering_model_args=ca,
                                        batch_complexes=a.batch_complexes)
torch.cuda.synchronize()
dt = time.time() - t0
if world > 1:
    t = torch.tensor([dt], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t.item())
if rank == 0:
    order = torch.argsort(best, descending=True)
    print(f'{a.complexes} complexes x {a.samples} samples x {a.steps} steps on {world} GPU(s): {dt:.2f} s, '
          f'{a.complexes * a.samples / dt:.1f} poses/s; succeeded on rank 0: {ok}; top-3 complexes by confidence: '
          f'{[(int(i), round(float(best[i]), 3)) for i in order[:3]]}')
if world > 1:
    dist.destroy_process_group()
