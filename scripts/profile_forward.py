"""One score-model forward (batch of 20 x 3dpf apo) between cudaProfilerStart/Stop, for ncu:
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file launches.csv python scripts/profile_forward.py
  ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:tpconv_umma -s 30 -c 1 -o prof python scripts/profile_forward.py
"""
import copy
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from diffdock_pocket_b200 import inputs, sampling as S, utils  # noqa: E402
from diffdock_pocket_b200.hetero import Batch  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else 'bf16'
n = int(sys.argv[2]) if len(sys.argv) > 2 else 20
dev = torch.device('cuda:0')
model, _, sa, _ = utils.build_models(dev, with_confidence=False)
model.conv_mode = mode
g = inputs.load_graph_npz(os.path.join(ROOT, 'tests', 'golden', '3dpf_apo.npz'))
np.random.seed(0)
torch.manual_seed(0)
dl = [copy.deepcopy(g) for _ in range(n)]
S.randomize_position(dl, False, False, sa.tr_sigma_max, flexible_sidechains=True)
with torch.no_grad():
    pl = model.make_plan(Batch.from_data_list(dl))
    ct = {k: torch.full((n,), 0.5) for k in ('tr', 'rot', 'tor', 'sc_tor')}
    for _ in range(2):
        model.run_plan(pl, ct)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    model.run_plan(pl, ct)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print('edges', {k: int(v.n_dev.item()) for k, v in pl.es.items()})
