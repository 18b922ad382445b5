"""Per-conv timing of one score-model forward (batch of n x 3dpf apo): CUDA events around every fused TP-conv launch.
  python scripts/conv_bench.py [mode] [n] [reps]
Prints achieved algorithmic TFLOP/s per edge set / layer and the total, plus the forward wall time."""
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
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
dev = torch.device('cuda:0')
model, _, sa, _ = utils.build_models(dev, with_confidence=False)
model.conv_mode = mode
model.group_convs = os.environ.get('DDP_NO_GROUP', '0') != '1'
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
    agg = {}
    for _ in range(reps):
        model.profile = []
        model.run_plan(pl, ct)
        torch.cuda.synchronize()
        for i, (e0, e1, convs) in enumerate(model.profile):
            ne = sum(int(es.n_dev.item()) for (_, es, _) in convs)
            nt = sum((int(es.n_dev.item()) + 127) // 128 for (_, es, _) in convs)
            fl = sum(2.0 * (9 * ns * ns + 3 * ns * w + w) * int(es.n_dev.item()) for (w, es, ns) in convs)
            a = agg.setdefault(i, [0.0, fl, ne, convs[0][0], nt, len(convs)])
            a[0] += e0.elapsed_time(e1) / reps
    model.profile = None
    tot_ms = sum(a[0] for a in agg.values())
    tot_fl = sum(a[1] for a in agg.values())
    for i, (ms, fl, ne, w, nt, nc) in agg.items():
        print(f'launch {i:2d}  convs={nc} W={w:5d} E={ne:7d} tiles={nt:5d}  {ms * 1e3:8.1f} us  {fl / ms / 1e9:7.1f} TFLOP/s')
    print(f'TOTAL conv {tot_ms:.3f} ms  {tot_fl / tot_ms / 1e9:.1f} TFLOP/s')
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(reps):
        model.run_plan(pl, ct)
    ev1.record()
    torch.cuda.synchronize()
    print(f'forward {ev0.elapsed_time(ev1) / reps:.3f} ms (eager launches)')
