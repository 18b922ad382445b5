"""Kernel-side time of ddp_knn_graph on the atom graph of a resident mini-batch (20 x 3dpf apo pocket, k = 8):
   python scripts/dbg/knn_time.py ; DDP_KNN_GRID=0 python scripts/dbg/knn_time.py   (plain filtered scan)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from diffdock_pocket_b200 import _lib, inputs  # noqa: E402
from diffdock_pocket_b200._lib import ptr  # noqa: E402

g = inputs.load_graph_npz('tests/golden/3dpf_apo.npz', name='x')
na, B, k = g['atom'].pos.shape[0], 20, 8
pos = torch.cat([g['atom'].pos for _ in range(B)]).float().cuda().contiguous()
p = (torch.arange(B + 1, dtype=torch.int32) * na).cuda()
n = pos.shape[0]
slab = torch.empty(n * (k + 1), dtype=torch.int32, device='cuda')
counts = torch.zeros(n + 1, dtype=torch.int32, device='cuda')
cap = n * (k + 1)
edge = torch.empty(2 * cap, dtype=torch.int32, device='cuda')
n_dev = torch.zeros(1, dtype=torch.int32, device='cuda')
L = _lib.lib()
call = lambda: _lib.check(L.ddp_knn_graph(ptr(pos), ptr(p), B, n, k, ptr(slab), k + 1, ptr(counts), ptr(edge), cap, ptr(n_dev), _lib.stream_ptr()), 'knn')
for _ in range(5):
    call()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50):
    call()
e1.record()
torch.cuda.synchronize()
print(f'DDP_KNN_GRID={os.environ.get("DDP_KNN_GRID", "1")}: {e0.elapsed_time(e1) * 1000 / 50:.1f} us per ddp_knn_graph call (search + scan + compaction), '
      f'{n} atoms in {B} pockets, {int(n_dev.item())} edges')
