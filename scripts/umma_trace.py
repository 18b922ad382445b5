"""Per-tile timeline of CTA 0 of the grouped layer-3 conv launch (MMA-issue warp vs epilogue warp 0), from the
clock64 trace hooks of the UMMA kernel.  The hooks are compiled out of the product library; build a traced copy first
(here, no GPU needed) and point DDP_LIB at it on the GPU box:
  DDP_NVCC_FLAGS=-DDDP_UMMA_TRACE DDP_LIB=$PWD/scripts/micro/libddp_trace.so python -c "from diffdock_pocket_b200 import _lib; _lib.build()"
  DDP_LIB=$PWD/scripts/micro/libddp_trace.so python scripts/umma_trace.py > gpurun_out/trace.txt"""
import copy
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from diffdock_pocket_b200 import _lib, inputs, sampling as S, utils  # noqa: E402
from diffdock_pocket_b200.hetero import Batch  # noqa: E402

n = 20
dev = torch.device('cuda:0')
model, _, sa, _ = utils.build_models(dev, with_confidence=False)
model.conv_mode = 'bf16'
g = inputs.load_graph_npz(os.path.join(ROOT, 'tests', 'golden', '3dpf_apo.npz'))
np.random.seed(0)
torch.manual_seed(0)
dl = [copy.deepcopy(g) for _ in range(n)]
S.randomize_position(dl, False, False, sa.tr_sigma_max, flexible_sidechains=True)
L = _lib.lib()
model.replay_launches = False          # the hook below needs every launch to go through the library wrapper
with torch.no_grad():
    pl = model.make_plan(Batch.from_data_list(dl))
    ct = {k: torch.full((n,), 0.5) for k in ('tr', 'rot', 'tor', 'sc_tor')}
    model.run_plan(pl, ct)
    torch.cuda.synchronize()
    slots = L.ddp_tpconv_umma_set_trace(None)
    assert slots > 0, 'library built without -DDDP_UMMA_TRACE (see the module docstring)'
    buf = torch.zeros(slots, dtype=torch.int64, device=dev)
    # trace only the 4th grouped launch (layer 3): enable, run forward with a hook counting launches
    orig = L.ddp_tpconv_umma_group
    cnt = [0]

    def hooked(*a):
        cnt[0] += 1
        L.ddp_tpconv_umma_set_trace(buf.data_ptr() if cnt[0] == 4 else None)
        return orig(*a)
    L.ddp_tpconv_umma_group = hooked
    model.run_plan(pl, ct)
    torch.cuda.synchronize()
    L.ddp_tpconv_umma_set_trace(None)
t = buf.cpu().numpy().reshape(3, -1, 8)
t0 = t[0, 0, 0]
print('iter | MMA: wait_empty_start empty_ok first_full issued | EPI warpgroup 0: ready_to_wait full_ok released red_done | EPI warpgroup 1: same   (cycles rel. to start)')
n_it = int((t[0, :, 0] != 0).sum())
print(f'{n_it} tile iterations traced (two MMA issuers: even iterations warp 5, odd iterations warp 7)')
for i in range(int(os.environ.get('TRACE_ROWS', '160'))):
    m, e, e2 = t[0, i], t[1, i], t[2, i]
    if m[0] == 0:
        break
    print(f'{i:4d} | ' + ' '.join(f'{int(v - t0):8d}' for v in m[:4]) + ' | ' + ' '.join(f'{int(v - t0):8d}' for v in e[:4])
          + ' | ' + ' '.join(f'{int(v - t0):8d}' for v in e2[:4]))
