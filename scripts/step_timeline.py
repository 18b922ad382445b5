"""Kernel timeline (CUPTI via torch.profiler) of ONE replayed step graph of a 20-sample mini-batch: start / duration of
every kernel and the idle gaps between them.  python scripts/step_timeline.py > gpurun_out/step_timeline.txt"""
import argparse
import os
import sys
from functools import partial

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from diffdock_pocket_b200 import diffusion_utils as du, sampling as ps, utils  # noqa: E402

args = argparse.Namespace(samples=20, batch_size=20, inference_steps=20, workload='3dpf_apo')
dev = torch.device('cuda:0')
model, conf, sa, ca = utils.build_models(dev, seed=0, with_confidence=False)
model.conv_mode = 'bf16'
g, dl0 = bench.workload(args, 0)
sch = du.get_t_schedule('expbeta', 20)
t2s = partial(du.t_to_sigma, args=sa)
with torch.no_grad():
    r = ps.StepRunner(model, dl0, True, False, use_graph=True)
    coefs = [ps.step_coefficients(i, 20, (sch,) * 4, t2s, sa, False, bench.TEMP['temp_sampling'], bench.TEMP['temp_psi'],
                                  bench.TEMP['temp_sigma_data'], True) for i in range(20)]
    z = torch.randn(r.n_extra - 8)
    for i in range(4):
        r.step(*coefs[i], z)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        r.step(*coefs[5], z)
        torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
t0 = ev[0].time_range.start
prev_end = t0
busy = 0.0
rows = []
for e in ev:
    s, d = e.time_range.start - t0, e.time_range.end - e.time_range.start
    gap = e.time_range.start - prev_end
    rows.append((s, d, gap, e.name[:70]))
    busy += d
    prev_end = max(prev_end, e.time_range.end)
total = prev_end - t0
print(f'{len(ev)} device activities, span {total:.1f} us, busy {busy:.1f} us, idle {total - busy:.1f} us')
for s, d, gap, n in rows:
    print(f'{s:9.1f} {d:8.1f} gap {gap:6.1f}  {n}')
