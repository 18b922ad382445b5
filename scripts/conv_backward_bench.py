"""Forward + backward of one TensorProductConvLayer (training path): this repo's autograd node (fused forward kernels,
ddp_tp_backward + library GEMMs) against PyTorch autograd through the oracle layer on the box's host cores -- the oracle's
FasterTensorProduct / conv layer are the reference's torch ops line by line (models/layers.py:8-85,
models/score_model.py:84-125; the oracle is a CPU restatement and not device-agnostic), i.e. what `loss.backward()`
does for one conv in the reference's CPU environment.
    python scripts/conv_backward_bench.py [n_edges] [case]        case: l3 (W = 10000, default) | l1 | lmax2
Test infrastructure: imports oracle/."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from diffdock_pocket_b200.score_model import TensorProductConvLayer  # noqa: E402
from oracle import e3nn_mini as E  # noqa: E402
from oracle.score_model_ref import TensorProductConvLayer as RefConv  # noqa: E402

n_edges = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
case = sys.argv[2] if len(sys.argv) > 2 else 'l3'
SEQ = ['60x0e', '60x0e + 10x1o', '60x0e + 10x1o + 10x1e', '60x0e + 10x1o + 10x1e + 60x0o']
in_ir, out_ir, faster = {'l1': (SEQ[1], SEQ[2], True), 'l3': (SEQ[3], SEQ[3], True), 'lmax2': (SEQ[3], SEQ[3], False)}[case]
sh_ir = '1x0e+1x1o' if faster else '1x0e+1x1o+1x2e'
dev = torch.device('cuda:0')
torch.manual_seed(0)
prod = TensorProductConvLayer(in_ir, sh_ir, out_ir, 180, residual=False, batch_norm=False, faster=faster).to(dev)
ref = RefConv(in_ir, sh_ir, out_ir, 180, residual=False, batch_norm=False, faster=faster)
ref.load_state_dict(prod.state_dict())
n = max(64, n_edges // 16)
x = torch.randn(n, E.Irreps(in_ir).dim, device=dev)
ei = torch.randint(0, n, (2, n_edges), device=dev)
ea = torch.randn(n_edges, 180, device=dev)
sh = E.spherical_harmonics(sh_ir, torch.randn(n_edges, 3)).to(dev)
probe = torch.randn(n, E.Irreps(out_ir).dim, device=dev)


def step(layer, d):
    xs, eas = x.to(d).clone().requires_grad_(True), ea.to(d).clone().requires_grad_(True)
    out = layer(xs, ei.to(d), eas, sh.to(d), out_nodes=n)
    params = [layer.fc[0].weight, layer.fc[0].bias, layer.fc[3].weight, layer.fc[3].bias]
    return torch.autograd.grad((out * probe.to(d)).sum(), [xs, eas] + params)


def timed(layer, reps):
    for _ in range(2):
        step(layer, dev)
    torch.cuda.synchronize()
    torch.cuda.reset_peak_memory_stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g = step(layer, dev)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, torch.cuda.max_memory_allocated() / 2 ** 30, g


for mode in ('fp32', 'bf16x3'):
    prod.conv_mode = mode
    ms, gib, gp = timed(prod, 5)
    print(f'this repo, forward {mode:6s} + backward: {ms:8.2f} ms per step   peak memory {gib:5.2f} GiB   ({n_edges} edges, case {case}, W = {prod.tp.weight_numel})')
import time  # noqa: E402
torch.set_num_threads(os.cpu_count())
step(ref, 'cpu')
t0 = time.time()
gr = step(ref, 'cpu')
print(f'torch autograd through the oracle layer, {os.cpu_count()} host cores: {(time.time() - t0) * 1000:8.1f} ms per step')
err = max(float((a.cpu() - b).abs().max() / b.abs().max().clamp(min=1e-30)) for a, b in zip(gp, gr))
print(f'max relative gradient difference {err:.2e}')
