"""Headline benchmark: docked poses/sec for 20-step reverse diffusion (BASELINE.json `metric`).

One bench "step" = one complete pass of the hot path over one batch of synthetic input: `samples` poses of
one complex through `inference_steps` reverse-diffusion steps (score-model forward + fused pose update per
mini-batch per step) plus the confidence pass.  Workload at N=1 is BASELINE.json configs[1] (3dpf ESMFold apo
pocket, explicit centre, 7 flexible residues, 40 samples, batch 20) built from tests/golden/3dpf_apo.npz with
seeded synthetic ESM features and random-init weights of the README big score model (no checkpoint offline).

  python bench.py --gpus N --steps K --warmup W            # this implementation (one rank per GPU under torchrun)
  python bench.py --impl reference ...                      # CPU arm: the oracle port on the host cores

`value`  : poses/s with the batch already resident in HBM (plans built before the timed region).
`e2e`    : poses/s through the public `sampling()` call with HOST graphs: collate + H2D upload of the batch,
           the whole loop, D2H of final poses and confidences inside the timed region.
`roofline`: the fused TP-conv kernel (tcgen05), achieved algorithmic TFLOP/s measured with CUDA events around
           its launches in one instrumented forward, against the measured bf16 peak (MEASURED_PEAKS.json).
Multi-GPU: complexes are sharded (each rank docks its own copy of the workload: weak scaling); the only
collective is one NCCL all-gather of final poses + confidences for ranking, after the loop.
"""
import argparse
import copy
import json
import os
import subprocess
import sys
import threading
import time
from functools import partial

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, 'tests', 'golden')
TEMP = dict(temp_sampling=[0.9766, 6.0774, 6.7616, 1.4488], temp_psi=[1.5103, 0.8141, 0.7662, 1.3396], temp_sigma_data=0.48884)


def parse():
    p = argparse.ArgumentParser()
    p.add_argument('--gpus', type=int, default=1)
    p.add_argument('--steps', type=int, default=3)
    p.add_argument('--warmup', type=int, default=3)
    p.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    p.add_argument('--mode', default='bf16', choices=['bf16', 'bf16x3', 'fp32'])
    p.add_argument('--samples', type=int, default=40)
    p.add_argument('--batch-size', type=int, default=20)
    p.add_argument('--inference-steps', type=int, default=20)
    p.add_argument('--workload', default='3dpf_apo')
    p.add_argument('--cpu-samples', type=int, default=4)
    p.add_argument('--cpu-steps', type=int, default=3)
    p.add_argument('--no-cpu-baseline', action='store_true')
    p.add_argument('--single-stream', action='store_true', help='run the mini-batches back to back on one stream (A/B of the two-stream overlap)')
    p.add_argument('--no-graph', action='store_true', help='launch kernels eagerly instead of replaying the captured step graph')
    return p.parse_args()


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        d = json.load(open(path))
        return d.get('bf16_tflops_sustained', 1414.3), d.get('hbm_gbs', 6452.2), 'measured (MEASURED_PEAKS.json, sustained bf16)'
    return 1590.0, 6650.0, 'fallback (B200_PROFILING.md)'


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
             'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')
        while not self.stop_flag:
            try:
                out = subprocess.run(['nvidia-smi', f'--query-gpu={q}', '--format=csv,noheader,nounits', '-i', str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(',')])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.rows:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace('.', '').isdigit())
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith('active') for r in self.rows if len(r) > 2 + i)]
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': float(self.rows[0][1]) if self.rows else None,
                'reasons': reasons, 'samples': len(self.rows)}


def workload(args, seed):
    from diffdock_pocket_b200 import inputs, sampling as S, utils
    g = inputs.load_graph_npz(os.path.join(GOLD, args.workload + '.npz'), name=args.workload)
    sa = utils.score_model_args()
    np.random.seed(seed)
    torch.manual_seed(seed)
    dl = [copy.deepcopy(g) for _ in range(args.samples)]
    S.randomize_position(dl, False, False, sa.tr_sigma_max, flexible_sidechains=True)
    return g, dl


def conv_flops(ns, w_numel):
    return 2.0 * ((3 * ns) * (3 * ns) + (3 * ns) * w_numel + w_numel)


def cpu_arm(args, n_samples, n_steps, threads):
    """Oracle port (pure PyTorch fp32 on the host cores) on a bounded sample, scaled to poses/s of the full workload."""
    from diffdock_pocket_b200 import so3, torus, utils
    from oracle import diffusion_ref as D, factory, sampling_ref as S
    torch.set_num_threads(threads)
    model, conf, sa, ca = utils.build_models(torch.device('cpu'), seed=0)
    om = factory.oracle_model(sa, model.state_dict(), so3.score_norm_np, torus.score_norm)
    _, dl = workload(args, 0)
    dl = dl[:n_samples]
    sch = D.get_t_schedule(args.inference_steps)
    torch.manual_seed(1)
    t0 = time.time()
    S.sampling(dl, om, args.inference_steps, sch, sch, sch, sch, partial(D.t_to_sigma, args=sa), sa, batch_size=args.batch_size,
               max_steps=n_steps, **TEMP)
    dt = time.time() - t0
    full = dt * (args.samples * args.inference_steps) / (n_samples * n_steps)
    return args.samples / full, dt


def run_reference(args):
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return
    threads = os.cpu_count()
    vals = []
    for _ in range(args.warmup if args.warmup < 1 else 1):
        cpu_arm(args, 1, 1, threads)
    t_all = 0.0
    for _ in range(args.steps):
        v, dt = cpu_arm(args, args.cpu_samples, args.cpu_steps, threads)
        vals.append(v)
        t_all += dt
    value = float(np.mean(vals))
    sample = f'{args.cpu_samples} samples x {args.cpu_steps} of {args.inference_steps} steps per bench step, scaled linearly to {args.samples} x {args.inference_steps}'
    print(json.dumps({
        'impl': 'reference', 'metric': 'docked poses/sec (20-step reverse diffusion)', 'value': value, 'unit': 'poses/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1000.0 * args.samples / value,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': f'{args.workload}: {args.samples} samples, batch {args.batch_size}, {args.inference_steps} steps (BASELINE.json configs[1])'},
        'cpu_baseline': {'value': value, 'unit': 'poses/s', 'cores': threads, 'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': 'poses/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}))


def main():
    args = parse()
    if args.impl == 'reference':
        return run_reference(args)
    import torch.distributed as dist
    from diffdock_pocket_b200 import _lib, diffusion_utils as du, sampling as ps, utils
    from diffdock_pocket_b200.hetero import Batch
    world = int(os.environ.get('WORLD_SIZE', 1))
    rank = int(os.environ.get('RANK', 0))
    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        # NCCL prints its version banner on stdout when the communicator comes up: keep stdout for the one JSON line
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group('nccl', device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    model, conf, sa, ca = utils.build_models(dev, seed=0)
    model.conv_mode = conf.conv_mode = args.mode
    g, dl0 = workload(args, rank)
    sch = du.get_t_schedule('expbeta', args.inference_steps)
    t2s = partial(du.t_to_sigma, args=sa)
    kw = dict(confidence_model=conf, filtering_model_args=ca, batch_size=args.batch_size, **TEMP)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def gather_rank(poses, confidence):
        """The single collective of the path: all-gather final poses + confidences, rank by confidence."""
        if world == 1:
            return torch.argsort(confidence, descending=True)
        pg = [torch.empty_like(poses) for _ in range(world)]
        cg = [torch.empty_like(confidence) for _ in range(world)]
        dist.all_gather(pg, poses)
        dist.all_gather(cg, confidence)
        return torch.argsort(torch.cat(cg), descending=True)

    # ---- e2e pass through the public API (host graphs in, host poses out) --------------------------------
    def e2e_step(dl):
        torch.manual_seed(7)
        out, c = ps.sampling(dl, model, args.inference_steps, sch, sch, sch, sch, dev, t2s, sa, concurrent_batches=not args.single_stream, **kw)
        poses = torch.stack([o['ligand'].pos for o in out]).to(dev)
        return gather_rank(poses, c.reshape(-1).to(dev)).cpu()

    # ---- resident pass: plans + pose states (+ captured step graphs) built once, inputs already in HBM ----------
    chunks = [list(range(i, min(i + args.batch_size, args.samples))) for i in range(0, args.samples, args.batch_size)]
    # independent mini-batches alternate between two streams, exactly as sampling() runs them
    streams = [torch.cuda.Stream(device=dev) for _ in range(2)] if len(chunks) > 1 and not args.single_stream else [None]
    with torch.no_grad():
        runners = [ps.StepRunner(model, [dl0[i] for i in idx], True, False, use_graph=not args.no_graph, stream=streams[k % len(streams)])
                   for k, idx in enumerate(chunks)]
        cplans = [conf.make_plan(Batch.from_data_list([dl0[i] for i in idx])) for idx in chunks]
    plans = [r.pl for r in runners]
    init = [(pl.lig_pos.clone(), pl.atom_pos.clone()) for pl in plans]
    N = args.samples
    T_tot, S_tot = sum(r.T for r in runners), sum(r.S for r in runners)
    noise = torch.randn(args.inference_steps, 6 * N + T_tot + S_tot, generator=torch.Generator().manual_seed(3))
    coefs = [ps.step_coefficients(t_idx, args.inference_steps, (sch,) * 4, t2s, sa, False, TEMP['temp_sampling'], TEMP['temp_psi'],
                                  TEMP['temp_sigma_data'], True) for t_idx in range(args.inference_steps)]

    def resident_step():
        with torch.no_grad():
            for pl, (lp, ap) in zip(plans, init):
                pl.lig_pos.copy_(lp)
                pl.atom_pos.copy_(ap)
            for r in runners:
                r.sync_in()
            for t_idx in range(args.inference_steps):
                t, coef = coefs[t_idx]
                z = noise[t_idx]
                s0 = t0 = c0 = 0
                for idx, r in zip(chunks, runners):
                    b = len(idx)
                    row = torch.cat([z[3 * s0:3 * (s0 + b)], z[3 * N + 3 * s0:3 * N + 3 * (s0 + b)], z[6 * N + t0:6 * N + t0 + r.T],
                                     z[6 * N + T_tot + c0:6 * N + T_tot + c0 + r.S]])
                    r.step(t, coef, row)
                    s0, t0, c0 = s0 + b, t0 + r.T, c0 + r.S
            cs = []
            for idx, r, cpl in zip(chunks, runners, cplans):
                with r.ctx():
                    cpl.lig_pos.copy_(r.pl.lig_pos)
                    cpl.atom_pos.copy_(r.pl.atom_pos)
                    zt = torch.zeros(len(idx))
                    cs.append(conf.run_plan(cpl, {'tr': zt, 'rot': zt, 'tor': zt, 'sc_tor': zt}).reshape(-1).clone())
            for r in runners:
                r.sync_out()
            return gather_rank(torch.cat([pl.lig_pos for pl in plans]).reshape(N, -1, 3), torch.cat(cs))

    # gpu_launches: kernels of ONE bench step, counted on an eager (non-graph) pass of identical work
    eager = [ps.StepRunner(model, [dl0[i] for i in idx], True, False, use_graph=False) for idx in chunks[:1]]
    _lib.COUNTS.clear()
    eager[0].step(*coefs[0], None)
    launches_per_step = _lib.launch_count() * len(chunks) * args.inference_steps
    _lib.COUNTS.clear()
    with torch.no_grad():
        zt = torch.zeros(len(chunks[0]))
        conf.run_plan(cplans[0], {'tr': zt, 'rot': zt, 'tor': zt, 'sc_tor': zt})
    launches_per_step += _lib.launch_count() * len(chunks)
    del eager
    for _ in range(args.warmup):
        resident_step()
    flush.fill_(1)
    sampler = ClockSampler(local)
    sampler.start()
    _lib.COUNTS.clear()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        resident_step()
        flush.fill_(0)                           # L2 flush between timed iterations (256 MiB > 126 MB L2)
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1) / args.steps
    launches = launches_per_step
    sampler.stop_flag = True
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * args.samples / (ms / 1000.0)

    # e2e: same metric through sampling() with host buffers
    n_e2e = max(1, min(args.steps, 3))
    e2e_inputs = [copy.deepcopy(dl0) for _ in range(n_e2e + 2)]      # host graphs (sampling() updates them in place)
    for _ in range(2):                                               # untimed: lazy initialisation, allocator, page cache
        e2e_step(e2e_inputs.pop())
    barrier()
    t0 = time.time()
    for _ in range(n_e2e):
        e2e_step(e2e_inputs.pop())
    barrier()
    e2e_ms = (time.time() - t0) * 1000.0 / n_e2e
    if world > 1:
        t = torch.tensor([e2e_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    h2d = sum(t.numel() * t.element_size() for pl in plans for t in (pl.lig_pos, pl.atom_pos, pl.rec_pos, pl.lig_static, pl.atom_static, pl.rec_static))
    h2d += sum(pl.NR * 1281 * 4 for pl in plans) + noise.numel() * 4 + sum(pl.step_in.numel() * 4 for pl in plans) * args.inference_steps
    d2h = sum((pl.NL + pl.NA) * 12 for pl in plans) + N * 4

    # ---- roofline of the dominant kernel: instrumented forward (events around every fused conv launch) -----
    roof = None
    if rank == 0:
        peak_tf, hbm, peak_src = peaks()
        pl = plans[0]
        ct = {k: torch.full((len(chunks[0]),), 0.5) for k in ('tr', 'rot', 'tor', 'sc_tor')}
        model.profile = []
        with torch.no_grad():
            model.run_plan(pl, ct)
        torch.cuda.synchronize()
        tot_ms, tot_fl, n_l = 0.0, 0.0, 0
        for (e0, e1, convs) in model.profile:
            tot_ms += e0.elapsed_time(e1)
            tot_fl += sum(conv_flops(ns, w_numel) * int(es.n_dev.item()) for (w_numel, es, ns) in convs)
            n_l += 1
        model.profile = None
        ach = tot_fl / (tot_ms * 1e-3) / 1e12 if tot_ms > 0 else 0.0
        # DRAM bytes of one conv launch (dram__bytes_read + write, ncu --set full of the grouped layer-3 launch of this
        # same batch: profiles/r1_ncu_umma_v27_summary.txt); far below the algorithmic 1.2 kB/edge because gathered
        # node rows and the weight image are served by L2
        traffic = 224.1e6 if (args.mode == 'bf16' and args.workload == '3dpf_apo' and args.batch_size == 20) else None
        roof = {'bound': 'tensor', 'achieved': ach, 'peak': peak_tf, 'unit': 'TFLOP/s', 'frac': ach / peak_tf, 'traffic': traffic,
                'kernel': 'tpconv_umma_kernel' if args.mode != 'fp32' else 'tpconv_fp32_kernel', 'launches_measured': n_l,
                'peak_source': peak_src, 'conv_share_of_forward_ms': tot_ms}
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, dt = cpu_arm(args, args.cpu_samples, args.cpu_steps, os.cpu_count())
        cpu = {'value': v, 'unit': 'poses/s', 'cores': os.cpu_count(), 'kind': 'port',
               'sample': f'{args.cpu_samples} samples x {args.cpu_steps} of {args.inference_steps} steps ({dt:.1f} s), scaled linearly'}
    if rank == 0:
        print(json.dumps({
            'metric': 'docked poses/sec (20-step reverse diffusion)', 'value': value, 'unit': 'poses/s', 'n_gpus': world,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': {'bf16': 'bf16', 'bf16x3': 'bf16x3 (fp32-grade)', 'fp32': 'f32'}[args.mode], 'data': 'synthetic',
            'config': {'workload': f'{args.workload}: {args.samples} samples, batch {args.batch_size}, {args.inference_steps} reverse-diffusion steps + confidence pass, per GPU (BASELINE.json configs[1])',
                       'weights': 'random init of the README big score model (ns=60 nv=10 6 layers lmax=1) + confidence model (no checkpoint offline)', 'conv_mode': args.mode, 'streams': len(streams),
                       'l2': 'weights + activations (>400 MB) exceed L2; 256 MiB flush between timed iterations'},
            'clocks': sampler.summary(), 'gpu_launches': launches,
            'e2e': {'value': world * args.samples / (e2e_ms / 1000.0), 'unit': 'poses/s', 'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(d2h)},
            'roofline': roof, 'cpu_baseline': cpu}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
