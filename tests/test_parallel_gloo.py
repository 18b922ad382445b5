"""world_size-2 gloo test (CPU) of the multi-GPU host logic: shard partition + the single final all-gather."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from diffdock_pocket_b200.parallel import gather_and_rank, shard_range


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_items, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    a, b = shard_range(n_items, rank, world)
    g = torch.Generator().manual_seed(0)
    poses = torch.randn(n_items, 5, 3, generator=g)
    conf = torch.randn(n_items, generator=g)
    all_p, all_c, order = gather_and_rank(poses[a:b], conf[a:b])
    ok = torch.equal(all_p, poses) and torch.equal(all_c, conf) and torch.equal(order, torch.argsort(conf, descending=True))
    q.put((rank, a, b, bool(ok)))
    dist.destroy_process_group()


def test_shard_range_matches_array_split():
    for n in (0, 1, 7, 40, 363):
        for w in (1, 2, 4, 8):
            want = [list(x) for x in np.array_split(np.arange(n), w)]
            got = [list(range(*shard_range(n, r, w))) for r in range(w)]
            assert got == want


def test_gather_and_rank_world2_gloo():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 7, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, 0, 4, True), (1, 4, 7, True)]


def test_inference_defaults_match_reference_parser():
    """The sampler flags that reach the hot path carry inference.py's defaults (:75-101)."""
    from diffdock_pocket_b200 import inference
    a = inference.default_args()
    assert (a.samples_per_complex, a.batch_size, a.inference_steps) == (10, 32, 30)
    assert abs(a.temp_sampling_tr - 0.9766350103728372) < 1e-15 and abs(a.temp_psi_sc_tor - 1.339614553802453) < 1e-15
    assert abs(a.temp_sigma_data - 0.48884149503636976) < 1e-15
