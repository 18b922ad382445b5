"""Multi-GPU plumbing: one process per GPU, complexes / samples sharded, ONE collective at the end.

The reference's multi-device inference is a spawn pool with ``np.array_split`` of the complex table and no
communication (inference.py:466-488).  Here each ``torch.distributed`` rank takes a contiguous shard
(``shard_range``, same split rule as ``np.array_split``), runs the whole reverse-diffusion loop locally with
no data-path collective, and the final ligand poses + confidences are all-gathered once (NCCL over NVLink on
GPUs, gloo in the CPU tests) so every rank can rank the poses (inference.py:216-219).
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_range(n_items, rank, world):
    """[start, stop) of rank's shard; identical partition to ``np.array_split(range(n_items), world)``."""
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def gather_and_rank(poses, confidence, group=None):
    """poses [n_local, n_atoms, 3], confidence [n_local] -> (all poses, all confidences, order) on every rank.
    Shards may have different sizes: they are padded to the largest shard for the fixed-size all_gather."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return poses, confidence, torch.argsort(confidence, descending=True)
    n_local = torch.tensor([poses.shape[0]], device=poses.device)
    sizes = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(sizes, n_local, group=group)
    sizes = [int(s.item()) for s in sizes]
    m = max(sizes)
    pp = torch.zeros((m,) + tuple(poses.shape[1:]), dtype=poses.dtype, device=poses.device)
    cc = torch.full((m,), float('-inf'), dtype=confidence.dtype, device=confidence.device)
    pp[:poses.shape[0]] = poses
    cc[:confidence.shape[0]] = confidence
    pg = [torch.empty_like(pp) for _ in range(world)]
    cg = [torch.empty_like(cc) for _ in range(world)]
    dist.all_gather(pg, pp, group=group)
    dist.all_gather(cg, cc, group=group)
    all_p = torch.cat([p[:s] for p, s in zip(pg, sizes)])
    all_c = torch.cat([c[:s] for c, s in zip(cg, sizes)])
    return all_p, all_c, torch.argsort(all_c, descending=True)
