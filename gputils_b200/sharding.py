"""Sharding of the `mats` axis across the GPUs of one box: host-side logic only.

Every batched op is independent per matrix, so rank g of G owns the contiguous block
[g*ceil(k/G), min(k, (g+1)*ceil(k/G))) of EVERY operand and runs the unchanged single-GPU launchers on it; there
is no collective on the data path. torch.distributed (NCCL on GPUs, gloo in the CPU tests) is used only to
all-gather result shards when the caller wants them everywhere and to combine the partial sums of the flat
reductions (normF / sumAbs / dotF are sums over all matrices; maxAbs / minAbs are max / min).
"""
from __future__ import annotations

import math


def shard_range(k: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous block partition of k matrices over `world` ranks; trailing ranks may be short or empty."""
    if world < 1 or not (0 <= rank < world) or k < 0:
        raise ValueError("bad shard arguments")
    per = math.ceil(k / world) if k else 0
    start = min(k, rank * per)
    return start, min(k, start + per)


def shard_sizes(k: int, world: int) -> list[int]:
    return [b - a for a, b in (shard_range(k, world, r) for r in range(world))]


def all_gather_shards(local, k: int, group=None):
    """Gathers the per-rank shards (first axis = matrices) into the full (k, ...) tensor on every rank.
    NCCL's all_gather needs equal counts, so shards are padded to ceil(k/G) and the padding dropped."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    per = math.ceil(k / world) if k else 0
    padded = torch.zeros((per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    padded[: local.shape[0]] = local
    out = torch.empty((world * per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, padded, group=group)
    pieces = [out[r * per: r * per + n] for r, n in enumerate(shard_sizes(k, world))]
    return torch.cat(pieces, dim=0)


def global_norm_f(local_sumsq: float, device="cpu", group=None) -> float:
    """Frobenius norm of a sharded tensor from the per-shard sums of squares (one scalar all-reduce)."""
    import torch
    import torch.distributed as dist
    t = torch.tensor([local_sumsq], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return float(t.sqrt())


def global_max_abs(local_max: float, device="cpu", group=None) -> float:
    import torch
    import torch.distributed as dist
    t = torch.tensor([local_max], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t)
