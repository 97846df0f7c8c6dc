"""Multi-GPU plumbing: envs are independent, so the path shards by global env index with NO per-step
collective; the only exchange is one all-reduce(sum) of the 8-double statistics block per rollout
(episodes, return sum, goals, env-steps, contacts; include/msoc.h msoc_stats) -- NCCL over NVLink on the
GPUs, gloo in the CPU tests.  The reference has no distributed code at all."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n_total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous shard [lo, hi) of rank `rank`: the first n_total % world ranks own one extra env."""
    base, rem = divmod(int(n_total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_stats(stats: torch.Tensor) -> torch.Tensor:
    """Sum the per-rank statistics block in place across the default process group (no-op without one)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
    return stats


def make_sharded_sim(n_total: int, config=None, seed: int = 0, device=None):
    """One BatchedSoccerSim per rank over its shard; the Philox spawn streams are keyed by the GLOBAL env
    index, so the union of the shards is bit-identical to a single-GPU run of n_total envs."""
    from .sim import BatchedSoccerSim
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    lo, hi = shard_range(n_total, rank, world)
    if device is None:
        device = f"cuda:{torch.cuda.current_device()}"
    return BatchedSoccerSim(hi - lo, config=config, device=device, seed=seed, global_env_offset=lo)
