"""Multi-GPU layout: cell instances are independent, so a job shards them across ranks (one process per GPU) with
no data-path communication. Instance g of the job uses Philox key seed_base + g, whatever the number of ranks,
so results do not depend on how the job is sharded. The only collective is the reduction of the episode
statistics vector (include/nbiot.h: nbiot_reduce_stats): SUM for the counters, MAX for the high-water marks, one all-gather."""
import numpy as np

from ._cabi import N_STATS, N_SUM_STATS


def shard_range(total, rank, world):
    """contiguous block [lo, hi) of instance indices owned by `rank`"""
    base, rem = divmod(int(total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_seeds(total, rank, world, seed_base=0):
    lo, hi = shard_range(total, rank, world)
    return np.arange(lo, hi, dtype=np.uint64) + np.uint64(seed_base)


def allreduce_stats(stats, group=None):
    """In-place all-reduce of an int64[N_STATS] tensor (CUDA with NCCL, CPU with gloo)."""
    import torch.distributed as dist
    assert stats.numel() == N_STATS
    if not (dist.is_available() and dist.is_initialized()):
        return stats
    # ONE collective: every rank's 24 values are gathered (192 bytes per rank: latency-bound either way), then the counters are
    # summed and the high-water marks maximised locally
    import torch
    world = dist.get_world_size(group)
    parts = [torch.empty_like(stats) for _ in range(world)]
    dist.all_gather(parts, stats.contiguous(), group=group)
    gathered = torch.stack(parts)
    stats[:N_SUM_STATS] = gathered[:, :N_SUM_STATS].sum(dim=0)
    stats[N_SUM_STATS:] = gathered[:, N_SUM_STATS:].max(dim=0).values
    return stats
