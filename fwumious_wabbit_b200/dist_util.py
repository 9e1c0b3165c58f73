"""Multi-GPU plumbing: one process per GPU, `torch.distributed` for barriers and the max-over-ranks
timing only.  The hot path itself has no exchange step: every rank trains an independent replica on
its own shard of the example stream (DESIGN.md "multi-GPU"), so no data-path collective exists."""
import os


def env_world():
    return int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))


def init(backend):
    """Initialise torch.distributed from the torchrun environment; returns (world, rank, local_rank, dist or None)."""
    world, rank, local_rank = env_world()
    if world == 1:
        return world, rank, local_rank, None
    import torch
    import torch.distributed as dist

    kw = {}
    if backend == "nccl":
        torch.cuda.set_device(local_rank)
        kw["device_id"] = torch.device("cuda", local_rank)
    dist.init_process_group(backend=backend, **kw)
    return world, rank, local_rank, dist


def shard(rank, world, n_per_rank):
    """Rank r trains on examples [r * n, (r + 1) * n) of the synthetic stream: disjoint, contiguous shards."""
    return rank * n_per_rank, n_per_rank


def max_over_ranks(x, dist, device="cpu"):
    if dist is None:
        return float(x)
    import torch

    t = torch.tensor([float(x)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def whole_job_rate(units_per_rank_per_step, steps, world, max_ms):
    """Aggregate throughput of the whole job: all ranks' units over the slowest rank's time."""
    return world * units_per_rank_per_step * steps / (max_ms * 1e-3)
