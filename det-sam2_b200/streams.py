"""Multi-GPU = independent video streams: the memory bank, prompts and outputs are per
session, so stream s runs on rank s mod world_size, weights are replicated, and NO collective sits on
the data path.  ``torch.distributed`` (host-side gloo: no NCCL communicator is ever created) is used only for the
timing barrier and the max-over-ranks reduction of the measured time.
"""
import os

import torch


def dist_env():
    """(rank, local_rank, world_size) from the torchrun environment (1 process per GPU)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def assign_streams(num_streams, world_size):
    """stream id -> rank (round robin).  Returns one list of stream ids per rank; a partition."""
    if num_streams < 0 or world_size < 1:
        raise ValueError("num_streams >= 0 and world_size >= 1 required")
    return [list(range(r, num_streams, world_size)) for r in range(world_size)]


def barrier(device=None):
    import torch.distributed as dist
    if device is not None and torch.device(device).type == "cuda":
        torch.cuda.synchronize(device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
    if device is not None and torch.device(device).type == "cuda":
        torch.cuda.synchronize(device)


def max_over_ranks(value, device="cpu"):
    """Largest `value` (a float, e.g. elapsed ms) over all ranks; identity for a single process."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def aggregate_throughput(units_per_rank, elapsed_ms_this_rank, device="cpu"):
    """Whole-job throughput: units processed by ALL ranks / slowest rank's time (weak scaling: every
    rank processes `units_per_rank`)."""
    import torch.distributed as dist
    world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
    ms = max_over_ranks(elapsed_ms_this_rank, device)
    return world * units_per_rank / (ms / 1e3), ms
