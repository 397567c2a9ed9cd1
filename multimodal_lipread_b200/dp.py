"""Data-parallel plumbing: one process per GPU, replicas of every weight, clips sharded across ranks, ONE
allreduce of the flat gradient per step.  The reference has no distributed code
(`torch.distributed` is never imported there); the DP oracle is "run the reference on each rank's shard with the
same weights and average the gradients".

torch.distributed is used for the rendezvous and the collective only (NCCL over NVLink on the GPU box, gloo on
CPU for the tests); the 1/world scaling is folded into the Adam kernel (`lr_adam_step(grad_scale=1/world)`).
"""
import os

import torch
import torch.distributed as dist


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def init(backend=None, device=None):
    """Join the job described by RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT (torchrun).  No-op for world 1."""
    rank, local_rank, world = env_rank_world()
    if world > 1 and not dist.is_initialized():
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        kw = {"device_id": device} if (backend == "nccl" and device is not None) else {}
        dist.init_process_group(backend, **kw)
    return rank, local_rank, world


def shard_range(n_clips, rank, world):
    """Contiguous clip range [lo, hi) of `rank`: rank r gets clips [r*B, (r+1)*B) of a global batch.
    A remainder is spread over the first ranks so that no clip is dropped."""
    base, rem = divmod(n_clips, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class GradAllReduce:
    """Callable handed to `model.train_step(grad_allreduce=...)`: sums the flat gradient over the ranks in place.
    Averaging (1/world) is left to the consumer so that it costs no extra pass."""

    def __init__(self, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.calls = 0

    def __call__(self, flat_grad):
        if self.world > 1:
            dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=self.group)
        self.calls += 1
        return flat_grad


def broadcast_parameters(flat_params, src=0, group=None):
    """Make every replica start from rank `src`'s weights (one broadcast of the flat buffer)."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.broadcast(flat_params, src=src, group=group)
    return flat_params


def max_over_ranks(values, device=None):
    """Element-wise max of a list of floats over all ranks (job time = slowest rank)."""
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.tolist()


def job_throughput(units_per_rank_per_step, steps, ms_total_max, world):
    """Whole-job units/s: everything all ranks processed divided by the slowest rank's time."""
    return units_per_rank_per_step * world * steps / (ms_total_max / 1e3)
