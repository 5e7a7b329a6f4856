"""Data-parallel exchange: one all-reduce of the flat fp32 gradient buffer per step (BatchNorm statistics
stay per rank, the only collective of the path is the gradient reduction that detectron2's DDP performs in 25 MB buckets).
NCCL over NVLink/NVSwitch on GPUs; the same function runs on gloo for the CPU tests."""
import torch
import torch.distributed as dist


def allreduce_gradients(flat_grad: torch.Tensor, average: bool = True, group=None):
    """in-place sum (or mean, as DDP) of the flat gradient buffer over all ranks; returns the async work handle's result"""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return flat_grad
    dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=group)
    if average:
        flat_grad.div_(dist.get_world_size(group))
    return flat_grad


def shard_batch(global_batch: int, rank: int, world: int):
    """contiguous, even split of a global batch (SOLVER.IMS_PER_BATCH) across ranks: [start, stop)"""
    if global_batch % world != 0:
        raise ValueError(f"global batch {global_batch} is not divisible by world size {world}")
    per = global_batch // world
    return rank * per, (rank + 1) * per
