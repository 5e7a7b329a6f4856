"""Data-parallel exchange of the path: BatchNorm statistics stay per rank, the only collective is the reduction of the
flat fp32 gradient buffer.  `GradientBuckets` splits that buffer into the contiguous slices that become final at three points of the
backward pass (head / neck / backbone weights; BatchNorm parameters and prediction biases with the last one) and launches each slice's NCCL
all-reduce on a communication stream as soon as its range has finished, so the transfer overlaps the rest of backward -- what detectron2's
DistributedDataParallel does with 25 MB buckets and autograd hooks, here on a static plan.  NCCL over NVLink / NVSwitch on GPUs; gloo on
the CPU for the host-logic tests.
"""
import torch
import torch.distributed as dist


def allreduce_gradients(flat_grad: torch.Tensor, average: bool = True, group=None):
    """in-place sum (or mean, as DDP) of the flat gradient buffer over all ranks"""
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


def bucket_slices(param_layout, total, order=("head.", "neck.", "backbone.")):
    """Element ranges of the flat gradient buffer that are complete after the backward of each part, in backward order.
    param_layout = [(name, offset, numel)] ascending; the convolution weights of a part are contiguous (plan order), the BatchNorm
    parameters / prediction biases follow all weights and are reduced with the LAST bucket.  Returns [[(lo, hi), ...] per part]."""
    is_weight = lambda n: n.endswith(".conv.weight") or ("_preds." in n and n.endswith(".weight"))
    first_tail = min((off for n, off, _ in param_layout if not is_weight(n)), default=total)
    out = []
    for prefix in order:
        offs = [(off, off + cnt) for n, off, cnt in param_layout if n.startswith(prefix) and is_weight(n) and off < first_tail]
        out.append([(min(o[0] for o in offs), max(o[1] for o in offs))] if offs else [])
    if first_tail < total:
        out[-1].append((first_tail, total))
    # padding between tensors belongs to the bucket that contains it; the slices must be disjoint and inside the buffer
    flat = sorted(r for b in out for r in b)
    assert all(a[1] <= b[0] for a, b in zip(flat, flat[1:])) and (not flat or flat[-1][1] <= total)
    return out


class GradientBuckets:
    """Bucketed, overlapped gradient all-reduce for one YoloxEngine plan.

        gb = GradientBuckets(engine)                # after dist.init_process_group
        gb.step_backward()                          # instead of engine.backward(): backward range by range, reduce as ranges finish
        gb.wait()                                   # the compute stream waits for the reductions (before the optimizer step)

    The sum over ranks is left in flat_grad (the mean's 1/world is folded into FlatOptimizer.grad_scale)."""

    PARTS = ("head", "neck", "backbone")

    def __init__(self, engine, group=None):
        self.eng, self.group = engine, group
        self.slices = bucket_slices(engine.param_layout, engine.flat_grad.numel())
        self.views = [[engine.flat_grad[lo:hi] for lo, hi in part] for part in self.slices]
        self.comm = torch.cuda.Stream(device=engine.dev) if engine.flat_grad.is_cuda else None
        self.events = [torch.cuda.Event() for _ in self.PARTS] if self.comm is not None else None
        self.works = []

    def reduce_part(self, i):
        """enqueue the all-reduce of bucket i behind everything issued so far on the current stream"""
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(self.group) == 1:
            return
        if self.comm is None:
            for v in self.views[i]:
                dist.all_reduce(v, group=self.group)
            return
        self.events[i].record(torch.cuda.current_stream())
        with torch.cuda.stream(self.comm):
            self.comm.wait_event(self.events[i])
            for v in self.views[i]:
                self.works.append(dist.all_reduce(v, group=self.group, async_op=True))

    def step_backward(self, accumulate=False):
        eng = self.eng
        for i, part in enumerate(self.PARTS):
            eng.backward(accumulate, eng.ranges[part], fresh=(i == 0))
            self.reduce_part(i)

    def wait(self):
        for w in self.works:
            w.wait()  # the current stream waits for the collective
        self.works = []
