"""Multi-view data parallelism (not present in the reference, which is single
process / single GPU).  One process per GPU, Gaussians replicated, one camera per rank per
step; the only exchange is a SUM all-reduce of the parameter gradients, done on one
flattened bucket (236 B/Gaussian) so NCCL sees a single large message over NVLink/NVSwitch.
Backend-agnostic: `nccl` on the GPUs, `gloo` in the CPU tests."""
import torch
import torch.distributed as dist


def view_index(step, rank, world):
    """camera index rendered by `rank` at `step` (8 cameras/step at world=8, config 5)"""
    return step * world + rank


def allreduce_grads(tensors, group=None, average=False):
    """In-place SUM (or mean) all-reduce of a list of gradient tensors through one flat bucket.
    Returns the number of bytes reduced."""
    tensors = [t for t in tensors if t is not None]
    if not tensors:
        return 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    flat = torch.cat([t.reshape(-1) for t in tensors])
    if world > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if average and world > 1:
        flat /= world
    off = 0
    for t in tensors:
        n = t.numel()
        t.copy_(flat[off:off + n].view_as(t))
        off += n
    return flat.numel() * flat.element_size()
