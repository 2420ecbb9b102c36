"""Multi-view data parallelism (not present in the reference, which is single
process / single GPU).  One process per GPU, Gaussians replicated, one camera per rank per
step; the only exchange is a SUM all-reduce of the parameter gradients (236 B/Gaussian), issued
as few large messages so NCCL moves them over NVLink/NVSwitch at full bandwidth.
Backend-agnostic: `nccl` on the GPUs, `gloo` in the CPU tests."""
import torch
import torch.distributed as dist


def view_index(step, rank, world):
    """camera index rendered by `rank` at `step` (8 cameras/step at world=8, config 5)"""
    return step * world + rank


def _coalesce(tensors):
    """Group tensors that are contiguous views tiling one storage range (the fused backward
    writes dL/dpws, dL/dshs, dL/dscales, dL/drots into one flat bucket) so the group can be
    reduced in place with a single collective and no staging copy."""
    groups, rest = {}, []
    for t in tensors:
        if t.is_contiguous() and t.numel() > 0:
            groups.setdefault((t.untyped_storage().data_ptr(), t.dtype), []).append(t)
        else:
            rest.append(t)
    flats = []
    for (_, dtype), ts in groups.items():
        ts.sort(key=lambda t: t.storage_offset())
        lo, hi = ts[0].storage_offset(), ts[0].storage_offset()
        ok = True
        for t in ts:
            if t.storage_offset() != hi:
                ok = False
                break
            hi += t.numel()
        if ok and len(ts) > 1:
            flats.append(torch.as_strided(ts[0], (hi - lo,), (1,), lo))
        else:
            rest.extend(ts)
    return flats, rest


def allreduce_grads(tensors, group=None, average=False):
    """In-place SUM (or mean) all-reduce of a list of gradient tensors.  Tensors that already
    share a flat bucket are reduced in place; the others go through one concatenated bucket.
    Returns the number of bytes reduced."""
    tensors = [t for t in tensors if t is not None]
    if not tensors:
        return 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    flats, rest = _coalesce(tensors)
    nbytes = 0
    for f in flats:
        if world > 1:
            dist.all_reduce(f, op=dist.ReduceOp.SUM, group=group)
            if average:
                f /= world
        nbytes += f.numel() * f.element_size()
    if rest:
        flat = torch.cat([t.reshape(-1) for t in rest])
        if world > 1:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
            if average:
                flat /= world
        off = 0
        for t in rest:
            n = t.numel()
            t.copy_(flat[off:off + n].view_as(t))
            off += n
        nbytes += flat.numel() * flat.element_size()
    return nbytes
