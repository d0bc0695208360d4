"""Multi-GPU plumbing for the IAF step: the batch is sharded across ranks (one process per
GPU), weights are replicated, and the only collective is a sum all-reduce of the ELBO
scalars (tf_train.py:126-142: ``tf.split`` over towers and ``tf.add_n(losses)``).  The
free-bits batch mean stays rank-local exactly as it is tower-local in the reference
(tf_train.py:79 runs inside the per-tower ``_forward``)."""
import torch
import torch.distributed as dist


def shard_range(n, rank, world):
    """Contiguous [lo, hi) slice of n samples owned by ``rank`` (even split, remainder to low ranks)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_scalars(values, group=None):
    """Sum-all-reduce a short list of scalar tensors in ONE collective (NCCL on GPU, gloo on CPU)."""
    dtype = torch.float64 if any(v.dtype == torch.float64 for v in values) else torch.float32
    buf = torch.stack([v.reshape(()).to(dtype) for v in values])
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    return list(buf.unbind(0))


def allreduce_grads(params, group=None, average=True):
    """Data-parallel training: the towers' gradients are averaged before the optimiser step
    (``average_grads``, tf_utils/common.py:78-115, called at tf_train.py:139-147).  All gradients travel in ONE flat
    bucket per dtype -- a few MB for the whole model, far below where NVLink bandwidth matters, so a single collective
    beats per-tensor calls on launch latency alone.  ``params``: iterable (or dict values) of tensors with ``.grad``;
    tensors without a gradient are skipped on every rank alike.  In place; returns the number of buckets reduced."""
    if isinstance(params, dict):
        params = [params[k] for k in sorted(params)]
    grads = [p.grad for p in params if p.grad is not None]
    if not grads or not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return 0
    world = dist.get_world_size(group)
    n = 0
    for dtype in sorted({g.dtype for g in grads}, key=str):
        gs = [g for g in grads if g.dtype == dtype]
        flat = torch.cat([g.reshape(-1) for g in gs])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        if average:
            flat.div_(world)
        off = 0
        for g in gs:
            g.copy_(flat[off:off + g.numel()].view_as(g))
            off += g.numel()
        n += 1
    return n
