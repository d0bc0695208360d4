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
