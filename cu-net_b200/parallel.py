"""Data parallelism: one process per GPU, one allreduce of the flat gradient bucket per step.

Replaces ``torch.nn.DataParallel`` (cu-net.py:59): the reference scatters ``--bs`` over the GPUs, re-broadcasts
all parameters every step, gathers outputs to GPU 0 and reduces gradients onto GPU 0.  Here every rank owns a full
replica, computes its shard's loss locally (local BatchNorm statistics, exactly like DataParallel's per-replica
statistics -- no SyncBN), pre-scales dLoss by 1/world and sum-allreduces ONE contiguous fp32 bucket over
NCCL/NVLink; the replicated RMSprop then applies identical updates everywhere.  With equal shards the averaged
gradient equals the reference's gather-then-global-mean gradient.
"""
import os


def env_world():
    """(rank, world_size, local_rank) from the torchrun environment."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def shard_batch(global_batch, rank, world, weak=False):
    """Per-rank batch.  weak=False: the reference's semantics, --bs is the GLOBAL batch (DataParallel scatter on
    dim 0, cu-net.py:59,84); weak=True: --bs per GPU.  Returns (start, size)."""
    if weak:
        return rank * global_batch, global_batch
    if global_batch % world:
        raise ValueError("global batch %d is not divisible by %d ranks (unequal shards would change the "
                         "gradient weighting relative to the reference)" % (global_batch, world))
    per = global_batch // world
    return rank * per, per


def allreduce_mean(flat, world, group=None):
    """Gradients that were computed with dLoss pre-scaled by 1/world: a SUM allreduce yields the mean."""
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return flat


def broadcast_params(flat, world, group=None, src=0):
    if world > 1:
        import torch.distributed as dist
        dist.broadcast(flat, src, group=group)
    return flat


def shard_range(n, rank, world):
    """[start, end) of rank's slice of a flat buffer of n elements (n is padded to a multiple of world by the
    engine, see Engine.n_params_padded)."""
    if n % world:
        raise ValueError("flat buffer of %d elements is not divisible by %d ranks" % (n, world))
    per = n // world
    return rank * per, (rank + 1) * per


def reduce_scatter_mean(flat, shard_out, world, group=None):
    """the gradient exchange as reduce-scatter -- every rank receives only the summed slice
    it will update (gradients pre-scaled by 1/world: sum == mean).  shard_out: [n / world]."""
    if world > 1:
        import torch.distributed as dist
        dist.reduce_scatter_tensor(shard_out, flat, op=dist.ReduceOp.SUM, group=group)
    else:
        shard_out.copy_(flat)
    return shard_out


def all_gather_params(flat, shard, world, group=None):
    """The updated parameter slices back into every rank's full flat parameter buffer."""
    if world > 1:
        import torch.distributed as dist
        dist.all_gather_into_tensor(flat, shard, group=group)
    else:
        flat.copy_(shard)
    return flat
