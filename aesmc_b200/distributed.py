"""Data-parallel plumbing: independent observation sequences (batch rows) are sharded across ranks,
one process per GPU.  There is no communication inside infer(); the only collectives
are (1) one flattened-gradient all-reduce per optimiser step (train.py, between backward() and
step()) and (2) the final log-evidence / ESS reductions.  Works with the nccl backend on GPUs and
with gloo on CPU tensors (used by the world_size-2 tests).
"""
import torch
import torch.distributed as dist


def is_active():
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_bounds(batch_size, rank=None, world_size=None):
    """Rows [lo, hi) owned by ``rank``: contiguous blocks, the first (batch_size % world) ranks get one
    extra row, so any batch size works and results are indexed by GLOBAL row."""
    if rank is None or world_size is None:
        rank, world_size = world()
    base, extra = divmod(batch_size, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(value, rank=None, world_size=None, dim=0):
    """Slice the batch axis of a tensor / list / dict of per-step observations for this rank."""
    if isinstance(value, dict):
        return {k: shard_batch(v, rank, world_size, dim) for k, v in value.items()}
    if isinstance(value, (list, tuple)):
        return type(value)(shard_batch(v, rank, world_size, dim) for v in value)
    lo, hi = shard_bounds(value.size(dim), rank, world_size)
    return value.narrow(dim, lo, hi - lo)


def shard_uniforms(uniforms, rank=None, world_size=None):
    """uniforms [T-1, B_global] -> this rank's columns, so resampling is invariant to world size."""
    lo, hi = shard_bounds(uniforms.shape[1], rank, world_size)
    return uniforms[:, lo:hi]


def all_reduce_gradients(parameters, local_batch, global_batch):
    """Sum gradients across ranks so that the result is the gradient of the GLOBAL batch-mean loss:
    each rank's loss is a mean over its local rows, hence the local_batch/global_batch weighting.
    One flattened all-reduce (the messages are tiny for these models: latency-bound)."""
    params = [p for p in parameters if p.grad is not None]
    if not params:
        return
    scale = float(local_batch) / float(global_batch)
    if not is_active():
        return
    flat = torch.cat([p.grad.reshape(-1) for p in params]) * scale
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    offset = 0
    for p in params:
        n = p.grad.numel()
        p.grad.copy_(flat[offset:offset + n].view_as(p.grad))
        offset += n


def gather_rows(local, global_batch):
    """All-gather per-row results (log-evidence, ESS) into global row order; ragged shards allowed."""
    if not is_active():
        return local
    rank, world_size = world()
    sizes = [shard_bounds(global_batch, r, world_size) for r in range(world_size)]
    width = max(hi - lo for lo, hi in sizes)
    padded = local.new_zeros((width,) + tuple(local.shape[1:]))
    padded[:local.shape[0]] = local
    pieces = [torch.empty_like(padded) for _ in range(world_size)]
    dist.all_gather(pieces, padded)
    return torch.cat([p[:hi - lo] for p, (lo, hi) in zip(pieces, sizes)], dim=0)


def global_mean(local_rows, global_batch):
    """Mean over the global batch of a per-row quantity held shard-wise (one scalar all-reduce)."""
    total = local_rows.sum()
    if is_active():
        dist.all_reduce(total, op=dist.ReduceOp.SUM)
    return total / global_batch
