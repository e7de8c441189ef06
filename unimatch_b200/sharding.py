"""Multi-GPU layout of the matching path: image pairs are independent, so a batch is sharded
contiguously across ranks (one process per GPU, full weight replica each) and the only collective is the gather of
the final predictions (NCCL over NVLink on the GPU box, gloo in the CPU tests)."""
import torch
import torch.distributed as dist


def shard_range(batch, rank, world):
    """[start, stop) of the pairs rank `rank` processes; remainders go to the first ranks."""
    base, rem = divmod(batch, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def gather_predictions(local, batch=None, async_op=False):
    """All ranks receive the predictions of the whole batch, in pair order.  `local`: [b_local, ...].

    `batch` = global number of pairs when the shards come from `shard_range` (equal shards take the single
    `all_gather_into_tensor`); without it the shard sizes are exchanged first, so unequal local batches can neither hang
    nor be silently mis-assembled.  `async_op=True` (equal shards only) returns (output, work handle): the collective runs on
    the communicator's stream, ordered after the current stream, and the caller waits when it needs the result -- this is
    how bench.py keeps the exchange off the critical path of the next forward."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return (local, None) if async_op else local
    world = dist.get_world_size()
    local = local.contiguous()
    if batch is None:
        n = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
        counts = [torch.zeros_like(n) for _ in range(world)]
        dist.all_gather(counts, n)
        counts = [int(c.item()) for c in counts]
    else:
        counts = [b - a for a, b in (shard_range(batch, r, world) for r in range(world))]
        if counts[dist.get_rank()] != local.shape[0]:
            raise ValueError("gather_predictions: local batch %d is not this rank's shard (%d of %d pairs)"
                             % (local.shape[0], counts[dist.get_rank()], batch))
    if len(set(counts)) == 1:
        out = torch.empty((world * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        work = dist.all_gather_into_tensor(out, local, async_op=async_op)
        return (out, work) if async_op else out
    if async_op:
        raise ValueError("gather_predictions: async_op needs equal shards")
    pad = max(counts)
    buf = torch.zeros((pad,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    buf[:local.shape[0]] = local
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf)
    return torch.cat([p[:c] for p, c in zip(parts, counts)], dim=0)
