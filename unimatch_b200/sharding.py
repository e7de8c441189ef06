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


def gather_predictions(local, batch=None):
    """All ranks receive the predictions of the whole batch, in pair order.  `local`: [b_local, ...]."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    local = local.contiguous()
    if batch is None or batch % world == 0:
        out = torch.empty((world * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local)
        return out
    sizes = [shard_range(batch, r, world) for r in range(world)]
    pad = max(b - a for a, b in sizes)
    buf = torch.zeros((pad,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    buf[:local.shape[0]] = local
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf)
    return torch.cat([p[:b - a] for p, (a, b) in zip(parts, sizes)], dim=0)
