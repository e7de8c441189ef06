"""The N > 1 path on CPU: two gloo ranks shard a batch of pairs, each runs its shard through the same function, and the
gathered predictions equal the single-process result (bench.py uses exactly these helpers with NCCL)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from unimatch_b200.sharding import gather_predictions, shard_range


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_forward(pairs):
    # any deterministic per-pair function: pairs never interact
    return torch.stack([p.flip(-1).cumsum(-1) * (1.0 + p.mean()) for p in pairs])


def _worker(rank, world, port, batch, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(7)
    data = torch.randn((batch, 2, 6, 9), generator=g)
    a, b = shard_range(batch, rank, world)
    local = _fake_forward(data[a:b])
    out = gather_predictions(local, batch)
    assert torch.equal(out, gather_predictions(local))      # shard sizes exchanged instead of given
    if batch % world == 0:                                  # the asynchronous form bench.py overlaps with the next forward
        out2, work = gather_predictions(local, batch, async_op=True)
        work.wait()
        assert torch.equal(out2, out)
    if rank == 0:
        q.put(out)
    try:                                                    # a shard that is not this rank's must be rejected, not mis-assembled
        gather_predictions(torch.zeros(99, 2), 6)
        ok = False
    except ValueError:
        ok = True
    dist.barrier()
    dist.destroy_process_group()
    assert ok


@pytest.mark.parametrize("batch", [8, 5])
def test_two_rank_sharding_matches_single_process(batch):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, batch, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    g = torch.Generator().manual_seed(7)
    data = torch.randn((batch, 2, 6, 9), generator=g)
    assert torch.equal(out, _fake_forward(data))


def test_shard_ranges_partition_the_batch():
    for batch in (1, 5, 8, 64):
        for world in (1, 2, 3, 8):
            r = [shard_range(batch, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == batch
            assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
