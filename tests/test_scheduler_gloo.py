"""Batch scheduler on CPU: world_size-2 (and 3, ragged) gloo process groups.

The CUDA kernels cannot run here, so ``propagate_fn`` is a stand-in that tags every row with the rank
that processed it; what is under test is the host logic the N>1 path relies on: the contiguous row
split, the ragged all-gather of complex results and of the per-row step counts."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from opticomlib_b200.scheduler import row_shard, all_shards


def test_row_shard_is_a_balanced_partition():
    for total in (0, 1, 7, 8, 4096, 4099):
        for world in (1, 2, 3, 8):
            shards = all_shards(total, world)
            assert shards[0].start == 0 and shards[-1].stop == total
            assert all(a.stop == b.start for a, b in zip(shards, shards[1:]))
            counts = [s.count for s in shards]
            assert max(counts) - min(counts) <= 1 and sum(counts) == total
    with pytest.raises(ValueError):
        row_shard(10, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from opticomlib_b200.scheduler import propagate_sharded
        base = torch.arange(total * n, dtype=torch.float64).reshape(total, n)
        full_in = torch.complex(base, -base)

        def rows_fn(shard):
            return full_in[shard.start:shard.stop].clone()

        def fake_propagate(block):                       # stand-in for devices.fiber_batch on this rank's GPU
            steps = np.arange(block.shape[0], dtype=np.int32) + 100 * (rank + 1)
            return block * (2.0 + 1j) + rank, steps

        out, steps = propagate_sharded(rows_fn, total, fake_propagate, gather=True)
        shards = all_shards(total, world)
        ok = out.shape == (total, n) and out.dtype == torch.complex128 and steps.shape == (total,)
        for s in shards:
            ok &= bool(torch.equal(out[s.start:s.stop], full_in[s.start:s.stop] * (2.0 + 1j) + s.rank))
            ok &= bool(torch.equal(steps[s.start:s.stop], torch.arange(s.count, dtype=torch.int32) + 100 * (s.rank + 1)))
        local, lsteps = propagate_sharded(rows_fn, total, fake_propagate, gather=False)
        ok &= local.shape[0] == shards[rank].count
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,total", [(2, 8), (2, 7), (3, 10)])
def test_sharded_propagation_and_ragged_gather(world, total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, 16, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    results = dict(q.get(timeout=5) for _ in range(world))
    assert results == {r: True for r in range(world)}
