"""Batch scheduler: shard independent waveforms over the GPUs of one box.

Rows of a batch (Monte-Carlo noise realisations, launch-power / length sweeps, received frames for
DBP) are independent -- the reference would simply loop over them, and the step-size maximum of
devices.py:1194 is per waveform -- so the data path needs NO collective: every rank propagates a
contiguous block of rows with its own plan.  ``torch.distributed`` (NCCL over NVLink on GPUs, gloo
in the CPU tests) is used only to gather the results and the per-row step counts.

One process per GPU, launched with ``torchrun``; ``LOCAL_RANK`` selects the device.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass(frozen=True)
class RowShard:
    rank: int
    world: int
    total: int
    start: int
    stop: int

    @property
    def count(self) -> int:
        return self.stop - self.start


def row_shard(total_rows: int, world: int, rank: int) -> RowShard:
    """Contiguous, balanced split: the first ``total % world`` ranks get one extra row."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad world/rank %r/%r" % (world, rank))
    if total_rows < 0:
        raise ValueError("total_rows < 0")
    base, extra = divmod(total_rows, world)
    start = rank * base + min(rank, extra)
    return RowShard(rank, world, total_rows, start, start + base + (1 if rank < extra else 0))


def all_shards(total_rows: int, world: int):
    return [row_shard(total_rows, world, r) for r in range(world)]


def gather_rows(local, total_rows: int, group=None):
    """All-gather row blocks ``local[count_r, ...]`` of every rank into ``[total_rows, ...]``.

    Ranks may hold different row counts (ragged split): blocks are padded to the largest count for the
    collective and trimmed afterwards.  Complex tensors travel as real views (NCCL has no complex type).
    """
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    shards = all_shards(total_rows, world)
    if local.shape[0] != shards[rank].count:
        raise ValueError("rank %d holds %d rows, its shard has %d" % (rank, local.shape[0], shards[rank].count))
    cplx = local.is_complex()
    payload = torch.view_as_real(local.contiguous()) if cplx else local.contiguous()
    width = max(s.count for s in shards)
    if payload.shape[0] < width:
        pad = torch.zeros((width - payload.shape[0],) + tuple(payload.shape[1:]), dtype=payload.dtype, device=payload.device)
        payload = torch.cat([payload, pad])
    out = torch.empty((world * width,) + tuple(payload.shape[1:]), dtype=payload.dtype, device=payload.device)
    dist.all_gather_into_tensor(out, payload, group=group)
    parts = [out[r * width: r * width + shards[r].count] for r in range(world)]
    full = torch.cat(parts)
    return torch.view_as_complex(full) if cplx else full


def propagate_sharded(rows_fn, total_rows: int, propagate_fn, gather: bool = True, group=None):
    """Run ``propagate_fn`` on this rank's block of rows and optionally gather everything.

    rows_fn(shard)        -> this rank's input block  [shard.count, (P,) N]  (built or loaded locally)
    propagate_fn(block)   -> (output block, steps[int32 shard.count])        (e.g. ``devices.fiber_batch``)
    Returns (out, steps): the whole batch on every rank when ``gather`` else the local block.
    """
    import torch
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized():
        world, rank = dist.get_world_size(group), dist.get_rank(group)
    else:
        world, rank = 1, 0
    shard = row_shard(total_rows, world, rank)
    out, steps = propagate_fn(rows_fn(shard))
    steps = torch.as_tensor(np.asarray(steps, dtype=np.int32), device=out.device if torch.is_tensor(out) else "cpu")
    if not torch.is_tensor(out):
        out = torch.from_numpy(np.asarray(out))
    if gather and world > 1:
        return gather_rows(out, total_rows, group), gather_rows(steps, total_rows, group)
    return out, steps


def bind_to_gpu_numa_node(device_index: int):
    """Pin this process to the CPUs of the NUMA node its GPU hangs off, so that host buffers it allocates afterwards (first
    touch, and the pinned allocations of the host pipelines) live in that node's memory.  ``torchrun`` does not do this; with
    eight ranks whose pinned buffers all land on one socket the H2D + D2H streams of the end-to-end path share that socket's
    memory controllers and the inter-socket link (measured on the 8-GPU box: 64.8 ms per propagation end to end against 47.9 ms
    device-resident; the run with the noise generated on the device, i.e. D2H only, sits at 51.0 ms).
    Returns ``(node, cpus)`` or None when the topology cannot be read (no sysfs, no permission): then nothing is changed."""
    import os
    try:
        import torch
        pr = torch.cuda.get_device_properties(device_index)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        with open("/sys/bus/pci/devices/%s/numa_node" % bdf) as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        allowed = cpus & os.sched_getaffinity(0)
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node, sorted(allowed)
    except Exception:
        return None

