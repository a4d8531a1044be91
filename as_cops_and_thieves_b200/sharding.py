"""World sharding across ranks.

Worlds are independent and the map is read-only, so the global world range is cut into
contiguous shards, one per rank / GPU, with no data-path collective.  Spawn randomness is keyed by
*global* world id (``include/cat_philox.h``), so the trajectories do not depend on the GPU count.
The only collectives in the whole system are (a) the MAPPO gradient all-reduce and (b) the
optional 3-double all-reduce of advantage statistics (``gae.compute_gae``).
"""
from __future__ import annotations

import os
from typing import Iterable, Tuple

import torch


def shard_range(n_global: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous [gid0, gid0 + n_local) of ``rank``; remainders go to the lowest ranks."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    base, rem = divmod(int(n_global), int(world_size))
    n_local = base + (1 if rank < rem else 0)
    gid0 = rank * base + min(rank, rem)
    return gid0, n_local


def dist_env() -> Tuple[int, int, int]:
    """(rank, local_rank, world_size) from the torchrun environment (1 process per GPU)."""
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def bind_to_gpu_numa_node(device_index: int) -> Tuple[int, ...]:
    """Pin this process to the CPU cores local to GPU ``device_index`` (best effort; returns the cores, () if
    unknown).  Call it BEFORE allocating pinned host buffers: with one process per GPU on a two-socket box, the
    first-touch policy then places the pinned pages on the GPU's own NUMA node, so the kernel's zero-copy
    stores (``CatWorlds.step_host``) and the DMA copies do not cross the socket interconnect."""
    try:
        props = torch.cuda.get_device_properties(device_index)
        bdf = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/local_cpulist") as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            if "-" in part:
                lo, hi = part.split("-")
                cpus.update(range(int(lo), int(hi) + 1))
            elif part:
                cpus.add(int(part))
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if cpus and cpus != allowed:
            os.sched_setaffinity(0, cpus)
        return tuple(sorted(cpus))
    except Exception:
        return ()


def allreduce_gradients(params: Iterable[torch.nn.Parameter], world_size: int, group=None) -> torch.Tensor:
    """MAPPO minibatch gradient all-reduce (skrl does the same when launched distributed;
    SURVEY.md §2.1, Appendix D): one flat fp32 bucket, one ``all_reduce``, mean over ranks.
    With NVSwitch the cost is latency-bound, so a single bucket beats per-tensor calls."""
    import torch.distributed as dist
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return torch.zeros(0)
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, group=group)
    flat.div_(world_size)
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n
    return flat
