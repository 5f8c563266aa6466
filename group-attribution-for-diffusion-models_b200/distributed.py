"""Sharding helpers for the multi-GPU path (one process per GPU, torch.distributed / NCCL).

The hot path shards by training example: every rank featurises and projects its own
examples with the same seed (hence the same P), holds Phi_r [N_r, k], and the only exchanges are
  * one all-reduce (sum) of the k x k Gram matrix  -- ``allreduce_sum_``
  * one all-gather of per-example score slices     -- ``allgather_cat`` (handles unequal shard sizes)
These wrappers work on any backend (NCCL on the GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def is_distributed() -> bool:
    return dist.is_available() and dist.is_initialized()


def world_and_rank(group=None):
    if not is_distributed():
        return 1, 0
    return dist.get_world_size(group), dist.get_rank(group)


def shard_range(n_total: int, world: int, rank: int):
    """Contiguous, balanced example range [lo, hi) of `rank`: the first n_total % world ranks get one extra."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} out of range for world size {world}")
    base, rem = divmod(int(n_total), int(world))
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def allreduce_sum_(t: torch.Tensor, group=None) -> torch.Tensor:
    if is_distributed() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def allgather_cat(t: torch.Tensor, n_total: int | None = None, dim: int = -1, group=None) -> torch.Tensor:
    """Concatenate every rank's slice along `dim`; slices may differ in length (shard_range layout)."""
    if not is_distributed() or dist.get_world_size(group) == 1:
        return t
    world = dist.get_world_size(group)
    dim = dim % t.dim()
    if n_total is None:
        sizes = [torch.zeros(1, dtype=torch.int64, device=t.device) for _ in range(world)]
        dist.all_gather(sizes, torch.tensor([t.shape[dim]], dtype=torch.int64, device=t.device), group=group)
        lens = [int(s.item()) for s in sizes]
    else:
        lens = [shard_range(n_total, world, r)[1] - shard_range(n_total, world, r)[0] for r in range(world)]
    mx = max(lens)
    moved = t.movedim(dim, 0).contiguous()
    if moved.shape[0] < mx:
        pad = torch.zeros((mx - moved.shape[0],) + tuple(moved.shape[1:]), dtype=t.dtype, device=t.device)
        moved = torch.cat([moved, pad], dim=0)
    parts = [torch.empty_like(moved) for _ in range(world)]
    dist.all_gather(parts, moved, group=group)
    out = torch.cat([p[:n] for p, n in zip(parts, lens)], dim=0)
    return out.movedim(0, dim)


def bind_to_gpu_numa(device_index: int) -> list[int] | None:
    """Pin this process to the CPUs NVML reports as local to the GPU (same NUMA node / PCIe root), so that pinned
    host buffers allocated afterwards are first-touched next to the GPU.  Matters when several ranks stream
    gradients host->device at once: without it all ranks' buffers tend to land on one socket.  Returns the CPU list,
    or None when NVML / sched_setaffinity are unavailable (nothing is changed then)."""
    import os

    try:
        import pynvml

        pynvml.nvmlInit()
        try:
            bus = torch.cuda.get_device_properties(device_index).pci_bus_id
            dom = getattr(torch.cuda.get_device_properties(device_index), "pci_domain_id", 0)
            dev = getattr(torch.cuda.get_device_properties(device_index), "pci_device_id", 0)
            handle = pynvml.nvmlDeviceGetHandleByPciBusId(f"{dom:08x}:{bus:02x}:{dev:02x}.0".encode())
        except Exception:
            handle = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        n_cpus = max(os.cpu_count() or 1, max(os.sched_getaffinity(0)) + 1)
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (n_cpus + 63) // 64)
        cpus = [64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1]
        allowed = os.sched_getaffinity(0)  # the container's cpuset
        cpus = [c for c in cpus if c in allowed]
        if not cpus or len(cpus) == len(allowed):
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:
        return None
