"""Batch sharding for the multi-GPU path.

The op has no cross-image term (the kernel index decomposes into an image index first,
/root/reference/codetr/csrc/ms_deform_attn.cu:226-232), so a batch shards by image with no
data-path collective: rank r of N owns a contiguous block of images.  The only communication is
the timing/accounting reduction done by the caller (``bench.py``) after the work.
"""
from __future__ import annotations

from typing import List, Tuple


def image_range(num_images: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Half-open range of images owned by ``rank``: contiguous blocks, sizes differ by at most one,
    lower ranks take the remainder."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError(f"bad rank/world_size {rank}/{world_size}")
    if num_images < 0:
        raise ValueError("num_images must be >= 0")
    base, extra = divmod(num_images, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def all_ranges(num_images: int, world_size: int) -> List[Tuple[int, int]]:
    return [image_range(num_images, r, world_size) for r in range(world_size)]


def shard_batch(tensors, rank: int, world_size: int):
    """Slice every batched tensor (dim 0 = image) to this rank's images; tensors without a batch
    dimension (``spatial_shapes``, ``level_start_index``) are replicated, i.e. passed as None -> None
    or returned unchanged when marked with ``replicate=True`` via a (tensor, True) tuple."""
    out = []
    for item in tensors:
        if isinstance(item, tuple):
            t, replicate = item
        else:
            t, replicate = item, False
        if t is None or replicate:
            out.append(t)
            continue
        lo, hi = image_range(t.shape[0], rank, world_size)
        out.append(t[lo:hi])
    return out


def weak_scaling_images(per_gpu_batch: int, world_size: int) -> int:
    """Weak scaling: per-GPU work is fixed, the job's image count grows with the GPU count."""
    return per_gpu_batch * world_size


def aggregate_throughput(elapsed_ms: float, images_this_rank: int, device=None):
    """Whole-job throughput from per-rank device timings: the job takes as long as its slowest rank
    (MAX over ranks of the CUDA-event time), and processes the SUM of the ranks' images.  Uses the
    default ``torch.distributed`` group when one is initialised (NCCL on GPUs, gloo in the CPU tests);
    this accounting reduction is the only collective of the path -- the data path itself has none.
    Returns ``(images_per_s, max_elapsed_ms, total_images)``."""
    import torch
    import torch.distributed as dist

    t = torch.tensor([float(elapsed_ms)], dtype=torch.float64, device=device)
    n = torch.tensor([int(images_this_rank)], dtype=torch.int64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(n, op=dist.ReduceOp.SUM)
    max_ms, total = float(t.item()), int(n.item())
    return (total / (max_ms * 1e-3) if max_ms > 0 else float("inf")), max_ms, total


def gpu_numa_cpus(device_index: int):
    """CPUs of the NUMA node the GPU hangs off (sysfs), or None when the platform does not say."""
    import os

    import torch

    try:
        props = torch.cuda.get_device_properties(device_index)
        bdf = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
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
        return cpus or None
    except Exception:
        return None


def bind_to_gpu_numa_node(device_index: int):
    """Pin this process to the CPUs next to its GPU, so that pinned host buffers allocated afterwards are
    first-touched on the GPU's NUMA node (host->device copies from the far node were measured at 17.8 GB/s
    against 48.8 GB/s from the near one on the same pool).  Returns the previous affinity set (pass it to
    ``os.sched_setaffinity(0, ...)`` to undo), or None when nothing was changed."""
    import os

    cpus = gpu_numa_cpus(device_index)
    if not cpus:
        return None
    previous = os.sched_getaffinity(0)
    try:
        os.sched_setaffinity(0, cpus)
    except OSError:
        return None
    return previous
