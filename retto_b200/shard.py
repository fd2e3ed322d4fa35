"""Multi-GPU = page sharding, no collective on the data path: pages are independent
(session.rs:75-106 reads no cross-page state), so rank r of N owns a slice of the page list and results
are gathered on the host by page index, which reproduces the sequential CLI order (main.rs:80-86).
Mixed page sizes are balanced with LPT (largest-processing-time first) on H*W."""
from __future__ import annotations

from typing import List, Sequence


def shard_indices(sizes: Sequence[int], world: int) -> List[List[int]]:
    """LPT assignment of page indices to `world` ranks by descending cost; each rank's list is sorted so a
    rank processes its pages in the original order.  Deterministic (ties by index)."""
    order = sorted(range(len(sizes)), key=lambda i: (-int(sizes[i]), i))
    loads = [0] * world
    out = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (loads[k], k))
        out[r].append(i)
        loads[r] += int(sizes[i])
    return [sorted(v) for v in out]


def gather_by_page(local_results, local_indices, n_pages: int, group=None):
    """all ranks end up with the full result list ordered by page index (torch.distributed object gather;
    host-side plumbing only — results are a few hundred bytes per page)."""
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        full = [None] * n_pages
        for i, r in zip(local_indices, local_results):
            full[i] = r
        return full
    payload = list(zip(local_indices, local_results))
    gathered = [None] * dist.get_world_size(group)
    dist.all_gather_object(gathered, payload, group=group)
    full = [None] * n_pages
    for part in gathered:
        for i, r in part:
            full[i] = r
    return full


def bind_host_to_gpu(device_index: int, pci_bus_id: str | None = None) -> int:
    """Pin the calling process to the CPU cores NVML reports as local to the GPU (same NUMA node / PCIe root), so
    that the pinned page buffers it allocates afterwards are placed on that node (first-touch) and each rank's
    host->device stream runs over its own memory controller.  Matters only for the host-resident (e2e) path with
    several ranks on one box: unbound, the 8 x 50 GB/s of page uploads share whichever node the ranks happened to
    start on.  Returns the number of cores bound to (0 = NVML / affinity unavailable, nothing changed)."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByPciBusId(pci_bus_id.encode() if isinstance(pci_bus_id, str) else pci_bus_id) \
            if pci_bus_id else pynvml.nvmlDeviceGetHandleByIndex(device_index)
        n_words = ((os.cpu_count() or 64) + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        cpus = {64 * i + b for i, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if not cpus:
            return 0
        os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:
        return 0
