"""Multi-GPU sharding of the rasterize path: one process per GPU, frames of a
camera sweep or screen bands of one frame are independent units, so ranks render without any
data-path collective; torch.distributed (NCCL over NVLink on GPUs, gloo in the CPU tests) is used
only to gather finished frames / bands to rank 0."""
import os
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist

TILE_H = 32  # GPU tile height (csrc/rx_device.cuh RX_TILE_H): bands start on tile rows


def bind_to_gpu_numa_node(device_index: int) -> Optional[List[int]]:
    """Pins the calling process to the host cores next to GPU `device_index` (NVML's ideal CPU affinity,
    i.e. the NUMA node its PCIe root port hangs off).  Host pixel buffers pinned afterwards are then
    first-touched on that node, so the D2H of finished frames does not cross the socket interconnect
    when 8 ranks drain their frames at once.  Returns the CPU list, or None when NVML or the topology
    is unavailable (nothing is changed then)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = device_index
        if visible:
            ids = [v.strip() for v in visible.split(",") if v.strip()]
            if device_index < len(ids) and ids[device_index].isdigit():
                idx = int(ids[device_index])
        h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cpus = [w * 64 + b for w, m in enumerate(words) for b in range(64) if (int(m) >> b) & 1]
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:
        return None


def shard_frames(n_frames: int, rank: int, world: int) -> List[int]:
    """Round-robin frame sharding: rank r renders frames r, r+world, ..."""
    return list(range(rank, n_frames, world))


def band_for_rank(height: int, rank: int, world: int, align: int = TILE_H) -> Tuple[int, int]:
    """Rows [y0, y1) of rank `rank` when a frame of `height` rows is split into `world` bands whose
    boundaries are multiples of `align` (the last band takes the remainder; bands may be empty)."""
    rows_aligned = (height + align - 1) // align
    per = (rows_aligned + world - 1) // world
    y0 = min(height, rank * per * align)
    y1 = min(height, (rank + 1) * per * align)
    return y0, y1


def all_bands(height: int, world: int, align: int = TILE_H) -> List[Tuple[int, int]]:
    return [band_for_rank(height, r, world, align) for r in range(world)]


def gather_bands_to_rank0(band: torch.Tensor, height: int, width: int, rank: int, world: int, align: int = TILE_H):
    """Each rank holds its band [rows, width, 4] uint8; rank 0 returns the full [height, width, 4]
    frame, other ranks return None.  Point-to-point sends (bands differ in size)."""
    if world == 1:
        return band
    bands = all_bands(height, world, align)
    if rank == 0:
        full = torch.empty((height, width, 4), dtype=torch.uint8, device=band.device)
        y0, y1 = bands[0]
        full[y0:y1].copy_(band)
        reqs = []
        for r in range(1, world):
            y0, y1 = bands[r]
            if y1 > y0:
                reqs.append(dist.irecv(full[y0:y1], src=r))
        for q in reqs:
            q.wait()
        return full
    if band.numel():
        dist.send(band.contiguous(), dst=0)
    return None


def gather_frames_to_rank0(frames: torch.Tensor, rank: int, world: int):
    """Each rank holds [F, H, W, 4]; rank 0 returns [world*F, H, W, 4] in round-robin frame order
    (frame i of rank r is global frame i*world + r), other ranks return None."""
    if world == 1:
        return frames
    if rank == 0:
        parts = [torch.empty_like(frames) for _ in range(world)]
        dist.gather(frames, parts, dst=0)
        stacked = torch.stack(parts, dim=1)  # [F, world, H, W, 4]
        return stacked.reshape((-1,) + tuple(frames.shape[1:]))
    dist.gather(frames, None, dst=0)
    return None
