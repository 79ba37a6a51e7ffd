"""Multi-GPU sharding of the rasterize path: one process per GPU, frames of a
camera sweep or screen bands of one frame are independent units, so ranks render without any
data-path collective; torch.distributed (NCCL over NVLink on GPUs, gloo in the CPU tests) is used
only to gather finished frames / bands to rank 0."""
import ctypes as C
import os
from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from . import _abi

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


class BandBalancer:
    """Cost-balanced row bands for the band split of one large frame.

    Equal-height bands are badly unbalanced on real views (the horizon band of the 1M-triangle scene holds more than
    half of all triangle references, the sky bands almost none), and a band split runs at the speed of its slowest
    rank.  The balancer keeps the bands contiguous (one send per rank to gather them) and moves their boundaries
    from the times the ranks measured for the previous frame.  It keeps a cost profile over the GPU tile rows of
    the frame; every measured band time (minus the rank-independent front-end time) rescales the profile over the
    rows of that band, so that boundaries that move from frame to frame resolve where the cost really sits (a band
    straddling sky and terrain is not uniform).  The new boundaries cut the cumulative profile at k/world of its
    total.  Successive frames of a camera path are coherent, so a few frames converge; `update` needs every rank's
    time (`all_gather_times`), which is the only communication and is off the data path."""

    def __init__(self, height: int, world: int, align: int = TILE_H):
        self.height, self.world, self.align = int(height), int(world), int(align)
        self.n_rows = (self.height + self.align - 1) // self.align          # tile rows
        self.profile = [1.0] * self.n_rows                                   # relative cost per tile row
        self.edges = [band_for_rank(height, r, world, align)[0] for r in range(world)] + [int(height)]
        self.best = (float("inf"), list(self.edges))

    def bands(self) -> List[Tuple[int, int]]:
        return [(self.edges[r], self.edges[r + 1]) for r in range(self.world)]

    def band(self, rank: int) -> Tuple[int, int]:
        return self.edges[rank], self.edges[rank + 1]

    def use_best(self) -> List[Tuple[int, int]]:
        """Freeze the boundaries that gave the smallest slowest-rank time so far (static views, benchmarks)."""
        self.edges = list(self.best[1])
        return self.bands()

    def update(self, times_ms: List[float]) -> List[Tuple[int, int]]:
        """New boundaries from the per-rank times of the frame just rendered with the current ones."""
        if len(times_ms) != self.world:
            raise ValueError("one time per rank")
        worst = max(times_ms)
        if worst < self.best[0]:
            self.best = (worst, list(self.edges))
        fixed = 0.9 * min(times_ms)   # setup on the replicated geometry costs every rank about the same
        measured = [max(t - fixed, 1e-6) for t in times_ms]
        scale = sum(measured) / max(sum(self.profile), 1e-30)
        for r in range(self.world):
            a, b = self.edges[r] // self.align, (self.edges[r + 1] + self.align - 1) // self.align
            if b <= a:
                continue
            predicted = sum(self.profile[a:b]) * scale
            k = (measured[r] / max(predicted, 1e-30)) ** 0.6   # damped: the ranks' corrections interact through `scale`
            for i in range(a, b):
                self.profile[i] *= k
        total = sum(self.profile)
        new_edges, acc, i = [0], 0.0, 0
        for k in range(1, self.world):
            target = total * k / self.world
            while i < self.n_rows and acc + self.profile[i] <= target:
                acc += self.profile[i]
                i += 1
            # cut inside tile row i when it helps: boundaries stay on tile rows, so round to the nearer side
            cut = i + (1 if i < self.n_rows and (target - acc) > 0.5 * self.profile[i] else 0)
            y = min(cut * self.align, self.height)
            new_edges.append(max(y, new_edges[-1]))
        new_edges.append(self.height)
        self.edges = new_edges
        return self.bands()


def all_gather_times(ms: float, device=None) -> List[float]:
    """Every rank's time for the last frame (one small all_gather; NCCL on GPUs, gloo on CPU)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return [float(ms)]
    t = torch.tensor([ms], dtype=torch.float64, device=device)
    out = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return [float(x.item()) for x in out]


def gather_ragged_bands_to_rank0(band: torch.Tensor, bands: List[Tuple[int, int]], height: int, width: int, rank: int, world: int):
    """Like `gather_bands_to_rank0` for arbitrary contiguous bands (the balancer's)."""
    if world == 1:
        return band
    if rank == 0:
        full = torch.empty((height, width, 4), dtype=torch.uint8, device=band.device)
        y0, y1 = bands[0]
        full[y0:y1].copy_(band[: y1 - y0])
        reqs = []
        for r in range(1, world):
            y0, y1 = bands[r]
            if y1 > y0:
                reqs.append(dist.irecv(full[y0:y1], src=r))
        for q in reqs:
            q.wait()
        return full
    y0, y1 = bands[rank]
    if y1 > y0:
        dist.send(band[: y1 - y0].contiguous(), dst=0)
    return None


def column_band_for_rank(width: int, rank: int, world: int, align: int = 32) -> Tuple[int, int]:
    """Columns [x0, x1) of rank `rank` when a frame of `width` columns is split into `world` column bands whose left
    edges are multiples of `align` (the GPU tile width).  Column bands cut every tile row -- the cheap sky rows and the
    expensive horizon rows alike -- into `world` pieces, so equal widths are balanced without any measurement."""
    cols_aligned = (width + align - 1) // align
    per = (cols_aligned + world - 1) // world
    return min(width, rank * per * align), min(width, (rank + 1) * per * align)


def gather_column_bands_to_rank0(band: torch.Tensor, height: int, width: int, rank: int, world: int, align: int = 32):
    """Each rank holds its column band [height, x1 - x0, 4] uint8; rank 0 returns the full [height, width, 4] frame."""
    if world == 1:
        return band
    if rank == 0:
        full = torch.empty((height, width, 4), dtype=torch.uint8, device=band.device)
        x0, x1 = column_band_for_rank(width, 0, world, align)
        full[:, x0:x1].copy_(band)
        parts = []
        for r in range(1, world):
            x0, x1 = column_band_for_rank(width, r, world, align)
            if x1 > x0:
                buf = torch.empty((height, x1 - x0, 4), dtype=torch.uint8, device=band.device)
                parts.append((x0, x1, buf, dist.irecv(buf, src=r)))
        for x0, x1, buf, req in parts:
            req.wait()
            full[:, x0:x1].copy_(buf)
        return full
    if band.numel():
        dist.send(band.contiguous(), dst=0)
    return None


# ---------------------------------------------------------------------------------------------------------------------
# rxc_mgpu_*: delivery to rank 0 inside the library (include/rxcuda.h).  The raster kernel of every rank writes its
# tiles straight into rank 0's buffer over NVLink (cudaIpc peer mapping); torch.distributed only carries the 128-byte
# ncclUniqueId from rank 0 to the others once.
# ---------------------------------------------------------------------------------------------------------------------
class _DevicePointer:
    """Lets torch wrap a raw device address (rank 0's delivery buffer) without copying."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 3, "strides": None}


def frame_regions(world: int, frames_per_rank: int, frame_bytes: int, base_offset: int = 0):
    """The regions of a frame-sharded step: rank r's `frames_per_rank` frames sit back to back at
    base_offset + r * frames_per_rank * frame_bytes (contiguous blocks per rank)."""
    arr = (_abi.rxc_mgpu_region * world)()
    for r in range(world):
        arr[r].rank, arr[r].rows = r, 1
        arr[r].offset = base_offset + r * frames_per_rank * frame_bytes
        arr[r].row_bytes = frames_per_rank * frame_bytes
        arr[r].pitch_bytes = 0
    return arr


def band_regions(bands: Sequence[Tuple[int, int, int, int]], width: int, base_offset: int = 0):
    """The regions of a band split: bands[r] = (y0, y1, x0, x1) of rank r inside a frame `width` pixels wide."""
    arr = (_abi.rxc_mgpu_region * len(bands))()
    for r, (y0, y1, x0, x1) in enumerate(bands):
        arr[r].rank = r
        arr[r].rows = max(0, y1 - y0) if x1 > x0 else 0
        arr[r].offset = base_offset + (y0 * width + x0) * 4
        arr[r].row_bytes = max(0, x1 - x0) * 4
        arr[r].pitch_bytes = width * 4
        if x0 == 0 and x1 == width:   # full rows: one contiguous block
            arr[r].row_bytes, arr[r].rows, arr[r].pitch_bytes = max(0, y1 - y0) * width * 4, 1 if y1 > y0 else 0, 0
    return arr


class Delivery:
    """One rxc_mgpu session of a DeviceContext: `Delivery(ctx, rank, world)` on every rank (collective), then per
    step `render(...)` any number of times, `deliver(regions)` once, and `release()` when rank 0 is done reading."""

    def __init__(self, ctx, rank: int, world: int, id_bytes: Optional[bytes] = None):
        self.ctx, self.rank, self.world = ctx, int(rank), int(world)
        self.lib = ctx.lib
        ident = (C.c_uint8 * _abi.RXC_MGPU_ID_BYTES)()
        if world > 1:
            if id_bytes is None:
                box = [None]
                if rank == 0:
                    ctx.check_static(self.lib.rxc_mgpu_unique_id(ident))
                    box[0] = bytes(ident)
                dist.broadcast_object_list(box, src=0)
                id_bytes = box[0]
            C.memmove(ident, id_bytes, _abi.RXC_MGPU_ID_BYTES)
        ctx.check(self.lib.rxc_mgpu_init(ctx.handle, ident, self.rank, self.world))
        self.bytes, self.ptr, self.mode = 0, None, _abi.RXC_MGPU_LOCAL

    def target(self, nbytes: int):
        """Collective: (re)allocates the delivery buffer on rank 0 and maps it everywhere.  Returns a uint8 tensor view
        of it on rank 0, None on the other ranks."""
        p, mode = C.c_void_p(), C.c_uint32()
        self.ctx.check(self.lib.rxc_mgpu_target(self.ctx.handle, int(nbytes), C.byref(p), C.byref(mode)))
        self.bytes, self.ptr, self.mode = int(nbytes), p.value, int(mode.value)
        if self.rank == 0:
            return torch.as_tensor(_DevicePointer(p.value, nbytes), device=torch.device("cuda", self.ctx.device))
        return None

    def render(self, batch, offset_bytes: int, frame_stride_bytes: Optional[int] = None, pitch_bytes: int = 0):
        """rxc_mgpu_rasterize of a prepared FrameBatch (asynchronous)."""
        stride = batch.stride if frame_stride_bytes is None else int(frame_stride_bytes)
        self.ctx.check(self.lib.rxc_mgpu_rasterize(self.ctx.handle, batch.frames, batch.n, int(offset_bytes), stride, int(pitch_bytes)))

    def deliver(self, regions=None):
        n = len(regions) if regions is not None else 0
        self.ctx.check(self.lib.rxc_mgpu_deliver(self.ctx.handle, regions, n))

    def release(self):
        self.ctx.check(self.lib.rxc_mgpu_release(self.ctx.handle))

    def status(self):
        mode, deliveries, timeouts = C.c_uint32(), C.c_uint32(), C.c_uint32()
        self.ctx.check(self.lib.rxc_mgpu_status(self.ctx.handle, C.byref(mode), C.byref(deliveries), C.byref(timeouts)))
        return {"mode": ("local", "peer", "nccl")[mode.value], "deliveries": int(deliveries.value), "timeouts": int(timeouts.value)}

    def close(self):
        self.ctx.check(self.lib.rxc_mgpu_shutdown(self.ctx.handle))
