#!/usr/bin/env python
"""End-to-end timing of the host-pixel path (render + D2H into pinned host memory) of the C ABI, for one frame
per call (the reference's `rasterize`) and for a batch, and a bit-for-bit check of the host result against the
device-resident render.  Slice size: env RXC_SLICE_MB (0 = whole frames).
usage: e2e_bench.py [workload:frames ...]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from rusterix_b200 import DeviceContext, Rasterizer

specs = sys.argv[1:] or ["map4k:1", "map4k:8", "sweep1080:1", "sweep1080:8", "teapot1080:1", "dense8k:1"]
ctx = DeviceContext.get(0)
for spec in specs:
    name, frames = spec.split(":"); frames = int(frames)
    cfg, frame_ids, desc = bench.build_workload(name, frames, 0, 1)
    rasts = [cfg.rasterizer(i) for i in frame_ids]
    dev = torch.empty((frames, cfg.height, cfg.width, 4), dtype=torch.uint8, device="cuda:0")
    host = torch.zeros((frames, cfg.height, cfg.width, 4), dtype=torch.uint8, pin_memory=True)
    batch = Rasterizer.prepare_batch(rasts, cfg.scene, cfg.width, cfg.height, cfg.tile_size, cfg.assets)
    batch.run(dev, sync=True)
    for _ in range(3):
        batch.run(host, sync=True)
    same = bool(torch.equal(dev.cpu(), host))
    n = 20
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        batch.run(host, sync=True)
    ms = (time.perf_counter() - t0) / n * 1e3
    t0 = time.perf_counter()
    for _ in range(n):
        batch.run(dev, sync=True)
    ms_dev = (time.perf_counter() - t0) / n * 1e3
    mb = host.numel() / 1e6
    # the same call with a pageable buffer, and with that buffer page-locked through the ABI (rxc_pin_host)
    import numpy as np
    page = np.zeros(host.shape, dtype=np.uint8)
    def wall(buf, n=10):
        batch.run(buf, sync=True)
        t = time.perf_counter()
        for _ in range(n):
            batch.run(buf, sync=True)
        return (time.perf_counter() - t) / n * 1e3
    ms_page = wall(page)
    ctx.pin_host(page)
    ms_reg = wall(page)
    ctx.unpin_host(page)
    print(f"{name:11s} x{frames}: e2e {ms:7.3f} ms ({mb / ms:6.1f} GB/s to host, {mb / 4 / ms * 1e-3 * 1e3:8.0f} Mpix/s)  device-resident {ms_dev:7.3f} ms  identical={same}  pageable {ms_page:7.3f} ms  rxc_pin_host {ms_reg:7.3f} ms")
