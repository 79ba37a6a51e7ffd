#!/usr/bin/env python
"""Device-resident timing of the bench workloads (k_raster and whole step), for kernel experiments.
usage: quick_bench.py [workload ...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from rusterix_b200 import DeviceContext, Rasterizer

names = sys.argv[1:] or ["map4k", "teapot1080", "dense8k", "sweep1080"]
ctx = DeviceContext.get(0)
for name in names:
    frames = 1 if name == "dense8k" else 8
    cfg, frame_ids, desc = bench.build_workload(name, frames, 0, 1)
    rasts = [cfg.rasterizer(i) for i in frame_ids]
    out = torch.empty((frames, cfg.height, cfg.width, 4), dtype=torch.uint8, device="cuda:0")
    batch = Rasterizer.prepare_batch(rasts, cfg.scene, cfg.width, cfg.height, cfg.tile_size, cfg.assets)
    for _ in range(3):
        batch.run(out, sync=True)
    ctx.reset_stats(); ctx.set_profiling(True)
    n = 30
    for _ in range(n):
        batch.run(out, sync=True)
    s = ctx.stats(); ctx.set_profiling(False)
    ms = [s.kernel_ms[i] / n for i in range(len(ctx.kernel_names()))]
    mpix = frames * cfg.width * cfg.height / 1e6
    print(f"{name:12s} raster {ms[7]:.4f} ms  step {sum(ms):.4f} ms  {mpix / sum(ms) * 1e3:9.0f} Mpix/s  others " + " ".join(f"{m:.3f}" for i, m in enumerate(ms) if i != 7))
