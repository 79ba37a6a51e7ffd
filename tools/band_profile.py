#!/usr/bin/env python
"""Per-band kernel times of the dense 8K frame (config D) on ONE GPU: what each rank of an R-way band split would
spend, front end vs raster.  usage: band_profile.py [R] [rows|cols]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from rusterix_b200 import DeviceContext, Rasterizer, mgpu

R = int(sys.argv[1]) if len(sys.argv) > 1 else 8
KIND = sys.argv[2] if len(sys.argv) > 2 else "rows"
ctx = DeviceContext.get(0)
cfg, frame_ids, desc = bench.build_workload("dense8k", 1, 0, 1)
W, H = cfg.width, cfg.height
names = ctx.kernel_names()
for r in range(R):
    if KIND == "rows":
        y0, y1 = mgpu.band_for_rank(H, r, R, 32); x0, x1 = 0, W
    else:
        y0, y1 = 0, H; x0, x1 = mgpu.column_band_for_rank(W, r, R)
    out = torch.empty((1, max(1, y1 - y0), max(1, x1 - x0), 4), dtype=torch.uint8, device="cuda:0")
    batch = Rasterizer.prepare_batch([cfg.rasterizer(frame_ids[0])], cfg.scene, W, H, cfg.tile_size, cfg.assets, band=(y0, y1, x0, x1))
    for _ in range(3):
        batch.run(out, sync=True)
    ctx.reset_stats(); ctx.set_profiling(True)
    n = 5
    for _ in range(n):
        batch.run(out, sync=True)
    s = ctx.stats(); ctx.set_profiling(False)
    ms = [s.kernel_ms[i] / n for i in range(len(names))]
    print(f"band {r} rows {y0}-{y1} cols {x0}-{x1}: raster {ms[7]:.3f}  front {sum(ms) - ms[7]:.3f}  total {sum(ms):.3f} ms   refs {s.last_binned_refs} visible {s.last_visible_tris}  " +
          " ".join(f"{names[i][2:]}={m:.3f}" for i, m in enumerate(ms) if m > 0.0005 and i != 7))
