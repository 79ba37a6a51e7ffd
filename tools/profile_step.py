#!/usr/bin/env python
"""Renders a few steps of one bench workload (device-resident) -- the short command ncu wraps.
usage: profile_step.py <workload> [steps] [frames]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from rusterix_b200 import DeviceContext, Rasterizer  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "map4k"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
frames = int(sys.argv[3]) if len(sys.argv) > 3 else (1 if name == "dense8k" else 8)
cfg, frame_ids, desc = bench.build_workload(name, frames, 0, 1)
rasts = [cfg.rasterizer(i) for i in frame_ids]
out = torch.empty((frames, cfg.height, cfg.width, 4), dtype=torch.uint8, device="cuda:0")
batch = Rasterizer.prepare_batch(rasts, cfg.scene, cfg.width, cfg.height, cfg.tile_size, cfg.assets)
for _ in range(steps):
    batch.run(out, sync=True)
s = DeviceContext.get(0).stats()
print(desc, "launches", s.kernel_launches, "binned", s.last_binned_refs, "large", s.last_large_tris, "visible", s.last_visible_tris)
