#!/usr/bin/env python
"""Renders the same frames repeatedly and checks that every result is bitwise identical to the first
(the kernels use atomics for list placement only; the image must not depend on scheduling).
usage: determinism.py [workload] [iterations]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from rusterix_b200 import Rasterizer

name = sys.argv[1] if len(sys.argv) > 1 else "teapot1080"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 100
frames = 1 if name == "dense8k" else 4
cfg, frame_ids, desc = bench.build_workload(name, frames, 0, 1)
rasts = [cfg.rasterizer(i) for i in frame_ids]
out = torch.empty((frames, cfg.height, cfg.width, 4), dtype=torch.uint8, device="cuda:0")
batch = Rasterizer.prepare_batch(rasts, cfg.scene, cfg.width, cfg.height, cfg.tile_size, cfg.assets)
batch.run(out, sync=True)
ref = out.clone()
bad = 0
for i in range(iters):
    out.zero_()
    batch.run(out, sync=True)
    d = int((out != ref).any(dim=-1).sum().item())
    if d:
        bad += 1
        print("iteration", i, "differs in", d, "pixels")
print(name, "iterations", iters, "nondeterministic results", bad)
sys.exit(1 if bad else 0)
