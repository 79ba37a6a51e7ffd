#!/usr/bin/env python
"""Whole-step device time (CUDA events around the step, no per-kernel events in between, so programmatic dependent
launch is not broken up) of the bench workloads and of the R column bands of the dense 8K frame.
Run once with RXC_PDL=1 and once with RXC_PDL=0 (or with RXC_RASTER_GROUPS=1 against unset: the frame groups of k_raster).  usage: pdl_ab.py [bands R] [workload ...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from rusterix_b200 import DeviceContext, Rasterizer, mgpu

args = sys.argv[1:]
R = 0
if args and args[0] == "bands":
    R = int(args[1]); args = args[2:]
names = args or ["map4k", "teapot1080", "dense8k", "sweep1080", "chunked1080", "game2d1080", "shaded1080"]
ctx = DeviceContext.get(0)
ctx.set_vm_jit(int(os.environ.get("RXC_VM_JIT", "2")))   # like bench.py: the kernels recompiled for the scene, before the first frame
stream = torch.cuda.Stream(device="cuda:0")               # kernels, L2 flush and events on ONE stream
torch.cuda.set_stream(stream)
ctx.set_stream(stream.cuda_stream)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda:0")
tag = "pdl=" + os.environ.get("RXC_PDL", "1") + " groups=" + os.environ.get("RXC_RASTER_GROUPS", "auto")


def timed(run, n=30):
    for _ in range(3):
        run(); flush.zero_()
    torch.cuda.synchronize()
    ms = []
    for _ in range(n):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); run(); b.record(); flush.zero_()
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    ms.sort()
    return sum(ms) / len(ms), ms[len(ms) // 2], ms[0]


for name in names:
    frames = 1 if name == "dense8k" else 32 if name == "sweep1080" else 8
    cfg, frame_ids, desc = bench.build_workload(name, frames, 0, 1)
    rasts = [cfg.rasterizer(i) for i in frame_ids]
    out = torch.empty((frames, cfg.height, cfg.width, 4), dtype=torch.uint8, device="cuda:0")
    batch = Rasterizer.prepare_batch(rasts, cfg.scene, cfg.width, cfg.height, cfg.tile_size, cfg.assets)
    mean, med, best = timed(lambda: batch.run(out, sync=False))
    print(f"{tag} {name:12s} step mean {mean:.4f} median {med:.4f} min {best:.4f} ms", flush=True)

if R:
    cfg, frame_ids, desc = bench.build_workload("dense8k", 1, 0, 1)
    W, H = cfg.width, cfg.height
    worst = 0.0
    for r in range(R):
        x0, x1 = mgpu.column_band_for_rank(W, r, R)
        out = torch.empty((1, H, max(1, x1 - x0), 4), dtype=torch.uint8, device="cuda:0")
        batch = Rasterizer.prepare_batch([cfg.rasterizer(frame_ids[0])], cfg.scene, W, H, cfg.tile_size, cfg.assets, band=(0, H, x0, x1))
        mean, med, best = timed(lambda: batch.run(out, sync=False))
        worst = max(worst, med)
        print(f"{tag} dense8k column band {r}/{R} [{x0},{x1}) step mean {mean:.4f} median {med:.4f} min {best:.4f} ms", flush=True)
    print(f"{tag} dense8k {R} column bands: slowest band median {worst:.4f} ms", flush=True)
    # what cost-balanced column bands would give (mgpu.BandBalancer over the tile columns, one band after the other on this GPU)
    bal = mgpu.BandBalancer(W, R, 32)
    for it in range(int(os.environ.get("AB_BALANCE", "0"))):
        times = []
        for r in range(R):
            x0, x1 = bal.band(r)
            if x1 <= x0:
                times.append(0.0); continue
            out = torch.empty((1, H, x1 - x0, 4), dtype=torch.uint8, device="cuda:0")
            batch = Rasterizer.prepare_batch([cfg.rasterizer(frame_ids[0])], cfg.scene, W, H, cfg.tile_size, cfg.assets, band=(0, H, x0, x1))
            times.append(timed(lambda: batch.run(out, sync=False), n=10)[1])
        print(f"{tag} balanced columns, iteration {it}: edges {bal.edges} slowest {max(times):.4f} ms  " + " ".join(f"{t:.3f}" for t in times), flush=True)
        bal.update(times)
