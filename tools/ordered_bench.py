import os, sys, time
sys.path.insert(0, os.getcwd())
import torch
import bench
from rusterix_b200 import DeviceContext, Rasterizer
ctx = DeviceContext.get(0)
for mode, jit in ((0, 2), (1, 0), (1, 2)):
    ctx.set_vm_state_mode(mode)
    ctx.set_vm_jit(jit)
    cfg, frame_ids, desc = bench.build_workload("shaded1080", 8, 0, 1)
    rasts = [cfg.rasterizer(i) for i in frame_ids]
    out = torch.empty((8, cfg.height, cfg.width, 4), dtype=torch.uint8, device="cuda:0")
    batch = Rasterizer.prepare_batch(rasts, cfg.scene, cfg.width, cfg.height, cfg.tile_size, cfg.assets)
    for _ in range(2):
        batch.run(out, sync=True)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    n = 5
    for _ in range(n):
        batch.run(out, sync=True)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / n
    print(f"state mode {mode} jit {jit}: {dt * 1e3:8.3f} ms per 8-frame 1080p step ({8 * cfg.width * cfg.height / dt / 1e6:9.0f} Mpixel/s), ordered frames so far {ctx.ordered_frames()}")
# the CPU port (oracle, all host threads) on one frame of the same workload, for scale
sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import oracle_ffi
cfg, frame_ids, desc = bench.build_workload("shaded1080", 1, 0, 1)
r = cfg.rasterizer(frame_ids[0])
oracle_ffi.rasterize(r, cfg.scene, cfg.assets, cfg.width, cfg.height, cfg.tile_size, want_planes=False, n_threads=0)
t0 = time.perf_counter()
for _ in range(3):
    oracle_ffi.rasterize(r, cfg.scene, cfg.assets, cfg.width, cfg.height, cfg.tile_size, want_planes=False, n_threads=0)
dt = (time.perf_counter() - t0) / 3
print(f"CPU port, {oracle_ffi.load().rxo_hardware_threads()} threads: {dt * 1e3:8.1f} ms per 1080p frame ({cfg.width * cfg.height / dt / 1e6:7.0f} Mpixel/s)")
