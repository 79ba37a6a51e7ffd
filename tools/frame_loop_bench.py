#!/usr/bin/env python
"""An engine's frame loop on the 1M-triangle scene: the world stays resident, one entity (a dynamic 3D batch) moves every frame.
Times the scene hand-over of a frame through rxc_set_scene (everything again) and through rxc_update_scene (only what follows the
unchanged batches), C-ABI call alone (host marshalling of the Python mirror excluded), plus the render.
usage: frame_loop_bench.py [frames]"""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from rusterix_b200 import Batch3D, CullMode, DeviceContext, PixelSource, marshal, scenes  # noqa: E402

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 20
cfg = scenes.dense(3840, 2160, 40)
ctx = DeviceContext.get(0)
out = torch.empty((cfg.height, cfg.width, 4), dtype=torch.uint8, device="cuda:0")


def entity(k):
    return (Batch3D.from_box(26.0 + 0.1 * k, 2.0, 10.0, 3.0, 4.5, 3.0).source(PixelSource.StaticTileIndex(0)).cull_mode(CullMode.Off)
            .with_computed_normals())


cfg.scene.d3_dynamic = [entity(0)]
r = cfg.rasterizer(0)
r.rasterize(cfg.scene, out, cfg.width, cfg.height, cfg.tile_size, cfg.assets)    # everything resident
n_static = len(marshal.submission_order(cfg.scene)[0]) - 1
res = {}
for mode in ("rxc_set_scene", "rxc_update_scene"):
    t_call, t_frame = [], []
    for k in range(1, frames + 1):
        cfg.scene.d3_dynamic = [entity(k)]
        m = marshal.marshal_scene(cfg.scene, 4, cfg.assets)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        st = ctx.lib.rxc_set_scene(ctx.handle, C.byref(m.struct)) if mode == "rxc_set_scene" else ctx.lib.rxc_update_scene(ctx.handle, C.byref(m.struct), n_static)
        t1 = time.perf_counter()
        assert st == 0, ctx.lib.rxc_last_error(ctx.handle)
        ctx._scene_key = (cfg.scene._uid, cfg.scene._generation, 4, cfg.scene.structure_key())   # the wrapper need not upload again
        ctx._geometry_keys = marshal.geometry_keys(cfg.scene)
        r.rasterize(cfg.scene, out, cfg.width, cfg.height, cfg.tile_size, cfg.assets)
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        t_call.append(t1 - t0); t_frame.append(t2 - t1)
    t_call.sort(); t_frame.sort()
    res[mode] = (t_call[len(t_call) // 2] * 1e3, t_frame[len(t_frame) // 2] * 1e3)
    print(f"{mode:18s} scene hand-over {res[mode][0]:8.3f} ms (median of {frames}), render of the frame {res[mode][1]:6.3f} ms")
print(f"1M-triangle world + 1 moving entity at 4K: {res['rxc_set_scene'][0] / res['rxc_update_scene'][0]:.1f}x faster hand-over; "
      f"frame loop {1e3 / sum(res['rxc_set_scene']):.0f} -> {1e3 / sum(res['rxc_update_scene']):.0f} frames/s (library time only)")
