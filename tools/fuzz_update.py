#!/usr/bin/env python
"""Differential fuzz of rxc_update_scene: random scenes (tests/test_fuzz_gpu.py), then a few "frames" in which the dynamic and overlay
batches are moved, dropped, duplicated or appended while the chunks and static batches stay -- every frame rendered after the host
mirror's automatic partial upload and again after a forced full rxc_set_scene: pixels, owner and depth must be identical.
usage: fuzz_update.py <first seed> <count>"""
import copy, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import test_fuzz_gpu as fz
from helpers import render_gpu
from rusterix_b200 import DeviceContext

first, count = int(sys.argv[1]), int(sys.argv[2])
ctx = DeviceContext.get(0)
bad, kept_total, frames = [], 0, 0
for seed in range(first, first + count):
    rng = np.random.default_rng(seed)
    scene, assets, r, w, h, ts = fz._scene(seed)
    base_lights = list(scene.dynamic_lights)

    def render(force_full):
        scene.dynamic_lights = list(base_lights)
        if force_full:
            ctx._scene_key = None
        return render_gpu(r, scene, assets, w, h, ts)

    render(True)
    for step in range(4):
        def mutate(lst):
            out = []
            for b in lst:
                k = rng.integers(0, 5)
                if k == 0:
                    continue                                   # dropped
                nb = copy.deepcopy(b)
                nb.vertices = np.array(nb.vertices, dtype=np.float32, copy=True)
                nb.vertices[:, :3] += rng.normal(0.0, 0.3, 3).astype(np.float32)   # moved
                out.append(nb)
                if k == 1:
                    out.append(copy.deepcopy(nb))              # duplicated (equal depths)
            return out
        scene.d3_dynamic = mutate(scene.d3_dynamic) + ([copy.deepcopy(scene.d3_static[0])] if scene.d3_static and rng.random() < 0.3 else [])
        scene.d3_overlay = mutate(scene.d3_overlay)
        a = render(False)
        kept_total += ctx.last_upload_kept
        b = render(True)
        frames += 1
        if not (np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2].view(np.uint32), b[2].view(np.uint32))):
            bad.append((seed, step))
print(f"rxc_update_scene fuzz: seeds {first}..{first + count - 1}, {frames} updated frames, {kept_total} batches kept in total; frames that differ from a full upload: {bad}")
