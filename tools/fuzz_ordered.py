#!/usr/bin/env python
"""Differential fuzz of the reference-order kernel (rxc_set_vm_state_mode 1): the seeded random scenes of tests/test_fuzz_gpu.py that
carry batch-shader programs, rendered in the reference's order with one Execution per API tile and compared with the FAITHFUL oracle
(owner and depth bit for bit).  usage: fuzz_ordered.py <first seed> <scenes with programs>"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import test_fuzz_gpu as fz
from helpers import render_gpu, render_oracle
from rusterix_b200 import DeviceContext, RxcError

first, want = int(sys.argv[1]), int(sys.argv[2])
ctx = DeviceContext.get(0)
ctx.set_vm_state_mode(1)
bad, done, skipped, npx, n1, worst, seed = [], 0, 0, 0, 0, 0, first
while done < want:
    scene, assets, r, w, h, ts = fz._scene(seed)
    seed += 1
    if not scene.shaders:
        continue
    n0 = ctx.ordered_frames()
    try:
        g = render_gpu(r, scene, assets, w, h, ts)
    except RxcError as e:
        if e.status == -3:      # tile_size too large for the mode
            skipped += 1
            continue
        raise
    if ctx.ordered_frames() == n0:      # no batch is bound to a program: the fast kernel rendered it
        continue
    scene2, assets2, r2, _, _, _ = fz._scene(seed - 1)   # a fresh scene: rasterize() appends the chunk lights on every call
    o = render_oracle(r2, scene2, assets2, w, h, ts)
    if (g[1] != o[1]).any() or (g[2].view(np.uint32) != o[2].view(np.uint32)).any():
        bad.append(seed - 1)
    d = np.abs(g[0].astype(np.int16) - o[0].astype(np.int16)).max(axis=-1)
    npx += d.size; n1 += int((d <= 1).sum()); worst = max(worst, int(d.max()))
    done += 1
print(f"reference-order mode, {done} scenes with programs from seed {first} (skipped {skipped}: tile too large): owner/depth mismatching seeds {bad}; "
      f"pixels {npx}, within 1 LSB {n1 / max(1, npx):.6f}, max diff {worst}")
