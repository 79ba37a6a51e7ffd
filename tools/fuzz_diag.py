#!/usr/bin/env python
"""Colour differences of one fuzz seed, broken down: usage fuzz_diag.py <seed>.  Renders the scene as is, then with each
light alone and with none, and prints how many pixels are off by more than 1 LSB in each case plus a few sample pixels."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import test_fuzz_gpu as fz
from helpers import render_gpu, render_oracle

seed = int(sys.argv[1])


def run(keep):
    scene, assets, r, w, h, ts = fz._scene(seed)
    if keep is not None:
        scene.lights = [l for i, l in enumerate(scene.lights) if i in keep]
    g = render_gpu(r, scene, assets, w, h, ts)
    o = render_oracle(r, scene, assets, w, h, ts)
    d = np.abs(g[0].astype(np.int16) - o[0].astype(np.int16)).max(axis=-1)
    return g, o, d, scene


g, o, d, scene = run(None)
print("seed", seed, "frame", d.shape, "pixels off by > 1:", int((d > 1).sum()), "covered", int((o[1] != 0xFFFFFFFF).sum()))
for i, l in enumerate(scene.lights):
    print(" light", i, {k: getattr(l, k) for k in dir(l) if not k.startswith("_") and not callable(getattr(l, k))})
ys, xs = np.nonzero(d > 1)
for y, x in list(zip(ys, xs))[:: max(1, len(ys) // 6)][:6]:
    print("  px", x, y, "gpu", g[0][y, x], "oracle", o[0][y, x], "owner", o[1][y, x], "depth", o[2][y, x])
n = len(scene.lights)
for keep in [()] + [(i,) for i in range(n)]:
    _, _, dd, _ = run(keep)
    print(" lights kept", keep, "-> pixels off by > 1:", int((dd > 1).sum()))
