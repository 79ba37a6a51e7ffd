#!/usr/bin/env python
"""Regenerates tests/golden/oracle_frames.json: SHA-256 of the pixel / owner / depth planes the CPU
oracle produces for small, seeded versions of BASELINE.json's configs.  The reference itself cannot
be run here (Rust, no toolchain) and ships no golden images, so these freeze the ORACLE: any later
edit to oracle/rx_oracle.cpp or to the scene generators that changes a single bit shows up."""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle_ffi  # noqa: E402
from rusterix_b200 import scenes  # noqa: E402

CASES = {
    "cube_200x150_t40": lambda: (scenes.cube(200, 150, 40, logo_size=64), 0),
    "cube_160x160_t200": lambda: (scenes.cube(160, 160, 200, logo_size=64), 0),
    "teapot_240x135_f0": lambda: (scenes.teapot(240, 135, 60, logo_size=64), 0),
    "teapot_240x135_f21": lambda: (scenes.teapot(240, 135, 60, logo_size=64), 21),
    "map_320x180": lambda: (scenes.map_config(320, 180, 40, logo_size=64), 0),
    "sweep_320x180_f1000": lambda: (scenes.sweep(320, 180, 40, logo_size=64), 1000),
    "dense_320x180_p4": lambda: (scenes.dense(320, 180, 40, patches=4), 0),
    # chunk path (opacity layer, surface ids, terrain, occlusion) and 2D game path
    "chunked_320x180_f0": lambda: (scenes.chunked_config(320, 180, 40), 0),
    "chunked_320x180_f5": lambda: (scenes.chunked_config(320, 180, 40), 5),
    "game2d_240x160": lambda: (scenes.game2d_config(240, 160), 0),
    # Rusteria VM programs on batches (the pixel digests also pin this image's libm: sin, pow, ...)
    "shaded_320x240_f0": lambda: (scenes.shaded_config(320, 240, 40), 0),
    "shaded_320x240_f9": lambda: (scenes.shaded_config(320, 240, 40), 9),
    # Sky node miss pass, directional sun, brush preview
    "sky_320x180_f1": lambda: (scenes.sky_config(320, 180, 40), 1),
    "sky_320x180_f6_dawn": lambda: (scenes.sky_config(320, 180, 40, hour=7.0), 6),
}


def digest(case):
    cfg, frame = CASES[case]()
    for chunk in cfg.scene.chunks.values():  # the per-call chunk-light append of rasterize() (src/rasterizer.rs:219-223)
        cfg.scene.dynamic_lights.extend(chunk.lights)
    px, ow, dp = oracle_ffi.rasterize(cfg.rasterizer(frame), cfg.scene, cfg.assets, cfg.width, cfg.height, cfg.tile_size, n_threads=1)
    return {"pixels": hashlib.sha256(px.tobytes()).hexdigest(), "owner": hashlib.sha256(ow.tobytes()).hexdigest(),
            "depth": hashlib.sha256(dp.tobytes()).hexdigest(), "covered": int((ow != 0xFFFFFFFF).sum())}


if __name__ == "__main__":
    out = {k: digest(k) for k in CASES}
    json.dump(out, open(os.path.join(ROOT, "tests", "golden", "oracle_frames.json"), "w"), indent=1, sort_keys=True)
    print(json.dumps(out, indent=1))
