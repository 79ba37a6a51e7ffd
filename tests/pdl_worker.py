"""Renders scenes that run every kind of kernel chain (separate front-end kernels, the cluster front end, general mode, batch
shaders) and prints a digest of the frames: tests/test_pdl_gpu.py runs it with RXC_PDL=1 and RXC_PDL=0 (the switch is read once per
process) and compares."""
import hashlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

from rusterix_b200 import scenes


def main():
    h = hashlib.sha256()
    for cfg in (scenes.dense(1280, 720, 40, patches=8), scenes.map_config(960, 540, 40, logo_size=256), scenes.chunked_config(640, 360)):
        for frame in (0, 1):
            px = np.zeros((cfg.height, cfg.width, 4), np.uint8)
            cfg.rasterizer(frame).rasterize(cfg.scene, px, cfg.width, cfg.height, cfg.tile_size, cfg.assets)
            h.update(px.tobytes())
    print("PDL_WORKER_OK", h.hexdigest(), flush=True)


if __name__ == "__main__":
    main()
