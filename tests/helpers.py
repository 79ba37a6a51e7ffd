"""Shared helpers for the parity tests: run the same host objects through the CUDA product path
(via the C ABI) and through the CPU oracle, and compare with the bar SURVEY.md section 8d states."""
import numpy as np

import oracle_ffi
from rusterix_b200 import Rasterizer

OWNER_NONE = 0xFFFFFFFF


def render_gpu(rast, scene, assets, width, height, tile_size, planes=True, band=None):
    rows = height if band is None else band[1] - band[0]
    pixels = np.zeros((rows, width, 4), dtype=np.uint8)
    owner = np.zeros((rows, width), dtype=np.uint32) if planes else None
    depth = np.zeros((rows, width), dtype=np.float32) if planes else None
    rast.rasterize(scene, pixels, width, height, tile_size, assets, owner=owner, depth=depth, band=band)
    return pixels, owner, depth


def render_oracle(rast, scene, assets, width, height, tile_size, planes=True, index_bytes=4):
    return oracle_ffi.rasterize(rast, scene, assets, width, height, tile_size, want_planes=planes, index_bytes=index_bytes)


def compare(gpu, ref, what="", pixel_frac=0.999):
    """Parity bar: owner and depth planes bit-exact; RGBA8 within +-1 LSB per channel on >= 99.9 %
    of the pixels (north_star).  Returns a dict of statistics."""
    gp, go, gd = gpu
    rp, ro, rd = ref
    stats = {}
    if go is not None and ro is not None:
        stats["owner_mismatch"] = int((go != ro).sum())
        assert stats["owner_mismatch"] == 0, f"{what}: {stats['owner_mismatch']} owner ids differ"
    if gd is not None and rd is not None:
        stats["depth_mismatch"] = int((gd.view(np.uint32) != rd.view(np.uint32)).sum())
        assert stats["depth_mismatch"] == 0, f"{what}: {stats['depth_mismatch']} depth values differ bitwise"
    diff = np.abs(gp.astype(np.int16) - rp.astype(np.int16)).max(axis=-1)
    stats["exact_frac"] = float((diff == 0).mean())
    stats["within1_frac"] = float((diff <= 1).mean())
    stats["max_diff"] = int(diff.max())
    assert stats["within1_frac"] >= pixel_frac, f"{what}: only {stats['within1_frac']:.5f} of pixels within 1 LSB (max diff {stats['max_diff']})"
    return stats
