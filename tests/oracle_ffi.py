"""ctypes binding of the CPU oracle (oracle/rx_oracle.cpp).  TEST INFRASTRUCTURE ONLY: imported by
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs, never by the
rusterix_b200 package."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
import sys

if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from rusterix_b200 import _abi, marshal  # noqa: E402  (struct definitions + marshalling only)

LIB = os.path.join(ROOT, "oracle", "_build", "librxoracle.so")
_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, capture_output=True)
        lib = C.CDLL(LIB)
        lib.rxo_rasterize.restype = C.c_int32
        lib.rxo_rasterize.argtypes = [C.POINTER(_abi.rxc_tile), C.c_uint32, C.POINTER(_abi.rxc_scene),
                                      C.POINTER(_abi.rxc_mapmini), C.POINTER(_abi.rxc_frame), C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_int32]
        lib.rxo_clip_and_project.restype = C.c_int32
        lib.rxo_clip_and_project.argtypes = [C.POINTER(_abi.rxc_batch3d), C.POINTER(_abi.rxc_frame), C.c_void_p,
                                             C.POINTER(C.c_uint32), C.c_void_p, C.c_void_p, C.c_void_p,
                                             C.POINTER(C.c_uint32), C.c_void_p, C.POINTER(C.c_uint32)]
        lib.rxo_hash_u32.restype = C.c_uint32
        lib.rxo_hash_u32.argtypes = [C.c_uint32]
        lib.rxo_edges_evaluate.restype = C.c_int32
        lib.rxo_edges_evaluate.argtypes = [C.c_void_p, C.c_float, C.c_float]
        lib.rxo_light_color_at.restype = C.c_int32
        lib.rxo_light_color_at.argtypes = [C.POINTER(_abi.rxc_light), C.c_void_p, C.c_uint32, C.c_int32, C.c_void_p]
        lib.rxo_light_radiance_at.restype = C.c_int32
        lib.rxo_light_radiance_at.argtypes = [C.POINTER(_abi.rxc_light), C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
        lib.rxo_texture_sample.argtypes = [C.POINTER(_abi.rxc_texture), C.c_float, C.c_float, C.c_uint32, C.c_uint32, C.c_void_p]
        lib.rxo_shade_background.argtypes = [C.POINTER(_abi.rxc_frame), C.c_float, C.c_float, C.c_void_p]
        lib.rxo_shade_fast_brdf.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.rxo_mat4_mul_vec4.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
        lib.rxo_hardware_threads.restype = C.c_uint32
        _lib = lib
    return _lib


def rasterize(rast, scene, assets, width, height, tile_size, want_planes=True, n_threads=0, index_bytes=4):
    """Run the oracle on the same host objects the product API takes.  Returns (pixels[h,w,4],
    owner[h,w] or None, depth[h,w] or None)."""
    lib = load()
    tiles = marshal.marshal_tiles(assets.tile_list)
    sc = marshal.marshal_scene(scene, index_bytes, assets)
    mm = marshal.marshal_mapmini(rast.mapmini)
    frame = marshal.make_frame(rast, scene, width, height, tile_size)
    pixels = np.zeros((height, width, 4), dtype=np.uint8)
    owner = np.zeros((height, width), dtype=np.uint32) if want_planes else None
    depth = np.zeros((height, width), dtype=np.float32) if want_planes else None
    st = lib.rxo_rasterize(tiles.struct, len(assets.tile_list), C.byref(sc.struct), C.byref(mm.struct), C.byref(frame),
                           pixels.ctypes.data, owner.ctypes.data if want_planes else None,
                           depth.ctypes.data if want_planes else None, n_threads)
    if st != 0:
        raise RuntimeError(f"oracle status {st}")
    return pixels, owner, depth


def clip_and_project(rast, scene, batch_index, width, height):
    """Stage outputs of Batch3D::clip_and_project for batch `batch_index` (submission order)."""
    lib = load()
    sc = marshal.marshal_scene(scene)
    frame = marshal.make_frame(rast, scene, width, height, 40)
    b = sc.struct.batches3d[batch_index]
    nv, nt = b.n_vertices, b.n_triangles
    projected = np.zeros((nv + 4 * nt + 1, 4), dtype=np.float32)
    cidx = np.zeros((3 * nt + 1, 3), dtype=np.uint32)
    edges = np.zeros((3 * nt + 1, 9), dtype=np.float32)
    visible = np.zeros(3 * nt + 1, dtype=np.uint8)
    bbox = np.zeros(4, dtype=np.float32)
    n_proj, n_clip, has_bbox = C.c_uint32(), C.c_uint32(), C.c_uint32()
    lib.rxo_clip_and_project(C.byref(b), C.byref(frame), projected.ctypes.data, C.byref(n_proj), cidx.ctypes.data,
                             edges.ctypes.data, visible.ctypes.data, C.byref(n_clip), bbox.ctypes.data, C.byref(has_bbox))
    return dict(projected=projected[:n_proj.value], clipped_indices=cidx[:n_clip.value], edges=edges[:n_clip.value],
                visible=visible[:n_clip.value], bbox=bbox if has_bbox.value else None)
