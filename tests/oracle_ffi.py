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
        lib.rxo_clip_and_project_attrs.restype = C.c_int32
        lib.rxo_clip_and_project_attrs.argtypes = [C.POINTER(_abi.rxc_batch3d), C.POINTER(_abi.rxc_frame), C.c_void_p, C.c_void_p]
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
        lib.rxo_mat4_mul_mat4.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
        lib.rxo_hardware_threads.restype = C.c_uint32
        lib.rxo_set_programs.restype = C.c_int32
        lib.rxo_set_programs.argtypes = [C.POINTER(C.c_void_p), C.c_void_p, C.c_uint32, C.POINTER(_abi.rxc_pattern), C.c_uint32,
                                         C.POINTER(_abi.rxc_pattern), C.c_uint32, C.c_void_p, C.c_uint32]
        lib.rxo_vm_execute.restype = C.c_uint32
        lib.rxo_vm_execute.argtypes = [C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]
        _lib = lib
    return _lib


def set_programs(scene, assets=None):
    """Hands the oracle the op TREES of the scene's programs (scene.shaders, then every chunk's) plus the
    pattern banks and the palette; rasterize() calls it before every frame."""
    lib = load()
    programs = list(scene.shaders)
    for chunk in scene.chunks.values():
        programs += list(chunk.shaders)
    trees = [p.encode_tree() for p in programs]
    ptrs = (C.c_void_p * max(1, len(trees)))(*[t.ctypes.data for t in trees])
    lens = np.array([len(t) for t in trees] + [0], dtype=np.uint32)

    def bank(b):
        arr = (_abi.rxc_pattern * max(1, len(b)))()
        keep = []
        for i, (w, h, data) in enumerate(b):
            d = np.ascontiguousarray(np.asarray(data, dtype=np.float32).reshape(h * w, 3))
            keep.append(d)
            arr[i].data, arr[i].width, arr[i].height = d.ctypes.data, int(w), int(h)
        return arr, keep
    pats, k1 = bank(scene.patterns)
    pats_n, k2 = bank(scene.patterns_normal)
    pal_src = list(assets.palette) if assets is not None else []
    palette = np.zeros((max(1, len(pal_src)), 4), dtype=np.float32)
    for i, col in enumerate(pal_src):
        if col is not None:
            palette[i] = [1.0, col[0], col[1], col[2]]
    st = lib.rxo_set_programs(ptrs, lens.ctypes.data, len(trees), pats, len(scene.patterns), pats_n, len(scene.patterns_normal),
                              palette.ctypes.data, len(pal_src))
    assert st == 0 and (k1 is not None) and (k2 is not None)


def set_vm_state_mode(per_fragment: bool):
    """False (default): the reference's per-tile Execution that is never reset; True: the device's documented
    deviation, a fresh Execution per fragment."""
    load().rxo_set_vm_state_mode(1 if per_fragment else 0)


def vm_execute(program: int, records):
    lib = load()
    rec = np.ascontiguousarray(records, dtype=np.float32).reshape(-1, 18)
    out = np.zeros((len(rec), 24), dtype=np.float32)
    faults = lib.rxo_vm_execute(int(program), len(rec), rec.ctypes.data, out.ctypes.data)
    return out, int(faults)


def rasterize(rast, scene, assets, width, height, tile_size, want_planes=True, n_threads=0, index_bytes=4):
    """Run the oracle on the same host objects the product API takes.  Returns (pixels[h,w,4],
    owner[h,w] or None, depth[h,w] or None)."""
    lib = load()
    if hasattr(rast, "prepare_render_graph"):
        rast.prepare_render_graph()   # src/rasterizer.rs:227-253 (host side)
    set_programs(scene, assets)
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


def project_scene(rast, scene, assets, width, height, tile_size=40):
    """What the reference's `Scene::project` (src/scene.rs:154-200) leaves in every 3D batch, computed by the oracle's
    clip_and_project: a list (submission order) of dicts with projected_vertices, clipped_uvs, clipped_normals (None
    when the batch has no normals), clipped_indices, edges, visible and bounding_box (None = Option::None).  This is
    the host-side input of the pre-projected entry (Rasterizer.rasterize_projected)."""
    lib = load()
    sc = marshal.marshal_scene(scene, 4, assets)
    frame = marshal.make_frame(rast, scene, width, height, tile_size)
    out = []
    for bi in range(sc.struct.n_batches3d):
        b = sc.struct.batches3d[bi]
        nv, nt = b.n_vertices, b.n_triangles
        projected = np.zeros((nv + 4 * nt + 1, 4), dtype=np.float32)
        cidx = np.zeros((3 * nt + 1, 3), dtype=np.uint32)
        edges = np.zeros((3 * nt + 1, 9), dtype=np.float32)
        visible = np.zeros(3 * nt + 1, dtype=np.uint8)
        bbox = np.zeros(4, dtype=np.float32)
        uvs = np.zeros((nv + 4 * nt + 1, 2), dtype=np.float32)
        nrm = np.zeros((nv + 4 * nt + 1, 3), dtype=np.float32)
        n_proj, n_clip, has_bbox = C.c_uint32(), C.c_uint32(), C.c_uint32()
        lib.rxo_clip_and_project(C.byref(b), C.byref(frame), projected.ctypes.data, C.byref(n_proj), cidx.ctypes.data,
                                 edges.ctypes.data, visible.ctypes.data, C.byref(n_clip), bbox.ctypes.data, C.byref(has_bbox))
        lib.rxo_clip_and_project_attrs(C.byref(b), C.byref(frame), uvs.ctypes.data, nrm.ctypes.data)
        out.append(dict(projected_vertices=projected[:n_proj.value].copy(), clipped_uvs=uvs[:n_proj.value].copy(),
                        clipped_normals=nrm[:n_proj.value].copy() if b.normals else None,
                        clipped_indices=cidx[:n_clip.value].copy(), edges=edges[:n_clip.value].copy(), visible=visible[:n_clip.value].copy(),
                        bounding_box=bbox.copy() if has_bbox.value else None))
    return out
