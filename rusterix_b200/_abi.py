"""ctypes mirror of include/rxcuda.h (the C ABI).  Field order and types must match the header 1:1;
tests/test_abi.py checks sizes/offsets against a C probe compiled from the header."""
import ctypes as C

RXC_ABI_VERSION = 6

RXC_OK = 0
RXC_ERR_INVALID = -1
RXC_ERR_CUDA = -2
RXC_ERR_UNSUPPORTED = -3
RXC_ERR_OOM = -4
RXC_ERR_INDEX = -5
RXC_ERR_NO_DEVICE = -6

STATUS_NAMES = {
    0: "RXC_OK", -1: "RXC_ERR_INVALID", -2: "RXC_ERR_CUDA", -3: "RXC_ERR_UNSUPPORTED",
    -4: "RXC_ERR_OOM", -5: "RXC_ERR_INDEX", -6: "RXC_ERR_NO_DEVICE",
}

RXC_N_KERNELS = 11


class rxc_texture(C.Structure):
    _fields_ = [("data", C.c_void_p), ("width", C.c_uint32), ("height", C.c_uint32)]


class rxc_tile(C.Structure):
    _fields_ = [("textures", C.POINTER(rxc_texture)), ("n_textures", C.c_uint32)]


class rxc_light(C.Structure):
    _fields_ = [
        ("light_type", C.c_uint32),
        ("position", C.c_float * 3),
        ("color", C.c_float * 3),
        ("intensity", C.c_float),
        ("emitting", C.c_uint32),
        ("start_distance", C.c_float),
        ("end_distance", C.c_float),
        ("flicker", C.c_float),
        ("direction", C.c_float * 3),
        ("cone_angle", C.c_float),
        ("normal", C.c_float * 3),
        ("width", C.c_float),
        ("height", C.c_float),
        ("from_linedef", C.c_uint32),
    ]


class rxc_batch3d(C.Structure):
    _fields_ = [
        ("vertices", C.c_void_p),
        ("uvs", C.c_void_p),
        ("normals", C.c_void_p),
        ("indices", C.c_void_p),
        ("n_vertices", C.c_uint32),
        ("n_triangles", C.c_uint32),
        ("index_bytes", C.c_uint32),
        ("mode", C.c_uint32),
        ("repeat_mode", C.c_uint32),
        ("cull_mode", C.c_uint32),
        ("source_kind", C.c_uint32),
        ("source_index", C.c_uint32),
        ("source_pixel", C.c_uint8 * 4),
        ("receives_light", C.c_uint32),
        ("ambient_color", C.c_float * 3),
        ("has_profile_id", C.c_uint32),
        ("profile_id", C.c_uint32),
        ("shader", C.c_int32),
        ("pass_", C.c_uint32),
        ("transform", C.c_float * 16),
        ("chunk", C.c_int32),
    ]


class rxc_batch2d(C.Structure):
    _fields_ = [
        ("vertices", C.c_void_p),
        ("uvs", C.c_void_p),
        ("indices", C.c_void_p),
        ("n_vertices", C.c_uint32),
        ("n_triangles", C.c_uint32),
        ("index_bytes", C.c_uint32),
        ("mode", C.c_uint32),
        ("repeat_mode", C.c_uint32),
        ("source_kind", C.c_uint32),
        ("source_index", C.c_uint32),
        ("source_pixel", C.c_uint8 * 4),
        ("receives_light", C.c_uint32),
        ("shader", C.c_int32),
        ("chunk", C.c_int32),
    ]


class rxc_sector(C.Structure):
    _fields_ = [("min", C.c_float * 2), ("max", C.c_float * 2), ("occlusion", C.c_float)]


class rxc_program(C.Structure):
    _fields_ = [
        ("code", C.c_void_p),
        ("n_words", C.c_uint32),
        ("entry", C.c_uint32),
        ("shade_locals", C.c_uint32),
        ("n_globals", C.c_uint32),
        ("sets_opacity", C.c_uint32),
    ]


class rxc_pattern(C.Structure):
    _fields_ = [("data", C.c_void_p), ("width", C.c_uint32), ("height", C.c_uint32)]


class rxc_chunk(C.Structure):
    _fields_ = [
        ("origin", C.c_int32 * 2),
        ("size", C.c_int32),
        ("occluded_sectors", C.POINTER(rxc_sector)),
        ("n_occluded_sectors", C.c_uint32),
        ("terrain_texture", C.POINTER(rxc_texture)),
        ("shader_base", C.c_uint32),
        ("n_shaders", C.c_uint32),
        ("shader_textures", C.POINTER(C.POINTER(rxc_texture))),
    ]


class rxc_linedef(C.Structure):
    _fields_ = [("start", C.c_float * 2), ("end", C.c_float * 2)]


class rxc_mapmini(C.Structure):
    _fields_ = [
        ("linedefs", C.POINTER(rxc_linedef)),
        ("n_linedefs", C.c_uint32),
        ("occluded_sectors", C.POINTER(rxc_sector)),
        ("n_occluded_sectors", C.c_uint32),
    ]


class rxc_scene(C.Structure):
    _fields_ = [
        ("batches3d", C.POINTER(rxc_batch3d)),
        ("n_batches3d", C.c_uint32),
        ("batches2d", C.POINTER(rxc_batch2d)),
        ("n_batches2d", C.c_uint32),
        ("lights", C.POINTER(rxc_light)),
        ("n_lights", C.c_uint32),
        ("dynamic_textures", C.POINTER(rxc_tile)),
        ("n_dynamic_textures", C.c_uint32),
        ("chunks", C.POINTER(rxc_chunk)),
        ("n_chunks", C.c_uint32),
        ("actor_tiles", C.POINTER(rxc_tile)),
        ("n_actor_tiles", C.c_uint32),
        ("shaders", C.POINTER(rxc_program)),
        ("n_shaders", C.c_uint32),
        ("n_scene_shaders", C.c_uint32),
        ("patterns", C.POINTER(rxc_pattern)),
        ("n_patterns", C.c_uint32),
        ("patterns_normal", C.POINTER(rxc_pattern)),
        ("n_patterns_normal", C.c_uint32),
        ("palette", C.c_void_p),
        ("n_palette", C.c_uint32),
    ]


class rxc_frame(C.Structure):
    _fields_ = [
        ("view", C.c_float * 16),
        ("projection", C.c_float * 16),
        ("inverse_view", C.c_float * 16),
        ("inverse_projection", C.c_float * 16),
        ("has_matrix2d", C.c_uint32),
        ("matrix2d", C.c_float * 9),
        ("width", C.c_uint32),
        ("height", C.c_uint32),
        ("tile_size", C.c_uint32),
        ("sample_mode", C.c_uint32),
        ("has_background_color", C.c_uint32),
        ("background_color", C.c_uint8 * 4),
        ("background_shader", C.c_uint32),
        ("grid_size", C.c_float),
        ("grid_subdivisions", C.c_float),
        ("grid_offset", C.c_float * 2),
        ("has_ambient", C.c_uint32),
        ("ambient", C.c_float * 4),
        ("animation_frame", C.c_uint64),
        ("time", C.c_float),
        ("hour", C.c_float),
        ("d2_active", C.c_uint32),
        ("d3_active", C.c_uint32),
        ("ignore_background_shader", C.c_uint32),
        ("preserve_transparency", C.c_uint32),
        ("matvec_mode", C.c_uint32),
        ("band_y0", C.c_uint32),
        ("band_y1", C.c_uint32),
        ("band_x0", C.c_uint32),
        ("band_x1", C.c_uint32),
        ("has_sun", C.c_uint32),
        ("sun_dir", C.c_float * 3),
        ("day_factor", C.c_float),
        ("has_sky", C.c_uint32),
        ("sky", (C.c_float * 4) * 6),
        ("sky_clouds", C.c_uint32),
        ("has_brush_preview", C.c_uint32),
        ("brush_position", C.c_float * 3),
        ("brush_radius", C.c_float),
        ("brush_falloff", C.c_float),
    ]


class rxc_stats(C.Structure):
    _fields_ = [
        ("frames", C.c_uint64),
        ("kernel_launches", C.c_uint64),
        ("launches", C.c_uint64 * RXC_N_KERNELS),
        ("kernel_ms", C.c_double * RXC_N_KERNELS),
        ("h2d_bytes", C.c_uint64),
        ("d2h_bytes", C.c_uint64),
        ("last_binned_refs", C.c_uint32),
        ("last_large_tris", C.c_uint32),
        ("last_clipped_tris", C.c_uint32),
        ("last_visible_tris", C.c_uint32),
    ]


class rxc_projected3d(C.Structure):
    _fields_ = [
        ("projected_vertices", C.c_void_p),
        ("clipped_uvs", C.c_void_p),
        ("clipped_normals", C.c_void_p),
        ("n_projected", C.c_uint32),
        ("n_clipped", C.c_uint32),
        ("clipped_indices", C.c_void_p),
        ("index_bytes", C.c_uint32),
        ("has_bounding_box", C.c_uint32),
        ("edges", C.c_void_p),
        ("visible", C.c_void_p),
        ("bounding_box", C.c_float * 4),
    ]


RXC_MGPU_ID_BYTES = 128
RXC_MGPU_LOCAL, RXC_MGPU_PEER, RXC_MGPU_NCCL = 0, 1, 2


class rxc_mgpu_region(C.Structure):
    _fields_ = [("rank", C.c_uint32), ("rows", C.c_uint32), ("offset", C.c_uint64), ("row_bytes", C.c_uint64), ("pitch_bytes", C.c_uint64)]


# every symbol include/rxcuda.h declares: (name, restype, argtypes)
EXPORTS = [
    ("rxc_abi_version", C.c_uint32, []),
    ("rxc_create", C.c_int32, [C.c_int32, C.POINTER(C.c_void_p)]),
    ("rxc_destroy", None, [C.c_void_p]),
    ("rxc_last_error", C.c_char_p, [C.c_void_p]),
    ("rxc_set_stream", C.c_int32, [C.c_void_p, C.c_void_p]),
    ("rxc_set_assets", C.c_int32, [C.c_void_p, C.POINTER(rxc_tile), C.c_uint32]),
    ("rxc_set_scene", C.c_int32, [C.c_void_p, C.POINTER(rxc_scene)]),
    ("rxc_update_scene", C.c_int32, [C.c_void_p, C.POINTER(rxc_scene), C.c_uint32]),
    ("rxc_set_lights", C.c_int32, [C.c_void_p, C.POINTER(rxc_light), C.c_uint32]),
    ("rxc_set_mapmini", C.c_int32, [C.c_void_p, C.POINTER(rxc_mapmini)]),
    ("rxc_rasterize", C.c_int32, [C.c_void_p, C.POINTER(rxc_frame), C.c_void_p, C.c_void_p, C.c_void_p]),
    ("rxc_rasterize_async", C.c_int32, [C.c_void_p, C.POINTER(rxc_frame), C.c_void_p, C.c_void_p, C.c_void_p]),
    ("rxc_rasterize_batch", C.c_int32, [C.c_void_p, C.POINTER(rxc_frame), C.c_uint32, C.c_void_p, C.c_uint64]),
    ("rxc_rasterize_batch_async", C.c_int32, [C.c_void_p, C.POINTER(rxc_frame), C.c_uint32, C.c_void_p, C.c_uint64]),
    ("rxc_rasterize_projected", C.c_int32, [C.c_void_p, C.POINTER(rxc_frame), C.POINTER(rxc_projected3d), C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]),
    ("rxc_synchronize", C.c_int32, [C.c_void_p]),
    ("rxc_pin_host", C.c_int32, [C.c_void_p, C.c_void_p, C.c_uint64]),
    ("rxc_unpin_host", C.c_int32, [C.c_void_p, C.c_void_p]),
    ("rxc_owner_base", C.c_int32, [C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32)]),
    ("rxc_selftest_div", C.c_int32, [C.c_void_p, C.c_uint64, C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)]),
    ("rxc_vm_execute", C.c_int32, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.POINTER(C.c_uint32)]),
    ("rxc_mgpu_unique_id", C.c_int32, [C.POINTER(C.c_uint8)]),
    ("rxc_mgpu_init", C.c_int32, [C.c_void_p, C.POINTER(C.c_uint8), C.c_uint32, C.c_uint32]),
    ("rxc_mgpu_shutdown", C.c_int32, [C.c_void_p]),
    ("rxc_mgpu_target", C.c_int32, [C.c_void_p, C.c_uint64, C.POINTER(C.c_void_p), C.POINTER(C.c_uint32)]),
    ("rxc_mgpu_rasterize", C.c_int32, [C.c_void_p, C.POINTER(rxc_frame), C.c_uint32, C.c_uint64, C.c_uint64, C.c_uint64]),
    ("rxc_mgpu_deliver", C.c_int32, [C.c_void_p, C.POINTER(rxc_mgpu_region), C.c_uint32]),
    ("rxc_mgpu_release", C.c_int32, [C.c_void_p]),
    ("rxc_mgpu_status", C.c_int32, [C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    ("rxc_vm_translate", C.c_int64, [C.POINTER(rxc_program), C.c_uint32, C.c_char_p, C.c_uint64, C.POINTER(C.c_uint32)]),
    ("rxc_vm_jit_compile", C.c_int64, [C.POINTER(rxc_program), C.c_uint32, C.c_int32, C.c_int32, C.c_char_p, C.c_uint32]),
    ("rxc_set_vm_jit", C.c_int32, [C.c_void_p, C.c_int32]),
    ("rxc_set_vm_state_mode", C.c_int32, [C.c_void_p, C.c_int32]),
    ("rxc_get_vm_state_mode", C.c_int32, [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_uint64)]),
    ("rxc_vm_state_report", C.c_int32, [C.POINTER(rxc_program), C.c_uint32, C.POINTER(C.c_uint8), C.c_int32, C.POINTER(C.c_uint32)]),
    ("rxc_vm_scene_state_report", C.c_int32, [C.c_void_p, C.POINTER(C.c_uint32), C.c_uint32, C.POINTER(C.c_uint32)]),
    ("rxc_vm_jit_info", C.c_int32, [C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64), C.c_char_p, C.c_uint32]),
    ("rxc_set_profiling", C.c_int32, [C.c_void_p, C.c_int32]),
    ("rxc_get_stats", C.c_int32, [C.c_void_p, C.POINTER(rxc_stats)]),
    ("rxc_reset_stats", C.c_int32, [C.c_void_p]),
    ("rxc_kernel_name", C.c_char_p, [C.c_uint32]),
]


def bind(lib):
    """Attach restype/argtypes for every exported symbol; raises AttributeError if one is missing."""
    for name, restype, argtypes in EXPORTS:
        fn = getattr(lib, name)
        fn.restype = restype
        fn.argtypes = argtypes
    return lib
