"""Scene / Assets / Rasterizer state -> the POD structs of include/rxcuda.h.

This is the Python twin of what the Rust `-sys` wrapper does: flatten the scene's
batch lists into submission order (reference src/rasterizer.rs:314-405, :501-553), point the PODs
at the host arrays, and keep those arrays alive for the duration of the call."""
import ctypes as C

import numpy as np

from . import _abi
from .types import Assets, Batch2D, Batch3D, CompiledLight, Scene, Tile
from .vekmath import to_cols


def _ptr(a: np.ndarray):
    return a.ctypes.data if a.size else None


class Marshalled:
    """Holds a ctypes struct plus everything it points to."""

    def __init__(self, struct, keep):
        self.struct = struct
        self.keep = keep


def marshal_tiles(tiles):
    keep = []
    arr = (_abi.rxc_tile * max(1, len(tiles)))()
    for i, tile in enumerate(tiles):
        tex = (_abi.rxc_texture * max(1, len(tile.textures)))()
        for j, t in enumerate(tile.textures):
            tex[j].data = t.data.ctypes.data
            tex[j].width = t.width
            tex[j].height = t.height
            keep.append(t.data)
        arr[i].textures = tex
        arr[i].n_textures = len(tile.textures)
        keep.append(tex)
    return Marshalled(arr, keep)


def marshal_lights(lights):
    arr = (_abi.rxc_light * max(1, len(lights)))()
    for i, l in enumerate(lights):
        o = arr[i]
        o.light_type = int(l.light_type)
        o.position[:] = [float(x) for x in l.position]
        o.color[:] = [float(x) for x in l.color]
        o.intensity = l.intensity
        o.emitting = 1 if l.emitting else 0
        o.start_distance = l.start_distance
        o.end_distance = l.end_distance
        o.flicker = l.flicker
        o.direction[:] = [float(x) for x in l.direction]
        o.cone_angle = l.cone_angle
        o.normal[:] = [float(x) for x in l.normal]
        o.width = l.width
        o.height = l.height
        o.from_linedef = 1 if l.from_linedef else 0
    return Marshalled(arr, [])


def marshal_sectors(sectors):
    arr = (_abi.rxc_sector * max(1, len(sectors)))()
    for i, (bbox, occlusion) in enumerate(sectors):
        arr[i].min[:] = [float(bbox.min[0]), float(bbox.min[1])]
        arr[i].max[:] = [float(bbox.max[0]), float(bbox.max[1])]
        arr[i].occlusion = float(occlusion)
    return arr


def marshal_mapmini(mapmini):
    m = _abi.rxc_mapmini()
    lines = (_abi.rxc_linedef * max(1, len(mapmini.linedefs)))()
    for i, l in enumerate(mapmini.linedefs):
        lines[i].start[:] = [float(l.start[0]), float(l.start[1])]
        lines[i].end[:] = [float(l.end[0]), float(l.end[1])]
    sectors = marshal_sectors(mapmini.occluded_sectors)
    m.linedefs = lines
    m.n_linedefs = len(mapmini.linedefs)
    m.occluded_sectors = sectors
    m.n_occluded_sectors = len(mapmini.occluded_sectors)
    return Marshalled(m, [lines, sectors])


def submission_order(scene: Scene):
    """The order in which the reference walks the batches of a scene (src/rasterizer.rs:314-405 and
    :501-553): (batch, pass tag, chunk index) for 3D and (batch, chunk index) for 2D."""
    b3, b2 = [], []
    for ci, chunk in enumerate(scene.chunks.values()):
        b3 += [(b, 4, ci) for b in chunk.batches3d_opacity]
        b3 += [(b, 3, ci) for b in chunk.batches3d]
        if chunk.terrain_batch3d is not None:
            b3.append((chunk.terrain_batch3d, 3, ci))
        b2 += [(b, ci) for b in chunk.batches2d]
        if chunk.terrain_batch2d is not None:
            b2.append((chunk.terrain_batch2d, ci))
    b3 += [(b, 0, -1) for b in scene.d3_static] + [(b, 1, -1) for b in scene.d3_dynamic] + [(b, 2, -1) for b in scene.d3_overlay]
    b2 += [(b, -1) for b in scene.d2_static] + [(b, -1) for b in scene.d2_dynamic]
    return b3, b2


def geometry_keys(scene: Scene):
    """Per 3D batch in submission order: the identity of its vertex / uv / normal / index arrays (DeviceContext.upload keeps
    the leading batches whose arrays are the ones already on the device)."""
    b3, _b2 = submission_order(scene)
    return [(id(b), id(b.vertices), id(b.uvs), id(getattr(b, "normals", None)), id(b.indices), len(b.vertices), len(b.indices)) for b, _p, _c in b3]


class _ActorTiles:
    """EntityTile / ItemTile sources are resolved on the host (the id -> IndexMap lookup of
    src/rasterizer.rs:1130-1177) to indices into one `actor_tiles` array."""

    def __init__(self, assets):
        self.assets = assets
        self.tiles = []
        self._index = {}

    def resolve(self, src) -> int:
        tile = self.assets.actor_tile(src) if self.assets is not None else None
        if tile is None or not tile.textures:
            return 0xFFFFFFFF
        k = id(tile)
        if k not in self._index:
            self._index[k] = len(self.tiles)
            self.tiles.append(tile)
        return self._index[k]


def _fill_source(o, src, actors):
    o.source_kind = int(src.kind)
    o.source_index = int(src.index)
    if int(src.kind) in (4, 5):
        o.source_index = actors.resolve(src)
    o.source_pixel[:] = list(src.pixel)


def marshal_scene(scene: Scene, index_bytes: int = 4, assets: Assets = None):
    """index_bytes=8 marshals indices as Rust `usize` triples (24 B/triangle) to exercise that path."""
    keep = []
    actors = _ActorTiles(assets)
    b3_list, b2_list = submission_order(scene)
    b3 = (_abi.rxc_batch3d * max(1, len(b3_list)))()
    for i, (b, pass_, chunk_index) in enumerate(b3_list):
        o = b3[i]
        idx = b.indices if index_bytes == 4 else np.ascontiguousarray(b.indices.astype(np.uint64))
        keep += [b.vertices, b.uvs, b.normals, idx]
        o.vertices = _ptr(b.vertices)
        o.uvs = _ptr(b.uvs)
        o.normals = _ptr(b.normals) if len(b.normals) else None
        o.indices = _ptr(idx)
        o.n_vertices = len(b.vertices)
        o.n_triangles = len(b.indices)
        o.index_bytes = index_bytes
        if len(b.uvs) != len(b.vertices):
            raise ValueError("Batch3D.uvs must have one entry per vertex")
        if len(b.normals) not in (0, len(b.vertices)):
            raise ValueError("Batch3D.normals must be empty or have one entry per vertex")
        o.mode = int(b.mode)
        o.repeat_mode = int(b.repeat_mode_)
        o.cull_mode = int(b.cull_mode_)
        _fill_source(o, b.source_, actors)
        o.receives_light = 1 if b.receives_light_ else 0
        o.ambient_color[:] = list(b.ambient_color_)
        o.has_profile_id = 0 if b.profile_id_ is None else 1
        o.profile_id = 0 if b.profile_id_ is None else b.profile_id_
        o.shader = -1 if b.shader_ is None else b.shader_
        o.pass_ = pass_
        o.transform[:] = to_cols(b.transform_3d).tolist()
        o.chunk = chunk_index
    b2 = (_abi.rxc_batch2d * max(1, len(b2_list)))()
    for i, (b, chunk_index) in enumerate(b2_list):
        o = b2[i]
        idx = b.indices if index_bytes == 4 else np.ascontiguousarray(b.indices.astype(np.uint64))
        keep += [b.vertices, b.uvs, idx]
        o.vertices = _ptr(b.vertices)
        o.uvs = _ptr(b.uvs)
        o.indices = _ptr(idx)
        o.n_vertices = len(b.vertices)
        o.n_triangles = len(b.indices)
        o.index_bytes = index_bytes
        o.mode = int(b.mode)
        o.repeat_mode = int(b.repeat_mode_)
        _fill_source(o, b.source_, actors)
        o.receives_light = 1 if b.receives_light_ else 0
        o.shader = -1 if b.shader_ is None else b.shader_
        o.chunk = chunk_index
    lights = marshal_lights(scene.all_lights())
    dyn = marshal_tiles(scene.dynamic_textures)
    act = marshal_tiles(actors.tiles)
    # Rusteria VM programs: scene.shaders, then every chunk's (rxc_chunk.shader_base)
    flat_programs = [p.flatten() for p in scene.shaders]
    chunk_base = []
    for chunk in scene.chunks.values():
        chunk_base.append(len(flat_programs))
        flat_programs += [p.flatten() for p in chunk.shaders]
    progs = (_abi.rxc_program * max(1, len(flat_programs)))()
    for i, fp in enumerate(flat_programs):
        o = progs[i]
        keep.append(fp.words)
        o.code = fp.words.ctypes.data if len(fp.words) else None
        o.n_words = len(fp.words)
        o.entry, o.shade_locals, o.n_globals = fp.entry, fp.shade_locals, fp.n_globals
        o.sets_opacity = 1 if fp.sets_opacity else 0

    def marshal_patterns(bank):
        arr = (_abi.rxc_pattern * max(1, len(bank)))()
        for i, (w, h, data) in enumerate(bank):
            d = np.ascontiguousarray(np.asarray(data, dtype=np.float32).reshape(h * w, 3))
            keep.append(d)
            arr[i].data, arr[i].width, arr[i].height = d.ctypes.data, int(w), int(h)
        return arr
    pats, pats_n = marshal_patterns(scene.patterns), marshal_patterns(scene.patterns_normal)
    pal_src = list(assets.palette) if assets is not None else []
    palette = np.zeros((max(1, len(pal_src)), 4), dtype=np.float32)
    for i, col in enumerate(pal_src):
        if col is not None:
            palette[i] = [1.0, col[0], col[1], col[2]]
    keep += [progs, pats, pats_n, palette]

    chunks = (_abi.rxc_chunk * max(1, len(scene.chunks)))()
    for ci, chunk in enumerate(scene.chunks.values()):
        c = chunks[ci]
        c.shader_base = chunk_base[ci]
        c.n_shaders = len(chunk.shaders)
        if any(t is not None for t in chunk.shader_textures):
            ptrs = (C.POINTER(_abi.rxc_texture) * max(1, len(chunk.shaders)))()
            for k, tex in enumerate(chunk.shader_textures[:len(chunk.shaders)]):
                if tex is not None:
                    t = _abi.rxc_texture()
                    t.data, t.width, t.height = tex.data.ctypes.data, tex.width, tex.height
                    ptrs[k] = C.pointer(t)
                    keep += [t, tex.data]
            c.shader_textures = ptrs
            keep.append(ptrs)
        c.origin[:] = [int(chunk.origin[0]), int(chunk.origin[1])]
        c.size = int(chunk.size)
        sectors = marshal_sectors(chunk.occluded_sectors)
        c.occluded_sectors = sectors
        c.n_occluded_sectors = len(chunk.occluded_sectors)
        keep.append(sectors)
        if chunk.terrain_texture is not None:
            t = _abi.rxc_texture()
            t.data = chunk.terrain_texture.data.ctypes.data
            t.width = chunk.terrain_texture.width
            t.height = chunk.terrain_texture.height
            c.terrain_texture = C.pointer(t)
            keep += [t, chunk.terrain_texture.data]
    s = _abi.rxc_scene()
    s.batches3d = b3
    s.n_batches3d = len(b3_list)
    s.batches2d = b2
    s.n_batches2d = len(b2_list)
    s.lights = lights.struct
    s.n_lights = len(scene.all_lights())
    s.dynamic_textures = dyn.struct
    s.n_dynamic_textures = len(scene.dynamic_textures)
    s.chunks = chunks
    s.n_chunks = len(scene.chunks)
    s.actor_tiles = act.struct
    s.n_actor_tiles = len(actors.tiles)
    s.shaders = progs
    s.n_shaders = len(flat_programs)
    s.n_scene_shaders = len(scene.shaders)
    s.patterns, s.n_patterns = pats, len(scene.patterns)
    s.patterns_normal, s.n_patterns_normal = pats_n, len(scene.patterns_normal)
    s.palette, s.n_palette = palette.ctypes.data, len(pal_src)
    keep += [b3, b2, lights, dyn, act, chunks]
    return Marshalled(s, keep)


def make_frame(rast, scene: Scene, width: int, height: int, tile_size: int, band=None) -> _abi.rxc_frame:
    """`rast` is a rusterix_b200.rasterizer.Rasterizer (fields mirror src/rasterizer.rs:35-88)."""
    f = _abi.rxc_frame()
    f.view[:] = to_cols(rast.view_matrix).tolist()
    f.projection[:] = to_cols(rast.projection_matrix).tolist()
    f.inverse_view[:] = to_cols(rast.inverse_view_matrix).tolist()
    f.inverse_projection[:] = to_cols(rast.inverse_projection_matrix).tolist()
    if rast.projection_matrix_2d is not None:
        f.has_matrix2d = 1
        f.matrix2d[:] = to_cols(np.asarray(rast.projection_matrix_2d, dtype=np.float32).reshape(3, 3)).tolist()
    f.width, f.height, f.tile_size = int(width), int(height), int(tile_size)
    f.sample_mode = int(rast.sample_mode_)
    if rast.background_color is not None:
        f.has_background_color = 1
        f.background_color[:] = list(rast.background_color)
    bg = scene.background
    if bg is not None:
        f.background_shader = int(bg.kind)
        if bg.kind == 2:
            f.grid_size = bg.grid_size
            f.grid_subdivisions = bg.subdivisions
            f.grid_offset[:] = list(bg.offset)
    if rast.ambient_color is not None:
        f.has_ambient = 1
        f.ambient[:] = [float(x) for x in rast.ambient_color]
    f.animation_frame = int(scene.animation_frame)
    f.time = rast.time_
    f.hour = rast.hour
    f.d2_active = 1 if rast.render_mode_.d2_active else 0
    f.d3_active = 1 if rast.render_mode_.d3_active else 0
    f.ignore_background_shader = 1 if rast.render_mode_.ignore_background_shader_ else 0
    f.preserve_transparency = 1 if rast.preserve_transparency else 0
    f.matvec_mode = int(rast.matvec_mode)
    if band is not None:
        f.band_y0, f.band_y1 = int(band[0]), int(band[1])
        if len(band) == 4:   # (y0, y1, x0, x1): a rectangle of the frame
            f.band_x0, f.band_x1 = int(band[2]), int(band[3])
    # render graph results (Rasterizer.prepare_render_graph) and the brush preview
    if getattr(rast, "sun_dir", None) is not None:
        f.has_sun = 1
        f.sun_dir[:] = [float(c) for c in rast.sun_dir]
        f.day_factor = float(rast.day_factor)
    sky = [n for n in getattr(rast, "render_miss", []) if getattr(n, "precomputed", None)]
    if sky:
        f.has_sky = 1
        for i, row in enumerate(sky[-1].precomputed):
            f.sky[i][:] = [float(c) for c in row]
        f.sky_clouds = 1 if sky[-1].clouds else 0
    bp = getattr(rast, "brush_preview", None)
    if bp is not None:
        f.has_brush_preview = 1
        f.brush_position[:] = [float(c) for c in bp.position]
        f.brush_radius, f.brush_falloff = float(bp.radius), float(bp.falloff)
    return f


def marshal_projected(projected, index_bytes: int = 4):
    """rxc_projected3d array from a list (one entry per 3D batch, submission order) of dicts / objects holding what
    `Scene::project` leaves in a Batch3D: projected_vertices [n,4], clipped_uvs [n,2], clipped_normals [n,3] or None,
    clipped_indices [m,3], edges [m,9] (a[3], b[3], c[3]), visible [m], bounding_box (x, y, width, height) or None."""
    n = len(projected)
    arr = (_abi.rxc_projected3d * max(1, n))()
    keep = []

    def get(p, k):
        return p[k] if isinstance(p, dict) else getattr(p, k)

    for i, p in enumerate(projected):
        pv = np.ascontiguousarray(get(p, "projected_vertices"), dtype=np.float32).reshape(-1, 4)
        uv = np.ascontiguousarray(get(p, "clipped_uvs"), dtype=np.float32).reshape(-1, 2)
        nr = get(p, "clipped_normals")
        nr = None if nr is None else np.ascontiguousarray(nr, dtype=np.float32).reshape(-1, 3)
        ci = np.ascontiguousarray(get(p, "clipped_indices"), dtype=np.uint64 if index_bytes == 8 else np.uint32).reshape(-1, 3)
        ed = np.ascontiguousarray(get(p, "edges"), dtype=np.float32).reshape(-1, 9)
        vi = np.ascontiguousarray(get(p, "visible"), dtype=np.uint8).reshape(-1)
        bb = get(p, "bounding_box")
        if len(uv) != len(pv) or (nr is not None and len(nr) != len(pv)) or len(ed) != len(ci) or len(vi) != len(ci):
            raise ValueError(f"projected batch {i}: array lengths disagree")
        keep += [pv, uv, nr, ci, ed, vi]
        a = arr[i]
        a.projected_vertices, a.clipped_uvs = pv.ctypes.data, uv.ctypes.data
        a.clipped_normals = nr.ctypes.data if nr is not None else None
        a.n_projected, a.n_clipped = len(pv), len(ci)
        a.clipped_indices, a.index_bytes = ci.ctypes.data, index_bytes
        a.edges, a.visible = ed.ctypes.data, vi.ctypes.data
        a.has_bounding_box = 0 if bb is None else 1
        if bb is not None:
            a.bounding_box[:] = [float(c) for c in bb]
    return Marshalled(arr, keep)
