"""Host-side mirror of the reference's data types on the rasterize path: Batch3D / Batch2D / Scene /
Texture / Tile / Assets / CompiledLight / PixelSource and the enums.  Same names, builder methods
and defaults as the Rust types so tests read like reference call sites; the data lives in numpy
arrays that are marshalled into the C-ABI PODs of include/rxcuda.h by `marshal.py`.

Reference: src/batch/batch3d.rs:15-137,421-479,812-842; src/batch/batch2d.rs:10-127;
src/scene.rs:8-150; src/texture.rs:46-54; src/map/tile.rs:83-110; src/map/light.rs:128-193,457-477;
src/map/pixelsource.rs:23-37; src/server/assets.rs:19."""
from __future__ import annotations

import enum
import itertools
import math
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

from . import vekmath

_UIDS = itertools.count(1)


class PrimitiveMode(enum.IntEnum):  # src/batch/mod.rs:5-15
    Triangles = 0
    Lines = 1
    LineStrip = 2
    LineLoop = 3


class CullMode(enum.IntEnum):  # src/batch/mod.rs:18-26
    Off = 0
    Front = 1
    Back = 2


class RepeatMode(enum.IntEnum):  # src/texture.rs:15-25
    ClampXY = 0
    RepeatXY = 1
    RepeatX = 2
    RepeatY = 3


class SampleMode(enum.IntEnum):  # src/texture.rs:6-12
    Nearest = 0
    Linear = 1


class LightType(enum.IntEnum):  # src/map/light.rs:7-14
    Point = 0
    Ambient = 1
    AmbientDaylight = 2
    Spot = 3
    Area = 4
    Daylight = 5


class MatVecMode(enum.IntEnum):
    FmaColumns = 0
    PlainRows = 1


class _SrcKind(enum.IntEnum):
    Other = 0
    StaticTile = 1
    DynamicTile = 2
    Pixel = 3
    EntityTile = 4
    ItemTile = 5
    Terrain = 6


@dataclass(frozen=True)
class PixelSource:
    """src/map/pixelsource.rs:23-37.  Use the constructors: PixelSource.Off, .StaticTileIndex(i) ..."""

    kind: int = _SrcKind.Other
    index: int = 0
    pixel: tuple = (0, 0, 0, 0)
    name: str = "Off"
    ident: object = None  # the Uuid of EntityTile / ItemTile

    @staticmethod
    def StaticTileIndex(i: int) -> "PixelSource":
        return PixelSource(_SrcKind.StaticTile, int(i), (0, 0, 0, 0), "StaticTileIndex")

    @staticmethod
    def DynamicTileIndex(i: int) -> "PixelSource":
        return PixelSource(_SrcKind.DynamicTile, int(i), (0, 0, 0, 0), "DynamicTileIndex")

    @staticmethod
    def Pixel(rgba: Sequence[int]) -> "PixelSource":
        return PixelSource(_SrcKind.Pixel, 0, tuple(int(c) for c in rgba), "Pixel")

    @staticmethod
    def Color(rgba: Sequence[int]) -> "PixelSource":
        # TheColor: the rasterizer's non-overlay paths do not read it (src/rasterizer.rs:1221, :757)
        return PixelSource(_SrcKind.Other, 0, tuple(int(c) for c in rgba), "Color")

    @staticmethod
    def EntityTile(id_: int, index: int) -> "PixelSource":
        return PixelSource(_SrcKind.EntityTile, int(index), (0, 0, 0, 0), "EntityTile", id_)

    @staticmethod
    def ItemTile(id_: int, index: int) -> "PixelSource":
        return PixelSource(_SrcKind.ItemTile, int(index), (0, 0, 0, 0), "ItemTile", id_)


PixelSource.Off = PixelSource()
PixelSource.Terrain = PixelSource(_SrcKind.Terrain, 0, (0, 0, 0, 0), "Terrain")


class Texture:
    """src/texture.rs:46-54: RGBA8 row-major."""

    def __init__(self, data, width: int, height: int):
        self.data = np.ascontiguousarray(np.asarray(data, dtype=np.uint8).reshape(-1))
        self.width = int(width)
        self.height = int(height)
        if self.data.size != self.width * self.height * 4:
            raise ValueError("Texture data must be width*height*4 bytes")

    @staticmethod
    def from_array(rgba: np.ndarray) -> "Texture":
        rgba = np.asarray(rgba, dtype=np.uint8)
        h, w, c = rgba.shape
        assert c == 4
        return Texture(rgba.reshape(-1), w, h)

    @staticmethod
    def from_image(path) -> "Texture":  # src/texture.rs:150-166 (decode -> RGBA8)
        from PIL import Image

        im = Image.open(path).convert("RGBA")
        return Texture.from_array(np.asarray(im))

    @staticmethod
    def white() -> "Texture":
        return Texture(np.full(4 * 100 * 100, 255, np.uint8), 100, 100)

    def as_array(self) -> np.ndarray:
        return self.data.reshape(self.height, self.width, 4)


class Tile:
    """src/map/tile.rs:83-110: animation frames."""

    def __init__(self, textures: List[Texture]):
        self.textures = list(textures)

    @staticmethod
    def from_texture(texture: Texture) -> "Tile":
        return Tile([texture])

    @staticmethod
    def from_textures(textures: List[Texture]) -> "Tile":
        return Tile(textures)


class Assets:
    """Only the member the rasterizer reads for static tiles: assets.tile_list (src/server/assets.rs:19)."""

    def __init__(self):
        self.tile_list: List[Tile] = []
        # src/server/assets.rs:28,34: id -> IndexMap<String, Tile>; here id -> list of (name, Tile)
        self.entity_tiles: dict = {}
        self.item_tiles: dict = {}
        self.palette: list = []   # assets.palette.colors (src/server/assets.rs:40): None or (r, g, b) in 0..1
        self._generation = 0
        self._uid = next(_UIDS)  # device-cache identity (id() can be reused after garbage collection)

    @staticmethod
    def default() -> "Assets":
        return Assets()

    def textures(self, tiles: List[Tile]) -> "Assets":  # builder, src/server/assets.rs:288-291
        self.tile_list = list(tiles)
        self._generation += 1
        return self

    def mark_dirty(self):
        self._generation += 1

    def actor_tile(self, src: "PixelSource") -> Optional[Tile]:
        """assets.entity_tiles.get(&id).get_index(index) (src/rasterizer.rs:1130-1152)."""
        table = self.entity_tiles if src.kind == _SrcKind.EntityTile else self.item_tiles
        seq = table.get(src.ident)
        if seq is None or not (0 <= src.index < len(seq)):
            return None
        return seq[src.index][1]


@dataclass
class CompiledLight:
    """src/map/light.rs:457-477; defaults are those of Light::compile (:128-193)."""

    light_type: LightType = LightType.Point
    position: Sequence[float] = (0.0, 0.0, 0.0)
    color: Sequence[float] = (1.0, 1.0, 1.0)
    intensity: float = 1.0
    emitting: bool = True
    start_distance: float = 1.0
    end_distance: float = 2.0
    flicker: float = 0.0
    direction: Sequence[float] = (0.0, 0.0, -1.0)
    cone_angle: float = math.pi / 4
    normal: Sequence[float] = (0.0, 1.0, 0.0)
    width: float = 1.0
    height: float = 1.0
    from_linedef: bool = False


class Light:
    """Builder subset of src/map/light.rs Light (new / with_* / compile)."""

    def __init__(self, light_type: LightType):
        self._c = CompiledLight(light_type=light_type)

    @staticmethod
    def new(light_type: LightType) -> "Light":
        return Light(light_type)

    def with_intensity(self, v):
        self._c.intensity = float(v)
        return self

    def with_color(self, c):
        self._c.color = tuple(float(x) for x in c)
        return self

    def with_position(self, p):
        self._c.position = tuple(float(x) for x in p)
        return self

    def with_start_distance(self, v):
        self._c.start_distance = float(v)
        return self

    def with_end_distance(self, v):
        self._c.end_distance = float(v)
        return self

    def with_flicker(self, v):
        self._c.flicker = float(v)
        return self

    def with_direction(self, d):
        d = np.asarray(d, dtype=np.float32)
        self._c.direction = tuple((d / np.float32(np.sqrt(np.dot(d, d)))).tolist())  # .normalized() (:149-155)
        return self

    def with_cone_angle(self, v):
        self._c.cone_angle = float(v)
        return self

    def with_normal(self, n):
        n = np.asarray(n, dtype=np.float32)
        self._c.normal = tuple((n / np.float32(np.sqrt(np.dot(n, n)))).tolist())
        return self

    def with_size(self, w, h):
        self._c.width, self._c.height = float(w), float(h)
        return self

    def with_emitting(self, e):
        self._c.emitting = bool(e)
        return self

    def compile(self) -> CompiledLight:
        return self._c


def _as_indices(indices) -> np.ndarray:
    a = np.asarray(indices)
    if a.size == 0:
        return np.zeros((0, 3), dtype=np.uint32)
    return np.ascontiguousarray(a.reshape(-1, 3).astype(np.uint32))


class Batch3D:
    """src/batch/batch3d.rs:15-78."""

    def __init__(self, vertices, indices, uvs):
        self.mode = PrimitiveMode.Triangles
        self.vertices = np.ascontiguousarray(np.asarray(vertices, dtype=np.float32).reshape(-1, 4))
        self.indices = _as_indices(indices)
        self.uvs = np.ascontiguousarray(np.asarray(uvs, dtype=np.float32).reshape(-1, 2))
        self.repeat_mode_ = RepeatMode.ClampXY
        self.cull_mode_ = CullMode.Off
        self.source_ = PixelSource.Off
        self.transform_3d = vekmath.identity4()
        self.receives_light_ = True
        self.normals = np.zeros((0, 3), dtype=np.float32)
        self.ambient_color_ = (0.0, 0.0, 0.0)
        self.shader_: Optional[int] = None
        self.profile_id_: Optional[int] = None
        self.pass_ = 0

    @staticmethod
    def empty() -> "Batch3D":
        return Batch3D(np.zeros((0, 4)), np.zeros((0, 3)), np.zeros((0, 2)))

    @staticmethod
    def new(vertices, indices, uvs) -> "Batch3D":
        return Batch3D(vertices, indices, uvs)

    @staticmethod
    def from_box(x, y, z, width, height, depth) -> "Batch3D":
        """Six faces, 24 vertices, 12 triangles; layout of src/batch/batch3d.rs:140-229."""
        x0, x1, y0, y1, z0, z1 = x, x + width, y, y + height, z, z + depth
        faces = [
            [(x0, y0, z0), (x1, y0, z0), (x1, y1, z0), (x0, y1, z0)],  # front
            [(x0, y0, z1), (x1, y0, z1), (x1, y1, z1), (x0, y1, z1)],  # back
            [(x0, y0, z0), (x0, y1, z0), (x0, y1, z1), (x0, y0, z1)],  # left
            [(x1, y0, z0), (x1, y1, z0), (x1, y1, z1), (x1, y0, z1)],  # right
            [(x0, y1, z0), (x1, y1, z0), (x1, y1, z1), (x0, y1, z1)],  # top
            [(x0, y0, z0), (x1, y0, z0), (x1, y0, z1), (x0, y0, z1)],  # bottom
        ]
        # per-face (second, third) corner of its two triangles, both starting at corner 0
        winding = [((1, 2), (2, 3)), ((2, 1), (3, 2)), ((1, 2), (2, 3)), ((2, 1), (3, 2)), ((1, 2), (2, 3)), ((3, 2), (2, 1))]
        verts, idx, uvs = [], [], []
        for fi, face in enumerate(faces):
            base = fi * 4
            verts += [(*p, 1.0) for p in face]
            uvs += [(0.0, 1.0), (1.0, 1.0), (1.0, 0.0), (0.0, 0.0)]
            for a, b in winding[fi]:
                idx.append((base, base + a, base + b))
        return Batch3D(verts, idx, uvs)

    @staticmethod
    def from_obj(path_or_text) -> "Batch3D":  # src/batch/batch3d.rs:407-419
        from .wavefront import Wavefront

        return Wavefront.parse(path_or_text).to_batch()

    # builder methods, src/batch/batch3d.rs:421-479
    def mode_(self, mode):
        self.mode = PrimitiveMode(mode)
        return self

    def repeat_mode(self, m):
        self.repeat_mode_ = RepeatMode(m)
        return self

    def cull_mode(self, m):
        self.cull_mode_ = CullMode(m)
        return self

    def source(self, s: PixelSource):
        self.source_ = s
        return self

    def shader(self, i: int):
        self.shader_ = int(i)
        return self

    def ambient_color(self, c):
        self.ambient_color_ = tuple(float(x) for x in c)
        return self

    def transform(self, m):
        self.transform_3d = np.asarray(m, dtype=np.float32).reshape(4, 4)
        return self

    def receives_light(self, v: bool):
        self.receives_light_ = bool(v)
        return self

    def profile_id(self, v: int):
        self.profile_id_ = int(v)
        return self

    def with_normals(self, normals):
        self.normals = np.ascontiguousarray(np.asarray(normals, dtype=np.float32).reshape(-1, 3))
        return self

    def with_computed_normals(self) -> "Batch3D":
        """Smooth vertex normals, src/batch/batch3d.rs:812-842 (float32, sequential accumulation)."""
        n = np.zeros((len(self.vertices), 3), dtype=np.float32)
        counts = np.zeros(len(self.vertices), dtype=np.uint32)
        p = self.vertices[:, :3]
        for i0, i1, i2 in self.indices.tolist():
            e1 = p[i1] - p[i0]
            e2 = p[i2] - p[i0]
            c = np.array(
                [e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]],
                dtype=np.float32,
            )
            with np.errstate(invalid="ignore", divide="ignore"):
                c = c / np.float32(np.sqrt(np.float32(c[0] * c[0] + c[1] * c[1] + c[2] * c[2])))
            n[i0] += c
            n[i1] += c
            n[i2] += c
            counts[i0] += 1
            counts[i1] += 1
            counts[i2] += 1
        for i in range(len(n)):
            if counts[i] > 0:
                v = n[i] / np.float32(counts[i])
                with np.errstate(invalid="ignore", divide="ignore"):
                    n[i] = v / np.float32(np.sqrt(np.float32(v[0] * v[0] + v[1] * v[1] + v[2] * v[2])))
        self.normals = n
        return self

    def with_computed_normals_fast(self) -> "Batch3D":
        """Vectorised variant for the large synthetic scenes (same formula; accumulation order of
        np.add.at instead of the sequential loop).  Normals are inputs to the path, not results."""
        p = self.vertices[:, :3]
        i0, i1, i2 = self.indices[:, 0], self.indices[:, 1], self.indices[:, 2]
        c = np.cross(p[i1] - p[i0], p[i2] - p[i0]).astype(np.float32)
        with np.errstate(invalid="ignore", divide="ignore"):
            c = (c / np.sqrt((c * c).sum(axis=1, dtype=np.float32))[:, None]).astype(np.float32)
        n = np.zeros((len(p), 3), dtype=np.float32)
        cnt = np.zeros(len(p), dtype=np.float32)
        for ix in (i0, i1, i2):
            np.add.at(n, ix, c)
            np.add.at(cnt, ix, 1.0)
        m = cnt > 0
        n[m] = n[m] / cnt[m, None]
        with np.errstate(invalid="ignore", divide="ignore"):
            n[m] = n[m] / np.sqrt((n[m] * n[m]).sum(axis=1, dtype=np.float32))[:, None]
        self.normals = np.ascontiguousarray(n.astype(np.float32))
        return self


class Batch2D:
    """src/batch/batch2d.rs:10-53."""

    def __init__(self, vertices, indices, uvs):
        self.mode = PrimitiveMode.Triangles
        self.vertices = np.ascontiguousarray(np.asarray(vertices, dtype=np.float32).reshape(-1, 2))
        self.indices = _as_indices(indices)
        self.uvs = np.ascontiguousarray(np.asarray(uvs, dtype=np.float32).reshape(-1, 2))
        self.repeat_mode_ = RepeatMode.ClampXY
        self.source_ = PixelSource.Off
        self.receives_light_ = True
        self.shader_: Optional[int] = None

    @staticmethod
    def empty() -> "Batch2D":
        return Batch2D(np.zeros((0, 2)), np.zeros((0, 3)), np.zeros((0, 2)))

    @staticmethod
    def new(vertices, indices, uvs) -> "Batch2D":
        return Batch2D(vertices, indices, uvs)

    @staticmethod
    def from_rectangle(x, y, width, height) -> "Batch2D":  # src/batch/batch2d.rs:109-127
        v = [(x, y), (x, y + height), (x + width, y + height), (x + width, y)]
        return Batch2D(v, [(0, 1, 2), (0, 2, 3)], [(0.0, 0.0), (0.0, 1.0), (1.0, 1.0), (1.0, 0.0)])

    def add_rectangle(self, x, y, width, height):
        base = len(self.vertices)
        v = np.array([(x, y), (x, y + height), (x + width, y + height), (x + width, y)], dtype=np.float32)
        self.vertices = np.ascontiguousarray(np.concatenate([self.vertices, v]))
        self.uvs = np.ascontiguousarray(
            np.concatenate([self.uvs, np.array([(0, 0), (0, 1), (1, 1), (1, 0)], dtype=np.float32)])
        )
        self.indices = np.ascontiguousarray(
            np.concatenate([self.indices, np.array([(base, base + 1, base + 2), (base, base + 2, base + 3)], np.uint32)])
        )
        return self

    def mode_(self, mode):
        self.mode = PrimitiveMode(mode)
        return self

    def repeat_mode(self, m):
        self.repeat_mode_ = RepeatMode(m)
        return self

    def source(self, s: PixelSource):
        self.source_ = s
        return self

    def shader(self, i: int):
        self.shader_ = int(i)
        return self

    def receives_light(self, v: bool):
        self.receives_light_ = bool(v)
        return self


class Scene:
    """src/scene.rs:8-50 (the members the rasterize path reads)."""

    def __init__(self):
        self.background: Optional[object] = None  # a shader from rusterix_b200.shader
        self.lights: List[CompiledLight] = []
        self.dynamic_lights: List[CompiledLight] = []
        self.d3_static: List[Batch3D] = []
        self.d3_dynamic: List[Batch3D] = []
        self.d3_overlay: List[Batch3D] = []
        self.d2_static: List[Batch2D] = []
        self.d2_dynamic: List[Batch2D] = []
        self.dynamic_textures: List[Tile] = []
        self.chunks: dict = {}  # (x, y) -> Chunk, iterated in insertion order (src/scene.rs:46)
        self.shaders: list = []            # src/scene.rs:42-43: rusterix_b200.vm.Program per add_shader
        self.shaders_with_opacity: List[bool] = []
        # the VM's pattern banks (rusteria/src/textures/patterns.rs:88-102) as (width, height, float32[h*w,3]); the
        # reference computes them once per process, the host passes them along with the scene
        self.patterns: list = []
        self.patterns_normal: list = []
        self.animation_frame = 0
        self._generation = 0
        self._uid = next(_UIDS)  # device-cache identity (id() can be reused after garbage collection)

    @staticmethod
    def empty() -> "Scene":
        return Scene()

    default = empty

    @staticmethod
    def from_static(d2: List[Batch2D], d3: List[Batch3D]) -> "Scene":  # src/scene.rs:83-91
        s = Scene()
        s.d2_static = list(d2)
        s.d3_static = list(d3)
        return s

    def background_(self, shader) -> "Scene":
        self.background = shader
        return self

    def lights_(self, lights: List[CompiledLight]) -> "Scene":
        self.lights = list(lights)
        return self

    def add_shader(self, program) -> Optional[int]:
        """src/scene.rs:108-133.  The reference compiles Rusteria source; the compiler stays on the host, so this
        mirror takes its output, a rusterix_b200.vm.Program."""
        if program is None:
            return None
        index = len(self.shaders)
        self.shaders_with_opacity.append(program.shader_supports_opacity())
        self.shaders.append(program)
        self._generation += 1
        return index

    def anim_tick(self):  # src/scene.rs:147-150
        self.animation_frame = (self.animation_frame + 1) & 0xFFFFFFFFFFFFFFFF

    def mark_dirty(self):
        """Tell the device cache that geometry / dynamic textures changed IN PLACE (a vertex array edited, a texture
        repainted).  Structural changes -- batches or tiles appended, removed or replaced, the usual per-frame
        `scene.d3_dynamic = ...` of the reference's callers -- are seen without it (structure_key)."""
        self._generation += 1

    def structure_key(self):
        """Cheap identity of what the scene currently holds: which batch objects (and which vertex / index arrays
        inside them) sit in which list, how many dynamic tiles and shaders there are.  The reference re-projects
        `&mut scene` on every rasterize() call; the device cache is re-uploaded when this key or the generation
        changes."""
        def bkey(b):
            return (id(b), id(b.vertices), id(b.indices), len(b.vertices), len(b.indices), id(getattr(b, "source_", None)))
        key = [tuple(bkey(b) for b in lst) for lst in (self.d2_static, self.d2_dynamic, self.d3_static, self.d3_dynamic, self.d3_overlay)]
        for ck, ch in self.chunks.items():
            key.append((ck, id(ch), tuple(bkey(b) for b in ch.batches2d), tuple(bkey(b) for b in ch.batches3d_opacity),
                        tuple(bkey(b) for b in ch.batches3d), id(ch.terrain_batch2d), id(ch.terrain_batch3d), id(ch.terrain_texture),
                        len(ch.occluded_sectors), len(ch.shaders)))
        key.append((len(self.dynamic_textures), tuple(id(t) for t in self.dynamic_textures), len(self.shaders)))
        return tuple(key)

    def all_lights(self) -> List[CompiledLight]:
        return list(self.lights) + list(self.dynamic_lights)


@dataclass
@dataclass
class BBox:  # src/map/bbox.rs:5-40
    min: tuple = (0.0, 0.0)
    max: tuple = (0.0, 0.0)

    @staticmethod
    def from_pos_size(pos, size) -> "BBox":
        return BBox((float(pos[0]), float(pos[1])), (float(pos[0]) + float(size[0]), float(pos[1]) + float(size[1])))


@dataclass
class CompiledLinedef:  # src/map/mini.rs: start / end of a wall segment
    start: tuple = (0.0, 0.0)
    end: tuple = (0.0, 0.0)


class MapMini:
    """src/map/mini.rs:22-56: what Rasterizer.mapmini contributes to the path -- linedefs for the 2D
    line-of-sight test (:88-95) and occluded sectors (:58-65)."""

    def __init__(self, linedefs=None, occluded_sectors=None):
        self.linedefs: List[CompiledLinedef] = list(linedefs or [])
        self.occluded_sectors: List[tuple] = list(occluded_sectors or [])  # (BBox, occlusion)

    @staticmethod
    def default() -> "MapMini":
        return MapMini()

    def key(self):
        return (tuple((l.start, l.end) for l in self.linedefs), tuple((b.min, b.max, o) for b, o in self.occluded_sectors))


class Chunk:
    """src/chunk.rs:23-57: the members on the rasterize path."""

    def __init__(self, origin=(0, 0), size=0):
        self.origin = (int(origin[0]), int(origin[1]))
        self.size = int(size)
        self.bbox = BBox.from_pos_size(self.origin, (size, size))
        self.batches2d: List[Batch2D] = []
        self.batches3d_opacity: List[Batch3D] = []
        self.batches3d: List[Batch3D] = []
        self.terrain_batch2d: Optional[Batch2D] = None
        self.terrain_batch3d: Optional[Batch3D] = None
        self.terrain_texture: Optional[Texture] = None
        self.lights: List[CompiledLight] = []
        self.occluded_sectors: List[tuple] = []  # (BBox, occlusion)
        self.shaders: list = []                  # src/chunk.rs: programs batch.shader indexes for this chunk's batches
        self.shader_textures: list = []          # Option<Texture> per shader: a baked result used instead of the program

    @staticmethod
    def new(origin, size) -> "Chunk":
        return Chunk(origin, size)


class VGrayGradientShader:  # src/shader/vgradient.rs:4-15
    kind: int = 1


@dataclass
class GridShader:  # src/shader/grid.rs:4-35
    kind: int = 2
    grid_size: float = 30.0
    subdivisions: float = 2.0
    offset: Sequence[float] = (0.0, 0.0)

    def set_parameter_f32(self, key, value):
        if key == "grid_size":
            self.grid_size = float(value)
        elif key == "subdivisions":
            self.subdivisions = float(value)

    def set_parameter_vec2(self, key, value):
        if key == "offset":
            self.offset = (float(value[0]), float(value[1]))


@dataclass
class RenderMode:  # src/rendermode.rs:3-52
    d2_active: bool = True
    d3_active: bool = True
    ignore_background_shader_: bool = False

    @staticmethod
    def render_all():
        return RenderMode(True, True, False)

    @staticmethod
    def render_2d():
        return RenderMode(True, False, False)

    @staticmethod
    def render_3d():
        return RenderMode(False, True, False)

    def ignore_background_shader(self, v: bool):
        self.ignore_background_shader_ = bool(v)
        return self


# ------------------------------------------------------------------------------------------------
# render graph: what Rasterizer::rasterize reads of it (src/rasterizer.rs:227-253, :419-461)
# ------------------------------------------------------------------------------------------------
def _f32(x):
    return np.float32(x)


@dataclass
class BrushPreview:  # src/rasterizer.rs:13-17
    position: Sequence[float] = (0.0, 0.0, 0.0)
    radius: float = 1.0
    falloff: float = 0.5


class SkyNode:
    """ShapeFX role Sky (src/shapestack/shapefx.rs): render_setup (:970-1057), render_ambient_color (:1086-1120).
    The per-pixel part, render_miss_d3 (:1122-1223), runs on the device.  `clouds=True` is the reference's
    behaviour (a noiselib Perlin cloud layer above the horizon), which the device path does not have."""

    def __init__(self, day_horizon=(0.87, 0.80, 0.70, 1.0), day_zenith=(0.36, 0.62, 0.98, 1.0),
                 night_horizon=(0.03, 0.04, 0.08, 1.0), night_zenith=(0.00, 0.01, 0.05, 1.0), clouds=True):
        self.day_horizon, self.day_zenith = tuple(day_horizon), tuple(day_zenith)
        self.night_horizon, self.night_zenith = tuple(night_horizon), tuple(night_zenith)
        self.clouds = bool(clouds)
        self.precomputed: list = []

    def render_setup(self, hour: float):
        h = _f32(hour)

        def smooth(x):   # ((x).clamp(0,2)/2).powi(2) * (3 - 2*(x).clamp(0,2)/2)
            c = _f32(min(max(x, _f32(0.0)), _f32(2.0)))
            return _f32(_f32(_f32(c / _f32(2.0)) * _f32(c / _f32(2.0))) * _f32(_f32(3.0) - _f32(_f32(_f32(2.0) * c) / _f32(2.0))))
        dawn, dusk = smooth(_f32(h - _f32(6.0))), smooth(_f32(_f32(20.0) - h))
        day_factor = _f32(0.0) if h < 6.0 else dawn if h < 8.0 else _f32(1.0) if h < 18.0 else dusk if h < 20.0 else _f32(0.0)
        t_day = _f32(min(max(_f32(_f32(h - _f32(6.0)) / _f32(14.0)), _f32(0.0)), _f32(1.0)))
        theta = _f32(t_day * _f32(np.pi))
        sun_dir = (float(_f32(np.cos(theta))), float(_f32(np.sin(theta))), 0.0)
        night, day = np.float32([0.1, 0.1, 0.15, 0.0]), np.float32([0.3, 0.3, 0.35, 0.0])
        tcl = _f32(min(max(day_factor, _f32(0.0)), _f32(1.0)))
        haze = [float(_f32(np.float64(tcl) * np.float64(_f32(d - n)) + np.float64(n))) for n, d in zip(night, day)]  # mul_add
        self.precomputed = [(sun_dir[0], sun_dir[1], sun_dir[2], float(day_factor)), tuple(haze), self.day_horizon, self.day_zenith,
                            self.night_horizon, self.night_zenith]
        return sun_dir, float(day_factor)

    def render_ambient_color(self):
        df = _f32(self.precomputed[0][3])
        tcl = _f32(min(max(df, _f32(0.0)), _f32(1.0)))
        out = []
        for i in range(3):
            day_avg = _f32(_f32(_f32(self.day_horizon[i]) * _f32(0.5)) + _f32(_f32(self.day_zenith[i]) * _f32(0.5)))
            night_avg = _f32(_f32(_f32(self.night_horizon[i]) * _f32(0.5)) + _f32(_f32(self.night_zenith[i]) * _f32(0.5)))
            c = _f32(np.float64(tcl) * np.float64(_f32(day_avg - night_avg)) + np.float64(night_avg))
            c = max(c, _f32(0.2))
            out.append(float(_f32(c * _f32(12.92)) if c <= 0.0031308 else _f32(_f32(_f32(1.055) * _f32(np.power(c, _f32(1.0 / 2.4)))) - _f32(0.055))))
        return (out[0], out[1], out[2], 1.0)


class RenderGraph:
    """The slice of ShapeFXGraph the rasterizer walks: the nodes reachable from the render node's miss terminal.
    Only Sky nodes act in render_miss_d3; Fog's render_hit_d3 is never called by rasterize() (the hit list is only
    set up, src/rasterizer.rs:227-233)."""

    def __init__(self, miss_nodes=None):
        self.miss_nodes: list = list(miss_nodes or [])
