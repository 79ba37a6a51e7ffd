"""Deterministic synthetic scenes for BASELINE.json's five configs.

No file of the reference is read: textures are procedural with the dimensions of the reference's
assets (logo 1024x1024, brick* 64x64, fence 64x80 with alpha holes, sky 256x128), the teapot is a
lathe mesh with the OBJ's triangle count (2256), and the map is a restatement of minigame/world.rxm
with exact integer coordinates.  Each builder returns a `Config` (scene, assets, camera set-up,
resolution, tile size, sampling) ready for `Rasterizer.setup(..).rasterize(..)`."""
from __future__ import annotations

import math
import os
from dataclasses import dataclass, field
from typing import Callable, List, Optional

import numpy as np

from .camera import D3FirstPCamera, D3OrbitCamera
from .rasterizer import Rasterizer
from .types import (Assets, BBox, Batch2D, Batch3D, Chunk, CompiledLinedef, CullMode, Light, LightType, MapMini,
                    PixelSource, PrimitiveMode, RenderMode, RepeatMode, SampleMode, Scene, Texture, Tile, VGrayGradientShader)
from . import vekmath

SEED = 0x52555354
# the reference's own input fixtures (mesh, textures), vendored under tests/golden/reference_assets with their provenance
REFERENCE_ASSETS = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "reference_assets")


def reference_asset(name: str) -> str:
    p = os.path.join(REFERENCE_ASSETS, name)
    if not os.path.exists(p):
        raise FileNotFoundError(f"{p}: the vendored reference fixtures are missing")
    return p


def have_reference_assets() -> bool:
    return os.path.exists(os.path.join(REFERENCE_ASSETS, "teapot.obj"))


def splitmix64(n: int, seed: int = SEED) -> np.ndarray:
    """n uint64 values of the splitmix64 sequence starting at `seed`."""
    with np.errstate(over="ignore"):
        idx = np.arange(1, n + 1, dtype=np.uint64)
        z = np.uint64(seed) + idx * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def rand01(n: int, seed: int) -> np.ndarray:
    return (splitmix64(n, seed) >> np.uint64(40)).astype(np.float32) / np.float32(1 << 24)


# ------------------------------------------------------------------------------------------------
# procedural textures
# ------------------------------------------------------------------------------------------------
def _rgba(r, g, b, a=None):
    h, w = r.shape
    out = np.empty((h, w, 4), dtype=np.uint8)
    out[..., 0] = np.clip(r, 0, 255)
    out[..., 1] = np.clip(g, 0, 255)
    out[..., 2] = np.clip(b, 0, 255)
    out[..., 3] = 255 if a is None else np.clip(a, 0, 255)
    return out


def tex_logo(size=1024) -> Texture:
    y, x = np.mgrid[0:size, 0:size].astype(np.float32)
    cx = (x - size / 2) / size
    cy = (y - size / 2) / size
    rad = np.sqrt(cx * cx + cy * cy)
    rings = (np.sin(rad * 60.0) * 0.5 + 0.5)
    checker = (((x // 64) + (y // 64)) % 2)
    noise = rand01(size * size, SEED + 1).reshape(size, size)
    r = 40 + 180 * rings + 20 * noise
    g = 60 + 120 * checker + 50 * (y / size)
    b = 90 + 140 * (x / size) + 20 * noise
    return Texture.from_array(_rgba(r, g, b))


def tex_brick(seed, base=(150, 70, 50), mortar=(190, 185, 170), w=64, h=64) -> Texture:
    y, x = np.mgrid[0:h, 0:w]
    row = y // 8
    xx = (x + (row % 2) * 8) % 16
    is_mortar = ((y % 8) == 0) | (xx == 0)
    n = rand01(w * h, seed).reshape(h, w)
    ch = []
    for c in range(3):
        ch.append(np.where(is_mortar, mortar[c], base[c] + 40 * (n - 0.5)))
    return Texture.from_array(_rgba(*ch))


def tex_lightpanel(w=64, h=64) -> Texture:
    y, x = np.mgrid[0:h, 0:w].astype(np.float32)
    d = np.maximum(np.abs(x - w / 2), np.abs(y - h / 2)) / (w / 2)
    v = 255 - 120 * d
    return Texture.from_array(_rgba(v, v, 0.8 * v))


def tex_fence(w=64, h=80) -> Texture:
    """Vertical bars + two rails; everything else alpha 0 (exercises the alpha test, SURVEY T-alpha)."""
    y, x = np.mgrid[0:h, 0:w]
    bar = (x % 16) < 4
    rail = ((y >= 12) & (y < 18)) | ((y >= 60) & (y < 66))
    solid = bar | rail
    n = rand01(w * h, SEED + 7).reshape(h, w)
    v = 90 + 50 * n
    a = np.where(solid, 255, 0)
    return Texture.from_array(_rgba(v, 0.8 * v, 0.5 * v, a))


def tex_sky(w=256, h=128) -> Texture:
    y, x = np.mgrid[0:h, 0:w].astype(np.float32)
    t = y / h
    n = rand01(w * h, SEED + 9).reshape(h, w)
    cloud = (np.sin(x * 0.09) * np.sin(y * 0.17 + 1.3) > 0.55) * 50
    return Texture.from_array(_rgba(70 + 120 * t + cloud, 120 + 90 * t + cloud, 215 + 30 * t + 10 * n))


def tex_checker_noise(i: int, size=64) -> Texture:
    y, x = np.mgrid[0:size, 0:size]
    n = rand01(size * size, SEED + 100 + i).reshape(size, size)
    k = 4 << (i % 3)
    chk = (((x // k) + (y // k)) % 2).astype(np.float32)
    hue = (i * 0.618034) % 1.0
    base = np.array([0.5 + 0.5 * math.cos(2 * math.pi * (hue + o)) for o in (0.0, 0.33, 0.67)]) * 200 + 30
    ch = [base[c] * (0.55 + 0.45 * chk) + 30 * (n - 0.5) for c in range(3)]
    return Texture.from_array(_rgba(*ch))


# ------------------------------------------------------------------------------------------------
@dataclass
class Config:
    name: str
    scene: Scene
    assets: Assets
    width: int
    height: int
    tile_size: int
    sample_mode: SampleMode
    ambient: Optional[tuple]
    camera: object
    cameras: Optional[Callable[[int], object]] = None  # frame index -> camera (sweeps)
    n_frames: int = 1
    notes: str = ""
    matrix2d: Optional[object] = None       # projection_matrix_2d (Mat3) of the 2D configs
    mapmini: Optional[MapMini] = None
    render_mode: Optional[RenderMode] = None
    time: float = 0.0                       # Rasterizer.time (the VM's `time`)
    render_graph: Optional[object] = None   # types.RenderGraph
    hour: float = 12.0
    brush_preview: Optional[object] = None  # types.BrushPreview

    def rasterizer(self, frame: int = 0) -> Rasterizer:
        cam = self.cameras(frame) if self.cameras is not None else self.camera
        r = Rasterizer.setup(self.matrix2d, cam.view_matrix(), cam.projection_matrix(float(self.width), float(self.height)))
        r.sample_mode(self.sample_mode)
        if self.ambient is not None:
            r.ambient(self.ambient)
        if self.mapmini is not None:
            r.mapmini = self.mapmini
        if self.render_mode is not None:
            r.render_mode(self.render_mode)
        r.time_ = float(self.time)
        r.render_graph, r.hour, r.brush_preview = self.render_graph, float(self.hour), self.brush_preview
        return r

    def counts(self):
        from .marshal import submission_order
        b3, _ = submission_order(self.scene)
        v = sum(len(b.vertices) for b, _p, _c in b3)
        t = sum(len(b.indices) for b, _p, _c in b3)
        return v, t

    def algorithmic_bytes(self) -> int:
        """B_alg of SURVEY.md section 8d / BASELINE.md: W*H*4 + V*36 + T*12 + referenced textures + L*72 + 256."""
        v, t = self.counts()
        used = set()
        for b in self.scene.d3_static + self.scene.d3_dynamic + self.scene.d3_overlay + self.scene.d2_static + self.scene.d2_dynamic:
            if b.source_.kind in (1, 2):
                used.add((b.source_.kind, b.source_.index))
        tex = 0
        for kind, idx in used:
            tiles = self.assets.tile_list if kind == 1 else self.scene.dynamic_textures
            if idx < len(tiles):
                tx = tiles[idx].textures[self.scene.animation_frame % len(tiles[idx].textures)]
                tex += tx.width * tx.height * 4
        return self.width * self.height * 4 + v * 36 + t * 12 + tex + len(self.scene.all_lights()) * 72 + 256


def _example_light():
    # examples/cube.rs:45-49,72-73 at t = 0: position (2, 0.8, 0)
    return Light.new(LightType.Point).with_intensity(1.0).with_color([1.0, 1.0, 0.95]).with_position([2.0, 0.8, 0.0]).compile()


def cube(width=800, height=600, tile_size=200, logo_size=1024, real=False) -> Config:
    """Config A (examples/cube.rs:27-90): textured box, CullMode::Off, 1 point light, 200x200 2D rect.
    real=True: the texture is the reference's images/logo.png (examples/cube.rs:36) instead of the procedural logo."""
    scene = Scene.from_static(
        [Batch2D.from_rectangle(0.0, 0.0, 200.0, 200.0)],
        [Batch3D.from_box(-0.5, -0.5, -0.5, 1.0, 1.0, 1.0).source(PixelSource.StaticTileIndex(0)).cull_mode(CullMode.Off).with_computed_normals()],
    ).lights_([_example_light()]).background_(VGrayGradientShader())
    assets = Assets.default().textures([Tile.from_texture(Texture.from_image(reference_asset("logo.png")) if real else tex_logo(logo_size))])
    cam = D3OrbitCamera.new()
    cam.set_parameter_f32("distance", 1.5)
    return Config("cube", scene, assets, width, height, tile_size, SampleMode.Nearest, (0.1,) * 4, cam)


def lathe_teapot(segs=47, bands=24) -> Batch3D:
    """A teapot-like lathe body with the triangle count of examples/teapot.obj (2*47*24 = 2256)."""
    prof = []
    for i in range(bands + 1):
        t = i / bands
        y = 3.15 * t
        r = 1.4 + 0.6 * math.sin(math.pi * min(1.0, t * 1.15)) - 1.1 * max(0.0, t - 0.8) * 5 * (t - 0.8) * 5 * 0.2
        r = max(0.05, r * (1.0 if t < 0.86 else (1.0 - (t - 0.86) / 0.14) * 0.9 + 0.1))
        prof.append((r, y))
    verts, idx = [], []
    for i, (r, y) in enumerate(prof):
        for s in range(segs):
            a = 2 * math.pi * s / segs
            verts.append((r * math.cos(a), y, r * math.sin(a), 1.0))
    for i in range(bands):
        for s in range(segs):
            a = i * segs + s
            b = i * segs + (s + 1) % segs
            c = (i + 1) * segs + s
            d = (i + 1) * segs + (s + 1) % segs
            idx.append((a, c, b))
            idx.append((b, c, d))
    v = np.asarray(verts, dtype=np.float32)
    return Batch3D(v, idx, v[:, :2].copy())  # UV defaults to (x, y): src/wavefront.rs:91-101


def teapot(width=1920, height=1080, tile_size=60, obj_path=None, logo_size=1024, n_frames=64, real=False) -> Config:
    """Config B (examples/obj.rs:28-83): OBJ mesh, RepeatXY, scaling(.35,-.35,.35), Linear, orbit sweep.
    real=True: the reference's examples/teapot.obj (1202 vertices, 2256 triangles) and images/logo.png; otherwise a
    lathe body with the same triangle count and the procedural logo."""
    if real and obj_path is None:
        obj_path = reference_asset("teapot.obj")
    mesh = Batch3D.from_obj(obj_path) if obj_path else lathe_teapot()
    mesh = (mesh.source(PixelSource.StaticTileIndex(0)).repeat_mode(RepeatMode.RepeatXY)
            .transform(vekmath.scaling_3d(0.35, -0.35, 0.35)).with_computed_normals())
    scene = Scene.from_static([Batch2D.from_rectangle(0.0, 0.0, 200.0, 200.0)], [mesh]).lights_([_example_light()]).background_(VGrayGradientShader())
    assets = Assets.default().textures([Tile.from_texture(Texture.from_image(reference_asset("logo.png")) if real else tex_logo(logo_size))])
    cam = D3OrbitCamera.new()
    cam.set_parameter_f32("distance", 1.5)

    def cams(k):
        c = D3OrbitCamera.new()
        c.set_parameter_f32("distance", 1.5)
        c.azimuth = math.pi / 2.0 + 2.0 * math.pi * k / n_frames
        return c

    return Config("teapot", scene, assets, width, height, tile_size, SampleMode.Linear, (0.8,) * 4, cam, cams, n_frames)


def _wall_quad(x0, z0, x1, z1, height, u0=0.0):
    length = math.hypot(x1 - x0, z1 - z0)
    verts = [(x0, 0.0, z0, 1.0), (x1, 0.0, z1, 1.0), (x1, height, z1, 1.0), (x0, height, z0, 1.0)]
    uvs = [(u0, height), (u0 + length, height), (u0 + length, 0.0), (u0, 0.0)]
    return verts, uvs


def _quads_batch(quads):
    verts, uvs, idx = [], [], []
    for qv, quv in quads:
        base = len(verts)
        verts += qv
        uvs += quv
        idx += [(base, base + 1, base + 2), (base, base + 2, base + 3)]
    return Batch3D(verts, idx, uvs)


MAP_TILES = ["logo", "brickwall", "lightpanel", "fence", "brickfloor", "sky"]


def map_assets(logo_size=1024, real=False) -> Assets:
    """Tiles of the map scene in the order map_scene() indexes them: logo, brickwall, lightpanel, fence, brickfloor, sky.
    real=True: the reference's own PNGs (images/logo.png, minigame/*.png) instead of the procedural stand-ins."""
    if real:
        return Assets.default().textures([Tile.from_texture(Texture.from_image(reference_asset(n)))
                                          for n in ("logo.png", "brickwall.png", "lightpanel.png", "fence.png", "brickfloor.png", "sky.png")])
    return Assets.default().textures([
        Tile.from_texture(tex_logo(logo_size)),
        Tile.from_texture(tex_brick(SEED + 2)),
        Tile.from_texture(tex_lightpanel()),
        Tile.from_texture(tex_fence()),
        Tile.from_texture(tex_brick(SEED + 3, base=(120, 110, 100), mortar=(70, 70, 70))),
        Tile.from_texture(tex_sky()),
    ])


def map_scene() -> Scene:
    """Restatement of minigame/world.rxm with exact integer coordinates."""
    H = 2.0
    room = [(0, 0), (15, 0), (15, 15), (10, 15), (9, 15), (0, 15), (0, 0)]
    brick, panel = [], []
    for (x0, z0), (x1, z1) in zip(room[:-1], room[1:]):
        q = _wall_quad(float(x0), float(z0), float(x1), float(z1), H)
        (panel if (x0, z0, x1, z1) == (10, 15, 9, 15) else brick).append(q)
    fence = [_wall_quad(6.0, 15.0, 6.0, 9.0, H), _wall_quad(6.0, 9.0, 0.0, 9.0, H)]
    floor = Batch3D(
        [(0.0, 0.0, 0.0, 1.0), (15.0, 0.0, 0.0, 1.0), (15.0, 0.0, 15.0, 1.0), (0.0, 0.0, 15.0, 1.0)],
        [(0, 1, 2), (0, 2, 3)],
        [(0.0, 0.0), (15.0, 0.0), (15.0, 15.0), (0.0, 15.0)],
    )
    sky = Batch3D.from_box(-32.5, -20.0, -32.5, 80.0, 60.0, 80.0)

    def fin(b, tile):
        return b.source(PixelSource.StaticTileIndex(tile)).repeat_mode(RepeatMode.RepeatXY).cull_mode(CullMode.Off).with_computed_normals()

    d3 = [fin(_quads_batch(brick), 1), fin(_quads_batch(panel), 2), fin(_quads_batch(fence), 3), fin(floor, 4),
          fin(sky, 5).receives_light(False)]
    logo = Batch2D.from_rectangle(0.0, 0.0, 200.0, 200.0).receives_light(False).source(PixelSource.StaticTileIndex(0))
    light = (Light.new(LightType.Point).with_color([1.0, 1.0, 0xBB / 255.0]).with_intensity(2.0)
             .with_start_distance(2.0).with_end_distance(13.0).with_position([9.0, 0.5, 15.0]).compile())
    return Scene.from_static([logo], d3).lights_([light])


def _firstp(pos, center):
    c = D3FirstPCamera.new()
    c.position = np.asarray(pos, dtype=np.float32)
    c.center = np.asarray(center, dtype=np.float32)
    return c


def map_config(width=3840, height=2160, tile_size=40, logo_size=1024, real=False) -> Config:
    """Config C (examples/map.rs:62-125): first-person camera inside the room, Nearest, ambient 1."""
    pos = np.array([6.0600824, 1.0, 4.5524735], dtype=np.float32)
    cam = _firstp(pos, pos + np.array([0.03489969, 0.0, 0.99939084], dtype=np.float32))
    return Config("map", map_scene(), map_assets(logo_size, real), width, height, tile_size, SampleMode.Nearest, (1.0,) * 4, cam)


def sweep(width=1920, height=1080, tile_size=40, n_frames=4096, logo_size=1024, real=False) -> Config:
    """Config E: n_frames views of the map scene on a circle of radius 4 around the room centre."""
    def cams(i):
        th = 2.0 * math.pi * i / n_frames
        return _firstp([7.5 + 4.0 * math.cos(th), 1.0, 7.5 + 4.0 * math.sin(th)], [7.5, 1.0, 7.5])

    return Config("sweep", map_scene(), map_assets(logo_size, real), width, height, tile_size, SampleMode.Nearest, (1.0,) * 4,
                  cams(0), cams, n_frames)


def _fbm(x, z, seed):
    def hash2(ix, iz):
        with np.errstate(over="ignore"):
            h = (ix.astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15)) ^ (iz.astype(np.uint64) * np.uint64(0xC2B2AE3D27D4EB4F)) ^ np.uint64(seed)
            h = (h ^ (h >> np.uint64(29))) * np.uint64(0xBF58476D1CE4E5B9)
            h = h ^ (h >> np.uint64(32))
        return (h >> np.uint64(40)).astype(np.float32) / np.float32(1 << 24)

    total = np.zeros_like(x, dtype=np.float32)
    amp, freq = 1.0, 1.0 / 8.0
    for _ in range(4):
        fx, fz = x * freq, z * freq
        ix, iz = np.floor(fx).astype(np.int64), np.floor(fz).astype(np.int64)
        tx, tz = (fx - ix).astype(np.float32), (fz - iz).astype(np.float32)
        tx, tz = tx * tx * (3 - 2 * tx), tz * tz * (3 - 2 * tz)
        a, b = hash2(ix, iz), hash2(ix + 1, iz)
        c, d = hash2(ix, iz + 1), hash2(ix + 1, iz + 1)
        total += amp * ((a * (1 - tx) + b * tx) * (1 - tz) + (c * (1 - tx) + d * tx) * tz)
        amp *= 0.5
        freq *= 2.0
    return total / 1.875


def dense(width=7680, height=4320, tile_size=40, patches=32, patch_verts=23) -> Config:
    """Config D: patches^2 batches tiling a 64x64-unit heightfield, (patch_verts-1)^2*2 triangles each
    (defaults: 1024 batches, 991,232 triangles, 541,696 vertices), 16 textures, 4 point + 2 spot + 1 area light."""
    extent = 64.0
    psize = extent / patches
    q = patch_verts - 1
    jj, ii = np.meshgrid(np.arange(patch_verts), np.arange(patch_verts))
    tri = []
    for i in range(q):
        for j in range(q):
            a = i * patch_verts + j
            tri.append((a, a + patch_verts, a + 1))
            tri.append((a + 1, a + patch_verts, a + patch_verts + 1))
    tri = np.asarray(tri, dtype=np.uint32)
    batches = []
    for pz in range(patches):
        for px in range(patches):
            x = (px * psize + jj.reshape(-1) * (psize / q)).astype(np.float32)
            z = (pz * psize + ii.reshape(-1) * (psize / q)).astype(np.float32)
            y = (1.5 * _fbm(x, z, SEED)).astype(np.float32)
            v = np.stack([x, y, z, np.ones_like(x)], axis=1)
            b = Batch3D(v, tri, np.stack([x, z], axis=1))
            b.source(PixelSource.StaticTileIndex((pz * patches + px) % 16)).repeat_mode(RepeatMode.RepeatXY).cull_mode(CullMode.Off)
            b.with_computed_normals_fast()
            batches.append(b)
    assets = Assets.default().textures([Tile.from_texture(tex_checker_noise(i)) for i in range(16)])
    lights = []
    cols = [(1.0, 0.85, 0.7), (0.7, 0.85, 1.0), (0.8, 1.0, 0.8), (1.0, 0.8, 1.0)]
    for k, (cx, cz) in enumerate([(8.0, 8.0), (56.0, 8.0), (8.0, 56.0), (56.0, 56.0)]):
        lights.append(Light.new(LightType.Point).with_position([cx, 3.0, cz]).with_color(cols[k]).with_intensity(1.5)
                      .with_start_distance(4.0).with_end_distance(24.0).compile())
    for sx in (20.0, 44.0):
        d = np.array([32.0 - sx, -6.0, 0.0])
        lights.append(Light.new(LightType.Spot).with_position([sx, 6.0, 32.0]).with_direction(d).with_cone_angle(math.pi / 6)
                      .with_color([1.0, 1.0, 0.9]).with_intensity(2.0).with_start_distance(2.0).with_end_distance(30.0).compile())
    lights.append(Light.new(LightType.Area).with_position([32.0, 5.0, 32.0]).with_normal([0.0, -1.0, 0.0]).with_size(4.0, 4.0)
                  .with_color([1.0, 0.95, 0.9]).with_intensity(0.5).with_start_distance(3.0).with_end_distance(20.0).compile())
    scene = Scene.from_static([], batches).lights_(lights)
    cam = _firstp([32.0, 6.0, -4.0], [32.0, 0.0, 32.0])
    return Config("dense", scene, assets, width, height, tile_size, SampleMode.Linear, (0.2,) * 4, cam)


BUILDERS = {"cube": cube, "teapot": teapot, "map": map_config, "dense": dense, "sweep": sweep}
BUILDERS["chunked"] = lambda **kw: chunked_config(**kw)
BUILDERS["game2d"] = lambda **kw: game2d_config(**kw)


# ------------------------------------------------------------------------------------------------
# the chunk path and the 2D game path (row f3)
# ------------------------------------------------------------------------------------------------
def tex_glass(w=32, h=32) -> Texture:
    """A window pane: opaque frame, translucent glass (alpha 96) with a tint gradient."""
    y, x = np.mgrid[0:h, 0:w].astype(np.float32)
    frame = (x < 2) | (x >= w - 2) | (y < 2) | (y >= h - 2) | (np.abs(x - w / 2) < 1) | (np.abs(y - h / 2) < 1)
    r = np.where(frame, 60, 90 + 2 * x)
    g = np.where(frame, 40, 150 + y)
    b = np.where(frame, 30, 220)
    a = np.where(frame, 255, 96)
    return Texture.from_array(_rgba(r, g, b, a))


def tex_terrain(seed, size=64, holes=True) -> Texture:
    """A baked terrain texture: blotchy grass/dirt; with `holes` a few texels are not opaque, which
    exercises the alpha test of PixelSource::Terrain (its texel comes from the world position)."""
    n = rand01(size * size, seed).reshape(size, size)
    y, x = np.mgrid[0:size, 0:size].astype(np.float32)
    blot = np.sin(x * 0.31 + seed % 7) * np.cos(y * 0.23) * 0.5 + 0.5
    r = 60 + 70 * blot + 30 * n
    g = 110 + 60 * (1 - blot) + 30 * n
    b = 40 + 30 * n
    a = np.full((size, size), 255.0)
    if holes:
        a[(x.astype(int) % 16 == 5) & (y.astype(int) % 16 == 9)] = 128
        a[(x.astype(int) % 16 == 11) & (y.astype(int) % 16 == 3)] = 0
    return Texture.from_array(_rgba(r, g, b, a))


def tex_sprite(seed, w=24, h=32) -> Texture:
    """A character sprite: opaque body, soft (partially transparent) outline, transparent background."""
    y, x = np.mgrid[0:h, 0:w].astype(np.float32)
    d = np.hypot((x - w / 2) / (w / 2), (y - h / 2) / (h / 2))
    a = np.clip((1.05 - d) * 6.0, 0.0, 1.0) * 255.0
    n = rand01(w * h, seed).reshape(h, w)
    return Texture.from_array(_rgba(200 * (1 - d) + 40 * n, 80 + 100 * n, 60 + 150 * d, a))


def _floor_quad(x0, z0, x1, z1, y=0.0):
    verts = [(x0, y, z0, 1.0), (x1, y, z0, 1.0), (x1, y, z1, 1.0), (x0, y, z1, 1.0)]
    uvs = [(x0, z0), (x1, z0), (x1, z1), (x0, z1)]
    return verts, uvs


def chunked_assets(logo_size=256) -> Assets:
    a = map_assets(logo_size)
    a.tile_list.append(Tile.from_texture(tex_glass()))           # 6
    a._generation += 1
    a.entity_tiles = {"hero": [("idle", Tile.from_textures([tex_sprite(SEED + 40), tex_sprite(SEED + 41)])),
                               ("walk", Tile.from_texture(tex_sprite(SEED + 42)))]}
    a.item_tiles = {"chest": [("closed", Tile.from_texture(tex_sprite(SEED + 43, 16, 16)))]}
    return a


def chunked_scene() -> Scene:
    """The 16x16 room of the map config as four 8x8 chunks (scene.chunks): per chunk opaque walls with a
    profile id, window panes in the opacity pass that share the profile id of the wall they sit in, a terrain
    batch textured from chunk.terrain_texture, occluded sectors and a chunk light; plus static batches with
    EntityTile / ItemTile sources (one of which does not resolve) and a 2D overlay of every primitive mode."""
    H = 2.0
    scene = Scene()

    def fin(b, tile, profile=None):
        b = b.source(PixelSource.StaticTileIndex(tile)).repeat_mode(RepeatMode.RepeatXY).cull_mode(CullMode.Off).with_computed_normals()
        return b if profile is None else b.profile_id(profile)

    profile = 100
    for cz in range(2):
        for cx in range(2):
            ox, oz = cx * 8, cz * 8
            ch = Chunk((ox, oz), 8)
            walls, panes = [], []
            # outer walls of the room that fall into this chunk, each with a window pane 2 cm in front of it
            if cz == 0:
                walls.append((_wall_quad(float(ox), 0.0, float(ox + 8), 0.0, H), profile))
                panes.append((_wall_quad(ox + 2.0, 0.02, ox + 6.0, 0.02, H), profile))
                profile += 1
            if cz == 1:
                walls.append((_wall_quad(float(ox + 8), 16.0, float(ox), 16.0, H), profile))
                panes.append((_wall_quad(ox + 6.0, 15.98, ox + 2.0, 15.98, H), profile))
                profile += 1
            if cx == 0:
                walls.append((_wall_quad(0.0, float(oz + 8), 0.0, float(oz), H), profile))
                profile += 1
            if cx == 1:
                walls.append((_wall_quad(16.0, float(oz), 16.0, float(oz + 8), H), profile))
                panes.append((_wall_quad(15.98, oz + 1.0, 15.98, oz + 7.0, H), profile + 1000))  # profile of no wall
                profile += 1
            for q, pid in walls:
                ch.batches3d.append(fin(_quads_batch([q]), 1, pid))
            for q, pid in panes:
                ch.batches3d_opacity.append(fin(_quads_batch([q]), 6, pid).repeat_mode(RepeatMode.ClampXY))
            # an inner fence (alpha tested) without a profile id
            ch.batches3d.append(fin(_quads_batch([_wall_quad(ox + 4.0, oz + 1.0, ox + 4.0, oz + 7.0, H)]), 3))
            t = _quads_batch([_floor_quad(float(ox), float(oz), float(ox + 8), float(oz + 8))])
            ch.terrain_batch3d = t.source(PixelSource.Terrain).cull_mode(CullMode.Off).with_computed_normals()
            ch.terrain_texture = tex_terrain(SEED + 50 + cx + 2 * cz, 64, holes=(cx != cz))
            ch.occluded_sectors = [(BBox((ox + 1.0, oz + 1.0), (ox + 3.5, oz + 4.0)), 0.35), (BBox((ox + 5.0, oz + 5.0), (ox + 7.0, oz + 7.5)), 0.0)]
            ch.lights = [Light.new(LightType.Point).with_color([1.0, 0.9 - 0.2 * cx, 0.7 + 0.3 * cz]).with_intensity(1.5)
                         .with_start_distance(1.0).with_end_distance(7.0).with_position([ox + 4.0, 1.5, oz + 4.0]).compile()]
            # 2D: a minimap tile of the chunk and its terrain as seen from above
            ch.batches2d.append(Batch2D.from_rectangle(20.0 + 40.0 * cx, 20.0 + 40.0 * cz, 36.0, 36.0).source(PixelSource.StaticTileIndex(4)))
            ch.terrain_batch2d = Batch2D.from_rectangle(120.0 + 40.0 * cx, 20.0 + 40.0 * cz, 36.0, 36.0).source(PixelSource.Terrain)
            scene.chunks[(cx, cz)] = ch

    sky = Batch3D.from_box(-32.5, -20.0, -32.5, 80.0, 60.0, 80.0)
    scene.d3_static.append(fin(sky, 5).receives_light(False))

    def billboard(x, z, w, h, src):
        verts = [(x - w / 2, 0.0, z, 1.0), (x + w / 2, 0.0, z, 1.0), (x + w / 2, h, z, 1.0), (x - w / 2, h, z, 1.0)]
        uvs = [(0.0, 1.0), (1.0, 1.0), (1.0, 0.0), (0.0, 0.0)]
        return Batch3D(verts, [(0, 1, 2), (0, 2, 3)], uvs).source(src).cull_mode(CullMode.Off).with_computed_normals()

    scene.d3_dynamic += [billboard(6.0, 9.0, 1.0, 1.6, PixelSource.EntityTile("hero", 0)),
                         billboard(9.5, 10.0, 1.0, 1.6, PixelSource.EntityTile("hero", 1)),
                         billboard(8.0, 11.0, 0.8, 0.8, PixelSource.ItemTile("chest", 0)),
                         billboard(7.0, 8.0, 1.0, 1.0, PixelSource.ItemTile("missing", 0)),   # does not resolve
                         billboard(10.0, 8.5, 1.0, 1.0, PixelSource.EntityTile("hero", 7))]   # index out of range
    # 2D overlay: sprites, a translucent stack, lines of every mode
    scene.d2_static.append(Batch2D.from_rectangle(300.0, 20.0, 48.0, 64.0).source(PixelSource.EntityTile("hero", 0)))
    scene.d2_static.append(Batch2D.from_rectangle(330.0, 40.0, 48.0, 64.0).source(PixelSource.ItemTile("chest", 0)).receives_light(False))
    scene.d2_static.append(Batch2D.from_rectangle(360.0, 30.0, 40.0, 40.0).source(PixelSource.ItemTile("nothing", 0)))
    pts = [(420.5, 20.2), (470.9, 28.0), (455.0, 70.7), (430.2, 60.1), (445.0, 45.0)]
    uv0 = [(0.0, 0.0)] * len(pts)
    scene.d2_dynamic.append(Batch2D.new(pts, [(0, 2, 0), (1, 3, 0), (4, 0, 0)], uv0).mode_(PrimitiveMode.Lines).source(PixelSource.Pixel((255, 220, 0, 255))))
    scene.d2_dynamic.append(Batch2D.new([(x + 70.0, y) for x, y in pts], [], uv0).mode_(PrimitiveMode.LineStrip))
    scene.d2_dynamic.append(Batch2D.new([(x + 140.0, y + 0.4) for x, y in pts], [], uv0).mode_(PrimitiveMode.LineLoop).source(PixelSource.Pixel((0, 255, 255, 128))))
    scene.lights = [Light.new(LightType.AmbientDaylight).with_color([0.4, 0.45, 0.6]).with_intensity(0.6).compile()]
    return scene


def chunked_config(width=1280, height=720, tile_size=40, n_frames=8) -> Config:
    def cams(i):
        th = 2.0 * math.pi * (i + 0.37) / n_frames
        return _firstp([8.0 + 3.0 * math.cos(th), 1.1, 8.0 + 3.0 * math.sin(th)], [8.0 - 2.0 * math.cos(th), 0.9, 8.0 - 2.0 * math.sin(th)])
    mm = MapMini(occluded_sectors=[(BBox((7.0, 7.0), (9.0, 9.0)), 0.5)])
    return Config("chunked", chunked_scene(), chunked_assets(), width, height, tile_size, SampleMode.Nearest, (0.9, 0.9, 1.0, 1.0),
                  cams(0), cameras=cams, n_frames=n_frames, mapmini=mm)


def game2d_scene(cols=24, rows=16, tile_px=1.0) -> Scene:
    """A top-down 2D game screen as the screen widget builds it (src/client/widget/screen.rs:60-94): a grid of
    map tiles (cols*rows quads = 2*cols*rows records), translucent decals that overlap in submission order, entity
    sprites, point lights with line-of-sight against mapmini linedefs, sector occlusion and line overlays."""
    scene = Scene()
    grid = Batch2D.empty()
    for r in range(rows):
        for c in range(cols):
            grid.add_rectangle(c * tile_px, r * tile_px, tile_px, tile_px)
    scene.d2_static.append(grid.source(PixelSource.StaticTileIndex(4)).repeat_mode(RepeatMode.RepeatXY))
    walls = Batch2D.empty()
    for c in range(4, 20):
        walls.add_rectangle(c * tile_px, 5 * tile_px, tile_px, tile_px)
    scene.d2_static.append(walls.source(PixelSource.StaticTileIndex(1)))
    decals = Batch2D.empty()
    for i in range(12):
        decals.add_rectangle(2.0 + 1.3 * i, 7.0 + 0.35 * i, 2.5, 2.5)   # overlapping, translucent: order matters
    scene.d2_static.append(decals.source(PixelSource.StaticTileIndex(6)))
    scene.d2_dynamic.append(Batch2D.from_rectangle(9.2, 9.1, 1.5, 2.0).source(PixelSource.EntityTile("hero", 0)))
    scene.d2_dynamic.append(Batch2D.from_rectangle(13.4, 3.2, 1.5, 2.0).source(PixelSource.EntityTile("hero", 1)).receives_light(False))
    scene.d2_dynamic.append(Batch2D.from_rectangle(16.0, 10.0, 1.0, 1.0).source(PixelSource.ItemTile("chest", 0)))
    pts = [(1.2, 1.1), (22.7, 2.3), (20.1, 14.6), (3.3, 13.2), (12.0, 8.0)]
    uv0 = [(0.0, 0.0)] * len(pts)
    scene.d2_dynamic.append(Batch2D.new(pts, [(0, 2, 0), (1, 3, 0)], uv0).mode_(PrimitiveMode.Lines).source(PixelSource.Pixel((255, 0, 0, 255))))
    scene.d2_dynamic.append(Batch2D.new(pts, [], uv0).mode_(PrimitiveMode.LineLoop))
    scene.lights = [
        Light.new(LightType.Point).with_color([1.0, 0.8, 0.5]).with_intensity(1.2).with_start_distance(1.0).with_end_distance(9.0).with_position([8.0, 0.0, 3.0]).compile(),
        Light.new(LightType.Point).with_color([0.4, 0.6, 1.0]).with_intensity(1.0).with_start_distance(0.5).with_end_distance(8.0).with_position([15.0, 0.0, 11.0]).compile(),
        Light.new(LightType.AmbientDaylight).with_color([0.5, 0.5, 0.5]).with_intensity(0.5).compile(),
    ]
    return scene


def game2d_config(width=960, height=640) -> Config:
    scale = float(width) / 24.0
    m = np.array([[scale, 0.0, 0.0], [0.0, scale, 0.0], [0.0, 0.0, 1.0]], dtype=np.float32)
    mm = MapMini(linedefs=[CompiledLinedef((4.0, 5.5), (20.0, 5.5)), CompiledLinedef((12.0, 8.0), (12.0, 13.0))],
                 occluded_sectors=[(BBox((0.0, 0.0), (6.0, 4.0)), 0.2), (BBox((18.0, 8.0), (24.0, 16.0)), 0.6)])
    cam = D3FirstPCamera.new()
    return Config("game2d", game2d_scene(), chunked_assets(), width, height, 40, SampleMode.Nearest, (0.25, 0.25, 0.3, 1.0), cam,
                  matrix2d=m, mapmini=mm, render_mode=RenderMode.render_2d())


# ------------------------------------------------------------------------------------------------
# batch shaders: Rusteria VM programs on 3D, opacity-pass and 2D batches
# ------------------------------------------------------------------------------------------------
def pattern_bank(size=64):
    """Stand-ins for rusteria's patterns() bank (textures/patterns.rs:27-37: value, fbm_value, perlin, fbm_perlin,
    bricks, tiles, blocks).  The reference computes its bank on the host once per process and the rasterizer only
    looks texels up; what matters here is that oracle and device read the same host-provided arrays."""
    y, x = np.mgrid[0:size, 0:size].astype(np.float32)
    bank = []
    for k in range(7):
        n = rand01(size * size, SEED + 900 + k).reshape(size, size).astype(np.float32)
        if k < 4:      # smooth noise: a few octaves of box-filtered white noise, tileable through np.roll
            v = np.zeros((size, size), np.float32)
            amp, tot = 1.0, 0.0
            for o in range(1 + k):
                r = 2 ** (3 - min(o, 3))
                s = sum(np.roll(np.roll(n, dx * (o + 1), 1), dy * (o + 1), 0) for dx in range(-r, r + 1) for dy in range(-r, r + 1)) / float((2 * r + 1) ** 2)
                v += amp * s
                tot += amp
                amp *= 0.5
            v = (v / tot - v.min() / tot) / max(1e-6, (v.max() - v.min()) / tot)
        elif k == 4:   # bricks
            row = (y // 8).astype(int)
            xx = (x + 8 * (row % 2)) % 16
            v = np.where((y % 8 < 1) | (xx < 1), 0.1, 0.6 + 0.4 * n)
        elif k == 5:   # tiles
            v = np.where((x % 16 < 2) | (y % 16 < 2), 0.0, 1.0)
        else:          # blocks
            v = ((x // 8 + y // 8) % 3) / 2.0
        v = v.astype(np.float32)
        bank.append((size, size, np.stack([v, 0.5 * v + 0.25 * n, 1.0 - v], axis=-1).reshape(-1, 3).astype(np.float32)))
    return bank


def shader_wood():
    """rusteria/examples/wood.rusteria (= the shade() of examples/cube_shaded.rs:46-102), lowered by hand the way the
    Rusteria compiler does it.  `vec2(1.5)` is (1.5, 0, 0): the parser pads a one-argument vec2 with a zero
    (rusteria/src/parser.rs:880-896; vec3 broadcasts instead).  Rendered over uv like rsia does, the oracle's VM
    reproduces the reference's own wood.png bit for bit (tests/test_rusteria_golden.py)."""
    from . import vm
    b = vm.Body()
    t = b.let(vm.time_ * 0.0)
    uv2 = b.let(vm.uv / 3.0 - vm.vec2(1.5, 0.0))
    n1 = b.let(vm.sample(uv2 + vm.vec2(t.x, 0.0), "fbm_perlin"))
    n2 = b.let(vm.sample(uv2 * 2.0 + vm.vec2(0.0, (t * 0.7).x), "fbm_perlin"))
    turb = b.let(0.65 * n1 + 0.35 * n2)
    turb_zm = b.let((turb - 0.5) * 2.0)
    r = b.let(vm.length(uv2))
    rings = b.let(r + 0.22 * turb_zm)
    waves = b.let(vm.sin(rings * 10.0))
    rings_mask = b.let(vm.pow_(1.0 - vm.abs_(waves), 3.0))
    grain_uv = b.let(vm.vec2(uv2.x * 8.0, uv2.y * 40.0))
    g = b.let(vm.sample(grain_uv + vm.vec2(0.0, (t * 0.5).x), "value"))
    grain = b.let((g - 0.5) * 2.0)
    b.set("Color", vm.mix((0.72, 0.52, 0.32), (0.45, 0.30, 0.16), rings_mask))
    b.set("Color", vm.color * (1.0 + 0.06 * grain))
    band = b.let(uv2.y + 0.15 * turb_zm)
    cathedral = b.let(vm.pow_(1.0 - vm.abs_(vm.sin(band * 6.0)), 4.0))
    b.set("Color", vm.mix(vm.color, vm.color * 0.9, cathedral * 0.2))
    b.set("Roughness", 0.6 + cathedral * 0.3)
    return vm.Program([b.code], 0, b.n_locals, 0)


def shader_marble():
    """rusteria/examples/marble.rusteria lowered by hand (golden: the reference's marble.png)."""
    from . import vm
    b = vm.Body()
    t = b.let(1.0)
    uv2 = b.let(vm.uv * 1.0)
    n1 = b.let(vm.sample(uv2 + vm.vec2(t, 0.0), "fbm_perlin"))
    n2 = b.let(vm.sample(uv2 * 2.0 + vm.vec2(0.0, t), "fbm_perlin"))
    turb = b.let(0.6 * n1 + 0.4 * n2)
    bands = b.let(uv2.x + turb * 0.6)
    s = b.let(vm.sin(bands * 8.0))
    veins = b.let(vm.pow_(1.0 - vm.abs_(s), 3.0))
    b.set("Color", vm.mix(vm.vec3(0.92, 0.93, 0.96), vm.vec3(0.18, 0.20, 0.24), veins))
    m = b.let(vm.sample(uv2 * 0.5 + vm.vec2(0.0, 0.0), "value"))
    b.set("Color", vm.color * (0.9 + 0.1 * m))
    return vm.Program([b.code], 0, b.n_locals, 0)


def shader_wood_ring():
    """rusteria/examples/wood_ring.rusteria lowered by hand: a user function (`pdelta`, whose value is the expression
    statement left on the stack), fract, nested vec2 (golden: the reference's wood_ring.png)."""
    from . import vm
    f = vm.Body(2)   # fn pdelta(a, c) { fract(a - c + 0.5) - 0.5; }
    f.code += (vm.fract(f.param(0) - f.param(1) + 0.5) - 0.5).ops
    b = vm.Body()
    t = b.let(vm.time_ * 0.1)
    uv0 = b.let(vm.fract(vm.uv))
    cx = b.let(0.5)
    cy = b.let(0.5)
    dx = b.let(vm.call(0, 2, uv0.x, cx))
    dy = b.let(vm.call(0, 2, uv0.y, cy))
    r = b.let(vm.length(vm.vec2(dx, dy)))
    w1 = b.let(vm.sample(vm.fract(uv0 * 2.0 + vm.vec2(t, 0.0)), "fbm_perlin"))
    w2 = b.let(vm.sample(vm.fract(uv0 * 4.0 + vm.vec2(0.0, t)), "fbm_perlin"))
    turb = b.let((0.6 * w1 + 0.4 * w2) - 0.5)
    ring_freq = b.let(14.0)
    ring_warp = b.let(0.05)
    phase = b.let(r + ring_warp * turb)
    waves = b.let(vm.sin(phase * ring_freq * 6.2831853))
    rings_mask = b.let(vm.pow_(1.0 - vm.abs_(waves), 6.0))
    grain_uv = b.let(vm.fract(vm.vec2(uv0.x * 8.0, uv0.y * 64.0)))
    g = b.let(vm.sample(grain_uv + vm.vec2(0.0, vm.fract(t)), "value"))
    grain = b.let((g - 0.5) * 2.0)
    b.set("Color", vm.mix(vm.vec3(0.72, 0.52, 0.32), vm.vec3(0.45, 0.30, 0.16), rings_mask))
    b.set("Color", vm.color * (1.0 + 0.05 * grain))
    b.set("Roughness", 0.6)
    return vm.Program([f.code, b.code], 1, b.n_locals, 0)


def rusteria_pattern_bank(directory):
    """rusteria's patterns() bank from its embedded PNGs (rusteria/src/textures/patterns.rs:136-175 via
    TexStorage::from_png_bytes, textures/mod.rs:64-82: channel / 255).  Patterns whose PNG is not in `directory`
    stay 1x1 zero textures, like an entry build_patterns() did not find."""
    from PIL import Image
    from . import vm
    bank = [(1, 1, np.zeros((1, 3), np.float32)) for _ in range(7)]
    for name, i in vm.PATTERN_INDEX.items():
        path = os.path.join(directory, name + ".png")
        if os.path.exists(path):
            im = np.asarray(Image.open(path).convert("RGB"))
            bank[i] = (im.shape[1], im.shape[0], (im.astype(np.float32) / np.float32(255.0)).reshape(-1, 3))
    return bank


def rsia_records(width, height, xs=None, ys=None):
    """Execution inputs of `rsia file.rusteria` (Rusteria::shade, rusteria/src/lib.rs:161-206): uv = (x / w, 1 - y / h, 0),
    everything else zero, for the pixels (xs, ys) or the whole image.  Returns [n, 18] float32 records."""
    if xs is None:
        ys, xs = np.mgrid[0:height, 0:width]
    xs, ys = np.asarray(xs).ravel(), np.asarray(ys).ravel()
    rec = np.zeros((xs.size, 18), np.float32)
    rec[:, 0] = xs.astype(np.float32) / np.float32(width)
    rec[:, 1] = np.float32(1.0) - ys.astype(np.float32) / np.float32(height)
    return rec


def rsia_pixels(colors):
    """RenderBuffer::save (rusteria/src/renderbuffer.rs:166-182): `(c * 255.0) as u8`, a saturating truncation."""
    c = np.nan_to_num(np.asarray(colors, np.float32) * np.float32(255.0), nan=0.0)
    return np.clip(np.trunc(c), 0, 255).astype(np.uint8)


def shader_holes():
    """Writes opacity at the top level of shade() (Program::shader_supports_opacity): a perforated sheet."""
    from . import vm
    b = vm.Body()
    m = b.let(vm.sample(vm.uv * 6.0, "tiles"))
    b.set("Opacity", m.x)
    b.set("Color", vm.color * (0.6 + 0.4 * vm.sample(vm.uv * 2.0, "bricks").x))
    b.set("Metallic", 0.7)
    b.set("Roughness", 0.25)
    return vm.Program([b.code], 0, b.n_locals, 0)


def shader_control_flow(emissive=False):
    """For / If / FunctionCall / Return / globals / swizzle writes: stripes counted in a loop through a helper.
    `emissive`: one branch also writes `emissive`, which the reference never resets (src/rasterizer.rs:310): the
    value then leaks into every later fragment of the screen tile, so frames with it are compared against the
    oracle's per-fragment-state mode (deviations)."""
    from . import vm
    f = vm.Body(n_params=2)          # fn band(x, k): if fract(x * k) < 0.5 { return 1 } 0.25
    inner = f.sub()
    inner.ret(1.0)
    f.if_(vm.fract(f.param(0) * f.param(1)) < 0.5, inner)
    f.code += vm.X.of(0.25).ops
    b = vm.Body()
    acc = b.let(0.0)
    i = b.let(0.0)
    init, incr, body = b.sub(), b.sub(), b.sub()
    incr.assign(i, i + 1.0)
    body.assign(acc, acc + vm.call(1, 2, vm.uv.x * 4.0 + vm.uv.y, i + 1.0) * 0.25)
    b.for_(init, i < 4.0, incr, body)
    b.set_global(0, acc)
    tint = b.let(vm.palette(2.0))
    b.code += tint.ops + vm.X.of(0.5).ops + [("SetComponents", [1])] + [("StoreLocal", tint.ops[0][1])]   # tint.y = 0.5
    then, other = b.sub(), b.sub()
    then.set("Color", vm.X([("LoadGlobal", 0)]) * tint)
    other.set("Color", vm.color * 0.5)
    if emissive:
        other.set("Emissive", vm.vec3(0.15, 0.0, 0.0) * vm.abs_(vm.sin(vm.hitpoint.x * 4.0)))
    b.if_(vm.normal.y > 0.5, then, other)
    return vm.Program([b.code, f.code], 0, b.n_locals, 1)


def shader_glass_tint():
    """For an opacity-pass batch: colour and opacity from the program (rasterizer.rs:1611-1645)."""
    from . import vm
    b = vm.Body()
    b.set("Color", vm.mix(vm.color, (0.1, 0.4, 0.8), 0.5 + 0.5 * vm.sin(vm.hitpoint.y * 6.0)))
    b.set("Opacity", 0.35 + 0.3 * vm.fract(vm.uv.x * 8.0))
    return vm.Program([b.code], 0, b.n_locals, 0)


def shader_2d_scanlines():
    """A 2D batch program (rasterizer.rs:760-797): sRGB in, sRGB out, alpha forced to 1."""
    from . import vm
    b = vm.Body()
    b.set("Color", vm.color * (0.55 + 0.45 * vm.step(0.5, vm.fract(vm.uv.y * 24.0))) + vm.vec3(0.0, 0.0, 0.1) * vm.cos(vm.hitpoint.x * 0.05))
    return vm.Program([b.code], 0, b.n_locals, 0)


def shaded_scene(logo_size=256, emissive=False) -> Scene:
    scene = Scene()
    scene.patterns = pattern_bank()
    scene.patterns_normal = pattern_bank()[:3]
    wood = scene.add_shader(shader_wood())
    holes = scene.add_shader(shader_holes())
    flow = scene.add_shader(shader_control_flow(emissive))
    scan = scene.add_shader(shader_2d_scanlines())

    def fin(b, tile):
        return b.source(PixelSource.StaticTileIndex(tile)).cull_mode(CullMode.Off).with_computed_normals()

    scene.d3_static.append(fin(Batch3D.from_box(-0.5, -0.5, -0.5, 1.0, 1.0, 1.0), 0).shader(wood))
    scene.d3_static.append(fin(Batch3D.from_box(-1.6, -0.4, -0.4, 0.8, 0.8, 0.8), 1).repeat_mode(RepeatMode.RepeatXY).shader(flow))
    # a perforated sheet in front of the boxes; what is behind shows through the holes the program cuts
    sheet = Batch3D([(-1.2, -0.7, 0.9, 1.0), (1.2, -0.7, 0.9, 1.0), (1.2, 0.7, 0.9, 1.0), (-1.2, 0.7, 0.9, 1.0)], [(0, 1, 2), (0, 2, 3)],
                    [(0.0, 0.0), (4.0, 0.0), (4.0, 2.0), (0.0, 2.0)])
    scene.d3_static.append(fin(sheet, 1).repeat_mode(RepeatMode.RepeatXY).shader(holes))
    scene.d3_static.append(fin(Batch3D.from_box(0.8, -0.45, -0.45, 0.9, 0.9, 0.9), 0).shader(7))   # no such program: nothing runs
    # one chunk: a wall shaded by the chunk's own program 0, a wall whose program 1 was baked into a texture,
    # and a pane in the opacity pass with program 2
    ch = Chunk((-4, -4), 8)
    ch.shaders = [shader_control_flow(emissive), shader_wood(), shader_glass_tint()]
    ch.shader_textures = [None, tex_brick(SEED + 77, (90, 140, 60)), None]
    ch.batches3d.append(fin(_quads_batch([_wall_quad(-2.5, -1.5, 2.5, -1.5, 2.0)]), 1).repeat_mode(RepeatMode.RepeatXY).shader(0))
    ch.batches3d.append(fin(_quads_batch([_wall_quad(2.5, -1.5, 2.5, 1.5, 2.0)]), 1).repeat_mode(RepeatMode.RepeatXY).shader(1))
    ch.batches3d_opacity.append(fin(_quads_batch([_wall_quad(-2.4, 1.4, -2.4, -1.4, 1.5)]), 3).repeat_mode(RepeatMode.RepeatXY).shader(2))
    scene.chunks[(0, 0)] = ch
    scene.d2_static.append(Batch2D.from_rectangle(8.0, 8.0, 120.0, 90.0).source(PixelSource.StaticTileIndex(0)).shader(scan))
    scene.d2_static.append(Batch2D.from_rectangle(100.0, 60.0, 60.0, 60.0).source(PixelSource.StaticTileIndex(3)).shader(scan))
    scene.lights = [_example_light(),
                    Light.new(LightType.Point).with_color([0.6, 0.7, 1.0]).with_intensity(1.2).with_start_distance(1.0).with_end_distance(6.0)
                    .with_position([-1.5, 1.2, 2.0]).compile()]
    scene.background_(VGrayGradientShader())
    return scene


def shaded_config(width=640, height=480, tile_size=40, n_frames=16, emissive=False) -> Config:
    assets = map_assets(256)
    assets.palette = [(0.1, 0.1, 0.1), None, (0.9, 0.7, 0.3), (0.2, 0.8, 0.4)]

    def cams(i):
        cam = D3OrbitCamera.new()
        cam.set_parameter_f32("distance", 3.2)
        cam.azimuth = math.pi / 2 + 2.0 * math.pi * i / n_frames * 0.35 - 0.3
        return cam
    cfg = Config("shaded", shaded_scene(emissive=emissive), assets, width, height, tile_size, SampleMode.Nearest, (0.25, 0.25, 0.3, 1.0), cams(0),
                 cameras=cams, n_frames=n_frames)
    cfg.time = 1.25
    return cfg


BUILDERS["shaded"] = lambda **kw: shaded_config(**kw)


# ------------------------------------------------------------------------------------------------
# render graph: Sky node on the pixels nothing covers, directional sun, brush preview
# ------------------------------------------------------------------------------------------------
def sky_config(width=640, height=360, tile_size=40, hour=16.5, n_frames=8) -> Config:
    from .types import BrushPreview, RenderGraph, SkyNode
    scene = Scene()

    def fin(b, tile):
        return b.source(PixelSource.StaticTileIndex(tile)).repeat_mode(RepeatMode.RepeatXY).cull_mode(CullMode.Off).with_computed_normals()

    scene.d3_static.append(fin(Batch3D.from_box(-0.5, 0.0, -0.5, 1.0, 1.0, 1.0), 0))
    scene.d3_static.append(fin(Batch3D.from_box(1.2, 0.0, 0.4, 0.6, 1.8, 0.6), 1))
    scene.d3_static.append(fin(_quads_batch([_wall_quad(-3.0, -2.0, 3.0, -2.0, 1.5), _wall_quad(-1.5, 1.5, -1.5, 0.2, 1.2)]), 3))  # fences: alpha holes show the sky
    scene.d3_static.append(fin(_quads_batch([_floor_quad(-2.0, -2.0, 2.5, 2.0)]), 2))
    scene.lights = [Light.new(LightType.Point).with_color([1.0, 0.8, 0.6]).with_intensity(0.8).with_start_distance(1.0).with_end_distance(5.0)
                    .with_position([0.5, 1.5, 1.5]).compile()]

    def cams(i):
        cam = D3OrbitCamera.new()
        cam.set_parameter_f32("distance", 4.5)
        cam.azimuth = math.pi / 2 + 2.0 * math.pi * i / n_frames
        cam.elevation = 0.25 + 0.05 * i
        return cam
    cfg = Config("sky", scene, map_assets(128), width, height, tile_size, SampleMode.Nearest, None, cams(0), cameras=cams, n_frames=n_frames)
    cfg.render_graph = RenderGraph([SkyNode(clouds=False)])
    cfg.hour = hour
    cfg.brush_preview = BrushPreview((3.0, 0.0, 1.0), 1.5, 0.6)
    return cfg


BUILDERS["sky"] = lambda **kw: sky_config(**kw)
