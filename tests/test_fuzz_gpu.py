"""Differential fuzzing of the CUDA path against the oracle: seeded random triangle soups (every cull / repeat /
sample mode, alpha-holed textures, near-plane crossings, slivers, shared edges, coincident depths, large and
sub-pixel triangles, random lights and 2D overlays) rendered through the C ABI and compared with the parity bar --
owner and depth planes bit for bit.  Run on the B200 box with `pytest -m gpu`."""
import math

import numpy as np
import pytest

from helpers import compare, render_gpu, render_oracle
from rusterix_b200 import (Assets, Batch2D, Batch3D, CullMode, D3FirstPCamera, Light, LightType, PixelSource, PrimitiveMode,
                           Rasterizer, RepeatMode, SampleMode, Scene, Texture, Tile, scenes)

pytestmark = pytest.mark.gpu


def _rng(seed):
    return np.random.default_rng(0x52555354 + seed)


def _texture(rng, holes):
    w, h = int(rng.choice([1, 2, 5, 16, 33, 64])), int(rng.choice([1, 3, 8, 16, 40]))
    px = rng.integers(0, 256, size=(h, w, 4), dtype=np.uint8)
    px[..., 3] = 255
    if holes:
        m = rng.random((h, w)) < 0.3
        px[..., 3][m] = rng.choice([0, 1, 128, 254], size=int(m.sum()))
    return Texture.from_array(px)


def _soup(rng, n, spread, size):
    """n triangles around the origin; a third of them share vertices (common edges), some are duplicated (equal depth)."""
    verts, tris, uvs = [], [], []
    for i in range(n):
        c = rng.normal(0.0, spread, 3)
        kind = rng.integers(0, 6)
        s = size * (0.02 if kind == 0 else 8.0 if kind == 1 else 1.0)
        p = [c + rng.normal(0.0, s, 3) for _ in range(3)]
        if kind == 2:       # sliver
            p[2] = p[0] + (p[1] - p[0]) * 0.5 + rng.normal(0.0, 1e-4, 3)
        base = len(verts)
        if kind == 3 and tris:   # share an edge with the previous triangle
            a, b, _ = tris[-1]
            verts.append(tuple(p[2]) + (1.0,))
            uvs.append(tuple(rng.random(2) * 3.0 - 1.0))
            tris.append((b, a, base))
            continue
        for q in p:
            verts.append(tuple(q) + (1.0,))
            uvs.append(tuple(rng.random(2) * 3.0 - 1.0))
        tris.append((base, base + 1, base + 2))
        if kind == 4:       # an exact duplicate drawn later: first drawn must keep the pixel
            tris.append((base, base + 1, base + 2))
    return np.asarray(verts, dtype=np.float32), tris, uvs


def _scene(seed):
    rng = _rng(seed)
    n_tex = int(rng.integers(1, 5))
    assets = Assets.default().textures([Tile.from_texture(_texture(rng, holes=bool(rng.integers(0, 2)))) for _ in range(n_tex)])
    scene = Scene()
    for b in range(int(rng.integers(1, 6))):
        v, t, uv = _soup(rng, int(rng.integers(1, 60)), float(rng.choice([0.5, 2.0, 6.0])), float(rng.choice([0.2, 1.0, 3.0])))
        batch = Batch3D(v, t, uv)
        src = PixelSource.StaticTileIndex(int(rng.integers(0, n_tex))) if rng.random() < 0.8 else PixelSource.Pixel((int(rng.integers(0, 256)), 90, 200, int(rng.choice([255, 255, 100]))))
        batch = batch.source(src).cull_mode(CullMode(int(rng.integers(0, 3)))).repeat_mode(RepeatMode(int(rng.integers(0, 4))))
        if rng.random() < 0.8:
            batch = batch.with_computed_normals()
        if rng.random() < 0.3:
            ang = rng.random() * 6.28
            m = np.eye(4, dtype=np.float32)
            m[0, 0], m[0, 2], m[2, 0], m[2, 2] = math.cos(ang), math.sin(ang), -math.sin(ang), math.cos(ang)
            m[:3, 3] = rng.normal(0.0, 1.0, 3)
            batch.transform_3d = m
        (scene.d3_static if rng.random() < 0.6 else scene.d3_dynamic if rng.random() < 0.7 else scene.d3_overlay).append(batch)
    if rng.random() < 0.35:   # general mode: a chunk with opacity-pass batches, surface ids, occluded sectors, a terrain texture
        from rusterix_b200 import BBox, Chunk
        ch = Chunk((int(rng.integers(-4, 2)), int(rng.integers(-4, 2))), int(rng.choice([4, 8])))
        for k in range(int(rng.integers(1, 4))):
            v, t, uv = _soup(rng, int(rng.integers(1, 20)), 2.0, float(rng.choice([0.5, 2.0])))
            b = Batch3D(v, t, uv).source(PixelSource.StaticTileIndex(int(rng.integers(0, n_tex)))).cull_mode(CullMode.Off).with_computed_normals()
            if rng.random() < 0.7:
                b = b.profile_id(int(rng.integers(0, 3)))
            (ch.batches3d_opacity if rng.random() < 0.5 else ch.batches3d).append(b)
        if rng.random() < 0.5:
            v, t, uv = _soup(rng, int(rng.integers(1, 8)), 2.0, 3.0)
            ch.terrain_batch3d = Batch3D(v, t, uv).source(PixelSource.Terrain).cull_mode(CullMode.Off).with_computed_normals()
            ch.terrain_texture = _texture(rng, holes=bool(rng.integers(0, 2)))
            while ch.terrain_texture.width < ch.size:      # pixels_per_tile = width / size must not be 0 ... it may be: clamp path
                break
        ch.occluded_sectors = [(BBox((float(rng.normal()), float(rng.normal())), (float(rng.normal() + 2), float(rng.normal() + 2))), float(rng.choice([0.0, 0.4, 1.0])))]
        scene.chunks[(0, 0)] = ch
    if rng.random() < 0.25:   # general mode + VM: programs on random batches (colour goes through libm; ownership through `holes` is exact)
        scene.patterns, scene.patterns_normal = scenes.pattern_bank(16), scenes.pattern_bank(16)[:2]
        progs = [scenes.shader_holes(), scenes.shader_control_flow(), scenes.shader_wood(), scenes.shader_glass_tint()]
        for pgm in progs:
            scene.add_shader(pgm)
        assets.palette = [(0.1, 0.1, 0.1), None, (0.9, 0.7, 0.3), (0.2, 0.8, 0.4)]
        for b in scene.d3_static + scene.d3_dynamic:
            if rng.random() < 0.6:
                b.shader(int(rng.integers(0, 5)))      # 4: past the list
    for _ in range(int(rng.integers(0, 4))):
        kind = LightType(int(rng.choice([0, 1, 3, 4, 5])))
        l = (Light.new(kind).with_intensity(float(rng.random() * 2)).with_color(list(rng.random(3))).with_position(list(rng.normal(0, 3, 3)))
             .with_start_distance(float(rng.random() * 2)).with_end_distance(float(2 + rng.random() * 8)))
        scene.lights.append(l.compile())
    for _ in range(int(rng.integers(0, 3))):
        x, y, w, h = rng.random(4) * np.array([200, 120, 150, 90]) - np.array([20, 20, 0, 0])
        b2 = Batch2D.from_rectangle(float(x), float(y), float(w), float(h)).source(PixelSource.StaticTileIndex(int(rng.integers(0, n_tex))))
        scene.d2_static.append(b2.receives_light(bool(rng.integers(0, 2))))
    if rng.random() < 0.3:
        pts = [tuple(rng.random(2) * np.array([260, 160]) - 10) for _ in range(4)]
        scene.d2_dynamic.append(Batch2D.new(pts, [], [(0.0, 0.0)] * 4).mode_(PrimitiveMode(int(rng.integers(2, 4)))))
    cam = D3FirstPCamera.new()
    pos = rng.normal(0.0, 2.5, 3)
    cam.set_parameter_vec3("position", pos.tolist())
    cam.set_parameter_vec3("center", (pos + rng.normal(0.0, 1.0, 3) + 1e-3).tolist())
    cam.set_parameter_f32("fov", float(rng.choice([40.0, 75.0, 110.0])))
    w, h = int(rng.choice([64, 97, 160, 255, 320])), int(rng.choice([48, 65, 120, 200]))
    r = Rasterizer.setup(None, cam.view_matrix(), cam.projection_matrix(float(w), float(h)))
    r.sample_mode(SampleMode(int(rng.integers(0, 2))))
    if rng.random() < 0.7:
        r.ambient(tuple(rng.random(3)) + (1.0,))
    return scene, assets, r, w, h, int(rng.choice([8, 40, 64, 500]))


@pytest.mark.parametrize("block", range(8))
def test_random_scenes(block):
    for seed in range(block * 12, block * 12 + 12):
        scene, assets, r, w, h, ts = _scene(seed)
        g = render_gpu(r, scene, assets, w, h, ts)
        o = render_oracle(r, scene, assets, w, h, ts)
        # tiny frames of random soups hold few pixels: a single texel flip of the +-1 LSB shading path is 0.1 %
        compare(g, o, f"fuzz seed {seed}", pixel_frac=0.99)
        # the pixels-only kernel variant (empty-tile path, sliced host output for large frames) writes the same bytes
        scene2, assets2, r2, _, _, _ = _scene(seed)   # a fresh scene: rasterize() appends the chunk lights on every call
        fast = render_gpu(r2, scene2, assets2, w, h, ts, planes=False)[0]
        assert np.array_equal(fast, g[0]), f"fuzz seed {seed}: pixels-only variant differs"


@pytest.mark.parametrize("seed", [20675, 20522])
def test_uv_of_triangles_crossing_the_near_plane(seed):
    """Regression: with fused multiply-adds in the deferred shade's perspective-correct UV sums these two scenes (a huge
    triangle crossing the near plane under a minified noise texture) had 32 % / 2.5 % of their pixels on another texel
    although owner and depth were exact.  The sums are the reference's unfused operations now: the full colour bar holds."""
    scene, assets, r, w, h, ts = _scene(seed)
    g = render_gpu(r, scene, assets, w, h, ts)
    o = render_oracle(r, scene, assets, w, h, ts)
    compare(g, o, f"fuzz seed {seed}", pixel_frac=0.999)
