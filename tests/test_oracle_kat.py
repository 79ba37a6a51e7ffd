"""Known-answer tests that pin the CPU oracle (oracle/rx_oracle.cpp) function by function.  The
reference ships no tests or fixtures for this path, so each expectation is
derived by hand / by an independent numpy-float32 restatement of the cited reference lines
(SURVEY 8c, k1..k10)."""
import ctypes as C
import math

import numpy as np
import pytest

import oracle_ffi
from rusterix_b200 import _abi
from rusterix_b200 import (Batch2D, Batch3D, CullMode, Light, LightType, PixelSource, Rasterizer, RenderMode, RepeatMode,
                           SampleMode, Scene, Texture, Tile, Assets, VGrayGradientShader, scenes)

f32 = np.float32
NONE = 0xFFFFFFFF


@pytest.fixture(scope="module")
def lib():
    return oracle_ffi.load()


def _arr(a, dt=np.float32):
    return np.ascontiguousarray(np.asarray(a, dtype=dt))


# ---- k1: Edges (src/edge.rs:12-36) ---------------------------------------------------------------
def test_k1_edges_sign_inclusive_and_nan(lib):
    v0 = _arr([[0, 0], [0, 4], [4, 0]])  # triangle (0,0) -> (0,4) -> (4,0), inside = all edges >= 0
    v1 = _arr([[0, 4], [4, 0], [0, 0]])
    abc = np.zeros(9, np.float32)
    lib.rxo_edges_new(v0.ctypes.data_as(C.c_void_p), v1.ctypes.data_as(C.c_void_p), abc.ctypes.data_as(C.c_void_p))
    assert abc.tolist() == [4, -4, 0, 0, -4, 4, 0, 16, 0]  # a = dy, b = -dx, c = x1*y0 - y1*x0

    def ev(x, y):
        return lib.rxo_edges_evaluate(abc.ctypes.data_as(C.c_void_p), x, y)

    assert ev(1, 1) == 1
    assert ev(0, 2) == 1 and ev(2, 0) == 1 and ev(2, 2) == 1  # all three edges are inclusive (no top-left rule)
    assert ev(0, 0) == 1 and ev(4, 0) == 1 and ev(0, 4) == 1  # vertices too
    assert ev(3, 2) == 0 and ev(-0.001, 1) == 0 and ev(1, -0.001) == 0
    assert ev(float("nan"), 1) == 1  # `result < 0.0` is false for NaN: NaN passes


# ---- k2: Texture::sample (src/texture.rs:203-232, 307-323, 414-460) ---------------------------------
def _ramp_texture(w=4, h=4):
    y, x = np.mgrid[0:h, 0:w]
    return np.stack([x * 60 + y, y * 60 + x, (x + y) * 30, 255 - x * 10 - y], axis=-1).astype(np.uint8)


def _ref_sample(img, u, v, sample_mode, repeat_mode):
    """Independent float32 restatement of the cited lines."""
    h, w = img.shape[:2]
    u, v = f32(u), f32(v)

    def clamp(x):
        return x if math.isnan(x) else f32(min(max(x, f32(0)), f32(1)))

    def wrap(x):
        return f32(x - f32(math.floor(x))) if not math.isnan(x) and not math.isinf(x) else f32(float("nan"))

    u = wrap(u) if repeat_mode in (1, 2) else clamp(u)
    v = wrap(v) if repeat_mode in (1, 3) else clamp(v)

    def rnd(x):  # f32::round, half away from zero; `as usize` saturates, NaN -> 0
        if math.isnan(x):
            return 0
        return max(0, int(math.floor(abs(x) + 0.5)) * (1 if x >= 0 else -1))

    if sample_mode == 0:
        tx = min(rnd(f32(u * f32(w - 1))), w - 1)
        ty = min(rnd(f32(v * f32(h - 1))), h - 1)
        return img[ty, tx].tolist()
    x = f32(u * f32(w - 1))
    y = f32(v * f32(h - 1))
    fx = 0.0 if math.isnan(x) else math.floor(x)
    fy = 0.0 if math.isnan(y) else math.floor(y)
    x0, y0 = max(0, int(fx)), max(0, int(fy))
    x1, y1 = min(x0 + 1, w - 1), min(y0 + 1, h - 1)
    dx, dy = f32(x - f32(math.floor(x))) if not math.isnan(x) else f32("nan"), f32(y - f32(math.floor(y))) if not math.isnan(y) else f32("nan")
    out = []
    for i in range(4):
        v00, v10, v01, v11 = (f32(img[y0, x0, i]), f32(img[y0, x1, i]), f32(img[y1, x0, i]), f32(img[y1, x1, i]))
        a = f32(v00 + f32(dx * f32(v10 - v00)))
        b = f32(v01 + f32(dx * f32(v11 - v01)))
        r = f32(a + f32(dy * f32(b - a)))
        out.append(0 if math.isnan(r) else min(255, max(0, int(math.floor(abs(r) + 0.5)))))
    return out


def test_k2_sampling_all_modes(lib):
    img = _ramp_texture()
    data = np.ascontiguousarray(img.reshape(-1))
    t = _abi.rxc_texture(data.ctypes.data, 4, 4)
    us = [0.0, 0.125, 1 / 3, 0.5, 0.49999, 0.875, 1.0, -1e-7, -0.3, 1.0 + 1e-6, 1.7, 2.0, -2.25, 0.1666667, 0.8333333]
    out = np.zeros(4, np.uint8)
    for repeat in range(4):
        for mode in range(2):
            for u in us:
                for v in (0.0, 0.4, 1.0, -0.2, 1.3):
                    lib.rxo_texture_sample(C.byref(t), C.c_float(u), C.c_float(v), mode, repeat, out.ctypes.data_as(C.c_void_p))
                    assert out.tolist() == _ref_sample(img, u, v, mode, repeat), (repeat, mode, u, v)
    # hand-checked anchors: nearest rounds half away from zero; linear clamps x1 at the edge even when repeating
    lib.rxo_texture_sample(C.byref(t), C.c_float(0.5), C.c_float(0.0), 0, 0, out.ctypes.data_as(C.c_void_p))
    assert out.tolist() == img[0, 2].tolist()  # 0.5*3 = 1.5 -> 2
    lib.rxo_texture_sample(C.byref(t), C.c_float(1.0), C.c_float(1.0), 1, 1, out.ctypes.data_as(C.c_void_p))
    assert out.tolist() == img[0, 0].tolist()  # RepeatXY: 1.0 - floor(1.0) = 0
    lib.rxo_texture_sample(C.byref(t), C.c_float(float("nan")), C.c_float(0.0), 0, 0, out.ctypes.data_as(C.c_void_p))
    assert out.tolist() == img[0, 0].tolist()  # NaN as usize == 0


# ---- k3: pixel conversions (src/lib.rs:55-79) --------------------------------------------------------
def test_k3_vec4_to_pixel_rounding(lib):
    out = np.zeros(4, np.uint8)

    def conv(x):
        v = _arr([x, x, x, x])
        lib.rxo_vec4_to_pixel(v.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
        return int(out[0])

    for k in range(256):
        assert conv(k / 255.0) == k
    assert conv(0.5) == 128          # 0.5*255 + 0.5 = 128.0 exactly (fused), truncated
    assert conv(-3.0) == 0 and conv(7.0) == 255 and conv(float("nan")) == 0  # max(0).min(1); NaN.max(0) = 0
    assert conv(np.nextafter(f32(0.5), f32(0))) == 127
    p = np.array([0, 51, 255, 128], np.uint8)
    v = np.zeros(4, np.float32)
    lib.rxo_pixel_to_vec4(p.ctypes.data_as(C.c_void_p), v.ctypes.data_as(C.c_void_p))
    inv = f32(1.0) / f32(255.0)
    assert v.tolist() == [f32(0) * inv, f32(51) * inv, f32(255) * inv, f32(128) * inv]  # multiply by INV_255, not divide


def test_hash_u32(lib):
    def ref(seed):
        m = 0xFFFFFFFF
        s = (seed ^ 61) ^ (seed >> 16)
        s = (s + (s << 3)) & m
        s ^= s >> 4
        s = (s * 0x27D4EB2D) & m
        s ^= s >> 15
        return s

    for seed in (0, 1, 2, 61, 12345, 0xFFFFFFFF, 0x80000000):
        assert lib.rxo_hash_u32(seed) == ref(seed)


# ---- k4 / k5: clip_and_project (src/batch/batch3d.rs:482-740) ----------------------------------------
def _identity_rast(w=100, h=100):
    return Rasterizer.setup(None, np.eye(4, dtype=np.float32), np.eye(4, dtype=np.float32))


def _one_tri_scene(verts, cull=CullMode.Off, idx=((0, 1, 2),)):
    v = [(x, y, z, 1.0) for x, y, z in verts]
    b = Batch3D(v, idx, [(0, 0), (1, 0), (0, 1)][: len(v)] + [(0, 0)] * max(0, len(v) - 3)).cull_mode(cull)
    b.with_normals(np.tile([0, 0, 1], (len(v), 1)))
    return Scene.from_static([], [b])


def test_k4_near_clip_slots_and_interpolation():
    # v2 is behind the z = -0.1 plane: quad (v0, v1, I12, I20) -> two fan triangles appended after the originals
    sc = _one_tri_scene([(0, 0, -0.5), (1, 0, -0.5), (0, 0, 0.3)])
    out = oracle_ffi.clip_and_project(_identity_rast(), sc, 0, 100, 100)
    assert out["clipped_indices"].tolist() == [[0, 1, 2], [3, 4, 5], [3, 5, 6]]
    assert out["visible"].tolist()[0] == 0  # the original keeps its slot but is not drawn (T-clipvis)
    assert len(out["projected"]) == 7
    near = f32(0.1)
    t12 = f32(f32(-near - f32(-0.5)) / f32(f32(0.3) - f32(-0.5)))
    t20 = f32(f32(-near - f32(0.3)) / f32(f32(-0.5) - f32(0.3)))
    assert abs(t12 - 0.5) < 1e-6 and abs(t20 - 0.5) < 1e-6
    i12 = [f32(1) + t12 * (f32(0) - f32(1)), f32(0), f32(-0.5) + t12 * (f32(0.3) - f32(-0.5))]
    # identity projection: sx = (x*0.5+0.5)*W, sy = (-y*0.5+0.5)*H, z, w=1
    assert out["projected"][5].tolist() == [f32((f32(i12[0]) * f32(0.5) + f32(0.5)) * f32(100)), 50.0, f32(i12[2]), 1.0]
    assert out["projected"][3].tolist() == out["projected"][0].tolist()  # inside vertices are re-appended
    assert abs(out["projected"][6][2] - (-0.1)) < 1e-6

    # two vertices behind -> a single triangle; all behind -> nothing appended, slot invisible
    sc = _one_tri_scene([(0, 0, -0.5), (1, 0, 0.3), (0, 1, 0.3)])
    out = oracle_ffi.clip_and_project(_identity_rast(), sc, 0, 100, 100)
    assert out["clipped_indices"].tolist() == [[0, 1, 2], [3, 4, 5]] and out["visible"].tolist() == [0, 1]
    sc = _one_tri_scene([(0, 0, 0.5), (1, 0, 0.3), (0, 1, 0.3)])
    out = oracle_ffi.clip_and_project(_identity_rast(), sc, 0, 100, 100)
    assert out["clipped_indices"].tolist() == [[0, 1, 2]] and out["visible"].tolist() == [0]


@pytest.mark.parametrize("cull,ccw_visible,cw_visible", [(CullMode.Off, 1, 1), (CullMode.Front, 1, 0), (CullMode.Back, 0, 1)])
def test_k5_cull_swap_truth_table(cull, ccw_visible, cw_visible):
    """Screen space has y down: a world-space CCW triangle has screen orientation < 0 ("not front")."""
    ccw = [(-0.5, -0.5, -0.5), (0.5, -0.5, -0.5), (0.0, 0.5, -0.5)]
    cw = [ccw[0], ccw[2], ccw[1]]
    for verts, expect in ((ccw, ccw_visible), (cw, cw_visible)):
        out = oracle_ffi.clip_and_project(_identity_rast(), _one_tri_scene(verts, cull), 0, 100, 100)
        assert out["visible"].tolist() == [expect]
        if expect:  # after the swap every edge function is >= 0 at the centroid
            e = out["edges"][0]
            cx, cy = out["projected"][:3, 0].mean(), out["projected"][:3, 1].mean()
            for i in range(3):
                assert e[i] * cx + e[3 + i] * cy + e[6 + i] >= 0
    bbox = out["bbox"]
    assert bbox.tolist() == [25.0, 25.0, 50.0, 50.0]  # Rect{x, y, width = max-min, height}


def test_frustum_aabb_early_out_clears_batch():
    sc = _one_tri_scene([(5, 5, -0.5), (6, 5, -0.5), (5, 6, -0.5)])  # x > w for all corners
    out = oracle_ffi.clip_and_project(_identity_rast(), sc, 0, 100, 100)
    assert out["bbox"] is None and len(out["edges"]) == 0 and len(out["projected"]) == 0


# ---- k6: CompiledLight (src/map/light.rs:491-677) ------------------------------------------------------
def _light(**kw):
    l = Light.new(kw.pop("t", LightType.Point))
    c = l.compile()
    for k, v in kw.items():
        setattr(c, k, v)
    from rusterix_b200 import marshal

    return marshal.marshal_lights([c]).struct


def _color_at(lib, ls, p, hash_=0, d2=False):
    out = np.zeros(3, np.float32)
    pt = _arr(p)
    ok = lib.rxo_light_color_at(ls, pt.ctypes.data_as(C.c_void_p), hash_, 1 if d2 else 0, out.ctypes.data_as(C.c_void_p))
    return ok, out.tolist()


def test_k6_light_types(lib):
    ls = _light(position=(0, 0, 0), color=(1.0, 0.5, 0.25), intensity=2.0, start_distance=1.0, end_distance=2.0)
    assert _color_at(lib, ls, (0.5, 0, 0)) == (1, [2.0, 1.0, 0.5])               # inside start: full
    assert _color_at(lib, ls, (1.5, 0, 0)) == (1, [1.0, 0.5, 0.25])              # smoothstep(end,start,1.5) = 0.5
    assert _color_at(lib, ls, (2.0, 0, 0))[0] == 0 and _color_at(lib, ls, (3, 0, 0))[0] == 0  # >= end: None
    ok, c = _color_at(lib, ls, (1.25, 0, 0))
    t = 0.75
    assert ok and abs(c[0] - 2.0 * t * t * (3 - 2 * t)) < 1e-6
    assert _color_at(lib, _light(emitting=False), (0, 0, 0))[0] == 0

    # spot: linear falloff, cone test acos(dir . to_point) > cone_angle -> None
    sp = _light(t=LightType.Spot, position=(0, 0, 0), direction=(0.0, 0.0, -1.0), cone_angle=math.pi / 4, start_distance=1.0,
                end_distance=3.0, intensity=1.0)
    assert _color_at(lib, sp, (0, 0, -2.0)) == (1, [0.5, 0.5, 0.5])
    assert _color_at(lib, sp, (0.5, 0, -1.0))[0] == 1 and _color_at(lib, sp, (1.2, 0, -1.0))[0] == 0
    # ambient ignores distance
    assert _color_at(lib, _light(t=LightType.Ambient, color=(0.2, 0.3, 0.4), intensity=0.5), (99, 99, 99)) == (1, [f32(0.2) * f32(0.5), f32(0.3) * f32(0.5), f32(0.4) * f32(0.5)])
    # area: < 0.1 returns the raw colour; otherwise (normal . dir)+ * smoothstep * w*h * intensity
    ar = _light(t=LightType.Area, position=(0, 1, 0), normal=(0.0, -1.0, 0.0), width=2.0, height=3.0, start_distance=1.0, end_distance=4.0,
                intensity=0.5, color=(1.0, 1.0, 1.0))
    assert _color_at(lib, ar, (0, 0.95, 0)) == (1, [1.0, 1.0, 1.0])
    ok, c = _color_at(lib, ar, (0, 0.5, 0))
    assert ok and abs(c[0] - 1.0 * 1.0 * 6.0 * 0.5) < 1e-6
    ok, c = _color_at(lib, ar, (0, 1.5, 0))
    assert ok and c[0] == 0.0  # behind the emitting side
    # daylight
    dl = _light(t=LightType.Daylight, position=(0, 5, 0), normal=(0.0, -1.0, 0.0), start_distance=1.0, end_distance=20.0, intensity=0.4)
    ok, c = _color_at(lib, dl, (0, 0, 0))
    t = (5.0 - 20.0) / (1.0 - 20.0)
    assert ok and abs(c[0] - 0.4 * t * t * (3 - 2 * t)) < 1e-6


def test_k6_flicker_hash(lib):
    ls = _light(position=(3.7, 2.2, 9.9), flicker=0.5, intensity=1.0, start_distance=5.0, end_distance=9.0, color=(1.0, 1.0, 1.0))
    h = lib.rxo_hash_u32(77)
    combined = (h + (3 + 2 + 9) * 100) & 0xFFFFFFFF
    fv = min(max(f32(combined) / f32(4294967296.0), f32(0)), f32(1))
    expect = f32(1.0) - f32(fv * f32(0.5))
    ok, c = _color_at(lib, ls, (3.7, 2.2, 9.0), hash_=h)
    assert ok and c[0] == f32(f32(1.0) * f32(1.0)) * expect


def test_radiance_applies_lambert_for_positional_lights_only(lib):
    ls = _light(position=(0, 2, 0), intensity=1.0, start_distance=5.0, end_distance=9.0)
    out = np.zeros(3, np.float32)
    p, n = _arr([0, 0, 0]), _arr([0, 1, 0])
    assert lib.rxo_light_radiance_at(ls, p.ctypes.data_as(C.c_void_p), n.ctypes.data_as(C.c_void_p), 0, out.ctypes.data_as(C.c_void_p)) == 1
    assert out.tolist() == [1.0, 1.0, 1.0]
    n = _arr([1, 0, 0])
    lib.rxo_light_radiance_at(ls, p.ctypes.data_as(C.c_void_p), n.ctypes.data_as(C.c_void_p), 0, out.ctypes.data_as(C.c_void_p))
    assert out.tolist() == [0.0, 0.0, 0.0]
    amb = _light(t=LightType.Ambient, intensity=1.0)
    lib.rxo_light_radiance_at(amb, p.ctypes.data_as(C.c_void_p), n.ctypes.data_as(C.c_void_p), 0, out.ctypes.data_as(C.c_void_p))
    assert out.tolist() == [1.0, 1.0, 1.0]


# ---- whole-frame behaviours: k7..k10 -----------------------------------------------------------------
def _quad_scene(z0=-0.5, z1=-0.5):
    """Two triangles sharing the diagonal of a screen-aligned quad (identity matrices)."""
    v = [(-0.5, -0.5, z0, 1.0), (0.5, -0.5, z0, 1.0), (0.5, 0.5, z1, 1.0), (-0.5, 0.5, z1, 1.0)]
    b = Batch3D(v, [(0, 1, 2), (0, 2, 3)], [(0, 0), (1, 0), (1, 1), (0, 1)]).source(PixelSource.Pixel([200, 100, 50, 255]))
    b.with_normals(np.tile([0, 0, 1], (4, 1)))
    return Scene.from_static([], [b])


def test_k7_shared_edge_ownership():
    """Both triangles cover the pixel centres on the shared diagonal (inclusive edges, no top-left
    rule); strict `z < zbuf` lets the first-drawn triangle keep exact ties."""
    sc = _quad_scene()
    px, owner, depth = oracle_ffi.rasterize(_identity_rast(), sc, Assets(), 64, 64, 16)
    inside = owner != NONE
    assert inside.sum() == 32 * 32  # pixel centres 16.5 .. 47.5
    assert set(np.unique(owner[inside]).tolist()) == {0, 1}
    diag = [(63 - i, i) for i in range(16, 48)]  # screen y is flipped: the diagonal runs (16,47)..(47,16)
    on_diag = np.array([owner[y, x] for x, y in [(i, 63 - i) for i in range(16, 48)]])
    assert (on_diag == 0).all()  # constant z: ties, first drawn wins
    assert np.all(depth[inside] == f32(-0.5))
    # with a depth gradient the two barycentric evaluations may differ by an ulp on the diagonal:
    # record the oracle's answer as the golden
    sc = _quad_scene(-0.3, -0.9)
    px, owner2, depth2 = oracle_ffi.rasterize(_identity_rast(), sc, Assets(), 64, 64, 16)
    on_diag2 = np.array([owner2[63 - i, i] for i in range(16, 48)])
    golden = np.load(oracle_ffi.ROOT + "/tests/golden/k7_diag_owner.npy")
    assert np.array_equal(on_diag2, golden)


def test_k8_tile_size_invariance():
    cfg = scenes.cube(200, 150, 40, logo_size=32)
    r = cfg.rasterizer()
    ref = oracle_ffi.rasterize(r, cfg.scene, cfg.assets, 200, 150, 8)
    for ts in (13, 40, 200, 4096):
        out = oracle_ffi.rasterize(r, cfg.scene, cfg.assets, 200, 150, ts)
        for a, b in zip(ref, out):
            assert np.array_equal(a, b)


def test_k9_miss_pixels_overwrite_background_shader():
    cfg = scenes.cube(120, 90, 40, logo_size=32)
    assert isinstance(cfg.scene.background, VGrayGradientShader)
    r = cfg.rasterizer().background([9, 9, 9, 255])
    px, owner, _ = oracle_ffi.rasterize(r, cfg.scene, cfg.assets, 120, 90, 40)
    miss = owner == NONE
    miss[:, :] &= True
    outside_rect = np.ones_like(miss)
    assert (px[miss][:, :3] == 0).all() and (px[miss][:, 3] == 255).all()  # src/rasterizer.rs:409-461
    # in 2D-only mode the gradient survives
    r2 = cfg.rasterizer().render_mode(RenderMode.render_2d())
    px2, _, _ = oracle_ffi.rasterize(r2, cfg.scene, cfg.assets, 120, 90, 40)
    assert px2[89, 100].tolist() == [int(89 / 90 * 128), int(89 / 90 * 128), int(89 / 90 * 128), 255]


def test_k10_alpha_tested_fragment_does_not_write_depth():
    """A fence (alpha 0 holes) in front of an opaque wall: through the holes the depth plane holds the
    wall's z, on the bars the fence's (src/rasterizer.rs:1408-1412)."""
    fence = np.zeros((4, 4, 4), np.uint8)
    fence[:, :2] = [255, 255, 255, 255]  # left half opaque, right half alpha 0
    wall = np.full((2, 2, 4), 255, np.uint8)
    assets = Assets().textures([Tile.from_texture(Texture.from_array(fence)), Tile.from_texture(Texture.from_array(wall))])

    def quad(z, tile):
        v = [(-0.8, -0.8, z, 1.0), (0.8, -0.8, z, 1.0), (0.8, 0.8, z, 1.0), (-0.8, 0.8, z, 1.0)]
        b = Batch3D(v, [(0, 1, 2), (0, 2, 3)], [(0, 0), (1, 0), (1, 1), (0, 1)]).source(PixelSource.StaticTileIndex(tile))
        return b.with_normals(np.tile([0, 0, 1], (4, 1)))

    sc = Scene.from_static([], [quad(-0.3, 0), quad(-0.6, 1)])  # the fence (z=-0.3 ... ndc -0.3) is drawn first
    px, owner, depth = oracle_ffi.rasterize(_identity_rast(), sc, assets, 80, 80, 40)
    # identity projection: z_ndc = z; smaller is nearer, so the wall at -0.6 is in FRONT numerically.
    # put the fence in front instead:
    sc = Scene.from_static([], [quad(-0.6, 0), quad(-0.3, 1)])
    px, owner, depth = oracle_ffi.rasterize(_identity_rast(), sc, assets, 80, 80, 40)
    left, right = (40, 20), (40, 60)  # (y, x): u < 0.5 opaque bars, u > 0.5 holes
    assert owner[left] in (0, 1) and depth[left] == f32(-0.6)
    assert owner[right] in (6, 7) and depth[right] == f32(-0.3)


def test_oracle_rejects_what_the_product_rejects():
    cfg = scenes.cube(32, 32, 16, logo_size=8)
    cfg.scene.d3_static[0].source(PixelSource.StaticTileIndex(5))   # past assets.tile_list: the reference panics
    with pytest.raises(RuntimeError):
        oracle_ffi.rasterize(cfg.rasterizer(), cfg.scene, cfg.assets, 32, 32, 16)
