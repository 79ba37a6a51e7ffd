"""CPU tests of the render-graph row: the host-side Sky node set-up (shapefx.rs:970-1120) and the
oracle's miss pass -- screen_ray, Sky render_miss_d3 without its cloud layer, brush preview -- against an independent
float32 restatement, plus the directional sun of the lighting block."""
import math

import numpy as np
import pytest

import oracle_ffi
from rusterix_b200 import scenes, vekmath
from rusterix_b200.types import BrushPreview, RenderGraph, SkyNode

F = np.float32


def test_sky_node_render_setup_known_answers():
    """shapefx.rs:970-1013: sun on a half circle between 06:00 and 20:00, smoothstep day factor over 6-8 h and 18-20 h."""
    s = SkyNode()
    sun, df = s.render_setup(13.0)              # t_day = 0.5 -> overhead
    assert df == 1.0 and abs(sun[0]) < 1e-6 and abs(sun[1] - 1.0) < 1e-6 and sun[2] == 0.0
    assert s.render_setup(3.0)[1] == 0.0 and s.render_setup(22.0)[1] == 0.0
    assert s.render_setup(7.0)[1] == pytest.approx(0.5)         # (1/2)^2 * (3 - 1)
    assert s.render_setup(19.0)[1] == pytest.approx(0.5)
    assert s.render_setup(6.0)[0] == pytest.approx((1.0, 0.0, 0.0), abs=1e-6)
    s.render_setup(12.0)
    np.testing.assert_allclose(s.precomputed[1], (0.3, 0.3, 0.35, 0.0), rtol=1e-6)   # day haze
    amb = s.render_ambient_color()                                                    # :1086-1120
    lin = [max(0.5 * a + 0.5 * b, 0.2) for a, b in zip(s.day_horizon[:3], s.day_zenith[:3])]
    np.testing.assert_allclose(amb[:3], [1.055 * c ** (1 / 2.4) - 0.055 for c in lin], rtol=1e-5)


def _mv(m, v):   # vek Mat4 * Vec4, the library's default convention (column accumulation with FMAs)
    r = np.zeros(4, np.float64)
    out = []
    for i in range(4):
        acc = F(F(m[i, 0]) * F(v[0]))
        for k in (1, 2, 3):
            acc = F(np.float64(F(m[i, k])) * np.float64(F(v[k])) + np.float64(acc))
        out.append(acc)
    return np.array(out, dtype=F)


def _ray(r, x, y, w, h):   # src/rasterizer.rs:1844-1870
    nx = F(F(F(2.0) * F(F(x) / F(w))) - F(1.0))
    ny = F(F(1.0) - F(F(2.0) * F(F(y) / F(h))))
    vn, vf = _mv(r.inverse_projection_matrix, [nx, ny, F(-1), F(1)]), _mv(r.inverse_projection_matrix, [nx, ny, F(1), F(1)])
    vn, vf = (vn / vn[3]).astype(F), (vf / vf[3]).astype(F)
    wn, wf = _mv(r.inverse_view_matrix, vn), _mv(r.inverse_view_matrix, vf)
    d = (wf[:3] - wn[:3]).astype(F)
    m = F(np.sqrt(F(F(F(d[0] * d[0]) + F(d[1] * d[1])) + F(d[2] * d[2]))))
    return wn[:3], (d / m).astype(F)


def _lerp(a, b, t):
    t = F(min(max(t, F(0)), F(1)))
    return F(np.float64(t) * np.float64(F(F(b) - F(a))) + np.float64(F(a)))


def _sky(pre, d):   # shapefx.rs:1122-1170
    sun, haze_c, day_h, day_z, night_h, night_z = [np.array(p, dtype=F) for p in pre]
    df = sun[3]
    up = F(min(max(d[1], F(-1)), F(1)))
    t = F(F(up + F(1)) * F(0.5))
    om = F(F(1) - up)
    haze = F(F(om * om) * om)
    c = []
    for i in range(4):
        v = _lerp(_lerp(night_h[i], night_z[i], t), _lerp(day_h[i], day_z[i], t), df)
        c.append(F(F(v * F(F(1) - F(haze * F(0.2)))) + F(F(haze_c[i] * haze) * F(0.3))))
    if df > 0:
        dt = F(F(F(d[0] * sun[0]) + F(d[1] * sun[1])) + F(d[2] * sun[2]))
        dist = F(max(F(F(1) - F(min(max(dt, F(-1)), F(1)))), F(0)))
        if dist < F(0.04):
            k = F(F(1) - F(dist / F(0.04)))
            g = F(F(k * k) * F(F(3) - F(F(2) * k)))
            for i, col in enumerate((1.0, 0.85, 0.6, 0.0)):
                c[i] = F(c[i] + F(F(F(col) * g) * df))
    return c


def _u8(x):
    x = F(min(max(x, F(0)), F(1))) if x == x else F(0)
    return int(F(np.float64(x) * 255.0 + 0.5))


def test_oracle_sky_miss_pixels_match_restatement():
    cfg = scenes.sky_config(96, 64, 32, hour=17.25)
    cfg.brush_preview = None
    cfg.scene.d3_static.clear()                  # every pixel is a miss
    r = cfg.rasterizer(2)
    px, ow, _ = oracle_ffi.rasterize(r, cfg.scene, cfg.assets, 96, 64, 32)
    assert (ow == 0xFFFFFFFF).all()
    pre = r.render_miss[-1].precomputed
    for (x, y) in [(0, 0), (95, 63), (48, 32), (10, 50), (80, 5), (33, 17)]:
        _, d = _ray(r, x, y, 96, 64)
        want = [_u8(c) for c in _sky(pre, d)]
        assert px[y, x].tolist() == want, (x, y)


def test_oracle_sun_glare_is_drawn_around_the_sun_direction():
    cfg = scenes.sky_config(128, 128, 64, hour=9.0)
    cfg.brush_preview = None
    cfg.scene.d3_static.clear()
    r = cfg.rasterizer(0)
    r.prepare_render_graph()
    sun = np.array(r.sun_dir, dtype=F)
    # aim the camera at the sun: the centre pixel gets the full glare (k ~ 1 -> + (1, .85, .6) * day_factor)
    from rusterix_b200 import D3FirstPCamera
    cam = D3FirstPCamera.new()
    cam.set_parameter_vec3("position", [0.0, 1.0, 0.0])
    cam.set_parameter_vec3("center", (np.array([0.0, 1.0, 0.0]) + sun).tolist())
    cfg.cameras, cfg.camera = None, cam
    r = cfg.rasterizer(0)
    px = oracle_ffi.rasterize(r, cfg.scene, cfg.assets, 128, 128, 64)[0]
    assert px[64, 64, 0] == 255 and px[64, 64, 1] >= 250
    assert px[5, 5, 0] < 200                      # away from the sun: the plain gradient


def test_oracle_brush_preview_blends_towards_white_on_the_ground_plane():
    cfg = scenes.sky_config(160, 90, 40)
    cfg.scene.d3_static.clear()
    with_brush = oracle_ffi.rasterize(cfg.rasterizer(1), cfg.scene, cfg.assets, 160, 90, 40)[0].astype(int)
    cfg.brush_preview = None
    without = oracle_ffi.rasterize(cfg.rasterizer(1), cfg.scene, cfg.assets, 160, 90, 40)[0].astype(int)
    changed = (with_brush != without).any(axis=-1)
    assert 0 < changed.sum() < 160 * 90 // 2
    assert (with_brush[changed][:, :3] >= without[changed][:, :3]).all()      # color*(1-b) + b >= color for color <= 1
    assert not changed[:20].any()                                              # rays above the horizon never hit y = 0


def test_oracle_brush_preview_whitens_terrain_texels_only():
    """rasterizer.rs:1193-1212: with a brush preview the terrain texel is blended towards white inside the brush radius
    before it is shaded.  Pixels owned by other batches keep their colour; terrain pixels change only near the brush."""
    from rusterix_b200 import marshal
    from rusterix_b200.types import BrushPreview

    cfg = scenes.chunked_config(320, 180)
    cfg.scene.d2_static.clear(); cfg.scene.d2_dynamic.clear()
    for ch in cfg.scene.chunks.values():
        ch.batches2d.clear(); ch.terrain_batch2d = None; ch.batches3d_opacity.clear()
    without, owner, depth = oracle_ffi.rasterize(cfg.rasterizer(0), cfg.scene, cfg.assets, 320, 180, 40)
    cfg.brush_preview = BrushPreview((8.0, 0.0, 8.0), 2.5, 0.5)
    with_brush, owner2, _ = oracle_ffi.rasterize(cfg.rasterizer(0), cfg.scene, cfg.assets, 320, 180, 40)
    assert np.array_equal(owner, owner2)                     # alpha is untouched: ownership does not move
    # owner id ranges of the terrain batches: 3 slots per triangle, in submission order
    b3, _b2 = marshal.submission_order(cfg.scene)
    terrain = np.zeros(owner.shape, dtype=bool)
    base = 0
    for entry in b3:
        batch = entry[0]
        n = 3 * len(batch.indices)
        if batch.source_.name == "Terrain":
            terrain |= (owner >= base) & (owner < base + n)
        base += n
    changed = (with_brush.astype(int) != without.astype(int)).any(axis=-1)
    assert changed.sum() > 50 and not changed[~terrain].any()
    assert (with_brush[changed][:, :3].astype(int) >= without[changed][:, :3].astype(int)).all()
    assert (terrain & ~changed).sum() > 50                   # terrain outside the radius is untouched


def test_sun_lights_the_scene_and_clouds_are_rejected():
    cfg = scenes.sky_config(120, 80, 40, hour=13.0)
    lit = oracle_ffi.rasterize(cfg.rasterizer(0), cfg.scene, cfg.assets, 120, 80, 40)
    cfg2 = scenes.sky_config(120, 80, 40, hour=23.0)                          # night: day_factor 0, no sun term
    dark = oracle_ffi.rasterize(cfg2.rasterizer(0), cfg2.scene, cfg2.assets, 120, 80, 40)
    covered = lit[1] != 0xFFFFFFFF
    assert np.array_equal(lit[1], dark[1])                                    # same geometry
    assert lit[0][covered][:, :3].astype(int).sum() > dark[0][covered][:, :3].astype(int).sum()
    cfg.render_graph = RenderGraph([SkyNode(clouds=True)])                    # the reference's cloud layer
    with pytest.raises(RuntimeError):
        oracle_ffi.rasterize(cfg.rasterizer(0), cfg.scene, cfg.assets, 120, 80, 40)
