"""GPU parity tests proper: the CUDA path, called through the C ABI, against the CPU oracle on the
same seeded inputs.  Run on the B200 box with `pytest -m gpu`."""
import math

import numpy as np
import pytest

from helpers import compare, render_gpu, render_oracle
from rusterix_b200 import (Batch2D, Batch3D, CullMode, Light, LightType, MatVecMode, PixelSource, Rasterizer, RenderMode,
                           RepeatMode, SampleMode, Scene, scenes)
from rusterix_b200 import GridShader, VGrayGradientShader

pytestmark = pytest.mark.gpu


def _run(cfg, frame=0, what=None, **kw):
    r = cfg.rasterizer(frame)
    for k, v in kw.items():
        setattr(r, k, v)
    g = render_gpu(r, cfg.scene, cfg.assets, cfg.width, cfg.height, cfg.tile_size)
    o = render_oracle(r, cfg.scene, cfg.assets, cfg.width, cfg.height, cfg.tile_size)
    return compare(g, o, what or cfg.name)


def test_cube_800x600_nearest():
    st = _run(scenes.cube(800, 600, 200, logo_size=256))
    assert st["exact_frac"] > 0.99


# ---- the BASELINE.json configs at their full sizes on the reference's own mesh and textures (tests/golden/reference_assets)
def test_config_a_cube_800x600_tile200_reference_logo():
    """Config A as the north star words it: 800x600, tile 200, Nearest, images/logo.png at 1024x1024."""
    st = _run(scenes.cube(800, 600, 200, real=True))
    assert st["exact_frac"] > 0.99


def test_config_a_cube_2000x2000_tile40_as_the_bench_file():
    """Config A in the shape of the reference's own bench (benches/rasterize_cube.rs:13-14,30): 2000x2000, tile 40."""
    st = _run(scenes.cube(2000, 2000, 40, real=True))
    assert st["exact_frac"] > 0.99


@pytest.mark.parametrize("frame", [0, 21, 47])
def test_config_b_teapot_obj_1920x1080_linear(frame):
    """Config B on examples/teapot.obj itself (1202 vertices, 2256 triangles, UV = (x, y), RepeatXY,
    scaling(.35, -.35, .35), bilinear sampling of images/logo.png) at 1920x1080, orbit camera."""
    cfg = scenes.teapot(1920, 1080, 60, real=True)
    assert cfg.counts() == (1202, 2256)
    _run(cfg, frame=frame)


def test_config_c_map_4k_reference_textures():
    """Config C with the minigame's own PNGs (brickwall, lightpanel, fence with its alpha holes, brickfloor, sky)."""
    st = _run(scenes.map_config(3840, 2160, 40, real=True))
    assert st["within1_frac"] >= 0.999


# ---- the pre-projected entry (rxc_rasterize_projected): the host's own Scene::project results, used verbatim --------------
def _run_projected(cfg, frame=0, host_mode=None, **kw):
    import oracle_ffi

    r = cfg.rasterizer(frame)
    for k, v in kw.items():
        setattr(r, k, v)
    host = cfg.rasterizer(frame)          # the "host" that ran Scene::project, possibly with another Mat4*Vec4 rounding
    if host_mode is not None:
        host.matvec_mode = host_mode
    proj = oracle_ffi.project_scene(host, cfg.scene, cfg.assets, cfg.width, cfg.height, cfg.tile_size)
    px = np.zeros((cfg.height, cfg.width, 4), dtype=np.uint8)
    ow = np.zeros((cfg.height, cfg.width), dtype=np.uint32)
    dp = np.zeros((cfg.height, cfg.width), dtype=np.float32)
    r.rasterize_projected(cfg.scene, proj, px, cfg.width, cfg.height, cfg.tile_size, cfg.assets, owner=ow, depth=dp)
    o = render_oracle(host, cfg.scene, cfg.assets, cfg.width, cfg.height, cfg.tile_size)
    return compare((px, ow, dp), o, cfg.name + " (pre-projected)"), (px, ow, dp)


@pytest.mark.parametrize("name", ["cube", "teapot", "map", "sweep", "dense", "chunked"])
def test_preprojected_entry_reproduces_the_oracle_from_its_own_projection(name):
    """Feed the device what the oracle's clip_and_project produced (projected_vertices, clipped_*, edges, bounding_box):
    the frame must equal the oracle's own, owner and depth bit for bit -- and the frame the device renders when it
    projects itself."""
    cfg = {"cube": lambda: scenes.cube(800, 600, 200, logo_size=256), "teapot": lambda: scenes.teapot(960, 540, 60, real=True),
           "map": lambda: scenes.map_config(960, 540, 40, logo_size=256), "sweep": lambda: scenes.sweep(640, 360, 40, logo_size=128),
           "dense": lambda: scenes.dense(1280, 720, 40, patches=8), "chunked": lambda: scenes.chunked_config(960, 540)}[name]()
    frame = 1024 if name == "sweep" else 3 if name in ("teapot", "chunked") else 0   # the sweep's walls cross the near plane
    st, got = _run_projected(cfg, frame)
    own = render_gpu(cfg.rasterizer(frame), cfg.scene, cfg.assets, cfg.width, cfg.height, cfg.tile_size)
    assert np.array_equal(got[1], own[1]) and np.array_equal(got[2].view(np.uint32), own[2].view(np.uint32))
    if not cfg.scene.chunks:   # every rasterize() call appends the chunk lights again (rasterizer.rs:219-223): colours drift by design
        assert np.array_equal(got[0], own[0])


def test_preprojected_entry_follows_the_hosts_vertex_bits_not_its_own_convention():
    """The point of the entry: the host projected with ANOTHER Mat4*Vec4 rounding (plain rows) than the device would
    have used (fused columns, the frame's matvec_mode).  Ownership must follow the host's bits."""
    cfg = scenes.teapot(960, 540, 60, real=True)
    a = render_oracle(cfg.rasterizer(7), cfg.scene, cfg.assets, cfg.width, cfg.height, cfg.tile_size)
    host = cfg.rasterizer(7); host.matvec_mode = MatVecMode.PlainRows
    b = render_oracle(host, cfg.scene, cfg.assets, cfg.width, cfg.height, cfg.tile_size)
    assert (a[2].view(np.uint32) != b[2].view(np.uint32)).any()      # the two conventions do differ on this frame
    _run_projected(cfg, 7, host_mode=MatVecMode.PlainRows)


def test_preprojected_entry_rejects_what_the_reference_would_panic_on():
    import oracle_ffi
    from rusterix_b200 import RxcError

    cfg = scenes.cube(320, 240, 40, logo_size=64)
    r = cfg.rasterizer(0)
    proj = oracle_ffi.project_scene(r, cfg.scene, cfg.assets, 320, 240, 40)
    px = np.zeros((240, 320, 4), dtype=np.uint8)
    with pytest.raises(RxcError):
        r.rasterize_projected(cfg.scene, proj + proj, px, 320, 240, 40, cfg.assets)       # one entry per 3D batch
    bad = [dict(p) for p in proj]
    bad[0]["clipped_indices"] = bad[0]["clipped_indices"].copy(); bad[0]["clipped_indices"][0, 0] = 10 ** 6
    with pytest.raises(RxcError):
        r.rasterize_projected(cfg.scene, bad, px, 320, 240, 40, cfg.assets)
    none = [dict(p, bounding_box=None) for p in proj]                                      # bounding_box None: batch skipped
    r.rasterize_projected(cfg.scene, none, px, 320, 240, 40, cfg.assets)
    ow = np.zeros((240, 320), dtype=np.uint32)
    r.rasterize_projected(cfg.scene, none, px, 320, 240, 40, cfg.assets, owner=ow)
    assert (ow == 0xFFFFFFFF).all()


def test_cube_odd_size_partial_tiles():
    _run(scenes.cube(333, 217, 40, logo_size=128))


def test_cube_plain_rows_matvec():
    _run(scenes.cube(400, 300, 60, logo_size=128), matvec_mode=MatVecMode.PlainRows)


@pytest.mark.parametrize("frame", [0, 5, 17, 40])
def test_teapot_linear_orbit(frame):
    _run(scenes.teapot(960, 540, 60, logo_size=256), frame=frame)


def test_map_nearest_alpha_fence():
    st = _run(scenes.map_config(960, 540, 40, logo_size=256))
    assert st["exact_frac"] > 0.99


def test_map_full_4k_owner_and_depth():
    cfg = scenes.map_config(3840, 2160, 40, logo_size=256)
    _run(cfg)


@pytest.mark.parametrize("i", [0, 100, 1024, 3000])
def test_sweep_frames(i):
    _run(scenes.sweep(640, 360, 40, logo_size=128), frame=i)


def test_dense_small():
    _run(scenes.dense(1280, 720, 40, patches=8))


def test_near_clip_camera_inside_geometry():
    """Camera close to a wall so triangles cross the z=-0.1 plane (SURVEY k4, T-clipvis)."""
    cfg = scenes.map_config(640, 360, 40, logo_size=64)
    cam = scenes._firstp([0.3, 1.0, 7.5], [5.0, 0.8, 7.7])
    cfg.camera = cam
    st = _run(cfg, what="near-clip")
    cam = scenes._firstp([7.5, 0.05, 7.5], [9.0, 0.0, 9.0])
    cfg.camera = cam
    _run(cfg, what="near-clip-floor")


@pytest.mark.parametrize("cull", [CullMode.Off, CullMode.Front, CullMode.Back])
def test_cull_modes(cull):
    cfg = scenes.cube(320, 240, 40, logo_size=64)
    cfg.scene.d3_static[0].cull_mode(cull)
    cfg.scene.mark_dirty()
    _run(cfg, what=f"cull-{cull.name}")


@pytest.mark.parametrize("repeat", list(RepeatMode))
@pytest.mark.parametrize("sample", list(SampleMode))
def test_repeat_and_sample_modes(repeat, sample):
    cfg = scenes.teapot(320, 240, 60, logo_size=64)
    cfg.scene.d3_static[0].repeat_mode(repeat)
    cfg.scene.mark_dirty()
    cfg.sample_mode = sample
    _run(cfg, what=f"{repeat.name}-{sample.name}")


def test_tile_size_invariance_and_scissor():
    cfg = scenes.cube(400, 300, 200, logo_size=64)
    r = cfg.rasterizer()
    ref = None
    for ts in (8, 40, 200, 1000):
        g = render_gpu(r, cfg.scene, cfg.assets, cfg.width, cfg.height, ts)
        o = render_oracle(r, cfg.scene, cfg.assets, cfg.width, cfg.height, ts)
        compare(g, o, f"tile_size {ts}")
        if ref is None:
            ref = g
        else:
            assert np.array_equal(ref[0], g[0]) and np.array_equal(ref[1], g[1])


def test_light_types():
    cfg = scenes.map_config(480, 270, 40, logo_size=64)
    L = []
    L.append(Light.new(LightType.Point).with_position([9, 0.5, 12]).with_intensity(2).with_start_distance(2).with_end_distance(13).with_flicker(0.6).compile())
    L.append(Light.new(LightType.Spot).with_position([4, 1.8, 8]).with_direction([0.2, -1.0, 0.5]).with_cone_angle(0.6).with_end_distance(9).compile())
    L.append(Light.new(LightType.Area).with_position([7, 1.9, 11]).with_normal([0, -1, 0]).with_size(2, 2).with_end_distance(8).with_intensity(0.7).compile())
    L.append(Light.new(LightType.Ambient).with_color([0.1, 0.05, 0.0]).compile())
    L.append(Light.new(LightType.AmbientDaylight).with_color([0.0, 0.05, 0.1]).compile())
    L.append(Light.new(LightType.Daylight).with_position([7, 6, 7]).with_normal([0, -1, 0]).with_end_distance(30).with_intensity(0.4).compile())
    L.append(Light.new(LightType.Point).with_position([1, 1, 1]).with_emitting(False).compile())
    cfg.scene.lights = L
    cfg.scene.animation_frame = 12345
    cfg.ambient = (0.2, 0.2, 0.25, 1.0)
    _run(cfg, what="light types")


def test_2d_only_backgrounds():
    for bg in (VGrayGradientShader(), GridShader(), GridShader(grid_size=17.0, subdivisions=3.0, offset=(5.0, -3.0)), None):
        cfg = scenes.cube(300, 200, 40, logo_size=64)
        cfg.scene.background = bg
        cfg.scene.d2_static = [Batch2D.from_rectangle(20.0, 30.0, 120.0, 90.0).source(PixelSource.StaticTileIndex(0)),
                               Batch2D.from_rectangle(100.0, 60.0, 150.0, 100.0).source(PixelSource.Pixel([200, 40, 40, 128]))]
        cfg.scene.mark_dirty()
        r = cfg.rasterizer().render_mode(RenderMode.render_2d()).background([10, 20, 30, 255])
        g = render_gpu(r, cfg.scene, cfg.assets, cfg.width, cfg.height, cfg.tile_size, planes=False)
        o = render_oracle(r, cfg.scene, cfg.assets, cfg.width, cfg.height, cfg.tile_size, planes=False)
        st = compare(g, o, f"2D bg {bg}", pixel_frac=1.0)
        assert st["max_diff"] == 0


def test_2d_matrix_lights_and_transparency():
    cfg = scenes.cube(320, 200, 40, logo_size=64)
    cfg.scene.d2_static = [Batch2D.from_rectangle(0.0, 0.0, 100.0, 60.0).source(PixelSource.StaticTileIndex(0)),
                           Batch2D.from_rectangle(30.0, 20.0, 80.0, 50.0).source(PixelSource.Pixel([255, 255, 0, 90])).receives_light(False)]
    cfg.scene.lights = [Light.new(LightType.Point).with_position([40, 0, 30]).with_start_distance(10).with_end_distance(90).with_intensity(1.3).compile()]
    cfg.scene.mark_dirty()
    m = np.array([[2.0, 0.0, 15.0], [0.0, 2.0, 9.0], [0.0, 0.0, 1.0]], dtype=np.float32)
    for preserve in (False, True):
        r = Rasterizer.setup(m, np.eye(4), np.eye(4)).render_mode(RenderMode.render_2d())
        r.preserve_transparency = preserve
        g = render_gpu(r, cfg.scene, cfg.assets, cfg.width, cfg.height, cfg.tile_size, planes=False)
        o = render_oracle(r, cfg.scene, cfg.assets, cfg.width, cfg.height, cfg.tile_size, planes=False)
        st = compare(g, o, "2D matrix+lights", pixel_frac=1.0)
        assert st["max_diff"] <= 1


def test_usize_indices_and_no_normals():
    cfg = scenes.cube(320, 240, 40, logo_size=64)
    cfg.scene.d3_static[0].normals = np.zeros((0, 3), dtype=np.float32)
    cfg.scene.mark_dirty()
    r = cfg.rasterizer()
    r.index_bytes = 8
    g = render_gpu(r, cfg.scene, cfg.assets, cfg.width, cfg.height, cfg.tile_size)
    o = render_oracle(r, cfg.scene, cfg.assets, cfg.width, cfg.height, cfg.tile_size, index_bytes=8)
    compare(g, o, "usize indices, no normals")


def test_band_rendering_matches_full_frame():
    cfg = scenes.map_config(640, 360, 40, logo_size=64)
    r = cfg.rasterizer()
    full = render_gpu(r, cfg.scene, cfg.assets, cfg.width, cfg.height, cfg.tile_size)
    for (y0, y1) in ((0, 96), (96, 200), (200, 360)):
        band = render_gpu(r, cfg.scene, cfg.assets, cfg.width, cfg.height, cfg.tile_size, band=(y0, y1))
        assert np.array_equal(band[0], full[0][y0:y1])
        assert np.array_equal(band[1], full[1][y0:y1])


def test_batch_api_matches_single_frames():
    cfg = scenes.sweep(320, 192, 40, n_frames=64, logo_size=64)
    frames = [0, 7, 31, 50]
    rs = [cfg.rasterizer(i) for i in frames]
    out = np.zeros((len(frames), cfg.height, cfg.width, 4), dtype=np.uint8)
    Rasterizer.rasterize_batch(rs, cfg.scene, out, cfg.width, cfg.height, cfg.tile_size, cfg.assets)
    for k, r in enumerate(rs):
        single = render_gpu(r, cfg.scene, cfg.assets, cfg.width, cfg.height, cfg.tile_size, planes=False)
        assert np.array_equal(out[k], single[0])


def test_device_output_tensor():
    import torch

    cfg = scenes.cube(320, 240, 40, logo_size=64)
    r = cfg.rasterizer()
    t = torch.zeros((cfg.height, cfg.width, 4), dtype=torch.uint8, device="cuda:0")
    r.rasterize(cfg.scene, t, cfg.width, cfg.height, cfg.tile_size, cfg.assets)
    host = render_gpu(r, cfg.scene, cfg.assets, cfg.width, cfg.height, cfg.tile_size, planes=False)
    assert np.array_equal(t.cpu().numpy(), host[0])


def test_errors_do_not_unwind():
    from rusterix_b200 import RxcError

    cfg = scenes.cube(64, 64, 40, logo_size=16)
    cfg.scene.d3_static[0].source(PixelSource.StaticTileIndex(7))
    cfg.scene.mark_dirty()
    with pytest.raises(RxcError) as e:
        render_gpu(cfg.rasterizer(), cfg.scene, cfg.assets, 64, 64, 40)
    assert e.value.status == -5
    cfg.scene.d3_static[0].source(PixelSource.StaticTileIndex(0))
    cfg.scene.mark_dirty()
    with pytest.raises(ValueError):  # the reference panics on a short slice (src/rasterizer.rs:572)
        cfg.rasterizer().rasterize(cfg.scene, np.zeros(10, np.uint8), 64, 64, 40, cfg.assets)
    render_gpu(cfg.rasterizer(), cfg.scene, cfg.assets, 64, 64, 40)  # the context is still usable


def test_fast_division_is_bit_exact():
    """The barycentric divisions use a residual-corrected multiply by RN(1/area) (rx_div_by) where
    the reference divides (rasterizer.rs:1768-1769); over the admitted operand ranges it must give
    the bits of div.rn.  2^32 operand pairs incl. hard mantissas and midpoint quotients."""
    from rusterix_b200 import DeviceContext

    assert DeviceContext.get(0).selftest_div(n_pairs=1 << 32) == 0
    assert DeviceContext.get(0).selftest_div(n_pairs=1 << 30, seed=12345) == 0


# ---- SURVEY 8f rows f2 (chunk path) and f3 (2D game path) ------------------------------------------
@pytest.mark.parametrize("frame", [0, 3, 5])
def test_chunk_path_opacity_surface_ids_terrain_occlusion(frame):
    """scene.chunks: opacity batches with surface ids (rasterizer.rs:1425-1690, :1041-1047, :464-495),
    terrain texels from the world position incl. their alpha test (:1178-1219, chunk.rs:135-151), sector
    occlusion (:1327-1365), chunk lights (:219-223), EntityTile / ItemTile sources that do and do not
    resolve (:1130-1177), and a 2D overlay with every primitive mode (:901-955)."""
    cfg = scenes.chunked_config(1280, 720)
    st = _run(cfg, frame)
    assert st["within1_frac"] >= 0.999


@pytest.mark.parametrize("frame", [0, 5])
def test_brush_preview_highlights_terrain_texels(frame):
    """brush_preview over Terrain batches: the sampled terrain texel goes towards white inside the brush radius, in
    d3_rasterize (rasterizer.rs:1193-1212) and in d3_rasterize_opacity (:1601-1620; one pane is given the Terrain source)."""
    from rusterix_b200.types import BrushPreview

    cfg = scenes.chunked_config(960, 540)
    cfg.brush_preview = BrushPreview((8.0, 0.0, 8.0), 3.5, 0.5)
    ch = cfg.scene.chunks[(0, 0)]
    ch.batches3d_opacity[0] = ch.batches3d_opacity[0].source(PixelSource.Terrain)
    cfg.scene.mark_dirty()
    st = _run(cfg, frame)
    assert st["within1_frac"] >= 0.999
    cfg.brush_preview = None
    plain = render_gpu(cfg.rasterizer(frame), cfg.scene, cfg.assets, cfg.width, cfg.height, cfg.tile_size)[0]
    cfg.brush_preview = BrushPreview((8.0, 0.0, 8.0), 3.5, 0.5)
    lit = render_gpu(cfg.rasterizer(frame), cfg.scene, cfg.assets, cfg.width, cfg.height, cfg.tile_size)[0]
    assert (plain != lit).any()


def test_chunk_path_linear_odd_size_and_preserve_transparency():
    cfg = scenes.chunked_config(1001, 563)
    cfg.sample_mode = SampleMode.Linear
    _run(cfg, 2, preserve_transparency=True)


def test_chunk_path_batch_api_matches_single_frames():
    cfg = scenes.chunked_config(640, 360)
    rasts = [cfg.rasterizer(i) for i in range(4)]
    import copy
    lights_before = list(cfg.scene.dynamic_lights)
    out = np.zeros((4, cfg.height, cfg.width, 4), dtype=np.uint8)
    Rasterizer.rasterize_batch(rasts, cfg.scene, out, cfg.width, cfg.height, cfg.tile_size, cfg.assets)
    for i in (0, 3):
        cfg.scene.dynamic_lights = list(cfg.scene.dynamic_lights)  # same light list as the sweep saw
        o = render_oracle(rasts[i], cfg.scene, cfg.assets, cfg.width, cfg.height, cfg.tile_size)
        compare((out[i], None, None), o, f"chunked sweep frame {i}")
    assert len(cfg.scene.dynamic_lights) == len(lights_before) + 4  # chunk lights appended once (rasterizer.rs:219-223)


@pytest.mark.parametrize("size", [(960, 640), (487, 333)])
def test_game2d_binned_records_lines_los_and_sectors(size):
    """2D-only render mode with 800+ records (sorted per-tile lists), translucent decals blended in
    submission order (rasterizer.rs:876-895), entity sprites, Bresenham lines (:1777-1821), point lights
    blocked by mapmini linedefs (mini.rs:67-95) and sector occlusion of ambient light (:806-836)."""
    cfg = scenes.game2d_config(*size)
    r = cfg.rasterizer(0)
    g = render_gpu(r, cfg.scene, cfg.assets, cfg.width, cfg.height, cfg.tile_size)
    o = render_oracle(r, cfg.scene, cfg.assets, cfg.width, cfg.height, cfg.tile_size)
    st = compare(g, o, "game2d")
    assert st["exact_frac"] >= 0.999  # the 2D path is integer blending of exact texels


def test_game2d_preserve_transparency_and_background():
    cfg = scenes.game2d_config(640, 480)
    cfg.scene.background = VGrayGradientShader()
    _run(cfg, 0, preserve_transparency=True)


# ------------------------------------------------------------------------------------------------
# batch shaders: the Rusteria VM on the device
# ------------------------------------------------------------------------------------------------
import oracle_ffi
import vm_programs
from rusterix_b200 import Assets, DeviceContext
from rusterix_b200 import vm as rvm
from rusterix_b200 import marshal as marshal_mod

# ops whose results go through libm (sin, cos, tan, atan, atan2, pow, ln, sincos): CUDA and glibc agree to a few ulp
_LIBM_PROGRAMS = {"libm", "wood", "control_flow", "glass", "scanlines"}


class _StateMode:
    def __init__(self, mode):
        self.mode = mode

    def __enter__(self):
        DeviceContext.get(0).set_vm_state_mode(self.mode)
        return DeviceContext.get(0)

    def __exit__(self, *exc):
        import os
        DeviceContext.get(0).set_vm_state_mode(int(os.environ.get("RXC_VM_STATE_MODE", "2")))


def _vm_scene():
    progs = vm_programs.all_programs()
    s = Scene()
    s.patterns, s.patterns_normal = scenes.pattern_bank(), scenes.pattern_bank()[:3]
    for p in progs.values():
        s.add_shader(p)
    a = Assets.default().textures([])
    a.palette = vm_programs.PALETTE
    return progs, s, a


@pytest.mark.parametrize("name", list(vm_programs.all_programs()))
def test_vm_device_matches_oracle_interpreter(name):
    """Every NodeOp, on 4096 random Execution states: the device runs the flat code, the oracle walks the tree."""
    progs, scene, assets = _vm_scene()
    ctx = DeviceContext.get(0)
    ctx.upload(scene, assets)
    oracle_ffi.set_programs(scene, assets)
    recs = vm_programs.records(4096, seed=11)
    k = list(progs).index(name)
    got, faults = ctx.vm_execute(k, recs)
    want, ofaults = oracle_ffi.vm_execute(k, recs)
    assert faults == 0 and ofaults == 0
    if name in _LIBM_PROGRAMS:
        np.testing.assert_allclose(got, want, rtol=2e-5, atol=2e-5)
    else:
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), np.abs(got - want).max()


def test_vm_device_limits_are_reported():
    """A program that overflows the 32-entry value stack: rxc_vm_execute counts the fault, rasterize returns an error."""
    from rusterix_b200 import RxcError
    deep = [("Push", (1.0, 1.0, 1.0))] * 40 + [("Add",)] * 39 + [("SetColor",)]
    s = Scene()
    s.add_shader(rvm.Program([deep], 0, 0, 0))
    a = Assets.default().textures([])
    ctx = DeviceContext.get(0)
    ctx.upload(s, a)
    _, faults = ctx.vm_execute(0, vm_programs.records(8))
    assert faults == 8
    cfg = scenes.cube(64, 64, 40, logo_size=16)
    cfg.scene.add_shader(rvm.Program([deep], 0, 0, 0))
    cfg.scene.d3_static[0].shader(0)
    with pytest.raises(RxcError) as e:
        render_gpu(cfg.rasterizer(), cfg.scene, cfg.assets, 64, 64, 40)
    assert e.value.status == -3
    cfg2 = scenes.cube(64, 64, 40, logo_size=16)
    render_gpu(cfg2.rasterizer(), cfg2.scene, cfg2.assets, 64, 64, 40)  # the context is still usable


@pytest.mark.parametrize("frame", [0, 5, 11])
def test_shaded_scene_parity(frame):
    """Programs on opaque 3D batches (incl. one that cuts holes through `opacity`), on a chunk's batches (its own
    program list, a baked shader texture), on an opacity-pass pane and on 2D batches."""
    cfg = scenes.shaded_config(640, 480, 40)
    st = _run(cfg, frame=frame)
    assert st["within1_frac"] > 0.999


def test_shaded_scene_linear_odd_size():
    cfg = scenes.shaded_config(501, 333, 64)
    cfg.sample_mode = SampleMode.Linear
    _run(cfg, frame=3)


def test_shaded_scene_is_tile_size_invariant_on_the_device():
    """Per-fragment Execution state: unlike the reference, the frame cannot depend on tile_size."""
    cfg = scenes.shaded_config(320, 240, 40)
    a = render_gpu(cfg.rasterizer(2), cfg.scene, cfg.assets, 320, 240, 16)
    b = render_gpu(cfg.rasterizer(2), cfg.scene, cfg.assets, 320, 240, 240)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_shaded_scene_with_emissive_against_per_fragment_oracle():
    """A program that writes `emissive` on one path only.  The reference leaks the value into later fragments of the
    tile; the device does not, so this frame is checked against the oracle in per-fragment-state mode."""
    cfg = scenes.shaded_config(480, 360, 40, emissive=True)
    oracle_ffi.set_vm_state_mode(True)
    try:
        with _StateMode(0):      # the fast kernel (by default such a scene is rendered in the reference's order, see below)
            _run(cfg, frame=1)
    finally:
        oracle_ffi.set_vm_state_mode(False)


# ------------------------------------------------------------------------------------------------
# batch shaders compiled at run time (rx_jit.cu): the same programs as straight-line code, NVRTC-compiled
# ------------------------------------------------------------------------------------------------
class _JitMode:
    def __init__(self, mode):
        self.mode = mode

    def __enter__(self):
        DeviceContext.get(0).set_vm_jit(self.mode)
        return DeviceContext.get(0)

    def __exit__(self, *exc):
        import os
        DeviceContext.get(0).set_vm_jit(int(os.environ.get("RXC_VM_JIT", "1")))


def test_vm_jit_kernel_equals_the_interpreter_on_every_op():
    """Every NodeOp program, 4096 random Execution states each: the generated code and the interpreter agree bit for bit
    (they call the same vm_un / vm_bin / vm_tern functions), and the generated code is what ran."""
    progs, scene, assets = _vm_scene()
    recs = vm_programs.records(4096, seed=23)
    with _JitMode(0) as ctx:
        launches = ctx.vm_jit_info()["launches"]
        ctx.upload(scene, assets)
        want = [ctx.vm_execute(k, recs) for k in range(len(progs))]
        assert ctx.vm_jit_info()["launches"] == launches
    with _JitMode(2) as ctx:
        before = ctx.vm_jit_info()
        ctx.upload(scene, assets)
        got = [ctx.vm_execute(k, recs) for k in range(len(progs))]
        info = ctx.vm_jit_info()
    assert info["translated"] == len(progs) and info["launches"] - before["launches"] == len(progs), (before, info)
    for name, (g, gf), (w, wf) in zip(progs, got, want):
        assert gf == wf == 0, name
        assert np.array_equal(g.view(np.uint32), w.view(np.uint32)), (name, np.abs(g - w).max())


@pytest.mark.parametrize("sample_mode", [SampleMode.Nearest, SampleMode.Linear])
def test_shaded_scene_jit_equals_interpreter(sample_mode):
    """The whole batch-shader scene (3D, chunk, opacity pane, 2D programs) through the recompiled raster kernel:
    pixels, owner and depth identical to the interpreter's frame; and parity with the oracle as before."""
    cfg = scenes.shaded_config(640, 480, 40)
    cfg.sample_mode = sample_mode
    with _JitMode(0):
        a = render_gpu(cfg.rasterizer(4), cfg.scene, cfg.assets, 640, 480, 40)
    with _JitMode(2) as ctx:
        before = ctx.vm_jit_info()
        b = render_gpu(cfg.rasterizer(4), cfg.scene, cfg.assets, 640, 480, 40)
        info = ctx.vm_jit_info()
        assert info["launches"] > before["launches"] and info["kernels"] >= 1 and info["translated"] >= 4, (before, info)
        st = _run(cfg, frame=4)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2].view(np.uint32), b[2].view(np.uint32))
    assert st["within1_frac"] > 0.999


def test_vm_jit_background_compile_switches_kernels_between_frames():
    """Default mode: the first frames run the interpreter while the worker thread compiles; once the kernel is there the
    next frame uses it -- and is the same frame."""
    import time
    cfg = scenes.shaded_config(320, 240, 40, emissive=True)    # a program set no other test compiles
    with _StateMode(0), _JitMode(1) as ctx:                     # (the fast kernel: in auto mode this scene goes to the reference-order kernel)
        first = render_gpu(cfg.rasterizer(1), cfg.scene, cfg.assets, 320, 240, 40, planes=False)
        t0 = time.time()
        while ctx.vm_jit_info()["pending"] and time.time() - t0 < 120:
            time.sleep(0.2)
        before = ctx.vm_jit_info()
        second = render_gpu(cfg.rasterizer(1), cfg.scene, cfg.assets, 320, 240, 40, planes=False)
        after = ctx.vm_jit_info()
    assert not before["pending"], before
    assert after["launches"] > before["launches"], (before, after)
    assert np.array_equal(first[0], second[0])


def test_vm_jit_declined_program_still_faults_through_the_interpreter():
    from rusterix_b200 import RxcError
    deep = [("Push", (1.0, 1.0, 1.0))] * 40 + [("Add",)] * 39 + [("SetColor",)]
    with _JitMode(2) as ctx:
        s = Scene()
        s.add_shader(rvm.Program([deep], 0, 0, 0))
        s.add_shader(rvm.Program([[("UV",), ("SetColor",)]], 0, 0, 0))
        a = Assets.default().textures([])
        ctx.upload(s, a)
        _, faults = ctx.vm_execute(0, vm_programs.records(8))
        assert faults == 8
        out, faults = ctx.vm_execute(1, vm_programs.records(8))
        assert faults == 0 and ctx.vm_jit_info()["translated"] == 1


def test_vm_state_report_names_the_scene_whose_frames_depend_on_the_tile_execution():
    """rxc_vm_scene_state_report: the batch-shader scene's programs cannot observe the reference's never-reset per-tile
    Execution (so the frame is compared with the FAITHFUL oracle, test_shaded_scene_parity); its `emissive` variant can --
    that is the one scene checked against the oracle in per-fragment-state mode."""
    ctx = DeviceContext.get(0)
    cfg = scenes.shaded_config(160, 120, 40)
    render_gpu(cfg.rasterizer(0), cfg.scene, cfg.assets, 160, 120, 40)
    plain = ctx.vm_state_report()
    cfg2 = scenes.shaded_config(160, 120, 40, emissive=True)
    render_gpu(cfg2.rasterizer(0), cfg2.scene, cfg2.assets, 160, 120, 40)
    leaky = ctx.vm_state_report()
    assert len(plain) == len(leaky) >= 4
    assert plain == [0] * len(plain), plain
    assert 1 in leaky and 2 not in leaky, leaky


# ------------------------------------------------------------------------------------------------
# rxc_set_vm_state_mode: the reference's order and per-tile Execution on the device (k_raster_ordered)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("tile_size", [40, 16, 64])
def test_reference_order_mode_reproduces_the_state_leak(tile_size):
    """The `emissive` variant of the batch-shader scene: one branch of a program writes `emissive`, the reference never resets its
    per-tile Execution, so every fragment shaded later in that tile glows -- a frame that depends on tile_size and on the order of
    the triangles.  In reference-order mode the device renders THAT frame: compared with the FAITHFUL oracle (no per-fragment
    switch), owner and depth bit for bit, colours within the bar.  The fast mode renders something else."""
    cfg = scenes.shaded_config(480, 360, tile_size, emissive=True)
    with _StateMode(1) as ctx:
        n0 = ctx.ordered_frames()
        st = _run(cfg, frame=1)
        assert ctx.ordered_frames() == n0 + 1
        ordered = render_gpu(cfg.rasterizer(1), cfg.scene, cfg.assets, 480, 360, tile_size)
    assert st["within1_frac"] > 0.999
    with _StateMode(0):
        fast = render_gpu(cfg.rasterizer(1), cfg.scene, cfg.assets, 480, 360, tile_size)
    assert np.array_equal(fast[1], ordered[1]) and np.array_equal(fast[2].view(np.uint32), ordered[2].view(np.uint32))   # same owners, same depth
    leak = np.abs(fast[0].astype(np.int16) - ordered[0].astype(np.int16)).max(axis=-1) > 1
    assert leak.mean() > 0.01, leak.mean()                                                                                 # ... other colours


def test_reference_order_mode_with_compiled_programs_equals_the_interpreted_one():
    """k_raster_ordered recompiled with the scene's programs as straight-line code (carried globals / locals included): the same
    frame as with the interpreter, bit for bit, and it is the kernel that ran."""
    cfg = scenes.shaded_config(320, 240, 40, emissive=True)
    with _StateMode(1), _JitMode(0):
        a = render_gpu(cfg.rasterizer(3), cfg.scene, cfg.assets, 320, 240, 40)
    with _StateMode(1), _JitMode(2) as ctx:
        before = ctx.vm_jit_info()["launches"]
        b = render_gpu(cfg.rasterizer(3), cfg.scene, cfg.assets, 320, 240, 40)
        assert ctx.vm_jit_info()["launches"] == before + 1, ctx.vm_jit_info()
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2].view(np.uint32), b[2].view(np.uint32))


def test_reference_order_mode_equals_the_fast_mode_where_nothing_leaks():
    """The plain batch-shader scene (3D, chunk, opacity pane and 2D programs that assign what they read): both kernels render the
    reference's frame -- same owners and depth, colours within 1 LSB of each other and of the oracle."""
    cfg = scenes.shaded_config(480, 360, 40)
    with _StateMode(0):
        fast = render_gpu(cfg.rasterizer(2), cfg.scene, cfg.assets, 480, 360, 40)
    with _StateMode(1):
        ordered = render_gpu(cfg.rasterizer(2), cfg.scene, cfg.assets, 480, 360, 40)
        st = _run(cfg, frame=2)
    assert st["within1_frac"] > 0.999
    assert np.array_equal(fast[1], ordered[1]) and np.array_equal(fast[2].view(np.uint32), ordered[2].view(np.uint32))
    d = np.abs(fast[0].astype(np.int16) - ordered[0].astype(np.int16)).max(axis=-1)
    assert (d <= 1).mean() > 0.999, (d <= 1).mean()


def test_state_mode_auto_takes_the_reference_order_only_where_the_report_asks_for_it():
    with _StateMode(2) as ctx:
        n0 = ctx.ordered_frames()
        plain = scenes.shaded_config(160, 120, 40)
        _run(plain, frame=0)
        assert ctx.ordered_frames() == n0                       # nothing to observe: the fast kernel
        leaky = scenes.shaded_config(160, 120, 40, emissive=True)
        _run(leaky, frame=0)                                    # ... compared with the faithful oracle
        assert ctx.ordered_frames() == n0 + 1
        cube = scenes.cube(160, 120, 40, logo_size=16)
        _run(cube)
        assert ctx.ordered_frames() == n0 + 1                   # no programs at all


def test_reference_order_mode_refuses_bands():
    from rusterix_b200 import RxcError
    cfg = scenes.shaded_config(160, 128, 40, emissive=True)
    with _StateMode(1):
        with pytest.raises(RxcError) as e:
            render_gpu(cfg.rasterizer(0), cfg.scene, cfg.assets, 160, 128, 40, band=(32, 96))
        assert e.value.status == -3


# ------------------------------------------------------------------------------------------------
# rxc_update_scene: the frame loop of an engine (the world stays, the dynamic batches change)
# ------------------------------------------------------------------------------------------------
def _entity_box(x, tile=0, y=0.2, z=6.0, size=0.8):
    return (Batch3D.from_box(x, y, z, size, 1.5 * size, size).source(PixelSource.StaticTileIndex(tile)).cull_mode(CullMode.Off)
            .with_computed_normals())


@pytest.mark.parametrize("name", ["map", "chunked", "dense"])
def test_update_scene_reuses_the_resident_world(name):
    """Replacing scene.d3_dynamic re-uploads only what follows the unchanged batches (chunks' and static batches' geometry stays
    on the device): the frame equals the one after a full rxc_set_scene, bit for bit, and far fewer bytes cross PCIe."""
    cfg = {"map": lambda: scenes.map_config(640, 360, 40, logo_size=64), "chunked": lambda: scenes.chunked_config(480, 270, 40),
           "dense": lambda: scenes.dense(640, 360, 40, patches=8, patch_verts=23)}[name]()
    ctx = DeviceContext.get(0)
    base_lights = list(cfg.scene.dynamic_lights)

    def render():
        cfg.scene.dynamic_lights = list(base_lights)
        return render_gpu(cfg.rasterizer(0), cfg.scene, cfg.assets, cfg.width, cfg.height, cfg.tile_size)

    kw = dict(y=2.0, z=10.0, size=3.0) if name == "dense" else {}
    off = 26.0 if name == "dense" else 0.0
    cfg.scene.d3_dynamic = [_entity_box(off + 5.0, **kw)]
    a = render()
    n_before = len(marshal_mod.submission_order(cfg.scene)[0]) - 1 - len(cfg.scene.d3_overlay)
    h0 = ctx.stats().h2d_bytes
    cfg.scene.d3_dynamic = [_entity_box(off + 6.5, **kw), _entity_box(off + 3.0, tile=1, **kw)]     # the entities moved, one more appeared
    b = render()
    updated_bytes = ctx.stats().h2d_bytes - h0
    assert ctx.last_upload_kept == n_before, (ctx.last_upload_kept, n_before)
    ctx._scene_key = None                                                    # the same scene through a full rxc_set_scene
    h1 = ctx.stats().h2d_bytes
    c = render()
    full_bytes = ctx.stats().h2d_bytes - h1
    assert ctx.last_upload_kept == 0
    assert np.array_equal(b[0], c[0]) and np.array_equal(b[1], c[1]) and np.array_equal(b[2].view(np.uint32), c[2].view(np.uint32))
    assert not np.array_equal(a[0], b[0])
    if name == "dense":
        assert updated_bytes * 4 < full_bytes, (updated_bytes, full_bytes)
    compare(b, render_oracle(cfg.rasterizer(0), cfg.scene, cfg.assets, cfg.width, cfg.height, cfg.tile_size), name + " after rxc_update_scene")


def test_update_scene_rejects_a_prefix_that_is_not_resident():
    from rusterix_b200 import RxcError, marshal as m
    cfg = scenes.cube(64, 64, 40, logo_size=16)
    ctx = DeviceContext.get(0)
    render_gpu(cfg.rasterizer(), cfg.scene, cfg.assets, 64, 64, 40)
    ms = m.marshal_scene(cfg.scene, 4, cfg.assets)
    import ctypes as C
    assert ctx.lib.rxc_update_scene(ctx.handle, C.byref(ms.struct), 5) == -1          # more batches than the scene has
    ctx._scene_key = None
    other = scenes.teapot(64, 64, 40, logo_size=16, n_frames=2)
    ms2 = m.marshal_scene(other.scene, 4, other.assets)
    assert ctx.lib.rxc_update_scene(ctx.handle, C.byref(ms2.struct), 1) == -1         # another batch (vertex count differs)
    render_gpu(cfg.rasterizer(), cfg.scene, cfg.assets, 64, 64, 40)                   # the context is still usable


# ------------------------------------------------------------------------------------------------
# raster kernels recompiled for the scene (rx_jit.cu): the scene's and the frame's constants folded in
# ------------------------------------------------------------------------------------------------
def _spec_cases():
    yield "cube", lambda: scenes.cube(800, 600, 200, logo_size=256), 0
    yield "teapot-linear", lambda: scenes.teapot(960, 540, 60, logo_size=256, n_frames=8), 3
    yield "map", lambda: scenes.map_config(1280, 720, 40, logo_size=256), 0
    yield "sweep", lambda: scenes.sweep(960, 540, 40, n_frames=64, logo_size=256), 17
    yield "dense", lambda: scenes.dense(1280, 720, 40, patches=8, patch_verts=23), 0
    yield "dense-long-lists", lambda: scenes.dense(640, 360, 40, patches=12, patch_verts=30), 0
    yield "chunked", lambda: scenes.chunked_config(640, 360, 40), 2
    yield "game2d", lambda: scenes.game2d_config(480, 320), 0
    yield "sky", lambda: scenes.sky_config(480, 270, 40, hour=16.5), 1


@pytest.mark.parametrize("name,make,frame", list(_spec_cases()), ids=[c[0] for c in _spec_cases()])
def test_scene_specialised_kernel_equals_the_generic_kernel(name, make, frame):
    """k_raster recompiled with the scene's constants (shade-descriptor bits shared by all batches, light count and type,
    ambient / sun / sky / 2D presence, no alpha test anywhere) renders the frame of the library's generic kernel bit for
    bit -- pixels, owner ids and depth -- with and without the owner / depth planes, and it is the kernel that ran."""
    cfg = make()
    base_lights = list(cfg.scene.dynamic_lights)

    def render(planes):
        cfg.scene.dynamic_lights = list(base_lights)   # rasterize() appends the chunks' lights on every call (src/rasterizer.rs:219-223)
        return render_gpu(cfg.rasterizer(frame), cfg.scene, cfg.assets, cfg.width, cfg.height, cfg.tile_size, planes=planes)

    with _JitMode(0):
        a, a2 = render(True), render(False)
    with _JitMode(2) as ctx:
        before = ctx.vm_jit_info()
        b, b2 = render(True), render(False)
        after = ctx.vm_jit_info()
    assert after["launches"] - before["launches"] >= 2, (before, after)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2].view(np.uint32), b[2].view(np.uint32))
    assert np.array_equal(a2[0], b2[0]) and np.array_equal(a[0], a2[0])


def test_scene_specialised_kernel_follows_light_and_frame_changes():
    """The signature is per launch: another light count, another light type, ambient switched off -- each picks (compiles)
    its own kernel; a kernel is never run on a frame it was not compiled for (the kernel checks and would report it)."""
    from rusterix_b200 import Light, LightType
    cfg = scenes.map_config(640, 360, 40, logo_size=64)
    extra = (Light.new(LightType.Spot).with_position([7.0, 1.8, 7.0]).with_color([0.6, 0.8, 1.0]).with_intensity(1.5)
             .with_start_distance(1.0).with_end_distance(9.0).with_direction([0.0, -1.0, 0.2]).with_cone_angle(0.9).compile())
    frames = {}
    for mode in (0, 2):
        with _JitMode(mode):
            out = []
            cfg.scene.dynamic_lights = []
            out.append(render_gpu(cfg.rasterizer(0), cfg.scene, cfg.assets, 640, 360, 40))
            cfg.scene.dynamic_lights = [extra]                       # two lights of two types
            out.append(render_gpu(cfg.rasterizer(0), cfg.scene, cfg.assets, 640, 360, 40))
            r = cfg.rasterizer(0)
            r.ambient_color = None                                    # frame signature: no ambient term
            out.append(render_gpu(r, cfg.scene, cfg.assets, 640, 360, 40))
            cfg.scene.dynamic_lights = []
            frames[mode] = out
    for a, b in zip(frames[0], frames[2]):
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2].view(np.uint32), b[2].view(np.uint32))
    assert not np.array_equal(frames[0][0][0], frames[0][1][0]) and not np.array_equal(frames[0][1][0], frames[0][2][0])


# ------------------------------------------------------------------------------------------------
# render graph: Sky node on uncovered pixels, directional sun, brush preview
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("frame,hour", [(0, 16.5), (3, 7.0), (5, 12.0), (6, 22.0)])
def test_sky_sun_and_brush_preview(frame, hour):
    cfg = scenes.sky_config(640, 360, 40, hour=hour)
    st = _run(cfg, frame=frame)
    assert st["within1_frac"] > 0.999


def test_sky_only_every_pixel_is_a_miss():
    cfg = scenes.sky_config(333, 211, 64, hour=18.5)
    cfg.scene.d3_static.clear()
    st = _run(cfg, frame=2)
    assert st["exact_frac"] > 0.999     # same IEEE arithmetic on both sides, no libm


def test_sky_cloud_layer_is_reported_unsupported():
    from rusterix_b200 import RxcError
    from rusterix_b200.types import RenderGraph, SkyNode
    cfg = scenes.sky_config(64, 64, 32)
    cfg.render_graph = RenderGraph([SkyNode(clouds=True)])
    with pytest.raises(RxcError) as e:
        render_gpu(cfg.rasterizer(), cfg.scene, cfg.assets, 64, 64, 32)
    assert e.value.status == -3


# ------------------------------------------------------------------------------------------------
# edge cases and full-size configurations
# ------------------------------------------------------------------------------------------------
def _tri_batch(verts, tris, uvs=None, tile=0):
    v = np.asarray(verts, dtype=np.float32)
    uvs = uvs if uvs is not None else [(0.1 * i, 0.07 * i) for i in range(len(v))]
    return Batch3D(v, tris, uvs).source(PixelSource.StaticTileIndex(tile)).cull_mode(CullMode.Off).with_computed_normals()


def test_empty_scene_and_empty_batches():
    """No batches at all, and batches without triangles or without vertices next to a drawn one."""
    cfg = scenes.cube(200, 120, 40, logo_size=16)
    cfg.scene.d3_static.clear()
    cfg.scene.d2_static.clear()
    st = _run(cfg, what="empty scene")
    assert st["exact_frac"] == 1.0
    cfg = scenes.cube(200, 120, 40, logo_size=16)
    cfg.scene.d3_static.insert(0, _tri_batch([(0, 0, 0, 1), (1, 0, 0, 1), (0, 1, 0, 1)], []))     # vertices, no triangles
    cfg.scene.d3_static.append(Batch3D(np.zeros((0, 4), np.float32), [], []).source(PixelSource.StaticTileIndex(0)))
    cfg.scene.d2_static.append(Batch2D.new([], [], []))
    _run(cfg, what="empty batches")


def test_degenerate_and_non_finite_geometry():
    """Zero-area and repeated-vertex triangles pass coverage but never the depth test (NaN z); NaN / Inf vertices and
    a vertex on the camera plane (w = 0) follow the formulas instead of being culled (SURVEY T-edge, T-cast)."""
    cfg = scenes.cube(320, 200, 40, logo_size=16)
    nan, inf = float("nan"), float("inf")
    verts = [(-0.9, -0.2, 0.3, 1), (0.9, -0.2, 0.3, 1), (0.0, 0.7, 0.3, 1),      # a plain triangle
             (-0.5, 0.0, 0.6, 1), (0.5, 0.0, 0.6, 1), (0.0, 0.0, 0.6, 1),        # collinear (zero area)
             (0.2, 0.2, 0.8, 1), (0.2, 0.2, 0.8, 1), (0.4, 0.5, 0.8, 1),         # repeated vertex
             (nan, 0.1, 0.2, 1), (0.3, nan, 0.2, 1), (0.1, 0.3, inf, 1),         # non-finite
             (0.0, 0.5, 0.0, 1), (-inf, 0.0, 0.5, 1)]
    tris = [(0, 1, 2), (3, 4, 5), (6, 7, 8), (9, 0, 1), (10, 1, 2), (11, 2, 0), (0, 1, 13), (12, 0, 2), (0, 0, 0)]
    cfg.scene.d3_static.append(_tri_batch(verts, tris))
    _run(cfg, what="degenerate / non-finite")
    # the camera sits exactly in the plane of a vertex: w = 0 after projection for the unclipped path
    cfg.camera = scenes._firstp([0.0, 0.5, 0.0], [0.0, 0.5, 1.0])
    _run(cfg, what="vertex at the eye")


def test_huge_offscreen_and_subpixel_triangles():
    cfg = scenes.cube(400, 300, 50, logo_size=16)
    big = 1.0e6
    verts = [(-big, -big, -3.0, 1), (big, -big, -3.0, 1), (0.0, big, -3.0, 1),           # covers the screen, far away
             (50.0, 50.0, 0.0, 1), (51.0, 50.0, 0.0, 1), (50.0, 51.0, 0.0, 1),           # completely off screen
             (0.1, 0.1, 0.55, 1), (0.1004, 0.1, 0.55, 1), (0.1, 0.1004, 0.55, 1),        # smaller than a pixel
             (0.3, 0.1, 0.55, 1), (0.3, 0.1 + 1e-7, 0.55, 1), (0.9, 0.1, 0.55, 1)]       # a sliver
    cfg.scene.d3_static.append(_tri_batch(verts, [(0, 1, 2), (3, 4, 5), (6, 7, 8), (9, 10, 11)]))
    _run(cfg, what="huge / offscreen / subpixel")


def test_frustum_rejected_batch_and_all_batches_behind_camera():
    cfg = scenes.cube(256, 160, 40, logo_size=16)
    behind = Batch3D.from_box(-0.5, -0.5, 5.0, 1.0, 1.0, 1.0).source(PixelSource.StaticTileIndex(0)).cull_mode(CullMode.Off).with_computed_normals()
    cfg.scene.d3_static.append(behind)          # on the far side of the orbit camera's back: AABB reject (batch3d.rs:492-552)
    _run(cfg, what="one rejected batch")
    cfg.scene.d3_static = [behind]
    cfg.scene.mark_dirty()
    cfg.camera = scenes._firstp([0.0, 0.0, 0.0], [0.0, 0.0, -1.0])
    _run(cfg, what="everything behind the camera")


def test_dense_full_8k_owner_and_depth():
    """BASELINE.json config D at its full size: 991,232 triangles, 7 lights, 7680x4320, Linear -- every pixel's
    owner and depth bit for bit, colours within 1 LSB."""
    st = _run(scenes.dense(7680, 4320, 40))
    assert st["within1_frac"] > 0.9999


def test_8k_bands_concatenate_to_the_full_frame():
    """Config D's sharding: 8 row bands rendered separately are the full frame (pixels and owners)."""
    from rusterix_b200 import mgpu
    cfg = scenes.dense(7680, 4320, 40, patches=16)
    r = cfg.rasterizer()
    full = render_gpu(r, cfg.scene, cfg.assets, cfg.width, cfg.height, cfg.tile_size)
    for (y0, y1) in mgpu.all_bands(cfg.height, 8):
        band = render_gpu(r, cfg.scene, cfg.assets, cfg.width, cfg.height, cfg.tile_size, band=(y0, y1))
        assert np.array_equal(band[0], full[0][y0:y1]) and np.array_equal(band[1], full[1][y0:y1])


def test_sweep_of_256_frames_in_one_call_is_deterministic_and_matches_single_frames():
    """Config E's unit of work: many cameras, one launch sequence (groups of frames share a workspace)."""
    cfg = scenes.sweep(480, 270, 40, n_frames=4096, logo_size=64)
    ids = list(range(0, 4096, 16))
    rs = [cfg.rasterizer(i) for i in ids]
    a = np.zeros((len(ids), cfg.height, cfg.width, 4), dtype=np.uint8)
    b = np.zeros_like(a)
    Rasterizer.rasterize_batch(rs, cfg.scene, a, cfg.width, cfg.height, cfg.tile_size, cfg.assets)
    Rasterizer.rasterize_batch(rs, cfg.scene, b, cfg.width, cfg.height, cfg.tile_size, cfg.assets)
    assert np.array_equal(a, b)
    for k in (0, 97, 255):
        single = render_gpu(rs[k], cfg.scene, cfg.assets, cfg.width, cfg.height, cfg.tile_size, planes=False)
        assert np.array_equal(a[k], single[0])


def test_config_e_sweep_of_1024_frames_in_contiguous_blocks_through_the_delivery_buffer():
    """Config E the way bench.py's sweep4096 block runs it: contiguous blocks of frames, 32 cameras per launch sequence,
    rendered straight into the rxc_mgpu delivery buffer (two-slot ring, release hand-shake; world 1 here, the multi-rank
    legs are in test_mgpu_gpu.py).  1024 frames of the 4096-frame path; every 37th is compared byte for byte with a
    single-frame render, and with the oracle for a few."""
    import torch
    from rusterix_b200 import DeviceContext, mgpu

    W, H, per_launch, n = 320, 180, 32, 1024
    cfg = scenes.sweep(W, H, 40, n_frames=4096, logo_size=64)
    fb = W * H * 4
    ctx = DeviceContext.get(0)
    dl = mgpu.Delivery(ctx, 0, 1)
    buf = dl.target(2 * per_launch * fb)
    first = 1536                                   # the block of "rank 3 of 8": frames 1536 .. 2559
    got = torch.empty((n, H, W, 4), dtype=torch.uint8, device="cuda:0")
    for k in range(n // per_launch):
        ids = range(first + k * per_launch, first + (k + 1) * per_launch)
        batch = Rasterizer.prepare_batch([cfg.rasterizer(i) for i in ids], cfg.scene, W, H, cfg.tile_size, cfg.assets)
        slot = k % 2
        if k >= 2:
            dl.release()
        dl.render(batch, slot * per_launch * fb)
        dl.deliver(mgpu.frame_regions(1, per_launch, fb, slot * per_launch * fb))
        ctx.synchronize()
        got[k * per_launch:(k + 1) * per_launch] = buf[slot * per_launch * fb:(slot + 1) * per_launch * fb].reshape(per_launch, H, W, 4)
    dl.close()
    got = got.cpu().numpy()
    assert len({got[i].tobytes() for i in range(0, n, 8)}) == n // 8          # the camera really moves
    for i in range(0, n, 37):
        single = render_gpu(cfg.rasterizer(first + i), cfg.scene, cfg.assets, W, H, cfg.tile_size, planes=False)
        assert np.array_equal(got[i], single[0]), i
    for i in (0, 511, 1023):
        o = render_oracle(cfg.rasterizer(first + i), cfg.scene, cfg.assets, W, H, cfg.tile_size, planes=False)
        diff = np.abs(got[i].astype(np.int16) - o[0].astype(np.int16)).max(axis=-1)
        assert (diff <= 1).mean() >= 0.999


def test_one_context_many_scenes_in_any_order():
    """The context keeps scene, textures, programs and a workspace resident: switching between scenes of different
    size, kernel mode (fast / general / general + VM), frame size and light count must never leak state."""
    order = [
        lambda: scenes.shaded_config(320, 240, 40),
        lambda: scenes.cube(200, 150, 40, logo_size=32),
        lambda: scenes.dense(640, 360, 40, patches=6),
        lambda: scenes.chunked_config(480, 270, 40),
        lambda: scenes.sky_config(320, 180, 40),
        lambda: scenes.game2d_config(240, 160),
        lambda: scenes.map_config(1280, 720, 40, logo_size=64),
        lambda: scenes.cube(64, 64, 200, logo_size=8),
        lambda: scenes.shaded_config(160, 120, 16, emissive=False),
        lambda: scenes.teapot(480, 270, 60, logo_size=64),
    ]
    for k, make in enumerate(order + order[::-1]):
        cfg = make()
        _run(cfg, frame=k % max(1, cfg.n_frames), what=f"#{k} {cfg.name}")


def test_lights_can_change_without_a_scene_upload():
    """examples/cube.rs:72-73 moves the light every frame: rxc_set_lights path, growing and shrinking lists."""
    cfg = scenes.cube(240, 180, 60, logo_size=32)
    base = cfg.scene.lights[0]
    for n in (1, 3, 0, 2):
        cfg.scene.lights = []
        for i in range(n):
            l = (Light.new(LightType.Point).with_intensity(1.0 + 0.3 * i).with_color([1.0, 0.9 - 0.2 * i, 0.6 + 0.1 * i])
                 .with_position([2.0 * math.cos(0.9 * i + n), 0.8, 2.0 * math.sin(0.9 * i + n)]).with_start_distance(1.0).with_end_distance(4.0))
            cfg.scene.lights.append(l.compile())
        _run(cfg, what=f"{n} lights")
    assert base is not None


def test_more_lights_than_the_shared_memory_table_and_more_large_triangles_than_the_cache():
    """40 lights (the raster kernel stages 16 per frame in shared memory, the rest of such a list is read from global
    memory) and 220 screen-filling triangles (160 large-triangle records are cached per frame, the two-level tile
    selection runs above 32, the remainder is walked from the large list)."""
    cfg = scenes.cube(384, 256, 64, logo_size=32)
    rng = np.random.default_rng(5)
    cfg.scene.lights = [Light.new(LightType(int(rng.choice([0, 3, 4])))).with_intensity(0.15).with_color(list(rng.random(3)))
                        .with_position(list(rng.normal(0, 2.0, 3))).with_start_distance(0.5).with_end_distance(6.0).compile() for _ in range(40)]
    verts, tris, uvs = [], [], []
    for i in range(220):
        z = -1.5 - 0.01 * i
        a = 0.1 * i
        base = len(verts)
        for (x, y) in ((-6, -5), (6, -5), (0, 7)):
            verts.append((x * math.cos(a) - y * math.sin(a), x * math.sin(a) + y * math.cos(a), z, 1.0))
            uvs.append((x * 0.3, y * 0.3))
        tris.append((base, base + 1, base + 2))
    big = Batch3D(np.asarray(verts, dtype=np.float32), tris, uvs).source(PixelSource.StaticTileIndex(0)).cull_mode(CullMode.Off) \
        .repeat_mode(RepeatMode.RepeatXY).with_computed_normals()
    cfg.scene.d3_static.append(big)
    st = _run(cfg, what="40 lights, 220 large triangles")
    from rusterix_b200 import DeviceContext
    assert DeviceContext.get(0).stats().last_large_tris > 160


@pytest.mark.parametrize("size", [(2600, 1700), (4099, 1031), (2048, 2176)])
def test_sliced_host_output_matches_device_output(size):
    """Host pixel buffers of large frames are rendered and drained as horizontal slices (rx_api.cu, pipelined host
    path): same bytes as the device-resident render, for heights that are not a multiple of the GPU tile, for a
    row band, and for a batch of frames."""
    import torch

    w, h = size
    cfg = scenes.map_config(w, h, 40, logo_size=64)
    r = cfg.rasterizer()
    dev = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda:0")
    r.rasterize(cfg.scene, dev, w, h, cfg.tile_size, cfg.assets)
    ref = dev.cpu().numpy()
    host = np.zeros((h, w, 4), dtype=np.uint8)
    r.rasterize(cfg.scene, host, w, h, cfg.tile_size, cfg.assets)
    assert np.array_equal(host, ref)
    y0, y1 = 64, h - 37
    band = np.zeros((y1 - y0, w, 4), dtype=np.uint8)
    r.rasterize(cfg.scene, band, w, h, cfg.tile_size, cfg.assets, band=(y0, y1))
    assert np.array_equal(band, ref[y0:y1])
    out = np.zeros((3, h, w, 4), dtype=np.uint8)
    Rasterizer.rasterize_batch([r, r, r], cfg.scene, out, w, h, cfg.tile_size, cfg.assets)
    for k in range(3):
        assert np.array_equal(out[k], ref)


def test_pin_host_buffer_same_frame_and_reusable():
    """rxc_pin_host / rxc_unpin_host: a caller-owned host buffer page-locked through the ABI receives the same bytes as a
    pageable one; pinning twice is harmless, unpinning makes the buffer pageable again."""
    from rusterix_b200 import DeviceContext

    cfg = scenes.map_config(1280, 720, 40, logo_size=64)
    r = cfg.rasterizer()
    pageable = np.zeros((720, 1280, 4), dtype=np.uint8)
    r.rasterize(cfg.scene, pageable, 1280, 720, cfg.tile_size, cfg.assets)
    ctx = DeviceContext.get(0)
    pinned = np.zeros((720, 1280, 4), dtype=np.uint8)
    ctx.pin_host(pinned)
    ctx.pin_host(pinned)
    try:
        for _ in range(2):
            pinned[:] = 0
            r.rasterize(cfg.scene, pinned, 1280, 720, cfg.tile_size, cfg.assets)
            assert np.array_equal(pinned, pageable)
    finally:
        ctx.unpin_host(pinned)
    r.rasterize(cfg.scene, pinned, 1280, 720, cfg.tile_size, cfg.assets)
    assert np.array_equal(pinned, pageable)


@pytest.mark.parametrize("make", [
    lambda: scenes.cube(800, 600, 200, logo_size=64), lambda: scenes.cube(333, 217, 40, logo_size=32),
    lambda: scenes.teapot(960, 540, 60, logo_size=64), lambda: scenes.dense(1280, 720, 40, patches=8),
    lambda: scenes.chunked_config(640, 360, 40), lambda: scenes.game2d_config(480, 320),
    lambda: scenes.shaded_config(320, 240, 40), lambda: scenes.sky_config(320, 180, 40), lambda: scenes.map_config(640, 360, 40, logo_size=64),
], ids=["cube", "cube_odd", "teapot", "dense", "chunked", "game2d", "shaded", "sky", "map"])
def test_empty_tile_fast_path_matches_the_full_path(make):
    """Tiles nothing can touch are filled with the miss colour without running the tile (k_raster, empty-tile path).  The
    kernel variant that also writes the owner / depth planes never takes that path: both must give the same pixels, for
    the whole frame, for a row band and for odd sizes with partial tiles."""
    def render(frame, **kw):
        cfg = make()   # a fresh scene per call: rasterize() appends the chunk lights to scene.dynamic_lights every time
        return render_gpu(cfg.rasterizer(frame), cfg.scene, cfg.assets, cfg.width, cfg.height, cfg.tile_size, **kw)[0]

    height = make().height
    for frame in (0, 3):
        full = render(frame, planes=True)
        assert np.array_equal(render(frame, planes=False), full)
        y0, y1 = 32, height - 45
        assert np.array_equal(render(frame, planes=False, band=(y0, y1)), full[y0:y1])


@pytest.mark.parametrize("make", [lambda: scenes.map_config(640, 360, 40, logo_size=64), lambda: scenes.dense(1280, 720, 40, patches=8),
                                  lambda: scenes.chunked_config(640, 360, 40), lambda: scenes.teapot(970, 543, 60, logo_size=64)],
                         ids=["map", "dense", "chunked", "teapot_odd"])
def test_column_bands_and_rectangles_match_the_full_frame(make):
    """rxc_frame.band_x0/x1: a rank of a column split renders columns [x0, x1) (x0 on a GPU tile) into a buffer of that
    width; rectangles combine both bands.  Pixels, owner ids and depth equal the full frame's, with and without planes."""
    def render(**kw):
        cfg = make()   # fresh scene per call (rasterize() appends the chunk lights every time)
        return render_gpu(cfg.rasterizer(2), cfg.scene, cfg.assets, cfg.width, cfg.height, cfg.tile_size, **kw)

    cfg = make()
    w, h = cfg.width, cfg.height
    full = render(planes=True)
    for (y0, y1, x0, x1) in ((0, h, 0, 96), (0, h, 96, 352), (0, h, 352, w), (64, h - 30, 160, w - 7), (32, 64, 32, 64)):
        px = np.zeros((y1 - y0, x1 - x0, 4), dtype=np.uint8)
        ow = np.zeros((y1 - y0, x1 - x0), dtype=np.uint32)
        dp = np.zeros((y1 - y0, x1 - x0), dtype=np.float32)
        c2 = make()
        c2.rasterizer(2).rasterize(c2.scene, px, w, h, c2.tile_size, c2.assets, owner=ow, depth=dp, band=(y0, y1, x0, x1))
        assert np.array_equal(px, full[0][y0:y1, x0:x1]) and np.array_equal(ow, full[1][y0:y1, x0:x1])
        assert np.array_equal(dp.view(np.uint32), full[2][y0:y1, x0:x1].view(np.uint32))
        c3 = make()
        fast = np.zeros((y1 - y0, x1 - x0, 4), dtype=np.uint8)
        c3.rasterizer(2).rasterize(c3.scene, fast, w, h, c3.tile_size, c3.assets, band=(y0, y1, x0, x1))
        assert np.array_equal(fast, full[0][y0:y1, x0:x1])


def test_small_triangle_pass_long_tile_lists():
    """Tiles whose binned list holds >= 128 records take the thread-per-record pass of k_raster (small triangles through a
    shared-memory atomicMin on (z, ordinal), the rest compacted for the warp walk): a quarter of a million triangles on a
    640x360 frame (about a thousand per tile), then the same kind of frame with alpha-holed textures and with batches
    drawn twice (exact depth ties: the first drawn must keep the pixel)."""
    from rusterix_b200 import Assets, DeviceContext, Tile
    st = _run(scenes.dense(640, 360, 40, patches=16), what="dense 16x16 patches on 640x360")
    assert DeviceContext.get(0).stats().last_binned_refs > 128 * 200
    cfg = scenes.dense(480, 270, 40, patches=10)
    cfg.assets = Assets.default().textures([Tile.from_texture(scenes.tex_terrain(i, holes=True)) for i in range(16)])
    cfg.sample_mode = SampleMode.Nearest
    cfg.scene.d3_static.extend(cfg.scene.d3_static[:40])
    cfg.scene.mark_dirty()
    _run(cfg, what="dense, holed textures, duplicated batches")


def test_small_triangle_pass_compacted_list_overflow():
    """More than 2048 records that are NOT small in one tile: the compacted shared list overflows and the warp walk reads
    the tile's whole list again, skipping the records the thread-per-record pass took."""
    cfg = scenes.cube(160, 120, 40, logo_size=16)
    rng = np.random.default_rng(11)
    verts, tris = [], []
    for i in range(18000):   # 16384 triangles and more select the kernel variant that has the pass
        small = i % 5 == 0 or i >= 6000
        cx, cy = rng.uniform(-0.12, 0.12, 2) if i < 6000 else rng.uniform(-1.0, 1.0, 2)
        s = rng.uniform(0.004, 0.05) if small else rng.uniform(0.1, 0.2)
        z = rng.uniform(-0.3, 0.3)
        base = len(verts)
        a = rng.uniform(0, 2 * math.pi)
        for k in range(3):
            verts.append((cx + s * math.cos(a + 2.1 * k), cy + s * math.sin(a + 2.1 * k), z, 1.0))
        tris.append((base, base + 1, base + 2))
    cfg.scene.d3_static.append(_tri_batch(verts, tris))
    cfg.scene.mark_dirty()
    _run(cfg, what="6000 overlapping triangles in a few tiles")
