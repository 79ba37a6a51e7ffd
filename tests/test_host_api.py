"""Host-side mirror of the reference API: builders, loaders, cameras, marshalling and the multi-GPU
sharding helpers (world_size-2 gloo on CPU).  No GPU, no compute calls into the CUDA library."""
import math
import os
import sys

import numpy as np
import pytest

from rusterix_b200 import (Assets, Batch2D, Batch3D, CullMode, D3FirstPCamera, D3OrbitCamera, Light, LightType, PixelSource,
                           Rasterizer, RepeatMode, SampleMode, Scene, Texture, Tile, marshal, mgpu, scenes, vekmath)
from rusterix_b200.wavefront import Wavefront


def test_from_box_layout():
    b = Batch3D.from_box(-0.5, -0.5, -0.5, 1.0, 1.0, 1.0)
    assert b.vertices.shape == (24, 4) and b.indices.shape == (12, 3) and b.uvs.shape == (24, 2)
    assert b.indices.tolist()[:4] == [[0, 1, 2], [0, 2, 3], [4, 6, 5], [4, 7, 6]]
    assert b.indices.tolist()[-2:] == [[20, 23, 22], [20, 22, 21]]
    assert b.vertices[6].tolist() == [0.5, 0.5, 0.5, 1.0] and b.uvs[0].tolist() == [0.0, 1.0]
    assert b.cull_mode_ == CullMode.Off and b.repeat_mode_ == RepeatMode.ClampXY and b.source_ == PixelSource.Off
    n = b.with_computed_normals().normals
    assert np.allclose(np.linalg.norm(n, axis=1), 1.0, atol=1e-6)
    assert n[0].tolist() == [0.0, 0.0, 1.0]  # (p1-p0) x (p2-p0) of the first face, as the reference computes it


def test_from_rectangle_and_builders():
    r = Batch2D.from_rectangle(1.0, 2.0, 10.0, 20.0)
    assert r.vertices.tolist() == [[1, 2], [1, 22], [11, 22], [11, 2]] and r.indices.tolist() == [[0, 1, 2], [0, 2, 3]]
    assert r.uvs.tolist() == [[0, 0], [0, 1], [1, 1], [1, 0]]
    r.receives_light(False).source(PixelSource.StaticTileIndex(3)).repeat_mode(RepeatMode.RepeatX)
    assert (r.receives_light_, r.source_.kind, r.source_.index, r.repeat_mode_) == (False, 1, 3, RepeatMode.RepeatX)


def test_wavefront_parse_defaults_uv_to_xy():
    obj = "# c\nv 0 0 0\nv 1.5 0 0\nv 0 2 0\nvn 0 0 1\nf 1/1/1 2//1 3\n"
    b = Wavefront.parse_string(obj).to_batch()
    assert b.indices.tolist() == [[0, 1, 2]] and b.vertices[1].tolist() == [1.5, 0, 0, 1]
    assert b.uvs.tolist() == [[0, 0], [1.5, 0], [0, 2]] and len(b.normals) == 0
    b2 = Wavefront.parse_string(obj + "vt 0.1 0.2\nvt 0.3 0.4\nvt 0.5 0.6\n").to_batch()
    assert np.allclose(b2.uvs, [[0.1, 0.2], [0.3, 0.4], [0.5, 0.6]])


def test_cameras_produce_rh_zero_to_one_projection():
    cam = D3OrbitCamera.new()
    cam.set_parameter_f32("distance", 1.5)
    eye = cam.eye_position()
    assert abs(np.linalg.norm(eye) - 1.5) < 1e-6
    v, p = cam.view_matrix(), cam.projection_matrix(800.0, 600.0)
    c = v @ np.array([0, 0, 0, 1], np.float32)  # the orbit centre sits on the -z axis of view space
    assert abs(c[0]) < 1e-6 and abs(c[1]) < 1e-6 and abs(c[2] + 1.5) < 1e-5
    for z, expect in ((-0.01, 0.0), (-100.0, 1.0)):
        q = p @ np.array([0, 0, z, 1], np.float32)
        assert abs(q[2] / q[3] - expect) < 1e-4
    fp = D3FirstPCamera.new()
    fp.position = np.array([1, 2, 3], np.float32)
    fp.center = np.array([1, 2, 4], np.float32)
    inv = vekmath.inverted(fp.view_matrix())
    assert np.allclose(inv[:3, 3], [1, 2, 3], atol=1e-6)  # camera_pos = inverse_view.cols[3]


def test_marshal_scene_order_and_usize_indices():
    cfg = scenes.map_config(64, 36, 40, logo_size=8)
    m = marshal.marshal_scene(cfg.scene)
    s = m.struct
    assert s.n_batches3d == 5 and s.n_batches2d == 1 and s.n_lights == 1
    assert [s.batches3d[i].source_index for i in range(5)] == [1, 2, 3, 4, 5]
    assert s.batches3d[0].n_triangles == 10 and s.batches3d[2].n_triangles == 4 and s.batches3d[4].n_triangles == 12
    m8 = marshal.marshal_scene(cfg.scene, index_bytes=8)
    assert m8.struct.batches3d[0].index_bytes == 8
    f = marshal.make_frame(cfg.rasterizer(), cfg.scene, 64, 36, 40)
    assert (f.width, f.height, f.tile_size, f.d2_active, f.d3_active, f.has_ambient) == (64, 36, 40, 1, 1, 1)
    assert f.view[12] == cfg.rasterizer().view_matrix[0, 3]  # column-major: m[c*4+r]


def test_algorithmic_bytes_matches_survey_formula():
    cfg = scenes.map_config(3840, 2160, 40)
    v, t = cfg.counts()
    tex = 1024 * 1024 * 4 + 4 * 64 * 64 * 4 - 64 * 64 * 4 + 64 * 80 * 4 + 256 * 128 * 4  # logo, brick, panel, fence, floor, sky
    assert (v, t) == (60, 30)
    assert cfg.algorithmic_bytes() == 3840 * 2160 * 4 + v * 36 + t * 12 + tex + 1 * 72 + 256


def test_short_pixel_buffer_is_rejected_before_any_device_work():
    cfg = scenes.cube(64, 64, 40, logo_size=8)
    from rusterix_b200.rasterizer import _buffer_pointer

    with pytest.raises(ValueError):
        _buffer_pointer(np.zeros(100, np.uint8), 64 * 64 * 4)


def test_band_and_frame_sharding():
    for h in (4320, 2160, 1080, 360, 17):
        for w in (1, 2, 4, 8):
            bands = mgpu.all_bands(h, w)
            assert bands[0][0] == 0 and bands[-1][1] == h
            for (a0, a1), (b0, b1) in zip(bands[:-1], bands[1:]):
                assert a1 == b0 and (a0 % 16 == 0 or a0 == h)
    assert mgpu.shard_frames(10, 1, 4) == [1, 5, 9]
    assert sorted(sum((mgpu.shard_frames(4096, r, 8) for r in range(8)), [])) == list(range(4096))


def _gloo_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        h, w = 50, 8
        full = (torch.arange(h * w * 4, dtype=torch.int64) % 251).to(torch.uint8).reshape(h, w, 4)
        y0, y1 = mgpu.band_for_rank(h, rank, world)
        got = mgpu.gather_bands_to_rank0(full[y0:y1].clone(), h, w, rank, world)
        ok = True
        if rank == 0:
            ok = ok and torch.equal(got, full)
        F = 3
        frames = torch.stack([torch.full((4, 4, 4), 10 * i + rank, dtype=torch.uint8) for i in range(F)])
        allf = mgpu.gather_frames_to_rank0(frames, rank, world)
        if rank == 0:
            ids = mgpu.shard_frames(F * world, 0, world)
            for g in range(F * world):
                ok = ok and int(allf[g, 0, 0, 0]) == 10 * (g // world) + (g % world)
        # cost-balanced bands: both ranks see both times, compute the same boundaries, and the ragged gather
        # reassembles the frame
        bal = mgpu.BandBalancer(h, world, align=2)
        for _ in range(3):
            y0, y1 = bal.band(rank)
            ms = 1.0 + 0.1 * sum(1 + 9 * (y >= 40) for y in range(y0, y1))   # rows 40.. are ten times as expensive
            times = mgpu.all_gather_times(ms)
            ok = ok and len(times) == world and abs(times[rank] - ms) < 1e-9
            bal.update(times)
        bands = bal.bands()
        ok = ok and bands[0][1] == bands[1][0] and bands[0][1] > 25   # the boundary moved towards the expensive rows
        y0, y1 = bands[rank]
        got = mgpu.gather_ragged_bands_to_rank0(full[y0:y1].clone(), bands, h, w, rank, world)
        if rank == 0:
            ok = ok and torch.equal(got, full)
        # column bands: tile-aligned left edges, together the whole width, reassembled on rank 0
        wide = (torch.arange(6 * 100 * 4, dtype=torch.int64) % 253).to(torch.uint8).reshape(6, 100, 4)
        x0, x1 = mgpu.column_band_for_rank(100, rank, world)
        got = mgpu.gather_column_bands_to_rank0(wide[:, x0:x1].clone(), 6, 100, rank, world)
        if rank == 0:
            ok = ok and torch.equal(got, wide)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_gather_to_rank0_world2_gloo():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    ps = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def test_column_bands_tile_the_width():
    for width, world in ((7680, 8), (1920, 8), (100, 3), (33, 4), (3840, 1)):
        bands = [mgpu.column_band_for_rank(width, r, world) for r in range(world)]
        assert bands[0][0] == 0 and bands[-1][1] == width and all(a[1] == b[0] for a, b in zip(bands, bands[1:]))
        assert all(x0 % 32 == 0 for x0, x1 in bands if x1 > x0)   # ranks beyond the width get an empty band and idle


def test_band_balancer_converges_on_a_skewed_cost_profile():
    """mgpu.BandBalancer: contiguous, tile-aligned bands that cover the frame; on a profile like the 1M-triangle
    view (cheap sky rows, an expensive horizon) the slowest rank ends close to the balanced optimum."""
    H, align = 4320, 32
    cost = np.full(H, 0.057e-3); cost[1632:2176] = 0.73e-3; cost[2176:2720] = 0.54e-3; cost[2720:3264] = 0.37e-3; cost[3264:] = 0.31e-3
    for world in (1, 2, 3, 4, 8):
        bal = mgpu.BandBalancer(H, world, align)
        first = None
        for _ in range(14):
            bands = bal.bands()
            assert bands[0][0] == 0 and bands[-1][1] == H
            assert all(a[1] == b[0] for a, b in zip(bands, bands[1:])) and all(y0 <= y1 for y0, y1 in bands)
            assert all(y % align == 0 for y0, y1 in bands[:-1] for y in (y0, y1))
            times = [0.14 + float(cost[y0:y1].sum()) for y0, y1 in bands]
            first = first if first is not None else max(times)
            bal.update(times)
        bands = bal.use_best()
        worst = max(0.14 + float(cost[y0:y1].sum()) for y0, y1 in bands)
        optimum = 0.14 + float(cost.sum()) / world
        assert worst <= first + 1e-12 and worst <= 1.12 * optimum, (world, worst, optimum)
    with pytest.raises(ValueError):
        mgpu.BandBalancer(H, 4).update([1.0, 2.0])


def test_band_balancer_over_tile_columns():
    """The same balancer over the tile COLUMNS of the 8K frame (bench.py band_split's fourth split): column bands cut every tile
    row alike, what differs is the triangles a band holds -- a mild bulge in the middle of the frame."""
    W, align = 7680, 32
    x = np.arange(W)
    cost = (0.08 + 0.07 * np.exp(-((x - 3600.0) / 1500.0) ** 2)) / 960.0   # ms per column: 0.08 ms per eighth at the rim, 0.15 in the middle
    for world in (2, 4, 8):
        bal = mgpu.BandBalancer(W, world, align)
        first = None
        for _ in range(12):
            bands = bal.bands()
            assert bands[0][0] == 0 and bands[-1][1] == W and all(a[1] == b[0] for a, b in zip(bands, bands[1:]))
            assert all(v % align == 0 for x0, x1 in bands[:-1] for v in (x0, x1))
            times = [0.065 + float(cost[x0:x1].sum()) * 8.0 / world for x0, x1 in bands]
            first = first if first is not None else max(times)
            bal.update(times)
        bands = bal.use_best()
        worst = max(0.065 + float(cost[x0:x1].sum()) * 8.0 / world for x0, x1 in bands)
        optimum = 0.065 + float(cost.sum()) * 8.0 / world / world
        assert worst <= first + 1e-12 and worst <= 1.05 * optimum, (world, worst, optimum)


def test_bresenham_membership_closed_form_matches_the_serial_walk():
    """The device decides per pixel whether rasterize_line_bresenham (rasterizer.rs:1777-1821) plots it
    (line_covers in rx_kernels.cu).  Same closed form, checked against the serial walk for every
    segment from the origin inside a 29x29 grid and for random long segments."""
    import random

    def serial(x0, y0, x1, y1):
        dx, dy = abs(x1 - x0), abs(y1 - y0)
        sx, sy = (1 if x0 < x1 else -1), (1 if y0 < y1 else -1)
        err, x, y, pts = dx - dy, x0, y0, set()
        while x != x1 or y != y1:
            pts.add((x, y))
            e2 = err * 2
            if e2 > -dy:
                err -= dy
                x += sx
            if e2 < dx:
                err += dx
                y += sy
        return pts

    def covers(x0, y0, x1, y1, x, y):
        a, b = abs(x1 - x0), abs(y1 - y0)
        u = (x - x0) if x0 < x1 else (x0 - x)
        v = (y - y0) if y0 < y1 else (y0 - y)
        if u < 0 or v < 0 or u > a or v > b or (u == a and v == b):
            return False
        if a >= b:
            return a != 0 and v == (2 * b * u + a - 1) // (2 * a)
        return u == (2 * a * v + b - 1) // (2 * b)

    R = 14
    for x1 in range(-R, R + 1):
        for y1 in range(-R, R + 1):
            pts = serial(0, 0, x1, y1)
            for x in range(-R - 1, R + 2):
                for y in range(-R - 1, R + 2):
                    assert covers(0, 0, x1, y1, x, y) == ((x, y) in pts), (x1, y1, x, y)
    rng = random.Random(7)
    for _ in range(200):
        x0, y0, x1, y1 = [rng.randint(-400, 400) for _ in range(4)]
        pts = serial(x0, y0, x1, y1)
        assert all(covers(x0, y0, x1, y1, x, y) for x, y in pts)
        for _ in range(300):
            x, y = rng.randint(min(x0, x1) - 2, max(x0, x1) + 2), rng.randint(min(y0, y1) - 2, max(y0, y1) + 2)
            assert covers(x0, y0, x1, y1, x, y) == ((x, y) in pts)


def test_submission_order_of_a_chunked_scene():
    """Per chunk: opacity batches, opaque batches, terrain; then static, dynamic, overlay
    (rasterizer.rs:314-405); 2D: per chunk batches2d, terrain2d; then static, dynamic (:501-553)."""
    from rusterix_b200 import marshal, scenes
    cfg = scenes.chunked_config(64, 64)
    b3, b2 = marshal.submission_order(cfg.scene)
    passes = [p for _b, p, _c in b3]
    chunks = [c for _b, _p, c in b3]
    n_chunk = sum(1 for c in chunks if c >= 0)
    assert chunks[:n_chunk] == sorted(chunks[:n_chunk]) and all(c == -1 for c in chunks[n_chunk:])
    for ci in range(4):
        ps = [p for p, c in zip(passes, chunks) if c == ci]
        assert ps == sorted(ps, reverse=True) and ps[0] == 4 and ps[-1] == 3   # opacity (4) before opaque/terrain (3)
    m = marshal.marshal_scene(cfg.scene, 4, cfg.assets)
    assert m.struct.n_chunks == 4 and m.struct.n_actor_tiles == 3   # hero/idle, hero/walk, chest/closed
    srcs = [(m.struct.batches3d[i].source_kind, m.struct.batches3d[i].source_index) for i in range(m.struct.n_batches3d)]
    assert (5, 0xFFFFFFFF) in srcs and (4, 0xFFFFFFFF) in srcs   # ItemTile("missing") and EntityTile("hero", 7)


def test_oracle_opacity_layer_and_surface_ids():
    """Known answers for the chunk path of the oracle: a wall with profile 7, a half-transparent pane
    in front of it with the same profile, and a back wall with profile 8.  Pixels under the pane skip
    the profile-7 wall (rasterizer.rs:1041-1047) and blend the pane over the back wall (:464-495)."""
    import oracle_ffi
    from rusterix_b200 import (Assets, Batch3D, Chunk, CullMode, PixelSource, Rasterizer, Scene, Texture, Tile)
    from rusterix_b200 import D3FirstPCamera

    def quad(z, x0, x1, colour, profile):
        v = [(x0, -1.0, z, 1.0), (x1, -1.0, z, 1.0), (x1, 1.0, z, 1.0), (x0, 1.0, z, 1.0)]
        b = Batch3D(v, [(0, 1, 2), (0, 2, 3)], [(0, 0), (1, 0), (1, 1), (0, 1)]).source(PixelSource.Pixel(colour)).cull_mode(CullMode.Off)
        return b.profile_id(profile).with_computed_normals()

    ch = Chunk((0, 0), 8)
    ch.batches3d_opacity.append(quad(-2.0, -0.5, 0.5, (200, 100, 0, 128), 7))
    ch.batches3d.append(quad(-3.0, -1.0, 1.0, (0, 255, 0, 255), 7))
    ch.batches3d.append(quad(-4.0, -2.0, 2.0, (0, 0, 255, 255), 8))
    scene = Scene()
    scene.chunks[(0, 0)] = ch
    cam = D3FirstPCamera.new()
    cam.position = np.array([0.0, 0.0, 0.0], dtype=np.float32)
    cam.center = np.array([0.0, 0.0, -1.0], dtype=np.float32)
    W = H = 64
    r = Rasterizer.setup(None, cam.view_matrix(), cam.projection_matrix(float(W), float(H))).ambient((1.0, 1.0, 1.0, 1.0))
    p, owner, depth = oracle_ffi.rasterize(r, scene, Assets.default(), W, H, 16)
    centre, side = p[H // 2, W // 2], p[H // 2, W // 2 + 12]
    # centre: pane over the BLUE back wall (green wall skipped); side: the green wall itself
    assert owner[H // 2, W // 2] == 3 * 4 and owner[H // 2, W // 2 + 12] == 3 * 2   # slots: batch2 tri0 / batch1 tri0 (+3 per tri capacity)
    assert side[1] > 150 and side[2] == 0
    assert centre[1] < 80 and centre[2] > 60 and centre[0] > 60 and centre[3] == 255


def test_empty_or_bad_bands_never_reach_the_abi_sentinel():
    """rxc_frame.band_y0 = band_y1 = 0 means "whole frame": an empty band must be an error on the host side, not a
    full-frame render into a band-sized buffer (a balancer may hand rank 0 an empty band)."""
    from rusterix_b200.rasterizer import _check_band

    _check_band(None, 640, 360); _check_band((0, 360), 640, 360); _check_band((32, 64, 0, 640), 640, 360)
    for bad in ((0, 0), (64, 64), (64, 32), (0, 361), (0, 32, 0, 0), (0, 32, 64, 64), (0, 32, 0, 641), (0,), (0, 1, 2)):
        with pytest.raises(ValueError):
            _check_band(bad, 640, 360)


def test_structural_scene_changes_invalidate_the_device_cache_key():
    """The reference re-projects `&mut scene` on every call; the device cache must notice batches that were appended,
    replaced or removed and dynamic tiles that were added, without an explicit mark_dirty()."""
    from rusterix_b200 import Batch2D, Batch3D, scenes

    scene = scenes.map_scene()
    k0 = scene.structure_key()
    assert k0 == scene.structure_key()
    scene.d3_dynamic.append(Batch3D.from_box(0, 0, 0, 1, 1, 1))
    k1 = scene.structure_key()
    assert k1 != k0
    scene.d3_dynamic = [Batch3D.from_box(0, 0, 0, 1, 1, 1)]          # the reference's per-frame reassignment
    k2 = scene.structure_key()
    assert k2 != k1
    scene.d2_dynamic.append(Batch2D.from_rectangle(0, 0, 4, 4))
    assert scene.structure_key() != k2
    k3 = scene.structure_key()
    scene.dynamic_textures.append(scenes.map_assets(16).tile_list[0])
    assert scene.structure_key() != k3


def test_marshalling_against_an_independent_walk_of_the_host_objects():
    """The oracle and the device both receive rxc_scene from marshal.py; a flattening / ordering mistake there would be
    invisible to every parity test.  This walks the chunked scene a second, deliberately dumb way -- nested loops
    written straight from src/rasterizer.rs:314-405 (3D), :501-553 (2D) and :219-223 (lights), no helper of marshal.py
    -- and compares what rxc_scene actually holds: batch order, pass tags, chunk indices, geometry pointers' contents,
    profile ids, resolved EntityTile / ItemTile sources and the light order after the per-call chunk-light append."""
    import ctypes as C

    from rusterix_b200 import marshal, scenes

    cfg = scenes.chunked_config(320, 180)
    scene, assets = cfg.scene, cfg.assets
    # what Rasterizer::rasterize does before anything else: chunk lights appended to dynamic_lights (:219-223)
    n_dyn_before = len(scene.dynamic_lights)
    for chunk in scene.chunks.values():
        scene.dynamic_lights.extend(chunk.lights)

    expect3, expect2 = [], []          # (batch object, pass tag, chunk index)
    ci = 0
    for _key, chunk in scene.chunks.items():
        for b in chunk.batches3d_opacity:
            expect3.append((b, 4, ci))                      # :318-329  d3_rasterize_opacity
        for b in chunk.batches3d:
            expect3.append((b, 3, ci))                      # :331-343
        if chunk.terrain_batch3d is not None:
            expect3.append((chunk.terrain_batch3d, 3, ci))  # :345-356
        ci += 1
    for b in scene.d3_static:
        expect3.append((b, 0, -1))                          # :359-371
    for b in scene.d3_dynamic:
        expect3.append((b, 1, -1))                          # :373-385
    for b in scene.d3_overlay:
        expect3.append((b, 2, -1))                          # :387-405
    ci = 0
    for _key, chunk in scene.chunks.items():
        for b in chunk.batches2d:
            expect2.append((b, ci))                         # :503-517
        if chunk.terrain_batch2d is not None:
            expect2.append((chunk.terrain_batch2d, ci))     # :519-531
        ci += 1
    for b in scene.d2_static:
        expect2.append((b, -1))                             # :534-543
    for b in scene.d2_dynamic:
        expect2.append((b, -1))                             # :545-553
    expect_lights = list(scene.lights) + list(scene.dynamic_lights)

    m = marshal.marshal_scene(scene, 4, assets)
    s = m.struct
    assert s.n_batches3d == len(expect3) and s.n_batches2d == len(expect2) and s.n_chunks == len(scene.chunks)
    assert len(scene.dynamic_lights) == n_dyn_before + sum(len(c.lights) for c in scene.chunks.values())

    def floats(ptr, n):
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_float)), (n,)).copy() if n else np.zeros(0, np.float32)

    def actor_bytes(index):
        t = s.actor_tiles[index].textures[0]
        return bytes(C.string_at(t.data, t.width * t.height * 4)), t.width, t.height

    for i, (b, tag, chunk_index) in enumerate(expect3):
        o = s.batches3d[i]
        assert (o.pass_, o.chunk) == (tag, chunk_index), i
        assert o.n_vertices == len(b.vertices) and o.n_triangles == len(b.indices), i
        assert np.array_equal(floats(o.vertices, o.n_vertices * 4), np.asarray(b.vertices, np.float32).reshape(-1)), i
        assert np.array_equal(floats(o.uvs, o.n_vertices * 2), np.asarray(b.uvs, np.float32).reshape(-1)), i
        assert (o.has_profile_id, o.profile_id) == ((0, 0) if b.profile_id_ is None else (1, b.profile_id_)), i
        assert np.array_equal(np.asarray(list(o.transform), np.float32).reshape(4, 4).T, np.asarray(b.transform_3d, np.float32)), i   # column-major
        src = b.source_
        assert o.source_kind == int(src.kind), i
        if src.name in ("EntityTile", "ItemTile"):       # resolved on the host: assets.entity_tiles[id].get_index(index) (:1130-1177)
            table = assets.entity_tiles if src.name == "EntityTile" else assets.item_tiles
            seq = table.get(src.ident)
            if seq is None or not (0 <= src.index < len(seq)):
                assert o.source_index == 0xFFFFFFFF, i
            else:
                tex = seq[src.index][1].textures[0]
                assert actor_bytes(o.source_index) == (tex.data.tobytes(), tex.width, tex.height), i
        elif src.name in ("StaticTileIndex", "DynamicTileIndex"):
            assert o.source_index == src.index, i
    for i, (b, chunk_index) in enumerate(expect2):
        o = s.batches2d[i]
        assert o.chunk == chunk_index and o.mode == int(b.mode) and o.n_vertices == len(b.vertices), i
        assert np.array_equal(floats(o.vertices, o.n_vertices * 2), np.asarray(b.vertices, np.float32).reshape(-1)), i
    assert s.n_lights == len(expect_lights)
    for i, l in enumerate(expect_lights):
        o = s.lights[i]
        assert o.light_type == int(l.light_type) and tuple(o.position) == tuple(np.float32(c) for c in l.position), i
        assert o.intensity == np.float32(l.intensity) and o.end_distance == np.float32(l.end_distance), i
    # chunk members: origin, size, sectors in order, terrain texture bytes
    for k, chunk in enumerate(scene.chunks.values()):
        c = s.chunks[k]
        assert tuple(c.origin) == tuple(chunk.origin) and c.size == chunk.size and c.n_occluded_sectors == len(chunk.occluded_sectors)
        for j, (bbox, occ) in enumerate(chunk.occluded_sectors):
            q = c.occluded_sectors[j]
            assert (tuple(q.min), tuple(q.max), q.occlusion) == (tuple(np.float32(v) for v in bbox.min), tuple(np.float32(v) for v in bbox.max), np.float32(occ))
        tt = c.terrain_texture.contents
        assert bytes(C.string_at(tt.data, tt.width * tt.height * 4)) == chunk.terrain_texture.data.tobytes()
