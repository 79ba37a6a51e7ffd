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
