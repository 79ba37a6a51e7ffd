#!/usr/bin/env python
"""Writes the scene dumps (format RXH1) the Rust harness under tests/rust_harness renders with the UNMODIFIED reference:
BASELINE.json configs A (both sizes), B (three orbit frames) and C, on the reference's own mesh and textures, plus
`vek` probe operands (Mat4*Vec4, Mat4*Mat4, inverted) taken from those scenes' matrices and vertices.
usage: rust_harness_export.py [outdir = tests/rust_harness/cases]
Then, on a machine with cargo:  cd tests/rust_harness && cargo run --release -- cases/*.rxh
and copy cases/*.rxo to tests/golden/rust/ -- tests/test_rust_reference.py picks them up."""
import os
import struct
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rusterix_b200 import scenes  # noqa: E402
from rusterix_b200.marshal import submission_order  # noqa: E402

N_PROBES = 256


def cases():
    return [("cube_800x600_ts200", scenes.cube(800, 600, 200, real=True), 0), ("cube_2000x2000_ts40", scenes.cube(2000, 2000, 40, real=True), 0),
            ("teapot_1080_f0", scenes.teapot(1920, 1080, 60, real=True), 0), ("teapot_1080_f21", scenes.teapot(1920, 1080, 60, real=True), 21),
            ("teapot_1080_f47", scenes.teapot(1920, 1080, 60, real=True), 47), ("map_4k", scenes.map_config(3840, 2160, 40, real=True), 0),
            ("sweep_1080_f1024", scenes.sweep(1920, 1080, 40, real=True), 1024)]


def col_major(m):
    return np.asarray(m, dtype=np.float32).reshape(4, 4).T.reshape(-1)   # vek Mat4.cols


def dump(cfg, frame):
    r = cfg.rasterizer(frame)
    if cfg.scene.chunks or r.projection_matrix_2d is not None:
        raise ValueError("the harness covers scenes without chunks / 2D matrix")
    out = [b"RXH1", struct.pack("<4I", cfg.width, cfg.height, cfg.tile_size, int(cfg.sample_mode))]
    amb = r.ambient_color
    out.append(struct.pack("<I4f", 1 if amb is not None else 0, *(amb or (0, 0, 0, 0))))
    out += [col_major(r.view_matrix).tobytes(), col_major(r.projection_matrix).tobytes()]
    tiles = cfg.assets.tile_list
    out.append(struct.pack("<I", len(tiles)))
    for t in tiles:
        tx = t.textures[0]
        out += [struct.pack("<2I", tx.width, tx.height), tx.data.tobytes()]
    lights = cfg.scene.all_lights()
    out.append(struct.pack("<I", len(lights)))
    for l in lights:
        out.append(struct.pack("<I19f2I", int(l.light_type), *l.position, *l.color, l.intensity, l.start_distance, l.end_distance, l.flicker,
                               *l.direction, l.cone_angle, *l.normal, l.width, l.height, 1 if l.emitting else 0, 1 if l.from_linedef else 0))
    b3, b2 = submission_order(cfg.scene)
    out.append(struct.pack("<I", len(b3)))
    probes_v = []
    for b, _pass, _chunk in b3:
        v = np.asarray(b.vertices, dtype=np.float32).reshape(-1, 4)
        idx = np.asarray(b.indices, dtype=np.uint32).reshape(-1, 3)
        uv = np.asarray(b.uvs, dtype=np.float32).reshape(-1, 2)
        nr = None if b.normals is None or len(b.normals) == 0 else np.asarray(b.normals, dtype=np.float32).reshape(-1, 3)
        src = b.source_
        out.append(struct.pack("<6I4B2I", len(v), len(idx), int(b.repeat_mode_), int(b.cull_mode_), int(src.kind) if int(src.kind) <= 3 else 0, int(src.index),
                               *[int(c) for c in src.pixel], 1 if b.receives_light_ else 0, 0 if nr is None else 1))
        out += [col_major(b.transform_3d).tobytes(), v.tobytes(), idx.tobytes(), uv.tobytes()]
        if nr is not None:
            out.append(nr.tobytes())
        probes_v.append((np.asarray(b.transform_3d, dtype=np.float32).reshape(4, 4), v))
    out.append(struct.pack("<I", len(b2)))
    for b, _chunk in b2:
        v = np.asarray(b.vertices, dtype=np.float32).reshape(-1, 2)
        idx = np.asarray(b.indices, dtype=np.uint32).reshape(-1, 3)
        uv = np.asarray(b.uvs, dtype=np.float32).reshape(-1, 2)
        src = b.source_
        out.append(struct.pack("<4I4BI", len(v), len(idx), int(src.kind) if int(src.kind) <= 3 else 0, int(src.index), *[int(c) for c in src.pixel],
                               1 if b.receives_light_ else 0))
        out += [v.tobytes(), idx.tobytes(), uv.tobytes()]
    # vek probes: view*model applied to real vertices, proj applied to the results, the matrix products and inverses
    view, proj = np.asarray(r.view_matrix, np.float32).reshape(4, 4), np.asarray(r.projection_matrix, np.float32).reshape(4, 4)
    rng = np.random.default_rng(0x52555354 + frame)
    mv, mm, mi = [], [], []
    for k in range(N_PROBES):
        model, verts = probes_v[k % len(probes_v)]
        vm = (view.astype(np.float64) @ model.astype(np.float64)).astype(np.float32)
        p = verts[int(rng.integers(len(verts)))]
        mat = (vm, proj, view, model)[k % 4]
        mv.append((mat, p if k % 4 != 1 else (vm.astype(np.float64) @ p.astype(np.float64)).astype(np.float32)))
        mm.append(((proj, view, proj, view)[k % 4], (view, model, vm, view)[k % 4]))
        mi.append((view, proj, vm, model)[k % 4] + (rng.standard_normal((4, 4)) * 1e-3 * (k >= 4)).astype(np.float32))
    out.append(struct.pack("<I", N_PROBES))
    for m, v in mv:
        out += [col_major(m).tobytes(), np.asarray(v, np.float32).tobytes()]
    for a, b in mm:
        out += [col_major(a).tobytes(), col_major(b).tobytes()]
    for a in mi:
        out.append(col_major(a).tobytes())
    return b"".join(out)


def main():
    outdir = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tests", "rust_harness", "cases")
    os.makedirs(outdir, exist_ok=True)
    for name, cfg, frame in cases():
        data = dump(cfg, frame)
        with open(os.path.join(outdir, name + ".rxh"), "wb") as fh:
            fh.write(data)
        print(name, len(data), "bytes")


if __name__ == "__main__":
    main()
