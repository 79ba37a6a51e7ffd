"""Minimal OBJ loader with the behaviour of the reference's Wavefront (src/wavefront.rs:34-101):
`v`, `vn`, `vt` and triangular `f` records; faces keep only the position index; `vn` is ignored by
`to_batch`; when the file has no `vt` the UV of a vertex defaults to its (x, y)."""
import os

import numpy as np


class Wavefront:
    def __init__(self, vertices, indices, normals, texture_coords):
        self.vertices = vertices
        self.indices = indices
        self.normals = normals
        self.texture_coords = texture_coords

    @staticmethod
    def parse(path_or_text) -> "Wavefront":
        if isinstance(path_or_text, (str, os.PathLike)) and "\n" not in str(path_or_text) and os.path.exists(path_or_text):
            with open(path_or_text, "r") as fh:
                text = fh.read()
        else:
            text = str(path_or_text)
        return Wavefront.parse_string(text)

    @staticmethod
    def parse_string(contents: str) -> "Wavefront":
        vertices, normals, tcs, indices = [], [], [], []
        for line in contents.splitlines():
            t = line.strip()
            if not t or t.startswith("#"):
                continue
            if t.startswith("v "):
                p = t.split()
                vertices.append((float(p[1]), float(p[2]), float(p[3]), 1.0))
            elif t.startswith("vn "):
                p = t.split()
                normals.append((float(p[1]), float(p[2]), float(p[3])))
            elif t.startswith("vt "):
                p = t.split()
                tcs.append((float(p[1]), float(p[2])))
            elif t.startswith("f "):
                p = t.split()
                indices.append(tuple(int(s.split("/")[0]) - 1 for s in p[1:4]))
        return Wavefront(vertices, indices, normals, tcs)

    def to_batch(self):
        from .types import Batch3D

        v = np.asarray(self.vertices, dtype=np.float32).reshape(-1, 4)
        uvs = v[:, :2].copy() if not self.texture_coords else np.asarray(self.texture_coords, dtype=np.float32)
        return Batch3D(v, self.indices, uvs)
