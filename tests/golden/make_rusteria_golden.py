#!/usr/bin/env python
"""Golden vectors for the Rusteria VM row, taken from OUTPUTS OF THE REFERENCE ITSELF.

`rusteria/examples/{wood,marble,wood_ring}.png` were written by the reference's own CLI (`rsia <name>.rusteria`,
rsia/src/main.rs: 800x800, Rusteria::shade -> RenderBuffer::save) from the shaders next to them, reading the pattern
textures embedded in the crate (`rusteria/embedded/{fbm_perlin,value}.png`).  This script copies the two pattern PNGs
(inputs) and every 4th pixel of every 4th row of the three images (200x200 outputs) into tests/golden/rusteria/.
tests/test_rusteria_golden.py re-runs the hand-lowered programs (rusterix_b200/scenes.py: shader_wood / shader_marble /
shader_wood_ring) on exactly those pixels through the oracle's VM (bit-exact) and the device VM.

The seven pattern textures embedded in the crate are themselves OUTPUTS of the reference VM: `make_textures.rusteria` at
the root of the reference generated them (`iterate(tex, "make_value_noise"); save(tex, "rusteria/embedded/value.png")`
...).  They are copied too (the small ones whole, fbm_value and perlin as every 2nd pixel), as goldens for the
hand-lowered generator programs of tests/rusteria_programs.py (For / If / Return / nested calls / swizzled assignment).
Run in the build container, where /root/reference exists:  python tests/golden/make_rusteria_golden.py"""
import os
import shutil

import numpy as np
from PIL import Image

REF = "/root/reference/rusteria"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "rusteria")
STRIDE = 4

if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    for name in ("fbm_perlin", "value"):
        shutil.copyfile(os.path.join(REF, "embedded", name + ".png"), os.path.join(OUT, name + ".png"))
    for name in ("wood", "marble", "wood_ring"):
        im = np.asarray(Image.open(os.path.join(REF, "examples", name + ".png")).convert("RGB"))
        assert im.shape == (800, 800, 3)
        Image.fromarray(np.ascontiguousarray(im[::STRIDE, ::STRIDE])).save(os.path.join(OUT, name + "_every4th.png"), optimize=True)
    for name in ("bricks", "tiles", "blocks"):
        shutil.copyfile(os.path.join(REF, "embedded", name + ".png"), os.path.join(OUT, name + ".png"))
    for name in ("fbm_value", "perlin"):
        im = np.asarray(Image.open(os.path.join(REF, "embedded", name + ".png")).convert("RGB"))
        assert im.shape == (512, 512, 3) and (im[..., 0] == im[..., 1]).all() and (im[..., 1] == im[..., 2]).all()
        Image.fromarray(np.ascontiguousarray(im[::2, ::2, 0])).save(os.path.join(OUT, name + "_every2nd.png"), optimize=True)
    print("wrote", sorted(os.listdir(OUT)))
