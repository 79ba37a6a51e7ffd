"""The oracle against the committed golden digests (tests/golden/oracle_frames.json, made by
tests/golden/make_golden.py) and against itself across thread counts."""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import make_golden  # noqa: E402

import oracle_ffi  # noqa: E402

GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "oracle_frames.json")))


@pytest.mark.parametrize("case", sorted(make_golden.CASES))
def test_oracle_matches_golden(case):
    assert make_golden.digest(case) == GOLD[case]


def test_oracle_threading_does_not_change_results():
    cfg, frame = make_golden.CASES["map_320x180"]()
    r = cfg.rasterizer(frame)
    a = oracle_ffi.rasterize(r, cfg.scene, cfg.assets, cfg.width, cfg.height, cfg.tile_size, n_threads=1)
    b = oracle_ffi.rasterize(r, cfg.scene, cfg.assets, cfg.width, cfg.height, cfg.tile_size, n_threads=7)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
