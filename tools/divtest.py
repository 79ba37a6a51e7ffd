#!/usr/bin/env python
"""Runs the device division self-test over several seeds and prints failing operand pairs."""
import os, sys, struct
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rusterix_b200 import DeviceContext
ctx = DeviceContext.get(0)
for seed in [12345] + list(range(1, int(sys.argv[1]) if len(sys.argv) > 1 else 8)):
    n = ctx.selftest_div(n_pairs=1 << 32, seed=seed)
    a, b = ctx.last_bad_pair
    fa, fb = struct.unpack("<f", struct.pack("<I", a))[0], struct.unpack("<f", struct.pack("<I", b))[0]
    print(seed, "mismatches", n, hex(a), hex(b), fa, fb)
