#!/usr/bin/env python
"""Aggregates the per-line instruction counts of sass_lines.py by enclosing function of rx_kernels.cu /
rx_device.cuh.  usage: hot_regions.py <report.ncu-rep> <kernel-substring> <pixels>"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, kernel, pixels = sys.argv[1], sys.argv[2], float(sys.argv[3])
out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sass_lines.py"), rep, kernel, "100000"], capture_output=True, text=True).stdout
marks = {}
for fn in ("rx_kernels.cu", "rx_device.cuh"):
    m = []
    for i, l in enumerate(open(os.path.join(ROOT, "rusterix_b200", "csrc", fn)).read().splitlines(), 1):
        g = re.match(r"(?:static\s+)?(?:template.*>\s*)?__(?:device|global|host)__.*?\b(\w+)\s*\(", l)
        if g and not l.startswith(" "):
            m.append((i, g.group(1)))
    marks[fn] = m
def region(f, l):
    if f not in marks:
        return f
    name = f + ":top"
    for a, n in marks[f]:
        if l >= a:
            name = n
        else:
            break
    return name
agg = collections.Counter()
tot = 0
for line in out.splitlines():
    g = re.match(r"total warp instructions (\d+)", line)
    if g:
        tot = int(g.group(1))
    g = re.match(r"\s*([\d.]+)% inst\s+([\d.]+)% samp\s+\('([^']+)', (\d+)\)", line)
    if g:
        agg[region(g.group(3), int(g.group(4)))] += float(g.group(1))
print(f"total warp instructions {tot}  = {tot * 32 / pixels:.0f} thread-instructions per pixel")
for k, v in agg.most_common(25):
    print(f"{v:6.1f}%  {v / 100 * tot * 32 / pixels:7.1f} instr/px  {k}")
