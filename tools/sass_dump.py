#!/usr/bin/env python
"""Dumps the SASS of one kernel of the in-tree library annotated with CUDA source lines.
usage: sass_dump.py <kernel-substring> > out.txt"""
import os, re, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
kernel = sys.argv[1]
with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "rusterix_b200", "librxcuda.so")], cwd=td, capture_output=True)
    cubin = [f for f in os.listdir(td) if f.startswith("rx_kernels")][0]
    txt = subprocess.run(["nvdisasm", "-g", "-c", cubin], cwd=td, capture_output=True, text=True).stdout
on, cur, last = False, None, None
for line in txt.splitlines():
    if line.startswith("//---") and ".text." in line:
        on = kernel in line
        continue
    if not on:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', line)
    if m:
        cur = f"{os.path.basename(m.group(1))}:{m.group(2)}{m.group(3)}"
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m:
        if cur != last:
            print(f"## {cur}")
            last = cur
        print(f"  {m.group(1)}  {m.group(2)}")
