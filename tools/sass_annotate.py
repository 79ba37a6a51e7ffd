#!/usr/bin/env python
"""Annotated SASS of the profiled kernel: per instruction the executed warp count (per tile-warp when
--per N is given), average active threads, source line.  usage: sass_annotate.py <report.ncu-rep> <kernel> [--per N]"""
import csv, io, os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import sass_lines

rep, kernel = sys.argv[1], sys.argv[2]
per = float(sys.argv[sys.argv.index("--per") + 1]) if "--per" in sys.argv else 1.0
page = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(page)))
pname = rows[0][1]
hdr = rows[1]
ci, ct, cs = hdr.index("Instructions Executed"), hdr.index("Avg. Threads Executed"), hdr.index("Source")
body = rows[2:]
lines = sass_lines.disasm_lines(kernel, len(body), pname)
for i, (r, ln) in enumerate(zip(body, lines)):
    n = int(r[ci] or 0)
    print(f"{i:5d} {n / per:10.2f} {r[ct]:>5s}  {ln[0][:14] if ln else '?':14s}:{ln[1] if ln else 0:<5d} {r[cs].strip()}")
