#!/usr/bin/env python
"""Writes one workload's entry of profiles/ncu_summary.json (what bench.py quotes as roofline.traffic / roofline.sm_issue) and the
condensed text profile from an `ncu --set full` report of k_raster.
usage: update_ncu_summary.py <workload> <report.ncu-rep> <frames per launch> <pixels per frame> <version string> <out.txt>"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import ncu_summary  # noqa: E402

workload, rep, frames, pixels, version, out_txt = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), sys.argv[5], sys.argv[6]
rows = [r for r in ncu_summary.summarize(rep) if "k_raster" in r.get("Kernel Name", "")]
row = rows[-1]


def num(key):
    v, unit = row[key].split()[0], (row[key].split() + [""])[1]
    return float(v) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}.get(unit, 1.0)


rd, wr = num("dram__bytes_read.sum"), num("dram__bytes_write.sum")
winst = int(num("smsp__inst_executed.sum"))
path = os.path.join(ROOT, "profiles", "ncu_summary.json")
summary = json.load(open(path))
summary[workload] = {
    "k_raster_dram_bytes_per_launch": int(rd + wr), "frames_per_launch": frames, "version": version,
    "smsp_issue_active_pct": round(num("smsp__issue_active.avg.pct_of_peak_sustained_active"), 2), "warp_instructions": winst,
    "thread_instructions_per_pixel": int(round(winst * 32.0 / (frames * pixels))), "kernel_ms_under_ncu": round(num("gpu__time_duration.sum"), 4),
    "dram_read_bytes": int(rd),
}
json.dump(summary, open(path, "w"), indent=1)
with open(out_txt, "w") as f:
    for k, v in row.items():
        f.write(f"{k:90s} {v}\n")
print(json.dumps(summary[workload], indent=1))
