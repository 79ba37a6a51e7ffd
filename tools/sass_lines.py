#!/usr/bin/env python
"""Joins an ncu SASS source page (per-instruction executed counts / stall samples) with nvdisasm line
info of the in-tree library, and prints the hottest CUDA source lines of one kernel.
usage: sass_lines.py <report.ncu-rep> <kernel-substring> [top N]"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def disasm_functions(kernel):
    """{mangled name: [source line of every instruction]} for the functions whose name contains `kernel`."""
    with tempfile.TemporaryDirectory() as td:
        jit = os.environ.get("RX_SASS_CUBIN")   # a kernel the batch-shader JIT compiled (cache file: u32 name length, name, cubin)
        if jit:
            data = open(jit, "rb").read()
            n = int.from_bytes(data[:4], "little")
            cubin = "jit.cubin"
            open(os.path.join(td, cubin), "wb").write(data[4 + n:] if n < 512 and data[4 + n:8 + n] == b"\x7fELF" else data)
        else:
            subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "rusterix_b200", "librxcuda.so")], cwd=td, capture_output=True)
            cubin = [f for f in os.listdir(td) if f.startswith("rx_kernels")][0]
        txt = subprocess.run(["nvdisasm", "-g", "-c", cubin], cwd=td, capture_output=True, text=True).stdout
    funcs, cur, name = {}, None, None
    for line in txt.splitlines():
        if line.startswith("//---") and ".text." in line:
            name = line.split(".text.")[1].split()[0]
            if kernel not in name:
                name = None
            else:
                funcs[name] = []
            continue
        if name is None:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', line)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
            funcs[name].append(cur)
    return funcs


def disasm_lines(kernel, n_profiled, profiled_name):
    """The instantiation that was profiled: same instruction count, and the template arguments of the
    demangled name in the report (`k_raster<0, 0, 1>`) in the same order as the mangled one (`ILi0ELb0ELi1E`)."""
    funcs = disasm_functions(kernel)
    args = re.findall(r"\d+", profiled_name[profiled_name.find("<"):profiled_name.find(">") + 1]) if "<" in profiled_name else []
    cands = [n for n, l in funcs.items() if len(l) == n_profiled]
    if args:
        want = [n for n in cands if re.findall(r"L[ibjm](\d+)E", n) == args]
        cands = want or cands
    if len(cands) != 1:
        print(f"warning: {len(cands)} disassembled functions match {profiled_name!r} ({n_profiled} instructions)", file=sys.stderr)
    if not cands:
        return max(funcs.values(), key=len)
    print(f"joined with {cands[0]}", file=sys.stderr)
    return funcs[cands[0]]


def main():
    rep, kernel = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    page = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(page)))
    hdr = rows[1]
    ci, cs = hdr.index("Instructions Executed"), hdr.index("# Samples")
    body = rows[2:]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(io.StringIO(raw)))
    pname = rr[2][rr[0].index("Kernel Name")] if len(rr) > 2 and "Kernel Name" in rr[0] else ""
    lines = disasm_lines(kernel, len(body), pname)
    if len(body) != len(lines):
        print(f"warning: {len(body)} profiled instructions vs {len(lines)} disassembled", file=sys.stderr)
    inst, samp = collections.Counter(), collections.Counter()
    for r, ln in zip(body, lines):
        inst[ln] += int(r[ci] or 0)
        samp[ln] += int(r[cs] or 0)
    ti, tsamp = sum(inst.values()), sum(samp.values())
    src = {}
    print(f"total warp instructions {ti}, samples {tsamp}")
    for ln, n in inst.most_common(top):
        if ln is None:
            text = "?"
        else:
            if ln[0] not in src:
                p = os.path.join(ROOT, "rusterix_b200", "csrc", ln[0])
                src[ln[0]] = open(p).read().splitlines() if os.path.exists(p) else []
            text = src[ln[0]][ln[1] - 1].strip()[:110] if ln[1] - 1 < len(src[ln[0]]) else ""
        print(f"{100.0 * n / ti:5.1f}% inst {100.0 * samp[ln] / max(1, tsamp):5.1f}% samp  {ln}  {text}")


if __name__ == "__main__":
    main()
