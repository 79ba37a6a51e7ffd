"""The C-ABI library loads without a GPU, exports every symbol include/rxcuda.h declares, and the
ctypes mirror agrees with the header's struct layout (checked with a C probe compiled by gcc)."""
import ctypes as C
import os
import re
import subprocess
import sys
import tempfile

import pytest

from rusterix_b200 import _abi, _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "rxcuda.h")


def test_library_loads_and_exports_every_declared_symbol():
    lib = _lib.load()
    text = open(HEADER).read()
    declared = set(re.findall(r"\b(rxc_[a-z_0-9]+)\s*\(", text))
    declared -= {"rxc_ctx"}
    assert declared == {name for name, _, _ in _abi.EXPORTS}
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.rxc_abi_version() == _abi.RXC_ABI_VERSION
    assert lib.rxc_kernel_name(7).decode() == "k_raster"


def test_struct_layout_matches_header():
    structs = ["rxc_texture", "rxc_tile", "rxc_light", "rxc_batch3d", "rxc_batch2d", "rxc_sector", "rxc_chunk", "rxc_linedef", "rxc_mapmini", "rxc_scene", "rxc_frame", "rxc_stats", "rxc_mgpu_region", "rxc_projected3d"]
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', "int main(void){"]
    for s in structs:
        cls = getattr(_abi, s)
        lines.append(f'printf("{s} %zu\\n", sizeof({s}));')
        for fname, _ in cls._fields_:
            cname = "pass" if fname == "pass_" else fname
            lines.append(f'printf("{s}.{fname} %zu\\n", offsetof({s}, {cname}));')
    lines.append("return 0;}")
    with tempfile.TemporaryDirectory() as td:
        src = os.path.join(td, "probe.c")
        exe = os.path.join(td, "probe")
        open(src, "w").write("\n".join(lines))
        subprocess.run(["gcc", "-std=c11", "-o", exe, src], check=True)
        out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout
    got = dict(l.split() for l in out.strip().splitlines())
    for s in structs:
        cls = getattr(_abi, s)
        assert int(got[s]) == C.sizeof(cls), s
        for fname, _ in cls._fields_:
            assert int(got[f"{s}.{fname}"]) == getattr(cls, fname).offset, f"{s}.{fname}"


def test_no_device_is_an_error_code_not_a_crash():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    lib = _lib.load()
    h = C.c_void_p()
    assert lib.rxc_create(0, C.byref(h)) == _abi.RXC_ERR_NO_DEVICE
    assert not h.value


def test_product_package_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "rusterix_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f == "build.py":  # builds the checker for the tests; building it is not using it
                continue
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle_ffi" not in text and "librxoracle" not in text and "rx_oracle" not in text, f


def test_rust_sys_crate_is_generated_from_the_checked_mirror():
    """rust/rusterix-cuda-sys/src/lib.rs is what tools/gen_rust_sys.py prints from _abi.py (which the test above checks
    against the header), names every entry point of the header, and mirrors every struct field in order."""
    import re
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    want = subprocess.run([sys.executable, os.path.join(root, "tools", "gen_rust_sys.py")], capture_output=True, text=True, check=True).stdout
    have = open(os.path.join(root, "rust", "rusterix-cuda-sys", "src", "lib.rs")).read()
    assert have == want, "stale: run `python tools/gen_rust_sys.py --write`"
    header = open(os.path.join(root, "include", "rxcuda.h")).read()
    declared = set(re.findall(r"^\w[\w\s\*]*?\b(rxc_\w+)\s*\(", header, flags=re.M))
    bound = set(re.findall(r"pub fn (rxc_\w+)\(", have))
    assert declared and declared == bound, declared ^ bound
    for s in (_abi.rxc_frame, _abi.rxc_scene, _abi.rxc_batch3d, _abi.rxc_light, _abi.rxc_stats):
        body = re.search(r"pub struct %s \{(.*?)\n\}" % s.__name__, have, flags=re.S).group(1)
        fields = re.findall(r"pub (\w+):", body)
        assert fields == [("pass" if f == "pass_" else f) for f, _ in s._fields_]


def test_rust_wrapper_sources_agree_with_the_sys_crate_and_the_header():
    """rust/rusterix-cuda/src/{cuda,lower}.rs cannot be compiled here (no rustc); tools/check_rust_wrapper.py checks what
    can be: call arities, struct-literal field sets, constants, and the NodeOp opcode table against vm.OPS and the
    reference's enum order."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "check_rust_wrapper.py")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "MISMATCH" not in r.stdout
