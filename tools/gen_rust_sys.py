#!/usr/bin/env python
"""Generates rust/rusterix-cuda-sys/src/lib.rs, the Rust mirror of include/rxcuda.h, from the ctypes mirror
(rusterix_b200/_abi.py) that tests/test_abi.py checks against the header through a C probe: same structs, same field
order and widths, same entry points.  tests/test_abi.py also checks that the committed file is what this script prints.
usage: gen_rust_sys.py [--write]"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rusterix_b200 import _abi  # noqa: E402

SCALARS = {C.c_uint8: "u8", C.c_int32: "i32", C.c_uint32: "u32", C.c_uint64: "u64", C.c_int64: "i64", C.c_float: "f32", C.c_double: "f64"}
STRUCTS = [v for v in vars(_abi).values() if isinstance(v, type) and issubclass(v, C.Structure) and v is not C.Structure]
# what the void pointers of the ABI point at (the header's declared pointee types)
VOID_FIELDS = {
    ("rxc_texture", "data"): "*const u8", ("rxc_pattern", "data"): "*const f32", ("rxc_program", "code"): "*const u32",
    ("rxc_batch3d", "vertices"): "*const f32", ("rxc_batch3d", "uvs"): "*const f32", ("rxc_batch3d", "normals"): "*const f32",
    ("rxc_batch3d", "indices"): "*const c_void", ("rxc_projected3d", "projected_vertices"): "*const f32", ("rxc_projected3d", "clipped_uvs"): "*const f32",
    ("rxc_projected3d", "clipped_normals"): "*const f32", ("rxc_projected3d", "clipped_indices"): "*const c_void", ("rxc_projected3d", "edges"): "*const f32",
    ("rxc_projected3d", "visible"): "*const u8", ("rxc_batch2d", "vertices"): "*const f32", ("rxc_batch2d", "uvs"): "*const f32",
    ("rxc_batch2d", "indices"): "*const c_void", ("rxc_scene", "palette"): "*const f32",
}
ARG_NAMES = {
    "rxc_create": ["device", "out"], "rxc_destroy": ["ctx"], "rxc_last_error": ["ctx"], "rxc_set_stream": ["ctx", "cuda_stream"],
    "rxc_set_assets": ["ctx", "tiles", "n_tiles"], "rxc_set_scene": ["ctx", "scene"], "rxc_set_lights": ["ctx", "lights", "n_lights"],
    "rxc_set_mapmini": ["ctx", "mapmini"], "rxc_rasterize": ["ctx", "frame", "pixels", "owner", "depth"],
    "rxc_rasterize_async": ["ctx", "frame", "pixels", "owner", "depth"],
    "rxc_rasterize_projected": ["ctx", "frame", "batches", "n_batches", "pixels", "owner", "depth"],
    "rxc_rasterize_batch": ["ctx", "frames", "n_frames", "pixels", "frame_stride_bytes"],
    "rxc_rasterize_batch_async": ["ctx", "frames", "n_frames", "pixels", "frame_stride_bytes"], "rxc_synchronize": ["ctx"],
    "rxc_owner_base": ["ctx", "batch", "base"], "rxc_selftest_div": ["ctx", "seed", "n_pairs", "mismatches", "bad_pair"],
    "rxc_vm_execute": ["ctx", "program", "n", "input", "output", "faults"], "rxc_set_profiling": ["ctx", "enabled"],
    "rxc_mgpu_unique_id": ["id"], "rxc_mgpu_init": ["ctx", "id", "rank", "world"], "rxc_mgpu_shutdown": ["ctx"],
    "rxc_mgpu_target": ["ctx", "bytes", "rank0_ptr", "mode"],
    "rxc_mgpu_rasterize": ["ctx", "frames", "n_frames", "offset_bytes", "frame_stride_bytes", "pitch_bytes"],
    "rxc_mgpu_deliver": ["ctx", "regions", "n_regions"], "rxc_mgpu_release": ["ctx"], "rxc_mgpu_status": ["ctx", "mode", "deliveries", "timeouts"],
    "rxc_vm_translate": ["programs", "n_programs", "source", "cap", "jit_index"],
    "rxc_vm_jit_compile": ["programs", "n_programs", "sample_mode", "planes", "log", "log_cap"],
    "rxc_set_vm_jit": ["ctx", "mode"], "rxc_update_scene": ["ctx", "scene", "keep_batches3d"],
    "rxc_set_vm_state_mode": ["ctx", "mode"], "rxc_get_vm_state_mode": ["ctx", "mode", "ordered_frames"],
    "rxc_vm_state_report": ["programs", "n_programs", "usage", "scene_has_3d", "report"],
    "rxc_vm_scene_state_report": ["ctx", "report", "cap", "n_programs"],
    "rxc_vm_jit_info": ["ctx", "n_translated", "kernels_compiled", "pending", "jit_launches", "log", "log_cap"],
    "rxc_get_stats": ["ctx", "out"], "rxc_pin_host": ["ctx", "ptr", "bytes"], "rxc_unpin_host": ["ctx", "ptr"], "rxc_reset_stats": ["ctx"], "rxc_kernel_name": ["kernel_class"],
}
# pointer arguments that are not the context handle: their Rust types
ARG_TYPES = {
    ("rxc_create", 1): "*mut *mut rxc_ctx", ("rxc_set_stream", 1): "*mut c_void", ("rxc_rasterize", 2): "*mut u8",
    ("rxc_rasterize", 3): "*mut u32", ("rxc_rasterize", 4): "*mut f32", ("rxc_rasterize_projected", 4): "*mut u8",
    ("rxc_rasterize_projected", 5): "*mut u32", ("rxc_rasterize_projected", 6): "*mut f32", ("rxc_rasterize_async", 2): "*mut u8",
    ("rxc_rasterize_async", 3): "*mut u32", ("rxc_rasterize_async", 4): "*mut f32", ("rxc_rasterize_batch", 3): "*mut u8",
    ("rxc_rasterize_batch_async", 3): "*mut u8", ("rxc_vm_translate", 2): "*mut c_char", ("rxc_vm_jit_info", 5): "*mut c_char", ("rxc_vm_state_report", 2): "*const u8", ("rxc_vm_jit_compile", 4): "*mut c_char", ("rxc_mgpu_init", 1): "*const u8", ("rxc_mgpu_target", 2): "*mut *mut c_void", ("rxc_pin_host", 1): "*mut c_void", ("rxc_unpin_host", 1): "*mut c_void", ("rxc_vm_execute", 3): "*const f32", ("rxc_vm_execute", 4): "*mut f32",
}


def rust_type(t, owner=None, field=None):
    if t in SCALARS:
        return SCALARS[t]
    if t is C.c_void_p:
        return VOID_FIELDS.get((owner, field), "*const c_void")
    if isinstance(t, type) and issubclass(t, C.Array):
        return "[%s; %d]" % (rust_type(t._type_, owner, field), t._length_)
    if isinstance(t, type) and issubclass(t, C._Pointer):
        return "*const " + rust_type(t._type_, owner, field)
    if isinstance(t, type) and issubclass(t, C.Structure):
        return t.__name__
    raise TypeError(t)


def header_enums():
    """(name, value) of every enumerator of the anonymous enums of include/rxcuda.h (RXC_MODE_*, RXC_SRC_*, RXVM_* ...)."""
    import re
    text = open(os.path.join(ROOT, "include", "rxcuda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    out = []
    for body in re.findall(r"\benum\s*\{(.*?)\}\s*;", text, flags=re.S):
        nxt = 0
        for item in body.split(","):
            item = item.strip()
            if not item:
                continue
            if "=" in item:
                name, val = [x.strip() for x in item.split("=")]
                nxt = int(val.rstrip("u"), 0)
            else:
                name = item
            out.append((name, nxt))
            nxt += 1
    return out


def generate():
    out = ["// GENERATED by tools/gen_rust_sys.py from rusterix_b200/_abi.py (the ctypes mirror that tests/test_abi.py checks",
           "// against include/rxcuda.h).  Do not edit: regenerate with `python tools/gen_rust_sys.py --write`.",
           "//", "// Raw bindings of the CUDA rasterizer's C ABI, version %d.  See INTEGRATION.md for the safe wrapper and the" % _abi.RXC_ABI_VERSION,
           "// `Rasterizer::rasterize_cuda` method a Rusterix maintainer adds on top.",
           "#![allow(non_camel_case_types)]", "use std::os::raw::{c_char, c_void};", "",
           "pub const RXC_ABI_VERSION: u32 = %d;" % _abi.RXC_ABI_VERSION, "pub const RXC_N_KERNELS: usize = %d;" % _abi.RXC_N_KERNELS]
    for code, name in sorted(_abi.STATUS_NAMES.items(), reverse=True):
        out.append("pub const %s: i32 = %d;" % (name, code))
    out.append("pub const RXC_MGPU_ID_BYTES: usize = %d;" % _abi.RXC_MGPU_ID_BYTES)
    out.append("// enumerators of include/rxcuda.h")
    for name, val in header_enums():
        out.append("pub const %s: u32 = %d;" % (name, val))
    out += ["", "#[repr(C)]", "pub struct rxc_ctx {", "    _private: [u8; 0],", "}"]
    for s in STRUCTS:
        out += ["", "#[repr(C)]", "#[derive(Clone, Copy)]", "pub struct %s {" % s.__name__]
        for fname, ftype in s._fields_:
            rname = "pass" if fname == "pass_" else fname   # `pass` is a Python keyword, not a Rust one
            out.append("    pub %s: %s," % (rname, rust_type(ftype, s.__name__, fname)))
        out.append("}")
    out += ["", 'extern "C" {']
    for name, restype, argtypes in _abi.EXPORTS:
        names = ARG_NAMES.get(name, [])
        args = []
        for i, t in enumerate(argtypes):
            if (name, i) in ARG_TYPES:
                rt = ARG_TYPES[(name, i)]
            elif t is C.c_void_p and i == 0 and names and names[0] == "ctx":
                rt = "*const rxc_ctx" if name in ("rxc_last_error", "rxc_owner_base") else "*mut rxc_ctx"
            elif isinstance(t, type) and issubclass(t, C._Pointer) and t._type_ in SCALARS:
                rt = "*mut " + SCALARS[t._type_]
            elif isinstance(t, type) and issubclass(t, C._Pointer) and t._type_ is _abi.rxc_stats:
                rt = "*mut rxc_stats"
            else:
                rt = rust_type(t)
            args.append("%s: %s" % (names[i] if i < len(names) else "a%d" % i, rt))
        ret = "" if restype is None else " -> " + ("*const c_char" if restype is C.c_char_p else rust_type(restype))
        out.append("    pub fn %s(%s)%s;" % (name, ", ".join(args), ret))
    out += ["}", ""]
    return "\n".join(out)


if __name__ == "__main__":
    text = generate()
    if "--write" in sys.argv:
        path = os.path.join(ROOT, "rust", "rusterix-cuda-sys", "src", "lib.rs")
        os.makedirs(os.path.dirname(path), exist_ok=True)
        open(path, "w").write(text)
        print("wrote", path)
    else:
        sys.stdout.write(text)
