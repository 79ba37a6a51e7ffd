"""Batch-shader JIT (rusterix_b200/csrc/rx_jit.cu): the translator and the NVRTC compilation, no GPU needed.
The GPU half (JIT kernel == interpreter kernel, bit for bit) is in test_parity_gpu.py."""
import ctypes as C
import os

import numpy as np
import pytest

import vm_programs
from rusterix_b200 import _abi, _lib, scenes
from rusterix_b200 import vm as rvm

NO_JIT = 0xFFFFFFFF


def _table(programs):
    flat = [p.flatten() for p in programs]
    arr = (_abi.rxc_program * max(1, len(flat)))()
    for i, fp in enumerate(flat):
        arr[i].code = fp.words.ctypes.data if len(fp.words) else None
        arr[i].n_words = len(fp.words)
        arr[i].entry, arr[i].shade_locals, arr[i].n_globals = fp.entry, fp.shade_locals, fp.n_globals
        arr[i].sets_opacity = 1 if fp.sets_opacity else 0
    return arr, flat


def _translate(programs):
    lib = _lib.load()
    arr, flat = _table(programs)
    idx = (C.c_uint32 * max(1, len(flat)))()
    n = lib.rxc_vm_translate(arr, len(flat), None, 0, idx)
    assert n > 0
    buf = C.create_string_buffer(int(n) + 1)
    assert lib.rxc_vm_translate(arr, len(flat), buf, len(buf), idx) == n
    return buf.value.decode(), list(idx)[:len(flat)]


def test_every_test_program_is_translated():
    progs = vm_programs.all_programs()
    src, idx = _translate(list(progs.values()))
    assert idx == list(range(len(progs)))
    for i in range(len(progs)):
        assert "vmj_p%d(" % i in src
    assert "vm_run_jit" in src and "vm_jit_may_bail" in src
    # straight-line code: no opcode fetch, every op is a compile-time constant
    assert "vm.code" not in src


def test_scene_shaders_are_translated():
    shaders = [scenes.shader_wood(), scenes.shader_marble(), scenes.shader_wood_ring(), scenes.shader_holes(),
               scenes.shader_control_flow(), scenes.shader_control_flow(True), scenes.shader_glass_tint(), scenes.shader_2d_scanlines()]
    _src, idx = _translate(shaders)
    assert idx == list(range(len(shaders)))


def test_programs_that_can_fault_stay_with_the_interpreter():
    """What the reference would panic on (or what has no static stack shape) is not translated: the interpreter keeps
    reporting it as a fault."""
    deep = [("Push", (1.0, 1.0, 1.0))] * 40 + [("Add",)] * 39 + [("SetColor",)]                  # value stack deeper than 32
    underflow = [("Add",), ("SetColor",)]                                                          # pops an empty stack
    uneven = [("UV",), ("If", [("Push", (1.0, 1.0, 1.0))], None), ("SetColor",)]                   # stack height differs at the merge
    good = [("UV",), ("SetColor",)]
    progs = [rvm.Program([deep], 0, 0, 0), rvm.Program([underflow], 0, 0, 0), rvm.Program([uneven], 0, 0, 0), rvm.Program([good], 0, 0, 0)]
    src, idx = _translate(progs)
    assert idx == [NO_JIT, NO_JIT, NO_JIT, 3]
    assert "vmj_p0(" not in src and "vmj_p3(" in src


def test_bare_return_is_only_translated_where_the_stack_below_is_empty():
    """A callee's `Return` with none of its own values on the stack pops a value of its CALLER (the value stack is shared,
    execution.rs:224-234): translated when every call site leaves nothing below it, left to the interpreter otherwise."""
    bare = [("Return",)]
    on_empty = rvm.Program([[("FunctionCall", 0, 0, 1), ("SetColor",)], bare], 0, 0, 0)
    on_value = rvm.Program([[("UV",), ("FunctionCall", 0, 0, 1), ("Add",), ("SetColor",)], bare], 0, 0, 0)
    _src, idx = _translate([on_empty, on_value])
    assert idx == [0, NO_JIT]


def test_palette_lookup_can_hand_over_to_the_interpreter():
    """PaletteIndex pushes nothing for a missing colour (execution.rs:735-742): the generated code returns 2 there."""
    progs = vm_programs.all_programs()
    name = next((n for n, p in progs.items() if any(op[0] == "PaletteIndex" for body in p.user_functions for op in _walk(body))), None)
    if name is None:
        pytest.skip("no test program uses PaletteIndex")
    src, idx = _translate([progs[name]])
    assert idx == [0] and "return 2;" in src and "case 0u: return true;" in src


def _walk(body):
    for op in body:
        yield op
        if op[0] in ("If", "For"):
            for arg in op[1:]:
                if isinstance(arg, list):
                    yield from _walk(arg)


def test_nvrtc_compiles_the_generated_kernel(tmp_path, monkeypatch):
    """The diagnostics kernel (k_vm_execute) with every test program as straight-line code, compiled for sm_100a here."""
    if not any(os.path.exists(os.path.join(d, "libnvrtc.so.12")) for d in ("/usr/local/cuda/lib64", "/usr/lib/x86_64-linux-gnu")):
        pytest.skip("libnvrtc.so.12 not installed")
    monkeypatch.setenv("RXC_JIT_CACHE", str(tmp_path))
    lib = _lib.load()
    arr, flat = _table(list(vm_programs.all_programs().values()))
    log = C.create_string_buffer(1 << 16)
    n = lib.rxc_vm_jit_compile(arr, len(flat), -1, 0, log, len(log))
    assert n > 0, log.value.decode()
    cached = [f for f in os.listdir(tmp_path) if f.endswith(".cubin")]
    assert len(cached) == 1
    # the second request is served from the disk cache
    assert lib.rxc_vm_jit_compile(arr, len(flat), -1, 0, log, len(log)) == n
    # an empty table has nothing to compile
    assert lib.rxc_vm_jit_compile(None, 0, -1, 0, log, len(log)) == 0
    # the reference-order kernel with the programs compiled (carried globals / locals: vmj_pN_ps)
    n2 = lib.rxc_vm_jit_compile(arr, len(flat), -2, 0, log, len(log))
    assert n2 > 0, log.value.decode()
    # the raster kernel itself (Nearest, pixels only, batch-shader mode) from the embedded sources: what a GPU box compiles at its first frame
    n3 = lib.rxc_vm_jit_compile(arr, len(flat), 0, 0, log, len(log))
    assert n3 > 0, log.value.decode()


# ---- which programs can observe the reference's per-tile Execution (rxc_vm_state_report) ----
def _state_report(programs, usage=None, scene_has_3d=True):
    lib = _lib.load()
    arr, flat = _table(programs)
    out = (C.c_uint32 * max(1, len(flat)))()
    use = None if usage is None else (C.c_uint8 * len(flat))(*usage)
    assert lib.rxc_vm_state_report(arr, len(flat), use, 1 if scene_has_3d else 0, out) == 0
    return list(out)[:len(flat)]


def test_state_report_passes_the_self_contained_shaders():
    """The example shaders assign everything they read: on them a fresh Execution per fragment and the reference's
    never-reset per-tile Execution compute the same pixels."""
    shaders = [scenes.shader_wood(), scenes.shader_marble(), scenes.shader_wood_ring(), scenes.shader_holes(),
               scenes.shader_control_flow(), scenes.shader_glass_tint()]
    assert _state_report(shaders, usage=[1] * len(shaders)) == [0] * len(shaders)


def test_state_report_flags_what_leaks_between_fragments():
    emissive = scenes.shader_control_flow(True)                                        # writes `emissive` on one path: every later fragment of the tile is lit by it
    stale_global = rvm.Program([[("LoadGlobal", 0), ("SetColor",), ("UV",), ("StoreGlobal", 0)]], 0, 0, 1)   # reads the global the previous fragment stored
    fresh_global = rvm.Program([[("UV",), ("StoreGlobal", 0), ("LoadGlobal", 0), ("SetColor",)]], 0, 0, 1)
    stale_local = rvm.Program([[("UV",), ("If", [("UV",), ("StoreLocal", 0)], None), ("LoadLocal", 0), ("SetColor",)]], 0, 1, 0)   # written on one path only
    writes_uvz = rvm.Program([[("Push", (0.0, 0.0, 7.0)), ("SetUV",)]], 0, 0, 0)       # leaves uv.z = 7 behind ...
    reads_uv = rvm.Program([[("UV",), ("SetColor",)]], 0, 0, 0)                        # ... where this one can see it
    assert _state_report([emissive]) == [1]
    assert _state_report([stale_global, fresh_global, stale_local]) == [1, 0, 1]
    assert _state_report([reads_uv]) == [0]
    assert _state_report([writes_uvz, reads_uv]) == [0, 1]
    # a 2D program reading the hit point sees the z the 3D pass left there -- unless the scene has no 3D batches
    reads_hit = rvm.Program([[("Hitpoint",), ("SetColor",)]], 0, 0, 0)
    assert _state_report([reads_hit], usage=[2], scene_has_3d=True) == [1]
    assert _state_report([reads_hit], usage=[2], scene_has_3d=False) == [0]
    assert _state_report([reads_hit], usage=[1], scene_has_3d=True) == [0]
    reads_hit_x = rvm.Program([[("Hitpoint",), ("GetComponents", [0]), ("SetColor",)]], 0, 0, 0)   # hitpoint.x is assigned in 2D too
    assert _state_report([reads_hit_x], usage=[2], scene_has_3d=True) == [0]
    # what the verifier declines is not analysable
    underflow = rvm.Program([[("Add",), ("SetColor",)]], 0, 0, 0)
    assert _state_report([underflow]) == [2]
