"""CPU tests of the Rusteria VM row: the op-tree interpreter of the oracle against an independent
Python restatement of rusteria/src/node/execution.rs, the lowering to flat code (what the device runs) against the
tree, hand-derived known answers for the reference's quirks, and the limits the flattening enforces."""
import numpy as np
import pytest

import oracle_ffi
import vm_programs
from rusterix_b200 import scenes, types, vm
from rusterix_b200.vm import X, Body, Program

BANK = scenes.pattern_bank()
BANK_N = scenes.pattern_bank()[:3]


def _scene_with(programs):
    s = types.Scene()
    s.patterns, s.patterns_normal = BANK, BANK_N
    for p in programs:
        s.add_shader(p)
    a = types.Assets()
    a.palette = vm_programs.PALETTE
    return s, a


def _run_py(prog, rec, flat):
    st = vm_programs.state_from_record(rec, BANK, BANK_N)
    if flat:
        vm.run_flat(prog.flatten(), st)
    else:
        vm.run_tree(prog, st)
    return vm_programs.state_outputs(st)


@pytest.mark.parametrize("name", list(vm_programs.all_programs()))
def test_flat_code_matches_tree(name):
    """Lowering If / For / FunctionCall / Return to jumps does not change any result (bit for bit)."""
    prog = vm_programs.all_programs()[name]
    for rec in vm_programs.records(12):
        a, b = _run_py(prog, rec, False), _run_py(prog, rec, True)
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), name


@pytest.mark.parametrize("name", list(vm_programs.all_programs()))
def test_oracle_vm_matches_python_restatement(name):
    progs = vm_programs.all_programs()
    scene, assets = _scene_with(list(progs.values()))
    oracle_ffi.set_programs(scene, assets)
    recs = vm_programs.records(24)
    out, faults = oracle_ffi.vm_execute(list(progs).index(name), recs)
    assert faults == 0
    want = np.stack([_run_py(progs[name], r, False) for r in recs])
    # sin/cos/pow... come from numpy here and from glibc in the oracle: equal to a few ulp; everything else is exact
    exact = name in ("arith", "logic_stack", "return_paths", "holes", "scanlines") and False
    if exact:
        assert np.array_equal(out.view(np.uint32), want.view(np.uint32))
    np.testing.assert_allclose(out, want, rtol=3e-6, atol=3e-6)


def _exec(prog, rec=None):
    scene, assets = _scene_with([prog])
    oracle_ffi.set_programs(scene, assets)
    r = np.zeros((1, 18), np.float32) if rec is None else np.asarray(rec, np.float32).reshape(1, 18)
    out, faults = oracle_ffi.vm_execute(0, r)
    return out[0].reshape(8, 3), faults


UVc, COLOR, NORMAL, ROUGH, METAL, EMISSIVE, OPACITY, BUMP = range(8)


def test_kat_execution_new_defaults():
    """execution.rs:58-77: roughness starts at 0.5, everything else at zero; an empty shade() changes nothing."""
    out, faults = _exec(Program([[]], 0, 0, 0))
    assert faults == 0
    assert out[ROUGH].tolist() == [0.5, 0.5, 0.5] and not out[[UVc, COLOR, NORMAL, METAL, EMISSIVE, OPACITY, BUMP]].any()


def test_kat_cos1_cos2_call_sin():
    """execution.rs:342-349: Cos1 / Cos2 evaluate sin."""
    b = Body()
    b.set("Color", X(vm.X.of((0.5, 1.0, 2.0)).ops + [("Cos1",)]))
    b.set("Emissive", X(vm.X.of((0.5, 1.0, 2.0)).ops + [("Cos2",)]))
    out, _ = _exec(Program([b.code], 0, 0, 0))
    np.testing.assert_allclose(out[COLOR], [np.sin(np.float32(0.5)), 0, 0], rtol=1e-6)
    np.testing.assert_allclose(out[EMISSIVE], [np.sin(np.float32(0.5)), np.sin(np.float32(1.0)), 0], rtol=1e-6)


def test_kat_swizzles():
    """execution.rs:135-183: one component broadcasts, two leave z = 0, four or none give zero; SetComponents of
    arity 4 writes nothing and out-of-range indices are skipped."""
    v = (1.0, 2.0, 3.0)
    cases = {(1,): [2, 2, 2], (2, 0): [3, 1, 0], (0, 1, 2): [1, 2, 3], (0, 1, 2, 0): [0, 0, 0], (5,): [0, 0, 0], (1, 9, 2): [2, 3, 0]}
    for sw, want in cases.items():
        b = Body()
        b.set("Color", X(X.of(v).ops + [("GetComponents", list(sw))]))
        out, _ = _exec(Program([b.code], 0, 0, 0))
        assert out[COLOR].tolist() == want, sw
    sets = {(2,): [7, 7, 1], (1, 0): [2, 1, 7], (2, 1, 0): [3, 2, 1], (0, 1, 2, 0): [7, 7, 7], (4, 1): [7, 2, 7]}
    for sw, want in sets.items():
        b = Body()
        b.set("Color", X(X.of((7.0, 7.0, 7.0)).ops + X.of(v).ops + [("SetComponents", list(sw))]))
        out, _ = _exec(Program([b.code], 0, 0, 0))
        assert out[COLOR].tolist() == want, sw


def test_kat_palette_miss_pushes_nothing():
    """execution.rs:735-742: a missing or None palette entry leaves the stack as it was."""
    b = Body()
    b.code += X.of((0.25, 0.5, 0.75)).ops + X.of(1.0).ops + [("PaletteIndex",)] + X.of(99.0).ops + [("PaletteIndex",)] + [("SetColor",)]
    b.set("Emissive", vm.palette(2.0))
    out, faults = _exec(Program([b.code], 0, 0, 0))
    assert faults == 0 and out[COLOR].tolist() == [0.25, 0.5, 0.75]
    np.testing.assert_array_equal(out[EMISSIVE], np.float32([0.9, 0.7, 0.3]))


def test_kat_loop_and_call():
    """sum_{i<5} f(i) with f(x) = x*x + 1 through FunctionCall, For and a global: 0+1+4+9+16 + 5 = 35."""
    f = Body(n_params=1)
    f.ret(f.param(0) * f.param(0) + 1.0)
    b = Body()
    i = b.let(0.0)
    init, incr, body = b.sub(), b.sub(), b.sub()
    incr.assign(i, i + 1.0)
    body.set_global(0, X([("LoadGlobal", 0)]) + vm.call(1, 1, i))
    b.for_(init, i < 5.0, incr, body)
    b.set("Color", X([("LoadGlobal", 0)]))
    prog = Program([b.code, f.code], 0, b.n_locals, 1)
    out, faults = _exec(prog)
    assert faults == 0 and out[COLOR].tolist() == [35.0, 35.0, 35.0]
    st = vm.VMState()
    vm.run_flat(prog.flatten(), st)
    assert st.color.tolist() == [35.0, 35.0, 35.0]


def test_kat_pattern_sample_wraps():
    """textures/mod.rs:131-146: u - floor(u), floor(u * w) as i32, wrapped into the texture."""
    w, h, data = BANK[4]
    for uvv in [(0.0, 0.0), (0.999, 0.5), (-0.25, 1.75), (3.5, -2.125), (1.0, 1.0)]:
        b = Body()
        b.set("Color", vm.sample(X.of((uvv[0], uvv[1], 0.0)), "bricks"))
        out, _ = _exec(Program([b.code], 0, 0, 0))
        u = np.float32(uvv[0]) - np.floor(np.float32(uvv[0]))
        v = np.float32(uvv[1]) - np.floor(np.float32(uvv[1]))
        x, y = int(np.floor(u * np.float32(w))) % w, int(np.floor(v * np.float32(h))) % h
        np.testing.assert_array_equal(out[COLOR], data[y * w + x])


def test_kat_state_leak_of_the_reference_is_in_the_oracle():
    """src/rasterizer.rs:310: one Execution per screen tile, never reset -- what a program writes to `emissive` is
    still there for the next fragment of the tile, shader or not.  The oracle keeps that (the frame depends on
    tile_size); the device starts every fragment from Execution::new.  A scene whose programs assign
    everything they read is tile-size invariant again, which is what the parity scenes are."""
    leaky = Body()
    leaky.set("Emissive", (0.3, 0.0, 0.0))
    cfg = scenes.cube(64, 64, 16, logo_size=16)
    cfg.scene.add_shader(Program([leaky.code], 0, 0, 0))
    cfg.scene.d3_static[0].shader(0)
    plain = scenes.cube(64, 64, 16, logo_size=16).scene.d3_static[0]
    plain.transform_3d = np.array([[1, 0, 0, 0.2], [0, 1, 0, 0.1], [0, 0, 1, -1.5], [0, 0, 0, 1]], dtype=np.float32)
    cfg.scene.d3_static.append(plain)                       # drawn after the shaded box, partly behind it
    a = oracle_ffi.rasterize(cfg.rasterizer(), cfg.scene, cfg.assets, 64, 64, 8)[0]
    b = oracle_ffi.rasterize(cfg.rasterizer(), cfg.scene, cfg.assets, 64, 64, 64)[0]
    assert (a != b).any()


def test_flatten_rejects_what_the_device_cannot_run():
    with pytest.raises(ValueError):
        Program([[("Push", (1, 1, 1)), ("Push", (1, 1, 1)), ("Alloc",)]], 0, 0, 0).flatten()
    rec = Body()
    rec.code += vm.call(0, 0).ops
    with pytest.raises(ValueError):
        Program([rec.code], 0, 0, 0).flatten()              # recursion
    with pytest.raises(ValueError):
        Program([[("LoadLocal", 3)]], 0, 2, 0).flatten()    # the reference would panic on the index
    with pytest.raises(ValueError):
        Program([[("FunctionCall", 0, 0, 5)]], 0, 0, 0).flatten()
    assert Program([[]], None, 0, 0).flatten().words.size == 0   # shade_index None: nothing to run


def test_shader_supports_opacity_scans_only_the_top_level():
    """program.rs:44-55."""
    b = Body()
    inner = b.sub()
    inner.set("Opacity", 0.5)
    b.if_(vm.uv.x > 0.5, inner)
    assert not Program([b.code], 0, 0, 0).shader_supports_opacity()
    assert scenes.shader_holes().shader_supports_opacity()


def test_shader_index_past_the_list_runs_nothing():
    """scene.shaders.get(i) == None (src/rasterizer.rs:1281-1300): same frame as without a shader."""
    cfg = scenes.cube(48, 48, 16, logo_size=8)
    want = oracle_ffi.rasterize(cfg.rasterizer(), cfg.scene, cfg.assets, 48, 48, 16)[0]
    cfg.scene.d3_static[0].shader(3)
    got = oracle_ffi.rasterize(cfg.rasterizer(), cfg.scene, cfg.assets, 48, 48, 16)[0]
    assert np.array_equal(want, got)
