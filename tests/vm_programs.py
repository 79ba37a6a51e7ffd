"""Programs that together execute every NodeOp the device VM accepts (reference rusteria/src/node/nodeop.rs:12-103),
for the interpreter parity tests: Python tree walk vs Python flat walk vs the C++ oracle vs the device."""
import numpy as np

from rusterix_b200 import scenes, vm
from rusterix_b200.vm import X, Body, Program


def _acc(b: Body, exprs):
    """color = sum of the expressions (each contributes all three lanes)."""
    total = b.let(0.0)
    for e in exprs:
        b.assign(total, total + e)
    return total


def prog_arith():
    b = Body()
    a, c = vm.uv * 3.0 - 1.0, vm.hitpoint * 0.5 + vm.vec3(0.25, -0.5, 2.0)
    t = _acc(b, [a + c, a - c, a * c, a / (c + 3.0), -a, vm.abs_(a), vm.floor(a * 3.0), vm.ceil(a * 3.0), vm.round_(a * 3.0), vm.fract(a * 3.0),
                 vm.mod(a * 5.0, vm.vec3(1.5, 2.0, 0.75)), vm.min_(a, c), vm.max_(a, c), vm.mix(a, c, vm.vec3(0.25, 0.5, 0.75)),
                 vm.smoothstep(0.0, 1.0, a), vm.step(a, c), vm.clamp(a * 2.0, -0.5, 0.5), vm.sqrt(vm.abs_(a)), vm.radians(a * 90.0), vm.degrees(a)])
    b.set("Color", t)
    b.set("Emissive", vm.vec3(vm.length(a), X(a.ops + [("Length2",)]).x, X(a.ops + [("Length3",)]).x))
    b.set("Bump", vm.vec3(vm.dot(a, c).x, X(a.ops + c.ops + [("Dot2",)]).x, X(a.ops + c.ops + [("Dot3",)]).x) + vm.cross(a, c) + vm.normalize(c))
    return Program([b.code], 0, b.n_locals, 0)


def prog_libm():
    b = Body()
    a = vm.uv * 6.0 - 3.0
    t = _acc(b, [vm.sin(a), vm.cos(a), vm.tan(a * 0.4), vm.atan(a), vm.atan2(a, vm.hitpoint + 0.1), vm.pow_(vm.abs_(a) + 0.1, vm.vec3(0.5, 2.0, 3.5)),
                 vm.log(vm.abs_(a) + 0.01), vm.rotate2d(a, 33.0), X(a.ops + [("Sin1",)]), X(a.ops + [("Sin2",)]), X(a.ops + [("Cos1",)]),
                 X(a.ops + [("Cos2",)])])
    b.set("Color", t)
    return Program([b.code], 0, b.n_locals, 0)


def prog_logic_stack():
    b = Body()
    a, c = vm.uv * 2.0 - 1.0, vm.hitpoint
    t = _acc(b, [a < c, a <= c, a > c, a >= c, a.eq(c), a.ne(c), (a < c).and_(a > -0.5), (a < c).or_(a > 0.5), (a < 0.0).not_(),
                 a.swz([2, 0]), a.swz([1, 1, 0]), a.swz([0, 1, 2, 0]), a.swz([5]), a.swz([1, 7, 2])])
    # Swap, Dup, Clear, Pack2/3, SetComponents with every arity, Print
    b.code += a.ops + c.ops + [("Swap",), ("Sub",)] + [("StoreLocal", b.n_locals)]
    sw = X([("LoadLocal", b.n_locals)]); b.n_locals += 1
    b.code += a.ops + [("Dup",), ("Mul",)] + [("StoreLocal", b.n_locals)]
    sq = X([("LoadLocal", b.n_locals)]); b.n_locals += 1
    b.code += c.ops + a.ops + [("Clear",)] + [("StoreLocal", b.n_locals)]
    kept = X([("LoadLocal", b.n_locals)]); b.n_locals += 1
    b.code += a.ops + [("Print",)]
    s1 = X(vm.vec3(0.0, 0.0, 0.0).ops + a.ops + [("SetComponents", [2])])
    s2 = X(vm.vec3(9.0, 9.0, 9.0).ops + a.ops + [("SetComponents", [1, 0])])
    s3 = X(vm.vec3(9.0, 9.0, 9.0).ops + a.ops + [("SetComponents", [2, 1, 0])])
    s4 = X(vm.vec3(7.0, 7.0, 7.0).ops + a.ops + [("SetComponents", [0, 1, 2, 0])])   # arity 4: nothing is written
    s5 = X(vm.vec3(7.0, 7.0, 7.0).ops + a.ops + [("SetComponents", [4, 1])])         # index 4 is ignored
    b.set("Color", t + sw + sq + kept)
    b.set("Emissive", s1 + s2 + s3 + s4 + s5)
    return Program([b.code], 0, b.n_locals, 0)


def prog_state():
    """Every Execution register read and written, pattern and palette lookups (hits, misses, wrap-around)."""
    b = Body()
    b.set("UV", vm.uv * 2.0)
    b.set("Normal", vm.normal + vm.vec3(0.1, 0.2, 0.3))
    b.set("Roughness", vm.roughness * 0.5 + vm.time_.x * 0.01)
    b.set("Metallic", vm.metallic + 0.25)
    b.set("Opacity", vm.opacity * 0.5 + 0.25)
    b.set("Bump", vm.bump + vm.sample_normal(vm.uv, "perlin") + vm.sample_normal(vm.uv, 5))          # 5: no such normal pattern
    b.set("Emissive", vm.emissive + vm.sample(vm.uv * 3.7 - 5.0, "bricks") + vm.sample(vm.uv, 11) + vm.sample(vm.uv, -3.0))
    tint = b.let((0.5, 0.5, 0.5))
    b.code += vm.X.of(9.0).ops + [("PaletteIndex",)]       # colour 9 does not exist: nothing is pushed
    b.code += vm.X.of(1.0).ops + [("PaletteIndex",)]       # colour 1 is None: nothing is pushed
    b.assign(tint, tint * vm.palette(3.0) + vm.palette(0.4))
    b.set("Color", vm.color * tint + vm.hitpoint * 0.1)
    return Program([b.code], 0, b.n_locals, 0)


def prog_return_paths():
    """Return with and without a value on the stack, a callee that falls off its end leaving a value / nothing,
    arguments popped into locals, and a Return inside a loop (device rule: leaves the function)."""
    f_value = Body(n_params=1)                 # fn 1: leaves x * 2 on the stack
    f_value.code += (f_value.param(0) * 2.0).ops
    f_empty = Body(n_params=1)                 # fn 2: nothing on the stack -> zero
    f_empty.code += f_empty.param(0).ops + [("Clear",)]
    f_ret = Body(n_params=2)                   # fn 3: early return from inside a loop
    i = f_ret.let(0.0)
    init, incr, body = f_ret.sub(), f_ret.sub(), f_ret.sub()
    incr.assign(i, i + 1.0)
    hit = body.sub()
    hit.ret(i + f_ret.param(1))
    body.if_(i.x * 0.25 > f_ret.param(0).x, hit)
    f_ret.for_(init, i < 16.0, incr, body)
    f_ret.close(body)
    f_ret.code += vm.X.of(-1.0).ops
    f_bare = Body(n_params=0)                  # fn 4: Return on an empty stack -> zero
    f_bare.code += [("Return",)]
    b = Body()
    b.set("Color", vm.call(1, 1, vm.uv) + vm.call(2, 1, vm.uv) + vm.call(3, 3, vm.uv.x * 3.0, vm.hitpoint))
    b.set("Metallic", vm.call(4, 0) + 0.125)   # called on an empty stack: a bare Return yields zero (it would pop the caller's operand otherwise)
    b.set("Emissive", vm.call(3, 3, vm.vec3(100.0, 0.0, 0.0), 5.0))    # the loop runs out: -1
    early = b.sub()
    early.ret(0.0)
    b.if_(vm.uv.y > 0.5, early)                # Return from shade() itself
    b.set("Bump", (1.0, 2.0, 3.0))
    return Program([b.code, f_value.code, f_empty.code, f_ret.code, f_bare.code], 0, b.n_locals, 0)


def all_programs():
    return {
        "arith": prog_arith(), "libm": prog_libm(), "logic_stack": prog_logic_stack(), "state": prog_state(),
        "return_paths": prog_return_paths(), "wood": scenes.shader_wood(), "holes": scenes.shader_holes(),
        "control_flow": scenes.shader_control_flow(), "glass": scenes.shader_glass_tint(), "scanlines": scenes.shader_2d_scanlines(),
    }


PALETTE = [(0.1, 0.1, 0.1), None, (0.9, 0.7, 0.3), (0.2, 0.8, 0.4)]


def records(n, seed=7):
    """n x 18 floats: uv, color, normal, hitpoint, time, opacity."""
    r = scenes.rand01(n * 18, scenes.SEED + seed).reshape(n, 18).astype(np.float32)
    r[:, 0:3] = r[:, 0:3] * 4.0 - 1.5          # uv beyond [0, 1): patterns wrap
    r[:, 6:9] = r[:, 6:9] * 2.0 - 1.0
    r[:, 9:12] = r[:, 9:12] * 8.0 - 4.0
    r[:, 12:15] = r[:, 12:13] * 10.0
    return r


def state_from_record(rec, bank, bank_normal):
    st = vm.VMState()
    st.uv, st.color, st.normal = rec[0:3].copy(), rec[3:6].copy(), rec[6:9].copy()
    st.hitpoint, st.time, st.opacity = rec[9:12].copy(), rec[12:15].copy(), rec[15:18].copy()
    st.patterns, st.patterns_normal, st.palette = bank, bank_normal, PALETTE
    return st


def state_outputs(st):
    return np.concatenate([st.uv, st.color, st.normal, st.roughness, st.metallic, st.emissive, st.opacity, st.bump]).astype(np.float32)
