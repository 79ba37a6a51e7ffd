"""make_textures.rusteria (the script at the root of the reference that GENERATED the pattern textures embedded in the
rusteria crate: `let tex = alloc(512, 512); iterate(tex, "make_value_noise"); save(tex, "rusteria/embedded/value.png")`
...) lowered by hand the way the Rusteria compiler lowers it (rusteria/src/compile.rs: binary ops push left then right;
`a op= b` is load a, b, op, store; a one-argument vec2 pads with zero, parser.rs:880-896; integer literals are floats,
parser.rs:843-852; `return` is NodeOp::Return, an `if` is cond + NodeOp::If, a `for` is NodeOp::For(init, cond, incr,
body); optimize() is a no-op).  Together with the PNGs the reference committed, these are golden vectors for the VM's
control flow: nested user functions, Return inside If, For loops, Mod, Rotate2D, Dot / Dot2 / Sin2, swizzles.

`Iterate` (execution.rs:664-714) runs the named function per texel with uv = (x / w, y / h, 0) and stores its colour;
`Save` writes clamp(v, 0, 1) * 255 truncated (textures/mod.rs:85-99)."""
import numpy as np

from rusterix_b200 import vm
from rusterix_b200.vm import Body, Program


def _value_functions():
    # fn value_hash(p, scale)
    h = Body(2)
    p, scale = h.param(0), h.param(1)
    h.assign(p, vm.mod(p, scale))
    p3 = h.let(vm.fract(p.swz([0, 1, 0]) * 0.13))
    h.assign(p3, p3 + vm.dot(p3, p3.swz([1, 2, 0]) + 3.333))
    h.ret(vm.fract((p3.x + p3.y) * p3.z))
    HASH, HASH_LOCALS = 0, h.n_locals

    # fn value_noise(x, scale)
    n = Body(2)
    x, scale = n.param(0), n.param(1)
    n.assign(x, x * scale)
    i = n.let(vm.floor(x))
    f = n.let(vm.fract(x))
    a = n.let(vm.call(HASH, HASH_LOCALS, i, scale))
    b = n.let(vm.call(HASH, HASH_LOCALS, i + vm.vec2(1.0, 0.0), scale))
    c = n.let(vm.call(HASH, HASH_LOCALS, i + vm.vec2(0.0, 1.0), scale))
    d = n.let(vm.call(HASH, HASH_LOCALS, i + vm.vec2(1.0, 1.0), scale))
    u = n.let(f * f * (3.0 - 2.0 * f))
    n.ret(vm.mix(a, b, u.x) + (c - a) * u.y * (1.0 - u.x) + (d - b) * u.x * u.y)
    NOISE, NOISE_LOCALS = 1, n.n_locals

    # fn fbm_value(x, scale)
    m = Body(2)
    x, scale = m.param(0), m.param(1)
    v = m.let(0.0)
    a = m.let(0.5)
    shift = m.let(vm.vec2(100.0, 0.0))          # vec2(100)
    init = m.sub(); i = init.let(0.0); m.close(init)
    incr = m.sub(); incr.assign(i, i + 1.0)
    body = m.sub()
    body.assign(v, v + a * vm.call(NOISE, NOISE_LOCALS, x, scale))
    body.assign(x, vm.rotate2d(x, 0.5) * 2.0 + shift)
    body.assign(a, a * 0.5)
    m.for_(init, i < 5.0, incr, body)
    m.ret(v)
    return [h.code, n.code, m.code], (NOISE, NOISE_LOCALS), (2, m.n_locals)


def make_value_noise():
    fns, (NOISE, NL), _ = _value_functions()
    s = Body()
    s.set("Color", vm.call(NOISE, NL, vm.uv, 20.0))
    return Program(fns + [s.code], len(fns), s.n_locals, 0)


def make_fbm_value_noise():
    fns, _, (FBM, FL) = _value_functions()
    s = Body()
    s.set("Color", vm.call(FBM, FL, vm.uv, 20.0))
    return Program(fns + [s.code], len(fns), s.n_locals, 0)


def iterate_records(width, height, xs=None, ys=None):
    """TexStorage::par_iterate_with (textures/mod.rs:44-62): uv = (x * (1 / w), y * (1 / h), 0)."""
    if xs is None:
        ys, xs = np.mgrid[0:height, 0:width]
    xs, ys = np.asarray(xs).ravel(), np.asarray(ys).ravel()
    rec = np.zeros((xs.size, 18), np.float32)
    rec[:, 0] = xs.astype(np.float32) * (np.float32(1.0) / np.float32(width))
    rec[:, 1] = ys.astype(np.float32) * (np.float32(1.0) / np.float32(height))
    return rec


def save_pixels(colors):
    """TexStorage::save_png (textures/mod.rs:85-99): `(v.clamp(0, 1) * 255.0) as u8`."""
    c = np.nan_to_num(np.asarray(colors, np.float32), nan=0.0)
    return np.trunc(np.clip(c, 0.0, 1.0) * np.float32(255.0)).astype(np.uint8)


_sin2, _dot2 = vm._fn("Sin2", 1), vm._fn("Dot2", 2)


def _perlin_functions():
    # fn grad2(i)
    g = Body(1)
    i = g.param(0)
    h = g.let(vm.floor(vm.fract(_sin2(_dot2(i, vm.vec2(127.1, 311.7))) * 43758.5453) * 8.0))
    for k, (gx, gy) in enumerate([(1.0, 0.0), (-1.0, 0.0), (0.0, 1.0), (0.0, -1.0), (0.707, 0.707), (-0.707, 0.707), (0.707, -0.707)]):
        t = g.sub()
        t.ret(vm.vec3(*[(-vm.X.of(-c) if c < 0 else c) for c in (gx, gy)], 0.0))   # a negative literal is Push |c|, Neg
        g.if_(h < float(k + 1), t)
    g.ret(vm.vec3(-vm.X.of(0.707), -vm.X.of(0.707), 0.0))
    GRAD, GRAD_LOCALS = 0, g.n_locals

    # fn fade(t)
    fd = Body(1)
    t = fd.param(0)
    fd.ret(t * t * t * (t * (t * 6.0 - 15.0) + 10.0))
    FADE, FADE_LOCALS = 1, fd.n_locals

    # fn perlin_noise(p, period)
    n = Body(2)
    p, period = n.param(0), n.param(1)
    n.assign(p, p * period)
    i = n.let(vm.floor(p))
    f = n.let(vm.fract(p))
    iw = n.let(vm.mod(i, period))
    iw10 = n.let(vm.mod(i + vm.vec2(1.0, 0.0), period))
    iw01 = n.let(vm.mod(i + vm.vec2(0.0, 1.0), period))
    iw11 = n.let(vm.mod(i + vm.vec2(1.0, 1.0), period))
    g00, g10, g01, g11 = (n.let(vm.call(GRAD, GRAD_LOCALS, w)) for w in (iw, iw10, iw01, iw11))
    d00 = n.let(vm.vec2(f.x, f.y))
    d10 = n.let(vm.vec2(f.x - 1.0, f.y))
    d01 = n.let(vm.vec2(f.x, f.y - 1.0))
    d11 = n.let(vm.vec2(f.x - 1.0, f.y - 1.0))
    v00, v10, v01, v11 = (n.let(_dot2(d, gg)) for d, gg in ((d00, g00), (d10, g10), (d01, g01), (d11, g11)))
    u = n.let(vm.call(FADE, FADE_LOCALS, f))
    nx0 = n.let(vm.mix(v00, v10, u.x))
    nx1 = n.let(vm.mix(v01, v11, u.x))
    nxy = n.let(vm.mix(nx0, nx1, u.y))
    n.ret(0.5 * vm.vec3(nxy + 1.0, nxy + 1.0, nxy + 1.0))       # vec3(e) evaluates e three times (parser.rs:912-940)
    NOISE, NOISE_LOCALS = 2, n.n_locals

    # fn perlin_fbm(p, scale)
    m = Body(2)
    p, scale = m.param(0), m.param(1)
    v = m.let(0.0)
    a = m.let(0.5)
    init = m.sub(); i = init.let(0.0); m.close(init)
    incr = m.sub(); incr.assign(i, i + 1.0)
    body = m.sub()
    body.assign(v, v + a * vm.call(NOISE, NOISE_LOCALS, p, scale))
    body.assign(p, p * 2.0 + vm.vec2(100.0, 0.0))             # vec2(100.0)
    body.assign(a, a * 0.5)
    m.for_(init, i < 5.0, incr, body)
    m.ret(vm.vec3(v, v, v))
    return [g.code, fd.code, n.code, m.code], (NOISE, NOISE_LOCALS), (3, m.n_locals)


def make_perlin_noise():
    fns, (NOISE, NL), _ = _perlin_functions()
    s = Body()
    s.set("Color", vm.call(NOISE, NL, vm.uv, 10.0))
    return Program(fns + [s.code], len(fns), s.n_locals, 0)


def make_perlin_fbm():
    fns, _, (FBM, FL) = _perlin_functions()
    s = Body()
    s.set("Color", vm.call(FBM, FL, vm.uv, 10.0))
    return Program(fns + [s.code], len(fns), s.n_locals, 0)


def _hash21():
    # fn hash21(p)
    h = Body(1)
    p = h.param(0)
    p3 = h.let(vm.fract(p * vm.vec3(0.1031, 0.1030, 0.0973)))
    h.assign(p3, p3 + vm.dot(p3, p3.swz([1, 2, 0]) + 33.33))
    h.ret(vm.fract((p3.swz([0, 0]) + p3.swz([1, 2])) * p3.swz([2, 1])))
    return h


def _set_swizzled(b: Body, local, comps, rhs):
    """`t.swz = rhs` (compile.rs:560-567): rhs, load t, Swap, SetComponents, store t."""
    b.code += vm.X.of(rhs).ops + local.ops + [("Swap",), ("SetComponents", list(comps)), ("StoreLocal", local.ops[0][1])]


def _op_swizzled(b: Body, local, comps, op, rhs):
    """`t.swz op= rhs` (compile.rs:568-578): load t, Dup, GetComponents, rhs, op, SetComponents, store t."""
    b.code += local.ops + [("Dup",), ("GetComponents", list(comps))] + vm.X.of(rhs).ops + [(op,), ("SetComponents", list(comps)),
                                                                                           ("StoreLocal", local.ops[0][1])]


def _bricks_or_tiles(ratio_v, gap_v, bevel_v, rounder, use_mod_fn, stagger):
    h = _hash21()
    HASH, HASH_LOCALS = 0, h.n_locals
    s_ = Body()
    ratio = s_.let(ratio_v)
    cell = s_.let(18.0)
    gap = s_.let(gap_v)
    bevel = s_.let(bevel_v)
    cycle = s_.let(vm.vec2(rounder(cell / ratio), cell))
    u = s_.let(vm.uv)
    w = s_.let(vm.vec2(ratio, 1.0))
    s_.assign(u, u * (cell / w))
    if stagger:
        _op_swizzled(s_, u, [0], "Add", 0.5 * vm.X(vm.floor(u.y).ops + [vm.push(2.0), ("Mod",)]))   # u.x += 0.5 * (floor(u.y) % 2.0)
    p = s_.let(vm.mod(u, cycle))                      # `u % cycle` and mod(u, cycle) are both NodeOp::Mod
    s_.assign(p, vm.floor(p))
    _set_swizzled(s_, p, [2], 0.0)
    id_ = s_.let(vm.call(HASH, HASH_LOCALS, p))
    s = s_.let(w * (vm.fract(u) - vm.X.of(1.0) / 2.0))
    a = s_.let(w / 2.0 - gap - vm.abs_(s))
    b = s_.let(a * 2.0 / bevel)
    m = s_.let(vm.min_(b.x, b.y))
    mask = s_.let(vm.clamp(m, 0.0, 1.0))
    s_.set("Color", vm.vec2(mask.x, id_.x))
    return Program([h.code, s_.code], 1, s_.n_locals, 0)


def make_bricks():
    return _bricks_or_tiles(3.0, 0.05, 0.0, vm.round_, False, True)


def make_tiles():
    return _bricks_or_tiles(1.0, 0.01, 0.1, vm.ceil, True, False)


def make_blocks():
    h = _hash21()
    HASH, HASH_LOCALS = 0, h.n_locals
    # fn s_box(p, b, rf)
    sb = Body(3)
    p_, b_, rf = sb.param(0), sb.param(1), sb.param(2)
    q = sb.let(vm.abs_(p_) - b_ + rf)
    sb.ret(vm.length(vm.max_(q, 0.0)) + vm.min_(vm.max_(q.x, q.y), 0.0) - rf)
    SBOX, SBOX_LOCALS = 1, sb.n_locals

    s = Body()
    gap = s.let(0.1)
    rotation = s.let(2.0)
    rounding = s.let(0.04)
    p = s.let(vm.uv)
    ip = s.let(vm.floor(vm.uv))
    s.assign(p, p - ip)
    last_l = s.let(0.0)
    l = s.let(vm.vec2(1.0, 1.0))
    r = s.let(vm.call(HASH, HASH_LOCALS, ip))

    def swap_xy(blk):
        blk.assign(p, vm.vec2(p.y, p.x))
        blk.assign(l, vm.vec2(l.y, l.x))

    init = s.sub(); i = init.let(0.0); s.close(init)
    incr = s.sub(); incr.assign(i, i + 1.0)
    body = s.sub()
    body.assign(r, vm.fract(vm.dot(l + r, vm.vec2(123.71, 439.43))) * 0.4 + (vm.X.of(1.0) - 0.4) / 2.0)
    body.assign(last_l, l)
    t1 = body.sub(); swap_xy(t1); body.if_(l.x > l.y, t1)
    t2 = body.sub(); _op_swizzled(t2, l, [0], "Div", r); _op_swizzled(t2, p, [0], "Div", r)
    e2 = body.sub(); _op_swizzled(e2, l, [0], "Div", 1.0 - r); _set_swizzled(e2, p, [0], (p.x - r) / (1.0 - r))
    body.if_(p.x < r, t2, e2)
    t3 = body.sub(); swap_xy(t3); body.if_(last_l.x > last_l.y, t3)
    s.for_(init, i < 6.0, incr, body)
    s.assign(p, p - 0.5)
    id_ = s.let(vm.call(HASH, HASH_LOCALS, ip + l))
    s.assign(p, vm.rotate2d(p, (id_ - 0.5) * rotation))
    th = s.let(l * 0.02 * gap)
    c = s.let(vm.call(SBOX, SBOX_LOCALS, p, 0.5 - th, rounding))
    c01 = s.let(vm.clamp(0.5 - c / (vm.X.of(2.0) * 0.5), 0.0, 1.0))
    mask = s.let(1.0)
    t4 = s.sub(); t4.assign(mask, 0.0); s.if_(c > 0.0, t4)
    s.set("Color", vm.vec3(mask.x, id_.x, c01.x))
    return Program([h.code, sb.code, s.code], 2, s.n_locals, 0)
