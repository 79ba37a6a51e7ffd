"""Host-side mirror of the Rusteria VM program model (reference rusteria/src/node/nodeop.rs:12-103,
program.rs:7-29) and the two serialisations this repo uses:

* `Program.flatten()`  -- the word stream of include/rxcuda.h (`rxc_program`): the NodeOp TREE of the
  shade function and of every function it can call, lowered to straight-line code with jumps.  This is
  what a Rust host would produce from `Program.user_functions` before calling `rxc_set_scene`.
* `Program.encode_tree()` -- the same tree kept as a tree (lengths instead of jumps) for the CPU oracle,
  whose interpreter recurses exactly like `Execution::execute` (execution.rs:109-768).

The Rusteria compiler (scanner/parser/compile.rs) stays on the host and is not restated: programs are
written here as NodeOp trees, the compiler's output format.

An op is a tuple: ("Add",), ("LoadLocal", 2), ("Push", (x, y, z)), ("GetComponents", [0, 1]),
("If", then_ops, else_ops_or_None), ("For", init, cond, incr, body), ("FunctionCall", arity, total_locals, index).
"""
import math
import struct
from typing import List, Optional, Sequence

import numpy as np

# opcode numbers: shared by the flat code (device), the tree code (oracle) and include/rxcuda.h (RXVM_*)
OPS = [
    "LoadGlobal", "StoreGlobal", "LoadLocal", "StoreLocal", "Swap", "GetComponents", "SetComponents", "If", "For",
    "Push", "FunctionCall", "Return", "Dup", "Clear", "Pack2", "Pack3", "Add", "Sub", "Mul", "Div", "Length", "Length2",
    "Length3", "Abs", "Sin", "Sin1", "Sin2", "Cos", "Cos1", "Cos2", "Tan", "Atan", "Atan2", "Rotate2D", "Dot", "Dot2",
    "Dot3", "Cross", "Normalize", "Floor", "Ceil", "Round", "Fract", "Mod", "Degrees", "Radians", "Min", "Max", "Mix",
    "Smoothstep", "Step", "Clamp", "Sqrt", "Pow", "Log", "Print", "Eq", "Ne", "Lt", "Le", "Gt", "Ge", "And", "Or", "Not",
    "Neg", "UV", "SetUV", "Normal", "SetNormal", "Hitpoint", "Time", "Sample", "SampleNormal", "Color", "SetColor",
    "Roughness", "SetRoughness", "Metallic", "SetMetallic", "Emissive", "SetEmissive", "Opacity", "SetOpacity", "Bump",
    "SetBump", "Alloc", "Iterate", "Save", "PaletteIndex",
    # flat code only (lowering of If / For / function bodies)
    "Jz", "Jmp", "Mark", "Trunc", "Unmark", "End",
]
OPCODE = {n: i for i, n in enumerate(OPS)}
HOST_ONLY = {"Alloc", "Iterate", "Save"}  # texture baking (execution.rs:643-733): not part of a shade() call

# device limits (rx_kernels.cu RXVM_*): programs that statically need more are rejected at flatten time
MAX_STACK = 32
MAX_LOCALS = 32       # per function frame
MAX_GLOBALS = 16
MAX_CALL_DEPTH = 8
MAX_LOOP_DEPTH = 8


def f32_bits(x: float) -> int:
    return struct.unpack("<I", struct.pack("<f", float(np.float32(x))))[0]


def _swizzle_get(sw: Sequence[int]) -> int:
    comps = [c for c in sw if c in (0, 1, 2)]       # execution.rs:139-147: other indices are skipped
    n = len(comps)
    if n == 0 or n > 3:
        return 7                                     # -> Value::broadcast(0.0) (execution.rs:150-155)
    a = n
    for i, c in enumerate(comps):
        a |= c << (3 + 2 * i)
    return a


def _swizzle_set(sw: Sequence[int]) -> int:
    n = len(sw) if 1 <= len(sw) <= 3 else 0          # execution.rs:163-168
    a = n
    for i, c in enumerate(list(sw)[:n]):
        a |= (c if c in (0, 1, 2) else 3) << (3 + 2 * i)
    return a


class Program:
    """rusteria::Program (program.rs:7-29): the members the rasterizer path reads."""

    def __init__(self, user_functions: List[list], shade_index: Optional[int] = 0, shade_locals: int = 0, globals_: int = 0):
        self.user_functions = user_functions
        self.shade_index = shade_index
        self.shade_locals = shade_locals
        self.globals = globals_

    # program.rs:44-55: only the top level of shade() is scanned
    def shader_supports_opacity(self) -> bool:
        if self.shade_index is None:
            return False
        return any(op[0] == "SetOpacity" for op in self.user_functions[self.shade_index])

    # ---------------------------------------------------------------- flat code (device)
    def _reachable(self):
        seen, order, todo = set(), [], [self.shade_index]
        while todo:
            f = todo.pop()
            if f in seen:
                continue
            if not (0 <= f < len(self.user_functions)):
                raise ValueError(f"FunctionCall to missing function {f} (the reference panics)")
            seen.add(f)
            order.append(f)

            def walk(ops):
                for op in ops:
                    if op[0] == "FunctionCall":
                        todo.append(op[3])
                    elif op[0] == "If":
                        walk(op[1])
                        if op[2] is not None:
                            walk(op[2])
                    elif op[0] == "For":
                        for part in op[1:5]:
                            walk(part)
            walk(self.user_functions[f])
        return order

    def _call_depth(self, f, stack=()):
        if f in stack:
            raise ValueError("recursive shader functions are not supported on the device")
        if not (0 <= f < len(self.user_functions)):
            raise ValueError(f"FunctionCall to missing function {f} (the reference panics)")
        d = 1

        def walk(ops):
            nonlocal d
            for op in ops:
                if op[0] == "FunctionCall":
                    d = max(d, 1 + self._call_depth(op[3], stack + (f,)))
                elif op[0] == "If":
                    walk(op[1])
                    if op[2] is not None:
                        walk(op[2])
                elif op[0] == "For":
                    for part in op[1:5]:
                        walk(part)
        walk(self.user_functions[f])
        return d

    def flatten(self) -> "FlatProgram":
        if self.shade_index is None:
            return FlatProgram(np.zeros(0, np.uint32), 0, 0, self.globals, False, None)
        if self._call_depth(self.shade_index) > MAX_CALL_DEPTH:
            raise ValueError("call depth exceeds the device limit")
        if self.globals > MAX_GLOBALS or self.shade_locals > MAX_LOCALS:
            raise ValueError("too many globals/locals for the device VM")
        words: List[int] = []
        fixups = []      # (word position of a CALL target, function index)
        offsets = {}

        def emit(name, a=0):
            assert 0 <= a < (1 << 24), (name, a)
            words.append(OPCODE[name] | (a << 8))

        def lower(ops, n_locals, loop_depth):
            for op in ops:
                name = op[0]
                if name in HOST_ONLY:
                    raise ValueError(f"{name} is a texture-baking op, not available in shade() on the device")
                if name in ("LoadLocal", "StoreLocal"):
                    if op[1] >= n_locals:
                        raise ValueError(f"{name}({op[1]}) outside the frame's {n_locals} locals (the reference panics)")
                    emit(name, op[1])
                elif name in ("LoadGlobal", "StoreGlobal"):
                    if op[1] >= self.globals:
                        raise ValueError(f"{name}({op[1]}) outside the program's {self.globals} globals (the reference panics)")
                    emit(name, op[1])
                elif name == "GetComponents":
                    emit(name, _swizzle_get(op[1]))
                elif name == "SetComponents":
                    emit(name, _swizzle_set(op[1]))
                elif name == "Push":
                    emit(name)
                    words.extend(f32_bits(c) for c in op[1])
                elif name == "FunctionCall":
                    arity, total, index = op[1], op[2], op[3]
                    if total > MAX_LOCALS or arity > 255:
                        raise ValueError("too many locals in a function call for the device VM")
                    emit(name, arity | (total << 8))
                    fixups.append((len(words), index))
                    words.append(0)
                elif name == "If":          # execution.rs:286-293
                    emit("Jz")
                    jz = len(words) - 1
                    lower(op[1], n_locals, loop_depth)
                    if op[2] is not None:
                        emit("Jmp")
                        jmp = len(words) - 1
                        words[jz] |= len(words) << 8
                        lower(op[2], n_locals, loop_depth)
                        words[jmp] |= len(words) << 8
                    else:
                        words[jz] |= len(words) << 8
                elif name == "For":         # execution.rs:259-285
                    if loop_depth + 1 > MAX_LOOP_DEPTH:
                        raise ValueError("loop nesting exceeds the device limit")
                    emit("Mark")
                    lower(op[1], n_locals, loop_depth + 1)
                    emit("Trunc")
                    top = len(words)
                    lower(op[2], n_locals, loop_depth + 1)
                    emit("Jz")
                    jz = len(words) - 1
                    emit("Trunc")
                    lower(op[4], n_locals, loop_depth + 1)
                    emit("Trunc")
                    lower(op[3], n_locals, loop_depth + 1)
                    emit("Trunc")
                    emit("Jmp", top)
                    words[jz] |= len(words) << 8
                    emit("Unmark")
                else:
                    if name not in OPCODE or name in ("Jz", "Jmp", "Mark", "Trunc", "Unmark", "End"):
                        raise ValueError(f"unknown op {name}")
                    emit(name)

        order = self._reachable()
        n_locals_of = {self.shade_index: self.shade_locals}

        def note_calls(ops):
            for op in ops:
                if op[0] == "FunctionCall":
                    n_locals_of[op[3]] = max(n_locals_of.get(op[3], 0), op[2])
                elif op[0] == "If":
                    note_calls(op[1])
                    if op[2] is not None:
                        note_calls(op[2])
                elif op[0] == "For":
                    for part in op[1:5]:
                        note_calls(part)
        for f in order:
            note_calls(self.user_functions[f])
        for f in order:
            offsets[f] = len(words)
            lower(self.user_functions[f], n_locals_of.get(f, 0), 0)
            emit("End")
        for pos, index in fixups:
            words[pos] = offsets[index]
        return FlatProgram(np.array(words, dtype=np.uint32), offsets[self.shade_index], self.shade_locals, self.globals,
                           self.shader_supports_opacity(), self)

    # ---------------------------------------------------------------- tree code (oracle)
    def encode_tree(self) -> np.ndarray:
        """[n_functions, shade_index or 0xFFFFFFFF, shade_locals, globals, sets_opacity, offset_0 .. offset_{n-1}, code...];
        a function is [n_words, ops...]; If = [op, n_then, n_else or 0xFFFFFFFF, then.., else..];
        For = [op, n_init, n_cond, n_incr, n_body, init.., cond.., incr.., body..]; Push = [op, x, y, z];
        FunctionCall = [op, arity, total_locals, index]; Get/SetComponents = [op, n, c0..]."""
        def enc(ops):
            out: List[int] = []
            for op in ops:
                name = op[0]
                code = OPCODE[name]
                if name in ("LoadLocal", "StoreLocal", "LoadGlobal", "StoreGlobal"):
                    out += [code, op[1]]
                elif name in ("GetComponents", "SetComponents"):
                    out += [code, len(op[1])] + [int(c) & 0xFF for c in op[1]]
                elif name == "Push":
                    out += [code] + [f32_bits(c) for c in op[1]]
                elif name == "FunctionCall":
                    out += [code, op[1], op[2], op[3]]
                elif name == "If":
                    t = enc(op[1])
                    e = enc(op[2]) if op[2] is not None else None
                    out += [code, len(t), 0xFFFFFFFF if e is None else len(e)] + t + (e or [])
                elif name == "For":
                    parts = [enc(p) for p in op[1:5]]
                    out += [code] + [len(p) for p in parts]
                    for p in parts:
                        out += p
                else:
                    out.append(code)
            return out
        bodies = [enc(f) for f in self.user_functions]
        head = [len(bodies), 0xFFFFFFFF if self.shade_index is None else self.shade_index, self.shade_locals, self.globals,
                1 if self.shader_supports_opacity() else 0]
        offs, code = [], []
        base = len(head) + len(bodies)
        for b in bodies:
            offs.append(base + len(code))
            code += [len(b)] + b
        return np.array(head + offs + code, dtype=np.uint32)


class FlatProgram:
    def __init__(self, words, entry, shade_locals, n_globals, sets_opacity, source):
        self.words = np.ascontiguousarray(words, dtype=np.uint32)
        self.entry = int(entry)
        self.shade_locals = int(shade_locals)
        self.n_globals = int(n_globals)
        self.sets_opacity = bool(sets_opacity)
        self.source = source


# ---------------------------------------------------------------------------------------------
# small assembler helpers for writing shade() bodies by hand
# ---------------------------------------------------------------------------------------------
def push(x, y=None, z=None):
    if y is None:
        y, z = x, x          # scalars are broadcast by the compiler (Value::broadcast)
    return ("Push", (float(x), float(y), float(0.0 if z is None else z)))


def ops(*names):
    return [(n,) for n in names]


# ---------------------------------------------------------------------------------------------
# Python interpreters (float32), used by the CPU tests to pin the flattening against the tree form
# ---------------------------------------------------------------------------------------------
F = np.float32


class VMState:
    """rusteria::Execution (execution.rs:8-55)."""

    def __init__(self, n_globals=0):
        z = lambda: np.zeros(3, F)
        self.globals = [z() for _ in range(n_globals)]
        self.locals: List[np.ndarray] = []
        self.stack: List[np.ndarray] = []
        self.return_value = None
        self.uv, self.color, self.metallic, self.emissive = z(), z(), z(), z()
        self.roughness = np.full(3, 0.5, F)
        self.opacity, self.bump, self.normal, self.hitpoint, self.time = z(), z(), z(), z(), z()
        self.patterns: list = []          # [(w, h, data[h*w,3])]
        self.patterns_normal: list = []
        self.palette: list = []           # [None or (r,g,b)]


def _v(x, y, z):
    return np.array([x, y, z], dtype=F)


def _sample(tex, uv):  # textures/mod.rs:20-24, 131-146
    w, h, data = tex
    with np.errstate(all="ignore"):
        u = F(uv[0]) - np.floor(F(uv[0]))
        v = F(uv[1]) - np.floor(F(uv[1]))
        x = int(np.floor(F(u * F(w)))) if np.isfinite(u) else 0
        y = int(np.floor(F(v * F(h)))) if np.isfinite(v) else 0
    x %= w
    y %= h
    return np.array(data[y * w + x], dtype=F)


def _as_usize(x):
    if not np.isfinite(x):
        return 0 if (np.isnan(x) or x < 0) else (1 << 63)
    return max(0, int(x))


def _simple(st: VMState, name: str):
    s = st.stack
    pop = s.pop
    with np.errstate(all="ignore"):
        if name == "Swap":
            b, a = pop(), pop(); s.append(b); s.append(a)
        elif name == "Clear":
            if s:
                pop()
        elif name == "Dup":
            if s:
                s.append(s[-1].copy())
        elif name == "Pack2":
            y, x = pop(), pop(); s.append(_v(x[0], y[0], 0))
        elif name == "Pack3":
            z, y, x = pop(), pop(), pop(); s.append(_v(x[0], y[0], z[0]))
        elif name in ("Add", "Sub", "Mul", "Div"):
            b, a = pop(), pop()
            s.append({"Add": a + b, "Sub": a - b, "Mul": a * b, "Div": a / b}[name].astype(F))
        elif name == "Length":
            a = pop(); s.append(np.full(3, np.sqrt(F(a[0] * a[0] + a[1] * a[1]) + a[2] * a[2]), F))
        elif name == "Length2":
            a = pop(); s.append(_v(np.sqrt(F(a[0] * a[0] + a[1] * a[1])), 0, 0))
        elif name == "Length3":
            a = pop(); s.append(_v(np.sqrt(F(F(a[0] * a[0] + a[1] * a[1]) + a[2] * a[2])), 0, 0))
        elif name == "Abs":
            s.append(np.abs(pop()))
        elif name == "Sin":
            s.append(np.sin(pop()).astype(F))
        elif name in ("Sin1", "Cos1"):      # Cos1 / Cos2 compute sin in the reference (execution.rs:342-349)
            a = pop(); s.append(_v(np.sin(a[0]), 0, 0))
        elif name in ("Sin2", "Cos2"):
            a = pop(); s.append(_v(np.sin(a[0]), np.sin(a[1]), 0))
        elif name == "Cos":
            s.append(np.cos(pop()).astype(F))
        elif name == "Tan":
            s.append(np.tan(pop()).astype(F))
        elif name == "Atan":
            s.append(np.arctan(pop()).astype(F))
        elif name == "Atan2":
            b, a = pop(), pop(); s.append(np.arctan2(a, b).astype(F))
        elif name == "Rotate2D":
            ang, v = pop(), pop()
            rad = F(ang[0] * F(math.pi / 180.0))
            sn, cs = F(np.sin(rad)), F(np.cos(rad))
            s.append(_v(F(v[0] * cs) - F(v[1] * sn), F(v[0] * sn) + F(v[1] * cs), v[2]))
        elif name == "Dot":
            b, a = pop(), pop(); s.append(np.full(3, F(F(a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]), F))
        elif name == "Dot2":
            b, a = pop(), pop(); s.append(_v(F(a[0] * b[0]) + F(a[1] * b[1]), 0, 0))
        elif name == "Dot3":
            b, a = pop(), pop(); s.append(_v(F(F(a[0] * b[0]) + F(a[1] * b[1])) + F(a[2] * b[2]), 0, 0))
        elif name == "Cross":
            b, a = pop(), pop()
            s.append(_v(F(a[1] * b[2]) - F(a[2] * b[1]), F(a[2] * b[0]) - F(a[0] * b[2]), F(a[0] * b[1]) - F(a[1] * b[0])))
        elif name == "Normalize":
            a = pop(); ln = np.sqrt(F(F(a[0] * a[0] + a[1] * a[1]) + a[2] * a[2]))
            s.append((a / ln).astype(F) if ln > 0 else a)
        elif name == "Floor":
            s.append(np.floor(pop()))
        elif name == "Ceil":
            s.append(np.ceil(pop()))
        elif name == "Round":
            a = pop(); s.append((np.sign(a) * np.floor(np.abs(a) + F(0.5))).astype(F))
        elif name == "Fract":
            a = pop(); s.append((a - np.floor(a)).astype(F))
        elif name == "Mod":
            b, a = pop(), pop(); s.append((a - b * np.floor(a / b)).astype(F))
        elif name == "Radians":
            s.append((pop() * F(math.pi / 180.0)).astype(F))
        elif name == "Degrees":
            s.append((pop() * F(180.0 / math.pi)).astype(F))
        elif name == "Min":
            b, a = pop(), pop(); s.append(np.fmin(a, b))
        elif name == "Max":
            b, a = pop(), pop(); s.append(np.fmax(a, b))
        elif name == "Mix":
            c, b, a = pop(), pop(), pop(); s.append((a + (b - a) * c).astype(F))
        elif name == "Smoothstep":
            c, b, a = pop(), pop(), pop()
            den = F(b[0] - a[0])
            t = F(F(c[0] - a[0]) / den) if den != 0 else F(0)
            t = F(0) if t < 0 else (F(1) if t > 1 else t)
            s.append(np.full(3, F(F(t * t) * F(F(3.0) - F(F(2.0) * t))), F))
        elif name == "Step":
            b, a = pop(), pop(); s.append((b >= a).astype(F))
        elif name == "Clamp":
            c, b, a = pop(), pop(), pop()
            s.append(np.where(a < b, b, np.where(a > c, c, a)).astype(F))
        elif name == "Sqrt":
            s.append(np.sqrt(pop()))
        elif name == "Log":
            s.append(np.log(pop()).astype(F))
        elif name == "Pow":
            b, a = pop(), pop(); s.append(np.power(a, b).astype(F))
        elif name in ("Eq", "Ne", "Lt", "Le", "Gt", "Ge"):
            b, a = pop(), pop()
            r = {"Eq": a[0] == b[0], "Ne": a[0] != b[0], "Lt": a[0] < b[0], "Le": a[0] <= b[0], "Gt": a[0] > b[0], "Ge": a[0] >= b[0]}[name]
            s.append(np.full(3, 1.0 if r else 0.0, F))
        elif name == "And":
            b, a = pop(), pop(); s.append(np.full(3, 1.0 if (a[0] != 0) and (b[0] != 0) else 0.0, F))
        elif name == "Or":
            b, a = pop(), pop(); s.append(np.full(3, 1.0 if (a[0] != 0) or (b[0] != 0) else 0.0, F))
        elif name == "Not":
            a = pop(); s.append(np.full(3, 1.0 if a[0] == 0 else 0.0, F))
        elif name == "Neg":
            s.append(-pop())
        elif name == "Print":
            pop()
        elif name in ("UV", "Normal", "Hitpoint", "Time", "Color", "Roughness", "Metallic", "Emissive", "Opacity", "Bump"):
            s.append(getattr(st, name.lower()).copy())
        elif name == "SetNormal":
            a = pop(); st.normal = (a / np.sqrt(F(F(a[0] * a[0] + a[1] * a[1]) + a[2] * a[2]))).astype(F)
        elif name in ("SetUV", "SetColor", "SetRoughness", "SetMetallic", "SetEmissive", "SetOpacity", "SetBump"):
            setattr(st, name[3:].lower(), pop())
        elif name == "Sample":
            b, a = pop(), pop()
            i = _as_usize(b[0])
            s.append(_sample(st.patterns[i], a) if i < len(st.patterns) else np.zeros(3, F))
        elif name == "SampleNormal":
            b, a = pop(), pop()
            i = _as_usize(b[0])
            s.append((_sample(st.patterns_normal[i], a) * F(2.0) - F(1.0)).astype(F) if i < len(st.patterns_normal) else np.zeros(3, F))
        elif name == "PaletteIndex":
            a = pop()
            i = _as_usize(a[0])
            if i < len(st.palette) and st.palette[i] is not None:
                s.append(np.array(st.palette[i], dtype=F))
        else:
            raise ValueError(name)


def run_tree(prog: Program, st: VMState):
    """Execution::shade (execution.rs:771-779) over the op tree."""
    st.stack.clear()
    st.return_value = None
    while len(st.locals) < prog.shade_locals:
        st.locals.append(np.zeros(3, F))
    del st.locals[prog.shade_locals:]
    if len(st.globals) != prog.globals:
        st.globals = (st.globals + [np.zeros(3, F) for _ in range(prog.globals)])[:prog.globals]

    def execute(code):
        for op in code:
            if st.return_value is not None:
                break
            name = op[0]
            if name == "LoadGlobal":
                st.stack.append(st.globals[op[1]].copy())
            elif name == "StoreGlobal":
                st.globals[op[1]] = st.stack.pop()
            elif name == "LoadLocal":
                st.stack.append(st.locals[op[1]].copy())
            elif name == "StoreLocal":
                st.locals[op[1]] = st.stack.pop()
            elif name == "GetComponents":
                v = st.stack.pop()
                r = [v[c] for c in op[1] if c in (0, 1, 2)]
                st.stack.append(np.full(3, r[0], F) if len(r) == 1 else _v(r[0], r[1], 0) if len(r) == 2 else _v(*r) if len(r) == 3 else np.zeros(3, F))
            elif name == "SetComponents":
                value, target = st.stack.pop(), st.stack.pop().copy()
                n = len(op[1]) if 1 <= len(op[1]) <= 3 else 0
                for i, idx in enumerate(op[1]):
                    if i >= n:
                        break
                    if idx in (0, 1, 2):
                        target[idx] = value[i]
                st.stack.append(target)
            elif name == "Push":
                st.stack.append(np.array(op[1], dtype=F))
            elif name == "FunctionCall":
                saved = st.locals
                st.locals = [np.zeros(3, F) for _ in range(op[2])]
                for i in reversed(range(op[1])):
                    if st.stack:
                        st.locals[i] = st.stack.pop()
                base = len(st.stack)
                execute(prog.user_functions[op[3]])
                if st.return_value is not None:
                    ret, st.return_value = st.return_value, None
                elif len(st.stack) > base:
                    ret = st.stack.pop()
                else:
                    ret = np.zeros(3, F)
                del st.stack[base:]
                st.locals = saved
                st.stack.append(ret)
            elif name == "Return":
                st.return_value = st.stack.pop() if st.stack else np.zeros(3, F)
                break
            elif name == "If":
                if st.stack.pop()[0] != 0.0:
                    execute(op[1])
                elif op[2] is not None:
                    execute(op[2])
            elif name == "For":
                base = len(st.stack)
                execute(op[1])
                del st.stack[base:]
                while True:
                    execute(op[2])
                    if st.return_value is not None:   # device rule: a Return inside a loop leaves the function
                        break
                    if st.stack.pop()[0] == 0.0:
                        break
                    del st.stack[base:]
                    execute(op[4]); del st.stack[base:]
                    execute(op[3]); del st.stack[base:]
            else:
                _simple(st, name)
    execute(prog.user_functions[prog.shade_index])


def run_flat(fp: FlatProgram, st: VMState):
    """Interpreter of the flat code, the algorithm of the device VM (rx_kernels.cu vm_run)."""
    w = fp.words
    st.stack.clear()
    st.locals = [np.zeros(3, F) for _ in range(fp.shade_locals)]
    st.globals = [np.zeros(3, F) for _ in range(fp.n_globals)]
    frames, marks = [], []
    pc = fp.entry
    ret_val = None
    budget = 1 << 20
    while budget > 0:
        budget -= 1
        word = int(w[pc]); pc += 1
        name, a = OPS[word & 0xFF], word >> 8
        if name == "LoadGlobal":
            st.stack.append(st.globals[a].copy())
        elif name == "StoreGlobal":
            st.globals[a] = st.stack.pop()
        elif name == "LoadLocal":
            st.stack.append(st.locals[a].copy())
        elif name == "StoreLocal":
            st.locals[a] = st.stack.pop()
        elif name == "GetComponents":
            v = st.stack.pop()
            n = a & 7
            c = [(a >> (3 + 2 * i)) & 3 for i in range(3)]
            st.stack.append(np.full(3, v[c[0]], F) if n == 1 else _v(v[c[0]], v[c[1]], 0) if n == 2 else _v(v[c[0]], v[c[1]], v[c[2]]) if n == 3 else np.zeros(3, F))
        elif name == "SetComponents":
            value, target = st.stack.pop(), st.stack.pop().copy()
            for i in range(a & 7):
                idx = (a >> (3 + 2 * i)) & 3
                if idx < 3:
                    target[idx] = value[i]
            st.stack.append(target)
        elif name == "Push":
            st.stack.append(np.array(w[pc:pc + 3], dtype=np.uint32).view(F).copy()); pc += 3
        elif name == "FunctionCall":
            target = int(w[pc]); pc += 1
            arity, total = a & 0xFF, a >> 8
            new_locals = [np.zeros(3, F) for _ in range(total)]
            for i in reversed(range(arity)):
                if st.stack:
                    new_locals[i] = st.stack.pop()
            frames.append((pc, st.locals, len(st.stack), len(marks)))
            st.locals = new_locals
            pc = target
        elif name in ("Return", "End"):
            if name == "Return":
                ret_val = st.stack.pop() if st.stack else np.zeros(3, F)
            if not frames:
                return
            rpc, saved, base, nmarks = frames.pop()
            if ret_val is not None:
                ret, ret_val = ret_val, None
            elif len(st.stack) > base:
                ret = st.stack.pop()
            else:
                ret = np.zeros(3, F)
            del st.stack[base:]
            del marks[nmarks:]
            st.locals = saved
            st.stack.append(ret)
            pc = rpc
        elif name == "Jz":
            if st.stack.pop()[0] == 0.0:
                pc = a
        elif name == "Jmp":
            pc = a
        elif name == "Mark":
            marks.append(len(st.stack))
        elif name == "Trunc":
            del st.stack[marks[-1]:]
        elif name == "Unmark":
            marks.pop()
        else:
            _simple(st, name)
    raise RuntimeError("VM op budget exhausted")


# ---------------------------------------------------------------------------------------------
# Expression builder: writes the postfix op sequences the Rusteria compiler emits for expressions
# (compile.rs), so shade() bodies can be written in Python instead of by hand.
# ---------------------------------------------------------------------------------------------
class X:
    """An expression = the op sequence that leaves its value on the stack."""

    def __init__(self, seq):
        self.ops = list(seq)

    @staticmethod
    def of(v) -> "X":
        if isinstance(v, X):
            return v
        if isinstance(v, (tuple, list)):
            c = list(v) + [0.0] * (3 - len(v))
            return X([("Push", (float(c[0]), float(c[1]), float(c[2])))])
        return X([push(v)])

    def _bin(self, o, name):
        return X(self.ops + X.of(o).ops + [(name,)])

    def _rbin(self, o, name):
        return X(X.of(o).ops + self.ops + [(name,)])

    def __add__(self, o): return self._bin(o, "Add")
    def __radd__(self, o): return self._rbin(o, "Add")
    def __sub__(self, o): return self._bin(o, "Sub")
    def __rsub__(self, o): return self._rbin(o, "Sub")
    def __mul__(self, o): return self._bin(o, "Mul")
    def __rmul__(self, o): return self._rbin(o, "Mul")
    def __truediv__(self, o): return self._bin(o, "Div")
    def __rtruediv__(self, o): return self._rbin(o, "Div")
    def __neg__(self): return X(self.ops + [("Neg",)])
    def __lt__(self, o): return self._bin(o, "Lt")
    def __le__(self, o): return self._bin(o, "Le")
    def __gt__(self, o): return self._bin(o, "Gt")
    def __ge__(self, o): return self._bin(o, "Ge")
    def eq(self, o): return self._bin(o, "Eq")
    def ne(self, o): return self._bin(o, "Ne")
    def and_(self, o): return self._bin(o, "And")
    def or_(self, o): return self._bin(o, "Or")
    def not_(self): return X(self.ops + [("Not",)])

    def swz(self, comps): return X(self.ops + [("GetComponents", list(comps))])
    x = property(lambda self: self.swz([0]))
    y = property(lambda self: self.swz([1]))
    z = property(lambda self: self.swz([2]))
    xy = property(lambda self: self.swz([0, 1]))


def _fn(name, arity):
    def f(*args):
        assert len(args) == arity, name
        seq = []
        for a in args:
            seq += X.of(a).ops
        return X(seq + [(name,)])
    return f


uv, normal, hitpoint, time_, color, roughness, metallic, emissive, opacity, bump = (X([(n,)]) for n in (
    "UV", "Normal", "Hitpoint", "Time", "Color", "Roughness", "Metallic", "Emissive", "Opacity", "Bump"))
sin, cos, tan, atan, abs_, floor, ceil, round_, fract, sqrt, log, length, normalize, radians, degrees = (_fn(n, 1) for n in (
    "Sin", "Cos", "Tan", "Atan", "Abs", "Floor", "Ceil", "Round", "Fract", "Sqrt", "Log", "Length", "Normalize", "Radians", "Degrees"))
atan2, dot, cross, mod, min_, max_, step, pow_, rotate2d = (_fn(n, 2) for n in (
    "Atan2", "Dot", "Cross", "Mod", "Min", "Max", "Step", "Pow", "Rotate2D"))
mix, smoothstep, clamp = (_fn(n, 3) for n in ("Mix", "Smoothstep", "Clamp"))
vec2 = lambda a, b: X(X.of(a).ops + X.of(b).ops + [("Pack2",)])
vec3 = lambda a, b, c: X(X.of(a).ops + X.of(b).ops + X.of(c).ops + [("Pack3",)])
PATTERN_INDEX = {"value": 0, "fbm_value": 1, "perlin": 2, "fbm_perlin": 3, "bricks": 4, "tiles": 5, "blocks": 6}  # patterns.rs:27-37


def sample(uv_expr, pattern):
    i = PATTERN_INDEX[pattern] if isinstance(pattern, str) else pattern
    return X(X.of(uv_expr).ops + [push(float(i)), ("Sample",)])


def sample_normal(uv_expr, pattern):
    i = PATTERN_INDEX[pattern] if isinstance(pattern, str) else pattern
    return X(X.of(uv_expr).ops + [push(float(i)), ("SampleNormal",)])


def palette(index):
    return X(X.of(index).ops + [("PaletteIndex",)])


def call(fn_index, total_locals, *args):
    seq = []
    for a in args:
        seq += X.of(a).ops
    return X(seq + [("FunctionCall", len(args), total_locals, fn_index)])


class Body:
    """Statement list of one function; `let` allocates a local slot like the compiler does."""

    def __init__(self, n_params=0):
        self.code = []
        self.n_locals = n_params

    def param(self, i):
        return X([("LoadLocal", i)])

    def let(self, expr) -> X:
        i = self.n_locals
        self.n_locals += 1
        self.code += X.of(expr).ops + [("StoreLocal", i)]
        return X([("LoadLocal", i)])

    def assign(self, local: X, expr):
        self.code += X.of(expr).ops + [("StoreLocal", local.ops[0][1])]

    def set(self, what, expr):
        self.code += X.of(expr).ops + [("Set" + what,)]

    def set_global(self, i, expr):
        self.code += X.of(expr).ops + [("StoreGlobal", i)]

    def ret(self, expr):
        self.code += X.of(expr).ops + [("Return",)]

    def if_(self, cond, then_body: "Body", else_body: "Body" = None):
        self.code += X.of(cond).ops + [("If", then_body.code, None if else_body is None else else_body.code)]

    def for_(self, init: "Body", cond, incr: "Body", body: "Body"):
        self.code += [("For", init.code, X.of(cond).ops, incr.code, body.code)]

    def sub(self) -> "Body":
        """A nested block sharing this function's local slots."""
        b = Body()
        b.n_locals = self.n_locals
        b._parent = self
        return b

    def close(self, child: "Body"):
        self.n_locals = max(self.n_locals, child.n_locals)
