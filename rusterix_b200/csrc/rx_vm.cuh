// rx_vm.cuh -- the Rusteria VM on the device: an interpreter of the flat code of include/rxcuda.h
// (rxc_program), one call per fragment.  It restates Execution::execute / Execution::shade
// (reference rusteria/src/node/execution.rs:109-779) op by op; control flow arrives lowered to jumps
// (rusterix_b200/vm.py, Program.flatten).
//
// Deviations from the reference, all documented in DESIGN.md:
//  * `Execution` is created once per screen tile and never reset (src/rasterizer.rs:310), so globals,
//    locals, emissive, bump, uv.z ... written by one fragment leak into the next fragment shaded in the same
//    tile.  In k_raster every fragment starts from Execution::new(): the state is per fragment.  Scenes whose programs
//    can observe the difference (rxj_state_report) are rendered by k_raster_ordered instead, which carries one
//    Execution per tile like the reference (rxc_set_vm_state_mode, on by default).
//  * a `Return` inside a `For` leaves the function (the reference pops the loop condition from an
//    empty or foreign stack after it: a panic or garbage).
//  * sin/cos/tan/atan/atan2/pow/ln come from CUDA's libm instead of the host's (<= 2 ulp apart).
#pragma once
#include "rx_device.cuh"

#define RXVM_STACK 32
#define RXVM_LOCAL_POOL 96     // locals of all live frames (<= 32 per frame)
#define RXVM_GLOBALS 16
#define RXVM_FRAMES 8
#define RXVM_MARKS 8
#define RXVM_BUDGET (1u << 20)

struct VmIO {   // the Execution fields the rasterizer reads and writes (execution.rs:27-55)
    f3 uv, color, normal, hitpoint, time, opacity, roughness, metallic, emissive, bump;
};
// What else an Execution carries from one shade() to the next when it is never reset (the reference's per-tile Execution,
// src/rasterizer.rs:310; only the reference-order kernel k_raster_ordered keeps one): `reset` and `shade` RESIZE the globals and
// the locals of shade()'s frame (execution.rs:103-107, :773), they do not clear them.
struct VmPersist {
    f3 globals[16];
    f3 locals[32];
    uint32_t n_globals, n_locals;
};

__device__ __forceinline__ void vm_io_reset(VmIO& io) {  // Execution::new, execution.rs:58-77
    const f3 z = {0.0f, 0.0f, 0.0f};
    io.uv = z; io.color = z; io.normal = z; io.hitpoint = z; io.time = z; io.opacity = z;
    io.roughness = {0.5f, 0.5f, 0.5f}; io.metallic = z; io.emissive = z; io.bump = z;
}

// `x as usize` (saturating, NaN -> 0) clamped to 2^31
__device__ __forceinline__ uint32_t vm_as_index(float x) {
    if (!(x == x) || x <= 0.0f) return 0u;
    if (x >= 2147483648.0f) return 0x80000000u;
    return (uint32_t)x;
}

// TexStorage::sample (rusteria/src/textures/mod.rs:20-24, :131-146)
__device__ __forceinline__ f3 vm_pattern_sample(const VmDev& vm, const DPattern& p, f3 uv) {
    float u = uv.x - floorf(uv.x), v = uv.y - floorf(uv.y);
    const float fx = floorf(u * (float)p.width), fy = floorf(v * (float)p.height);
    int x = (fx == fx) ? (int)fminf(fmaxf(fx, -2147483648.0f), 2147483520.0f) : 0;  // `as i32`
    int y = (fy == fy) ? (int)fminf(fmaxf(fy, -2147483648.0f), 2147483520.0f) : 0;
    x %= (int)p.width; if (x < 0) x += (int)p.width;
    y %= (int)p.height; if (y < 0) y += (int)p.height;
    const float* d = vm.pattern_data + ((size_t)p.off + (size_t)y * p.width + (size_t)x) * 3;
    return {__ldg(d), __ldg(d + 1), __ldg(d + 2)};
}

// Transcendental and other rarely executed component-wise ops.  The interpreter calls them out of line (vm_math1 / vm_math2):
// inlined three components at a time they are most of its code size, and the dispatch loop then no longer fits the instruction
// cache.  Generated code (rx_jit.cu) inlines them: `op` is a constant there and only that op's arithmetic remains.
__device__ __forceinline__ f3 vm_math1_inline(uint32_t op, f3 a) {
    switch (op) {
        case RXVM_SIN: return {sinf(a.x), sinf(a.y), sinf(a.z)};
        case RXVM_SIN1: case RXVM_COS1: return {sinf(a.x), 0.0f, 0.0f};          // Cos1/Cos2 call sin (execution.rs:342-349)
        case RXVM_SIN2: case RXVM_COS2: return {sinf(a.x), sinf(a.y), 0.0f};
        case RXVM_COS: return {cosf(a.x), cosf(a.y), cosf(a.z)};
        case RXVM_TAN: return {tanf(a.x), tanf(a.y), tanf(a.z)};
        case RXVM_ATAN: return {atanf(a.x), atanf(a.y), atanf(a.z)};
        case RXVM_SQRT: return {sqrtf(a.x), sqrtf(a.y), sqrtf(a.z)};
        case RXVM_LOG: return {logf(a.x), logf(a.y), logf(a.z)};
        case RXVM_ROUND: return {roundf(a.x), roundf(a.y), roundf(a.z)};
        case RXVM_CEIL: return {ceilf(a.x), ceilf(a.y), ceilf(a.z)};
        case RXVM_RADIANS: return {a.x * 0.017453292519943295f, a.y * 0.017453292519943295f, a.z * 0.017453292519943295f};
        case RXVM_DEGREES: return {a.x * 57.29577951308232f, a.y * 57.29577951308232f, a.z * 57.29577951308232f};
        case RXVM_LENGTH: { const float m = sqrtf(a.x * a.x + a.y * a.y + a.z * a.z); return {m, m, m}; }
        case RXVM_LENGTH2: return {sqrtf(a.x * a.x + a.y * a.y), 0.0f, 0.0f};
        case RXVM_LENGTH3: return {sqrtf(a.x * a.x + a.y * a.y + a.z * a.z), 0.0f, 0.0f};
        case RXVM_NORMALIZE: { const float m = sqrtf(a.x * a.x + a.y * a.y + a.z * a.z); if (m > 0.0f) return {a.x / m, a.y / m, a.z / m}; return a; }
        default: return a;
    }
}

__device__ __forceinline__ f3 vm_math2_inline(uint32_t op, f3 a, f3 b) {
    switch (op) {
        case RXVM_ATAN2: return {atan2f(a.x, b.x), atan2f(a.y, b.y), atan2f(a.z, b.z)};
        case RXVM_POW: return {powf(a.x, b.x), powf(a.y, b.y), powf(a.z, b.z)};
        case RXVM_MOD: return {a.x - b.x * floorf(a.x / b.x), a.y - b.y * floorf(a.y / b.y), a.z - b.z * floorf(a.z / b.z)};
        case RXVM_DIV: return {a.x / b.x, a.y / b.y, a.z / b.z};
        case RXVM_ROTATE2D: {         // execution.rs:375-383: a = vector, b = angle in degrees
            const float rad = b.x * 0.017453292519943295f;
            float sn, cs;
            sincosf(rad, &sn, &cs);
            return {a.x * cs - a.y * sn, a.x * sn + a.y * cs, a.z};
        }
        case RXVM_CROSS: return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
        default: return a;
    }
}

__device__ __noinline__ f3 vm_math1(uint32_t op, f3 a) { return vm_math1_inline(op, a); }
__device__ __noinline__ f3 vm_math2(uint32_t op, f3 a, f3 b) { return vm_math2_inline(op, a, b); }

// ---- op semantics shared by the interpreter below and by the code the JIT generates (rx_jit.cpp) ---------------------
// One function per arity; `op` is a compile-time constant in generated code (the switch folds away) and a run-time
// value in the interpreter.  Everything is the reference's arithmetic, op by op (execution.rs:296-742).
#define VM_BOOL3(x) ((x) ? f3{1.0f, 1.0f, 1.0f} : f3{0.0f, 0.0f, 0.0f})
template <bool INL = false>
__device__ __forceinline__ f3 vm_un(uint32_t op, f3 a) {
    switch (op) {
        case RXVM_ABS: return {fabsf(a.x), fabsf(a.y), fabsf(a.z)};
        case RXVM_FLOOR: return {floorf(a.x), floorf(a.y), floorf(a.z)};
        case RXVM_FRACT: return {a.x - floorf(a.x), a.y - floorf(a.y), a.z - floorf(a.z)};
        case RXVM_NOT: return VM_BOOL3(a.x == 0.0f);
        case RXVM_NEG: return {-a.x, -a.y, -a.z};
        default: return INL ? vm_math1_inline(op, a) : vm_math1(op, a);   // the transcendental / rarely executed ones
    }
}
template <bool INL = false>
__device__ __forceinline__ f3 vm_bin(uint32_t op, f3 a, f3 b) {
    switch (op) {
        case RXVM_ADD: return {a.x + b.x, a.y + b.y, a.z + b.z};
        case RXVM_SUB: return {a.x - b.x, a.y - b.y, a.z - b.z};
        case RXVM_MUL: return {a.x * b.x, a.y * b.y, a.z * b.z};
        case RXVM_PACK2: return {a.x, b.x, 0.0f};
        case RXVM_DOT: { const float d = a.x * b.x + a.y * b.y + a.z * b.z; return {d, d, d}; }
        case RXVM_DOT2: return {a.x * b.x + a.y * b.y, 0.0f, 0.0f};
        case RXVM_DOT3: return {a.x * b.x + a.y * b.y + a.z * b.z, 0.0f, 0.0f};
        case RXVM_MIN: return {fminf(a.x, b.x), fminf(a.y, b.y), fminf(a.z, b.z)};
        case RXVM_MAX: return {fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z)};
        case RXVM_STEP: return {b.x >= a.x ? 1.0f : 0.0f, b.y >= a.y ? 1.0f : 0.0f, b.z >= a.z ? 1.0f : 0.0f};
        case RXVM_EQ: return VM_BOOL3(a.x == b.x);
        case RXVM_NE: return VM_BOOL3(a.x != b.x);
        case RXVM_LT: return VM_BOOL3(a.x < b.x);
        case RXVM_LE: return VM_BOOL3(a.x <= b.x);
        case RXVM_GT: return VM_BOOL3(a.x > b.x);
        case RXVM_GE: return VM_BOOL3(a.x >= b.x);
        case RXVM_AND: return VM_BOOL3((a.x != 0.0f) & (b.x != 0.0f));
        case RXVM_OR: return VM_BOOL3((a.x != 0.0f) | (b.x != 0.0f));
        default: return INL ? vm_math2_inline(op, a, b) : vm_math2(op, a, b);   // Atan2, Pow, Mod, Div, Rotate2D, Cross
    }
}
// Generated code names its ops as template arguments: the cheap ones are inlined (`OP` folds the switch away), the libm-sized ones
// (sin / cos / tan / atan / ln, atan2 / pow / the sincos of Rotate2D: a few hundred instructions each with their slow paths) get ONE
// out-of-line instance per op that all call sites of all programs share.  Measured on the batch-shader scene: the same kernel time
// as with everything inlined (0.945 vs 0.944 ms; RXVM_JIT_INLINE_LIBM=1), a smaller kernel and a fifth less compile time.
#ifndef RXVM_JIT_INLINE_LIBM
#define RXVM_JIT_INLINE_LIBM 0
#endif
template <uint32_t OP> __device__ __noinline__ f3 vm_math1_shared(f3 a) { return vm_math1_inline(OP, a); }
template <uint32_t OP> __device__ __noinline__ f3 vm_math2_shared(f3 a, f3 b) { return vm_math2_inline(OP, a, b); }
template <uint32_t OP> __device__ __forceinline__ f3 vm_un_c(f3 a) {
    constexpr bool libm = OP == RXVM_SIN || OP == RXVM_SIN1 || OP == RXVM_SIN2 || OP == RXVM_COS || OP == RXVM_COS1 || OP == RXVM_COS2 || OP == RXVM_TAN ||
                          OP == RXVM_ATAN || OP == RXVM_LOG;
    if (libm && !RXVM_JIT_INLINE_LIBM) return vm_math1_shared<OP>(a);
    return vm_un<true>(OP, a);
}
template <uint32_t OP> __device__ __forceinline__ f3 vm_bin_c(f3 a, f3 b) {
    constexpr bool libm = OP == RXVM_ATAN2 || OP == RXVM_POW || OP == RXVM_ROTATE2D;
    if (libm && !RXVM_JIT_INLINE_LIBM) return vm_math2_shared<OP>(a, b);
    return vm_bin<true>(OP, a, b);
}
__device__ __forceinline__ f3 vm_tern(uint32_t op, f3 a, f3 b, f3 c) {
    switch (op) {
        case RXVM_PACK3: return {a.x, b.x, c.x};
        case RXVM_MIX: return {a.x + (b.x - a.x) * c.x, a.y + (b.y - a.y) * c.y, a.z + (b.z - a.z) * c.z};
        case RXVM_SMOOTHSTEP: {       // execution.rs:458-476
            const float denom = b.x - a.x;
            float u = denom != 0.0f ? (c.x - a.x) / denom : 0.0f;
            if (u < 0.0f) u = 0.0f; else if (u > 1.0f) u = 1.0f;
            const float sm = u * u * (3.0f - 2.0f * u);
            return {sm, sm, sm};
        }
        case RXVM_CLAMP: return {rx_clamp(a.x, b.x, c.x), rx_clamp(a.y, b.y, c.y), rx_clamp(a.z, b.z, c.z)};
        default: return a;
    }
}
__device__ __forceinline__ f3 vm_get_components(uint32_t a24, f3 t) {   // execution.rs:135-157
    const float c[4] = {t.x, t.y, t.z, 0.0f};
    const uint32_t n = a24 & 7u, i0 = (a24 >> 3) & 3u, i1 = (a24 >> 5) & 3u, i2 = (a24 >> 7) & 3u;
    if (n == 1u) return {c[i0], c[i0], c[i0]};
    if (n == 2u) return {c[i0], c[i1], 0.0f};
    if (n == 3u) return {c[i0], c[i1], c[i2]};
    return {0.0f, 0.0f, 0.0f};
}
__device__ __forceinline__ f3 vm_set_components(uint32_t a24, f3 d, f3 value) {   // execution.rs:158-183
    const float c[3] = {value.x, value.y, value.z};
    const uint32_t n = a24 & 7u;
    for (uint32_t i = 0; i < n && i < 3u; ++i) {
        const uint32_t idx = (a24 >> (3u + 2u * i)) & 3u;
        if (idx == 0u) d.x = c[i]; else if (idx == 1u) d.y = c[i]; else if (idx == 2u) d.z = c[i];
    }
    return d;
}
__device__ __forceinline__ f3 vm_sample_op(const VmDev& vm, bool normal_bank, f3 a, f3 b) {   // execution.rs:625-649
    const uint32_t i = vm_as_index(b.x);
    if (!normal_bank) return i < vm.n_patterns ? vm_pattern_sample(vm, vm.patterns[i], a) : f3{0.0f, 0.0f, 0.0f};
    if (i >= vm.n_patterns_normal) return {0.0f, 0.0f, 0.0f};
    const f3 nm3 = vm_pattern_sample(vm, vm.patterns[vm.n_patterns + i], a);
    return {nm3.x * 2.0f - 1.0f, nm3.y * 2.0f - 1.0f, nm3.z * 2.0f - 1.0f};
}

#ifdef RXVM_JIT
// Straight-line C++ generated from the scene's programs (rx_jit.cpp), compiled with NVRTC into its own copy of the VM
// kernels: vm_run_jit(vm, jit_index, io) runs program `jit_index` without fetching or dispatching a single op
// (returns 1 = done, 0 = device limit hit, 2 = hand over to the interpreter).
#include "rx_vm_generated.inc"
#endif

#if defined(RXVM_JIT) && RXVM_JIT_COMPLETE
// every program of the scene is generated code: no interpreter in this kernel
__device__ __forceinline__ bool vm_run(const VmDev& vm, const DProgram& P, VmIO& io) { return vm_run_jit(vm, P.jit_index, io) == 1; }
template <bool PERSIST>
__device__ __forceinline__ bool vm_run_t(const VmDev& vm, const DProgram& P, VmIO& io, VmPersist* ps) {
    return (PERSIST ? vm_run_jit_ps(vm, P.jit_index, io, *ps) : vm_run_jit(vm, P.jit_index, io)) == 1;
}
#else
// Runs the shade function of program P on `io`.  Returns false when a device limit was hit (stack,
// frames, op budget) or the code is malformed; the reference would have panicked or looped.
// PERSIST: globals and shade()'s locals start from `ps` and are left there (see VmPersist); always interpreted.
template <bool PERSIST>
__device__ __noinline__ bool vm_run_t(const VmDev& vm, const DProgram& P, VmIO& io, VmPersist* ps) {
    // The value stack keeps its top in registers (`t`): positions 1..sp-1 live in stack[1..sp-1], position sp in
    // `t`; stack[0] is a dummy that absorbs the spill of an empty stack's top.  A unary op then touches no memory,
    // a binary op loads one operand, a push stores one (half the local-memory traffic of a stack held in memory).
#ifdef RXVM_JIT
    if (P.jit_index != 0xFFFFFFFFu) {   // the program as straight-line code; 2 = its stack took a shape the translator did not verify
        const bool may_bail = vm_jit_may_bail(P.jit_index);
        VmIO saved;
        if (may_bail) saved = io;
        const int st = PERSIST ? vm_run_jit_ps(vm, P.jit_index, io, *ps) : vm_run_jit(vm, P.jit_index, io);   // (ps is only written when the program ends)
        if (st != 2) return st == 1;
        if (may_bail) io = saved;          // ... so the interpreter runs it from the start
    }
#endif
    f3 stack[RXVM_STACK + 1];
    f3 locals[RXVM_LOCAL_POOL];
    f3 globals[RXVM_GLOBALS];
    uint32_t fr_pc[RXVM_FRAMES], fr_lb[RXVM_FRAMES], fr_sb[RXVM_FRAMES], fr_mk[RXVM_FRAMES], fr_nl[RXVM_FRAMES];
    uint32_t marks[RXVM_MARKS];
    const f3 zero = {0.0f, 0.0f, 0.0f};
    const uint32_t n_words = P.n_words, n_globals = P.n_globals, shade_locals = P.shade_locals;
    if (n_words == 0u) return true;                       // shade_index is None
    if (shade_locals > 32u || n_globals > RXVM_GLOBALS) return false;
    for (uint32_t i = 0; i < n_globals; ++i) globals[i] = (PERSIST && i < ps->n_globals) ? ps->globals[i] : zero;
    for (uint32_t i = 0; i < shade_locals; ++i) locals[i] = (PERSIST && i < ps->n_locals) ? ps->locals[i] : zero;
    const uint32_t* __restrict__ code = vm.code + P.code_off;
    uint32_t pc = P.entry, sp = 0, lb = 0, nl = shade_locals, nf = 0, nm = 0;
    bool have_ret = false;
    f3 ret = zero, t = zero;
    stack[0] = zero;

#define VM_NEED(n) if (sp < (uint32_t)(n)) return false
#define VM_ROOM() if (sp >= RXVM_STACK) return false
#define VM_PUSH_RAW(v) { stack[sp] = t; ++sp; t = (v); }
#define VM_POP_RAW(dst) { dst = t; --sp; t = stack[sp]; }
#define VM_UN(expr) { VM_NEED(1); const f3 a = t; t = expr; break; }
#define VM_BIN(expr) { VM_NEED(2); const f3 b = t, a = stack[sp - 1]; --sp; t = expr; break; }
#define VM_MAP1(fn) VM_UN((f3{fn(a.x), fn(a.y), fn(a.z)}))
#define VM_PUSHV(v) { VM_ROOM(); VM_PUSH_RAW(v); break; }
#define VM_POPTO(dst) { VM_NEED(1); VM_POP_RAW(dst); break; }
#define VM_B(x) ((x) ? f3{1.0f, 1.0f, 1.0f} : f3{0.0f, 0.0f, 0.0f})

    for (uint32_t budget = RXVM_BUDGET; budget != 0u; --budget) {
        if (pc >= n_words) return false;
        const uint32_t w = __ldg(code + pc++);
        const uint32_t a24 = w >> 8, op = w & 0xFFu;
        // nvcc lowers the 100-way switch below to a compare tree (7 levels, ~23 instructions per op); the five ops that
        // make up ~60 % of what shader programs execute are tested first, most frequent first
        if (op == RXVM_PUSH) {
            if (pc + 3u > n_words) return false;
            VM_ROOM();
            VM_PUSH_RAW((f3{__uint_as_float(__ldg(code + pc)), __uint_as_float(__ldg(code + pc + 1)), __uint_as_float(__ldg(code + pc + 2))}));
            pc += 3u;
            continue;
        }
        if (op == RXVM_LOAD_LOCAL) { if (a24 >= nl) return false; VM_ROOM(); VM_PUSH_RAW(locals[lb + a24]); continue; }
        if (op == RXVM_MUL) { VM_NEED(2); const f3 a = stack[sp - 1]; --sp; t = {a.x * t.x, a.y * t.y, a.z * t.z}; continue; }
        if (op == RXVM_ADD) { VM_NEED(2); const f3 a = stack[sp - 1]; --sp; t = {a.x + t.x, a.y + t.y, a.z + t.z}; continue; }
        if (op == RXVM_STORE_LOCAL) { if (a24 >= nl) return false; VM_NEED(1); VM_POP_RAW(locals[lb + a24]); continue; }
        switch (op) {
            case RXVM_LOAD_GLOBAL: if (a24 >= n_globals) return false; VM_PUSHV(globals[a24]);
            case RXVM_STORE_GLOBAL: if (a24 >= n_globals) return false; VM_POPTO(globals[a24]);
            case RXVM_SWAP: { VM_NEED(2); const f3 x = stack[sp - 1]; stack[sp - 1] = t; t = x; break; }
            case RXVM_GET_COMPONENTS: VM_NEED(1); t = vm_get_components(a24, t); break;
            case RXVM_SET_COMPONENTS: { VM_NEED(2); const f3 d = stack[sp - 1]; --sp; t = vm_set_components(a24, d, t); break; }
            case RXVM_CLEAR: if (sp) { --sp; t = stack[sp]; } break;
            case RXVM_DUP: if (sp) { VM_ROOM(); stack[sp] = t; ++sp; } break;
            case RXVM_FUNCTION_CALL: {    // execution.rs:186-223
                if (pc >= n_words || nf >= RXVM_FRAMES) return false;
                const uint32_t target = __ldg(code + pc++);
                const uint32_t arity = a24 & 0xFFu, total = a24 >> 8;
                const uint32_t nlb = lb + nl;
                if (total > 32u || nlb + total > RXVM_LOCAL_POOL || target >= n_words) return false;
                for (uint32_t i = 0; i < total; ++i) locals[nlb + i] = zero;
                for (uint32_t i = arity; i-- > 0u;)
                    if (sp) { f3 v; VM_POP_RAW(v); if (i < total) locals[nlb + i] = v; else return false; }
                fr_pc[nf] = pc; fr_lb[nf] = lb; fr_nl[nf] = nl; fr_sb[nf] = sp; fr_mk[nf] = nm; ++nf;
                lb = nlb; nl = total; pc = target;
                break;
            }
            case RXVM_RETURN:             // execution.rs:224-234
                if (sp) { VM_POP_RAW(ret); } else ret = zero;
                have_ret = true;
                // fall through: leave the function
            case RXVM_END: {
                if (nf == 0u) {
                    if (PERSIST) {   // (at the outermost frame lb is 0: locals[0 .. shade_locals) are shade()'s)
                        for (uint32_t i = 0; i < n_globals; ++i) ps->globals[i] = globals[i];
                        for (uint32_t i = 0; i < shade_locals; ++i) ps->locals[i] = locals[i];
                        ps->n_globals = n_globals; ps->n_locals = shade_locals;
                    }
                    return true;
                }
                --nf;
                const uint32_t base = fr_sb[nf];
                f3 r = zero;
                if (have_ret) { r = ret; have_ret = false; }
                else if (sp > base) { VM_POP_RAW(r); }
                if (sp > base) { sp = base; t = stack[sp]; }
                nm = fr_mk[nf]; lb = fr_lb[nf]; nl = fr_nl[nf]; pc = fr_pc[nf];
                VM_ROOM();
                VM_PUSH_RAW(r);
                break;
            }
            case RXVM_JZ: { VM_NEED(1); f3 v; VM_POP_RAW(v); if (v.x == 0.0f) { if (a24 > n_words) return false; pc = a24; } break; }
            case RXVM_JMP: if (a24 > n_words) return false; pc = a24; break;
            case RXVM_MARK: if (nm >= RXVM_MARKS) return false; marks[nm++] = sp; break;
            case RXVM_TRUNC: if (nm == 0u) return false; if (sp > marks[nm - 1]) { sp = marks[nm - 1]; t = stack[sp]; } break;
            case RXVM_UNMARK: if (nm == 0u) return false; --nm; break;
            case RXVM_PACK3: case RXVM_MIX: case RXVM_SMOOTHSTEP: case RXVM_CLAMP: {
                VM_NEED(3);
                const f3 c = t, b = stack[sp - 1], a = stack[sp - 2];
                sp -= 2;
                t = vm_tern(op, a, b, c);
                break;
            }
            case RXVM_ABS: case RXVM_FLOOR: case RXVM_FRACT: case RXVM_NOT: case RXVM_NEG:
            case RXVM_SIN: case RXVM_SIN1: case RXVM_COS1: case RXVM_SIN2: case RXVM_COS2: case RXVM_COS: case RXVM_TAN: case RXVM_ATAN:
            case RXVM_SQRT: case RXVM_LOG: case RXVM_ROUND: case RXVM_CEIL: case RXVM_RADIANS: case RXVM_DEGREES: case RXVM_LENGTH:
            case RXVM_LENGTH2: case RXVM_LENGTH3: case RXVM_NORMALIZE:
                VM_NEED(1); t = vm_un(op, t); break;
            case RXVM_PACK2: case RXVM_SUB: case RXVM_DOT: case RXVM_DOT2: case RXVM_DOT3: case RXVM_MIN: case RXVM_MAX: case RXVM_STEP:
            case RXVM_EQ: case RXVM_NE: case RXVM_LT: case RXVM_LE: case RXVM_GT: case RXVM_GE: case RXVM_AND: case RXVM_OR:
            case RXVM_ATAN2: case RXVM_POW: case RXVM_MOD: case RXVM_DIV: case RXVM_ROTATE2D: case RXVM_CROSS: {
                VM_NEED(2); const f3 a = stack[sp - 1]; --sp; t = vm_bin(op, a, t); break;
            }
            case RXVM_PRINT: { VM_NEED(1); --sp; t = stack[sp]; break; }
            case RXVM_UV: VM_PUSHV(io.uv)
            case RXVM_SET_UV: VM_POPTO(io.uv)
            case RXVM_NORMAL: VM_PUSHV(io.normal)
            case RXVM_SET_NORMAL: { VM_NEED(1); f3 v; VM_POP_RAW(v); io.normal = rx_normalize3(v); break; }   // .normalized(), execution.rs:583
            case RXVM_HITPOINT: VM_PUSHV(io.hitpoint)
            case RXVM_TIME: VM_PUSHV(io.time)
            case RXVM_COLOR: VM_PUSHV(io.color)
            case RXVM_SET_COLOR: VM_POPTO(io.color)
            case RXVM_ROUGHNESS: VM_PUSHV(io.roughness)
            case RXVM_SET_ROUGHNESS: VM_POPTO(io.roughness)
            case RXVM_METALLIC: VM_PUSHV(io.metallic)
            case RXVM_SET_METALLIC: VM_POPTO(io.metallic)
            case RXVM_EMISSIVE: VM_PUSHV(io.emissive)
            case RXVM_SET_EMISSIVE: VM_POPTO(io.emissive)
            case RXVM_OPACITY: VM_PUSHV(io.opacity)
            case RXVM_SET_OPACITY: VM_POPTO(io.opacity)
            case RXVM_BUMP: VM_PUSHV(io.bump)
            case RXVM_SET_BUMP: VM_POPTO(io.bump)
            case RXVM_SAMPLE: case RXVM_SAMPLE_NORMAL: {
                VM_NEED(2); const f3 a = stack[sp - 1]; --sp; t = vm_sample_op(vm, op == RXVM_SAMPLE_NORMAL, a, t); break;
            }
            case RXVM_PALETTE_INDEX: {    // execution.rs:735-742: nothing is pushed for a missing colour
                VM_NEED(1);
                f3 a; VM_POP_RAW(a);
                const uint32_t i = vm_as_index(a.x);
                if (i < vm.n_palette) {
                    const float4 c = __ldg(vm.palette + i);
                    if (c.x != 0.0f) { VM_ROOM(); VM_PUSH_RAW((f3{c.y, c.z, c.w})); }
                }
                break;
            }
            default: return false;        // If / For (not lowered), Alloc / Iterate / Save, unknown
        }
    }
    return false;  // op budget exhausted
#undef VM_NEED
#undef VM_ROOM
#undef VM_PUSH_RAW
#undef VM_POP_RAW
#undef VM_UN
#undef VM_BIN
#undef VM_MAP1
#undef VM_PUSHV
#undef VM_POPTO
#undef VM_B
}
__device__ __forceinline__ bool vm_run(const VmDev& vm, const DProgram& P, VmIO& io) { return vm_run_t<false>(vm, P, io, nullptr); }
#endif  // !(RXVM_JIT && RXVM_JIT_COMPLETE)
