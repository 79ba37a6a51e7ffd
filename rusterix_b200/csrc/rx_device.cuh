// rx_device.cuh -- device-side data layout and exact-arithmetic helpers shared by the kernels.
//
// Everything that decides coverage, depth or the alpha test is written with one rounding per
// reference operation (the library is compiled with -fmad=false; FMAs appear only as explicit
// __fmaf_rn where the reference has mul_add or where vek's Mat*Vec does).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "rxcuda.h"

#define RX_TILE_W 32             // screen tile of one raster CTA
#define RX_TILE_H 32
#define RX_TILE_THREADS 256       // 8 warps; warp w owns a 16x8 region, every thread 2x2 pixels of it (stride 8, 4)
#ifndef RX_DX
#define RX_DX 8            // a thread's 2x2 pixels: (x, x + RX_DX) x (y, y + 4).  8: at a given k the 32 lanes of a warp cover a compact
                           // 8x4 block (coherent coverage, textures and light culls).  1 (horizontally adjacent pairs, which the
                           // packed pair shade needs) was measured: teapot +50 %, dense 8K +23 %, map 4K +10 % -- a small triangle's
                           // footprint then spreads over twice the k iterations at half the lane utilisation
#endif
#define RX_REGION_W 16
#define RX_REGION_H 8
#define RX_CHUNK_TRIS 256          // triangles of one batch handled by one setup CTA
#define RX_LARGE_TILES 24          // a triangle covering more GPU tiles than this goes to the large list
#define RX_OWNER_NONE 0xFFFFFFFFu
#define RX_NEAR_PLANE 0.1f         // reference src/batch/batch3d.rs:563

// ---------------------------------------------------------------------------------------------
// HBM layout (see DESIGN.md "Data layout")
// ---------------------------------------------------------------------------------------------
struct DTex {            // one texture frame in the texel arena
    uint64_t offset;     // byte offset of RGBA8 data in the arena
    uint32_t width, height;
    uint32_t all_opaque; // every alpha == 255 -> the alpha test can be skipped
    uint32_t pad;
};
struct DTile {           // an animated tile = n_frames consecutive DTex
    uint32_t first, n_frames;
};

struct DBatch3 {         // static per 3D batch (uploaded by rxc_set_scene)
    uint32_t v_off, n_verts;
    uint32_t t_off, n_tris;
    uint32_t owner_base;         // 3 * t_off : slot/ordinal of triangle 0
    uint32_t cull_mode, repeat_mode;
    uint32_t source_kind, source_index;
    uint32_t source_pixel;       // RGBA packed little endian
    uint32_t has_normals;
    uint32_t chunk_first, n_chunks;
    uint32_t orphan_off, n_orphans; // vertices no triangle references (they still count for the bbox)
    int32_t chunk;               // index of the batch's chunk, -1 = none
    float ambient[3];
    int32_t program;             // absolute index into VmDev::programs of batch.shader, -1 = none
    float transform[16];
    float aabb_min[3], aabb_max[3]; // object-space AABB, NaN-ignoring min/max (batch3d.rs:494-507)
    uint32_t profile_id;         // batch.profile_id when RX_BF_HAS_PROFILE
    uint32_t bflags;             // RX_BF_*
};
#define RX_BF_HAS_PROFILE 1u
#define RX_BF_OPACITY 2u         // a chunk.batches3d_opacity batch (rasterizer.rs:1425-1690)
#define RX_BF_UNIT_W 4u          // every vertex has w == 1 and finite coordinates: the vertices lie inside the object AABB's hull

struct DBatch2 {         // static per 2D batch
    uint32_t v_off, n_verts;
    uint32_t t_off, n_tris;      // triangles, or line segments for line modes
    uint32_t mode, repeat_mode;
    uint32_t source_kind, source_index;
    uint32_t source_pixel;
    uint32_t receives_light;
    uint32_t rec_off;            // first record of this batch in the per-frame 2D record array
    int32_t chunk;               // index of the batch's chunk, -1 = none
    uint32_t n_recs;             // records per frame: triangles, or line segments (by mode)
    int32_t program;             // absolute index into VmDev::programs of batch.shader, -1 = none
    uint32_t pad[2];
};

struct DSector {         // one (BBox, occlusion) entry (chunk.rs:41, mini.rs:33)
    float minx, miny, maxx, maxy;
    float occlusion;
    float pad[3];
};
struct DChunkInfo {      // what shading reads of a Chunk; entry [n_chunks] holds the mapmini sectors
    uint32_t sector_off, n_sectors;
    int32_t origin_x, origin_y;
    int32_t size;
    uint32_t terrain_tex;    // DTex index of chunk.terrain_texture or 0xFFFFFFFF
    uint32_t pad[2];
};

struct DLight {          // rxc_light + the per-frame flicker factor (light.rs:656-672)
    uint32_t light_type, emitting, from_linedef;
    float range2;   // per frame: squared distance from which the light contributes nothing (0: never, +inf: no range)
    float px, py, pz, intensity;
    float cr, cg, cb, flicker_factor;
    float start_distance, end_distance, cone_angle, width;
    float dx, dy, dz, height;
    float nx, ny, nz, inv_range;  // inv_range = 1/(start - end), filled per frame (smoothstep / linear falloff)
};

// Rusteria VM residency (rxc_program / rxc_pattern of include/rxcuda.h)
struct DProgram {
    uint32_t code_off, n_words;  // words of this program inside VmDev::code
    uint32_t entry, shade_locals, n_globals, sets_opacity;
    uint32_t jit_index;          // which generated function runs this program in a JIT-compiled kernel, 0xFFFFFFFF = the interpreter
    uint32_t pad;
};
struct DPattern {
    uint32_t off;                // first Value (3 floats) inside VmDev::pattern_data
    uint32_t width, height, pad;
};
struct VmDev {
    const uint32_t* code;
    const DProgram* programs;
    const float* pattern_data;
    const DPattern* patterns;    // patterns, then normal patterns
    const float4* palette;       // (present, r, g, b)
    uint32_t n_programs, n_patterns, n_patterns_normal, n_palette;
};

struct DChunk {          // work item of the setup kernel: <= RX_CHUNK_TRIS triangles of one batch
    uint32_t batch, first_tri, n_tris, pad;
};

// per (frame, 3D batch) state, written by k_frame_setup / k_tri_setup / k_batch_finalize
struct __align__(16) DFrameBatch {
    float view_model[16];
    uint32_t bb_minx, bb_maxx, bb_miny, bb_maxy; // order-preserving uint keys of the float bbox
    int32_t sc_x0, sc_x1, sc_y0, sc_y1;          // pixel scissor equivalent to the per-tile bbox reject
    uint32_t rejected;                           // frustum-AABB early-out or nothing-to-draw
    uint32_t tex;                                // DTex index for this frame (animation frame applied)
    uint32_t alpha_test;                         // texture has non-opaque texels
    uint32_t n_new_tris;                         // near-clip output triangles of this batch
    // what the deferred shade reads per owner, packed into two 16 B loads
    uint32_t sd_tex_word;                        // texel arena offset / 4 of this frame's texture
    uint32_t sd_wh;                              // width | height << 16   (textures <= 65535 per side)
    uint32_t sd_flags;                           // RX_SD_* bits
    uint32_t sd_pixel;                           // constant texel for Pixel / other sources
    float sd_ambient[3];                         // batch.ambient_color
    int32_t sd_chunk;                            // chunk index (occlusion, terrain), -1 = none
    uint32_t sd_profile;                         // profile id (valid with RX_SD_HAS_PROFILE)
    int32_t sd_program;                          // VM program of the batch (RX_SD_SHADER)
    uint32_t sd_pad[2];
};
#define RX_SD_TEXTURED 1u
#define RX_SD_REPEAT_X 2u
#define RX_SD_REPEAT_Y 4u
#define RX_SD_NORMALS 8u
#define RX_SD_TERRAIN 16u      // texel = chunk.sample_terrain_texture(world.xz)
#define RX_SD_HAS_PROFILE 32u
#define RX_SD_OPACITY 64u
#define RX_SD_SHADER 128u      // a Rusteria VM program shades the fragment (rasterizer.rs:1226-1293)
#define RX_SD_VM_OPACITY 256u  // ... and it writes opacity: the alpha test runs the program

struct DFrameBatch2 {
    uint32_t tex;          // DTex index or 0xFFFFFFFF (transparent texel)
    uint32_t lit;          // lighting branch taken (rasterizer.rs:799-802)
    uint32_t terrain;      // source is PixelSource::Terrain: `tex` is the chunk's terrain texture
    int32_t program;       // VM program of the batch, -1 = none
};

// Visibility record: everything the per-pixel coverage/depth test reads (96 B, 16 B aligned)
struct __align__(16) TriVis {
    float ax, ay, bx, by;      // unswapped projected v0, v1 (xy)
    float cx, cy, rarea, spare;// v2 ; rarea = RN(1/area) for the exact fast division (see rx_div_by)
    float area, iz0, iz1, iz2; // area = ac.x*ab.y - ac.y*ab.x ; 1/z per vertex
    float ea[3], eb[3], ec[3]; // edge equations of the (possibly swapped) triangle
    uint32_t bbx;              // x0 | x1<<16  (x1 exclusive), after scissor
    uint32_t bby;              // y0 | y1<<16
    uint32_t meta;             // batch index | opacity<<29 | fastdiv_ok<<30 | alpha_test<<31
};
#define RX_META_ALPHA 0x80000000u
#define RX_META_FASTDIV 0x40000000u
#define RX_META_OPACITY 0x20000000u
#define RX_META_BATCH 0x1FFFFFFFu
static_assert(sizeof(TriVis) == 96, "TriVis must be 96 bytes");

// Shading record: attributes only the alpha test and the final shade read (80 B)
struct __align__(16) TriShade {
    float uw0, vw0, uw1, vw1;  // uv_i / w_i
    float uw2, vw2, rw0, rw1;  // 1 / w_i
    float rw2, n0x, n0y, n0z;
    float n1x, n1y, n1z, n2x;
    float n2y, n2z, pad0, pad1;
};
static_assert(sizeof(TriShade) == 80, "TriShade must be 80 bytes");

// 2D triangle / line record (per frame, submission order)
struct __align__(16) Tri2D {
    float ax, ay, bx, by;
    float cx, cy, u0, v0;
    float u1, v1, u2, v2;
    float ea[3], eb[3], ec[3];
    uint32_t bbx, bby;         // scissored pixel bbox; empty when rejected
    uint32_t batch;            // 2D batch index
    uint32_t kind;             // 0 triangle; 1 line segment, whose integer end points x0,y0,x1,y1
                               // (`p as isize`, rasterizer.rs:1785-1788) are the bits of ax,ay,bx,by
    uint32_t pad[3];
};
static_assert(sizeof(Tri2D) == 112, "Tri2D must be 112 bytes");

struct DClip {             // one near-clipped source triangle (compact list)
    uint32_t tri;          // global original triangle index
    uint32_t chunk;        // setup chunk it came from
    uint32_t local_off;    // exclusive prefix of new-triangle counts inside the chunk
    uint32_t batch;
};

// per-frame scalar state
struct DFrame {
    float view[16], proj[16], inv_view[16], inv_proj[16];
    float mat2d[9];
    uint32_t has_mat2d;
    float cam[3];
    float width_f, height_f;
    int32_t width, height;        // full frame
    int32_t band_y0, band_y1;     // rows rendered
    int32_t band_x0, band_x1;     // columns rendered (the output buffer holds the rectangle, pitch band_x1 - band_x0)
    int32_t tiles_x, tiles_y;     // GPU tiles covering the rectangle
    uint32_t tile_size;           // API tile size (scissor computation)
    uint32_t sample_mode;
    uint32_t has_bg_color, bg_color;
    uint32_t bg_shader;
    float grid_size, grid_subdiv, grid_off[2];
    uint32_t has_ambient;
    float ambient[4];
    uint32_t hash_anim;
    uint32_t d2_active, d3_active, ignore_bg_shader, preserve_transparency;
    uint32_t matvec_mode;
    float trans2d[2], scale2d;
    uint64_t animation_frame;
    // screen_to_world (rasterizer.rs:1707-1727) folded by the host in double precision:
    // (hx,hy,hz,hw) = s2w * (px+.5, py+.5, z, 1), world = (hx,hy,hz)/hw.  Row-major 4x4.
    float s2w[16];
    float time;                   // Rasterizer.time (VM `time`)
    // render graph results (rxc_frame: sun, Sky node, brush preview)
    float sun_radiance;           // max(day_factor, 0) when the sun lights the frame, else 0 (rasterizer.rs:1342-1347)
    float sun_l[3];               // (-sun_dir).normalized()
    uint32_t has_sky, has_brush;
    float sky[6][4];
    float brush_pos[3], brush_radius, brush_falloff;
    uint32_t preprojected;        // rxc_rasterize_projected: the host's own Scene::project results replace the device front end
};

// per-frame counters (zeroed by k_frame_setup)
struct DCounters {
    uint32_t n_clip;        // entries in the clip list
    uint32_t n_new_slots;   // emitted near-clip triangles (compact slot list)
    uint32_t n_large;       // entries in the large-triangle list
    uint32_t list_cursor;   // next free entry of the tile-list arena
    uint32_t overflow;      // bit0 tile-list arena, bit1 large list, bit2 clip list
    uint32_t n_visible;     // statistics
    uint32_t list_cursor2;  // next free entry of the 2D tile-list arena
    uint32_t pad;
};

// ---------------------------------------------------------------------------------------------
// exact helpers
// ---------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t rx_float_key(float f) {  // monotone float -> uint
    uint32_t b;
#ifdef __CUDA_ARCH__
    b = __float_as_uint(f);
#else
    memcpy(&b, &f, 4);
#endif
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__host__ __device__ __forceinline__ float rx_key_float(uint32_t k) {
    uint32_t b = (k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k;
#ifdef __CUDA_ARCH__
    return __uint_as_float(b);
#else
    float f;
    memcpy(&f, &b, 4);
    return f;
#endif
}

#ifdef __CUDACC__
struct f4 { float x, y, z, w; };
struct f3 { float x, y, z; };

__device__ __forceinline__ f4 rx_matvec4(const float* __restrict__ m, f4 v, uint32_t mode) {
    f4 r;
    if (mode == RXC_MATVEC_PLAIN_ROWS) {
        r.x = ((m[0] * v.x + m[4] * v.y) + m[8] * v.z) + m[12] * v.w;
        r.y = ((m[1] * v.x + m[5] * v.y) + m[9] * v.z) + m[13] * v.w;
        r.z = ((m[2] * v.x + m[6] * v.y) + m[10] * v.z) + m[14] * v.w;
        r.w = ((m[3] * v.x + m[7] * v.y) + m[11] * v.z) + m[15] * v.w;
    } else {
        r.x = __fmaf_rn(m[12], v.w, __fmaf_rn(m[8], v.z, __fmaf_rn(m[4], v.y, m[0] * v.x)));
        r.y = __fmaf_rn(m[13], v.w, __fmaf_rn(m[9], v.z, __fmaf_rn(m[5], v.y, m[1] * v.x)));
        r.z = __fmaf_rn(m[14], v.w, __fmaf_rn(m[10], v.z, __fmaf_rn(m[6], v.y, m[2] * v.x)));
        r.w = __fmaf_rn(m[15], v.w, __fmaf_rn(m[11], v.z, __fmaf_rn(m[7], v.y, m[3] * v.x)));
    }
    return r;
}
__device__ __forceinline__ void rx_matmat4(const float* A, const float* B, float* R, uint32_t mode) {
    for (int c = 0; c < 4; ++c) {
        f4 col = {B[c * 4 + 0], B[c * 4 + 1], B[c * 4 + 2], B[c * 4 + 3]};
        f4 r = rx_matvec4(A, col, mode);
        R[c * 4 + 0] = r.x; R[c * 4 + 1] = r.y; R[c * 4 + 2] = r.z; R[c * 4 + 3] = r.w;
    }
}
__device__ __forceinline__ f3 rx_matvec3(const float* m, f3 v, uint32_t mode) {
    f3 r;
    if (mode == RXC_MATVEC_PLAIN_ROWS) {
        r.x = (m[0] * v.x + m[3] * v.y) + m[6] * v.z;
        r.y = (m[1] * v.x + m[4] * v.y) + m[7] * v.z;
        r.z = (m[2] * v.x + m[5] * v.y) + m[8] * v.z;
    } else {
        r.x = __fmaf_rn(m[6], v.z, __fmaf_rn(m[3], v.y, m[0] * v.x));
        r.y = __fmaf_rn(m[7], v.z, __fmaf_rn(m[4], v.y, m[1] * v.x));
        r.z = __fmaf_rn(m[8], v.z, __fmaf_rn(m[5], v.y, m[2] * v.x));
    }
    return r;
}

// batch3d.rs:691-700
__device__ __forceinline__ f4 rx_project(const float* proj, f4 v, float vw, float vh, uint32_t mode) {
    f4 r = rx_matvec4(proj, v, mode);
    float w = r.w;
    f4 o;
    o.x = ((r.x / w) * 0.5f + 0.5f) * vw;
    o.y = ((-r.y / w) * 0.5f + 0.5f) * vh;
    o.z = r.z / w;
    o.w = w;
    return o;
}

// a / b, correctly rounded, given rb = RN(1/b) (Markstein's residual correction: two fused
// corrections of q = a*rb).  Exact only while no intermediate under/overflows: callers guarantee
// |b| in [2^-20, 2^44] and a == 0 or |a| in [2^-80, 2^62], so the quotient stays normal (see make_tri);
// checked against div.rn on
// the device by rxc_selftest_div.
__device__ __forceinline__ float rx_div_by(float a, float b, float rb) {
    float q = a * rb;
    float r = __fmaf_rn(-b, q, a);
    q = __fmaf_rn(r, rb, q);
    r = __fmaf_rn(-b, q, a);
    return __fmaf_rn(r, rb, q);
}

// ---- packed fp32 (sm_100 FFMA2 / FMUL2 / FADD2: one issue slot for two lanes of fp32) -------------------------------
// The coverage / depth arithmetic is one rounded IEEE operation per reference operation, never contracted.  For the
// scalar code -fmad=false guarantees that.  ptxas does NOT extend the guarantee to packed operations: a mul.rn.f32x2
// followed by an add.rn.f32x2 (or an fma.rn.f32x2 it recognises as one of those) is contracted into FFMA2 whatever
// --fmad says (checked in SASS).  So every packed PRODUCT that feeds a sum is issued as a true fused multiply-add with
// the addend -0.0 held in a register the compiler cannot see through (a kernel argument): RN(a*b + (-0)) == RN(a*b)
// bit for bit (a zero product keeps its sign: (+0) + (-0) = +0, (-0) + (-0) = -0), and an FFMA2 cannot be contracted
// with the FADD2 that consumes it.  Sums and differences are fma(a, +-1, b), exact by construction.
__device__ __forceinline__ float2 rx_fma2(float2 a, float2 b, float2 c) {
    float2 d;
    asm("{\n .reg .b64 ra, rb, rc, rd;\n mov.b64 ra, {%2, %3};\n mov.b64 rb, {%4, %5};\n mov.b64 rc, {%6, %7};\n fma.rn.f32x2 rd, ra, rb, rc;\n mov.b64 {%0, %1}, rd;\n}"
        : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return d;
}
__device__ __forceinline__ float2 rx_bc2(float s) { return make_float2(s, s); }
__device__ __forceinline__ float2 rx_mul2(float2 a, float2 b, float nz) { return rx_fma2(a, b, rx_bc2(nz)); }       // RN(a*b), nz = -0.0f (opaque)
__device__ __forceinline__ float2 rx_add2(float2 a, float2 b) { return rx_fma2(a, rx_bc2(1.0f), b); }                // RN(a+b)
__device__ __forceinline__ float2 rx_sub2(float2 a, float2 b) { return rx_fma2(b, rx_bc2(-1.0f), a); }               // RN(a-b)
__device__ __forceinline__ float rx_mul1(float a, float b, float nz) { return __fmaf_rn(a, b, nz); }                 // scalar product that may feed a packed sum
// rx_div_by on two numerators
__device__ __forceinline__ float2 rx_div_by2(float2 a, float b, float rb, float nz) {
    float2 q = rx_mul2(a, rx_bc2(rb), nz);
    float2 r = rx_fma2(q, rx_bc2(-b), a);
    q = rx_fma2(r, rx_bc2(rb), q);
    r = rx_fma2(q, rx_bc2(-b), a);
    return rx_fma2(r, rx_bc2(rb), q);
}

__device__ __forceinline__ float rx_clamp(float x, float lo, float hi) {  // Rust f32::clamp (NaN stays)
    if (x < lo) return lo;
    if (x > hi) return hi;
    return x;
}
// `as usize`-style saturating conversions clamped to [0, limit]
__device__ __forceinline__ int rx_sat_int(float x, int limit) {
    if (!(x == x)) return 0;
    if (x <= 0.0f) return 0;
    if (x >= (float)limit) return limit;
    return (int)x;
}
// `x.round() as usize` clamped to [0, limit] for x >= 0 or NaN (texture.rs:311-312; u and v are wrapped or
// clamped to [0, 1] before): round half away from zero == floor + (frac >= 0.5); the saturating
// cvt.rzi maps NaN to 0 like `as usize`.
__device__ __forceinline__ int rx_round_index(float x, int limit) {
    float r = floorf(x);
    if (x - r >= 0.5f) r += 1.0f;
    return min(max(__float2int_rz(r), 0), limit);
}
__device__ __forceinline__ uint32_t rx_as_u32(float x) {
    if (!(x == x)) return 0u;
    if (x <= 0.0f) return 0u;
    if (x >= 4294967296.0f) return 0xFFFFFFFFu;
    return (uint32_t)x;
}
__device__ __forceinline__ uint32_t rx_as_u8(float x) {
    if (!(x == x)) return 0u;
    if (x <= 0.0f) return 0u;
    if (x >= 255.0f) return 255u;
    return (uint32_t)x;
}
__device__ __forceinline__ uint32_t rx_f32_to_u8_saturated(float x) {  // lib.rs:65-68
    float y = __fmaf_rn(fminf(fmaxf(x, 0.0f), 1.0f), 255.0f, 0.5f);
    return (uint32_t)__float2int_rz(y);  // y is in [0.5, 255.5]; a NaN x became 0 in fmaxf
}

__device__ __forceinline__ float rx_dot3(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ f3 rx_sub3(f3 a, f3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ f3 rx_add3(f3 a, f3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ f3 rx_mul3(f3 a, f3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
__device__ __forceinline__ f3 rx_scale3(f3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
__device__ __forceinline__ f3 rx_normalize3(f3 a) {
    float m = sqrtf(rx_dot3(a, a));
    return {a.x / m, a.y / m, a.z / m};
}

// texture.rs:203-232 wrap/clamp of one coordinate
__device__ __forceinline__ float rx_wrap(float u, bool repeat) {
    return repeat ? (u - floorf(u)) : rx_clamp(u, 0.0f, 1.0f);
}
// texture.rs:203-232, 307-323, 414-460.  Returns RGBA packed little endian.
__device__ __forceinline__ uint32_t rx_sample_tex(const uint32_t* __restrict__ tex, int W, int H, float u, float v,
                                                  uint32_t sample_mode, bool repeat_x, bool repeat_y) {
    u = rx_wrap(u, repeat_x);
    v = rx_wrap(v, repeat_y);
    if (sample_mode == RXC_SAMPLE_NEAREST) {
        int tx = rx_round_index(u * ((float)W - 1.0f), W - 1);
        int ty = rx_round_index(v * ((float)H - 1.0f), H - 1);
        return __ldg(tex + ty * W + tx);
    }
    float x = u * ((float)W - 1.0f);
    float y = v * ((float)H - 1.0f);
    float fx = floorf(x), fy = floorf(y);
    int x0 = rx_sat_int(fx, W - 1);
    int y0 = rx_sat_int(fy, H - 1);
    int x1 = min(x0 + 1, W - 1);
    int y1 = min(y0 + 1, H - 1);
    float dx = x - fx, dy = y - fy;
    uint32_t c00 = __ldg(tex + y0 * W + x0), c10 = __ldg(tex + y0 * W + x1);
    uint32_t c01 = __ldg(tex + y1 * W + x0), c11 = __ldg(tex + y1 * W + x1);
    uint32_t out = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float v00 = (float)((c00 >> (8 * i)) & 0xFF), v10 = (float)((c10 >> (8 * i)) & 0xFF);
        float v01 = (float)((c01 >> (8 * i)) & 0xFF), v11 = (float)((c11 >> (8 * i)) & 0xFF);
        float a = v00 + dx * (v10 - v00);
        float b = v01 + dx * (v11 - v01);
        float r = a + dy * (b - a);
        out |= rx_as_u8(roundf(r)) << (8 * i);
    }
    return out;
}
__device__ __forceinline__ uint32_t rx_sample(const uint8_t* __restrict__ arena, const DTex& t, float u, float v,
                                              uint32_t sample_mode, uint32_t repeat_mode) {
    return rx_sample_tex(reinterpret_cast<const uint32_t*>(arena + t.offset), (int)t.width, (int)t.height, u, v, sample_mode,
                         repeat_mode == RXC_REPEAT_REPEAT_XY || repeat_mode == RXC_REPEAT_REPEAT_X,
                         repeat_mode == RXC_REPEAT_REPEAT_XY || repeat_mode == RXC_REPEAT_REPEAT_Y);
}

__device__ __forceinline__ float rx_smoothstep(float e0, float e1, float x) {  // light.rs:674-677
    float t = rx_clamp((x - e0) / (e1 - e0), 0.0f, 1.0f);
    return t * t * (3.0f - 2.0f * t);
}

// light.rs:491-653 CompiledLight::color_at.  flicker_factor is precomputed per frame.
__device__ __forceinline__ bool rx_light_color_at(const DLight& l, f3 point, bool d2, f3* out) {
    if (!l.emitting) return false;
    const f3 lp = {l.px, l.py, l.pz};
    const f3 col = {l.cr, l.cg, l.cb};
    switch (l.light_type) {
        case RXC_LIGHT_POINT: {
            f3 d = rx_sub3(point, lp);
            float distance = sqrtf(rx_dot3(d, d));
            if (distance >= l.end_distance) return false;
            float inten = l.intensity;
            if (!(distance <= l.start_distance)) inten = l.intensity * rx_smoothstep(l.end_distance, l.start_distance, distance);
            *out = {col.x * inten * l.flicker_factor, col.y * inten * l.flicker_factor, col.z * inten * l.flicker_factor};
            return true;
        }
        case RXC_LIGHT_AMBIENT:
        case RXC_LIGHT_AMBIENT_DAYLIGHT:
            *out = {col.x * l.intensity * l.flicker_factor, col.y * l.intensity * l.flicker_factor, col.z * l.intensity * l.flicker_factor};
            return true;
        case RXC_LIGHT_SPOT: {
            f3 d = rx_sub3(point, lp);
            float distance = sqrtf(rx_dot3(d, d));
            if (distance >= l.end_distance) return false;
            float att = (distance <= l.start_distance) ? 1.0f : 1.0f - ((distance - l.start_distance) / (l.end_distance - l.start_distance));
            f3 dir = {d.x / distance, d.y / distance, d.z / distance};
            f3 ld = {l.dx, l.dy, l.dz};
            float angle = acosf(rx_dot3(ld, dir));
            if (angle > l.cone_angle) return false;
            float inten = l.intensity * att;
            *out = {col.x * inten * l.flicker_factor, col.y * inten * l.flicker_factor, col.z * inten * l.flicker_factor};
            return true;
        }
        case RXC_LIGHT_AREA: {
            f3 tp = rx_sub3(point, lp);
            float distance = sqrtf(rx_dot3(tp, tp));
            if (distance >= l.end_distance) return false;
            if (distance < 0.1f) { *out = col; return true; }
            float datt = (distance <= l.start_distance) ? 1.0f : rx_smoothstep(l.end_distance, l.start_distance, distance);
            float area = l.width * l.height;
            f3 dir = {tp.x / distance, tp.y / distance, tp.z / distance};
            float att;
            if (l.from_linedef) {
                att = datt * area * l.intensity;
            } else if (d2) {
                float dxn = fabsf(tp.x / (l.width * 0.5f));
                float dyn = fabsf(tp.y / (l.height * 0.5f));
                float ax = fmaxf(1.0f - dxn, 0.0f);
                float ay = fmaxf(1.0f - dyn, 0.0f);
                att = ax * ay * datt * l.intensity;
            } else {
                f3 n = {l.nx, l.ny, l.nz};
                float aatt = fmaxf(rx_dot3(n, dir), 0.0f);
                att = aatt * datt * area * l.intensity;
            }
            *out = {col.x * att, col.y * att, col.z * att};
            return true;
        }
        default: {  // Daylight
            f3 tp = rx_sub3(point, lp);
            float distance = sqrtf(rx_dot3(tp, tp));
            if (distance >= l.end_distance) return false;
            f3 dir = {tp.x / distance, tp.y / distance, tp.z / distance};
            f3 n = {l.nx, l.ny, l.nz};
            float aatt = fmaxf(rx_dot3(n, dir), 0.0f);
            float datt = (distance <= l.start_distance) ? 1.0f : rx_smoothstep(l.end_distance, l.start_distance, distance);
            float att = aatt * datt * l.intensity;
            *out = {col.x * att, col.y * att, col.z * att};
            return true;
        }
    }
}
#endif  // __CUDACC__
