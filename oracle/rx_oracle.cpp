// rx_oracle.cpp -- CPU ORACLE (test infrastructure, NOT the product).
//
// A restatement, in plain C++17, of the reference algorithm of markusmoenig/Rusterix's CPU tile
// rasterizer `Rasterizer::setup(..).rasterize(..)` (crate rusterix 0.2.8, Rust).  Every function
// cites the reference file:line it follows.  It is deliberately naive: it keeps the reference's
// structure (project every batch, split the screen into `tile_size` tiles, every tile walks every
// triangle of every bbox-overlapping batch in submission order, shades every z-passing fragment,
// serial compose), because its only jobs are (1) to be the checker the CUDA path is compared with
// and (2) to be the "port" CPU baseline timed next to the GPU numbers.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
// this library.  The product (rusterix_b200/, include/rxcuda.h) never links or calls it.
//
// PARITY STATUS.  Rasterizer proper: UNPINNED.  The reference is pure Rust and cannot be built in this image (no
// cargo/rustc), it ships no tests, golden images or fixtures for the rasterize path, and the
// rounding of vek 0.17.2's Mat*Vec / Mat*Mat (Cargo.lock:3826) cannot be verified offline.  That part is pinned only
// by hand-derived known-answer tests (tests/test_oracle_kat.py) and by the source text it restates; the Mat*Vec
// convention is selectable (frame.matvec_mode).
// Rusteria VM (the Execution::execute restatement below): PINNED by output of the reference itself.  The reference
// repository ships ten images its own VM rendered together with the programs that rendered them
// (rusteria/examples/*.png by the rsia CLI; rusteria/embedded/*.png by make_textures.rusteria); this interpreter
// reproduces eight of them bit for bit and the two sin-hashed ones up to the platform libm
// (tests/test_rusteria_golden.py, fixtures under tests/golden/rusteria/).
//
// Build: g++ -O2 -std=c++17 -ffp-contract=off -fno-fast-math -mfma -shared -fPIC (oracle/Makefile).
// -ffp-contract=off matters: Rust never fuses a*b+c, GCC's default would.

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <optional>
#include <thread>
#include <vector>

#include "../include/rxcuda.h"

namespace {

// ---------------------------------------------------------------------------------------------
// Rust scalar semantics
// ---------------------------------------------------------------------------------------------
inline float rmin(float a, float b) { return std::fmin(a, b); }  // f32::min ignores NaN
inline float rmax(float a, float b) { return std::fmax(a, b); }  // f32::max ignores NaN
inline float rclamp(float x, float lo, float hi) {               // f32::clamp keeps NaN
    if (x < lo) return lo;
    if (x > hi) return hi;
    return x;
}
inline size_t as_usize(float x) {  // `as usize`: saturating, NaN -> 0, truncation
    if (!(x == x)) return 0;
    if (x <= 0.0f) return 0;
    if (x >= 18446744073709551616.0f) return std::numeric_limits<size_t>::max();
    return (size_t)x;
}
inline int64_t as_isize(float x) {
    if (!(x == x)) return 0;
    if (x >= 9223372036854775808.0f) return std::numeric_limits<int64_t>::max();
    if (x <= -9223372036854775808.0f) return std::numeric_limits<int64_t>::min();
    return (int64_t)x;
}
inline int32_t as_i32(float x) {
    if (!(x == x)) return 0;
    if (x >= 2147483648.0f) return std::numeric_limits<int32_t>::max();
    if (x <= -2147483648.0f) return std::numeric_limits<int32_t>::min();
    return (int32_t)x;
}
inline uint32_t as_u32(float x) {
    if (!(x == x)) return 0;
    if (x <= 0.0f) return 0;
    if (x >= 4294967296.0f) return 0xFFFFFFFFu;
    return (uint32_t)x;
}
inline uint8_t as_u8(float x) {
    if (!(x == x)) return 0;
    if (x <= 0.0f) return 0;
    if (x >= 255.0f) return 255;
    return (uint8_t)x;
}

struct V2 { float x, y; };
struct V3 { float x, y, z; };
struct V4 { float x, y, z, w; };

inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator*(V3 a, V3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
inline V3 operator*(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline V3 operator/(V3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }
inline V3 operator-(V3 a) { return {-a.x, -a.y, -a.z}; }
// vek: dot = (a*b).sum(), magnitude = sqrt(dot(self,self)), normalized = self / magnitude
inline float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline float magnitude(V3 a) { return std::sqrt(dot(a, a)); }
inline V3 normalized(V3 a) { return a / magnitude(a); }

// vek Mat4 (column-major storage).  Op order of Mat*Vec is unpinned; both candidate
// conventions are implemented and selected by frame.matvec_mode.
struct M4 { float m[16]; };  // m[c*4+r]
inline V4 mat4_mul_vec4(const M4& M, V4 v, uint32_t mode) {
    const float* m = M.m;
    if (mode == RXC_MATVEC_PLAIN_ROWS) {
        V4 r;
        r.x = ((m[0] * v.x + m[4] * v.y) + m[8] * v.z) + m[12] * v.w;
        r.y = ((m[1] * v.x + m[5] * v.y) + m[9] * v.z) + m[13] * v.w;
        r.z = ((m[2] * v.x + m[6] * v.y) + m[10] * v.z) + m[14] * v.w;
        r.w = ((m[3] * v.x + m[7] * v.y) + m[11] * v.z) + m[15] * v.w;
        return r;
    }
    float r[4];
    for (int i = 0; i < 4; ++i) {
        float acc = m[i] * v.x;
        acc = std::fmaf(m[4 + i], v.y, acc);
        acc = std::fmaf(m[8 + i], v.z, acc);
        acc = std::fmaf(m[12 + i], v.w, acc);
        r[i] = acc;
    }
    return {r[0], r[1], r[2], r[3]};
}
inline M4 mat4_mul_mat4(const M4& A, const M4& B, uint32_t mode) {  // columns of B through A
    M4 R;
    for (int c = 0; c < 4; ++c) {
        V4 col = {B.m[c * 4 + 0], B.m[c * 4 + 1], B.m[c * 4 + 2], B.m[c * 4 + 3]};
        V4 r = mat4_mul_vec4(A, col, mode);
        R.m[c * 4 + 0] = r.x; R.m[c * 4 + 1] = r.y; R.m[c * 4 + 2] = r.z; R.m[c * 4 + 3] = r.w;
    }
    return R;
}
inline V3 mat3_mul_vec3(const float* m, V3 v, uint32_t mode) {  // m[c*3+r]
    if (mode == RXC_MATVEC_PLAIN_ROWS) {
        return {(m[0] * v.x + m[3] * v.y) + m[6] * v.z, (m[1] * v.x + m[4] * v.y) + m[7] * v.z,
                (m[2] * v.x + m[5] * v.y) + m[8] * v.z};
    }
    float r[3];
    for (int i = 0; i < 3; ++i) {
        float acc = m[i] * v.x;
        acc = std::fmaf(m[3 + i], v.y, acc);
        acc = std::fmaf(m[6 + i], v.z, acc);
        r[i] = acc;
    }
    return {r[0], r[1], r[2]};
}

// ---------------------------------------------------------------------------------------------
// src/lib.rs:50-79  pixel conversions
// ---------------------------------------------------------------------------------------------
const float INV_255 = 1.0f / 255.0f;
inline V4 pixel_to_vec4(const uint8_t* p) {  // src/lib.rs:55-62
    return {(float)p[0] * INV_255, (float)p[1] * INV_255, (float)p[2] * INV_255, (float)p[3] * INV_255};
}
inline uint8_t f32_to_u8_saturated(float x) {  // src/lib.rs:65-68 (true FMA, truncation)
    float y = std::fmaf(rmin(rmax(x, 0.0f), 1.0f), 255.0f, 0.5f);
    return (uint8_t)as_i32(y);
}
inline void vec4_to_pixel(V4 v, uint8_t* out) {  // src/lib.rs:72-79
    out[0] = f32_to_u8_saturated(v.x);
    out[1] = f32_to_u8_saturated(v.y);
    out[2] = f32_to_u8_saturated(v.z);
    out[3] = f32_to_u8_saturated(v.w);
}

// src/rasterizer.rs:19-33
inline float srgb_to_linear_fast(float x) {
    float x2 = x * x;
    return (0.6975f * x2 + 0.3025f) * x;
}
inline float linear_to_srgb_fast(float x) {
    float s = std::sqrt(x);
    return 1.055f * s - 0.055f * s * s;
}

// ---------------------------------------------------------------------------------------------
// src/edge.rs:1-36
// ---------------------------------------------------------------------------------------------
struct Edges {
    float a[3], b[3], c[3];
    bool visible;
};
inline Edges edges_new(const float v0[3][2], const float v1[3][2], bool visible) {  // src/edge.rs:12-25
    Edges e;
    for (int i = 0; i < 3; ++i) {
        e.a[i] = v1[i][1] - v0[i][1];
        e.b[i] = v0[i][0] - v1[i][0];
        e.c[i] = v1[i][0] * v0[i][1] - v1[i][1] * v0[i][0];
    }
    e.visible = visible;
    return e;
}
inline bool edges_evaluate(const Edges& e, float px, float py) {  // src/edge.rs:28-36
    for (int i = 0; i < 3; ++i) {
        float result = e.a[i] * px + e.b[i] * py + e.c[i];
        if (result < 0.0f) return false;
    }
    return true;
}

struct Rect { float x, y, width, height; };  // src/rect.rs:5-10

// ---------------------------------------------------------------------------------------------
// src/texture.rs:203-232, 307-323, 414-460
// ---------------------------------------------------------------------------------------------
inline void sample_nearest(const rxc_texture& t, float u, float v, uint8_t out[4]) {  // :307-323
    size_t tx = as_usize(std::round(u * ((float)t.width - 1.0f)));
    size_t ty = as_usize(std::round(v * ((float)t.height - 1.0f)));
    tx = std::min(tx, (size_t)t.width - 1);
    ty = std::min(ty, (size_t)t.height - 1);
    size_t idx = (ty * t.width + tx) * 4;
    std::memcpy(out, t.data + idx, 4);
}
inline void sample_linear(const rxc_texture& t, float u, float v, uint8_t out[4]) {  // :414-460
    float x = u * ((float)t.width - 1.0f);
    float y = v * ((float)t.height - 1.0f);
    size_t x0 = as_usize(std::floor(x));
    size_t x1 = std::min(x0 + 1, (size_t)t.width - 1);
    size_t y0 = as_usize(std::floor(y));
    size_t y1 = std::min(y0 + 1, (size_t)t.height - 1);
    // the reference would panic on an out-of-range x0/y0; u,v in [0,1] keeps them in range
    x0 = std::min(x0, (size_t)t.width - 1);
    y0 = std::min(y0, (size_t)t.height - 1);
    float dx = x - std::floor(x);
    float dy = y - std::floor(y);
    const uint8_t* c00 = t.data + (y0 * t.width + x0) * 4;
    const uint8_t* c10 = t.data + (y0 * t.width + x1) * 4;
    const uint8_t* c01 = t.data + (y1 * t.width + x0) * 4;
    const uint8_t* c11 = t.data + (y1 * t.width + x1) * 4;
    for (int i = 0; i < 4; ++i) {
        float v00 = c00[i], v10 = c10[i], v01 = c01[i], v11 = c11[i];
        float a = v00 + dx * (v10 - v00);
        float b = v01 + dx * (v11 - v01);
        float r = a + dy * (b - a);
        out[i] = as_u8(std::round(r));
    }
}
inline void texture_sample(const rxc_texture& t, float u, float v, uint32_t sample_mode, uint32_t repeat_mode,
                           uint8_t out[4]) {  // :203-232
    switch (repeat_mode) {
        case RXC_REPEAT_CLAMP_XY: u = rclamp(u, 0.0f, 1.0f); v = rclamp(v, 0.0f, 1.0f); break;
        case RXC_REPEAT_REPEAT_XY: u = u - std::floor(u); v = v - std::floor(v); break;
        case RXC_REPEAT_REPEAT_X: u = u - std::floor(u); v = rclamp(v, 0.0f, 1.0f); break;
        default: u = rclamp(u, 0.0f, 1.0f); v = v - std::floor(v); break;
    }
    if (sample_mode == RXC_SAMPLE_NEAREST) sample_nearest(t, u, v, out);
    else sample_linear(t, u, v, out);
}

// ---------------------------------------------------------------------------------------------
// src/map/light.rs:491-677  CompiledLight
// ---------------------------------------------------------------------------------------------
inline float smoothstep(float edge0, float edge1, float x) {  // :674-677
    float t = rclamp((x - edge0) / (edge1 - edge0), 0.0f, 1.0f);
    return t * t * (3.0f - 2.0f * t);
}
inline void apply_flicker(const rxc_light& l, const float color[3], float intensity, float flicker, uint32_t hash,
                          float out[3]) {  // :656-672
    float factor;
    if (flicker > 0.0f) {
        uint32_t s = (as_u32(l.position[0]) + as_u32(l.position[1]) + as_u32(l.position[2])) * 100u;
        uint32_t combined = hash + s;  // wrapping
        float fv = rclamp((float)combined / 4294967296.0f /* u32::MAX as f32 */, 0.0f, 1.0f);
        factor = 1.0f - fv * flicker;
    } else {
        factor = 1.0f;
    }
    out[0] = color[0] * intensity * factor;
    out[1] = color[1] * intensity * factor;
    out[2] = color[2] * intensity * factor;
}
inline V3 lpos(const rxc_light& l) { return {l.position[0], l.position[1], l.position[2]}; }

inline bool light_color_at(const rxc_light& l, V3 point, uint32_t hash, bool d2, float out[3]) {  // :491-502
    if (!l.emitting) return false;
    switch (l.light_type) {
        case RXC_LIGHT_POINT: {  // :535-552
            float distance = magnitude(point - lpos(l));
            if (distance >= l.end_distance) return false;
            if (distance <= l.start_distance) {
                apply_flicker(l, l.color, l.intensity, l.flicker, hash, out);
                return true;
            }
            float att = smoothstep(l.end_distance, l.start_distance, distance);
            apply_flicker(l, l.color, l.intensity * att, l.flicker, hash, out);
            return true;
        }
        case RXC_LIGHT_AMBIENT:
        case RXC_LIGHT_AMBIENT_DAYLIGHT:  // :554-557
            apply_flicker(l, l.color, l.intensity, l.flicker, hash, out);
            return true;
        case RXC_LIGHT_SPOT: {  // :559-580
            float distance = magnitude(point - lpos(l));
            if (distance >= l.end_distance) return false;
            float att = (distance <= l.start_distance)
                            ? 1.0f
                            : 1.0f - ((distance - l.start_distance) / (l.end_distance - l.start_distance));
            V3 dir_to_point = normalized(point - lpos(l));
            V3 d = {l.direction[0], l.direction[1], l.direction[2]};
            float angle = std::acos(dot(d, dir_to_point));
            if (angle > l.cone_angle) return false;
            apply_flicker(l, l.color, l.intensity * att, l.flicker, hash, out);
            return true;
        }
        case RXC_LIGHT_AREA: {  // :582-628
            V3 to_point = point - lpos(l);
            float distance = magnitude(to_point);
            if (distance >= l.end_distance) return false;
            if (distance < 0.1f) { out[0] = l.color[0]; out[1] = l.color[1]; out[2] = l.color[2]; return true; }
            float datt = (distance <= l.start_distance) ? 1.0f : smoothstep(l.end_distance, l.start_distance, distance);
            float area = l.width * l.height;
            V3 direction = normalized(to_point);
            float att;
            if (l.from_linedef) {
                att = datt * area * l.intensity;
            } else if (d2) {
                float dxn = std::fabs(to_point.x / (l.width * 0.5f));
                float dyn = std::fabs(to_point.y / (l.height * 0.5f));
                float ax = rmax(1.0f - dxn, 0.0f);
                float ay = rmax(1.0f - dyn, 0.0f);
                att = ax * ay * datt * l.intensity;
            } else {
                V3 n = {l.normal[0], l.normal[1], l.normal[2]};
                float aatt = rmax(dot(n, direction), 0.0f);
                att = aatt * datt * area * l.intensity;
            }
            out[0] = l.color[0] * att; out[1] = l.color[1] * att; out[2] = l.color[2] * att;
            return true;
        }
        default: {  // Daylight :630-653
            V3 to_point = point - lpos(l);
            float distance = magnitude(to_point);
            if (distance >= l.end_distance) return false;
            V3 direction = normalized(to_point);
            V3 n = {l.normal[0], l.normal[1], l.normal[2]};
            float aatt = rmax(dot(n, direction), 0.0f);
            float datt = (distance <= l.start_distance) ? 1.0f : smoothstep(l.end_distance, l.start_distance, distance);
            float att = aatt * datt * l.intensity;
            out[0] = l.color[0] * att; out[1] = l.color[1] * att; out[2] = l.color[2] * att;
            return true;
        }
    }
}
inline bool light_radiance_at(const rxc_light& l, V3 point, V3 normal, uint32_t hash, V3* out) {  // :504-533
    float c[3];
    if (!light_color_at(l, point, hash, false, c)) return false;
    V3 incoming = {c[0], c[1], c[2]};
    if (l.light_type == RXC_LIGHT_AMBIENT || l.light_type == RXC_LIGHT_AMBIENT_DAYLIGHT ||
        l.light_type == RXC_LIGHT_DAYLIGHT) {
        *out = incoming;
        return true;
    }
    V3 dir_to_light = normalized(lpos(l) - point);
    float lambert = rmax(dot(normal, dir_to_light), 0.0f);
    *out = incoming * lambert;
    return true;
}

// ---------------------------------------------------------------------------------------------
// src/shader/vgradient.rs:11-14, src/shader/grid.rs:36-108
// ---------------------------------------------------------------------------------------------
inline void shade_vgray(float uvy, uint8_t out[4]) {
    uint8_t i = as_u8(rclamp(uvy * 128.0f, 0.0f, 128.0f));
    out[0] = out[1] = out[2] = i;
    out[3] = 255;
}
inline void shade_grid(const rxc_frame& f, V2 uv, V2 screen, uint8_t out[4]) {
    auto closest_mul = [](V2 delta, V2 value) -> V2 {
        return {delta.x * std::round(value.x / delta.x), delta.y * std::round(value.y / delta.y)};
    };
    auto mul_dist = [&](V2 delta, V2 value) -> V2 {
        V2 c = closest_mul(delta, value);
        return {std::fabs(value.x - c.x), std::fabs(value.y - c.y)};
    };
    V2 position = {uv.x * screen.x, uv.y * screen.y};
    V2 origin = {screen.x / 2.0f + f.grid_offset[0], screen.y / 2.0f + f.grid_offset[1]};
    V2 grid_size = {f.grid_size, f.grid_size};
    V2 sub_grid_div = {f.grid_subdivisions, f.grid_subdivisions};
    const V4 bg_color = {0.05f, 0.05f, 0.05f, 1.0f};
    const V4 line_color = {0.15f, 0.15f, 0.15f, 1.0f};
    const V4 sub_line_color = {0.11f, 0.11f, 0.11f, 1.0f};
    const float th = 1.0f, sth = 1.0f;
    // align_pixel(origin, 1): thickness 1 is odd
    V2 aligned_origin = {std::round(origin.x - 0.5f) + 0.5f, std::round(origin.y - 0.5f) + 0.5f};
    V2 rel_p = {position.x - aligned_origin.x, position.y - aligned_origin.y};
    V2 dist = mul_dist(grid_size, rel_p);
    if (rmin(dist.x, dist.y) <= th * 0.5f) { vec4_to_pixel(line_color, out); return; }
    V2 dist_to_floor = {std::fabs(rel_p.x - grid_size.x * std::floor(rel_p.x / grid_size.x)),
                        std::fabs(rel_p.y - grid_size.y * std::floor(rel_p.y / grid_size.y))};
    V2 sub_size = {grid_size.x / std::round(sub_grid_div.x), grid_size.y / std::round(sub_grid_div.y)};
    V2 sub_dist = mul_dist(sub_size, dist_to_floor);
    V2 rc = {std::round(dist.x / sub_size.x), std::round(dist.y / sub_size.y)};
    V2 extra = {grid_size.x - sub_size.x * sub_grid_div.x, grid_size.y - sub_size.y * sub_grid_div.y};
    V2 sd = {rc.x == sub_grid_div.x ? sub_dist.x + extra.x : sub_dist.x,
             rc.y == sub_grid_div.y ? sub_dist.y + extra.y : sub_dist.y};
    if (rmin(sd.x, sd.y) <= sth * 0.5f) { vec4_to_pixel(sub_line_color, out); return; }
    vec4_to_pixel(bg_color, out);
}

// ---------------------------------------------------------------------------------------------
// Batch state: the projection caches of Batch3D / Batch2D (src/batch/batch3d.rs:28-62)
// ---------------------------------------------------------------------------------------------
struct Tri { size_t i0, i1, i2; };

inline Tri read_tri(const void* indices, uint32_t index_bytes, size_t t) {
    if (index_bytes == 8) {
        const uint64_t* p = (const uint64_t*)indices + t * 3;
        return {(size_t)p[0], (size_t)p[1], (size_t)p[2]};
    }
    const uint32_t* p = (const uint32_t*)indices + t * 3;
    return {p[0], p[1], p[2]};
}

struct Batch3DState {
    const rxc_batch3d* b = nullptr;
    std::vector<V4> projected_vertices;
    std::vector<Tri> clipped_indices;
    std::vector<V2> clipped_uvs;
    std::vector<V3> clipped_normals;
    std::vector<Edges> edges;
    std::optional<Rect> bounding_box;
    uint32_t owner_base = 0;
};

inline bool is_front_facing(const V4& v0, const V4& v1, const V4& v2) {  // src/batch/batch3d.rs:743-746
    float orientation = (v1.x - v0.x) * (v2.y - v0.y) - (v1.y - v0.y) * (v2.x - v0.x);
    return orientation > 0.0f;
}

// src/batch/batch3d.rs:482-740
void clip_and_project(Batch3DState& s, const M4& view_matrix, const M4& projection_matrix, float viewport_width,
                      float viewport_height, uint32_t mode) {
    const rxc_batch3d& b = *s.b;
    const V4* vertices = (const V4*)b.vertices;
    const V2* uvs = (const V2*)b.uvs;
    const V3* normals = (const V3*)b.normals;
    M4 transform;
    std::memcpy(transform.m, b.transform, sizeof(transform.m));

    M4 mvp = mat4_mul_mat4(mat4_mul_mat4(projection_matrix, view_matrix, mode), transform, mode);  // :490

    if (b.n_vertices != 0) {  // :493-552
        float min_x = INFINITY, min_y = INFINITY, min_z = INFINITY;
        float max_x = -INFINITY, max_y = -INFINITY, max_z = -INFINITY;
        for (uint32_t i = 0; i < b.n_vertices; ++i) {
            min_x = rmin(min_x, vertices[i].x); min_y = rmin(min_y, vertices[i].y); min_z = rmin(min_z, vertices[i].z);
            max_x = rmax(max_x, vertices[i].x); max_y = rmax(max_y, vertices[i].y); max_z = rmax(max_z, vertices[i].z);
        }
        const float corners[8][4] = {
            {min_x, min_y, min_z, 1.0f}, {min_x, min_y, max_z, 1.0f}, {min_x, max_y, min_z, 1.0f},
            {min_x, max_y, max_z, 1.0f}, {max_x, min_y, min_z, 1.0f}, {max_x, min_y, max_z, 1.0f},
            {max_x, max_y, min_z, 1.0f}, {max_x, max_y, max_z, 1.0f}};
        bool out_l = true, out_r = true, out_b = true, out_t = true, out_n = true, out_f = true;
        for (auto& c : corners) {
            V4 v = mat4_mul_vec4(mvp, {c[0], c[1], c[2], c[3]}, mode);
            float w = v.w;
            out_l &= v.x < -w; out_r &= v.x > w;
            out_b &= v.y < -w; out_t &= v.y > w;
            out_n &= v.z < -w; out_f &= v.z > w;
        }
        if (out_l || out_r || out_b || out_t || out_n || out_f) {
            s.projected_vertices.clear(); s.clipped_indices.clear(); s.clipped_uvs.clear();
            s.clipped_normals.clear(); s.edges.clear(); s.bounding_box.reset();
            return;
        }
    }

    M4 view_model = mat4_mul_mat4(view_matrix, transform, mode);  // :555
    std::vector<V4> view_space_vertices;
    view_space_vertices.reserve(b.n_vertices);
    for (uint32_t i = 0; i < b.n_vertices; ++i) view_space_vertices.push_back(mat4_mul_vec4(view_model, vertices[i], mode));

    const float near_plane = 0.1f;  // :563

    s.clipped_indices.clear(); s.clipped_uvs.clear(); s.clipped_normals.clear();
    for (uint32_t t = 0; t < b.n_triangles; ++t) s.clipped_indices.push_back(read_tri(b.indices, b.index_bytes, t));
    s.clipped_uvs.assign(uvs, uvs + b.n_vertices);
    if (normals) s.clipped_normals.assign(normals, normals + b.n_vertices);

    std::vector<V4> new_vertices;
    std::vector<V2> new_uvs;
    std::vector<V3> new_normals;
    std::vector<bool> edge_visibility(b.n_triangles, true);

    for (uint32_t triangle_idx = 0; triangle_idx < b.n_triangles; ++triangle_idx) {  // :586-681
        Tri tri = read_tri(b.indices, b.index_bytes, triangle_idx);
        V4 v0 = view_space_vertices[tri.i0], v1 = view_space_vertices[tri.i1], v2 = view_space_vertices[tri.i2];

        if (b.cull_mode != RXC_CULL_OFF) {  // :592-600
            float orient = (v1.x - v0.x) * (v2.y - v0.y) - (v1.y - v0.y) * (v2.x - v0.x);
            bool is_front = orient > 0.0f;
            if (b.cull_mode == RXC_CULL_BACK && is_front) continue;
            if (b.cull_mode == RXC_CULL_FRONT && !is_front) continue;
        }
        V2 uv0 = uvs[tri.i0], uv1 = uvs[tri.i1], uv2 = uvs[tri.i2];
        // the reference indexes self.normals unconditionally (:605-607, panics when empty);
        // the ABI allows an empty normal list and substitutes zero vectors.
        V3 n0 = normals ? normals[tri.i0] : V3{0, 0, 0};
        V3 n1 = normals ? normals[tri.i1] : V3{0, 0, 0};
        V3 n2 = normals ? normals[tri.i2] : V3{0, 0, 0};

        bool in0 = v0.z < -near_plane, in1 = v1.z < -near_plane, in2 = v2.z < -near_plane;
        if (in0 && in1 && in2) continue;
        edge_visibility[triangle_idx] = false;
        if (!in0 && !in1 && !in2) continue;

        const V4 vv[3] = {v0, v1, v2};
        const V2 uu[3] = {uv0, uv1, uv2};
        const V3 nn[3] = {n0, n1, n2};
        std::vector<size_t> ci;
        for (int i = 0; i < 3; ++i) {  // :630-669
            V4 current = vv[i]; V2 uv_current = uu[i]; V3 n_current = nn[i];
            V4 next = vv[(i + 1) % 3]; V2 uv_next = uu[(i + 1) % 3]; V3 n_next = nn[(i + 1) % 3];
            if (current.z < -near_plane) {
                new_vertices.push_back(current); new_uvs.push_back(uv_current); new_normals.push_back(n_current);
                ci.push_back(b.n_vertices + new_vertices.size() - 1);
                edge_visibility.push_back(true);
            }
            if ((current.z < -near_plane) != (next.z < -near_plane)) {
                float t = (-near_plane - current.z) / (next.z - current.z);
                V4 inter = {current.x + t * (next.x - current.x), current.y + t * (next.y - current.y),
                            current.z + t * (next.z - current.z), current.w + t * (next.w - current.w)};
                V2 iuv = {uv_current.x + t * (uv_next.x - uv_current.x), uv_current.y + t * (uv_next.y - uv_current.y)};
                V3 in = normalized(n_current * (1.0f - t) + n_next * t);
                new_vertices.push_back(inter); new_uvs.push_back(iuv); new_normals.push_back(in);
                ci.push_back(b.n_vertices + new_vertices.size() - 1);
                edge_visibility.push_back(true);
            }
        }
        for (size_t i = 1; i + 1 < ci.size(); ++i) s.clipped_indices.push_back({ci[0], ci[i], ci[i + 1]});  // :672-678
    }

    view_space_vertices.insert(view_space_vertices.end(), new_vertices.begin(), new_vertices.end());  // :684-686
    s.clipped_uvs.insert(s.clipped_uvs.end(), new_uvs.begin(), new_uvs.end());
    if (normals) s.clipped_normals.insert(s.clipped_normals.end(), new_normals.begin(), new_normals.end());

    s.projected_vertices.clear();  // :689-700
    for (const V4& v : view_space_vertices) {
        V4 r = mat4_mul_vec4(projection_matrix, v, mode);
        float w = r.w;
        s.projected_vertices.push_back(
            {((r.x / w) * 0.5f + 0.5f) * viewport_width, ((-r.y / w) * 0.5f + 0.5f) * viewport_height, r.z / w, w});
    }

    {  // :703, :749-768
        float min_x = INFINITY, max_x = -INFINITY, min_y = INFINITY, max_y = -INFINITY;
        for (const V4& v : s.projected_vertices) {
            min_x = rmin(min_x, v.x); max_x = rmax(max_x, v.x);
            min_y = rmin(min_y, v.y); max_y = rmax(max_y, v.y);
        }
        s.bounding_box = Rect{min_x, min_y, max_x - min_x, max_y - min_y};
    }

    s.edges.clear();  // :706-739
    for (size_t triangle_idx = 0; triangle_idx < s.clipped_indices.size(); ++triangle_idx) {
        const Tri& tri = s.clipped_indices[triangle_idx];
        V4 v0 = s.projected_vertices[tri.i0], v1 = s.projected_vertices[tri.i1], v2 = s.projected_vertices[tri.i2];
        bool visible;
        switch (b.cull_mode) {
            case RXC_CULL_OFF:
                if (is_front_facing(v0, v1, v2)) std::swap(v1, v2);
                visible = true;
                break;
            case RXC_CULL_FRONT: visible = !is_front_facing(v0, v1, v2); break;
            default:
                if (is_front_facing(v0, v1, v2)) { std::swap(v1, v2); visible = true; }
                else visible = false;
                break;
        }
        bool ev = (triangle_idx < edge_visibility.size() ? (bool)edge_visibility[triangle_idx] : true) && visible;
        const float a[3][2] = {{v0.x, v0.y}, {v1.x, v1.y}, {v2.x, v2.y}};
        const float bb[3][2] = {{v1.x, v1.y}, {v2.x, v2.y}, {v0.x, v0.y}};
        s.edges.push_back(edges_new(a, bb, ev));
    }
}

struct Batch2DState {
    const rxc_batch2d* b = nullptr;
    std::vector<V2> projected_vertices;
    std::optional<Rect> bounding_box;
    std::vector<Edges> edges;
};

// src/batch/batch2d.rs:373-425
void project2d(Batch2DState& s, const float* matrix /*nullable*/, uint32_t mode) {
    const rxc_batch2d& b = *s.b;
    const V2* vertices = (const V2*)b.vertices;
    s.projected_vertices.clear();
    float min_x = INFINITY, max_x = -INFINITY, min_y = INFINITY, max_y = -INFINITY;
    for (uint32_t i = 0; i < b.n_vertices; ++i) {
        V2 p = vertices[i];
        if (matrix) {
            V3 r = mat3_mul_vec3(matrix, {p.x, p.y, 1.0f}, mode);
            p = {r.x, r.y};
        }
        min_x = rmin(min_x, p.x); max_x = rmax(max_x, p.x);
        min_y = rmin(min_y, p.y); max_y = rmax(max_y, p.y);
        s.projected_vertices.push_back(p);
    }
    s.bounding_box = Rect{min_x, min_y, max_x - min_x, max_y - min_y};
    s.edges.clear();
    for (uint32_t t = 0; t < b.n_triangles; ++t) {
        Tri tri = read_tri(b.indices, b.index_bytes, t);
        V2 v0 = s.projected_vertices[tri.i0], v1 = s.projected_vertices[tri.i1], v2 = s.projected_vertices[tri.i2];
        const float a[3][2] = {{v0.x, v0.y}, {v1.x, v1.y}, {v2.x, v2.y}};
        const float bb[3][2] = {{v1.x, v1.y}, {v2.x, v2.y}, {v0.x, v0.y}};
        s.edges.push_back(edges_new(a, bb, true));
    }
}

// ---------------------------------------------------------------------------------------------
// The rasterizer proper
// ---------------------------------------------------------------------------------------------
struct TileRect { size_t x, y, width, height; };  // src/rasterizer.rs:2013-2019

// ---------------------------------------------------------------------------------------------
// rusteria::Execution + Program (rusteria/src/node/execution.rs:8-779, program.rs:7-29).
// Programs arrive as the op TREE (rusterix_b200/vm.py Program.encode_tree), and execute() recurses over it
// exactly like the reference; the device runs the same programs lowered to jumps, so the two interpreters
// share nothing but the opcode numbers of include/rxcuda.h.
// ---------------------------------------------------------------------------------------------
struct VmProgram {
    const uint32_t* t = nullptr;        // the encoded tree
    uint32_t n_functions() const { return t[0]; }
    bool has_shade() const { return t && t[1] != 0xFFFFFFFFu; }
    uint32_t shade_index() const { return t[1]; }
    uint32_t shade_locals() const { return t[2]; }
    uint32_t globals() const { return t[3]; }
    const uint32_t* function(uint32_t i, uint32_t* n) const { const uint32_t* f = t + t[5 + i]; *n = f[0]; return f + 1; }
};
struct VmBank {                          // set by rxo_set_programs (test infrastructure, not thread safe)
    std::vector<std::vector<uint32_t>> trees;
    std::vector<VmProgram> programs;
    std::vector<rxc_pattern> patterns, patterns_normal;
    std::vector<std::vector<float>> pattern_store;
    std::vector<float> palette;          // n * (present, r, g, b)
};
VmBank g_vm;
// 0 (default) = the reference: one Execution per screen tile, never reset (src/rasterizer.rs:310).
// 1 = the device's documented deviation: every fragment starts from Execution::new() (rxo_set_vm_state_mode).
int g_vm_fresh_state = 0;

inline float f32_from_bits(uint32_t b) { float f; std::memcpy(&f, &b, 4); return f; }

struct Execution {
    std::vector<V3> globals, locals, stack;
    std::vector<std::vector<V3>> locals_stack;
    bool has_return = false; V3 return_value{0, 0, 0};
    V3 uv{0, 0, 0}, color{0, 0, 0}, roughness{0.5f, 0.5f, 0.5f}, metallic{0, 0, 0}, emissive{0, 0, 0}, opacity{0, 0, 0},
        bump{0, 0, 0}, normal{0, 0, 0}, hitpoint{0, 0, 0}, time{0, 0, 0};
    bool fault = false;                  // the reference would have panicked (pop of an empty stack, bad index)

    V3 pop() { if (stack.empty()) { fault = true; return {0, 0, 0}; } V3 v = stack.back(); stack.pop_back(); return v; }
    static V3 map1(V3 a, float (*fn)(float)) { return {fn(a.x), fn(a.y), fn(a.z)}; }
    static V3 splat(float x) { return {x, x, x}; }

    // TexStorage::sample, rusteria/src/textures/mod.rs:20-24, :131-146
    static V3 pattern_sample(const rxc_pattern& p, V3 uv) {
        float u = uv.x - std::floor(uv.x), v = uv.y - std::floor(uv.y);
        int32_t x = as_i32(std::floor(u * (float)p.width)), y = as_i32(std::floor(v * (float)p.height));
        auto rem = [](int32_t a, int32_t m) { int32_t r = a % m; return r < 0 ? r + m : r; };
        x = rem(x, (int32_t)p.width); y = rem(y, (int32_t)p.height);
        const float* d = p.data + ((size_t)y * p.width + (size_t)x) * 3;
        return {d[0], d[1], d[2]};
    }

    void reset(size_t var_size) { if (var_size != globals.size()) globals.resize(var_size, V3{0, 0, 0}); }  // execution.rs:103-107

    // execution.rs:109-746
    void execute(const uint32_t* code, uint32_t n, const VmProgram& program) {
        for (uint32_t pc = 0; pc < n;) {
            if (has_return || fault) break;
            const uint32_t op = code[pc++];
            switch (op) {
                case RXVM_LOAD_GLOBAL: { uint32_t i = code[pc++]; if (i >= globals.size()) { fault = true; break; } stack.push_back(globals[i]); break; }
                case RXVM_STORE_GLOBAL: { uint32_t i = code[pc++]; if (i >= globals.size()) { fault = true; break; } globals[i] = pop(); break; }
                case RXVM_LOAD_LOCAL: { uint32_t i = code[pc++]; if (i >= locals.size()) { fault = true; break; } stack.push_back(locals[i]); break; }
                case RXVM_STORE_LOCAL: { uint32_t i = code[pc++]; if (i >= locals.size()) { fault = true; break; } locals[i] = pop(); break; }
                case RXVM_SWAP: { V3 b = pop(), a = pop(); stack.push_back(b); stack.push_back(a); break; }
                case RXVM_GET_COMPONENTS: {
                    uint32_t len = code[pc++];
                    V3 v = pop();
                    float r[8]; uint32_t k = 0;
                    for (uint32_t i = 0; i < len; ++i) {
                        uint32_t index = code[pc++];
                        if (index > 2) continue;
                        if (k < 8) r[k] = index == 0 ? v.x : index == 1 ? v.y : v.z;
                        ++k;
                    }
                    V3 pushed = k == 1 ? splat(r[0]) : k == 2 ? V3{r[0], r[1], 0.0f} : k == 3 ? V3{r[0], r[1], r[2]} : splat(0.0f);
                    stack.push_back(pushed);
                    break;
                }
                case RXVM_SET_COMPONENTS: {
                    uint32_t len = code[pc++];
                    V3 value = pop(), target = pop();
                    float comps[3] = {value.x, value.y, value.z};
                    uint32_t ncomp = (len >= 1 && len <= 3) ? len : 0;
                    for (uint32_t i = 0; i < len; ++i) {
                        uint32_t idx = code[pc + i];
                        if (i >= ncomp) break;
                        if (idx == 0) target.x = comps[i]; else if (idx == 1) target.y = comps[i]; else if (idx == 2) target.z = comps[i];
                    }
                    pc += len;
                    stack.push_back(target);
                    break;
                }
                case RXVM_PUSH: stack.push_back({f32_from_bits(code[pc]), f32_from_bits(code[pc + 1]), f32_from_bits(code[pc + 2])}); pc += 3; break;
                case RXVM_CLEAR: if (!stack.empty()) stack.pop_back(); break;
                case RXVM_FUNCTION_CALL: {  // :186-223
                    uint32_t arity = code[pc++], total_locals = code[pc++], index = code[pc++];
                    locals_stack.push_back(locals);
                    locals.assign(total_locals, V3{0, 0, 0});
                    for (uint32_t i = arity; i-- > 0;)
                        if (!stack.empty()) { V3 arg = stack.back(); stack.pop_back(); if (i < locals.size()) locals[i] = arg; else fault = true; }
                    size_t stack_base = stack.size();
                    if (index >= program.n_functions()) { fault = true; break; }
                    uint32_t fn_n; const uint32_t* body = program.function(index, &fn_n);
                    execute(body, fn_n, program);
                    V3 ret{0, 0, 0};
                    if (has_return) { ret = return_value; has_return = false; }
                    else if (stack.size() > stack_base) { ret = stack.back(); stack.pop_back(); }
                    if (stack.size() > stack_base) stack.resize(stack_base);
                    locals = locals_stack.back(); locals_stack.pop_back();
                    stack.push_back(ret);
                    break;
                }
                case RXVM_RETURN: {  // :224-234
                    V3 v{0, 0, 0};
                    if (!stack.empty()) { v = stack.back(); stack.pop_back(); }
                    return_value = v; has_return = true;
                    break;
                }
                case RXVM_PACK2: { V3 y = pop(), x = pop(); stack.push_back({x.x, y.x, 0.0f}); break; }
                case RXVM_PACK3: { V3 z = pop(), y = pop(), x = pop(); stack.push_back({x.x, y.x, z.x}); break; }
                case RXVM_DUP: if (!stack.empty()) stack.push_back(stack.back()); break;
                case RXVM_FOR: {  // :259-285.  A Return inside the loop leaves the function (the reference pops the
                                  // condition from an empty or foreign stack after it).
                    uint32_t n_init = code[pc], n_cond = code[pc + 1], n_incr = code[pc + 2], n_body = code[pc + 3];
                    const uint32_t *init = code + pc + 4, *cond = init + n_init, *incr = cond + n_cond, *body = incr + n_incr;
                    pc += 4 + n_init + n_cond + n_incr + n_body;
                    size_t base = stack.size();
                    size_t iter = 0;
                    execute(init, n_init, program);
                    if (stack.size() > base) stack.resize(base);
                    for (;;) {
                        execute(cond, n_cond, program);
                        if (has_return || fault) break;
                        V3 z = pop();
                        if (z.x == 0.0f) break;
                        if (stack.size() > base) stack.resize(base);
                        execute(body, n_body, program);
                        if (stack.size() > base) stack.resize(base);
                        execute(incr, n_incr, program);
                        if (stack.size() > base) stack.resize(base);
                        if (++iter > 10000000) { fault = true; break; }
                    }
                    break;
                }
                case RXVM_IF: {  // :286-293
                    uint32_t n_then = code[pc], n_else = code[pc + 1];
                    const uint32_t* then_code = code + pc + 2;
                    const uint32_t* else_code = then_code + n_then;
                    pc += 2 + n_then + (n_else == 0xFFFFFFFFu ? 0 : n_else);
                    bool value = pop().x != 0.0f;
                    if (value) execute(then_code, n_then, program);
                    else if (n_else != 0xFFFFFFFFu) execute(else_code, n_else, program);
                    break;
                }
                case RXVM_ADD: { V3 b = pop(), a = pop(); stack.push_back(a + b); break; }
                case RXVM_SUB: { V3 b = pop(), a = pop(); stack.push_back(a - b); break; }
                case RXVM_MUL: { V3 b = pop(), a = pop(); stack.push_back(a * b); break; }
                case RXVM_DIV: { V3 b = pop(), a = pop(); stack.push_back({a.x / b.x, a.y / b.y, a.z / b.z}); break; }
                case RXVM_LENGTH: { V3 a = pop(); stack.push_back(splat(magnitude(a))); break; }
                case RXVM_LENGTH2: { V3 a = pop(); stack.push_back({std::sqrt(a.x * a.x + a.y * a.y), 0.0f, 0.0f}); break; }
                case RXVM_LENGTH3: { V3 a = pop(); stack.push_back({std::sqrt(a.x * a.x + a.y * a.y + a.z * a.z), 0.0f, 0.0f}); break; }
                case RXVM_ABS: stack.push_back(map1(pop(), [](float x) { return std::fabs(x); })); break;
                case RXVM_SIN: stack.push_back(map1(pop(), [](float x) { return std::sin(x); })); break;
                case RXVM_SIN1: case RXVM_COS1: { V3 a = pop(); stack.push_back({std::sin(a.x), 0.0f, 0.0f}); break; }   // :342-345 (sin!)
                case RXVM_SIN2: case RXVM_COS2: { V3 a = pop(); stack.push_back({std::sin(a.x), std::sin(a.y), 0.0f}); break; }
                case RXVM_COS: stack.push_back(map1(pop(), [](float x) { return std::cos(x); })); break;
                case RXVM_NORMALIZE: { V3 a = pop(); float len = magnitude(a); stack.push_back(len > 0.0f ? a / len : a); break; }
                case RXVM_TAN: stack.push_back(map1(pop(), [](float x) { return std::tan(x); })); break;
                case RXVM_ATAN: stack.push_back(map1(pop(), [](float x) { return std::atan(x); })); break;
                case RXVM_ATAN2: { V3 b = pop(), a = pop(); stack.push_back({std::atan2(a.x, b.x), std::atan2(a.y, b.y), std::atan2(a.z, b.z)}); break; }
                case RXVM_ROTATE2D: {
                    V3 angle = pop(), v = pop();
                    float rad = angle.x * 0.017453292519943295f;   // f32::to_radians
                    float sn = std::sin(rad), c = std::cos(rad);
                    stack.push_back({v.x * c - v.y * sn, v.x * sn + v.y * c, v.z});
                    break;
                }
                case RXVM_DOT: { V3 b = pop(), a = pop(); stack.push_back(splat(dot(a, b))); break; }
                case RXVM_DOT2: { V3 b = pop(), a = pop(); stack.push_back({a.x * b.x + a.y * b.y, 0.0f, 0.0f}); break; }
                case RXVM_DOT3: { V3 b = pop(), a = pop(); stack.push_back({a.x * b.x + a.y * b.y + a.z * b.z, 0.0f, 0.0f}); break; }
                case RXVM_CROSS: { V3 b = pop(), a = pop(); stack.push_back({a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}); break; }
                case RXVM_FLOOR: stack.push_back(map1(pop(), [](float x) { return std::floor(x); })); break;
                case RXVM_CEIL: stack.push_back(map1(pop(), [](float x) { return std::ceil(x); })); break;
                case RXVM_ROUND: stack.push_back(map1(pop(), [](float x) { return std::round(x); })); break;
                case RXVM_FRACT: stack.push_back(map1(pop(), [](float x) { return x - std::floor(x); })); break;
                case RXVM_MOD: {
                    V3 b = pop(), a = pop();
                    stack.push_back({a.x - b.x * std::floor(a.x / b.x), a.y - b.y * std::floor(a.y / b.y), a.z - b.z * std::floor(a.z / b.z)});
                    break;
                }
                case RXVM_RADIANS: stack.push_back(map1(pop(), [](float x) { return x * 0.017453292519943295f; })); break;
                case RXVM_DEGREES: stack.push_back(map1(pop(), [](float x) { return x * 57.29577951308232f; })); break;
                case RXVM_MIN: { V3 b = pop(), a = pop(); stack.push_back({rmin(a.x, b.x), rmin(a.y, b.y), rmin(a.z, b.z)}); break; }
                case RXVM_MAX: { V3 b = pop(), a = pop(); stack.push_back({rmax(a.x, b.x), rmax(a.y, b.y), rmax(a.z, b.z)}); break; }
                case RXVM_MIX: { V3 c = pop(), b = pop(), a = pop(); stack.push_back(a + (b - a) * c); break; }
                case RXVM_SMOOTHSTEP: {
                    V3 c = pop(), b = pop(), a = pop();
                    float denom = b.x - a.x;
                    float t = denom != 0.0f ? (c.x - a.x) / denom : 0.0f;
                    if (t < 0.0f) t = 0.0f; else if (t > 1.0f) t = 1.0f;
                    stack.push_back(splat(t * t * (3.0f - 2.0f * t)));
                    break;
                }
                case RXVM_STEP: { V3 b = pop(), a = pop(); stack.push_back({b.x >= a.x ? 1.0f : 0.0f, b.y >= a.y ? 1.0f : 0.0f, b.z >= a.z ? 1.0f : 0.0f}); break; }
                case RXVM_CLAMP: { V3 c = pop(), b = pop(), a = pop(); stack.push_back({rclamp(a.x, b.x, c.x), rclamp(a.y, b.y, c.y), rclamp(a.z, b.z, c.z)}); break; }
                case RXVM_SQRT: stack.push_back(map1(pop(), [](float x) { return std::sqrt(x); })); break;
                case RXVM_LOG: stack.push_back(map1(pop(), [](float x) { return std::log(x); })); break;
                case RXVM_POW: { V3 b = pop(), a = pop(); stack.push_back({std::pow(a.x, b.x), std::pow(a.y, b.y), std::pow(a.z, b.z)}); break; }
                case RXVM_EQ: { V3 b = pop(), a = pop(); stack.push_back(splat(a.x == b.x ? 1.0f : 0.0f)); break; }
                case RXVM_NE: { V3 b = pop(), a = pop(); stack.push_back(splat(a.x != b.x ? 1.0f : 0.0f)); break; }
                case RXVM_LT: { V3 b = pop(), a = pop(); stack.push_back(splat(a.x < b.x ? 1.0f : 0.0f)); break; }
                case RXVM_LE: { V3 b = pop(), a = pop(); stack.push_back(splat(a.x <= b.x ? 1.0f : 0.0f)); break; }
                case RXVM_GT: { V3 b = pop(), a = pop(); stack.push_back(splat(a.x > b.x ? 1.0f : 0.0f)); break; }
                case RXVM_GE: { V3 b = pop(), a = pop(); stack.push_back(splat(a.x >= b.x ? 1.0f : 0.0f)); break; }
                case RXVM_AND: { V3 b = pop(), a = pop(); stack.push_back(splat(((a.x != 0.0f) & (b.x != 0.0f)) ? 1.0f : 0.0f)); break; }
                case RXVM_OR: { V3 b = pop(), a = pop(); stack.push_back(splat(((a.x != 0.0f) | (b.x != 0.0f)) ? 1.0f : 0.0f)); break; }
                case RXVM_NOT: { V3 a = pop(); stack.push_back(splat(a.x == 0.0f ? 1.0f : 0.0f)); break; }
                case RXVM_NEG: stack.push_back(-pop()); break;
                case RXVM_PRINT: pop(); break;
                case RXVM_UV: stack.push_back(uv); break;
                case RXVM_SET_UV: uv = pop(); break;
                case RXVM_NORMAL: stack.push_back(normal); break;
                case RXVM_SET_NORMAL: normal = normalized(pop()); break;
                case RXVM_HITPOINT: stack.push_back(hitpoint); break;
                case RXVM_TIME: stack.push_back(time); break;
                case RXVM_COLOR: stack.push_back(color); break;
                case RXVM_SET_COLOR: color = pop(); break;
                case RXVM_ROUGHNESS: stack.push_back(roughness); break;
                case RXVM_SET_ROUGHNESS: roughness = pop(); break;
                case RXVM_METALLIC: stack.push_back(metallic); break;
                case RXVM_SET_METALLIC: metallic = pop(); break;
                case RXVM_EMISSIVE: stack.push_back(emissive); break;
                case RXVM_SET_EMISSIVE: emissive = pop(); break;
                case RXVM_OPACITY: stack.push_back(opacity); break;
                case RXVM_SET_OPACITY: opacity = pop(); break;
                case RXVM_BUMP: stack.push_back(bump); break;
                case RXVM_SET_BUMP: bump = pop(); break;
                case RXVM_SAMPLE: {  // :625-633
                    V3 b = pop(), a = pop();
                    size_t i = as_usize(b.x);
                    stack.push_back(i < g_vm.patterns.size() ? pattern_sample(g_vm.patterns[i], a) : V3{0, 0, 0});
                    break;
                }
                case RXVM_SAMPLE_NORMAL: {  // :634-649
                    V3 b = pop(), a = pop();
                    size_t i = as_usize(b.x);
                    if (i < g_vm.patterns_normal.size()) { V3 nm = pattern_sample(g_vm.patterns_normal[i], a); stack.push_back(nm * 2.0f - V3{1, 1, 1}); }
                    else stack.push_back({0, 0, 0});
                    break;
                }
                case RXVM_PALETTE_INDEX: {  // :735-742
                    V3 a = pop();
                    size_t i = as_usize(a.x);
                    if (i < g_vm.palette.size() / 4 && g_vm.palette[i * 4] != 0.0f)
                        stack.push_back({g_vm.palette[i * 4 + 1], g_vm.palette[i * 4 + 2], g_vm.palette[i * 4 + 3]});
                    break;
                }
                default: fault = true; break;  // Alloc / Iterate / Save: texture baking, not reachable from the rasterizer path
            }
        }
    }

    // execution.rs:771-779
    void shade(const VmProgram& program) {
        stack.clear();
        has_return = false;
        locals.resize(program.shade_locals(), V3{0, 0, 0});
        uint32_t n; const uint32_t* body = program.function(program.shade_index(), &n);
        execute(body, n, program);
    }
};

struct Raster {
    const rxc_tile* tiles; uint32_t n_tiles;
    const rxc_scene* scene;
    const rxc_frame* f;
    std::vector<Batch3DState>* b3;
    std::vector<Batch2DState>* b2;
    M4 inverse_view, inverse_projection;
    V3 camera_pos;
    float width, height;
    uint32_t hash_anim;
    V2 translationd2; float scaled2;
    const rxc_mapmini* mapmini = nullptr;

    // src/map/bbox.rs:35-40 + src/chunk.rs:154-161 / src/map/mini.rs:58-65
    static float sector_occlusion(const rxc_sector* sectors, uint32_t n, V2 at) {
        for (uint32_t i = 0; i < n; ++i) {
            const rxc_sector& b = sectors[i];
            if (at.x >= b.min[0] && at.x <= b.max[0] && at.y >= b.min[1] && at.y <= b.max[1]) return b.occlusion;
        }
        return 1.0f;
    }
    float get_occlusion(int32_t chunk, V2 at) const {  // :1327-1331, :808-812
        if (chunk >= 0) return sector_occlusion(scene->chunks[chunk].occluded_sectors, scene->chunks[chunk].n_occluded_sectors, at);
        if (mapmini) return sector_occlusion(mapmini->occluded_sectors, mapmini->n_occluded_sectors, at);
        return 1.0f;
    }
    // src/map/mini.rs:67-95
    static bool segments_intersect(V2 a1, V2 a2, V2 b1, V2 b2) {
        float d = (a2.x - a1.x) * (b2.y - b1.y) - (a2.y - a1.y) * (b2.x - b1.x);
        if (d == 0.0f) return false;
        float u = ((b1.x - a1.x) * (b2.y - b1.y) - (b1.y - a1.y) * (b2.x - b1.x)) / d;
        float v = ((b1.x - a1.x) * (a2.y - a1.y) - (b1.y - a1.y) * (a2.x - a1.x)) / d;
        return u >= 0.0f && u <= 1.0f && v >= 0.0f && v <= 1.0f;
    }
    bool is_visible(V2 from, V2 to) const {
        if (!mapmini) return true;
        for (uint32_t i = 0; i < mapmini->n_linedefs; ++i) {
            const rxc_linedef& l = mapmini->linedefs[i];
            if (segments_intersect(from, to, {l.start[0], l.start[1]}, {l.end[0], l.end[1]})) return false;
        }
        return true;
    }
    // src/chunk.rs:135-151 (scale = 1) + src/texture.rs:527-538
    void sample_terrain_texture(int32_t chunk, V2 world_pos, uint8_t out[4]) const {
        const rxc_chunk& c = scene->chunks[chunk];
        float local_x = (world_pos.x / 1.0f) - (float)c.origin[0];
        float local_y = (world_pos.y / 1.0f) - (float)c.origin[1];
        out[0] = out[1] = out[2] = out[3] = 0;
        if (!c.terrain_texture) return;
        const rxc_texture& t = *c.terrain_texture;
        int32_t pixels_per_tile = (int32_t)t.width / c.size;
        float pixel_x = local_x * (float)pixels_per_tile;
        float pixel_y = local_y * (float)pixels_per_tile;
        uint32_t px = as_u32(rclamp(std::floor(pixel_x), 0.0f, (float)t.width - 1.0f));
        uint32_t py = as_u32(rclamp(std::floor(pixel_y), 0.0f, (float)t.height - 1.0f));
        size_t x = std::min(px, t.width - 1), y = std::min(py, t.height - 1);
        std::memcpy(out, t.data + (y * t.width + x) * 4, 4);
    }

    // src/rasterizer.rs:1707-1727
    V3 screen_to_world(float x, float y, float z_ndc) const {
        float x_ndc = 2.0f * (x / width) - 1.0f;
        float y_ndc = 1.0f - 2.0f * (y / height);
        V4 ndc = {x_ndc, y_ndc, z_ndc, 1.0f};
        V4 vs = mat4_mul_vec4(inverse_projection, ndc, f->matvec_mode);
        vs = {vs.x / vs.w, vs.y / vs.w, vs.z / vs.w, vs.w / vs.w};
        V4 ws = mat4_mul_vec4(inverse_view, vs, f->matvec_mode);
        return {ws.x, ws.y, ws.z};
    }

    // src/rasterizer.rs:1844-1870
    void screen_ray(float x, float y, V3* origin, V3* dir) const {
        float ndc_x = 2.0f * (x / width) - 1.0f;
        float ndc_y = 1.0f - 2.0f * (y / height);
        V4 vn = mat4_mul_vec4(inverse_projection, {ndc_x, ndc_y, -1.0f, 1.0f}, f->matvec_mode);
        V4 vf = mat4_mul_vec4(inverse_projection, {ndc_x, ndc_y, 1.0f, 1.0f}, f->matvec_mode);
        vn = {vn.x / vn.w, vn.y / vn.w, vn.z / vn.w, vn.w / vn.w};
        vf = {vf.x / vf.w, vf.y / vf.w, vf.z / vf.w, vf.w / vf.w};
        V4 wn = mat4_mul_vec4(inverse_view, vn, f->matvec_mode);
        V4 wf = mat4_mul_vec4(inverse_view, vf, f->matvec_mode);
        *origin = {wn.x, wn.y, wn.z};
        *dir = normalized(V3{wf.x, wf.y, wf.z} - *origin);
    }
    // vek Vec4::lerp: clamped factor, mul_add (see shade_fast_brdf)
    static float lerp1(float from, float to, float t) { return std::fmaf(rclamp(t, 0.0f, 1.0f), to - from, from); }
    // ShapeFX Sky render_miss_d3, src/shapestack/shapefx.rs:1122-1170 (the cloud layer :1172-1219 needs noiselib:
    // frames that ask for it are rejected by validate())
    void render_miss_sky(V4& color, V3 dir) const {
        const float* sun = f->sky[0];
        const float* haze_color = f->sky[1];
        const float *day_h = f->sky[2], *day_z = f->sky[3], *night_h = f->sky[4], *night_z = f->sky[5];
        float day_factor = sun[3];
        float up = rclamp(dir.y, -1.0f, 1.0f);
        float t = (up + 1.0f) * 0.5f;
        float c[4];
        for (int i = 0; i < 4; ++i) c[i] = lerp1(lerp1(night_h[i], night_z[i], t), lerp1(day_h[i], day_z[i], t), day_factor);
        float om = 1.0f - up;
        float haze = om * om * om;  // powi(3)
        for (int i = 0; i < 4; ++i) {
            float fog = haze_color[i] * haze * 0.3f;
            c[i] = c[i] * (1.0f - haze * 0.2f) + fog;
        }
        if (day_factor > 0.0f) {
            float d = rclamp(dot(dir, V3{sun[0], sun[1], sun[2]}), -1.0f, 1.0f);
            float dist = rmax(1.0f - d, 0.0f);
            if (dist < 0.04f) {
                float k = 1.0f - dist / 0.04f;
                float glare = k * k * (3.0f - 2.0f * k);
                const float g[4] = {1.0f, 0.85f, 0.6f, 0.0f};
                for (int i = 0; i < 4; ++i) c[i] += g[i] * glare * day_factor;
            }
        }
        color = {c[0], c[1], c[2], c[3]};
    }
    // src/rasterizer.rs:434-456
    void brush_preview(V4& color, V3 origin, V3 dir) const {
        if (!(std::fabs(dir.y) > 1e-5f)) return;
        float t = -origin.y / dir.y;
        if (!(t > 0.0f)) return;
        V3 world = origin + dir * t;
        float dist = magnitude(world - V3{f->brush_position[0], f->brush_position[1], f->brush_position[2]});
        if (!(dist < f->brush_radius)) return;
        float normalized_d = dist / f->brush_radius;
        float falloff = rclamp(f->brush_falloff, 0.001f, 1.0f);
        float fade = rclamp((1.0f - normalized_d) / falloff, 0.0f, 1.0f);
        float blend = 0.2f + 0.6f * fade;
        color.x = rmin(color.x * (1.0f - blend) + blend, 1.0f);
        color.y = rmin(color.y * (1.0f - blend) + blend, 1.0f);
        color.z = rmin(color.z * (1.0f - blend) + blend, 1.0f);
    }

    // src/rasterizer.rs:1754-1773 (and :1731-1750, same arithmetic on [f32;2])
    static void barycentric_weights(float ax, float ay, float bx, float by, float cx, float cy, float px, float py,
                                    float w[3]) {
        float ac[2] = {cx - ax, cy - ay};
        float ab[2] = {bx - ax, by - ay};
        float ap[2] = {px - ax, py - ay};
        float pc[2] = {cx - px, cy - py};
        float pb[2] = {bx - px, by - py};
        float area = ac[0] * ab[1] - ac[1] * ab[0];
        float alpha = (pc[0] * pb[1] - pc[1] * pb[0]) / area;
        float beta = (ac[0] * ap[1] - ac[1] * ap[0]) / area;
        float gamma = 1.0f - alpha - beta;
        w[0] = alpha; w[1] = beta; w[2] = gamma;
    }

    // src/rasterizer.rs:1875-1951
    static float roughness_to_shininess(float r) {
        float a = rmax(r * r, 1e-4f);
        return rclamp(2.0f / a - 2.0f, 1.0f, 2048.0f);
    }
    static V3 schlick_fresnel(V3 f0, float cos_theta) {
        float one_minus = 1.0f - rclamp(cos_theta, 0.0f, 1.0f);
        float x = one_minus * one_minus * one_minus * one_minus * one_minus;
        return f0 + (V3{1, 1, 1} - f0) * x;
    }
    static float pow32_fast(float x, float y) {
        if (x <= 0.0f) return 0.0f;
        return std::exp2(y * std::log2(x));
    }
    static float blinn_phong_spec(V3 n, V3 l, V3 v, float shininess) {
        V3 h = normalized(l + v);
        float n_dot_h = rmax(dot(n, h), 0.0f);
        return pow32_fast(n_dot_h, shininess);
    }
    static V3 shade_fast_brdf(V3 base_color, float roughness, float metallic, V3 emissive, V3 n, V3 v, V3 l,
                              V3 light_radiance) {
        float n_dot_l = rmax(dot(n, l), 0.0f);
        if (n_dot_l <= 0.0f) return emissive;
        // vek Vec3::lerp(from,to,t) = lerp_unclamped(from,to,clamped01(t)) = t.mul_add(to-from, from)
        float t = rclamp(metallic, 0.0f, 1.0f);
        V3 f0 = {std::fmaf(t, base_color.x - 0.04f, 0.04f), std::fmaf(t, base_color.y - 0.04f, 0.04f),
                 std::fmaf(t, base_color.z - 0.04f, 0.04f)};
        V3 kd = base_color * (1.0f - metallic);
        kd = kd * (1.0f - rmax(f0.x, rmax(f0.y, f0.z)));
        float shininess = roughness_to_shininess(roughness);
        float spec_b = blinn_phong_spec(n, l, v, shininess);
        float n_dot_v = rmax(dot(n, v), 0.0f);
        V3 fr = schlick_fresnel(f0, n_dot_v);
        V3 diffuse = kd * n_dot_l;
        V3 specular = fr * spec_b * n_dot_l;
        return (diffuse + specular) * light_radiance + emissive;
    }

    // `chunk.shaders.get(i)` / `scene.shaders.get(i)` (src/rasterizer.rs:1281-1285): nullptr when nothing runs
    const VmProgram* program_for(int32_t shader, int32_t chunk) const {
        if (shader < 0) return nullptr;
        uint32_t index;
        if (chunk >= 0) {
            const rxc_chunk& c = scene->chunks[chunk];
            if ((uint32_t)shader >= c.n_shaders) return nullptr;
            index = c.shader_base + (uint32_t)shader;
        } else {
            if ((uint32_t)shader >= scene->n_scene_shaders) return nullptr;
            index = (uint32_t)shader;
        }
        if (index >= g_vm.programs.size() || !g_vm.programs[index].has_shade()) return nullptr;
        return &g_vm.programs[index];
    }
    // `chunk.shader_textures.get(i)` (src/rasterizer.rs:1227-1236)
    const rxc_texture* baked_texture(int32_t shader, int32_t chunk) const {
        if (shader < 0 || chunk < 0) return nullptr;
        const rxc_chunk& c = scene->chunks[chunk];
        if (!c.shader_textures || (uint32_t)shader >= c.n_shaders) return nullptr;
        return c.shader_textures[shader];
    }

    // Texel fetch shared by d3 and d2.  Returns false when the reference would panic.
    const rxc_texture* tile_frame(uint32_t kind, uint32_t index) const {
        const rxc_tile* t = nullptr;
        if (kind == RXC_SRC_STATIC_TILE) { if (index < n_tiles) t = &tiles[index]; }
        else if (kind == RXC_SRC_DYNAMIC_TILE) { if (index < scene->n_dynamic_textures) t = &scene->dynamic_textures[index]; }
        else { if (index < scene->n_actor_tiles) t = &scene->actor_tiles[index]; }  // EntityTile / ItemTile, host-resolved
        if (!t || t->n_textures == 0) return nullptr;
        return &t->textures[f->animation_frame % t->n_textures];  // :1104-1105
    }

    // 3D texel of a fragment, src/rasterizer.rs:1100-1222 (and :1512-1598 for the opacity pass)
    void texel_3d(const rxc_batch3d& batch, float u, float v, V3 world, uint8_t texel[4]) const {
        switch (batch.source_kind) {
            case RXC_SRC_STATIC_TILE:
            case RXC_SRC_DYNAMIC_TILE: {
                const rxc_texture* t = tile_frame(batch.source_kind, batch.source_index);
                texture_sample(*t, u, v, f->sample_mode, batch.repeat_mode, texel);
                break;
            }
            case RXC_SRC_PIXEL: std::memcpy(texel, batch.source_pixel, 4); break;
            case RXC_SRC_ENTITY_TILE:
            case RXC_SRC_ITEM_TILE: {
                const rxc_texture* t = tile_frame(batch.source_kind, batch.source_index);
                if (t) texture_sample(*t, u, v, f->sample_mode, batch.repeat_mode, texel);
                else texel[0] = texel[1] = texel[2] = texel[3] = 0;
                break;
            }
            case RXC_SRC_TERRAIN:
                if (batch.chunk >= 0) {
                    sample_terrain_texture(batch.chunk, {world.x, world.z}, texel);
                    if (f->has_brush_preview) {   // :1193-1212 (and :1601-1620): the texel goes towards white inside the brush
                        float dist = magnitude(world - V3{f->brush_position[0], f->brush_position[1], f->brush_position[2]});
                        if (dist < f->brush_radius) {
                            float normalized_d = dist / f->brush_radius;
                            float falloff = rclamp(f->brush_falloff, 0.001f, 1.0f);
                            float fade = rclamp((1.0f - normalized_d) / falloff, 0.0f, 1.0f);
                            float blend = 0.2f + 0.6f * fade;
                            for (int i = 0; i < 3; ++i) texel[i] = as_u8(rmin((float)texel[i] * (1.0f - blend) + 255.0f * blend, 255.0f));
                        }
                    }
                } else { texel[0] = 255; texel[1] = 0; texel[2] = 0; texel[3] = 255; }
                break;
            default: texel[0] = texel[1] = texel[2] = 0; texel[3] = 255; break;
        }
    }

    // src/rasterizer.rs:964-1420.  surface_id: 0 = None, else Some(value - 1) widened to 64 bits
    void d3_rasterize(std::vector<uint8_t>& buffer, std::vector<float>& z_buffer, std::vector<uint32_t>& owner,
                      const std::vector<uint64_t>& surface_id, const TileRect& tile, const Batch3DState& s,
                      Execution& execution) const {
        const rxc_batch3d& batch = *s.b;
        const uint64_t profile = batch.has_profile_id ? (uint64_t)batch.profile_id + 1u : 0u;
        if (!s.bounding_box) return;
        const Rect& bbox = *s.bounding_box;
        if (!(bbox.x < (float)(tile.x + tile.width) && (bbox.x + bbox.width) > (float)tile.x &&
              bbox.y < (float)(tile.y + tile.height) && (bbox.y + bbox.height) > (float)tile.y))
            return;
        for (size_t triangle_index = 0; triangle_index < s.edges.size(); ++triangle_index) {
            const Edges& edges = s.edges[triangle_index];
            if (!edges.visible) continue;
            const Tri& tri = s.clipped_indices[triangle_index];
            const V4 v0 = s.projected_vertices[tri.i0], v1 = s.projected_vertices[tri.i1], v2 = s.projected_vertices[tri.i2];
            const V2 uv0 = s.clipped_uvs[tri.i0], uv1 = s.clipped_uvs[tri.i1], uv2 = s.clipped_uvs[tri.i2];

            float min_xf = rmin(v0.x, rmin(v1.x, v2.x)), max_xf = rmax(v0.x, rmax(v1.x, v2.x));
            float min_yf = rmin(v0.y, rmin(v1.y, v2.y)), max_yf = rmax(v0.y, rmax(v1.y, v2.y));
            size_t min_x = as_usize(rmax(std::floor(min_xf), (float)tile.x));
            size_t max_x = as_usize(rmin(std::ceil(max_xf), (float)(tile.x + tile.width)));
            size_t min_y = as_usize(rmax(std::floor(min_yf), (float)tile.y));
            size_t max_y = as_usize(rmin(std::ceil(max_yf), (float)(tile.y + tile.height)));

            for (size_t ty = min_y; ty < max_y; ++ty) {
                for (size_t tx = min_x; tx < max_x; ++tx) {
                    float p[2] = {(float)tx + 0.5f, (float)ty + 0.5f};
                    if (!edges_evaluate(edges, p[0], p[1])) continue;
                    {   // :1041-1047 wall geometry behind an opacity batch of the same profile is skipped
                        size_t idx = (ty - tile.y) * tile.width + (tx - tile.x);
                        if (surface_id[idx] != 0 && surface_id[idx] == profile) continue;
                    }
                    float w[3];
                    barycentric_weights(v0.x, v0.y, v1.x, v1.y, v2.x, v2.y, p[0], p[1], w);
                    float alpha = w[0], beta = w[1], gamma = w[2];
                    float one_over_z = 1.0f / v0.z * alpha + 1.0f / v1.z * beta + 1.0f / v2.z * gamma;
                    float z = 1.0f / one_over_z;
                    size_t zidx = (ty - tile.y) * tile.width + (tx - tile.x);
                    if (!(z < z_buffer[zidx])) continue;

                    float interpolated_u = (uv0.x / v0.w) * alpha + (uv1.x / v1.w) * beta + (uv2.x / v2.w) * gamma;
                    float interpolated_v = (uv0.y / v0.w) * alpha + (uv1.y / v1.w) * beta + (uv2.y / v2.w) * gamma;
                    float interpolated_reciprocal_w = (1.0f / v0.w) * alpha + (1.0f / v1.w) * beta + (1.0f / v2.w) * gamma;
                    interpolated_u /= interpolated_reciprocal_w;
                    interpolated_v /= interpolated_reciprocal_w;

                    V3 world = screen_to_world(p[0], p[1], z);

                    V3 normal;
                    if (batch.normals) {  // :1083-1099
                        V3 n0 = s.clipped_normals[tri.i0], n1 = s.clipped_normals[tri.i1], n2 = s.clipped_normals[tri.i2];
                        normal = normalized(n0 * alpha + n1 * beta + n2 * gamma);
                        V3 view_dir = normalized(camera_pos - world);
                        if (dot(normal, view_dir) < 0.0f) normal = -normal;
                    } else {
                        normal = {0, 0, 0};
                    }

                    uint8_t texel[4];
                    texel_3d(batch, interpolated_u, interpolated_v, world, texel);  // :1101-1222

                    if (g_vm_fresh_state) execution = Execution();
                    V4 color = pixel_to_vec4(texel);
                    const rxc_texture* baked = baked_texture(batch.shader, batch.chunk);
                    if (batch.shader >= 0 && baked) {  // :1227-1262: a baked shader texture replaces the texel
                        uint8_t t2[4];
                        texture_sample(*baked, interpolated_u, interpolated_v, f->sample_mode, batch.repeat_mode, t2);
                        color = pixel_to_vec4(t2);
                        color.x = srgb_to_linear_fast(color.x);
                        color.y = srgb_to_linear_fast(color.y);
                        color.z = srgb_to_linear_fast(color.z);
                        execution.color = {color.x, color.y, color.z};
                        execution.opacity.x = color.w;
                        execution.roughness.x = 0.5f;
                        execution.metallic.x = 0.0f;
                        execution.normal = normal;
                    } else {  // :1263-1317
                        color.x = srgb_to_linear_fast(color.x);
                        color.y = srgb_to_linear_fast(color.y);
                        color.z = srgb_to_linear_fast(color.z);
                        execution.color = {color.x, color.y, color.z};
                        execution.opacity.x = (float)texel[3] / 255.0f;
                        execution.normal = normal;
                        execution.roughness.x = 0.5f;
                        execution.metallic.x = 0.0f;
                        if (const VmProgram* program = program_for(batch.shader, batch.chunk)) {  // :1281-1300
                            execution.uv.x = interpolated_u / 4.0f;
                            execution.uv.y = interpolated_v / 4.0f;
                            execution.hitpoint = world;
                            execution.time = {f->time, f->time, f->time};
                            execution.reset(program->globals());
                            execution.shade(*program);
                        }
                    }

                    V3 mat_base = execution.color;  // :1319-1323
                    normal = normalized(execution.normal);
                    float mat_roughness = rclamp(execution.roughness.x, 0.0f, 1.0f);
                    float mat_metallic = rclamp(execution.metallic.x, 0.0f, 1.0f);
                    V3 mat_emissive = execution.emissive;

                    V3 lit = {0, 0, 0};
                    float occlusion = get_occlusion(batch.chunk, {world.x, world.z});  // :1327-1331
                    if (occlusion > 0.0f) {  // :1334-1365
                        if (f->has_ambient) {
                            float hemi = 0.5f * (normal.y + 1.0f);
                            V3 kd = mat_base * (1.0f - mat_metallic) * (1.0f - 0.04f);
                            lit = lit + V3{f->ambient[0], f->ambient[1], f->ambient[2]} * kd * hemi;
                        }
                        if (f->has_sun && f->day_factor > 0.0f) {  // :1342-1361 directional sun
                            V3 ldir = normalized(-V3{f->sun_dir[0], f->sun_dir[1], f->sun_dir[2]});
                            float r = rmax(f->day_factor, 0.0f);
                            lit = lit + shade_fast_brdf(mat_base, mat_roughness, mat_metallic, {0, 0, 0}, normal,
                                                        normalized(camera_pos - world), ldir, {r, r, r});
                        }
                        lit.x *= occlusion; lit.y *= occlusion; lit.z *= occlusion;
                    }
                    float hemi = 0.5f * (normal.y + 1.0f);  // :1368-1370
                    V3 kd = mat_base * (1.0f - mat_metallic) * (1.0f - 0.04f);
                    lit = lit + V3{batch.ambient_color[0], batch.ambient_color[1], batch.ambient_color[2]} * kd * hemi;

                    for (uint32_t li = 0; li < scene->n_lights; ++li) {  // :1373-1391
                        const rxc_light& light = scene->lights[li];
                        V3 radiance;
                        if (!light_radiance_at(light, world, normal, hash_anim, &radiance)) continue;
                        V3 ldir = normalized(lpos(light) - world);
                        lit = lit + shade_fast_brdf(mat_base, mat_roughness, mat_metallic, {0, 0, 0}, normal,
                                                    normalized(camera_pos - world), ldir, radiance);
                    }
                    lit = lit + mat_emissive;

                    color.x = linear_to_srgb_fast(lit.x);  // :1400-1404
                    color.y = linear_to_srgb_fast(lit.y);
                    color.z = linear_to_srgb_fast(lit.z);
                    color.w = execution.opacity.x;
                    vec4_to_pixel(color, texel);

                    if (texel[3] == 255) {  // :1408-1412
                        size_t idx = zidx * 4;
                        std::memcpy(&buffer[idx], texel, 4);
                        z_buffer[zidx] = z;
                        owner[zidx] = s.owner_base + (uint32_t)triangle_index;
                    }
                }
            }
        }
    }

    // src/rasterizer.rs:1425-1690: the opacity layer of a chunk.  Same coverage and depth arithmetic as
    // d3_rasterize against its own z-buffer; no lighting, no alpha test; always writes colour, z and the
    // batch's profile id into surface_id.
    void d3_rasterize_opacity(std::vector<uint8_t>& buffer, std::vector<float>& z_buffer, std::vector<uint64_t>& surface_id,
                              const TileRect& tile, const Batch3DState& s, Execution& execution) const {
        const rxc_batch3d& batch = *s.b;
        const uint64_t profile = batch.has_profile_id ? (uint64_t)batch.profile_id + 1u : 0u;
        if (!s.bounding_box) return;
        const Rect& bbox = *s.bounding_box;
        if (!(bbox.x < (float)(tile.x + tile.width) && (bbox.x + bbox.width) > (float)tile.x &&
              bbox.y < (float)(tile.y + tile.height) && (bbox.y + bbox.height) > (float)tile.y))
            return;
        for (size_t triangle_index = 0; triangle_index < s.edges.size(); ++triangle_index) {
            const Edges& edges = s.edges[triangle_index];
            if (!edges.visible) continue;
            const Tri& tri = s.clipped_indices[triangle_index];
            const V4 v0 = s.projected_vertices[tri.i0], v1 = s.projected_vertices[tri.i1], v2 = s.projected_vertices[tri.i2];
            const V2 uv0 = s.clipped_uvs[tri.i0], uv1 = s.clipped_uvs[tri.i1], uv2 = s.clipped_uvs[tri.i2];
            float min_xf = rmin(v0.x, rmin(v1.x, v2.x)), max_xf = rmax(v0.x, rmax(v1.x, v2.x));
            float min_yf = rmin(v0.y, rmin(v1.y, v2.y)), max_yf = rmax(v0.y, rmax(v1.y, v2.y));
            size_t min_x = as_usize(rmax(std::floor(min_xf), (float)tile.x));
            size_t max_x = as_usize(rmin(std::ceil(max_xf), (float)(tile.x + tile.width)));
            size_t min_y = as_usize(rmax(std::floor(min_yf), (float)tile.y));
            size_t max_y = as_usize(rmin(std::ceil(max_yf), (float)(tile.y + tile.height)));
            for (size_t ty = min_y; ty < max_y; ++ty) {
                for (size_t tx = min_x; tx < max_x; ++tx) {
                    float p[2] = {(float)tx + 0.5f, (float)ty + 0.5f};
                    if (!edges_evaluate(edges, p[0], p[1])) continue;
                    float w[3];
                    barycentric_weights(v0.x, v0.y, v1.x, v1.y, v2.x, v2.y, p[0], p[1], w);
                    float alpha = w[0], beta = w[1], gamma = w[2];
                    float one_over_z = 1.0f / v0.z * alpha + 1.0f / v1.z * beta + 1.0f / v2.z * gamma;
                    float z = 1.0f / one_over_z;
                    size_t zidx = (ty - tile.y) * tile.width + (tx - tile.x);
                    if (!(z < z_buffer[zidx])) continue;
                    float interpolated_u = (uv0.x / v0.w) * alpha + (uv1.x / v1.w) * beta + (uv2.x / v2.w) * gamma;
                    float interpolated_v = (uv0.y / v0.w) * alpha + (uv1.y / v1.w) * beta + (uv2.y / v2.w) * gamma;
                    float interpolated_reciprocal_w = (1.0f / v0.w) * alpha + (1.0f / v1.w) * beta + (1.0f / v2.w) * gamma;
                    interpolated_u /= interpolated_reciprocal_w;
                    interpolated_v /= interpolated_reciprocal_w;
                    V3 world = screen_to_world(p[0], p[1], z);
                    uint8_t texel[4];
                    texel_3d(batch, interpolated_u, interpolated_v, world, texel);  // :1512-1598
                    V4 color = pixel_to_vec4(texel);                                // :1600-1608
                    color.x = srgb_to_linear_fast(color.x);
                    color.y = srgb_to_linear_fast(color.y);
                    color.z = srgb_to_linear_fast(color.z);
                    if (g_vm_fresh_state) execution = Execution();
                    execution.color = {color.x, color.y, color.z};
                    execution.opacity.x = (float)texel[3] / 255.0f;
                    if (const VmProgram* program = program_for(batch.shader, batch.chunk)) {  // :1611-1637
                        execution.normal = {0, 0, 0};
                        execution.uv.x = interpolated_u / 4.0f;
                        execution.uv.y = interpolated_v / 4.0f;
                        execution.hitpoint = world;
                        execution.time = {f->time, f->time, f->time};
                        execution.roughness.x = 0.5f;
                        execution.metallic.x = 0.0f;
                        execution.reset(program->globals());
                        execution.shade(*program);
                    }
                    color.x = linear_to_srgb_fast(execution.color.x);  // :1639-1643
                    color.y = linear_to_srgb_fast(execution.color.y);
                    color.z = linear_to_srgb_fast(execution.color.z);
                    color.w = execution.opacity.x;
                    vec4_to_pixel(color, texel);
                    std::memcpy(&buffer[zidx * 4], texel, 4);  // :1647-1651
                    z_buffer[zidx] = z;
                    surface_id[zidx] = profile;
                }
            }
        }
    }

    // src/rasterizer.rs:1777-1821
    static void rasterize_line_bresenham(const V2& p0, const V2& p1, std::vector<uint8_t>& buffer, const TileRect& tile,
                                         const uint8_t color[4]) {
        int64_t x0 = as_isize(p0.x), y0 = as_isize(p0.y), x1 = as_isize(p1.x), y1 = as_isize(p1.y);
        int64_t dx = std::llabs(x1 - x0), dy = std::llabs(y1 - y0);
        int64_t sx = x0 < x1 ? 1 : -1, sy = y0 < y1 ? 1 : -1;
        int64_t err = dx - dy;
        int64_t x = x0, y = y0;
        while (x != x1 || y != y1) {
            size_t tx = (size_t)(x - (int64_t)tile.x);
            size_t ty = (size_t)(y - (int64_t)tile.y);
            if (tx < tile.width && ty < tile.height) {
                size_t idx = (ty * tile.width + tx) * 4;
                std::memcpy(&buffer[idx], color, 4);
            }
            int64_t e2 = err * 2;
            if (e2 > -dy) { err -= dy; x += sx; }
            if (e2 < dx) { err += dx; y += sy; }
        }
    }

    // src/rasterizer.rs:584-959
    void d2_rasterize(std::vector<uint8_t>& buffer, const TileRect& tile, const Batch2DState& s, Execution& execution) const {
        const rxc_batch2d& batch = *s.b;
        if (!s.bounding_box) return;
        const Rect& bbox = *s.bounding_box;
        const float pad = 0.5f;
        if (!(bbox.x < (float)(tile.x + tile.width) + pad && (bbox.x + bbox.width) > (float)tile.x - pad &&
              bbox.y < (float)(tile.y + tile.height) + pad && (bbox.y + bbox.height) > (float)tile.y - pad))
            return;
        const V2* uvs = (const V2*)batch.uvs;
        uint8_t line_color[4] = {255, 255, 255, 255};
        if (batch.source_kind == RXC_SRC_PIXEL) std::memcpy(line_color, batch.source_pixel, 4);
        switch (batch.mode) {
            case RXC_MODE_TRIANGLES: break;
            case RXC_MODE_LINES:  // :901-918
                for (uint32_t t = 0; t < batch.n_triangles; ++t) {
                    Tri tri = read_tri(batch.indices, batch.index_bytes, t);
                    rasterize_line_bresenham(s.projected_vertices[tri.i0], s.projected_vertices[tri.i1], buffer, tile, line_color);
                }
                return;
            case RXC_MODE_LINE_STRIP:  // :919-936 (the reference underflows on an empty vertex list)
                for (size_t i = 0; i + 1 < s.projected_vertices.size(); ++i)
                    rasterize_line_bresenham(s.projected_vertices[i], s.projected_vertices[i + 1], buffer, tile, line_color);
                return;
            default:  // :937-955
                for (size_t i = 0; i < s.projected_vertices.size(); ++i)
                    rasterize_line_bresenham(s.projected_vertices[i],
                                             s.projected_vertices[(i + 1) % s.projected_vertices.size()], buffer, tile,
                                             line_color);
                return;
        }
        for (size_t triangle_index = 0; triangle_index < s.edges.size(); ++triangle_index) {
            const Edges& edges = s.edges[triangle_index];
            Tri tri = read_tri(batch.indices, batch.index_bytes, triangle_index);
            V2 v0 = s.projected_vertices[tri.i0], v1 = s.projected_vertices[tri.i1], v2 = s.projected_vertices[tri.i2];
            V2 uv0 = uvs[tri.i0], uv1 = uvs[tri.i1], uv2 = uvs[tri.i2];
            float min_xf = rmin(v0.x, rmin(v1.x, v2.x)), max_xf = rmax(v0.x, rmax(v1.x, v2.x));
            float min_yf = rmin(v0.y, rmin(v1.y, v2.y)), max_yf = rmax(v0.y, rmax(v1.y, v2.y));
            size_t min_x = as_usize(rmax(std::floor(min_xf), (float)tile.x));
            size_t max_x = as_usize(rmin(std::ceil(max_xf), (float)(tile.x + tile.width)));
            size_t min_y = as_usize(rmax(std::floor(min_yf), (float)tile.y));
            size_t max_y = as_usize(rmin(std::ceil(max_yf), (float)(tile.y + tile.height)));
            for (size_t ty = min_y; ty < max_y; ++ty) {
                for (size_t tx = min_x; tx < max_x; ++tx) {
                    float p[2] = {(float)tx + 0.5f, (float)ty + 0.5f};
                    // :641-652 wrap: unreachable for tx in [tile.x, tile.x+tile.width), kept for fidelity
                    if (p[0] >= (float)(tile.x + tile.width)) p[0] -= (float)tile.width;
                    else if (p[0] < (float)tile.x) p[0] += (float)tile.width;
                    if (p[1] >= (float)(tile.y + tile.height)) p[1] -= (float)tile.height;
                    else if (p[1] < (float)tile.y) p[1] += (float)tile.height;

                    if (!(edges.visible && edges_evaluate(edges, p[0], p[1]))) continue;
                    float w[3];
                    barycentric_weights(v0.x, v0.y, v1.x, v1.y, v2.x, v2.y, p[0], p[1], w);
                    float u = uv0.x * w[0] + uv1.x * w[1] + uv2.x * w[2];
                    float v = uv0.y * w[0] + uv1.y * w[1] + uv2.y * w[2];

                    // :664-670
                    V2 gsp = {(float)tx - width / 2.0f - (translationd2.x - width / 2.0f),
                              (float)ty - height / 2.0f - (translationd2.y - height / 2.0f)};
                    V2 world = {gsp.x / scaled2, gsp.y / scaled2};

                    uint8_t texel[4] = {0, 0, 0, 0};
                    switch (batch.source_kind) {  // :672-758
                        case RXC_SRC_STATIC_TILE:
                        case RXC_SRC_DYNAMIC_TILE: {
                            const rxc_texture* t = tile_frame(batch.source_kind, batch.source_index);
                            if (t) texture_sample(*t, u, v, f->sample_mode, batch.repeat_mode, texel);
                            break;
                        }
                        case RXC_SRC_PIXEL: std::memcpy(texel, batch.source_pixel, 4); break;
                        case RXC_SRC_ENTITY_TILE:
                        case RXC_SRC_ITEM_TILE: {
                            const rxc_texture* t = tile_frame(batch.source_kind, batch.source_index);
                            if (t) texture_sample(*t, u, v, f->sample_mode, batch.repeat_mode, texel);
                            break;
                        }
                        case RXC_SRC_TERRAIN:
                            if (batch.chunk >= 0) sample_terrain_texture(batch.chunk, world, texel);
                            break;
                        default: break;
                    }

                    if (const VmProgram* program = program_for(batch.shader, batch.chunk)) {  // :760-797
                        if (g_vm_fresh_state) execution = Execution();
                        V4 color = pixel_to_vec4(texel);
                        execution.uv.x = u / 4.0f;
                        execution.uv.y = v / 4.0f;
                        execution.color = {color.x, color.y, color.z};
                        execution.hitpoint.x = world.x;
                        execution.hitpoint.y = world.y;
                        execution.time = {f->time, f->time, f->time};
                        execution.roughness.x = 0.5f;
                        execution.metallic.x = 0.0f;
                        execution.reset(program->globals());
                        execution.shade(*program);
                        color = {execution.color.x, execution.color.y, execution.color.z, 1.0f};
                        vec4_to_pixel(color, texel);
                    }

                    if ((batch.receives_light && scene->n_lights != 0) || f->has_ambient) {  // :799-873
                        float acc[3] = {0, 0, 0};
                        if (f->has_ambient) {
                            float occlusion = get_occlusion(batch.chunk, world);
                            acc[0] += f->ambient[0] * occlusion; acc[1] += f->ambient[1] * occlusion; acc[2] += f->ambient[2] * occlusion;
                        }
                        for (uint32_t li = 0; li < scene->n_lights; ++li) {
                            const rxc_light& light = scene->lights[li];
                            float lc[3];
                            if (!light_color_at(light, {world.x, 0.0f, world.y}, hash_anim, true, lc)) continue;
                            if (light.light_type == RXC_LIGHT_AMBIENT_DAYLIGHT) {
                                float occlusion = get_occlusion(batch.chunk, world);
                                lc[0] *= occlusion; lc[1] *= occlusion; lc[2] *= occlusion;
                            }
                            if (light.light_type != RXC_LIGHT_AMBIENT && light.light_type != RXC_LIGHT_AMBIENT_DAYLIGHT &&
                                !is_visible(world, {light.position[0], light.position[2]}))  // src/map/mini.rs:88-95
                                continue;
                            acc[0] += lc[0]; acc[1] += lc[1]; acc[2] += lc[2];
                        }
                        for (int i = 0; i < 3; ++i) {
                            acc[i] = rclamp(acc[i], 0.0f, 1.0f);
                            texel[i] = as_u8(rclamp(((float)texel[i] / 255.0f) * acc[i] * 255.0f, 0.0f, 255.0f));
                        }
                    }

                    size_t idx = ((ty - tile.y) * tile.width + (tx - tile.x)) * 4;  // :876-895
                    if (texel[3] == 255) {
                        std::memcpy(&buffer[idx], texel, 4);
                    } else {
                        float src_alpha = (float)texel[3] / 255.0f;
                        float dst_alpha = 1.0f - src_alpha;
                        for (int i = 0; i < 3; ++i)
                            buffer[idx + i] = as_u8(((float)texel[i] * src_alpha) + ((float)buffer[idx + i] * dst_alpha));
                        if (!f->preserve_transparency) buffer[idx + 3] = 255;
                        else buffer[idx + 3] = std::max(buffer[idx + 3], texel[3]);
                    }
                }
            }
        }
    }

    // the body of the per-tile closure, src/rasterizer.rs:275-556
    void render_tile(const TileRect& tile, std::vector<uint8_t>& buffer, std::vector<float>& z_buffer,
                     std::vector<uint32_t>& owner) const {
        buffer.assign(tile.width * tile.height * 4, 0);
        if (f->has_background_color)
            for (size_t i = 0; i < buffer.size(); i += 4) std::memcpy(&buffer[i], f->background_color, 4);
        std::vector<uint8_t> buffer_opacity(tile.width * tile.height * 4, 0);
        z_buffer.assign(tile.width * tile.height, 1.0f);
        std::vector<float> z_buffer_opacity(tile.width * tile.height, 1.0f);
        std::vector<uint64_t> surface_id(tile.width * tile.height, 0);  // :290 (0 = None)
        owner.assign(tile.width * tile.height, 0xFFFFFFFFu);

        if (!f->ignore_background_shader && f->background_shader != RXC_BG_NONE) {  // :292-308
            V2 screen_size = {width, height};
            for (size_t ty = 0; ty < tile.height; ++ty)
                for (size_t tx = 0; tx < tile.width; ++tx) {
                    V2 uv = {(float)(tile.x + tx) / screen_size.x, (float)(tile.y + ty) / screen_size.y};
                    uint8_t* px = &buffer[(ty * tile.width + tx) * 4];
                    if (f->background_shader == RXC_BG_VGRAY_GRADIENT) shade_vgray(uv.y, px);
                    else shade_grid(*f, uv, screen_size, px);
                }
        }

        Execution execution;  // :310

        if (f->d3_active) {
            // batches arrive in submission order: per chunk opacity, opaque, terrain (:314-357), then
            // static, dynamic, overlay (:360-405)
            for (const Batch3DState& s : *b3) {
                if (s.b->pass == RXC_PASS_CHUNK_OPACITY) d3_rasterize_opacity(buffer_opacity, z_buffer_opacity, surface_id, tile, s, execution);
                else d3_rasterize(buffer, z_buffer, owner, surface_id, tile, s, execution);
            }
            for (size_t i = 0; i < z_buffer.size(); ++i) {  // :409-461
                if (z_buffer[i] == 1.0f) {
                    V4 color = {0.0f, 0.0f, 0.0f, 1.0f};
                    const size_t tx = i % tile.width, ty = i / tile.width;
                    if (f->has_sky || f->has_brush_preview) {
                        V3 origin, dir;
                        screen_ray((float)(tile.x + tx), (float)(tile.y + ty), &origin, &dir);
                        if (f->has_sky) render_miss_sky(color, dir);
                        if (f->has_brush_preview) brush_preview(color, origin, dir);
                    }
                    vec4_to_pixel(color, &buffer[i * 4]);
                }
                if (z_buffer_opacity[i] < 1.0f && z_buffer[i] > z_buffer_opacity[i]) {  // :464-495
                    const size_t idx = i * 4;
                    float src_r = (float)buffer_opacity[idx], src_g = (float)buffer_opacity[idx + 1], src_b = (float)buffer_opacity[idx + 2];
                    float src_a = (float)buffer_opacity[idx + 3] / 255.0f;
                    float dst_r = (float)buffer[idx], dst_g = (float)buffer[idx + 1], dst_b = (float)buffer[idx + 2];
                    float dst_a = (float)buffer[idx + 3] / 255.0f;
                    float inv_a = 1.0f - src_a;
                    float out_r = src_r * src_a + dst_r * inv_a;
                    float out_g = src_g * src_a + dst_g * inv_a;
                    float out_b = src_b * src_a + dst_b * inv_a;
                    float out_a = !f->preserve_transparency ? 1.0f : rclamp(src_a + dst_a * inv_a, 0.0f, 1.0f);
                    buffer[idx] = as_u8(rclamp(out_r, 0.0f, 255.0f));
                    buffer[idx + 1] = as_u8(rclamp(out_g, 0.0f, 255.0f));
                    buffer[idx + 2] = as_u8(rclamp(out_b, 0.0f, 255.0f));
                    buffer[idx + 3] = as_u8(rclamp(out_a * 255.0f, 0.0f, 255.0f));
                }
            }
        }
        if (f->d2_active)
            for (const Batch2DState& s : *b2) d2_rasterize(buffer, tile, s, execution);  // :501-553
    }
};

uint32_t hash_u32(uint32_t seed) {  // src/rasterizer.rs:199-207
    uint32_t state = seed;
    state = (state ^ 61u) ^ (state >> 16);
    state = state + (state << 3);
    state ^= state >> 4;
    state = state * 0x27d4eb2du;
    state ^= state >> 15;
    return state;
}

int32_t validate(const rxc_tile* tiles, uint32_t n_tiles, const rxc_scene* scene, const rxc_frame* f) {
    if (!scene || !f || f->width == 0 || f->height == 0 || f->tile_size == 0) return RXC_ERR_INVALID;
    if (f->has_sky && f->sky_clouds) return RXC_ERR_UNSUPPORTED;  // noiselib's perlin_noise_2d is not available
    (void)tiles;
    for (uint32_t i = 0; i < scene->n_batches3d; ++i) {
        const rxc_batch3d& b = scene->batches3d[i];
        if (b.source_kind > RXC_SRC_TERRAIN) return RXC_ERR_INVALID;
        if (b.chunk >= (int32_t)scene->n_chunks) return RXC_ERR_INDEX;
        if (b.source_kind == RXC_SRC_TERRAIN && b.chunk >= 0 && scene->chunks[b.chunk].terrain_texture && scene->chunks[b.chunk].size == 0)
            return RXC_ERR_INDEX;  // the reference divides by chunk.size (src/chunk.rs:140)
        if (b.index_bytes != 4 && b.index_bytes != 8) return RXC_ERR_INVALID;
        if (b.source_kind == RXC_SRC_STATIC_TILE && (b.source_index >= n_tiles || tiles[b.source_index].n_textures == 0)) return RXC_ERR_INDEX;
        if (b.source_kind == RXC_SRC_DYNAMIC_TILE && (b.source_index >= scene->n_dynamic_textures || scene->dynamic_textures[b.source_index].n_textures == 0)) return RXC_ERR_INDEX;
        for (uint32_t t = 0; t < b.n_triangles; ++t) {
            Tri tri = read_tri(b.indices, b.index_bytes, t);
            if (tri.i0 >= b.n_vertices || tri.i1 >= b.n_vertices || tri.i2 >= b.n_vertices) return RXC_ERR_INDEX;
        }
    }
    for (uint32_t i = 0; i < scene->n_batches2d; ++i) {
        const rxc_batch2d& b = scene->batches2d[i];
        if (b.source_kind > RXC_SRC_TERRAIN) return RXC_ERR_INVALID;
        if (b.chunk >= (int32_t)scene->n_chunks) return RXC_ERR_INDEX;
        if (b.source_kind == RXC_SRC_TERRAIN && b.chunk >= 0 && scene->chunks[b.chunk].terrain_texture && scene->chunks[b.chunk].size == 0)
            return RXC_ERR_INDEX;
        if (b.index_bytes != 4 && b.index_bytes != 8) return RXC_ERR_INVALID;
        for (uint32_t t = 0; t < b.n_triangles; ++t) {
            Tri tri = read_tri(b.indices, b.index_bytes, t);
            if (b.mode == RXC_MODE_TRIANGLES && (tri.i0 >= b.n_vertices || tri.i1 >= b.n_vertices || tri.i2 >= b.n_vertices)) return RXC_ERR_INDEX;
            if (b.mode == RXC_MODE_LINES && (tri.i0 >= b.n_vertices || tri.i1 >= b.n_vertices)) return RXC_ERR_INDEX;
        }
    }
    return RXC_OK;
}

}  // namespace

extern "C" {

// Rasterizer::rasterize, src/rasterizer.rs:185-580.  n_threads <= 0: hardware concurrency
// (rayon's default global pool); tiles are handed out dynamically like rayon's work stealing.
// `pixels` gets width*height*4 bytes; `owner` / `depth` (optional) width*height entries.
int32_t rxo_rasterize(const rxc_tile* tiles, uint32_t n_tiles, const rxc_scene* scene, const rxc_mapmini* mapmini,
                      const rxc_frame* f, uint8_t* pixels, uint32_t* owner_out, float* depth_out, int32_t n_threads) {
    int32_t st = validate(tiles, n_tiles, scene, f);
    if (st != RXC_OK) return st;
    if (!pixels) return RXC_ERR_INVALID;

    Raster r;
    r.tiles = tiles; r.n_tiles = n_tiles; r.scene = scene; r.f = f; r.mapmini = mapmini;
    r.width = (float)f->width; r.height = (float)f->height;  // :194-195
    r.hash_anim = hash_u32((uint32_t)f->animation_frame);   // :208
    std::memcpy(r.inverse_view.m, f->inverse_view, sizeof(r.inverse_view.m));
    std::memcpy(r.inverse_projection.m, f->inverse_projection, sizeof(r.inverse_projection.m));
    r.camera_pos = {f->inverse_view[12], f->inverse_view[13], f->inverse_view[14]};  // :98-102
    r.translationd2 = {0.0f, 0.0f}; r.scaled2 = 1.0f;                                // :104-110
    if (f->has_matrix2d) { r.translationd2 = {f->matrix2d[6], f->matrix2d[7]}; r.scaled2 = f->matrix2d[0]; }

    // scene.project, src/scene.rs:154-200
    M4 view, proj;
    std::memcpy(view.m, f->view, sizeof(view.m));
    std::memcpy(proj.m, f->projection, sizeof(proj.m));
    std::vector<Batch3DState> b3(scene->n_batches3d);
    std::vector<Batch2DState> b2(scene->n_batches2d);
    uint32_t base = 0;
    for (uint32_t i = 0; i < scene->n_batches3d; ++i) {
        b3[i].b = &scene->batches3d[i];
        b3[i].owner_base = base;
        base += 3u * scene->batches3d[i].n_triangles;
        clip_and_project(b3[i], view, proj, r.width, r.height, f->matvec_mode);
    }
    for (uint32_t i = 0; i < scene->n_batches2d; ++i) {
        b2[i].b = &scene->batches2d[i];
        project2d(b2[i], f->has_matrix2d ? f->matrix2d : nullptr, f->matvec_mode);
    }
    r.b3 = &b3; r.b2 = &b2;

    const size_t width = f->width, height = f->height, tile_size = f->tile_size;
    std::vector<TileRect> tile_rects;  // :256-268
    for (size_t y = 0; y < height; y += tile_size)
        for (size_t x = 0; x < width; x += tile_size)
            tile_rects.push_back({x, y, std::min(tile_size, width - x), std::min(tile_size, height - y)});

    std::vector<std::vector<uint8_t>> tile_buffers(tile_rects.size());
    std::vector<std::vector<float>> tile_z(owner_out || depth_out ? tile_rects.size() : 0);
    std::vector<std::vector<uint32_t>> tile_owner(owner_out || depth_out ? tile_rects.size() : 0);

    unsigned nt = n_threads > 0 ? (unsigned)n_threads : std::max(1u, std::thread::hardware_concurrency());
    std::atomic<size_t> next{0};
    auto worker = [&]() {
        std::vector<float> z; std::vector<uint32_t> ow;
        for (;;) {
            size_t i = next.fetch_add(1);
            if (i >= tile_rects.size()) break;
            r.render_tile(tile_rects[i], tile_buffers[i], z, ow);
            if (!tile_z.empty()) { tile_z[i] = z; tile_owner[i] = ow; }
        }
    };
    if (nt == 1) worker();
    else {
        std::vector<std::thread> pool;
        for (unsigned t = 0; t < nt; ++t) pool.emplace_back(worker);
        for (auto& t : pool) t.join();
    }

    for (size_t i = 0; i < tile_rects.size(); ++i) {  // :560-579 serial compose
        const TileRect& tile = tile_rects[i];
        for (size_t row = 0; row < tile.height; ++row) {
            size_t dst = (tile.y + row) * width + tile.x;
            std::memcpy(pixels + dst * 4, &tile_buffers[i][row * tile.width * 4], tile.width * 4);
            if (owner_out) std::memcpy(owner_out + dst, &tile_owner[i][row * tile.width], tile.width * 4);
            if (depth_out) std::memcpy(depth_out + dst, &tile_z[i][row * tile.width], tile.width * 4);
        }
    }
    return RXC_OK;
}

// The programs, pattern banks and palette the next rxo_rasterize / rxo_vm_execute calls use.  trees[i] is
// Program.encode_tree() of program i (scene.shaders, then every chunk's, like rxc_scene.shaders).
int32_t rxo_set_programs(const uint32_t* const* trees, const uint32_t* n_words, uint32_t n_programs, const rxc_pattern* patterns,
                         uint32_t n_patterns, const rxc_pattern* patterns_normal, uint32_t n_patterns_normal, const float* palette,
                         uint32_t n_palette) {
    g_vm = VmBank{};
    g_vm.trees.resize(n_programs);
    g_vm.programs.resize(n_programs);
    for (uint32_t i = 0; i < n_programs; ++i) {
        g_vm.trees[i].assign(trees[i], trees[i] + n_words[i]);
        g_vm.programs[i].t = g_vm.trees[i].data();
    }
    auto copy_bank = [&](const rxc_pattern* src, uint32_t n, std::vector<rxc_pattern>& dst) {
        for (uint32_t i = 0; i < n; ++i) {
            g_vm.pattern_store.emplace_back(src[i].data, src[i].data + (size_t)src[i].width * src[i].height * 3);
            dst.push_back(rxc_pattern{nullptr, src[i].width, src[i].height});
        }
    };
    copy_bank(patterns, n_patterns, g_vm.patterns);
    copy_bank(patterns_normal, n_patterns_normal, g_vm.patterns_normal);
    size_t k = 0;
    for (auto& p : g_vm.patterns) p.data = g_vm.pattern_store[k++].data();
    for (auto& p : g_vm.patterns_normal) p.data = g_vm.pattern_store[k++].data();
    if (n_palette) g_vm.palette.assign(palette, palette + (size_t)n_palette * 4);
    return RXC_OK;
}

void rxo_set_vm_state_mode(int32_t per_fragment) { g_vm_fresh_state = per_fragment ? 1 : 0; }

// Execution::shade of program `program` on n records, same record layout as rxc_vm_execute (18 floats in,
// 24 floats out); every record starts from Execution::new.  Returns the number of records that faulted.
uint32_t rxo_vm_execute(uint32_t program, uint32_t n, const float* in, float* out) {
    if (program >= g_vm.programs.size()) return n;
    const VmProgram& P = g_vm.programs[program];
    uint32_t faults = 0;
    for (uint32_t i = 0; i < n; ++i) {
        const float* r = in + (size_t)i * 18;
        Execution e;
        e.uv = {r[0], r[1], r[2]}; e.color = {r[3], r[4], r[5]}; e.normal = {r[6], r[7], r[8]};
        e.hitpoint = {r[9], r[10], r[11]}; e.time = {r[12], r[13], r[14]}; e.opacity = {r[15], r[16], r[17]};
        if (P.has_shade()) { e.reset(P.globals()); e.shade(P); }
        if (e.fault) ++faults;
        const V3 v[8] = {e.uv, e.color, e.normal, e.roughness, e.metallic, e.emissive, e.opacity, e.bump};
        float* o = out + (size_t)i * 24;
        for (int k = 0; k < 8; ++k) { o[3 * k] = v[k].x; o[3 * k + 1] = v[k].y; o[3 * k + 2] = v[k].z; }
    }
    return faults;
}

// Stage output of Batch3D::clip_and_project for one batch (setup-kernel parity tests).
// Arrays are caller allocated; capacities: projected 4*(n_vertices + 4*n_triangles) floats,
// clipped_indices 3*(3*n_triangles) u32, edges 9*(3*n_triangles) floats (a0..2,b0..2,c0..2),
// visible 3*n_triangles bytes, bbox 4 floats.  Returns counts through the out params.
int32_t rxo_clip_and_project(const rxc_batch3d* batch, const rxc_frame* f, float* projected, uint32_t* n_projected,
                             uint32_t* clipped_indices, float* edges, uint8_t* visible, uint32_t* n_clipped,
                             float* bbox, uint32_t* has_bbox) {
    Batch3DState s;
    s.b = batch;
    M4 view, proj;
    std::memcpy(view.m, f->view, sizeof(view.m));
    std::memcpy(proj.m, f->projection, sizeof(proj.m));
    clip_and_project(s, view, proj, (float)f->width, (float)f->height, f->matvec_mode);
    *n_projected = (uint32_t)s.projected_vertices.size();
    for (size_t i = 0; i < s.projected_vertices.size(); ++i) std::memcpy(projected + i * 4, &s.projected_vertices[i], 16);
    *n_clipped = (uint32_t)s.edges.size();
    for (size_t i = 0; i < s.edges.size(); ++i) {
        clipped_indices[i * 3 + 0] = (uint32_t)s.clipped_indices[i].i0;
        clipped_indices[i * 3 + 1] = (uint32_t)s.clipped_indices[i].i1;
        clipped_indices[i * 3 + 2] = (uint32_t)s.clipped_indices[i].i2;
        std::memcpy(edges + i * 9 + 0, s.edges[i].a, 12);
        std::memcpy(edges + i * 9 + 3, s.edges[i].b, 12);
        std::memcpy(edges + i * 9 + 6, s.edges[i].c, 12);
        visible[i] = s.edges[i].visible ? 1 : 0;
    }
    *has_bbox = s.bounding_box ? 1 : 0;
    if (s.bounding_box) { bbox[0] = s.bounding_box->x; bbox[1] = s.bounding_box->y; bbox[2] = s.bounding_box->width; bbox[3] = s.bounding_box->height; }
    return RXC_OK;
}

// The attributes clip_and_project leaves next to projected_vertices (batch3d.rs:602-607, :648-681): clipped_uvs and
// clipped_normals, n_projected entries each (normals: zeros when the batch has none -- the reference would have panicked).
// Call after rxo_clip_and_project with the same arguments; capacities as there.
int32_t rxo_clip_and_project_attrs(const rxc_batch3d* batch, const rxc_frame* f, float* clipped_uvs, float* clipped_normals) {
    Batch3DState s;
    s.b = batch;
    M4 view, proj;
    std::memcpy(view.m, f->view, sizeof(view.m));
    std::memcpy(proj.m, f->projection, sizeof(proj.m));
    clip_and_project(s, view, proj, (float)f->width, (float)f->height, f->matvec_mode);
    for (size_t i = 0; i < s.projected_vertices.size(); ++i) {
        V2 uv = i < s.clipped_uvs.size() ? s.clipped_uvs[i] : V2{0.0f, 0.0f};
        V3 n = i < s.clipped_normals.size() ? s.clipped_normals[i] : V3{0.0f, 0.0f, 0.0f};
        clipped_uvs[i * 2] = uv.x; clipped_uvs[i * 2 + 1] = uv.y;
        clipped_normals[i * 3] = n.x; clipped_normals[i * 3 + 1] = n.y; clipped_normals[i * 3 + 2] = n.z;
    }
    return RXC_OK;
}

// ---- known-answer-test hooks: each exposes one reference function unchanged ---------------------
void rxo_edges_new(const float v0[6], const float v1[6], float out_abc[9]) {
    float a[3][2], b[3][2];
    std::memcpy(a, v0, 24); std::memcpy(b, v1, 24);
    Edges e = edges_new(a, b, true);
    std::memcpy(out_abc, e.a, 12); std::memcpy(out_abc + 3, e.b, 12); std::memcpy(out_abc + 6, e.c, 12);
}
int32_t rxo_edges_evaluate(const float abc[9], float px, float py) {
    Edges e; std::memcpy(e.a, abc, 12); std::memcpy(e.b, abc + 3, 12); std::memcpy(e.c, abc + 6, 12); e.visible = true;
    return edges_evaluate(e, px, py) ? 1 : 0;
}
void rxo_texture_sample(const rxc_texture* t, float u, float v, uint32_t sample_mode, uint32_t repeat_mode, uint8_t out[4]) {
    texture_sample(*t, u, v, sample_mode, repeat_mode, out);
}
void rxo_vec4_to_pixel(const float v[4], uint8_t out[4]) { vec4_to_pixel({v[0], v[1], v[2], v[3]}, out); }
void rxo_pixel_to_vec4(const uint8_t p[4], float out[4]) { V4 v = pixel_to_vec4(p); std::memcpy(out, &v, 16); }
int32_t rxo_light_color_at(const rxc_light* l, const float point[3], uint32_t hash, int32_t d2, float out[3]) {
    return light_color_at(*l, {point[0], point[1], point[2]}, hash, d2 != 0, out) ? 1 : 0;
}
int32_t rxo_light_radiance_at(const rxc_light* l, const float point[3], const float normal[3], uint32_t hash, float out[3]) {
    V3 r;
    if (!light_radiance_at(*l, {point[0], point[1], point[2]}, {normal[0], normal[1], normal[2]}, hash, &r)) return 0;
    out[0] = r.x; out[1] = r.y; out[2] = r.z;
    return 1;
}
uint32_t rxo_hash_u32(uint32_t seed) { return hash_u32(seed); }
void rxo_barycentric(const float a[2], const float b[2], const float c[2], const float p[2], float w[3]) {
    Raster::barycentric_weights(a[0], a[1], b[0], b[1], c[0], c[1], p[0], p[1], w);
}
void rxo_shade_background(const rxc_frame* f, float u, float v, uint8_t out[4]) {
    if (f->background_shader == RXC_BG_VGRAY_GRADIENT) shade_vgray(v, out);
    else shade_grid(*f, {u, v}, {(float)f->width, (float)f->height}, out);
}
void rxo_shade_fast_brdf(const float base[3], float roughness, float metallic, const float n[3], const float v[3],
                         const float l[3], const float radiance[3], float out[3]) {
    V3 r = Raster::shade_fast_brdf({base[0], base[1], base[2]}, roughness, metallic, {0, 0, 0}, {n[0], n[1], n[2]},
                                   {v[0], v[1], v[2]}, {l[0], l[1], l[2]}, {radiance[0], radiance[1], radiance[2]});
    out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
void rxo_mat4_mul_vec4(const float m[16], const float v[4], uint32_t mode, float out[4]) {
    M4 M; std::memcpy(M.m, m, 64);
    V4 r = mat4_mul_vec4(M, {v[0], v[1], v[2], v[3]}, mode);
    std::memcpy(out, &r, 16);
}
void rxo_mat4_mul_mat4(const float a[16], const float b[16], uint32_t mode, float out[16]) {
    M4 A, B; std::memcpy(A.m, a, 64); std::memcpy(B.m, b, 64);
    M4 R = mat4_mul_mat4(A, B, mode);
    std::memcpy(out, R.m, 64);
}
uint32_t rxo_hardware_threads(void) { return std::max(1u, std::thread::hardware_concurrency()); }

}  // extern "C"
