// rx_kernels.cu -- the sm_100a kernels of the rasterize path.
//
//   k_frame_setup     per (frame, batch) matrices, frustum-AABB reject, texture frame, shade descriptor (source,
//                     repeat flags, VM program), light flicker, 2D batch projection + records, counter zeroing
//   k_tri_setup       per original triangle: view transform, early cull, near classification,
//                     projection, cull/swap, edge equations, depth/uv reciprocals, pixel bbox, flat-normal flag;
//                     batch screen bbox by block reduction + atomics; near-clip counting + block scan
//   k_batch_finalize  per batch: scan of the per-chunk clip counts, API-tile scissor from the bbox
//   k_clip_emit       per near-clipped triangle: Sutherland-Hodgman, fan triangles at ordered slots
//   k_bin_count       final (scissored) bbox, tile counts, large-triangle list
//   k_tile_alloc      per tile: list space from an atomic arena cursor
//   k_bin_fill        tile lists
//   k_front_small     all of the above back to back in one CTA per frame (tiny scenes)
//   k_front_cluster   all of the above in one thread-block cluster per frame, cluster barriers between the phases
//   k_bin_large, k_bin2d, k_list_sort   general mode: every triangle and every 2D record in per-tile lists sorted
//                     by submission ordinal (chunk opacity layer, surface ids, many 2D records)
//   k_raster          persistent, one CTA per 32x32 tile at a time, 2x2 pixels per thread: warp-private walk of the
//                     triangle records, per-pixel exact edge/depth test, alpha test, deferred shading of the
//                     owner (or the Rusteria VM program of its batch, rx_vm.cuh), miss pass (sky, brush preview),
//                     opacity blend, 2D pass, 128-bit RGBA8 stores
//   k_vm_execute, k_selftest_div   diagnostics
//
// Reference cites are to /root/reference (markusmoenig/Rusterix).
#ifndef __CUDACC_RTC__
#include <set>
#endif

#include "rx_kernels.cuh"
#include "rx_vm.cuh"

#include <math_constants.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

namespace {

// Programmatic dependent launch (sm_90+): every kernel of the frame's chain lets its successor become resident at once
// (`launch_dependents`: the next grid's CTAs are placed as soon as every CTA of this one has started) and touches global memory
// only after `wait` (all prerequisite grids complete, their writes visible) -- the drain / launch / ramp between two dependent
// short kernels shrinks to the release of CTAs that are already there.  Both are no-ops for a launch without the
// programmatic-serialization attribute (rxk_* below set it unless RXC_PDL=0).
__device__ __forceinline__ void pdl_enter() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}

__device__ __forceinline__ uint32_t lanemask_lt() {
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// The reference rejects a batch per API tile (rasterizer.rs:978-983 / :594-600).  Over one axis the
// accepted tiles form a contiguous range; return the pixel range [o0, o1) they cover.
__device__ void scissor_1d(float bx, float bw, int dim, int ts, float pad, int* o0, int* o1) {
    const int nt = (dim + ts - 1) / ts;
    const float far_edge = bx + bw;
    int lo = 0, hi = nt;
    while (lo < hi) {  // first tile with  bx < (tile.x + tile.w) as f32 (+ pad)
        int mid = (lo + hi) >> 1;
        long long e = (long long)(mid + 1) * ts;
        float h = (float)(e < dim ? (int)e : dim) + pad;
        if (bx < h) hi = mid; else lo = mid + 1;
    }
    const int kA = lo;
    lo = 0; hi = nt;
    while (lo < hi) {  // first tile where  bx + bw > tile.x as f32 (- pad)  fails
        int mid = (lo + hi) >> 1;
        float l = (float)((long long)mid * ts) - pad;
        if (far_edge > l) lo = mid + 1; else hi = mid;
    }
    const int kB = lo - 1;
    if (kA > kB) { *o0 = 0; *o1 = 0; return; }
    long long e = (long long)(kB + 1) * ts;
    *o0 = kA * ts;
    *o1 = (int)(e < dim ? e : dim);
}

// DTile of a tile source: static tiles, then dynamic tiles, then the host-resolved entity/item tiles.
// Returns false when the source has no tile (EntityTile/ItemTile not found, index out of range).
__device__ __forceinline__ bool source_tile(const SceneDev& S, uint32_t kind, uint32_t index, DTile* out) {
    uint32_t base, n;
    if (kind == RXC_SRC_STATIC_TILE) { base = 0u; n = S.n_static_tiles; }
    else if (kind == RXC_SRC_DYNAMIC_TILE) { base = S.n_static_tiles; n = S.n_dynamic_tiles; }
    else { base = S.n_static_tiles + S.n_dynamic_tiles; n = S.n_actor_tiles; }
    if (index >= n) return false;
    *out = S.tiles[base + index];
    return out->n_frames != 0u;
}
__device__ __forceinline__ bool is_tile_source(uint32_t kind) {
    return kind == RXC_SRC_STATIC_TILE || kind == RXC_SRC_DYNAMIC_TILE || kind == RXC_SRC_ENTITY_TILE || kind == RXC_SRC_ITEM_TILE;
}

__device__ __forceinline__ void edge_eq(float x0, float y0, float x1, float y1, float* a, float* b, float* c) {
    *a = y1 - y0;               // edge.rs:18
    *b = x0 - x1;               // edge.rs:19
    *c = x1 * y0 - y1 * x0;     // edge.rs:20
}

// pixel range visited for a primitive spanning [mn, mx] on one axis (rasterizer.rs:1014-1017, union over tiles)
__device__ __forceinline__ void pixel_range(float mn, float mx, int dim, int* lo, int* hi) {
    *lo = (mn == mn) ? rx_sat_int(floorf(mn), dim) : 0;
    *hi = (mx == mx) ? rx_sat_int(ceilf(mx), dim) : dim;
}

struct ClipVert { f4 p; float u, v; f3 n; };

// batch3d.rs:626-669.  Returns the number of polygon vertices (0, 3 or 4).
__device__ int clip_polygon(const f4 vv[3], const float2 uv[3], const f3 nn[3], ClipVert out[4]) {
    int n = 0;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const int j = (i + 1) % 3;
        const f4 cur = vv[i], nxt = vv[j];
        const bool cin = cur.z < -RX_NEAR_PLANE, nin = nxt.z < -RX_NEAR_PLANE;
        if (cin) {
            out[n].p = cur; out[n].u = uv[i].x; out[n].v = uv[i].y; out[n].n = nn[i];
            ++n;
        }
        if (cin != nin) {
            float t = (-RX_NEAR_PLANE - cur.z) / (nxt.z - cur.z);
            out[n].p = {cur.x + t * (nxt.x - cur.x), cur.y + t * (nxt.y - cur.y), cur.z + t * (nxt.z - cur.z),
                        cur.w + t * (nxt.w - cur.w)};
            out[n].u = uv[i].x + t * (uv[j].x - uv[i].x);
            out[n].v = uv[i].y + t * (uv[j].y - uv[i].y);
            f3 a = rx_scale3(nn[i], 1.0f - t), b = rx_scale3(nn[j], t);
            out[n].n = rx_normalize3(rx_add3(a, b));
            ++n;
        }
    }
    return n;
}

// The records of a visible triangle from its unswapped projected vertices and its final edge equations
// (rasterizer.rs:989-1076: pixel range, barycentric constants, 1/z, uv/w, 1/w).  Returns false when no pixel centre
// can be visited.
__device__ bool tri_records(const f4 P[3], const float2 uv[3], const f3 nn[3], const float ea[3], const float eb[3], const float ec[3],
                            int W, int H, uint32_t meta, TriVis* tv, TriShade* ts, uint32_t* bbx, uint32_t* bby) {
    int x0, x1, y0, y1;
    pixel_range(fminf(P[0].x, fminf(P[1].x, P[2].x)), fmaxf(P[0].x, fmaxf(P[1].x, P[2].x)), W, &x0, &x1);
    pixel_range(fminf(P[0].y, fminf(P[1].y, P[2].y)), fmaxf(P[0].y, fmaxf(P[1].y, P[2].y)), H, &y0, &y1);
    if (x0 >= x1 || y0 >= y1) return false;  // no pixel centre can be visited
    *bbx = (uint32_t)x0 | ((uint32_t)x1 << 16);
    *bby = (uint32_t)y0 | ((uint32_t)y1 << 16);

    TriVis r;
#pragma unroll
    for (int i = 0; i < 3; ++i) { r.ea[i] = ea[i]; r.eb[i] = eb[i]; r.ec[i] = ec[i]; }
    r.ax = P[0].x; r.ay = P[0].y; r.bx = P[1].x; r.by = P[1].y; r.cx = P[2].x; r.cy = P[2].y;
    const float acx = P[2].x - P[0].x, acy = P[2].y - P[0].y;
    const float abx = P[1].x - P[0].x, aby = P[1].y - P[0].y;
    r.area = acx * aby - acy * abx;  // rasterizer.rs:1767
    r.rarea = 1.0f / r.area; r.spare = 0.0f;
    r.iz0 = 1.0f / P[0].z; r.iz1 = 1.0f / P[1].z; r.iz2 = 1.0f / P[2].z;  // rasterizer.rs:1054-1055
    // The per-pixel barycentric divisions by `area` may use rx_div_by when nothing can leave the range
    // in which the residual corrections are exact: finite coordinates up to 2^30 (numerators <= 2^62;
    // pixel centres are k+0.5 < 2^14, so a non-zero (vertex - centre) is >= 2^-26), ac components zero
    // or >= 2^-20 (non-zero numerators >= 2^-76) and |area| in [2^-20, 2^44].
    {
        const float big = 1073741824.0f, tiny = 9.5367431640625e-07f;
        bool ok = fabsf(r.ax) <= big && fabsf(r.ay) <= big && fabsf(r.bx) <= big && fabsf(r.by) <= big && fabsf(r.cx) <= big &&
                  fabsf(r.cy) <= big;  // false for NaN / Inf
        ok = ok && (acx == 0.0f || fabsf(acx) >= tiny) && (acy == 0.0f || fabsf(acy) >= tiny);
        ok = ok && fabsf(r.area) >= tiny && fabsf(r.area) <= 17592186044416.0f;
        if (ok) meta |= RX_META_FASTDIV;
    }
    r.bbx = *bbx; r.bby = *bby; r.meta = meta;
    *tv = r;

    TriShade s;
    s.uw0 = uv[0].x / P[0].w; s.vw0 = uv[0].y / P[0].w;  // rasterizer.rs:1062-1072
    s.uw1 = uv[1].x / P[1].w; s.vw1 = uv[1].y / P[1].w;
    s.uw2 = uv[2].x / P[2].w; s.vw2 = uv[2].y / P[2].w;
    s.rw0 = 1.0f / P[0].w; s.rw1 = 1.0f / P[1].w; s.rw2 = 1.0f / P[2].w;
    s.n0x = nn[0].x; s.n0y = nn[0].y; s.n0z = nn[0].z;
    s.n1x = nn[1].x; s.n1y = nn[1].y; s.n1z = nn[1].z;
    s.n2x = nn[2].x; s.n2y = nn[2].y; s.n2z = nn[2].z;
    s.pad0 = 0.0f; s.pad1 = 0.0f;
    // flat triangle: the blend n*alpha + n*beta + n*gamma normalises to n/|n| (rasterizer.rs:1083-1092); the
    // shade reads the unit normal from n0 and skips the interpolation
    if (nn[0].x == nn[1].x && nn[0].y == nn[1].y && nn[0].z == nn[1].z && nn[0].x == nn[2].x && nn[0].y == nn[2].y && nn[0].z == nn[2].z) {
        const f3 u = rx_normalize3(nn[0]);
        s.n0x = u.x; s.n0y = u.y; s.n0z = u.z;
        s.pad0 = __uint_as_float(1u);
    }
    *ts = s;
    return true;
}

// batch3d.rs:706-739 + the per-triangle constants of rasterizer.rs:989-1076.
// Returns visibility; fills the records and the raw pixel bbox when visible.
__device__ bool make_tri(const f4 P[3], const float2 uv[3], const f3 nn[3], uint32_t cull_mode, bool edge_vis, int W, int H,
                         uint32_t meta, TriVis* tv, TriShade* ts, uint32_t* bbx, uint32_t* bby) {
    const f4 v0 = P[0];
    f4 v1 = P[1], v2 = P[2];
    const float orientation = (v1.x - v0.x) * (v2.y - v0.y) - (v1.y - v0.y) * (v2.x - v0.x);  // batch3d.rs:743-746
    const bool front = orientation > 0.0f;
    bool visible;
    bool swap = false;
    if (cull_mode == RXC_CULL_OFF) { swap = front; visible = true; }
    else if (cull_mode == RXC_CULL_FRONT) { visible = !front; }
    else { swap = front; visible = front; }
    visible = visible && edge_vis;
    if (!visible) return false;
    if (swap) { f4 t = v1; v1 = v2; v2 = t; }

    float ea[3], eb[3], ec[3];
    edge_eq(v0.x, v0.y, v1.x, v1.y, &ea[0], &eb[0], &ec[0]);
    edge_eq(v1.x, v1.y, v2.x, v2.y, &ea[1], &eb[1], &ec[1]);
    edge_eq(v2.x, v2.y, v0.x, v0.y, &ea[2], &eb[2], &ec[2]);
    return tri_records(P, uv, nn, ea, eb, ec, W, H, meta, tv, ts, bbx, bby);
}

// record flags of every triangle of a batch: the opacity layer never alpha-tests (rasterizer.rs:1647-1651)
__device__ __forceinline__ uint32_t tri_meta_flags(const DFrameBatch& FB) {
    if (FB.sd_flags & RX_SD_OPACITY) return RX_META_OPACITY;
    return FB.alpha_test ? RX_META_ALPHA : 0u;
}

// virtual block coordinates: the front-end bodies run either as their own kernels (one CTA per block) or
// back to back inside k_front_small (one CTA per frame walks the blocks)
struct Blk { uint32_t x, y, nx; };

struct MinMax {
    float mnx, mxx, mny, mxy;
    __device__ void init() { mnx = CUDART_INF_F; mxx = -CUDART_INF_F; mny = CUDART_INF_F; mxy = -CUDART_INF_F; }
    __device__ void add(float x, float y) {  // f32::min / f32::max ignore NaN (batch3d.rs:755-760)
        mnx = fminf(mnx, x); mxx = fmaxf(mxx, x); mny = fminf(mny, y); mxy = fmaxf(mxy, y);
    }
};

__device__ void block_minmax_to_keys(MinMax m, uint32_t* kminx, uint32_t* kmaxx, uint32_t* kminy, uint32_t* kmaxy) {
    __shared__ float s_red[4][32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        m.mnx = fminf(m.mnx, __shfl_xor_sync(0xFFFFFFFFu, m.mnx, o));
        m.mxx = fmaxf(m.mxx, __shfl_xor_sync(0xFFFFFFFFu, m.mxx, o));
        m.mny = fminf(m.mny, __shfl_xor_sync(0xFFFFFFFFu, m.mny, o));
        m.mxy = fmaxf(m.mxy, __shfl_xor_sync(0xFFFFFFFFu, m.mxy, o));
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
    if (lane == 0) { s_red[0][warp] = m.mnx; s_red[1][warp] = m.mxx; s_red[2][warp] = m.mny; s_red[3][warp] = m.mxy; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < nw; ++w) {
            m.mnx = fminf(m.mnx, s_red[0][w]); m.mxx = fmaxf(m.mxx, s_red[1][w]);
            m.mny = fminf(m.mny, s_red[2][w]); m.mxy = fmaxf(m.mxy, s_red[3][w]);
        }
        atomicMin(kminx, rx_float_key(m.mnx)); atomicMax(kmaxx, rx_float_key(m.mxx));
        atomicMin(kminy, rx_float_key(m.mny)); atomicMax(kmaxy, rx_float_key(m.mxy));
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// k_frame_setup
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void d_frame_setup(const SceneDev& S, const Workspace& Wk, uint32_t tiles_per_frame, Blk bk) {
    const uint32_t f = bk.y;
    const DFrame& F = Wk.frames[f];
    const uint32_t tid = threadIdx.x;

    if (bk.x == 0) {
        if (tid == 0) {
            DCounters z = {};
            Wk.counters[f] = z;
        }
        if (f == 0 && tid < RX_RASTER_COUNTERS) Wk.raster_counter[tid] = 0u;
        // light flicker is a per-frame constant of the light (light.rs:656-672)
        for (uint32_t i = tid; i < S.n_lights; i += blockDim.x) {
            DLight l = S.lights[i];
            float factor = 1.0f;
            if (l.flicker_factor /* holds `flicker` in the scene copy */ > 0.0f) {
                uint32_t s = (rx_as_u32(l.px) + rx_as_u32(l.py) + rx_as_u32(l.pz)) * 100u;
                uint32_t combined = F.hash_anim + s;
                float fv = rx_clamp((float)combined / 4294967296.0f, 0.0f, 1.0f);
                factor = 1.0f - fv * l.flicker_factor;
            }
            l.flicker_factor = factor;
            l.inv_range = 1.0f / (l.start_distance - l.end_distance);
            // early range cull of the deferred shade: radiance_at returns None from end_distance on (light.rs:536-653) for
            // every positional type; the slack keeps the decision with the `dist >= end_distance` test behind it
            if (!l.emitting) l.range2 = 0.0f;
            else if (l.light_type == RXC_LIGHT_AMBIENT || l.light_type == RXC_LIGHT_AMBIENT_DAYLIGHT) l.range2 = CUDART_INF_F;
            else l.range2 = l.end_distance > 0.0f ? l.end_distance * l.end_distance * 1.00001f : (l.end_distance <= 0.0f ? 0.0f : CUDART_INF_F);  // NaN: no cull
            // spot lights: the per-frame copy carries cos(cone_angle) in `width` (an area light's field) for the deferred shade's
            // cone test; < 0 rejects every direction, > pi none, NaN none (light.rs:566-569 compares acos(..) > cone_angle)
            if (l.light_type == RXC_LIGHT_SPOT)
                l.width = l.cone_angle < 0.0f ? CUDART_INF_F : l.cone_angle > 3.14159274f ? -CUDART_INF_F : cosf(l.cone_angle);
            Wk.lights[(size_t)f * Wk.lights_stride + i] = l;
        }
        return;
    }

    if (bk.x <= S.n_b2) {  // one CTA per 2D batch: batch2d.rs:373-425
        const uint32_t b = bk.x - 1;
        const DBatch2& B = S.b2[b];
        Tri2D* recs = Wk.tri2d + (size_t)f * Wk.tri2d_stride + B.rec_off;
        const bool active = F.d2_active != 0;
        MinMax mm; mm.init();
        if (active) {
            for (uint32_t i = tid; i < B.n_verts; i += blockDim.x) {
                float2 p = S.pos2[B.v_off + i];
                if (F.has_mat2d) { f3 r = rx_matvec3(F.mat2d, {p.x, p.y, 1.0f}, F.matvec_mode); p.x = r.x; p.y = r.y; }
                mm.add(p.x, p.y);
            }
        }
        __shared__ uint32_t s_key[4];
        if (tid == 0) {
            s_key[0] = rx_float_key(CUDART_INF_F); s_key[1] = rx_float_key(-CUDART_INF_F);
            s_key[2] = rx_float_key(CUDART_INF_F); s_key[3] = rx_float_key(-CUDART_INF_F);
        }
        __syncthreads();
        block_minmax_to_keys(mm, &s_key[0], &s_key[1], &s_key[2], &s_key[3]);
        const float mnx = rx_key_float(s_key[0]), mxx = rx_key_float(s_key[1]);
        const float mny = rx_key_float(s_key[2]), mxy = rx_key_float(s_key[3]);
        int sx0, sx1, sy0, sy1;
        scissor_1d(mnx, mxx - mnx, F.width, (int)F.tile_size, 0.5f, &sx0, &sx1);   // rasterizer.rs:594-600
        scissor_1d(mny, mxy - mny, F.height, (int)F.tile_size, 0.5f, &sy0, &sy1);
        sy0 = max(sy0, F.band_y0); sy1 = min(sy1, F.band_y1);
        sx0 = max(sx0, F.band_x0); sx1 = min(sx1, F.band_x1);
        if (tid == 0) {
            DFrameBatch2 fb2;
            fb2.tex = 0xFFFFFFFFu; fb2.terrain = 0u; fb2.program = B.program;
            if (is_tile_source(B.source_kind)) {
                DTile t;  // rasterizer.rs:674-733: a missing tile samples as transparent
                if (source_tile(S, B.source_kind, B.source_index, &t)) fb2.tex = t.first + (uint32_t)(F.animation_frame % t.n_frames);
            } else if (B.source_kind == RXC_SRC_TERRAIN && B.chunk >= 0) {  // :746-752
                fb2.tex = S.chunk_info[B.chunk].terrain_tex;
                fb2.terrain = 1u;
            }
            fb2.lit = ((B.receives_light && S.n_lights != 0) || F.has_ambient) ? 1u : 0u;  // rasterizer.rs:799-802
            Wk.fb2[(size_t)f * Wk.fb2_stride + b] = fb2;
        }
        for (uint32_t t = tid; t < B.n_recs; t += blockDim.x) {
            Tri2D r = {};
            r.batch = b; r.kind = B.mode == RXC_MODE_TRIANGLES ? 0u : 1u;
            if (active && sx0 < sx1 && sy0 < sy1) {
                if (B.mode == RXC_MODE_TRIANGLES) {
                    const uint32_t i0 = S.idx2[(size_t)(B.t_off + t) * 3 + 0], i1 = S.idx2[(size_t)(B.t_off + t) * 3 + 1],
                                   i2 = S.idx2[(size_t)(B.t_off + t) * 3 + 2];
                    float2 p[3] = {S.pos2[i0], S.pos2[i1], S.pos2[i2]};
                    if (F.has_mat2d) {
                        for (int k = 0; k < 3; ++k) { f3 q = rx_matvec3(F.mat2d, {p[k].x, p[k].y, 1.0f}, F.matvec_mode); p[k].x = q.x; p[k].y = q.y; }
                    }
                    const float2 t0 = S.uv2[i0], t1 = S.uv2[i1], t2 = S.uv2[i2];
                    r.ax = p[0].x; r.ay = p[0].y; r.bx = p[1].x; r.by = p[1].y; r.cx = p[2].x; r.cy = p[2].y;
                    r.u0 = t0.x; r.v0 = t0.y; r.u1 = t1.x; r.v1 = t1.y; r.u2 = t2.x; r.v2 = t2.y;
                    edge_eq(p[0].x, p[0].y, p[1].x, p[1].y, &r.ea[0], &r.eb[0], &r.ec[0]);
                    edge_eq(p[1].x, p[1].y, p[2].x, p[2].y, &r.ea[1], &r.eb[1], &r.ec[1]);
                    edge_eq(p[2].x, p[2].y, p[0].x, p[0].y, &r.ea[2], &r.eb[2], &r.ec[2]);
                    int x0, x1, y0, y1;
                    pixel_range(fminf(p[0].x, fminf(p[1].x, p[2].x)), fmaxf(p[0].x, fmaxf(p[1].x, p[2].x)), F.width, &x0, &x1);
                    pixel_range(fminf(p[0].y, fminf(p[1].y, p[2].y)), fmaxf(p[0].y, fmaxf(p[1].y, p[2].y)), F.height, &y0, &y1);
                    x0 = max(x0, sx0); x1 = min(x1, sx1); y0 = max(y0, sy0); y1 = min(y1, sy1);
                    if (x0 < x1 && y0 < y1) {
                        r.bbx = (uint32_t)x0 | ((uint32_t)x1 << 16);
                        r.bby = (uint32_t)y0 | ((uint32_t)y1 << 16);
                    }
                } else {
                    // line segment t of the batch (rasterizer.rs:901-955): Lines uses the first two indices of
                    // triple t, LineStrip vertices (t, t+1), LineLoop (t, (t+1) % n)
                    uint32_t i0, i1;
                    if (B.mode == RXC_MODE_LINES) { i0 = S.idx2[(size_t)(B.t_off + t) * 3 + 0]; i1 = S.idx2[(size_t)(B.t_off + t) * 3 + 1]; }
                    else { i0 = B.v_off + t; i1 = B.v_off + (t + 1u) % B.n_verts; }
                    float2 p[2] = {S.pos2[i0], S.pos2[i1]};
                    if (F.has_mat2d) {
                        for (int k = 0; k < 2; ++k) { f3 q = rx_matvec3(F.mat2d, {p[k].x, p[k].y, 1.0f}, F.matvec_mode); p[k].x = q.x; p[k].y = q.y; }
                    }
                    // `p as isize` (rasterizer.rs:1785-1788): truncation, NaN -> 0, saturating (here to +-2^30)
                    auto as_isize = [](float v) { return (v == v) ? (int)fminf(fmaxf(v, -1073741824.0f), 1073741824.0f) : 0; };
                    const int x0 = as_isize(p[0].x), y0 = as_isize(p[0].y), x1 = as_isize(p[1].x), y1 = as_isize(p[1].y);
                    r.ax = __int_as_float(x0); r.ay = __int_as_float(y0); r.bx = __int_as_float(x1); r.by = __int_as_float(y1);
                    int bx0 = max(max(min(x0, x1), 0), sx0), bx1 = min(min(max(x0, x1) + 1, F.width), sx1);
                    int by0 = max(max(min(y0, y1), 0), sy0), by1 = min(min(max(y0, y1) + 1, F.height), sy1);
                    if (bx0 < bx1 && by0 < by1) {
                        r.bbx = (uint32_t)bx0 | ((uint32_t)bx1 << 16);
                        r.bby = (uint32_t)by0 | ((uint32_t)by1 << 16);
                    }
                }
            }
            recs[t] = r;
        }
        return;
    }

    // remaining CTAs: per-batch frame state, and zeroing of the per-tile counters of this frame
    const uint32_t zb = bk.x - 1 - S.n_b2, nzb = bk.nx - 1 - S.n_b2;
    for (uint32_t b = zb * blockDim.x + tid; b < S.n_b3; b += nzb * blockDim.x) {  // per (frame, 3D batch) state
        const DBatch3& B = S.b3[b];
        DFrameBatch fb;
        float pv[16], mvp[16];
        rx_matmat4(F.proj, F.view, pv, F.matvec_mode);          // batch3d.rs:490
        rx_matmat4(pv, B.transform, mvp, F.matvec_mode);
        rx_matmat4(F.view, B.transform, fb.view_model, F.matvec_mode);  // batch3d.rs:555
        bool rejected = false;
        if (B.n_verts != 0 && !F.preprojected) {  // batch3d.rs:493-552 (pre-projected frames: the host's bounding_box decides, k_tri_setup_projected)
            bool ol = true, orr = true, ob = true, ot = true, on = true, of = true;
            for (int c = 0; c < 8; ++c) {
                f4 v = {(c & 4) ? B.aabb_max[0] : B.aabb_min[0], (c & 2) ? B.aabb_max[1] : B.aabb_min[1],
                        (c & 1) ? B.aabb_max[2] : B.aabb_min[2], 1.0f};
                f4 r = rx_matvec4(mvp, v, F.matvec_mode);
                float w = r.w;
                ol &= r.x < -w; orr &= r.x > w; ob &= r.y < -w; ot &= r.y > w; on &= r.z < -w; of &= r.z > w;
            }
            rejected = ol || orr || ob || ot || on || of;
        }
        // Band / rectangle rendering (one rank of a band split): a batch none of whose triangles can touch the rectangle
        // is dropped here, before k_tri_setup loads a single vertex of it -- this is what shards the front end of a band
        // split across the ranks.  Conservative: every vertex has w = 1 and lies inside the object AABB (RX_BF_UNIT_W),
        // all eight corners are in front of the near-clip plane (view z < -0.11: no triangle of the batch is clipped, every
        // clip w is positive), so every projected vertex is a convex combination of the projected corners; the bounds get
        // a margin for the rounding of the two different matrix chains.  Only pixels outside the rectangle can differ.
        const bool sub_rect = F.band_y0 > 0 || F.band_y1 < F.height || F.band_x0 > 0 || F.band_x1 < F.width;
        if (sub_rect && !rejected && !F.preprojected && (B.bflags & RX_BF_UNIT_W)) {
            bool front = true;
            float sx0 = CUDART_INF_F, sx1 = -CUDART_INF_F, sy0 = CUDART_INF_F, sy1 = -CUDART_INF_F;
            for (int c = 0; c < 8; ++c) {
                const f4 v = {(c & 4) ? B.aabb_max[0] : B.aabb_min[0], (c & 2) ? B.aabb_max[1] : B.aabb_min[1],
                              (c & 1) ? B.aabb_max[2] : B.aabb_min[2], 1.0f};
                const f4 vv = rx_matvec4(fb.view_model, v, F.matvec_mode);
                const f4 r = rx_matvec4(mvp, v, F.matvec_mode);
                front = front && vv.z < -0.11f && r.w > 1e-3f;
                const float x = ((r.x / r.w) * 0.5f + 0.5f) * F.width_f, y = ((-r.y / r.w) * 0.5f + 0.5f) * F.height_f;
                sx0 = fminf(sx0, x); sx1 = fmaxf(sx1, x); sy0 = fminf(sy0, y); sy1 = fmaxf(sy1, y);
            }
            if (front && sx0 == sx0 && sx1 == sx1 && sy0 == sy0 && sy1 == sy1) {
                const float mx = 2.0f + 1e-4f * fmaxf(fabsf(sx0), fabsf(sx1)), my = 2.0f + 1e-4f * fmaxf(fabsf(sy0), fabsf(sy1));
                if (sy1 + my < (float)F.band_y0 || sy0 - my > (float)F.band_y1 || sx1 + mx < (float)F.band_x0 || sx0 - mx > (float)F.band_x1) rejected = true;
            }
        }
        fb.tex = 0xFFFFFFFFu;
        fb.alpha_test = 0;
        fb.sd_tex_word = 0; fb.sd_wh = 0; fb.sd_pad[0] = fb.sd_pad[1] = 0;
        fb.sd_chunk = B.chunk; fb.sd_profile = B.profile_id;
        fb.sd_flags = B.has_normals ? RX_SD_NORMALS : 0u;
        if (B.bflags & RX_BF_HAS_PROFILE) fb.sd_flags |= RX_SD_HAS_PROFILE;
        if (B.bflags & RX_BF_OPACITY) fb.sd_flags |= RX_SD_OPACITY;
        fb.sd_pixel = 0xFF000000u;  // rasterizer.rs:1221
        fb.sd_ambient[0] = B.ambient[0]; fb.sd_ambient[1] = B.ambient[1]; fb.sd_ambient[2] = B.ambient[2];
        if (B.repeat_mode == RXC_REPEAT_REPEAT_XY || B.repeat_mode == RXC_REPEAT_REPEAT_X) fb.sd_flags |= RX_SD_REPEAT_X;
        if (B.repeat_mode == RXC_REPEAT_REPEAT_XY || B.repeat_mode == RXC_REPEAT_REPEAT_Y) fb.sd_flags |= RX_SD_REPEAT_Y;
        if (B.source_kind == RXC_SRC_PIXEL) {
            fb.sd_pixel = B.source_pixel;
        } else if (is_tile_source(B.source_kind)) {
            DTile t;
            if (source_tile(S, B.source_kind, B.source_index, &t)) {
                fb.tex = t.first + (uint32_t)(F.animation_frame % t.n_frames);  // rasterizer.rs:1104-1105
                const DTex tx = S.tex[fb.tex];
                fb.alpha_test = tx.all_opaque ? 0u : 1u;
                fb.sd_tex_word = (uint32_t)(tx.offset >> 2);
                fb.sd_wh = tx.width | (tx.height << 16);
                fb.sd_flags |= RX_SD_TEXTURED;
            } else {
                fb.sd_pixel = 0u;  // EntityTile / ItemTile that does not resolve: [0,0,0,0] (:1146-1151)
            }
        } else if (B.source_kind == RXC_SRC_TERRAIN) {  // :1178-1219
            if (B.chunk < 0) {
                fb.sd_pixel = 0xFF0000FFu;  // [255, 0, 0, 255]
            } else {
                const uint32_t tt = S.chunk_info[B.chunk].terrain_tex;
                if (tt == 0xFFFFFFFFu) {
                    fb.sd_pixel = 0u;       // chunk without a terrain texture: [0,0,0,0] (chunk.rs:150)
                } else {
                    const DTex tx = S.tex[tt];
                    fb.tex = tt;
                    fb.alpha_test = tx.all_opaque ? 0u : 1u;
                    fb.sd_tex_word = (uint32_t)(tx.offset >> 2);
                    fb.sd_wh = tx.width | (tx.height << 16);
                    fb.sd_flags |= RX_SD_TERRAIN;
                }
            }
        }
        fb.sd_program = -1;
        if (B.program >= 0 && (uint32_t)B.program < S.vm.n_programs && S.vm.programs[B.program].n_words != 0u) {  // rasterizer.rs:1226-1293
            fb.sd_program = B.program;
            fb.sd_flags |= RX_SD_SHADER;
            if (S.vm.programs[B.program].sets_opacity) { fb.sd_flags |= RX_SD_VM_OPACITY; fb.alpha_test = 1u; }  // opacity is the program's
        }
        // an opaque-pass batch whose constant texel is not opaque can never write (rasterizer.rs:1408)
        if (!(fb.sd_flags & (RX_SD_TEXTURED | RX_SD_TERRAIN | RX_SD_OPACITY | RX_SD_VM_OPACITY)) && (fb.sd_pixel >> 24) != 255u) rejected = true;
        fb.bb_minx = rx_float_key(CUDART_INF_F); fb.bb_maxx = rx_float_key(-CUDART_INF_F);
        fb.bb_miny = rx_float_key(CUDART_INF_F); fb.bb_maxy = rx_float_key(-CUDART_INF_F);
        fb.sc_x0 = fb.sc_x1 = fb.sc_y0 = fb.sc_y1 = 0;
        fb.rejected = (rejected || !F.d3_active) ? 1u : 0u;
        fb.n_new_tris = 0;
        Wk.fb[(size_t)f * Wk.fb_stride + b] = fb;
    }

    uint32_t* tc = Wk.tile_count + (size_t)f * Wk.tile_stride;
    uint32_t* tf = Wk.tile_fill + (size_t)f * Wk.tile_stride;
    for (uint32_t i = zb * blockDim.x + tid; i < tiles_per_frame; i += nzb * blockDim.x) { tc[i] = 0u; tf[i] = 0u; }
    if (S.general) {
        uint32_t* tc2 = Wk.tile_count2 + (size_t)f * Wk.tile_stride;
        uint32_t* tf2 = Wk.tile_fill2 + (size_t)f * Wk.tile_stride;
        for (uint32_t i = zb * blockDim.x + tid; i < tiles_per_frame; i += nzb * blockDim.x) { tc2[i] = 0u; tf2[i] = 0u; }
    }
}

// ---------------------------------------------------------------------------------------------
// k_tri_setup : one CTA per chunk (<= 256 triangles of one batch), one thread per triangle
// ---------------------------------------------------------------------------------------------
struct TriLoad {
    f4 vv[3];       // view space
    float2 uv[3];
    f3 nn[3];
};

__device__ __forceinline__ void load_tri(const SceneDev& S, const DBatch3& B, const float* view_model, uint32_t mode,
                                         uint32_t tri, TriLoad* T) {
    const uint32_t i0 = S.idx[(size_t)tri * 3 + 0], i1 = S.idx[(size_t)tri * 3 + 1], i2 = S.idx[(size_t)tri * 3 + 2];
    const uint32_t ii[3] = {i0, i1, i2};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float4 p = __ldg(S.pos + ii[k]);
        T->vv[k] = rx_matvec4(view_model, {p.x, p.y, p.z, p.w}, mode);  // batch3d.rs:557-560
        T->uv[k] = __ldg(S.uv + ii[k]);
        if (B.has_normals) T->nn[k] = {__ldg(S.nrm + (size_t)ii[k] * 3), __ldg(S.nrm + (size_t)ii[k] * 3 + 1), __ldg(S.nrm + (size_t)ii[k] * 3 + 2)};
        else T->nn[k] = {0.0f, 0.0f, 0.0f};
    }
}

__device__ __forceinline__ void d_tri_setup(const SceneDev& S, const Workspace& Wk, Blk bk) {
    const uint32_t f = bk.y;
    const DFrame& F = Wk.frames[f];
    const DChunk ch = S.chunks[bk.x];
    const DBatch3& B = S.b3[ch.batch];
    DFrameBatch& FB = Wk.fb[(size_t)f * Wk.fb_stride + ch.batch];
    const uint32_t tid = threadIdx.x;
    uint32_t* chunk_total = Wk.chunk_new_total + (size_t)f * Wk.chunk_stride + bk.x;

    if (FB.rejected) {  // uniform for the CTA
        if (tid == 0) *chunk_total = 0u;
        // bins of a rejected batch must read as empty
        if (tid < ch.n_tris) {
            TriBin e = {0u, 0u, 0u, ch.batch};
            Wk.bins[(size_t)f * Wk.bins_stride + ch.first_tri + tid] = e;
        }
        return;
    }

    MinMax mm; mm.init();
    uint32_t n_new = 0;      // near-clip output triangles of this thread's triangle
    bool vis = false;

    if (tid < ch.n_tris) {
        const uint32_t tri = ch.first_tri + tid;           // global original triangle index
        const uint32_t local = tri - B.t_off;
        const uint32_t slot = B.owner_base + local;
        TriLoad T;
        load_tri(S, B, FB.view_model, F.matvec_mode, tri, &T);

        bool early_cull = false;  // batch3d.rs:592-600
        if (B.cull_mode != RXC_CULL_OFF) {
            float orient = (T.vv[1].x - T.vv[0].x) * (T.vv[2].y - T.vv[0].y) - (T.vv[1].y - T.vv[0].y) * (T.vv[2].x - T.vv[0].x);
            bool is_front = orient > 0.0f;
            early_cull = (B.cull_mode == RXC_CULL_BACK && is_front) || (B.cull_mode == RXC_CULL_FRONT && !is_front);
        }
        const bool in0 = T.vv[0].z < -RX_NEAR_PLANE, in1 = T.vv[1].z < -RX_NEAR_PLANE, in2 = T.vv[2].z < -RX_NEAR_PLANE;
        const bool all_in = in0 && in1 && in2, all_out = !in0 && !in1 && !in2;
        const bool edge_vis = early_cull || all_in;                // batch3d.rs:613-618
        const bool mixed = !early_cull && !all_in && !all_out;

        f4 P[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {                               // batch3d.rs:689-700
            P[k] = rx_project(F.proj, T.vv[k], F.width_f, F.height_f, F.matvec_mode);
            mm.add(P[k].x, P[k].y);
        }
        TriBin bin = {0u, 0u, slot, ch.batch};
        const uint32_t meta = ch.batch | tri_meta_flags(FB);  // make_tri adds RX_META_FASTDIV
        TriVis tv; TriShade tsh;
        vis = make_tri(P, T.uv, T.nn, B.cull_mode, edge_vis, F.width, F.height, meta, &tv, &tsh, &bin.bbx, &bin.bby);
        // band rendering (a rank of a row split): a triangle whose pixel rows miss the band is dropped before its
        // 176 B of records are written; the batch bbox above still saw it, like the reference's projected_vertices
        if (vis && ((int)(bin.bby >> 16) <= F.band_y0 || (int)(bin.bby & 0xFFFFu) >= F.band_y1 || (int)(bin.bbx >> 16) <= F.band_x0 ||
                    (int)(bin.bbx & 0xFFFFu) >= F.band_x1)) { vis = false; bin.bbx = 0u; bin.bby = 0u; }
        if (vis) {
            Wk.vis[(size_t)f * Wk.slot_stride + slot] = tv;
            Wk.shade[(size_t)f * Wk.slot_stride + slot] = tsh;
        }
        Wk.bins[(size_t)f * Wk.bins_stride + tri] = bin;

        if (mixed) {
            ClipVert poly[4];
            const int nv = clip_polygon(T.vv, T.uv, T.nn, poly);
            for (int k = 0; k < nv; ++k) {   // the appended vertices are part of projected_vertices (bbox)
                f4 q = rx_project(F.proj, poly[k].p, F.width_f, F.height_f, F.matvec_mode);
                mm.add(q.x, q.y);
            }
            n_new = nv >= 3 ? (uint32_t)(nv - 2) : 0u;
        }
    }

    // vertices no triangle references still belong to projected_vertices (batch3d.rs:749-768)
    if (bk.x == B.chunk_first) {
        for (uint32_t o = tid; o < B.n_orphans; o += blockDim.x) {
            const float4 p = __ldg(S.pos + S.orphans[B.orphan_off + o]);
            f4 vvv = rx_matvec4(FB.view_model, {p.x, p.y, p.z, p.w}, F.matvec_mode);
            f4 q = rx_project(F.proj, vvv, F.width_f, F.height_f, F.matvec_mode);
            mm.add(q.x, q.y);
        }
    }

    // block exclusive scan of n_new in {0,1,2}: two ballots per warp + warp totals in shared memory
    __shared__ uint32_t s_wtot[RX_CHUNK_TRIS / 32];
    const uint32_t lane = tid & 31, warp = tid >> 5;
    const uint32_t b0 = __ballot_sync(0xFFFFFFFFu, n_new & 1u), b1 = __ballot_sync(0xFFFFFFFFu, n_new & 2u);
    const uint32_t wprefix = __popc(b0 & lanemask_lt()) + 2u * __popc(b1 & lanemask_lt());
    if (lane == 0) s_wtot[warp] = __popc(b0) + 2u * __popc(b1);
    __syncthreads();
    uint32_t base = 0, total = 0;
    for (uint32_t w = 0; w < RX_CHUNK_TRIS / 32; ++w) { if (w < warp) base += s_wtot[w]; total += s_wtot[w]; }
    if (tid == 0) *chunk_total = total;
    if (n_new) {
        const uint32_t k = atomicAdd(&Wk.counters[f].n_clip, 1u);
        if (k < Wk.clip_stride) {
            DClip c = {ch.first_tri + tid, bk.x, base + wprefix, ch.batch};
            Wk.clip[(size_t)f * Wk.clip_stride + k] = c;
        } else {
            atomicOr(&Wk.counters[f].overflow, 4u);
        }
    }
    const uint32_t nvis = __popc(__ballot_sync(0xFFFFFFFFu, vis));
    if (lane == 0 && nvis) atomicAdd(&Wk.counters[f].n_visible, nvis);

    block_minmax_to_keys(mm, &FB.bb_minx, &FB.bb_maxx, &FB.bb_miny, &FB.bb_maxy);
}

#ifndef __CUDACC_RTC__   // (the JIT build recompiles k_raster only)
// ---------------------------------------------------------------------------------------------
// k_tri_setup_projected : rxc_rasterize_projected.  The host ran Scene::project itself (its own vek arithmetic) and hands
// over projected_vertices / clipped_indices / clipped_uvs / clipped_normals / edges / bounding_box per batch; one thread
// per clipped triangle builds the same records k_tri_setup / k_clip_emit build, with the host's edge equations and
// visibility verbatim; one thread per batch turns bounding_box into the scissor of rasterizer.rs:978-983.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_tri_setup_projected(SceneDev S, Workspace Wk, ProjectedDev Pj) {
    pdl_enter();
    const DFrame& F = Wk.frames[0];
    DCounters& C = Wk.counters[0];
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < S.n_b3) {
        DFrameBatch& FB = Wk.fb[i];
        const float* bb = Pj.bbox + (size_t)i * 5;
        if (bb[0] == 0.0f) FB.rejected = 1u;   // bounding_box == None: d3_rasterize returns at once (rasterizer.rs:976-977)
        if (!FB.rejected) {
            int x0, x1, y0, y1;
            scissor_1d(bb[1], bb[3], F.width, (int)F.tile_size, 0.0f, &x0, &x1);  // rasterizer.rs:978-983
            scissor_1d(bb[2], bb[4], F.height, (int)F.tile_size, 0.0f, &y0, &y1);
            FB.sc_x0 = max(x0, F.band_x0); FB.sc_x1 = min(x1, F.band_x1);
            FB.sc_y0 = max(y0, F.band_y0); FB.sc_y1 = min(y1, F.band_y1);
        }
    }
    const uint32_t total = max(S.n_tris, Pj.n_clipped);   // k_bin_count streams bins [0, max(n_tris, n_clipped))
    if (i == 0) C.n_new_slots = Pj.n_clipped > S.n_tris ? Pj.n_clipped - S.n_tris : 0u;
    if (i >= total) return;
    TriBin bin = {0u, 0u, 0u, 0u};
    if (i < Pj.n_clipped) {
        const uint32_t w0 = Pj.info[2 * i], slot = Pj.info[2 * i + 1];
        const uint32_t b = w0 & 0x7FFFFFFFu;
        bin.slot = slot; bin.batch = b;
        const DFrameBatch& FB = Wk.fb[b];
        const float* bb = Pj.bbox + (size_t)b * 5;
        if ((w0 >> 31) && bb[0] != 0.0f && F.d3_active) {
            const uint32_t i0 = Pj.idx[3 * i], i1 = Pj.idx[3 * i + 1], i2 = Pj.idx[3 * i + 2];
            const float4 p0 = Pj.pv[i0], p1 = Pj.pv[i1], p2 = Pj.pv[i2];
            const f4 P[3] = {{p0.x, p0.y, p0.z, p0.w}, {p1.x, p1.y, p1.z, p1.w}, {p2.x, p2.y, p2.z, p2.w}};
            const float2 uv[3] = {Pj.uv[i0], Pj.uv[i1], Pj.uv[i2]};
            const f3 nn[3] = {{Pj.nrm[3 * i0], Pj.nrm[3 * i0 + 1], Pj.nrm[3 * i0 + 2]}, {Pj.nrm[3 * i1], Pj.nrm[3 * i1 + 1], Pj.nrm[3 * i1 + 2]},
                              {Pj.nrm[3 * i2], Pj.nrm[3 * i2 + 1], Pj.nrm[3 * i2 + 2]}};
            const float* e = Pj.edges + (size_t)i * 9;
            const float ea[3] = {e[0], e[1], e[2]}, eb[3] = {e[3], e[4], e[5]}, ec[3] = {e[6], e[7], e[8]};
            TriVis tv; TriShade tsh;
            uint32_t bbx = 0u, bby = 0u;
            // a batch whose constant texel can never be written was already marked by k_frame_setup (FB.rejected)
            if (!FB.rejected && tri_records(P, uv, nn, ea, eb, ec, F.width, F.height, b | tri_meta_flags(FB), &tv, &tsh, &bbx, &bby) &&
                !((int)(bby >> 16) <= F.band_y0 || (int)(bby & 0xFFFFu) >= F.band_y1 || (int)(bbx >> 16) <= F.band_x0 || (int)(bbx & 0xFFFFu) >= F.band_x1)) {
                bin.bbx = bbx; bin.bby = bby;
                Wk.vis[slot] = tv;
                Wk.shade[slot] = tsh;
                atomicAdd(&C.n_visible, 1u);
            }
        }
    }
    Wk.bins[i] = bin;
}

#endif

// ---------------------------------------------------------------------------------------------
// k_batch_finalize : one warp per (frame, batch)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void d_batch_finalize(const SceneDev& S, const Workspace& Wk, Blk bk) {
    const uint32_t f = bk.y;
    const uint32_t b = bk.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const uint32_t lane = threadIdx.x & 31;
    if (b >= S.n_b3) return;
    const DFrame& F = Wk.frames[f];
    const DBatch3& B = S.b3[b];
    DFrameBatch& FB = Wk.fb[(size_t)f * Wk.fb_stride + b];
    const uint32_t* tot = Wk.chunk_new_total + (size_t)f * Wk.chunk_stride + B.chunk_first;
    uint32_t* basep = Wk.chunk_new_base + (size_t)f * Wk.chunk_stride + B.chunk_first;
    uint32_t carry = 0;
    for (uint32_t c0 = 0; c0 < B.n_chunks; c0 += 32) {
        const uint32_t c = c0 + lane;
        uint32_t v = (c < B.n_chunks) ? tot[c] : 0u;
        uint32_t incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t n = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if ((int)lane >= o) incl += n;
        }
        if (c < B.n_chunks) basep[c] = carry + incl - v;
        carry += __shfl_sync(0xFFFFFFFFu, incl, 31);
    }
    if (lane == 0) {
        FB.n_new_tris = carry;
        if (!FB.rejected) {
            // Rect{x: min_x, y: min_y, width: max_x - min_x, height: max_y - min_y} (batch3d.rs:762-767)
            const float mnx = rx_key_float(FB.bb_minx), mxx = rx_key_float(FB.bb_maxx);
            const float mny = rx_key_float(FB.bb_miny), mxy = rx_key_float(FB.bb_maxy);
            int x0, x1, y0, y1;
            scissor_1d(mnx, mxx - mnx, F.width, (int)F.tile_size, 0.0f, &x0, &x1);  // rasterizer.rs:978-983
            scissor_1d(mny, mxy - mny, F.height, (int)F.tile_size, 0.0f, &y0, &y1);
            FB.sc_x0 = max(x0, F.band_x0); FB.sc_x1 = min(x1, F.band_x1);
            FB.sc_y0 = max(y0, F.band_y0); FB.sc_y1 = min(y1, F.band_y1);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// k_clip_emit : grid-stride over the compact list of near-clipped triangles
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void d_clip_emit(const SceneDev& S, const Workspace& Wk, Blk bk) {
    const uint32_t f = bk.y;
    const DFrame& F = Wk.frames[f];
    DCounters& C = Wk.counters[f];
    const uint32_t n = min(C.n_clip, Wk.clip_stride);
    const uint32_t new_cap = Wk.bins_stride - S.n_tris;
    for (uint32_t k = bk.x * blockDim.x + threadIdx.x; k < n; k += bk.nx * blockDim.x) {
        const DClip c = Wk.clip[(size_t)f * Wk.clip_stride + k];
        const DBatch3& B = S.b3[c.batch];
        const DFrameBatch& FB = Wk.fb[(size_t)f * Wk.fb_stride + c.batch];
        TriLoad T;
        load_tri(S, B, FB.view_model, F.matvec_mode, c.tri, &T);
        ClipVert poly[4];
        const int nv = clip_polygon(T.vv, T.uv, T.nn, poly);
        f4 Q[4];
        for (int i = 0; i < nv; ++i) Q[i] = rx_project(F.proj, poly[i].p, F.width_f, F.height_f, F.matvec_mode);
        const uint32_t first = B.owner_base + B.n_tris + Wk.chunk_new_base[(size_t)f * Wk.chunk_stride + c.chunk] + c.local_off;
        const uint32_t meta = c.batch | tri_meta_flags(FB);
        for (int j = 1; j + 1 < nv; ++j) {  // fan (c0, cj, cj+1): batch3d.rs:672-678
            const f4 P[3] = {Q[0], Q[j], Q[j + 1]};
            const float2 uv[3] = {{poly[0].u, poly[0].v}, {poly[j].u, poly[j].v}, {poly[j + 1].u, poly[j + 1].v}};
            const f3 nn[3] = {poly[0].n, poly[j].n, poly[j + 1].n};
            const uint32_t slot = first + (uint32_t)(j - 1);
            TriBin bin = {0u, 0u, slot, c.batch};
            TriVis tv; TriShade tsh;
            if (make_tri(P, uv, nn, B.cull_mode, true, F.width, F.height, meta, &tv, &tsh, &bin.bbx, &bin.bby) &&
                !((int)(bin.bby >> 16) <= F.band_y0 || (int)(bin.bby & 0xFFFFu) >= F.band_y1 || (int)(bin.bbx >> 16) <= F.band_x0 ||
                  (int)(bin.bbx & 0xFFFFu) >= F.band_x1)) {
                const uint32_t pos = atomicAdd(&C.n_new_slots, 1u);
                if (pos < new_cap) {
                    Wk.vis[(size_t)f * Wk.slot_stride + slot] = tv;
                    Wk.shade[(size_t)f * Wk.slot_stride + slot] = tsh;
                    Wk.bins[(size_t)f * Wk.bins_stride + S.n_tris + pos] = bin;
                    atomicAdd(&C.n_visible, 1u);
                } else {
                    atomicOr(&C.overflow, 4u);
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// binning
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool bin_tile_range(const DFrame& F, uint32_t bbx, uint32_t bby, int* tx0, int* tx1, int* ty0, int* ty1) {
    const int x0 = bbx & 0xFFFF, x1 = bbx >> 16, y0 = bby & 0xFFFF, y1 = bby >> 16;
    if (x0 >= x1 || y0 >= y1) return false;
    *tx0 = (x0 - F.band_x0) / RX_TILE_W; *tx1 = (x1 - 1 - F.band_x0) / RX_TILE_W;
    *ty0 = (y0 - F.band_y0) / RX_TILE_H; *ty1 = (y1 - 1 - F.band_y0) / RX_TILE_H;
    return true;
}

// Warp-aggregated tile counters: the lanes that are at the same tile in the same step share one atomic
// (`__match_any_sync` over whoever is converged here).  Neighbouring triangles of a mesh fall into the same tiles, and at
// the horizon of a dense mesh thousands of them into the same few -- one atomic per tile and warp instead of one per
// triangle and tile.  tile_slot() returns the caller's own entry of the group's reservation.
__device__ __forceinline__ void tile_count(uint32_t* tc, int t) {
    const unsigned peers = __match_any_sync(__activemask(), t);
    if ((uint32_t)(__ffs(peers) - 1) == (threadIdx.x & 31u)) atomicAdd(&tc[t], (uint32_t)__popc(peers));
}
__device__ __forceinline__ uint32_t tile_slot(uint32_t* tf, int t) {
    const unsigned peers = __match_any_sync(__activemask(), t);
    const uint32_t lane = threadIdx.x & 31u, leader = (uint32_t)(__ffs(peers) - 1);
    uint32_t pos = 0u;
    if (lane == leader) pos = atomicAdd(&tf[t], (uint32_t)__popc(peers));
    return __shfl_sync(peers, pos, (int)leader) + (uint32_t)__popc(peers & ((1u << lane) - 1u));
}

__device__ __forceinline__ void d_bin_count(const SceneDev& S, const Workspace& Wk, Blk bk) {
    const uint32_t f = bk.y;
    const DFrame& F = Wk.frames[f];
    DCounters& C = Wk.counters[f];
    const uint32_t new_cap = Wk.bins_stride - S.n_tris;
    const uint32_t total = S.n_tris + min(C.n_new_slots, new_cap);
    TriBin* bins = Wk.bins + (size_t)f * Wk.bins_stride;
    uint32_t* tc = Wk.tile_count + (size_t)f * Wk.tile_stride;
    for (uint32_t i = bk.x * blockDim.x + threadIdx.x; i < total; i += bk.nx * blockDim.x) {
        TriBin b = bins[i];
        int x0 = b.bbx & 0xFFFF, x1 = b.bbx >> 16, y0 = b.bby & 0xFFFF, y1 = b.bby >> 16;
        if (x0 >= x1 || y0 >= y1) continue;
        const DFrameBatch& FB = Wk.fb[(size_t)f * Wk.fb_stride + b.batch];
        x0 = max(x0, FB.sc_x0); x1 = min(x1, FB.sc_x1); y0 = max(y0, FB.sc_y0); y1 = min(y1, FB.sc_y1);
        if (x0 >= x1 || y0 >= y1) { bins[i].bbx = 0u; bins[i].bby = 0u; continue; }
        // the scissor almost never clips a triangle of its own batch (it is the union of the API tiles that overlap the
        // batch bbox): the records are rewritten only when it, or the band, did
        const uint32_t nbx = (uint32_t)x0 | ((uint32_t)x1 << 16), nby = (uint32_t)y0 | ((uint32_t)y1 << 16);
        bool dirty = nbx != b.bbx || nby != b.bby;
        if (dirty) {
            b.bbx = nbx; b.bby = nby;
            TriVis* tv = Wk.vis + (size_t)f * Wk.slot_stride + b.slot;
            tv->bbx = nbx; tv->bby = nby;
        }
        int tx0, tx1, ty0, ty1;
        bin_tile_range(F, b.bbx, b.bby, &tx0, &tx1, &ty0, &ty1);
        const int nt = (tx1 - tx0 + 1) * (ty1 - ty0 + 1);
        if (nt > RX_LARGE_TILES) {  // general mode: k_bin_large bins these into the ordered lists, a warp per triangle
            const uint32_t k = atomicAdd(&C.n_large, 1u);
            if (k < Wk.large_stride) Wk.large[(size_t)f * Wk.large_stride + k] = b.slot;
            else atomicOr(&C.overflow, 2u);
            b.batch |= 0x80000000u;
            dirty = true;
        } else if (nt == 1) {
            tile_count(tc, ty0 * F.tiles_x + tx0);
        } else {
            for (int ty = ty0; ty <= ty1; ++ty)
                for (int tx = tx0; tx <= tx1; ++tx) tile_count(tc, ty * F.tiles_x + tx);
        }
        if (dirty) bins[i] = b;
    }
}

__device__ __forceinline__ void d_tile_alloc(const Workspace& Wk, uint32_t tiles_per_frame, int which, int pow2, Blk bk) {
    const uint32_t f = bk.y;
    DCounters& C = Wk.counters[f];
    const uint32_t i = bk.x * blockDim.x + threadIdx.x;
    if (i >= tiles_per_frame) return;
    uint32_t* tc = (which ? Wk.tile_count2 : Wk.tile_count) + (size_t)f * Wk.tile_stride;
    uint32_t* tb = (which ? Wk.tile_base2 : Wk.tile_base) + (size_t)f * Wk.tile_stride;
    const uint32_t cap = which ? Wk.list2_stride : Wk.list_stride;
    const uint32_t n = tc[i];
    uint32_t base = 0;
    if (n) {
        uint32_t want = n;
        if (pow2 && n > 1u) want = 1u << (32 - __clz(n - 1u));  // room for the sort's padding
        base = atomicAdd(which ? &C.list_cursor2 : &C.list_cursor, want);
        if (base + want > cap) { atomicOr(&C.overflow, which ? 8u : 1u); tc[i] = 0u; base = 0; }
    }
    tb[i] = base;
}

__device__ __forceinline__ void d_bin_fill(const SceneDev& S, const Workspace& Wk, Blk bk) {
    const uint32_t f = bk.y;
    const DFrame& F = Wk.frames[f];
    DCounters& C = Wk.counters[f];
    const uint32_t new_cap = Wk.bins_stride - S.n_tris;
    const uint32_t total = S.n_tris + min(C.n_new_slots, new_cap);
    const TriBin* bins = Wk.bins + (size_t)f * Wk.bins_stride;
    const uint32_t* tc = Wk.tile_count + (size_t)f * Wk.tile_stride;
    const uint32_t* tb = Wk.tile_base + (size_t)f * Wk.tile_stride;
    uint32_t* tf = Wk.tile_fill + (size_t)f * Wk.tile_stride;
    uint32_t* lists = Wk.lists + (size_t)f * Wk.list_stride;
    for (uint32_t i = bk.x * blockDim.x + threadIdx.x; i < total; i += bk.nx * blockDim.x) {
        const TriBin b = bins[i];
        if (b.batch & 0x80000000u) continue;
        int tx0, tx1, ty0, ty1;
        if (!bin_tile_range(F, b.bbx, b.bby, &tx0, &tx1, &ty0, &ty1)) continue;
        if (tx0 == tx1 && ty0 == ty1) {   // single tile, the common case of a dense mesh
            const int t = ty0 * F.tiles_x + tx0;
            if (tc[t] == 0u) continue;  // list dropped on arena overflow
            lists[tb[t] + tile_slot(tf, t)] = b.slot;
            continue;
        }
        for (int ty = ty0; ty <= ty1; ++ty)
            for (int tx = tx0; tx <= tx1; ++tx) {
                const int t = ty * F.tiles_x + tx;
                if (tc[t] == 0u) continue;  // list dropped on arena overflow
                lists[tb[t] + tile_slot(tf, t)] = b.slot;
            }
    }
}

// ---------------------------------------------------------------------------------------------
// general mode: 2D record binning (one warp per record, lanes over the tiles of its bbox), binning of the large
// triangles into the ordered lists (one warp per triangle, skipping tiles the conservative overlap test excludes),
// and a warp-level list sort for the fused front end
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t rect_overlaps(const TriVis& T, int tx0, int ty0, int tx1, int ty1);

__device__ __forceinline__ void d_bin2d(const SceneDev& S, const Workspace& Wk, int fill, Blk bk) {
    const uint32_t f = bk.y;
    const DFrame& F = Wk.frames[f];
    const uint32_t lane = threadIdx.x & 31, wpg = (bk.nx * blockDim.x) >> 5;
    uint32_t* tc = Wk.tile_count2 + (size_t)f * Wk.tile_stride;
    const uint32_t* tb = Wk.tile_base2 + (size_t)f * Wk.tile_stride;
    uint32_t* tf = Wk.tile_fill2 + (size_t)f * Wk.tile_stride;
    uint32_t* lists = Wk.lists2 + (size_t)f * Wk.list2_stride;
    for (uint32_t r = (bk.x * blockDim.x + threadIdx.x) >> 5; r < S.n_rec2d; r += wpg) {
        const Tri2D& T = Wk.tri2d[(size_t)f * Wk.tri2d_stride + r];
        int tx0, tx1, ty0, ty1;
        if (!bin_tile_range(F, T.bbx, T.bby, &tx0, &tx1, &ty0, &ty1)) continue;
        const int w = tx1 - tx0 + 1, n = w * (ty1 - ty0 + 1);
        for (int i = (int)lane; i < n; i += 32) {
            const int t = (ty0 + i / w) * F.tiles_x + tx0 + i % w;
            if (!fill) { atomicAdd(&tc[t], 1u); continue; }
            if (tc[t] == 0u) continue;  // list dropped on arena overflow
            lists[tb[t] + atomicAdd(&tf[t], 1u)] = r;
        }
    }
}

__device__ __forceinline__ void d_bin_large(const SceneDev& S, const Workspace& Wk, int fill, Blk bk) {
    const uint32_t f = bk.y;
    const DFrame& F = Wk.frames[f];
    const DCounters& C = Wk.counters[f];
    const uint32_t n_large = min(C.n_large, Wk.large_stride);
    const uint32_t lane = threadIdx.x & 31, wpg = (bk.nx * blockDim.x) >> 5;
    uint32_t* tc = Wk.tile_count + (size_t)f * Wk.tile_stride;
    const uint32_t* tb = Wk.tile_base + (size_t)f * Wk.tile_stride;
    uint32_t* tf = Wk.tile_fill + (size_t)f * Wk.tile_stride;
    uint32_t* lists = Wk.lists + (size_t)f * Wk.list_stride;
    for (uint32_t k = (bk.x * blockDim.x + threadIdx.x) >> 5; k < n_large; k += wpg) {
        const uint32_t slot = Wk.large[(size_t)f * Wk.large_stride + k];
        const TriVis& T = Wk.vis[(size_t)f * Wk.slot_stride + slot];
        int tx0, tx1, ty0, ty1;
        if (!bin_tile_range(F, T.bbx, T.bby, &tx0, &tx1, &ty0, &ty1)) continue;
        const int w = tx1 - tx0 + 1, n = w * (ty1 - ty0 + 1);
        for (int i = (int)lane; i < n; i += 32) {
            const int tx = tx0 + i % w, ty = ty0 + i / w;
            const int px0 = F.band_x0 + tx * RX_TILE_W, py0 = F.band_y0 + ty * RX_TILE_H;
            if (rect_overlaps(T, px0, py0, min(px0 + RX_TILE_W, F.band_x1), min(py0 + RX_TILE_H, F.band_y1)) == 0u) continue;
            const int t = ty * F.tiles_x + tx;
            if (!fill) { atomicAdd(&tc[t], 1u); continue; }
            if (tc[t] == 0u) continue;  // list dropped on arena overflow
            lists[tb[t] + atomicAdd(&tf[t], 1u)] = slot;
        }
    }
}

// One warp per tile, tiles strided over the warps of the (virtual) grid: lists of fewer than two entries are done,
// the others are sorted ascending in place (bitonic over the power-of-two allocation, padded with 0xFFFFFFFF).  The
// small scenes that take the fused front end have short lists; k_list_sort (a CTA per tile) serves the big ones.
__device__ __forceinline__ void d_list_sort_warp(const Workspace& Wk, int which, uint32_t tiles_per_frame, Blk bk) {
    const uint32_t f = bk.y, lane = threadIdx.x & 31, wpg = (bk.nx * blockDim.x) >> 5;
    const uint32_t* tc = (which ? Wk.tile_count2 : Wk.tile_count) + (size_t)f * Wk.tile_stride;
    const uint32_t* tb = (which ? Wk.tile_base2 : Wk.tile_base) + (size_t)f * Wk.tile_stride;
    uint32_t* lists = which ? Wk.lists2 + (size_t)f * Wk.list2_stride : Wk.lists + (size_t)f * Wk.list_stride;
    for (uint32_t t0 = ((bk.x * blockDim.x + threadIdx.x) >> 5) * 32u; t0 < tiles_per_frame; t0 += wpg * 32u) {
        // 32 tiles per warp and round: one coalesced read of their counts, then the warp visits the ones with work
        const uint32_t mine = t0 + lane < tiles_per_frame ? tc[t0 + lane] : 0u;
        uint32_t todo = __ballot_sync(0xFFFFFFFFu, mine >= 2u);
        while (todo) {
            const int b = __ffs(todo) - 1;
            todo &= todo - 1u;
            const uint32_t n = __shfl_sync(0xFFFFFFFFu, mine, b);
            uint32_t* d = lists + tb[t0 + b];
            const uint32_t P = 1u << (32 - __clz(n - 1u));
            if (P <= 32u) {   // in registers
                uint32_t v = lane < n ? d[lane] : 0xFFFFFFFFu;
                for (uint32_t k = 2; k <= P; k <<= 1)
                    for (uint32_t j = k >> 1; j > 0; j >>= 1) {
                        const uint32_t o = __shfl_xor_sync(0xFFFFFFFFu, v, j);
                        const bool up = (lane & k) == 0u, low = (lane & j) == 0u;
                        v = (low == up) ? min(v, o) : max(v, o);
                    }
                if (lane < n) d[lane] = v;
            } else {          // in place, in global memory (L1 of this SM), warp-synchronous
                for (uint32_t i = n + lane; i < P; i += 32) d[i] = 0xFFFFFFFFu;
                __syncwarp();
                for (uint32_t k = 2; k <= P; k <<= 1)
                    for (uint32_t j = k >> 1; j > 0; j >>= 1) {
                        for (uint32_t i = lane; i < P; i += 32) {
                            const uint32_t ixj = i ^ j;
                            if (ixj > i) {
                                const uint32_t a = d[i], c = d[ixj];
                                if ((a > c) == ((i & k) == 0u)) { d[i] = c; d[ixj] = a; }
                            }
                        }
                        __syncwarp();
                    }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// the front-end as kernels ...
// ---------------------------------------------------------------------------------------------
#ifndef __CUDACC_RTC__
__global__ void __launch_bounds__(256) k_frame_setup(SceneDev S, Workspace Wk, uint32_t tiles_per_frame) { pdl_enter(); d_frame_setup(S, Wk, tiles_per_frame, Blk{blockIdx.x, blockIdx.y, gridDim.x}); }
__global__ void __launch_bounds__(RX_CHUNK_TRIS) k_tri_setup(SceneDev S, Workspace Wk) { pdl_enter(); d_tri_setup(S, Wk, Blk{blockIdx.x, blockIdx.y, gridDim.x}); }
__global__ void __launch_bounds__(256) k_batch_finalize(SceneDev S, Workspace Wk) { pdl_enter(); d_batch_finalize(S, Wk, Blk{blockIdx.x, blockIdx.y, gridDim.x}); }
__global__ void __launch_bounds__(128) k_clip_emit(SceneDev S, Workspace Wk) { pdl_enter(); d_clip_emit(S, Wk, Blk{blockIdx.x, blockIdx.y, gridDim.x}); }
__global__ void __launch_bounds__(256) k_bin_count(SceneDev S, Workspace Wk) { pdl_enter(); d_bin_count(S, Wk, Blk{blockIdx.x, blockIdx.y, gridDim.x}); }
__global__ void __launch_bounds__(256) k_tile_alloc(Workspace Wk, uint32_t tiles_per_frame, int which, int pow2) { pdl_enter(); d_tile_alloc(Wk, tiles_per_frame, which, pow2, Blk{blockIdx.x, blockIdx.y, gridDim.x}); }
__global__ void __launch_bounds__(256) k_bin_fill(SceneDev S, Workspace Wk) { pdl_enter(); d_bin_fill(S, Wk, Blk{blockIdx.x, blockIdx.y, gridDim.x}); }
__global__ void __launch_bounds__(256) k_bin2d(SceneDev S, Workspace Wk, int fill) { pdl_enter(); d_bin2d(S, Wk, fill, Blk{blockIdx.x, blockIdx.y, gridDim.x}); }
__global__ void __launch_bounds__(256) k_bin_large(SceneDev S, Workspace Wk, int fill) { pdl_enter(); d_bin_large(S, Wk, fill, Blk{blockIdx.x, blockIdx.y, gridDim.x}); }

// ... and fused for small scenes: one CTA per frame runs every front-end phase back to back (7 launches and
// their gaps cost ~70 us per call, more than rasterising a 1080p frame of such a scene).  Phases communicate
// through global memory written and read by this CTA only; __syncthreads orders them.
__global__ void __launch_bounds__(256) k_front_small(SceneDev S, Workspace Wk, uint32_t tiles_per_frame, uint32_t n_frame_blocks) {
    pdl_enter();
    const uint32_t f = blockIdx.x;
    for (uint32_t v = 0; v < n_frame_blocks; ++v) { d_frame_setup(S, Wk, tiles_per_frame, Blk{v, f, n_frame_blocks}); __syncthreads(); }
    if (S.n_tris == 0u) return;
    for (uint32_t v = 0; v < S.n_chunks; ++v) { d_tri_setup(S, Wk, Blk{v, f, S.n_chunks}); __syncthreads(); }
    const uint32_t nfb = (S.n_b3 + 7u) / 8u;
    for (uint32_t v = 0; v < nfb; ++v) d_batch_finalize(S, Wk, Blk{v, f, nfb});
    __syncthreads();
    d_clip_emit(S, Wk, Blk{0u, f, 1u});
    __syncthreads();
    d_bin_count(S, Wk, Blk{0u, f, 1u});
    __syncthreads();
    // nothing binned (every visible triangle went to the large list): all tile lists stay empty
    if (Wk.counters[f].n_visible == Wk.counters[f].n_large) {
        for (uint32_t i = threadIdx.x; i < tiles_per_frame; i += blockDim.x) Wk.tile_base[(size_t)f * Wk.tile_stride + i] = 0u;
        return;
    }
    const uint32_t ntb = (tiles_per_frame + 255u) / 256u;
    for (uint32_t v = 0; v < ntb; ++v) d_tile_alloc(Wk, tiles_per_frame, 0, 0, Blk{v, f, ntb});
    __syncthreads();
    d_bin_fill(S, Wk, Blk{0u, f, 1u});
}

// ... and fused for mid-sized scenes: one thread-block CLUSTER per frame.  The CTAs of the cluster share out the
// virtual blocks of every phase and meet at the hardware cluster barrier (barrier.cluster arrive.release /
// wait.acquire, which also orders the global-memory hand-over between phases) instead of at a kernel boundary:
// seven dependent launches (~10 us each of drain + launch + ramp) become six barriers of well under a microsecond.
__global__ void __launch_bounds__(256) k_front_cluster(SceneDev S, Workspace Wk, uint32_t tiles_per_frame, uint32_t n_frame_blocks, uint32_t stop_phase) {
    pdl_enter();
    cg::cluster_group cl = cg::this_cluster();
    const uint32_t r = cl.block_rank(), R = cl.num_blocks();   // the cluster spans grid.x; grid.y = frames
    const uint32_t f = blockIdx.y;
    const bool general = S.general != 0u, d2 = general && S.n_rec2d != 0u, d3 = S.n_tris != 0u;   // uniform over the grid
    const uint32_t ntb = (tiles_per_frame + 255u) / 256u;
    for (uint32_t v = r; v < n_frame_blocks; v += R) { d_frame_setup(S, Wk, tiles_per_frame, Blk{v, f, n_frame_blocks}); __syncthreads(); }
    if (!d3 && !d2) return;
    cl.sync();
    if (stop_phase == 1u) return;   // profiling aid (RXC_FRONT_STOP): leave after this many phases
    // the 2D chain of general mode (count, allocate, fill, sort) rides along with the first 3D phases
    if (d3) for (uint32_t v = r; v < S.n_chunks; v += R) { d_tri_setup(S, Wk, Blk{v, f, S.n_chunks}); __syncthreads(); }
    if (d2) d_bin2d(S, Wk, 0, Blk{r, f, R});
    cl.sync();
    if (stop_phase == 2u) return;   // profiling aid (RXC_FRONT_STOP): leave after this many phases
    if (d3) { const uint32_t nfb = (S.n_b3 + 7u) / 8u; for (uint32_t v = r; v < nfb; v += R) d_batch_finalize(S, Wk, Blk{v, f, nfb}); }
    if (d2) for (uint32_t v = r; v < ntb; v += R) d_tile_alloc(Wk, tiles_per_frame, 1, 1, Blk{v, f, ntb});
    cl.sync();
    if (stop_phase == 3u) return;   // profiling aid (RXC_FRONT_STOP): leave after this many phases
    if (d3) d_clip_emit(S, Wk, Blk{r, f, R});
    if (d2) d_bin2d(S, Wk, 1, Blk{r, f, R});
    cl.sync();
    if (stop_phase == 4u) return;   // profiling aid (RXC_FRONT_STOP): leave after this many phases
    if (d3) d_bin_count(S, Wk, Blk{r, f, R});
    if (d2) d_list_sort_warp(Wk, 1, tiles_per_frame, Blk{r, f, R});
    if (!d3) return;
    cl.sync();
    if (stop_phase == 5u) return;   // profiling aid (RXC_FRONT_STOP): leave after this many phases
    if (general) {   // the ordered lists hold every triangle: the large ones are binned too
        d_bin_large(S, Wk, 0, Blk{r, f, R});
        cl.sync();
    if (stop_phase == 6u) return;   // profiling aid (RXC_FRONT_STOP): leave after this many phases
    } else if (Wk.counters[f].n_visible == Wk.counters[f].n_large) {
        // nothing binned (every visible triangle went to the large list): all tile lists stay empty
        for (uint32_t i = r * blockDim.x + threadIdx.x; i < tiles_per_frame; i += R * blockDim.x) Wk.tile_base[(size_t)f * Wk.tile_stride + i] = 0u;
        return;
    }
    for (uint32_t v = r; v < ntb; v += R) d_tile_alloc(Wk, tiles_per_frame, 0, general ? 1 : 0, Blk{v, f, ntb});
    cl.sync();
    if (stop_phase == 7u) return;   // profiling aid (RXC_FRONT_STOP): leave after this many phases
    d_bin_fill(S, Wk, Blk{r, f, R});
    if (general) {
        d_bin_large(S, Wk, 1, Blk{r, f, R});
        cl.sync();
    if (stop_phase == 8u) return;   // profiling aid (RXC_FRONT_STOP): leave after this many phases
        d_list_sort_warp(Wk, 0, tiles_per_frame, Blk{r, f, R});
    }
}

#endif  // !__CUDACC_RTC__

// ---------------------------------------------------------------------------------------------
// k_raster
// ---------------------------------------------------------------------------------------------
#ifndef RX_RASTER_MIN_BLOCKS
#define RX_RASTER_MIN_BLOCKS 4  // resident CTAs per SM the register allocation is bounded for
#endif
#ifndef RX_OPAQUE
#define RX_OPAQUE 2
#endif
#ifndef RX_BULK_STORE
#define RX_BULK_STORE 0   // 1: finished tiles leave shared memory through the bulk-copy (TMA) engine, 32 x cp.async.bulk of one
                          // 128 B row per tile.  Measured slower than 256 x STG.128 (map 4K 1.41 -> 1.50 ms): off.
#endif
#ifndef RX_PACKED_SHADE
#define RX_PACKED_SHADE 0       // deferred shade of two adjacent pixels of one triangle per packed fp32 instruction (shade_owner_pair, needs
                                // RX_DX = 1).  Measured and rejected: the instruction count did not move (map 4K 1.173e9 vs 1.178e9 warp
                                // instructions) and the kernel slowed down 11 % on top of the RX_DX = 1 layout's own loss
#endif
#ifndef RX_PACKED_FRAGMENTS
#define RX_PACKED_FRAGMENTS 0   // barycentrics / depth of two pixels per packed fp32 instruction (test_fragment_pair: FFMA2 / FADD2, bit-exact --
                                // all GPU parity tests pass with it).  Measured and rejected: -3.7 % instructions but only -1.9 % time on map 4K
                                // when every covered pair takes it (teapot +3 %, dense 8K +2 %: half-covered pairs do double work), and
                                // +8 % on map 4K when half-covered pairs fall back to the scalar test (two code paths, 32 B more spills)
#endif
#ifndef RX_DEPTH_CULL
#define RX_DEPTH_CULL 0         // fast mode: records that lie entirely behind an opaque record covering the warp's whole region are dropped
                                // before the per-pixel tests (conservative bounds of the interpolated 1/z, see region_iz_bounds).  Correct
                                // (128 GPU parity tests, owner and depth bit for bit) and measured slower: the bounds cost every record what the
                                // culled sky fragments save -- map 4K 1.312 -> 1.320 ms, dense 8K 0.828 -> 0.839, teapot 0.213 -> 0.210
#endif
// ---- scene / frame specialisation (rx_jit.cu): a kernel recompiled for one scene knows these as constants ------------------
// RX_SPEC_FLAGS_KNOWN / _VALUE: shade-descriptor flag bits that are the same on every batch; RX_SPEC_NLIGHTS; RX_SPEC_LIGHT_TYPE
// (every light has this type); RX_SPEC_FRAME_KNOWN / _VALUE over the RX_FS_* bits below.  Undefined = read at run time.
#define RX_FS_AMBIENT 1u
#define RX_FS_SUN 2u
#define RX_FS_SKY_OR_BRUSH 4u
#define RX_FS_D3 8u
#define RX_FS_D2 16u
#define RX_FS_SECTORS 32u
#ifndef RX_SPEC_FLAGS_KNOWN
#define RX_SPEC_FLAGS_KNOWN 0u
#define RX_SPEC_FLAGS_VALUE 0u
#endif
#ifndef RX_SPEC_FRAME_KNOWN
#define RX_SPEC_FRAME_KNOWN 0u
#define RX_SPEC_FRAME_VALUE 0u
#endif
#ifdef RX_SPEC_NO_ALPHA           // no texture of the scene has a texel with alpha < 255 and no program writes opacity: no alpha test anywhere
#define RX_K_META_ALPHA(m) 0u
#else
#define RX_K_META_ALPHA(m) ((m) & RX_META_ALPHA)
#endif
#define RX_K_FLAGS(x) (((x) & ~RX_SPEC_FLAGS_KNOWN) | RX_SPEC_FLAGS_VALUE)
#define RX_K_FS(bit, x) ((RX_SPEC_FRAME_KNOWN & (bit)) ? ((RX_SPEC_FRAME_VALUE & (bit)) != 0u) : (bool)(x))
#ifdef RX_SPEC_NLIGHTS
#define RX_K_NLIGHTS(x) ((uint32_t)RX_SPEC_NLIGHTS)
#else
#define RX_K_NLIGHTS(x) (x)
#endif
#ifdef RX_SPEC_LIGHT_TYPE
#define RX_K_LIGHT_TYPE(x) ((uint32_t)RX_SPEC_LIGHT_TYPE)
#else
#define RX_K_LIGHT_TYPE(x) (x)
#endif

#define RX_LARGE_CACHE 160   // large-triangle records kept in shared memory across the tiles of a frame
#define RX_COLOR_STRIDE 40   // words per tile row in shared memory: the 4 rows a warp writes hit disjoint banks

// Visibility state of the 2x2 pixels of a thread: pixel k is (px0 + RX_DX*(k&1), py0 + 4*(k>>1)).
struct Vis4 {
    float z[4];
    uint32_t own[4];     // owner slot
    float al[4], be[4];  // barycentrics of the owner at the pixel
};

// ---- shading-only fast math --------------------------------------------------------------------
// Everything below feeds only the final RGBA8 of a pixel whose owner is already decided (coverage,
// depth and the alpha test are exact).  The parity bar for colour is +-1 LSB, so shading uses the
// SFU approximations (rsqrt/rcp/sqrt, <= 2 ulp) and explicit FMAs.
__device__ __forceinline__ float fast_rcp(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float fast_sqrt(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float fast_rsqrt(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float fdot3(f3 a, f3 b) { return __fmaf_rn(a.x, b.x, __fmaf_rn(a.y, b.y, a.z * b.z)); }
__device__ __forceinline__ f3 fnormalize3(f3 a) { const float i = fast_rsqrt(fdot3(a, a)); return {a.x * i, a.y * i, a.z * i}; }
// smoothstep(end, start, x) with inv_range = 1/(start - end)  (light.rs:674-677)
__device__ __forceinline__ float fsmooth(const DLight& l, float x) {
    const float t = rx_clamp((x - l.end_distance) * l.inv_range, 0.0f, 1.0f);
    return t * t * __fmaf_rn(-2.0f, t, 3.0f);
}

// f32_to_u8_saturated (lib.rs:65-68) for an already decided owner: one multiply and a saturating
// round-to-nearest conversion.  Differs from trunc(fma(x, 255, 0.5)) only when x*255 rounds onto k+0.5.
__device__ __forceinline__ uint32_t fast_u8(float x) {
    uint32_t r;
    asm("cvt.rni.sat.u8.f32 %0, %1;" : "=r"(r) : "f"(x * 255.0f));
    return r;
}
// linear_to_srgb_fast (rasterizer.rs:28-33: 1.055 s - 0.055 s^2 with s = sqrt(x)) and f32_to_u8_saturated in one: the 255 is folded into
// the polynomial, s * (269.025 - 14.025 s), then the saturating round-to-nearest conversion -- sqrt, FMA, multiply, convert
__device__ __forceinline__ uint32_t fast_srgb_u8(float x) {
    const float s = fast_sqrt(x);
    uint32_t r;
    asm("cvt.rni.sat.u8.f32 %0, %1;" : "=r"(r) : "f"(s * __fmaf_rn(-14.025f, s, 269.025f)));
    return r;
}
// Texture::sample_nearest (texture.rs:307-323) for shading: floor(u*(W-1) + 0.5) instead of round() and
// a saturating clamp; the exact version (rx_sample_tex) stays on the alpha-test path, which decides ownership.
__device__ __forceinline__ float fast_wrap(float u, bool repeat) { return repeat ? (u - floorf(u)) : __saturatef(u); }
__device__ __forceinline__ uint32_t sample_nearest_fast(const uint32_t* __restrict__ tex, int W, int H, float u, float v,
                                                        bool repeat_x, bool repeat_y) {
    u = fast_wrap(u, repeat_x);
    v = fast_wrap(v, repeat_y);
    const int tx = __float2int_rd(__fmaf_rn(u, (float)(W - 1), 0.5f));  // u in [0, 1] -> [0, W-1]; NaN -> 0
    const int ty = __float2int_rd(__fmaf_rn(v, (float)(H - 1), 0.5f));
    return __ldg(tex + ty * W + tx);
}
// Texture::sample_linear (texture.rs:414-460) for shading: same taps and weights, FMA lerps, channels
// rounded by the saturating conversion.  Only the owner of a pixel is shaded, and a fragment owns a pixel only if
// its exactly sampled texel has alpha 255 (rasterizer.rs:1408): the alpha channel is not filtered again, it is 255.
__device__ __forceinline__ uint32_t sample_linear_fast(const uint32_t* __restrict__ tex, int W, int H, float u, float v,
                                                       bool repeat_x, bool repeat_y) {
    u = fast_wrap(u, repeat_x);
    v = fast_wrap(v, repeat_y);
    const float x = u * (float)(W - 1), y = v * (float)(H - 1);
    const float fx = floorf(x), fy = floorf(y);
    const int x0 = __float2int_rz(fx), y0 = __float2int_rz(fy);  // in [0, W-1]; NaN -> 0
    const int x1 = min(x0 + 1, W - 1), y1 = min(y0 + 1, H - 1);
    const float dx = x - fx, dy = y - fy;
    const uint32_t c00 = __ldg(tex + y0 * W + x0), c10 = __ldg(tex + y0 * W + x1);
    const uint32_t c01 = __ldg(tex + y1 * W + x0), c11 = __ldg(tex + y1 * W + x1);
    uint32_t out = 0xFF000000u;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const float v00 = (float)((c00 >> (8 * i)) & 0xFF), v10 = (float)((c10 >> (8 * i)) & 0xFF);
        const float v01 = (float)((c01 >> (8 * i)) & 0xFF), v11 = (float)((c11 >> (8 * i)) & 0xFF);
        const float a = __fmaf_rn(dx, v10 - v00, v00), b = __fmaf_rn(dx, v11 - v01, v01);
        uint32_t c;
        asm("cvt.rni.sat.u8.f32 %0, %1;" : "=r"(c) : "f"(__fmaf_rn(dy, b - a, a)));
        out |= c << (8 * i);
    }
    return out;
}

// What the deferred shade reads of the frame, staged in shared memory once per frame (DFrame lives in
// global memory: every field read there costs an address move and a load per pixel).
struct __align__(16) ShadeConst {
    float s2w[16];
    float cam[3];
    uint32_t has_ambient;
    float ambient[3];
    uint32_t n_lights;
    float sun_l[3];          // unit vector towards the sun
    float sun_radiance;      // 0 = no sun
    float brush[4];          // brush preview: position, radius (rasterizer.rs:13-17)
    float brush_falloff;
    uint32_t has_brush;
    uint32_t pad[2];
};
#define RX_SMEM_LIGHTS 16   // lights of the frame staged in shared memory (more: read from global memory)

// Rasterizer::screen_to_world exactly as the reference composes it (rasterizer.rs:1707-1727): used where the
// world position decides ownership (alpha test of a terrain texel); the shade uses the folded DFrame::s2w.
__device__ __forceinline__ f3 screen_to_world_exact(const DFrame& F, float x, float y, float z) {
    const float x_ndc = 2.0f * (x / F.width_f) - 1.0f;
    const float y_ndc = 1.0f - 2.0f * (y / F.height_f);
    f4 vs = rx_matvec4(F.inv_proj, {x_ndc, y_ndc, z, 1.0f}, F.matvec_mode);
    vs = {vs.x / vs.w, vs.y / vs.w, vs.z / vs.w, vs.w / vs.w};
    const f4 ws = rx_matvec4(F.inv_view, vs, F.matvec_mode);
    return {ws.x, ws.y, ws.z};
}

// Chunk::sample_terrain_texture with scale 1 (chunk.rs:135-151) + Texture::get_pixel (texture.rs:527-538)
__device__ __forceinline__ uint32_t terrain_sample(const uint8_t* __restrict__ arena, uint32_t tex_word, uint32_t wh, const DChunkInfo& ci,
                                                   float wx, float wy) {
    const int W = (int)(wh & 0xFFFFu), H = (int)(wh >> 16);
    const float local_x = (wx / 1.0f) - (float)ci.origin_x, local_y = (wy / 1.0f) - (float)ci.origin_y;
    const int ppt = W / ci.size;  // size != 0 is validated at upload
    const float pixel_x = local_x * (float)ppt, pixel_y = local_y * (float)ppt;
    const uint32_t px = rx_as_u32(rx_clamp(floorf(pixel_x), 0.0f, (float)W - 1.0f));
    const uint32_t py = rx_as_u32(rx_clamp(floorf(pixel_y), 0.0f, (float)H - 1.0f));
    const uint32_t x = min(px, (uint32_t)(W - 1)), y = min(py, (uint32_t)(H - 1));
    return __ldg(reinterpret_cast<const uint32_t*>(arena) + tex_word + y * (uint32_t)W + x);
}

// Brush preview on a terrain texel (rasterizer.rs:1193-1212 in d3_rasterize, :1601-1620 in d3_rasterize_opacity): inside
// the brush radius the RGB channels go 20..80 % of the way to white, `as u8` truncation.  Alpha is untouched, so the
// alpha test (ownership) never sees it.  bp = brush position xyz, radius.
__device__ __forceinline__ uint32_t brush_terrain(uint32_t texel, f3 world, const float* bp, float brush_falloff) {
    const f3 dv = {world.x - bp[0], world.y - bp[1], world.z - bp[2]};
    const float dist = sqrtf(rx_dot3(dv, dv));
    if (!(dist < bp[3])) return texel;
    const float nd = dist / bp[3];
    const float fade = rx_clamp((1.0f - nd) / rx_clamp(brush_falloff, 0.001f, 1.0f), 0.0f, 1.0f);
    const float blend = 0.2f + 0.6f * fade;
    uint32_t out = texel & 0xFF000000u;
#pragma unroll
    for (int i = 0; i < 3; ++i)
        out |= rx_as_u8(fminf((float)((texel >> (8 * i)) & 0xFFu) * (1.0f - blend) + 255.0f * blend, 255.0f)) << (8 * i);
    return out;
}

// first sector containing the point (chunk.rs:154-161, mini.rs:58-65, bbox.rs:35-40); chunk < 0 = the mapmini
__device__ __forceinline__ float sector_occlusion(const SceneDev& S, int chunk, float x, float y) {
    const DChunkInfo ci = S.chunk_info[chunk >= 0 ? (uint32_t)chunk : S.n_scene_chunks];
    for (uint32_t i = 0; i < ci.n_sectors; ++i) {
        const DSector b = S.sectors[ci.sector_off + i];
        if (x >= b.minx && x <= b.maxx && y >= b.miny && y <= b.maxy) return b.occlusion;
    }
    return 1.0f;
}

// opacity-layer state of the thread's 2x2 pixels (general mode): nearest opacity fragment so far and the
// surface id it wrote (rasterizer.rs:1647-1651)
struct Opa4 {
    float z[4];
    uint32_t own[4];
    uint32_t sid[4];
    uint32_t some;   // bit k: sid[k] is Some(..)
};

// barycentrics and depth of a fragment, the arithmetic of test_fragment (rasterizer.rs:1754-1773, :1054-1056)
__device__ __forceinline__ float fragment_depth(const float4 q0, const float4 q1, const float4 q2, uint32_t meta, float fpx, float fpy,
                                                float* alpha_out, float* beta_out) {
    const float acx = q1.x - q0.x, acy = q1.y - q0.y;
    const float apx = fpx - q0.x, apy = fpy - q0.y;
    const float pcx = q1.x - fpx, pcy = q1.y - fpy, pbx = q0.z - fpx, pby = q0.w - fpy;
    const float na = pcx * pby - pcy * pbx, nb = acx * apy - acy * apx;
    float alpha, beta;
    if (meta & RX_META_FASTDIV) { alpha = rx_div_by(na, q2.x, q1.z); beta = rx_div_by(nb, q2.x, q1.z); }
    else { alpha = na / q2.x; beta = nb / q2.x; }
    const float gamma = 1.0f - alpha - beta;
    const float one_over_z = q2.y * alpha + q2.z * beta + q2.w * gamma;
    *alpha_out = alpha; *beta_out = beta;
    return 1.0f / one_over_z;
}

// CompiledLight::radiance_at for the 3D path (light.rs:504-653): incoming colour times Lambert for
// positional lights.  `ldir`/`dist` are the unit vector and distance from the point to the light.
__device__ __forceinline__ bool light_radiance_fast(const DLight& l, float n_dot_l, f3 ldir, float dist, f3* out) {
    if (!l.emitting) return false;
    float att;      // scalar applied to the light colour
    bool lambert = true;
    switch (RX_K_LIGHT_TYPE(l.light_type)) {
        case RXC_LIGHT_POINT:
            if (dist >= l.end_distance) return false;
            att = l.intensity * l.flicker_factor;
            if (!(dist <= l.start_distance)) att *= fsmooth(l, dist);
            break;
        case RXC_LIGHT_AMBIENT:
        case RXC_LIGHT_AMBIENT_DAYLIGHT:
            att = l.intensity * l.flicker_factor;
            lambert = false;
            break;
        case RXC_LIGHT_SPOT: {
            if (dist >= l.end_distance) return false;
            // 1 - (dist - start)/(end - start)
            const float a = (dist <= l.start_distance) ? 1.0f : __fmaf_rn(dist - l.start_distance, l.inv_range, 1.0f);
            // direction_to_point = -ldir; angle = acos(dir . direction_to_point) > cone_angle -> None
            const float c = -(l.dx * ldir.x + l.dy * ldir.y + l.dz * ldir.z);
            // acos(c) > cone_angle  <=>  c < cos(cone_angle) for c in [-1, 1]; outside, acos is NaN and the light passes
            if (c < l.width && c >= -1.0f && c <= 1.0f) return false;
            att = l.intensity * a * l.flicker_factor;
            break;
        }
        case RXC_LIGHT_AREA: {
            if (dist >= l.end_distance) return false;
            if (dist < 0.1f) { att = 1.0f; break; }
            const float d = (dist <= l.start_distance) ? 1.0f : fsmooth(l, dist);
            const float area = l.width * l.height;
            if (l.from_linedef) att = d * area * l.intensity;
            else att = fmaxf(-(l.nx * ldir.x + l.ny * ldir.y + l.nz * ldir.z), 0.0f) * d * area * l.intensity;
            break;
        }
        default: {  // Daylight: no Lambert term (light.rs:513-519)
            if (dist >= l.end_distance) return false;
            const float d = (dist <= l.start_distance) ? 1.0f : fsmooth(l, dist);
            att = fmaxf(-(l.nx * ldir.x + l.ny * ldir.y + l.nz * ldir.z), 0.0f) * d * l.intensity;
            lambert = false;
            break;
        }
    }
    if (lambert) att *= n_dot_l;  // light.rs:529-532
    *out = {l.cr * att, l.cg * att, l.cb * att};
    return true;
}

// Texture::sample (texture.rs:203-460) through the per-(frame,batch) descriptor of DFrameBatch.
__device__ __forceinline__ uint32_t sample_desc(const uint8_t* __restrict__ arena, uint32_t tex_word, uint32_t wh, uint32_t flags,
                                                float u, float v, uint32_t sample_mode) {
    return rx_sample_tex(reinterpret_cast<const uint32_t*>(arena) + tex_word, (int)(wh & 0xFFFFu), (int)(wh >> 16), u, v,
                         sample_mode, (flags & RX_SD_REPEAT_X) != 0u, (flags & RX_SD_REPEAT_Y) != 0u);
}

// rasterizer.rs:1062-1404 + :1875-1951 for the owning fragment of a pixel; returns RGBA8.
// K, lights and kd_lut are in shared memory (kd_lut[c] = srgb_to_linear_fast(c / 255) * 0.96).
__device__ __forceinline__ uint32_t shade_owner(const SceneDev& S, const ShadeConst& K, const DLight* lights, const float* kd_lut,
                                                const DFrameBatch& FB, const TriShade* __restrict__ shp, float alpha, float beta,
                                                float z, float fpx, float fpy, uint32_t sample_mode) {
    const float4* sq = reinterpret_cast<const float4*>(shp);
    const float4 s0 = __ldg(sq), s1 = __ldg(sq + 1), s2 = __ldg(sq + 2);
    const uint4 d0 = __ldg(reinterpret_cast<const uint4*>(&FB.sd_tex_word));
    const float4 d1 = __ldg(reinterpret_cast<const float4*>(&FB.sd_ambient[0]));
    const uint32_t flags = RX_K_FLAGS(d0.z);
    const float gamma = 1.0f - alpha - beta;

    uint32_t texel = d0.w;
    if (flags & RX_SD_TEXTURED) {
        // perspective-correct UV, rasterizer.rs:1062-1076.  The owner is decided; the quotients are
        // faithful (rcp + one residual correction) instead of div.rn.
        // The three sums are the reference's own unfused operations: on a triangle that crosses the near plane the terms
        // cancel, and a fused sum then differs from the reference's by far more than an ulp -- on a minified noise texture
        // that is another texel (fuzz seed 20675: 32 % of the frame).
        const float iu = s0.x * alpha + s0.z * beta + s1.x * gamma;
        const float iv = s0.y * alpha + s0.w * beta + s1.y * gamma;
        const float irw = s1.z * alpha + s1.w * beta + s2.x * gamma;
        const float rr = fast_rcp(irw);
        float u = iu * rr, v = iv * rr;
        u = __fmaf_rn(__fmaf_rn(-irw, u, iu), rr, u);
        v = __fmaf_rn(__fmaf_rn(-irw, v, iv), rr, v);
        const uint32_t* tex = reinterpret_cast<const uint32_t*>(S.arena) + d0.x;
        const int W = (int)(d0.y & 0xFFFFu), H = (int)(d0.y >> 16);
        if (sample_mode == RXC_SAMPLE_NEAREST) texel = sample_nearest_fast(tex, W, H, u, v, (flags & RX_SD_REPEAT_X) != 0u, (flags & RX_SD_REPEAT_Y) != 0u);
        else texel = sample_linear_fast(tex, W, H, u, v, (flags & RX_SD_REPEAT_X) != 0u, (flags & RX_SD_REPEAT_Y) != 0u);
    }

    // screen_to_world, rasterizer.rs:1707-1727, folded into one projective map (DFrame::s2w)
    const float4 m0 = *reinterpret_cast<const float4*>(&K.s2w[0]), m1 = *reinterpret_cast<const float4*>(&K.s2w[4]);
    const float4 m2 = *reinterpret_cast<const float4*>(&K.s2w[8]), m3 = *reinterpret_cast<const float4*>(&K.s2w[12]);
    const float4 kc = *reinterpret_cast<const float4*>(&K.cam[0]), ka = *reinterpret_cast<const float4*>(&K.ambient[0]);
    const float hx = __fmaf_rn(m0.z, z, __fmaf_rn(m0.y, fpy, __fmaf_rn(m0.x, fpx, m0.w)));
    const float hy = __fmaf_rn(m1.z, z, __fmaf_rn(m1.y, fpy, __fmaf_rn(m1.x, fpx, m1.w)));
    const float hz = __fmaf_rn(m2.z, z, __fmaf_rn(m2.y, fpy, __fmaf_rn(m2.x, fpx, m2.w)));
    const float hw = __fmaf_rn(m3.z, z, __fmaf_rn(m3.y, fpy, __fmaf_rn(m3.x, fpx, m3.w)));
    const float ihw = fast_rcp(hw);
    const f3 world = {hx * ihw, hy * ihw, hz * ihw};
    const f3 view_dir = fnormalize3({kc.x - world.x, kc.y - world.y, kc.z - world.z});

    if (flags & RX_SD_TERRAIN) {
        texel = terrain_sample(S.arena, d0.x, d0.y, S.chunk_info[__float_as_int(d1.w)], world.x, world.z);
        if (RX_K_FS(RX_FS_SKY_OR_BRUSH, K.has_brush) && K.has_brush) texel = brush_terrain(texel, world, K.brush, K.brush_falloff);
    }

    f3 normal = {0.0f, 0.0f, 0.0f};
    if (flags & RX_SD_NORMALS) {  // rasterizer.rs:1083-1099
        const float4 s4 = __ldg(sq + 4);
        if (__float_as_uint(s4.z) != 0u) {
            normal = {s2.y, s2.z, s2.w};  // the three vertex normals are equal: n0 holds the unit normal (make_tri)
        } else {
            const float4 s3 = __ldg(sq + 3);
            normal = {__fmaf_rn(s3.w, gamma, __fmaf_rn(s3.x, beta, s2.y * alpha)),
                      __fmaf_rn(s4.x, gamma, __fmaf_rn(s3.y, beta, s2.z * alpha)),
                      __fmaf_rn(s4.y, gamma, __fmaf_rn(s3.z, beta, s2.w * alpha))};
            normal = fnormalize3(normal);
        }
        if (fdot3(normal, view_dir) < 0.0f) normal = {-normal.x, -normal.y, -normal.z};
    } else {
        normal = fnormalize3(normal);  // Vec3::zero().normalized() is NaN in the reference (:1320)
    }

    // roughness 0.5, metallic 0 (no batch shader): f0 = 0.04, kd = srgb_to_linear_fast(texel) * 0.96 (rasterizer.rs:20-25),
    // shininess = 2/0.25 - 2 = 6
    const f3 kd = {kd_lut[texel & 0xFFu], kd_lut[(texel >> 8) & 0xFFu], kd_lut[(texel >> 16) & 0xFFu]};
    const float hemi = __fmaf_rn(0.5f, normal.y, 0.5f);
    f3 amb = {d1.x, d1.y, d1.z};                                                        // :1368-1370
    const float4 ks = *reinterpret_cast<const float4*>(&K.sun_l[0]);
    float occ = 1.0f;  // :1327-1365: the sky and sun terms are scaled by the sector occlusion (0 when it is not > 0)
    const bool has_ambient = RX_K_FS(RX_FS_AMBIENT, K.has_ambient), has_sun = RX_K_FS(RX_FS_SUN, ks.w > 0.0f);
    if ((has_ambient || has_sun) && RX_K_FS(RX_FS_SECTORS, S.n_sectors)) { const float o = sector_occlusion(S, __float_as_int(d1.w), world.x, world.z); occ = o > 0.0f ? o : 0.0f; }
    if (has_ambient) amb = {__fmaf_rn(ka.x, occ, amb.x), __fmaf_rn(ka.y, occ, amb.y), __fmaf_rn(ka.z, occ, amb.z)};
    f3 lit = {amb.x * kd.x * hemi, amb.y * kd.y * hemi, amb.z * kd.z * hemi};

    const float n_dot_v = fmaxf(fdot3(normal, view_dir), 0.0f);
    const float om = 1.0f - fminf(n_dot_v, 1.0f);
    const float om2 = om * om;
    const float fr = __fmaf_rn(1.0f - 0.04f, om2 * om2 * om, 0.04f);  // schlick_fresnel with f0 = 0.04 (:1882-1887)
    if (has_sun) {  // directional sun, rasterizer.rs:1342-1361
        const f3 ldir = {ks.x, ks.y, ks.z};
        const float n_dot_l = fmaxf(fdot3(normal, ldir), 0.0f);
        if (n_dot_l > 0.0f) {
            const f3 h = fnormalize3(rx_add3(ldir, view_dir));
            const float n_dot_h = fmaxf(fdot3(normal, h), 0.0f);
            const float h2 = n_dot_h * n_dot_h;
            const float spec = fr * (h2 * h2 * h2);
            const float w = n_dot_l * ks.w * occ;
            lit = {__fmaf_rn(kd.x + spec, w, lit.x), __fmaf_rn(kd.y + spec, w, lit.y), __fmaf_rn(kd.z + spec, w, lit.z)};
        }
    }
    const uint32_t n_lights = RX_K_NLIGHTS(__float_as_uint(ka.w));
    for (uint32_t li = 0; li < n_lights; ++li) {  // rasterizer.rs:1373-1391
        const DLight& L = lights[li];
        const f3 to_l = {L.px - world.x, L.py - world.y, L.pz - world.z};
        const float d2 = fdot3(to_l, to_l);
        if (d2 >= L.range2) continue;   // out of range (or not emitting): before the normalisation and the type dispatch
        const float inv_d = fast_rsqrt(d2);
        const f3 ldir = {to_l.x * inv_d, to_l.y * inv_d, to_l.z * inv_d};
        const float n_dot_l = fmaxf(fdot3(normal, ldir), 0.0f);
        if (!(n_dot_l > 0.0f)) continue;   // shade_fast_brdf adds nothing then (rasterizer.rs:1912-1951), whatever the radiance
        f3 radiance;
        if (!light_radiance_fast(L, n_dot_l, ldir, d2 * inv_d, &radiance)) continue;
        const f3 h = fnormalize3(rx_add3(ldir, view_dir));
        const float n_dot_h = fmaxf(fdot3(normal, h), 0.0f);
        const float h2 = n_dot_h * n_dot_h;
        const float spec = fr * (h2 * h2 * h2);  // pow32_fast(n.h, 6) = exp2(6 log2 x), :1895-1908
        lit = {__fmaf_rn((kd.x + spec) * n_dot_l, radiance.x, lit.x), __fmaf_rn((kd.y + spec) * n_dot_l, radiance.y, lit.y),
               __fmaf_rn((kd.z + spec) * n_dot_l, radiance.z, lit.z)};
    }
    const uint32_t a8 = texel >> 24;  // f32_to_u8_saturated(a / 255) == a for every u8 a
    return fast_srgb_u8(lit.x) | (fast_srgb_u8(lit.y) << 8) | (fast_srgb_u8(lit.z) << 16) | (a8 << 24);
}

// shade_owner for TWO horizontally adjacent pixels (fpx, fpx + 1; same y) owned by the SAME triangle, in packed fp32
// (FFMA2 / FMUL2 / FADD2: one issue slot per two pixels): the same formulas in the same order as shade_owner, lane by
// lane; only the texel fetches, the table look-ups, the SFU calls and the per-light radiance stay scalar.  The caller
// takes this path for lit, textured-or-constant, non-terrain owners with vertex normals when neither the sun nor sector
// occlusion is in play (everything else goes through shade_owner pixel by pixel).  nz = -0.0f (see rx_mul2).
#if RX_PACKED_SHADE
__device__ __forceinline__ uint2 shade_owner_pair(const SceneDev& S, const ShadeConst& K, const DLight* lights, const float* kd_lut,
                                                  const DFrameBatch& FB, const TriShade* __restrict__ shp, float2 alpha, float2 beta, float2 z,
                                                  float fpx, float fpy, uint32_t sample_mode, float nz) {
    const float4* sq = reinterpret_cast<const float4*>(shp);
    const float4 s0 = __ldg(sq), s1 = __ldg(sq + 1), s2 = __ldg(sq + 2);
    const uint4 d0 = __ldg(reinterpret_cast<const uint4*>(&FB.sd_tex_word));
    const float4 d1 = __ldg(reinterpret_cast<const float4*>(&FB.sd_ambient[0]));
    const uint32_t flags = d0.z;
    const float2 gamma = rx_sub2(rx_sub2(rx_bc2(1.0f), alpha), beta);
    const float2 X = make_float2(fpx, fpx + 1.0f);

    uint32_t texel0 = d0.w, texel1 = d0.w;
    if (flags & RX_SD_TEXTURED) {   // the reference's own unfused sums (see shade_owner), two pixels per instruction
        const float2 iu = rx_add2(rx_add2(rx_mul2(alpha, rx_bc2(s0.x), nz), rx_mul2(beta, rx_bc2(s0.z), nz)), rx_mul2(gamma, rx_bc2(s1.x), nz));
        const float2 iv = rx_add2(rx_add2(rx_mul2(alpha, rx_bc2(s0.y), nz), rx_mul2(beta, rx_bc2(s0.w), nz)), rx_mul2(gamma, rx_bc2(s1.y), nz));
        const float2 irw = rx_add2(rx_add2(rx_mul2(alpha, rx_bc2(s1.z), nz), rx_mul2(beta, rx_bc2(s1.w), nz)), rx_mul2(gamma, rx_bc2(s2.x), nz));
        const float2 rr = make_float2(fast_rcp(irw.x), fast_rcp(irw.y));
        const float2 nirw = make_float2(-irw.x, -irw.y);
        float2 u = rx_fma2(iu, rr, rx_bc2(-0.0f)), v = rx_fma2(iv, rr, rx_bc2(-0.0f));
        u = rx_fma2(rx_fma2(nirw, u, iu), rr, u);
        v = rx_fma2(rx_fma2(nirw, v, iv), rr, v);
        const uint32_t* tex = reinterpret_cast<const uint32_t*>(S.arena) + d0.x;
        const int W = (int)(d0.y & 0xFFFFu), H = (int)(d0.y >> 16);
        const bool rpx = (flags & RX_SD_REPEAT_X) != 0u, rpy = (flags & RX_SD_REPEAT_Y) != 0u;
        if (sample_mode == RXC_SAMPLE_NEAREST) { texel0 = sample_nearest_fast(tex, W, H, u.x, v.x, rpx, rpy); texel1 = sample_nearest_fast(tex, W, H, u.y, v.y, rpx, rpy); }
        else { texel0 = sample_linear_fast(tex, W, H, u.x, v.x, rpx, rpy); texel1 = sample_linear_fast(tex, W, H, u.y, v.y, rpx, rpy); }
    }

    // screen_to_world folded into one projective map (DFrame::s2w)
    const float4 m0 = *reinterpret_cast<const float4*>(&K.s2w[0]), m1 = *reinterpret_cast<const float4*>(&K.s2w[4]);
    const float4 m2 = *reinterpret_cast<const float4*>(&K.s2w[8]), m3 = *reinterpret_cast<const float4*>(&K.s2w[12]);
    const float4 kc = *reinterpret_cast<const float4*>(&K.cam[0]), ka = *reinterpret_cast<const float4*>(&K.ambient[0]);
    const float2 Y = rx_bc2(fpy);
    const float2 hx = rx_fma2(rx_bc2(m0.z), z, rx_fma2(rx_bc2(m0.y), Y, rx_fma2(rx_bc2(m0.x), X, rx_bc2(m0.w))));
    const float2 hy = rx_fma2(rx_bc2(m1.z), z, rx_fma2(rx_bc2(m1.y), Y, rx_fma2(rx_bc2(m1.x), X, rx_bc2(m1.w))));
    const float2 hz = rx_fma2(rx_bc2(m2.z), z, rx_fma2(rx_bc2(m2.y), Y, rx_fma2(rx_bc2(m2.x), X, rx_bc2(m2.w))));
    const float2 hw = rx_fma2(rx_bc2(m3.z), z, rx_fma2(rx_bc2(m3.y), Y, rx_fma2(rx_bc2(m3.x), X, rx_bc2(m3.w))));
    const float2 ihw = make_float2(fast_rcp(hw.x), fast_rcp(hw.y));
    const float2 zero2 = rx_bc2(-0.0f);
    const float2 wx = rx_fma2(hx, ihw, zero2), wy = rx_fma2(hy, ihw, zero2), wz = rx_fma2(hz, ihw, zero2);
    // view_dir = normalize(cam - world)
    float2 vx = rx_sub2(rx_bc2(kc.x), wx), vy = rx_sub2(rx_bc2(kc.y), wy), vz = rx_sub2(rx_bc2(kc.z), wz);
    {
        const float2 dd = rx_fma2(vx, vx, rx_fma2(vy, vy, rx_fma2(vz, vz, zero2)));
        const float2 iv2 = make_float2(fast_rsqrt(dd.x), fast_rsqrt(dd.y));
        vx = rx_fma2(vx, iv2, zero2); vy = rx_fma2(vy, iv2, zero2); vz = rx_fma2(vz, iv2, zero2);
    }
    // normal (rasterizer.rs:1083-1099); the caller guarantees RX_SD_NORMALS
    float2 nx, ny, nzv;
    {
        const float4 s4 = __ldg(sq + 4);
        if (__float_as_uint(s4.z) != 0u) {
            nx = rx_bc2(s2.y); ny = rx_bc2(s2.z); nzv = rx_bc2(s2.w);
        } else {
            const float4 s3 = __ldg(sq + 3);
            nx = rx_fma2(rx_bc2(s3.w), gamma, rx_fma2(rx_bc2(s3.x), beta, rx_fma2(alpha, rx_bc2(s2.y), zero2)));
            ny = rx_fma2(rx_bc2(s4.x), gamma, rx_fma2(rx_bc2(s3.y), beta, rx_fma2(alpha, rx_bc2(s2.z), zero2)));
            nzv = rx_fma2(rx_bc2(s4.y), gamma, rx_fma2(rx_bc2(s3.z), beta, rx_fma2(alpha, rx_bc2(s2.w), zero2)));
            const float2 nn = rx_fma2(nx, nx, rx_fma2(ny, ny, rx_fma2(nzv, nzv, zero2)));
            const float2 inn = make_float2(fast_rsqrt(nn.x), fast_rsqrt(nn.y));
            nx = rx_fma2(nx, inn, zero2); ny = rx_fma2(ny, inn, zero2); nzv = rx_fma2(nzv, inn, zero2);
        }
        const float2 ndv = rx_fma2(nx, vx, rx_fma2(ny, vy, rx_fma2(nzv, vz, zero2)));
        const float2 sg = make_float2(ndv.x < 0.0f ? -1.0f : 1.0f, ndv.y < 0.0f ? -1.0f : 1.0f);   // face the camera
        nx = rx_fma2(nx, sg, zero2); ny = rx_fma2(ny, sg, zero2); nzv = rx_fma2(nzv, sg, zero2);
    }
    const float2 kdr = make_float2(kd_lut[texel0 & 0xFFu], kd_lut[texel1 & 0xFFu]);
    const float2 kdg = make_float2(kd_lut[(texel0 >> 8) & 0xFFu], kd_lut[(texel1 >> 8) & 0xFFu]);
    const float2 kdb = make_float2(kd_lut[(texel0 >> 16) & 0xFFu], kd_lut[(texel1 >> 16) & 0xFFu]);
    const float2 hemi = rx_fma2(rx_bc2(0.5f), ny, rx_bc2(0.5f));
    f3 amb = {d1.x, d1.y, d1.z};
    if (K.has_ambient) amb = {ka.x + amb.x, ka.y + amb.y, ka.z + amb.z};   // occlusion is 1 on this path
    float2 lr = rx_fma2(rx_fma2(kdr, rx_bc2(amb.x), zero2), hemi, zero2);
    float2 lg = rx_fma2(rx_fma2(kdg, rx_bc2(amb.y), zero2), hemi, zero2);
    float2 lb = rx_fma2(rx_fma2(kdb, rx_bc2(amb.z), zero2), hemi, zero2);

    float2 fr;
    {
        const float2 d = rx_fma2(nx, vx, rx_fma2(ny, vy, rx_fma2(nzv, vz, zero2)));
        const float2 om = make_float2(1.0f - fminf(fmaxf(d.x, 0.0f), 1.0f), 1.0f - fminf(fmaxf(d.y, 0.0f), 1.0f));
        const float2 om2 = rx_fma2(om, om, zero2);
        fr = rx_fma2(rx_bc2(1.0f - 0.04f), rx_fma2(rx_fma2(om2, om2, zero2), om, zero2), rx_bc2(0.04f));
    }
    const uint32_t n_lights = __float_as_uint(ka.w);
    for (uint32_t li = 0; li < n_lights; ++li) {  // rasterizer.rs:1373-1391
        const DLight& L = lights[li];
        const float2 tx = rx_sub2(rx_bc2(L.px), wx), ty = rx_sub2(rx_bc2(L.py), wy), tz = rx_sub2(rx_bc2(L.pz), wz);
        const float2 d2 = rx_fma2(tx, tx, rx_fma2(ty, ty, rx_fma2(tz, tz, zero2)));
        const float range2 = L.range2;
        bool a0 = d2.x < range2, a1 = d2.y < range2;
        if (!(a0 || a1)) continue;
        const float2 inv_d = make_float2(fast_rsqrt(d2.x), fast_rsqrt(d2.y));
        const float2 lx = rx_fma2(tx, inv_d, zero2), ly = rx_fma2(ty, inv_d, zero2), lz = rx_fma2(tz, inv_d, zero2);
        const float2 ndl_raw = rx_fma2(nx, lx, rx_fma2(ny, ly, rx_fma2(nzv, lz, zero2)));
        const float2 ndl = make_float2(fmaxf(ndl_raw.x, 0.0f), fmaxf(ndl_raw.y, 0.0f));
        a0 = a0 && ndl.x > 0.0f; a1 = a1 && ndl.y > 0.0f;
        if (!(a0 || a1)) continue;
        f3 r0 = {0.0f, 0.0f, 0.0f}, r1 = {0.0f, 0.0f, 0.0f};   // a pixel the light does not reach adds exactly nothing
        const float2 dist = rx_fma2(d2, inv_d, zero2);
        if (a0 && !light_radiance_fast(L, ndl.x, {lx.x, ly.x, lz.x}, dist.x, &r0)) r0 = {0.0f, 0.0f, 0.0f};
        if (a1 && !light_radiance_fast(L, ndl.y, {lx.y, ly.y, lz.y}, dist.y, &r1)) r1 = {0.0f, 0.0f, 0.0f};
        float2 hx2 = rx_add2(lx, vx), hy2 = rx_add2(ly, vy), hz2 = rx_add2(lz, vz);
        const float2 hh = rx_fma2(hx2, hx2, rx_fma2(hy2, hy2, rx_fma2(hz2, hz2, zero2)));
        const float2 ih = make_float2(fast_rsqrt(hh.x), fast_rsqrt(hh.y));
        hx2 = rx_fma2(hx2, ih, zero2); hy2 = rx_fma2(hy2, ih, zero2); hz2 = rx_fma2(hz2, ih, zero2);
        const float2 ndh_raw = rx_fma2(nx, hx2, rx_fma2(ny, hy2, rx_fma2(nzv, hz2, zero2)));
        const float2 ndh = make_float2(fmaxf(ndh_raw.x, 0.0f), fmaxf(ndh_raw.y, 0.0f));
        const float2 h2 = rx_fma2(ndh, ndh, zero2);
        const float2 spec = rx_fma2(fr, rx_fma2(rx_fma2(h2, h2, zero2), h2, zero2), zero2);
        lr = rx_fma2(rx_fma2(rx_add2(kdr, spec), ndl, zero2), make_float2(r0.x, r1.x), lr);
        lg = rx_fma2(rx_fma2(rx_add2(kdg, spec), ndl, zero2), make_float2(r0.y, r1.y), lg);
        lb = rx_fma2(rx_fma2(rx_add2(kdb, spec), ndl, zero2), make_float2(r0.z, r1.z), lb);
    }
    auto l2s2 = [&](float2 x) {   // linear_to_srgb_fast, rasterizer.rs:28-33, then * 255 for the saturating conversion
        const float2 sq2 = make_float2(fast_sqrt(x.x), fast_sqrt(x.y));
        const float2 r = rx_fma2(rx_fma2(sq2, rx_bc2(-0.055f), zero2), sq2, rx_fma2(sq2, rx_bc2(1.055f), zero2));
        return rx_fma2(r, rx_bc2(255.0f), zero2);
    };
    const float2 cr = l2s2(lr), cg = l2s2(lg), cb = l2s2(lb);
    auto u8 = [](float x) { uint32_t r; asm("cvt.rni.sat.u8.f32 %0, %1;" : "=r"(r) : "f"(x)); return r; };
    return make_uint2(u8(cr.x) | (u8(cg.x) << 8) | (u8(cb.x) << 16) | (texel0 & 0xFF000000u),
                      u8(cr.y) | (u8(cg.y) << 8) | (u8(cb.y) << 16) | (texel1 & 0xFF000000u));
}
#endif

// ---- Rusteria VM batch shaders (general mode) -------------------------------------------------------
// Everything a program can observe is derived with the reference's own arithmetic (exact divisions, no
// approximations): a program may quantise its inputs (floor, step, pattern lookups), which would amplify
// the +-1 ulp of the fast shading path into visible differences.

__device__ __forceinline__ float srgb_to_linear_exact(float x) { const float x2 = x * x; return (0.6975f * x2 + 0.3025f) * x; }  // rasterizer.rs:20-25
__device__ __forceinline__ float linear_to_srgb_exact(float x) { const float s = sqrtf(x); return 1.055f * s - 0.055f * s * s; }   // :28-33
__device__ __forceinline__ uint32_t pack_pixel(float r, float g, float b, float a) {  // vec4_to_pixel, lib.rs:72-79
    return rx_f32_to_u8_saturated(r) | (rx_f32_to_u8_saturated(g) << 8) | (rx_f32_to_u8_saturated(b) << 16) | (rx_f32_to_u8_saturated(a) << 24);
}

// What a 3D fragment hands to its program, derived with the reference's own arithmetic (rasterizer.rs:1062-1222 for the opaque
// pass, :1500-1600 for the opacity pass): perspective-correct uv, world position, interpolated / flipped normal, texel.
struct Frag3D { float u, v; f3 world, normal; uint32_t texel; };
__device__ __forceinline__ Frag3D vm_inputs_3d(const SceneDev& S, const DFrame& F, const DFrameBatch& FB, const TriShade& sh, float alpha, float beta,
                                               float z, float fpx, float fpy, uint32_t sample_mode, bool opacity_pass) {
    Frag3D g;
    const float gamma = 1.0f - alpha - beta;
    float u = sh.uw0 * alpha + sh.uw1 * beta + sh.uw2 * gamma;
    float v = sh.vw0 * alpha + sh.vw1 * beta + sh.vw2 * gamma;
    const float irw = sh.rw0 * alpha + sh.rw1 * beta + sh.rw2 * gamma;
    u = u / irw; v = v / irw;
    const f3 world = screen_to_world_exact(F, fpx, fpy, z);
    f3 normal = {0.0f, 0.0f, 0.0f};
    if (!opacity_pass && (FB.sd_flags & RX_SD_NORMALS)) {  // :1083-1099
        if (__float_as_uint(sh.pad0) != 0u) {
            normal = {sh.n0x, sh.n0y, sh.n0z};  // flat triangle: make_tri stored the unit normal
        } else {
            normal = rx_normalize3({sh.n0x * alpha + sh.n1x * beta + sh.n2x * gamma, sh.n0y * alpha + sh.n1y * beta + sh.n2y * gamma,
                                    sh.n0z * alpha + sh.n1z * beta + sh.n2z * gamma});
        }
        const f3 view_dir = rx_normalize3({F.cam[0] - world.x, F.cam[1] - world.y, F.cam[2] - world.z});
        if (rx_dot3(normal, view_dir) < 0.0f) normal = {-normal.x, -normal.y, -normal.z};
    }
    uint32_t texel = FB.sd_pixel;
    if (FB.sd_flags & RX_SD_TEXTURED) texel = sample_desc(S.arena, FB.sd_tex_word, FB.sd_wh, FB.sd_flags, u, v, sample_mode);
    else if (FB.sd_flags & RX_SD_TERRAIN) {
        texel = terrain_sample(S.arena, FB.sd_tex_word, FB.sd_wh, S.chunk_info[FB.sd_chunk], world.x, world.z);
        if (F.has_brush) { const float bp[4] = {F.brush_pos[0], F.brush_pos[1], F.brush_pos[2], F.brush_radius}; texel = brush_terrain(texel, world, bp, F.brush_falloff); }
    }
    g.u = u; g.v = v; g.world = world; g.normal = normal; g.texel = texel;
    return g;
}
__device__ __forceinline__ f3 texel_linear_exact(uint32_t texel) {   // pixel_to_vec4 (lib.rs:55-62) + srgb_to_linear_fast
    const float inv255 = 1.0f / 255.0f;
    return {srgb_to_linear_exact((float)(texel & 0xFFu) * inv255), srgb_to_linear_exact((float)((texel >> 8) & 0xFFu) * inv255),
            srgb_to_linear_exact((float)((texel >> 16) & 0xFFu) * inv255)};
}

// A 3D fragment of a batch with a VM program, up to and including the program (rasterizer.rs:1062-1317 for the
// opaque pass, :1500-1637 for the opacity pass).  Returns false when the program hit a device limit.
__device__ __noinline__ bool vm_fragment_3d(const SceneDev& S, const DFrame& F, const DFrameBatch& FB, const TriShade& sh, float alpha, float beta,
                                            float z, float fpx, float fpy, uint32_t sample_mode, bool opacity_pass, VmIO* io, f3* world_out) {
    const Frag3D g = vm_inputs_3d(S, F, FB, sh, alpha, beta, z, fpx, fpy, sample_mode, opacity_pass);
    *world_out = g.world;
    vm_io_reset(*io);
    io->color = texel_linear_exact(g.texel);
    io->opacity.x = (float)(g.texel >> 24) / 255.0f;
    io->normal = g.normal;
    io->uv = {g.u / 4.0f, g.v / 4.0f, 0.0f};
    io->hitpoint = g.world;
    io->time = {F.time, F.time, F.time};
    if (FB.sd_program < 0 || (uint32_t)FB.sd_program >= S.vm.n_programs) return true;
    return vm_run(S.vm, S.vm.programs[FB.sd_program], *io);
}

// the opaque 3D pass behind a program: lighting with the material the program produced (rasterizer.rs:1319-1404).
// What the program READS is exact (vm_fragment_3d); what happens to its outputs afterwards only feeds the pixel's RGBA8 and is
// continuous in them, so it runs in the shading-only fast arithmetic of shade_owner (rsqrt / ex2 / lg2 approximations, FMAs;
// +-1 LSB): shade_fast_brdf (:1912-1951) with the per-fragment terms (f0, kd, shininess, Fresnel) hoisted out of the light loop.
__device__ __forceinline__ uint32_t vm_light_3d(const SceneDev& S, const DFrame& F, const DLight* lights, const DFrameBatch& FB, const VmIO& io, f3 world) {
    const f3 base = io.color;
    const f3 normal = fnormalize3(io.normal);
    const float rough = rx_clamp(io.roughness.x, 0.0f, 1.0f), metal = rx_clamp(io.metallic.x, 0.0f, 1.0f);
    const f3 view_dir = fnormalize3({F.cam[0] - world.x, F.cam[1] - world.y, F.cam[2] - world.z});
    // shade_fast_brdf, the terms that do not depend on the light
    const f3 f0 = {__fmaf_rn(metal, base.x - 0.04f, 0.04f), __fmaf_rn(metal, base.y - 0.04f, 0.04f), __fmaf_rn(metal, base.z - 0.04f, 0.04f)};
    const float kscale = (1.0f - metal) * (1.0f - fmaxf(f0.x, fmaxf(f0.y, f0.z)));
    const f3 kd_b = {base.x * kscale, base.y * kscale, base.z * kscale};
    const float shininess = rx_clamp(__fmaf_rn(2.0f, fast_rcp(fmaxf(rough * rough, 1e-4f)), -2.0f), 1.0f, 2048.0f);
    const float n_dot_v = fmaxf(fdot3(normal, view_dir), 0.0f);
    const float om = 1.0f - fminf(n_dot_v, 1.0f);
    const float om2 = om * om, x5 = om2 * om2 * om;
    const f3 fr = {__fmaf_rn(1.0f - f0.x, x5, f0.x), __fmaf_rn(1.0f - f0.y, x5, f0.y), __fmaf_rn(1.0f - f0.z, x5, f0.z)};
    auto brdf = [&](f3 ldir, float n_dot_l, f3 radiance, f3& acc) {   // (kd n.l + F spec n.l) * radiance, n.l > 0
        const f3 h = fnormalize3(rx_add3(ldir, view_dir));
        const float n_dot_h = fmaxf(fdot3(normal, h), 0.0f);
        float spec = 0.0f;
        if (n_dot_h > 0.0f) {
            float lg, ex;
            asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"(n_dot_h));
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"(shininess * lg));
            spec = ex;
        }
        acc = {__fmaf_rn(__fmaf_rn(fr.x, spec, kd_b.x) * n_dot_l, radiance.x, acc.x), __fmaf_rn(__fmaf_rn(fr.y, spec, kd_b.y) * n_dot_l, radiance.y, acc.y),
               __fmaf_rn(__fmaf_rn(fr.z, spec, kd_b.z) * n_dot_l, radiance.z, acc.z)};
    };
    f3 lit = {0.0f, 0.0f, 0.0f};
    const float occlusion = S.n_sectors ? sector_occlusion(S, FB.sd_chunk, world.x, world.z) : 1.0f;
    const float hemi = __fmaf_rn(0.5f, normal.y, 0.5f);
    const float ka = (1.0f - metal) * (1.0f - 0.04f);
    const f3 kd = {base.x * ka, base.y * ka, base.z * ka};
    if (occlusion > 0.0f) {
        if (F.has_ambient) lit = {F.ambient[0] * kd.x * hemi, F.ambient[1] * kd.y * hemi, F.ambient[2] * kd.z * hemi};
        if (F.sun_radiance > 0.0f) {  // :1342-1361
            const f3 ldir = {F.sun_l[0], F.sun_l[1], F.sun_l[2]};
            const float n_dot_l = fmaxf(fdot3(normal, ldir), 0.0f);
            if (n_dot_l > 0.0f) brdf(ldir, n_dot_l, {F.sun_radiance, F.sun_radiance, F.sun_radiance}, lit);
        }
        lit = {lit.x * occlusion, lit.y * occlusion, lit.z * occlusion};
    }
    lit = {__fmaf_rn(FB.sd_ambient[0] * kd.x, hemi, lit.x), __fmaf_rn(FB.sd_ambient[1] * kd.y, hemi, lit.y), __fmaf_rn(FB.sd_ambient[2] * kd.z, hemi, lit.z)};
    for (uint32_t li = 0; li < S.n_lights; ++li) {   // :1373-1391, the loop of shade_owner
        const DLight& L = lights[li];
        const f3 to_l = {L.px - world.x, L.py - world.y, L.pz - world.z};
        const float d2 = fdot3(to_l, to_l);
        if (d2 >= L.range2) continue;
        const float inv_d = fast_rsqrt(d2);
        const f3 ldir = {to_l.x * inv_d, to_l.y * inv_d, to_l.z * inv_d};
        const float n_dot_l = fmaxf(fdot3(normal, ldir), 0.0f);
        if (!(n_dot_l > 0.0f)) continue;
        f3 radiance;
        if (!light_radiance_fast(L, n_dot_l, ldir, d2 * inv_d, &radiance)) continue;
        brdf(ldir, n_dot_l, radiance, lit);
    }
    lit = rx_add3(lit, io.emissive);
    return fast_srgb_u8(lit.x) | (fast_srgb_u8(lit.y) << 8) | (fast_srgb_u8(lit.z) << 16) | (rx_f32_to_u8_saturated(io.opacity.x) << 24);
}
__device__ __noinline__ uint32_t shade_owner_vm(const SceneDev& S, const DFrame& F, const DLight* lights, const DFrameBatch& FB, const TriShade& sh,
                                                float alpha, float beta, float z, float fpx, float fpy, uint32_t sample_mode, uint32_t* fault) {
    VmIO io;
    f3 world;
    if (!vm_fragment_3d(S, F, FB, sh, alpha, beta, z, fpx, fpy, sample_mode, false, &io, &world)) *fault = 1u;
    return vm_light_3d(S, F, lights, FB, io, world);
}

// 2D fragment behind a program (rasterizer.rs:760-797): sRGB texel in, colour out, alpha forced to 1
__device__ __noinline__ uint32_t shade_2d_vm(const SceneDev& S, const DFrame& F, int program, uint32_t texel, float u, float v, float wx, float wy,
                                             uint32_t* fault, VmIO* carried = nullptr, VmPersist* ps = nullptr) {
    if (program < 0 || (uint32_t)program >= S.vm.n_programs || S.vm.programs[program].n_words == 0u) return texel;
    VmIO fresh;
    VmIO& io = carried ? *carried : fresh;   // carried: the tile's never-reset Execution (k_raster_ordered); the assignments below are the reference's
    if (!carried) vm_io_reset(io);
    const float inv255 = 1.0f / 255.0f;
    io.color = {(float)(texel & 0xFFu) * inv255, (float)((texel >> 8) & 0xFFu) * inv255, (float)((texel >> 16) & 0xFFu) * inv255};
    io.uv.x = u / 4.0f; io.uv.y = v / 4.0f;
    io.hitpoint.x = wx; io.hitpoint.y = wy;
    io.time = {F.time, F.time, F.time};
    io.roughness.x = 0.5f; io.metallic.x = 0.0f;
    const bool ok = ps ? vm_run_t<true>(S.vm, S.vm.programs[program], io, ps) : vm_run(S.vm, S.vm.programs[program], io);
    if (!ok) *fault = 1u;
    return pack_pixel(io.color.x, io.color.y, io.color.z, 1.0f);
}

// The opacity layer's pixel (rasterizer.rs:1500-1645): texel -> linear -> sRGB, no lighting; alpha = texel alpha.
// Barycentrics are recomputed from the owner's record with the arithmetic of the visibility pass.
template <bool VM>
__device__ __forceinline__ uint32_t shade_opacity(const SceneDev& S, const DFrame& F, const DFrameBatch* __restrict__ fbs,
                                                  const TriVis* __restrict__ vis, const TriShade* __restrict__ shade, uint32_t owner,
                                                  float fpx, float fpy, uint32_t sample_mode, uint32_t* fault) {
    const float4* q = reinterpret_cast<const float4*>(vis + owner);
    const float4 q0 = __ldg(q), q1 = __ldg(q + 1), q2 = __ldg(q + 2);
    const uint32_t meta = __ldg(&vis[owner].meta);
    const DFrameBatch& FB = fbs[meta & RX_META_BATCH];
    float alpha, beta;
    const float z = fragment_depth(q0, q1, q2, meta, fpx, fpy, &alpha, &beta);
    if (VM && (FB.sd_flags & RX_SD_SHADER)) {  // rasterizer.rs:1611-1645: the program's colour and opacity, no lighting
        VmIO io;
        f3 world;
        if (!vm_fragment_3d(S, F, FB, shade[owner], alpha, beta, z, fpx, fpy, sample_mode, true, &io, &world)) *fault = 1u;
        return pack_pixel(linear_to_srgb_exact(io.color.x), linear_to_srgb_exact(io.color.y), linear_to_srgb_exact(io.color.z), io.opacity.x);
    }
    const float gamma = 1.0f - alpha - beta;
    uint32_t texel = FB.sd_pixel;
    if (FB.sd_flags & RX_SD_TEXTURED) {
        const TriShade& sh = shade[owner];
        const float iu = sh.uw0 * alpha + sh.uw1 * beta + sh.uw2 * gamma;
        const float iv = sh.vw0 * alpha + sh.vw1 * beta + sh.vw2 * gamma;
        const float irw = sh.rw0 * alpha + sh.rw1 * beta + sh.rw2 * gamma;
        texel = sample_desc(S.arena, FB.sd_tex_word, FB.sd_wh, FB.sd_flags, iu / irw, iv / irw, sample_mode);
    } else if (FB.sd_flags & RX_SD_TERRAIN) {
        const float hx = __fmaf_rn(F.s2w[2], z, __fmaf_rn(F.s2w[1], fpy, __fmaf_rn(F.s2w[0], fpx, F.s2w[3])));
        const float hz = __fmaf_rn(F.s2w[10], z, __fmaf_rn(F.s2w[9], fpy, __fmaf_rn(F.s2w[8], fpx, F.s2w[11])));
        const float hw = __fmaf_rn(F.s2w[14], z, __fmaf_rn(F.s2w[13], fpy, __fmaf_rn(F.s2w[12], fpx, F.s2w[15])));
        texel = terrain_sample(S.arena, FB.sd_tex_word, FB.sd_wh, S.chunk_info[FB.sd_chunk], hx / hw, hz / hw);
        if (F.has_brush) {
            const float hy = __fmaf_rn(F.s2w[6], z, __fmaf_rn(F.s2w[5], fpy, __fmaf_rn(F.s2w[4], fpx, F.s2w[7])));
            const float bp[4] = {F.brush_pos[0], F.brush_pos[1], F.brush_pos[2], F.brush_radius};
            texel = brush_terrain(texel, {hx / hw, hy / hw, hz / hw}, bp, F.brush_falloff);
        }
    }
    auto rt = [](uint32_t c) {  // linear_to_srgb_fast(srgb_to_linear_fast(c / 255)), rasterizer.rs:19-33
        const float x = (float)c * (1.0f / 255.0f);
        const float l = __fmaf_rn(0.6975f, x * x, 0.3025f) * x;
        const float r = fast_sqrt(l);
        return rx_f32_to_u8_saturated(__fmaf_rn(-0.055f * r, r, 1.055f * r));
    };
    return rt(texel & 0xFF) | (rt((texel >> 8) & 0xFF) << 8) | (rt((texel >> 16) & 0xFF) << 16) | (texel & 0xFF000000u);
}

// "Blend Opacity", rasterizer.rs:464-495: src-over of the opacity layer onto the resolved pixel
__device__ __forceinline__ uint32_t blend_opacity(uint32_t src, uint32_t dst, bool preserve_transparency) {
    const float src_a = (float)(src >> 24) / 255.0f, dst_a = (float)(dst >> 24) / 255.0f;
    const float inv_a = 1.0f - src_a;
    uint32_t out = 0;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const float sc = (float)((src >> (8 * i)) & 0xFF), dc = (float)((dst >> (8 * i)) & 0xFF);
        out |= rx_as_u8(rx_clamp(sc * src_a + dc * inv_a, 0.0f, 255.0f)) << (8 * i);
    }
    const float out_a = !preserve_transparency ? 1.0f : rx_clamp(src_a + dst_a * inv_a, 0.0f, 1.0f);
    return out | (rx_as_u8(rx_clamp(out_a * 255.0f, 0.0f, 255.0f)) << 24);
}

// A pixel no geometry covered when a Sky node or a brush preview is active (rasterizer.rs:409-461): screen_ray
// (:1844-1870), ShapeFX Sky render_miss_d3 without its cloud layer (shapefx.rs:1122-1170), brush preview (:434-456).
__device__ __noinline__ uint32_t miss_color(const DFrame* __restrict__ Fp, int px, int py) {
    const DFrame& F = *Fp;
    const float x = (float)px, y = (float)py;   // the reference passes the pixel corner here, not its centre
    const float ndc_x = 2.0f * (x / F.width_f) - 1.0f, ndc_y = 1.0f - 2.0f * (y / F.height_f);
    f4 vn = rx_matvec4(F.inv_proj, {ndc_x, ndc_y, -1.0f, 1.0f}, F.matvec_mode);
    f4 vf = rx_matvec4(F.inv_proj, {ndc_x, ndc_y, 1.0f, 1.0f}, F.matvec_mode);
    vn = {vn.x / vn.w, vn.y / vn.w, vn.z / vn.w, vn.w / vn.w};
    vf = {vf.x / vf.w, vf.y / vf.w, vf.z / vf.w, vf.w / vf.w};
    const f4 wn = rx_matvec4(F.inv_view, vn, F.matvec_mode), wf = rx_matvec4(F.inv_view, vf, F.matvec_mode);
    const f3 origin = {wn.x, wn.y, wn.z};
    const f3 dir = rx_normalize3({wf.x - wn.x, wf.y - wn.y, wf.z - wn.z});
    float c[4] = {0.0f, 0.0f, 0.0f, 1.0f};
    auto lerp1 = [](float from, float to, float t) { return __fmaf_rn(rx_clamp(t, 0.0f, 1.0f), to - from, from); };  // vek lerp
    if (F.has_sky) {
        const float day_factor = F.sky[0][3];
        const float up = rx_clamp(dir.y, -1.0f, 1.0f);
        const float t = (up + 1.0f) * 0.5f;
        const float om = 1.0f - up;
        const float haze = om * om * om;
        for (int i = 0; i < 4; ++i) {
            const float v = lerp1(lerp1(F.sky[4][i], F.sky[5][i], t), lerp1(F.sky[2][i], F.sky[3][i], t), day_factor);
            const float fog = F.sky[1][i] * haze * 0.3f;
            c[i] = v * (1.0f - haze * 0.2f) + fog;
        }
        if (day_factor > 0.0f) {
            const float d = rx_clamp(rx_dot3(dir, {F.sky[0][0], F.sky[0][1], F.sky[0][2]}), -1.0f, 1.0f);
            const float dist = fmaxf(1.0f - d, 0.0f);
            if (dist < 0.04f) {
                const float k = 1.0f - dist / 0.04f;
                const float glare = k * k * (3.0f - 2.0f * k);
                c[0] += 1.0f * glare * day_factor; c[1] += 0.85f * glare * day_factor; c[2] += 0.6f * glare * day_factor;
                c[3] += 0.0f * glare * day_factor;
            }
        }
    }
    if (F.has_brush && fabsf(dir.y) > 1e-5f) {
        const float t = -origin.y / dir.y;
        if (t > 0.0f) {
            const f3 world = {origin.x + dir.x * t, origin.y + dir.y * t, origin.z + dir.z * t};
            const f3 dv = {world.x - F.brush_pos[0], world.y - F.brush_pos[1], world.z - F.brush_pos[2]};
            const float dist = sqrtf(rx_dot3(dv, dv));
            if (dist < F.brush_radius) {
                const float nd = dist / F.brush_radius;
                const float fade = rx_clamp((1.0f - nd) / rx_clamp(F.brush_falloff, 0.001f, 1.0f), 0.0f, 1.0f);
                const float blend = 0.2f + 0.6f * fade;
                for (int i = 0; i < 3; ++i) c[i] = fminf(c[i] * (1.0f - blend) + blend, 1.0f);
            }
        }
    }
    return rx_f32_to_u8_saturated(c[0]) | (rx_f32_to_u8_saturated(c[1]) << 8) | (rx_f32_to_u8_saturated(c[2]) << 16) |
           (rx_f32_to_u8_saturated(c[3]) << 24);
}

// src/shader/vgradient.rs:11-14 and src/shader/grid.rs:36-108
__device__ uint32_t shade_background(const DFrame& F, int px, int py) {
    const float uvx = (float)px / F.width_f, uvy = (float)py / F.height_f;  // rasterizer.rs:296-302
    if (F.bg_shader == RXC_BG_VGRAY_GRADIENT) {
        const uint32_t i = rx_as_u8(rx_clamp(uvy * 128.0f, 0.0f, 128.0f));
        return i | (i << 8) | (i << 16) | 0xFF000000u;
    }
    auto pix = [](float g) { uint32_t c = rx_f32_to_u8_saturated(g); return c | (c << 8) | (c << 16) | (rx_f32_to_u8_saturated(1.0f) << 24); };
    const float gsz = F.grid_size, sdiv = F.grid_subdiv;
    const float posx = uvx * F.width_f, posy = uvy * F.height_f;
    const float orgx = F.width_f / 2.0f + F.grid_off[0], orgy = F.height_f / 2.0f + F.grid_off[1];
    const float aox = roundf(orgx - 0.5f) + 0.5f, aoy = roundf(orgy - 0.5f) + 0.5f;
    const float rpx = posx - aox, rpy = posy - aoy;
    auto mul_dist = [](float delta, float value) { return fabsf(value - delta * roundf(value / delta)); };
    const float dx = mul_dist(gsz, rpx), dy = mul_dist(gsz, rpy);
    if (fminf(dx, dy) <= 1.0f * 0.5f) return pix(0.15f);
    const float dfx = fabsf(rpx - gsz * floorf(rpx / gsz)), dfy = fabsf(rpy - gsz * floorf(rpy / gsz));
    const float ssz = gsz / roundf(sdiv);
    float sdx = mul_dist(ssz, dfx), sdy = mul_dist(ssz, dfy);
    const float rcx = roundf(dx / ssz), rcy = roundf(dy / ssz);
    const float extra = gsz - ssz * sdiv;
    if (rcx == sdiv) sdx = sdx + extra;
    if (rcy == sdiv) sdy = sdy + extra;
    if (fminf(sdx, sdy) <= 1.0f * 0.5f) return pix(0.11f);
    return pix(0.05f);
}

// MapMini::is_visible (mini.rs:67-95): false when the segment from -> to crosses any linedef
__device__ __forceinline__ bool los_visible(const SceneDev& S, float fx, float fy, float tx, float ty) {
    for (uint32_t i = 0; i < S.n_linedefs; ++i) {
        const float4 l = __ldg(S.linedefs + i);  // b1 = (l.x, l.y), b2 = (l.z, l.w)
        const float d = (tx - fx) * (l.w - l.y) - (ty - fy) * (l.z - l.x);
        if (d == 0.0f) continue;
        const float u = ((l.x - fx) * (l.w - l.y) - (l.y - fy) * (l.z - l.x)) / d;
        const float v = ((l.x - fx) * (ty - fy) - (l.y - fy) * (tx - fx)) / d;
        if (u >= 0.0f && u <= 1.0f && v >= 0.0f && v <= 1.0f) return false;
    }
    return true;
}

// one 2D triangle fragment: rasterizer.rs:655-895.  `color` is the tile buffer pixel (RGBA8).
template <bool VM>
__device__ __forceinline__ uint32_t shade_2d(const SceneDev& S, const DFrame& F, const DLight* __restrict__ lights, const Tri2D& T,
                                             const DBatch2& B, const DFrameBatch2& FB, int px, int py, float fpx, float fpy,
                                             uint32_t sample_mode, uint32_t color, uint32_t* fault, VmIO* carried = nullptr, VmPersist* ps = nullptr) {
    // barycentric_weights_2d, rasterizer.rs:1731-1750
    const float acx = T.cx - T.ax, acy = T.cy - T.ay, abx = T.bx - T.ax, aby = T.by - T.ay;
    const float apx = fpx - T.ax, apy = fpy - T.ay, pcx = T.cx - fpx, pcy = T.cy - fpy, pbx = T.bx - fpx, pby = T.by - fpy;
    const float area = acx * aby - acy * abx;
    const float w0 = (pcx * pby - pcy * pbx) / area;
    const float w1 = (acx * apy - acy * apx) / area;
    const float w2 = 1.0f - w0 - w1;
    const float u = T.u0 * w0 + T.u1 * w1 + T.u2 * w2;
    const float v = T.v0 * w0 + T.v1 * w1 + T.v2 * w2;
    // rasterizer.rs:664-670
    const float gx = (float)px - F.width_f / 2.0f - (F.trans2d[0] - F.width_f / 2.0f);
    const float gy = (float)py - F.height_f / 2.0f - (F.trans2d[1] - F.height_f / 2.0f);
    const float wx = gx / F.scale2d, wy = gy / F.scale2d;

    uint32_t texel = 0u;
    if (FB.tex != 0xFFFFFFFFu) {
        const DTex tx = S.tex[FB.tex];
        if (FB.terrain) texel = terrain_sample(S.arena, (uint32_t)(tx.offset >> 2), tx.width | (tx.height << 16), S.chunk_info[B.chunk], wx, wy);
        else texel = rx_sample(S.arena, tx, u, v, sample_mode, B.repeat_mode);
    } else if (B.source_kind == RXC_SRC_PIXEL) {
        texel = B.source_pixel;
    }

    if (VM && FB.program >= 0) texel = shade_2d_vm(S, F, FB.program, texel, u, v, wx, wy, fault, carried, ps);  // rasterizer.rs:760-797

    if (FB.lit) {  // rasterizer.rs:799-873
        float acc[3] = {0.0f, 0.0f, 0.0f};
        const float occlusion = S.n_sectors ? sector_occlusion(S, B.chunk, wx, wy) : 1.0f;  // :806-812
        if (F.has_ambient) { acc[0] += F.ambient[0] * occlusion; acc[1] += F.ambient[1] * occlusion; acc[2] += F.ambient[2] * occlusion; }
        for (uint32_t li = 0; li < S.n_lights; ++li) {
            const DLight& L = lights[li];
            f3 lc;
            if (!rx_light_color_at(L, {wx, 0.0f, wy}, true, &lc)) continue;
            if (L.light_type == RXC_LIGHT_AMBIENT_DAYLIGHT) { lc.x *= occlusion; lc.y *= occlusion; lc.z *= occlusion; }  // :826-836
            if (L.light_type != RXC_LIGHT_AMBIENT && L.light_type != RXC_LIGHT_AMBIENT_DAYLIGHT && S.n_linedefs &&
                !los_visible(S, wx, wy, L.px, L.pz))  // :838-846, light.position_2d() = (x, z)
                continue;
            acc[0] += lc.x; acc[1] += lc.y; acc[2] += lc.z;
        }
        uint32_t out = texel & 0xFF000000u;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const float a = rx_clamp(acc[i], 0.0f, 1.0f);
            const float t = (float)((texel >> (8 * i)) & 0xFF);
            out |= rx_as_u8(rx_clamp((t / 255.0f) * a * 255.0f, 0.0f, 255.0f)) << (8 * i);
        }
        texel = out;
    }
    const uint32_t ta = texel >> 24;  // rasterizer.rs:876-895
    if (ta == 255u) return texel;
    const float src_alpha = (float)ta / 255.0f, dst_alpha = 1.0f - src_alpha;
    uint32_t out = 0;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const float s = (float)((texel >> (8 * i)) & 0xFF), d = (float)((color >> (8 * i)) & 0xFF);
        out |= rx_as_u8((s * src_alpha) + (d * dst_alpha)) << (8 * i);
    }
    const uint32_t da = color >> 24;
    out |= (F.preserve_transparency ? max(da, ta) : 255u) << 24;
    return out;
}

// Is pixel (x, y) plotted by rasterize_line_bresenham (rasterizer.rs:1777-1821) for the segment
// (x0,y0) -> (x1,y1)?  With a = |dx|, b = |dy| and (u, v) the steps from the start along each axis, the
// walk (err = a - b; x-step iff 2*err > -b; y-step iff 2*err < a) visits, for a >= b, exactly
// v = floor((2*b*u + a - 1) / (2*a)) for u in [0, a), and symmetrically for b > a; the end point is excluded.
// (Checked against the serial walk for every segment of a 29x29 grid, tests/test_host_api.py.)
__device__ __forceinline__ bool line_covers(int x0, int y0, int x1, int y1, int x, int y) {
    const long long a = llabs((long long)x1 - x0), b = llabs((long long)y1 - y0);
    const long long u = x0 < x1 ? (long long)x - x0 : (long long)x0 - x;
    const long long v = y0 < y1 ? (long long)y - y0 : (long long)y0 - y;
    if (u < 0 || v < 0 || u > a || v > b || (u == a && v == b)) return false;
    if (a >= b) return a != 0 && v == (2 * b * u + a - 1) / (2 * a);
    return u == (2 * a * v + b - 1) / (2 * b);
}

// Conservative rectangle-vs-triangle overlap: bbox, then for each edge the rectangle corner where the
// edge function is largest.  The per-pixel test is `fl(a*px + b*py + c) < 0 -> outside` (edge.rs:28-36);
// its rounding error is below 3 * 2^-24 * (|a|*X + |b|*Y + |c|), so a corner value below
// -2e-6 * (|a|*X + |b|*Y + |c|) proves every pixel centre of the rectangle fails.  NaNs keep the triangle.
// Returns 0 = no pixel centre of the rectangle can be covered, 1 = maybe, 2 = the bbox contains the
// rectangle and every edge function is provably >= 0 at every pixel centre of it (the corner where it is
// smallest is above +margin), so the per-pixel edge evaluation can be skipped.
__device__ __forceinline__ uint32_t rect_overlaps(const TriVis& T, int tx0, int ty0, int tx1, int ty1) {
    const int x0 = T.bbx & 0xFFFF, x1 = T.bbx >> 16, y0 = T.bby & 0xFFFF, y1 = T.bby >> 16;
    if (x0 >= tx1 || x1 <= tx0 || y0 >= ty1 || y1 <= ty0) return 0u;
    const float xmin = (float)max(tx0, x0) + 0.5f, xmax = (float)min(tx1, x1) - 0.5f;
    const float ymin = (float)max(ty0, y0) + 0.5f, ymax = (float)min(ty1, y1) - 0.5f;
    bool inside = x0 <= tx0 && x1 >= tx1 && y0 <= ty0 && y1 >= ty1;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const float a = T.ea[i], b = T.eb[i], c = T.ec[i];
        const float ax_hi = a * (a >= 0.0f ? xmax : xmin), ax_lo = a * (a >= 0.0f ? xmin : xmax);
        const float by_hi = b * (b >= 0.0f ? ymax : ymin), by_lo = b * (b >= 0.0f ? ymin : ymax);
        const float m = (fabsf(a) * xmax + fabsf(b) * ymax + fabsf(c)) * 2e-6f;
        if (ax_hi + by_hi + c < -m) return 0u;
        inside = inside && (ax_lo + by_lo + c > m);  // false for NaN
    }
    return inside ? 2u : 1u;
}

// texel of a z-passing fragment of an alpha-tested batch (rasterizer.rs:1062-1222): exact arithmetic, it decides
// ownership (:1408).  Out of line: rare, and the visibility loop is instantiated once per pixel of the 2x2.
template <bool VM>
__device__ __noinline__ uint32_t alpha_test_texel(const SceneDev& S, const uint8_t* __restrict__ arena, const DChunkInfo* __restrict__ chunk_info,
                                                  const DFrame* __restrict__ F, const DFrameBatch* __restrict__ FB,
                                                  const TriShade* __restrict__ sh, float alpha, float beta, float z, float fpx, float fpy,
                                                  uint32_t sample_mode) {
    if (VM && (FB->sd_flags & RX_SD_VM_OPACITY)) {  // the program decides the opacity (rasterizer.rs:1398-1408)
        VmIO io;
        f3 world;
        vm_fragment_3d(S, *F, *FB, *sh, alpha, beta, z, fpx, fpy, sample_mode, false, &io, &world);  // a fault is reported by the resolve
        return rx_f32_to_u8_saturated(io.opacity.x) << 24;
    }
    if (FB->sd_flags & RX_SD_TERRAIN) {
        const f3 world = screen_to_world_exact(*F, fpx, fpy, z);
        return terrain_sample(arena, FB->sd_tex_word, FB->sd_wh, chunk_info[FB->sd_chunk], world.x, world.z);
    }
    const float gamma = 1.0f - alpha - beta;
    const float iu = sh->uw0 * alpha + sh->uw1 * beta + sh->uw2 * gamma;
    const float iv = sh->vw0 * alpha + sh->vw1 * beta + sh->vw2 * gamma;
    const float irw = sh->rw0 * alpha + sh->rw1 * beta + sh->rw2 * gamma;
    return sample_desc(arena, FB->sd_tex_word, FB->sd_wh, FB->sd_flags, iu / irw, iv / irw, sample_mode);
}

// depth + alpha test of one covered pixel (rasterizer.rs:1051-1060, :1408)
template <bool VM>
__device__ __forceinline__ void test_fragment(const SceneDev& S, const DFrame& F, const DFrameBatch* __restrict__ fbs,
                                              const TriShade* __restrict__ shade, const float4 q0, const float4 q1, const float4 q2,
                                              uint32_t meta, uint32_t slot, float fpx, float fpy, uint32_t sample_mode, float& best_z,
                                              uint32_t& best, float& best_al, float& best_be) {
    // barycentric_weights_3d, rasterizer.rs:1754-1773 (a = q0.xy, b = q0.zw, c = q1.xy)
    const float acx = q1.x - q0.x, acy = q1.y - q0.y;
    const float apx = fpx - q0.x, apy = fpy - q0.y;
    const float pcx = q1.x - fpx, pcy = q1.y - fpy, pbx = q0.z - fpx, pby = q0.w - fpy;
    const float na = pcx * pby - pcy * pbx, nb = acx * apy - acy * apx;
    float alpha, beta;
    if (meta & RX_META_FASTDIV) { alpha = rx_div_by(na, q2.x, q1.z); beta = rx_div_by(nb, q2.x, q1.z); }
    else { alpha = na / q2.x; beta = nb / q2.x; }
    const float gamma = 1.0f - alpha - beta;
    const float one_over_z = q2.y * alpha + q2.z * beta + q2.w * gamma;  // :1054-1056
    const float z = 1.0f / one_over_z;
    // sequential `z < zbuf` in submission order == lexicographic min of (z, ordinal)
    const bool pass_z = (z < best_z) || (z == best_z && best != RX_OWNER_NONE && slot < best);
    if (!pass_z) return;
    if (RX_K_META_ALPHA(meta)) {  // alpha test: texel alpha must be 255 to write (:1408)
        const uint32_t texel = alpha_test_texel<VM>(S, S.arena, S.chunk_info, &F, fbs + (meta & RX_META_BATCH), shade + slot, alpha, beta, z,
                                                         fpx, fpy, sample_mode);
        if ((texel >> 24) != 255u) return;
    }
    best_z = z; best = slot; best_al = alpha; best_be = beta;
}

// The same for the two pixels of a row of the thread's 2x2 (x = fxa and fxb, same y): the barycentric and depth
// arithmetic of both in packed fp32 (rx_fma2: FFMA2 / FMUL2 / FADD2, one issue slot for two pixels), every operation
// the same single rounding as in test_fragment.  `pm` = which of the two are covered (bit 0: a, bit 1: b).
template <bool VM>
__device__ __forceinline__ void test_fragment_pair(const SceneDev& S, const DFrame& F, const DFrameBatch* __restrict__ fbs,
                                                   const TriShade* __restrict__ shade, const float4 q0, const float4 q1, const float4 q2,
                                                   uint32_t meta, uint32_t slot, float fxa, float fxb, float fpy, uint32_t sample_mode, float nz,
                                                   float& bz0, uint32_t& bo0, float& bal0, float& bbe0, float& bz1, uint32_t& bo1, float& bal1, float& bbe1) {
    const float2 X = make_float2(fxa, fxb);
    const float acx = q1.x - q0.x, acy = q1.y - q0.y;
    const float apy = fpy - q0.y, pcy = q1.y - fpy, pby = q0.w - fpy;
    const float2 apx = rx_add2(X, rx_bc2(-q0.x));          // fpx - a.x
    const float2 pcx = rx_sub2(rx_bc2(q1.x), X);           // c.x - fpx
    const float2 pbx = rx_sub2(rx_bc2(q0.z), X);           // b.x - fpx
    const float2 na = rx_sub2(rx_mul2(pcx, rx_bc2(pby), nz), rx_mul2(pbx, rx_bc2(pcy), nz));   // pcx*pby - pcy*pbx
    const float2 nb = rx_sub2(rx_bc2(rx_mul1(acx, apy, nz)), rx_mul2(apx, rx_bc2(acy), nz));   // acx*apy - acy*apx
    float2 alpha, beta;
    if (meta & RX_META_FASTDIV) { alpha = rx_div_by2(na, q2.x, q1.z, nz); beta = rx_div_by2(nb, q2.x, q1.z, nz); }
    else { alpha = make_float2(na.x / q2.x, na.y / q2.x); beta = make_float2(nb.x / q2.x, nb.y / q2.x); }
    const float2 gamma = rx_sub2(rx_sub2(rx_bc2(1.0f), alpha), beta);
    const float2 ooz = rx_add2(rx_add2(rx_mul2(alpha, rx_bc2(q2.y), nz), rx_mul2(beta, rx_bc2(q2.z), nz)), rx_mul2(gamma, rx_bc2(q2.w), nz));  // :1054-1056
    // sequential `z < zbuf` in submission order == lexicographic min of (z, ordinal); alpha test (:1408)
    auto commit = [&](float z, float al, float be, float fx, float& bz, uint32_t& bo, float& bal, float& bbe) {
        const bool pass_z = (z < bz) || (z == bz && bo != RX_OWNER_NONE && slot < bo);
        if (!pass_z) return;
        if (meta & RX_META_ALPHA) {
            const uint32_t texel = alpha_test_texel<VM>(S, S.arena, S.chunk_info, &F, fbs + (meta & RX_META_BATCH), shade + slot, al, be, z, fx, fpy,
                                                             sample_mode);
            if ((texel >> 24) != 255u) return;
        }
        bz = z; bo = slot; bal = al; bbe = be;
    };
    commit(1.0f / ooz.x, alpha.x, beta.x, fxa, bz0, bo0, bal0, bbe0);
    commit(1.0f / ooz.y, alpha.y, beta.y, fxb, bz1, bo1, bal1, bbe1);
}

// coverage of one staged triangle over the thread's 2x2 pixels (rasterizer.rs:1020-1036), then the
// depth test of the covered ones.  `valid` masks pixels outside the frame.
// Conservative bounds [lo, hi] of the 1/z a record's covered pixels can compute (test_fragment: iz0*alpha + iz1*beta + iz2*gamma).
// The weights of a pixel that passed the edge tests sum to one and lie in [-d, 1 + d], d = the rounding of the edge functions and
// of the barycentric numerators (a few ulp of a coordinate product) relative to the area; slivers (d large) get no bounds.
__device__ __forceinline__ bool region_iz_bounds(const TriVis& T, float* lo, float* hi) {
    const float mn = fminf(T.iz0, fminf(T.iz1, T.iz2)), mx = fmaxf(T.iz0, fmaxf(T.iz1, T.iz2));
    const float mc = fmaxf(fmaxf(fmaxf(fabsf(T.ax), fabsf(T.ay)), fmaxf(fabsf(T.bx), fabsf(T.by))), fmaxf(fabsf(T.cx), fabsf(T.cy)));
    const float d = (mc * mc) * 2e-6f * fabsf(T.rarea);
    const float pad = __fmaf_rn(2.0f * d, mx - mn, mx * 2e-6f);
    *lo = mn - pad; *hi = mx + pad;
    return (mn > 0.0f) && (d < 0.05f) && (mx < 3.0e38f) && (mn - pad > 0.0f);   // false for NaN
}

template <bool GENERAL, bool VM>
__device__ __forceinline__ void process_record(const SceneDev& S, const DFrame& F, const DFrameBatch* __restrict__ fbs,
                                               const TriShade* __restrict__ shade, const TriVis* Tp, uint32_t slot, bool full, int px0,
                                               int py0, float fx0, float fy0, uint32_t valid, uint32_t sample_mode, float nz, Vis4& V, Opa4& O) {
    const float4* q = reinterpret_cast<const float4*>(Tp);
    const float4 q5 = q[5];
    const uint32_t bbx = __float_as_uint(q5.y), bby = __float_as_uint(q5.z), meta = __float_as_uint(q5.w);
    uint32_t m = valid;   // `full` (uniform over the warp): the bbox contains the warp's region and every edge passes everywhere
    if (!full) {
        const int x0 = (int)(bbx & 0xFFFFu), x1 = (int)(bbx >> 16), y0 = (int)(bby & 0xFFFFu), y1 = (int)(bby >> 16);
        const bool cx0 = px0 >= x0 && px0 < x1, cx1 = px0 + RX_DX >= x0 && px0 + RX_DX < x1;
        const bool cy0 = py0 >= y0 && py0 < y1, cy1 = py0 + 4 >= y0 && py0 + 4 < y1;
        m &= ((cx0 && cy0) ? 1u : 0u) | ((cx1 && cy0) ? 2u : 0u) | ((cx0 && cy1) ? 4u : 0u) | ((cx1 && cy1) ? 8u : 0u);
    }
    if (!m) return;
    const float fx1 = fx0 + (float)RX_DX, fy1 = fy0 + 4.0f;
    if (!full) {   // Edges::evaluate, edge.rs:28-36: (a*px + b*py) + c < 0 -> outside (a NaN result passes)
        const float4 q3 = q[3], q4 = q[4];
        const float ea[3] = {q3.x, q3.y, q3.z}, eb[3] = {q3.w, q4.x, q4.y}, ec[3] = {q4.z, q4.w, q5.x};
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const float ax0 = ea[i] * fx0, ax1 = ea[i] * fx1, by0 = eb[i] * fy0, by1 = eb[i] * fy1;
            if ((ax0 + by0) + ec[i] < 0.0f) m &= ~1u;
            if ((ax1 + by0) + ec[i] < 0.0f) m &= ~2u;
            if ((ax0 + by1) + ec[i] < 0.0f) m &= ~4u;
            if ((ax1 + by1) + ec[i] < 0.0f) m &= ~8u;
        }
        if (!m) return;
    }
    const float4 q0 = q[0], q1 = q[1], q2 = q[2];
    if (GENERAL) {
        // records arrive in submission order: the opacity layer and the surface ids evolve sequentially
        // (rasterizer.rs:314-357, :1041-1047, :1647-1651)
        const DFrameBatch& FB = fbs[meta & RX_META_BATCH];
        const uint32_t profile = FB.sd_profile;
        const bool has_profile = (FB.sd_flags & RX_SD_HAS_PROFILE) != 0u;
        if (meta & RX_META_OPACITY) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (!(m & (1u << k))) continue;
                float al, be;
                const float z = fragment_depth(q0, q1, q2, meta, (k & 1) ? fx1 : fx0, (k & 2) ? fy1 : fy0, &al, &be);
                if (z < O.z[k]) {
                    O.z[k] = z; O.own[k] = slot; O.sid[k] = profile;
                    O.some = has_profile ? (O.some | (1u << k)) : (O.some & ~(1u << k));
                }
            }
            return;
        }
        if (has_profile) {  // wall geometry behind an opacity batch of the same profile is skipped
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if ((O.some & (1u << k)) && O.sid[k] == profile) m &= ~(1u << k);
        }
    }
#if RX_PACKED_FRAGMENTS
    // pixels 0,1 share y = fy0 and pixels 2,3 y = fy1: each covered pair goes through the packed test
#pragma unroll
    for (int pr = 0; pr < 2; ++pr) {
        const uint32_t pm = (m >> (2 * pr)) & 3u;
        const float fy = pr ? fy1 : fy0;
        if (pm == 3u) {
            test_fragment_pair<VM>(S, F, fbs, shade, q0, q1, q2, meta, slot, fx0, fx1, fy, sample_mode, nz, V.z[2 * pr], V.own[2 * pr], V.al[2 * pr],
                                   V.be[2 * pr], V.z[2 * pr + 1], V.own[2 * pr + 1], V.al[2 * pr + 1], V.be[2 * pr + 1]);
        } else if (pm == 1u) {   // one pixel of the pair: the scalar test (small triangles mostly come this way)
            test_fragment<VM>(S, F, fbs, shade, q0, q1, q2, meta, slot, fx0, fy, sample_mode, V.z[2 * pr], V.own[2 * pr], V.al[2 * pr], V.be[2 * pr]);
        } else if (pm == 2u) {
            test_fragment<VM>(S, F, fbs, shade, q0, q1, q2, meta, slot, fx1, fy, sample_mode, V.z[2 * pr + 1], V.own[2 * pr + 1], V.al[2 * pr + 1],
                              V.be[2 * pr + 1]);
        }
    }
#else
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (m & (1u << k))
            test_fragment<VM>(S, F, fbs, shade, q0, q1, q2, meta, slot, (k & 1) ? fx1 : fx0, (k & 2) ? fy1 : fy0, sample_mode, V.z[k],
                          V.own[k], V.al[k], V.be[k]);
    }
#endif
}

// ---- small-triangle pass of long tile lists (fast mode) -------------------------------------------------
// A tile at the horizon of a dense mesh holds thousands of triangles whose pixel boxes are a pixel or two wide.
// Walked a record per warp (process_record: 128 pixel slots per record) such a tile is one long serial chain in the
// one or two warps that own those pixels.  Here every THREAD takes one record of the list instead: a record whose
// box, clipped to the tile, has at most RX_SMALL_MAX_PIX pixels is rasterised by that thread alone -- same edge,
// barycentric, depth and alpha arithmetic as process_record / test_fragment -- and competes for its pixels with a
// 64-bit shared-memory atomicMin on (order(z), slot).  That is the reference's sequential `z < zbuf` in submission
// order (rasterizer.rs:1051-1060): the winner is the lexicographic minimum of (z, ordinal) over the fragments that
// pass the alpha test, whatever the order they arrive in.  The other records are compacted into a shared list that
// the warp walk then reads instead of the tile's whole list.
#define RX_BIG_CAP 2048u        // compacted list of the other records (aliases the upper half of s_state)

__device__ __forceinline__ uint32_t z_order_bits(float z) {   // monotone float -> uint map (z is never NaN here)
    const uint32_t u = __float_as_uint(z);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float z_from_order_bits(uint32_t k) { return __uint_as_float((k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k); }
// c_inv_w[w] = 65536 / w + 1: p / w == (p * c_inv_w[w]) >> 16 for p < 1024 and 1 <= w <= 32 (an integer division costs ~25 instructions)
__constant__ uint32_t c_inv_w[33] = {0, 65537, 32769, 21846, 16385, 13108, 10923, 9363, 8193, 7282, 6554, 5958, 5462, 5042, 4682, 4370, 4097, 3856, 3641, 3450, 3277, 3121, 2979, 2850, 2731, 2622, 2521, 2428, 2341, 2260, 2185, 2115, 2049};
#define RX_KEY_NONE ((((unsigned long long)0xBF800000u) << 32) | 0xFFFFFFFFull)   // (order(1.0f), RX_OWNER_NONE): z_buffer starts at 1.0

// clipped pixel count of a record's box inside the tile; 0 = the record cannot touch the tile
__device__ __forceinline__ int small_box(uint32_t bbx, uint32_t bby, int tx0, int ty0, int tx1, int ty1, int* x0, int* y0, int* x1, int* y1) {
    *x0 = max((int)(bbx & 0xFFFFu), tx0); *x1 = min((int)(bbx >> 16), tx1);
    *y0 = max((int)(bby & 0xFFFFu), ty0); *y1 = min((int)(bby >> 16), ty1);
    const int w = *x1 - *x0, h = *y1 - *y0;
    return (w <= 0 || h <= 0) ? 0 : w * h;
}

__device__ __noinline__ void small_triangle_pass(const SceneDev& S, const DFrame& F, const DFrameBatch* __restrict__ fbs,
                                                 const TriVis* __restrict__ vis, const TriShade* __restrict__ shade,
                                                 const uint32_t* __restrict__ list, uint32_t n_list, int tx0, int ty0, int tx1, int ty1,
                                                 uint32_t sample_mode, int max_pix, uint32_t gshift, unsigned long long* s_key, uint32_t* s_big,
                                                 uint32_t* s_nbig) {
    // a group of 2^gshift neighbouring lanes per record: they read the same record (one transaction) and share out the
    // pixels of its box, pixel p = sub + j * group -> (x0 + p % w, y0 + p / w)
    const uint32_t group = 1u << gshift, sub = threadIdx.x & (group - 1u), per_round = RX_TILE_THREADS >> gshift;
    for (uint32_t i = threadIdx.x >> gshift; i < n_list; i += per_round) {
        const uint32_t slot = __ldg(list + i);
        const float4* q = reinterpret_cast<const float4*>(vis + slot);
        const float4 q5 = __ldg(q + 5);
        const uint32_t meta = __float_as_uint(q5.w);
        int x0, y0, x1, y1;
        const int npix = small_box(__float_as_uint(q5.y), __float_as_uint(q5.z), tx0, ty0, tx1, ty1, &x0, &y0, &x1, &y1);
        if (npix == 0) continue;
        if (npix > max_pix) {
            if (sub == 0u) {
                const uint32_t k = atomicAdd(s_nbig, 1u);
                if (k < RX_BIG_CAP) s_big[k] = slot;
            }
            continue;
        }
        const float4 q3 = __ldg(q + 3), q4 = __ldg(q + 4);
        const float4 q0 = __ldg(q), q1 = __ldg(q + 1), q2 = __ldg(q + 2);
        const uint32_t w = (uint32_t)(x1 - x0);
        const uint32_t iw = c_inv_w[w];
#pragma unroll 1
        for (uint32_t p = sub; p < (uint32_t)npix; p += group) {
            const uint32_t row = (p * iw) >> 16;
            const int x = x0 + (int)(p - row * w), y = y0 + (int)row;
            const float fpx = (float)x + 0.5f, fpy = (float)y + 0.5f;   // rasterizer.rs:1022
            // Edges::evaluate, edge.rs:28-36: (a*px + b*py) + c < 0 -> outside
            if ((q3.x * fpx + q3.w * fpy) + q4.z < 0.0f) continue;
            if ((q3.y * fpx + q4.x * fpy) + q4.w < 0.0f) continue;
            if ((q3.z * fpx + q4.y * fpy) + q5.x < 0.0f) continue;
            float al, be;
            const float z = fragment_depth(q0, q1, q2, meta, fpx, fpy, &al, &be);
            if (!(z < 1.0f)) continue;   // the first `z < zbuf` of a pixel is against 1.0; NaN fails
            const unsigned long long key = ((unsigned long long)z_order_bits(z) << 32) | slot;
            unsigned long long* cell = s_key + (y - ty0) * RX_TILE_W + (x - tx0);
            if (key >= *reinterpret_cast<volatile unsigned long long*>(cell)) continue;
            if (RX_K_META_ALPHA(meta)) {  // alpha test: texel alpha must be 255 to write (:1408)
                const uint32_t texel = alpha_test_texel<false>(S, S.arena, S.chunk_info, &F, fbs + (meta & RX_META_BATCH), shade + slot, al, be, z,
                                                               fpx, fpy, sample_mode);
                if ((texel >> 24) != 255u) continue;
            }
            atomicMin(cell, key);
        }
    }
}

// one 2D record against one pixel: triangle (rasterizer.rs:640-895) or Bresenham line (:901-955, :1777-1821)
template <bool VM>
__device__ __forceinline__ uint32_t apply_2d(const SceneDev& S, const DFrame& F, const DLight* __restrict__ lights,
                                            const DFrameBatch2* __restrict__ fb2, const Tri2D& T, int px, int py, uint32_t sample_mode,
                                            uint32_t color, uint32_t* fault, VmIO* carried = nullptr, VmPersist* ps = nullptr) {
    const int x0 = T.bbx & 0xFFFF, x1 = T.bbx >> 16, y0 = T.bby & 0xFFFF, y1 = T.bby >> 16;
    if (px < x0 || px >= x1 || py < y0 || py >= y1) return color;
    const DBatch2& B = S.b2[T.batch];
    if (T.kind != 0u) {
        if (!line_covers(__float_as_int(T.ax), __float_as_int(T.ay), __float_as_int(T.bx), __float_as_int(T.by), px, py)) return color;
        return B.source_kind == RXC_SRC_PIXEL ? B.source_pixel : 0xFFFFFFFFu;  // crate::WHITE
    }
    const float fpx = (float)px + 0.5f, fpy = (float)py + 0.5f;
    if ((T.ea[0] * fpx + T.eb[0] * fpy + T.ec[0]) < 0.0f) return color;
    if ((T.ea[1] * fpx + T.eb[1] * fpy + T.ec[1]) < 0.0f) return color;
    if ((T.ea[2] * fpx + T.eb[2] * fpy + T.ec[2]) < 0.0f) return color;
    return shade_2d<VM>(S, F, lights, T, B, fb2[T.batch], px, py, fpx, fpy, sample_mode, color, fault, carried, ps);
}

// SAMPLE: 0 nearest / 1 linear for every frame of the launch, 2 = read it per frame.  PLANES: owner/depth outputs.
// GENERAL: the tile lists are sorted by submission ordinal and hold every triangle (no large list): chunk opacity
// batches with their surface ids are evaluated sequentially, and 2D records come from sorted per-tile lists.
// MODE: 0 = fast path, 1 = general (below), 2 = general + Rusteria VM programs on batches, 3 = fast path with the
// thread-per-record pass of long tile lists (small_triangle_pass; its own instantiation because the extra code costs the
// plain fast path 6 % on the map scene -- scenes with few triangles never have such lists and keep MODE 0).
template <int SAMPLE, bool PLANES, int MODE>
__global__ void __launch_bounds__(RX_TILE_THREADS, RX_RASTER_MIN_BLOCKS) k_raster(SceneDev S, Workspace Wk, const __grid_constant__ RasterOut out, uint32_t n_frames,
                                                               uint32_t tile0, uint32_t tiles_per_frame, uint32_t counter) {
    constexpr bool GENERAL = MODE == 1 || MODE == 2, VM = MODE == 2, SMALL = MODE == 3;
    __shared__ __align__(16) TriVis s_large[GENERAL ? 1 : RX_LARGE_CACHE];
    __shared__ uint32_t s_large_slot[GENERAL ? 1 : RX_LARGE_CACHE];
    __shared__ uint16_t s_sel[GENERAL ? 1 : RX_LARGE_CACHE];
    __shared__ uint32_t s_nsel;
    __shared__ int32_t s_work[4];   // frame (-1 = done), tile x0, tile y0, tile index
    __shared__ uint32_t s_grp[2];   // frame groups: the group this CTA takes tiles from, and how many groups it has seen run dry
#if RX_TMA_STORE
    // two dense 32x32 tiles in the tensor map's 128 B swizzle (16 B chunk c of row r sits at chunk c ^ (r & 7)), one draining
    __shared__ __align__(1024) uint32_t s_color[2 * RX_TILE_H * RX_TILE_W];
#define RX_CIDX(row, col) ((row) * RX_TILE_W + (((((col) >> 2) ^ ((row) & 7)) << 2) | ((col) & 3)))
#else
    __shared__ __align__(16) uint32_t s_color[(RX_BULK_STORE ? 2 : 1) * RX_TILE_H * RX_COLOR_STRIDE];  // bulk store: two tiles, one draining
#define RX_CIDX(row, col) ((row) * RX_COLOR_STRIDE + (col))
#endif
    __shared__ float4 s_state[4 * RX_TILE_THREADS];  // (z, owner, alpha, beta) of pixel k of thread t at [k*256 + t]
    __shared__ float2 s_ostate[GENERAL ? 4 * RX_TILE_THREADS : 1];  // (z, owner) of the opacity layer
    __shared__ ShadeConst s_k;                       // frame constants of the deferred shade
    __shared__ __align__(16) DLight s_lights[RX_SMEM_LIGHTS];
    __shared__ float s_kd[256];                      // srgb_to_linear_fast(c / 255) * (1 - 0.04), rasterizer.rs:20-25
    __shared__ int s_union[8];                       // pixel bbox of the frame's cached large triangles [0..3] and of its 2D records [4..7]
    __shared__ int s_can_be_empty;                   // per frame: some tile may be untouched (see the empty-tile path below)
    __shared__ struct { const DFrame* F; const DLight* lights_g; const TriVis* vis; const TriShade* shade; const DFrameBatch* fbs; const uint32_t* large; const uint32_t* tile_count; const uint32_t* tile_base; const uint32_t* lists; uint32_t n_large; } s_p;
    __shared__ uint32_t s_nbig;                      // small-triangle pass: length of the compacted list of the other records
    __shared__ __align__(8) unsigned long long s_mbar;   // completion of the bulk copies that stage the frame's large-triangle records

    uint32_t tid = threadIdx.x;
#if RX_OPAQUE >= 3
    asm volatile("" : "+r"(tid));
#endif
    const uint32_t lane = tid & 31, warp = tid >> 5;
    // warp w covers a 16x8 region (2 across, 4 down); lane (lx, ly) of the 8x4 lane grid owns the
    // pixels (lx + RX_DX*i, ly + 4j) of the region (RX_DX = 1: lx is even, the thread owns two horizontally adjacent pairs)
    const int rbx = (int)(warp & 1) * RX_REGION_W, rby = (int)(warp >> 1) * RX_REGION_H;
    int lx = (int)(lane & 7) * (RX_DX == 1 ? 2 : 1) + rbx, ly = (int)(lane >> 3) + rby;   // pixel offset of the thread inside the tile
#if RX_OPAQUE >= 4
    asm volatile("" : "+r"(lx), "+r"(ly));
#endif
#if RX_TMA_STORE
    // the thread's first pixel in the swizzled tile; with RX_DX = 8 pixel k of its 2x2 sits at (cbase ^ (k&1)<<3 ^ (k>>1)<<4) + (k>>1)*128:
    // +8 columns flips chunk bit 1, +4 rows flips bit 2 of (row & 7) and with it chunk bit 2
    int cbase = ly * RX_TILE_W + ((((lx >> 2) ^ (ly & 7)) << 2) | (lx & 3));
#if RX_DX == 1
#define RX_CPIX(k) (((cbase + ((k) & 1)) ^ (((k) >> 1) << 4)) + ((k) >> 1) * (4 * RX_TILE_W))   // lx is even: the neighbour shares the 16 B chunk
#else
#define RX_CPIX(k) ((cbase ^ (((k) & 1) << 3) ^ (((k) >> 1) << 4)) + ((k) >> 1) * (4 * RX_TILE_W))
#endif
#else
    int cbase = ly * RX_COLOR_STRIDE + lx;   // the thread's first pixel in s_color
#define RX_CPIX(k) (cbase + ((k) & 1) * RX_DX + ((k) >> 1) * (4 * RX_COLOR_STRIDE))
#endif
#if RX_OPAQUE >= 2
    asm volatile("" : "+r"(cbase));
#endif
    const uint32_t total = n_frames * tiles_per_frame;
    uint32_t cached_frame = 0xFFFFFFFFu, n_cached = 0;
    uint32_t mbar_phase = 0u;
    if (!GENERAL && tid == 0) {   // visible to everybody (and to the async proxy) after the first tile's barrier
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&s_mbar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    uint32_t coff = 0;   // which half of s_color this tile uses (bulk store: the other half may still be draining)
    {
        const float x = (float)tid * (1.0f / 255.0f);
        s_kd[tid] = (__fmaf_rn(0.6975f, x * x, 0.3025f) * x) * (1.0f - 0.04f);  // visible after the first tile's barrier
    }
    pdl_enter();   // everything above is private to the CTA; the front end's results are read from here on

    // Frame groups (launches of many frames): what a CTA stages per frame (pointer block, shade constants, lights, the large
    // records, the union boxes: three barriers and a bulk-copy round trip) is paid per (CTA, frame) pair, and with ONE counter
    // over all tiles every CTA walks through every frame.  With G counters over G runs of consecutive frames a CTA stays with
    // the frames of its group (blockIdx % G) and only moves on, group by group, when they are done.
    const uint32_t n_groups = Wk.raster_groups > 1u ? Wk.raster_groups : 1u;
    if (tid == 0) { s_grp[0] = blockIdx.x % n_groups; s_grp[1] = 0u; }
    for (;;) {
        if (tid == 0) {
            uint32_t work = 0xFFFFFFFFu, f_base = 0u;
            if (n_groups == 1u) {
                const uint32_t w = atomicAdd(Wk.raster_counter + counter, 1u);
                if (w < total) work = w;
            } else {
                uint32_t g = s_grp[0], dry = s_grp[1];
                while (dry < n_groups) {
                    const uint32_t f0 = (uint32_t)(((unsigned long long)g * n_frames) / n_groups), f1 = (uint32_t)(((unsigned long long)(g + 1u) * n_frames) / n_groups);
                    const uint32_t w = atomicAdd(Wk.raster_counter + g, 1u);
                    if (w < (f1 - f0) * tiles_per_frame) { work = w; f_base = f0; break; }
                    g = g + 1u == n_groups ? 0u : g + 1u; ++dry;   // that group is done: help the next one
                }
                s_grp[0] = g; s_grp[1] = dry;
            }
            if (work == 0xFFFFFFFFu) {
                s_work[0] = -1;
            } else {
                const uint32_t df = work / tiles_per_frame, f = f_base + df, tile = tile0 + (work - df * tiles_per_frame);
                const uint32_t tiles_x = (uint32_t)Wk.frames[f].tiles_x;
                const uint32_t ty = tile / tiles_x;
                s_work[0] = (int32_t)f; s_work[1] = (int32_t)((tile - ty * tiles_x) * RX_TILE_W); s_work[2] = (int32_t)(ty * RX_TILE_H);
                s_work[3] = (int32_t)tile;
            }
            s_nsel = 0u;
        }
        __syncthreads();
        if (s_work[0] < 0) break;
        const uint32_t f = (uint32_t)s_work[0], tile = (uint32_t)s_work[3];
        // per-frame base pointers, staged below when the frame changes (f is uniform, but the compiler cannot know)
        if (f != cached_frame) {  // uniform for the CTA: stage what every tile of this frame reads
            const DFrame& Fg = Wk.frames[f];
            const DLight* lights_gg = Wk.lights + (size_t)f * Wk.lights_stride;
            const TriVis* visg = Wk.vis + (size_t)f * Wk.slot_stride;
            const uint32_t n_largeg = (GENERAL || !Fg.d3_active) ? 0u : min(Wk.counters[f].n_large, Wk.large_stride);
            const uint32_t* largeg = Wk.large + (size_t)f * Wk.large_stride;
            if (tid == 32) {   // the frame's base pointers: one shared-memory load per tile instead of the 64-bit index arithmetic
                s_p.F = &Fg; s_p.lights_g = lights_gg; s_p.vis = visg; s_p.shade = Wk.shade + (size_t)f * Wk.slot_stride;
                s_p.fbs = Wk.fb + (size_t)f * Wk.fb_stride; s_p.large = largeg; s_p.n_large = n_largeg;
                s_p.tile_count = Wk.tile_count + (size_t)f * Wk.tile_stride; s_p.tile_base = Wk.tile_base + (size_t)f * Wk.tile_stride;
                s_p.lists = Wk.lists + (size_t)f * Wk.list_stride;
            }
            if (tid < 16) s_k.s2w[tid] = Fg.s2w[tid];
            else if (tid < 19) s_k.cam[tid - 16] = Fg.cam[tid - 16];
            else if (tid == 19) s_k.has_ambient = Fg.has_ambient;
            else if (tid < 23) s_k.ambient[tid - 20] = Fg.ambient[tid - 20];
            else if (tid == 23) s_k.n_lights = S.n_lights;
            else if (tid < 27) s_k.sun_l[tid - 24] = Fg.sun_l[tid - 24];
            else if (tid == 27) s_k.sun_radiance = Fg.sun_radiance;
            else if (tid < 31) s_k.brush[tid - 28] = Fg.brush_pos[tid - 28];
            else if (tid == 31) s_k.brush[3] = Fg.brush_radius;
            else if (tid == 32) s_k.brush_falloff = Fg.brush_falloff;
            else if (tid == 33) s_k.has_brush = Fg.has_brush;
            for (uint32_t i = tid; i < min(S.n_lights, (uint32_t)RX_SMEM_LIGHTS) * (uint32_t)(sizeof(DLight) / 4); i += RX_TILE_THREADS)
                reinterpret_cast<uint32_t*>(s_lights)[i] = __ldg(reinterpret_cast<const uint32_t*>(lights_gg) + i);
            if (!GENERAL) {
                // The frame's large-triangle records are staged in shared memory by the bulk-copy (TMA) engine: one
                // cp.async.bulk of a 96 B record per lane, completion counted in bytes on an mbarrier the CTA then waits on
                // (every thread of the previous tile is past its reads of s_large: the barrier after the work fetch).
                n_cached = min(n_largeg, (uint32_t)RX_LARGE_CACHE);
                if (n_cached) {
                    const uint32_t mbar = (uint32_t)__cvta_generic_to_shared(&s_mbar);
                    if (tid == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(n_cached * (uint32_t)sizeof(TriVis)) : "memory");
                    for (uint32_t r = tid; r < n_cached; r += RX_TILE_THREADS) {
                        const uint32_t slot = __ldg(largeg + r);
                        s_large_slot[r] = slot;
                        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                     ::"r"((uint32_t)__cvta_generic_to_shared(&s_large[r])), "l"(visg + slot), "r"((uint32_t)sizeof(TriVis)), "r"(mbar) : "memory");
                    }
                    uint32_t done = 0u;
                    while (!done) {
                        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                                     : "=r"(done) : "r"(mbar), "r"(mbar_phase) : "memory");
                    }
                    mbar_phase ^= 1u;
                }
            }
#ifdef RX_SPEC_ACTIVE
            // A kernel compiled for one scene / frame signature (rx_jit.cu) checks, once per frame and CTA, that what it was told is
            // what it finds: a mismatch would render silently wrong pixels, so it is reported like an overflow (bit 5) instead.
            {
                bool bad = false;
                const DFrameBatch* fbg = Wk.fb + (size_t)f * Wk.fb_stride;
                for (uint32_t b = tid; b < S.n_b3; b += RX_TILE_THREADS) {
                    bad = bad || (fbg[b].sd_flags & RX_SPEC_FLAGS_KNOWN) != RX_SPEC_FLAGS_VALUE;
#ifdef RX_SPEC_NO_ALPHA
                    bad = bad || fbg[b].alpha_test != 0u;
#endif
                }
#ifdef RX_SPEC_LIGHT_TYPE
                for (uint32_t i = tid; i < S.n_lights; i += RX_TILE_THREADS) bad = bad || lights_gg[i].light_type != (uint32_t)RX_SPEC_LIGHT_TYPE;
#endif
                if (tid == 0) {
#ifdef RX_SPEC_NLIGHTS
                    bad = bad || S.n_lights != (uint32_t)RX_SPEC_NLIGHTS;
#endif
                    const uint32_t fs = (Fg.has_ambient ? RX_FS_AMBIENT : 0u) | (Fg.sun_radiance > 0.0f ? RX_FS_SUN : 0u) | ((Fg.has_sky | Fg.has_brush) ? RX_FS_SKY_OR_BRUSH : 0u) |
                                        (Fg.d3_active ? RX_FS_D3 : 0u) | ((Fg.d2_active && S.n_rec2d != 0u) ? RX_FS_D2 : 0u) | (S.n_sectors ? RX_FS_SECTORS : 0u);
                    bad = bad || (fs & RX_SPEC_FRAME_KNOWN) != RX_SPEC_FRAME_VALUE;
                }
                if (bad) atomicOr(&Wk.counters[f].overflow, 32u);
            }
#endif
            cached_frame = f;
            __syncthreads();
            if (!GENERAL && warp < 2u) {
                // what an empty tile must not touch: the union of the cached largeg triangles' pixel boxes (warp 0) and
                // of the (at most 32, in this mode) 2D records' (warp 1)
                int x0 = 0x7FFFFFFF, y0 = 0x7FFFFFFF, x1 = 0, y1 = 0;
                const uint32_t n = warp == 0u ? n_cached : ((Fg.d2_active && S.n_rec2d) ? S.n_rec2d : 0u);
                const Tri2D* recs2 = Wk.tri2d + (size_t)f * Wk.tri2d_stride;
                for (uint32_t i = lane; i < n; i += 32) {
                    const uint32_t bx = warp == 0u ? s_large[i].bbx : recs2[i].bbx, by = warp == 0u ? s_large[i].bby : recs2[i].bby;
                    if ((bx & 0xFFFFu) < (bx >> 16) && (by & 0xFFFFu) < (by >> 16)) {
                        x0 = min(x0, (int)(bx & 0xFFFFu)); x1 = max(x1, (int)(bx >> 16));
                        y0 = min(y0, (int)(by & 0xFFFFu)); y1 = max(y1, (int)(by >> 16));
                    }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    x0 = min(x0, __shfl_xor_sync(0xFFFFFFFFu, x0, o)); y0 = min(y0, __shfl_xor_sync(0xFFFFFFFFu, y0, o));
                    x1 = max(x1, __shfl_xor_sync(0xFFFFFFFFu, x1, o)); y1 = max(y1, __shfl_xor_sync(0xFFFFFFFFu, y1, o));
                }
                if (lane == 0) { int* u = s_union + 4 * warp; u[0] = x0; u[1] = y0; u[2] = x1; u[3] = y1; }
            }
            __syncthreads();
            if (tid == 0) {
                // no tile of this frame can be empty when the largeg triangles' boxes already cover the band (a sky box,
                // the walls of a room): then the per-tile test is skipped altogether
                bool can = Fg.d3_active && !(Fg.has_sky | Fg.has_brush);
                if (!GENERAL) can = can && n_largeg == n_cached && !(s_union[0] <= Fg.band_x0 && s_union[1] <= Fg.band_y0 && s_union[2] >= Fg.band_x1 && s_union[3] >= Fg.band_y1);
                s_can_be_empty = can ? 1 : 0;
            }
            __syncthreads();
        }
        // (the VM variant, far over its register budget already, is faster recomputing them: 3.50 vs 3.64 ms on the batch-shader scene)
        const DFrame& F = VM ? Wk.frames[f] : *s_p.F;
        const DLight* lights_g = VM ? Wk.lights + (size_t)f * Wk.lights_stride : s_p.lights_g;
        const TriVis* vis = VM ? Wk.vis + (size_t)f * Wk.slot_stride : s_p.vis;
        const TriShade* shade = VM ? Wk.shade + (size_t)f * Wk.slot_stride : s_p.shade;
        const DFrameBatch* fbs = VM ? Wk.fb + (size_t)f * Wk.fb_stride : s_p.fbs;
        const uint32_t n_large = s_p.n_large;
        const uint32_t* large = s_p.large;
        const DLight* lights = S.n_lights <= (uint32_t)RX_SMEM_LIGHTS ? s_lights : lights_g;

        const uint32_t smode = SAMPLE == 2 ? F.sample_mode : (uint32_t)SAMPLE;
        const int fw = F.band_x1, fy1 = F.band_y1;   // right / bottom bound of the rendered rectangle
        const int bx0 = F.band_x0, fpitch = F.band_x1 - F.band_x0;   // its left edge and the row pitch of the owner / depth planes
        const int ppitch = (int)out.pitch;   // row pitch of the pixel buffer in pixels: the rectangle's width, or the full frame's when a band is written in place (rxc_mgpu_rasterize)
        const int tx0 = F.band_x0 + s_work[1], ty0 = F.band_y0 + s_work[2];
        const int tx1 = min(tx0 + RX_TILE_W, fw), ty1 = min(ty0 + RX_TILE_H, fy1);
        if (!PLANES && s_can_be_empty) {
            // Empty tile: no binned triangle, no large triangle and no 2D record can touch it -> every pixel is the miss
            // colour vec4_to_pixel((0,0,0,1)) (rasterizer.rs:409-417).  Sparse scenes (an object in front of nothing) are
            // mostly such tiles, and the full tile prologue + resolve costs them ~600 warp-instructions per warp.
            bool empty = s_p.tile_count[tile] == 0u;
            if (GENERAL) {
                empty = empty && !(F.d2_active && S.n_rec2d != 0u && Wk.tile_count2[(size_t)f * Wk.tile_stride + tile] != 0u);
            } else {
                empty = empty &&
                        !(s_union[0] < tx1 && s_union[2] > tx0 && s_union[1] < ty1 && s_union[3] > ty0) &&
                        !(s_union[4] < tx1 && s_union[6] > tx0 && s_union[5] < ty1 && s_union[7] > ty0);
            }
            if (empty) {   // uniform over the CTA
                uint8_t* frame_px = out.pixels + (size_t)f * out.frame_stride;
                if (out.vec_store && tx0 + RX_TILE_W <= fw && ty0 + RX_TILE_H <= fy1) {
                    const int r = (int)tid >> 3, c4 = (int)tid & 7;
                    reinterpret_cast<uint4*>(frame_px + ((size_t)(ty0 - F.band_y0 + r) * (size_t)ppitch + (size_t)(tx0 - bx0)) * 4)[c4] =
                        make_uint4(0xFF000000u, 0xFF000000u, 0xFF000000u, 0xFF000000u);
                } else {
                    for (int i = (int)tid; i < RX_TILE_W * RX_TILE_H; i += RX_TILE_THREADS) {
                        const int r = i >> 5, c = i & 31;
                        if (tx0 + c < fw && ty0 + r < fy1)
                            reinterpret_cast<uint32_t*>(frame_px)[(size_t)(ty0 - F.band_y0 + r) * (size_t)ppitch + (size_t)(tx0 - bx0 + c)] = 0xFF000000u;
                    }
                }
                __syncthreads();  // s_work is rewritten by the next tile
                continue;
            }
        }
        const int rx0 = tx0 + rbx, ry0 = ty0 + rby, rx1 = min(rx0 + RX_REGION_W, fw), ry1 = min(ry0 + RX_REGION_H, fy1);
        const bool region_ok = rx0 < rx1 && ry0 < ry1;
        int px0 = tx0 + lx, py0 = ty0 + ly;
        float fx0 = (float)px0 + 0.5f, fy0 = (float)py0 + 0.5f;  // rasterizer.rs:1022
        uint32_t valid = ((px0 < fw && py0 < fy1) ? 1u : 0u) | ((px0 + RX_DX < fw && py0 < fy1) ? 2u : 0u) |
                         ((px0 < fw && py0 + 4 < fy1) ? 4u : 0u) | ((px0 + RX_DX < fw && py0 + 4 < fy1) ? 8u : 0u);
#if RX_OPAQUE >= 1
        asm volatile("" : "+r"(px0), "+r"(py0));
#endif
#if RX_OPAQUE >= 2
        asm volatile("" : "+f"(fx0), "+f"(fy0), "+r"(valid));
#endif

        Vis4 V;  // z_buffer starts at 1.0 (rasterizer.rs:287)
        Opa4 O;  // z_buffer_opacity starts at 1.0, surface_id at None (:288-290)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            V.z[k] = 1.0f; V.own[k] = RX_OWNER_NONE; V.al[k] = 0.0f; V.be[k] = 0.0f;
            O.z[k] = 1.0f; O.own[k] = RX_OWNER_NONE; O.sid[k] = 0u;
        }
        O.some = 0u;

        if (RX_K_FS(RX_FS_D3, F.d3_active)) {
            // Three sources of triangle records, one walk: (0) the large triangles cached in shared memory (culled
            // per tile by the CTA when there are many), (1) the rest of the large list, (2) the tile's binned list.
            // Warp-private: every lane fetches one record and tests it against the warp's region, the survivors are
            // then read by the whole warp (broadcast loads) -- no staging, no CTA barrier.  Lanes and mask bits are
            // visited in list order, which general mode relies on.
            if (!GENERAL && n_cached > 32u) {
                if (tid < n_cached && rect_overlaps(s_large[tid], tx0, ty0, tx1, ty1) != 0u) s_sel[atomicAdd(&s_nsel, 1u)] = (uint16_t)tid;
                __syncthreads();
            }
            const uint32_t n_list = VM ? Wk.tile_count[(size_t)f * Wk.tile_stride + tile] : s_p.tile_count[tile];
            const uint32_t* list = VM ? Wk.lists + (size_t)f * Wk.list_stride + Wk.tile_base[(size_t)f * Wk.tile_stride + tile] : s_p.lists + s_p.tile_base[tile];
            // long list (uniform over the CTA): a thread per record for the small triangles, the others compacted (see
            // small_triangle_pass).  s_key / s_big alias s_state, which is only written by the resolve below.
            unsigned long long* s_key = reinterpret_cast<unsigned long long*>(s_state);
            uint32_t* s_big = reinterpret_cast<uint32_t*>(s_state + 2 * RX_TILE_THREADS);
            const bool small_pass = SMALL && n_list >= Wk.small_min_list;
            if (small_pass) {
#pragma unroll
                for (int k = 0; k < 4; ++k) s_key[k * RX_TILE_THREADS + tid] = RX_KEY_NONE;
                if (tid == 0) s_nbig = 0u;
                __syncthreads();
                small_triangle_pass(S, F, fbs, vis, shade, list, n_list, tx0, ty0, tx1, ty1, smode, (int)Wk.small_max_pix, Wk.small_gshift, s_key, s_big, &s_nbig);
                __syncthreads();
            }
#if RX_DEPTH_CULL
            uint32_t iz_cut = 0u;   // bits of the largest lower 1/z bound of an opaque record covering the warp's whole region
#endif
#pragma unroll 1
            for (int pass = GENERAL ? 2 : 0; pass < 3; ++pass) {
                // pass 2 after the small-triangle pass: the compacted list; if that overflowed, the whole list with the
                // small records skipped
                const bool from_big = SMALL && pass == 2 && small_pass && s_nbig <= RX_BIG_CAP;
                const bool skip_small = SMALL && pass == 2 && small_pass && !from_big;
                const uint32_t n_src = pass == 0 ? (n_cached > 32u ? s_nsel : n_cached) : pass == 1 ? n_large - n_cached : from_big ? s_nbig : n_list;
                const uint32_t* src = pass == 1 ? large + n_cached : list;
#pragma unroll 1
                for (uint32_t base = 0; base < n_src; base += 32) {
                    const uint32_t i = base + lane;
                    uint32_t slot = 0u, ov = 0u;
#if RX_DEPTH_CULL
                    float iz_lo = 0.0f, iz_hi = 0.0f;
                    bool cull_ok = false, occluder = false;
#endif
                    const TriVis* rp = vis;   // shared (pass 0) or global memory
                    if (i < n_src) {
                        if (!GENERAL && pass == 0) {
                            const uint32_t r = n_cached > 32u ? (uint32_t)s_sel[i] : i;
                            rp = &s_large[r]; slot = s_large_slot[r];
                        } else {
                            slot = from_big ? s_big[i] : __ldg(src + i); rp = vis + slot;
                        }
                        if (region_ok) ov = rect_overlaps(*rp, rx0, ry0, rx1, ry1);
#if RX_DEPTH_CULL
                        if (!GENERAL && ov) {
                            cull_ok = region_iz_bounds(*rp, &iz_lo, &iz_hi);
                            occluder = cull_ok && ov == 2u && !(rp->meta & (RX_META_ALPHA | RX_META_OPACITY));
                        }
#endif
                        if (skip_small) {
                            int a0, b0, a1, b1;
                            if (small_box(rp->bbx, rp->bby, tx0, ty0, tx1, ty1, &a0, &b0, &a1, &b1) <= (int)Wk.small_max_pix) ov = 0u;
                        }
                    }
#if RX_DEPTH_CULL
                    if (!GENERAL) {
                        // the nearest "everything of this region is at least this near" of the opaque full-cover records seen so far in
                        // this tile (positive floats order like their bits); a record whose 1/z stays below it loses the depth test
                        // at every pixel of the region, whatever the order (fast mode: the owner is the minimum over all fragments)
                        if (__any_sync(0xFFFFFFFFu, occluder)) iz_cut = max(iz_cut, __reduce_max_sync(0xFFFFFFFFu, occluder ? __float_as_uint(iz_lo) : 0u));
                        if (cull_ok && iz_hi < __uint_as_float(iz_cut)) ov = 0u;
                    }
#endif
                    uint32_t mask = __ballot_sync(0xFFFFFFFFu, ov != 0u);
                    const uint32_t fullm = __ballot_sync(0xFFFFFFFFu, ov == 2u);
                    while (mask) {
                        const int b = __ffs(mask) - 1;
                        mask &= mask - 1u;
                        const uint32_t rr = __shfl_sync(0xFFFFFFFFu, slot, b);
                        const TriVis* Tp = reinterpret_cast<const TriVis*>(__shfl_sync(0xFFFFFFFFu, reinterpret_cast<unsigned long long>(rp), b));
                        process_record<GENERAL, VM>(S, F, fbs, shade, Tp, rr, (fullm >> b) & 1u, px0, py0, fx0, fy0, valid, smode, Wk.neg_zero, V, O);
                    }
                }
            }
            if (small_pass) {   // merge the small-triangle winners into the walk's state: the same (z, ordinal) rule as test_fragment
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const unsigned long long key = s_key[(ly + ((k >> 1) << 2)) * RX_TILE_W + lx + (k & 1) * RX_DX];
                    const uint32_t sslot = (uint32_t)key;
                    if (sslot != RX_OWNER_NONE) {
                        const float zs = z_from_order_bits((uint32_t)(key >> 32));
                        if ((zs < V.z[k]) || (zs == V.z[k] && V.own[k] != RX_OWNER_NONE && sslot < V.own[k])) {
                            const float4* q = reinterpret_cast<const float4*>(vis + sslot);
                            V.z[k] = fragment_depth(__ldg(q), __ldg(q + 1), __ldg(q + 2), __float_as_uint(__ldg(q + 5).w),
                                                    fx0 + ((k & 1) ? (float)RX_DX : 0.0f), fy0 + ((k & 2) ? 4.0f : 0.0f), &V.al[k], &V.be[k]);
                            V.own[k] = sslot;
                        }
                    }
                }
                __syncthreads();   // every key is read before s_state is written
            }
        }

        uint32_t vm_fault = 0u;
        // resolve, one pixel of the 2x2 at a time (the visibility state goes through shared memory so the
        // shading code is not unrolled and does not hold it in registers): deferred shade of the owner,
        // miss pass (rasterizer.rs:409-461) or the 2D-only background, opacity blend (:464-495)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            s_state[k * RX_TILE_THREADS + tid] = make_float4(V.z[k], __uint_as_float(V.own[k]), V.al[k], V.be[k]);
            if (GENERAL) s_ostate[k * RX_TILE_THREADS + tid] = make_float2(O.z[k], __uint_as_float(O.own[k]));
        }
#pragma unroll 1
        for (int k = 0; k < 4; ++k) {
            const int px = px0 + (k & 1) * RX_DX, py = py0 + ((k >> 1) << 2);
            const float fpx = fx0 + ((k & 1) ? (float)RX_DX : 0.0f), fpy = fy0 + ((k & 2) ? 4.0f : 0.0f);  // exact: small integers + 0.5
            const float4 st = s_state[k * RX_TILE_THREADS + tid];
            const uint32_t owner = __float_as_uint(st.y);
            uint32_t color;
#if RX_PACKED_SHADE && RX_DX == 1
            if (!GENERAL && !(k & 1) && owner != RX_OWNER_NONE && F.d3_active) {
                // the horizontally adjacent pixel has the same owner (it mostly does): both are shaded in packed arithmetic
                const float4 st1 = s_state[(k + 1) * RX_TILE_THREADS + tid];
                if (__float_as_uint(st1.y) == owner) {
                    const uint32_t b = __ldg(&vis[owner].meta) & RX_META_BATCH;
                    const uint32_t fl = fbs[b].sd_flags;
                    if ((fl & (RX_SD_NORMALS | RX_SD_TERRAIN)) == RX_SD_NORMALS && s_k.sun_radiance <= 0.0f && (S.n_sectors == 0u || !s_k.has_ambient)) {
                        const uint2 c2 = shade_owner_pair(S, s_k, lights, s_kd, fbs[b], shade + owner, make_float2(st.z, st1.z), make_float2(st.w, st1.w),
                                                          make_float2(st.x, st1.x), fpx, fpy, smode, Wk.neg_zero);
                        s_color[coff + RX_CPIX(k)] = c2.x;
                        s_color[coff + RX_CPIX(k + 1)] = c2.y;
                        if (PLANES) {
                            if (px < fw && py < fy1) {
                                const size_t o = (size_t)(py - F.band_y0) * (size_t)fpitch + (size_t)(px - bx0);
                                if (out.owner) out.owner[o] = owner;
                                if (out.depth) out.depth[o] = st.x;
                            }
                            if (px + 1 < fw && py < fy1) {
                                const size_t o = (size_t)(py - F.band_y0) * (size_t)fpitch + (size_t)(px + 1 - bx0);
                                if (out.owner) out.owner[o] = owner;
                                if (out.depth) out.depth[o] = st1.x;
                            }
                        }
                        ++k;   // the pair is done
                        continue;
                    }
                }
            }
#endif
            if (RX_K_FS(RX_FS_D3, F.d3_active)) {
                if (owner != RX_OWNER_NONE) {
                    const uint32_t b = __ldg(&vis[owner].meta) & RX_META_BATCH;
                    if (VM && (fbs[b].sd_flags & RX_SD_SHADER))
                        color = shade_owner_vm(S, F, lights_g, fbs[b], shade[owner], st.z, st.w, st.x, fpx, fpy, smode, &vm_fault);
                    else
                        color = shade_owner(S, s_k, lights, s_kd, fbs[b], shade + owner, st.z, st.w, st.x, fpx, fpy, smode);
                } else {
                    color = 0xFF000000u;  // vec4_to_pixel((0,0,0,1))
                    if (RX_K_FS(RX_FS_SKY_OR_BRUSH, F.has_sky | F.has_brush)) color = miss_color(&F, px, py);
                }
                if (GENERAL) {
                    const float2 os = s_ostate[k * RX_TILE_THREADS + tid];
                    if (os.x < 1.0f && st.x > os.x)
                        color = blend_opacity(shade_opacity<VM>(S, F, fbs, vis, shade, __float_as_uint(os.y), fpx, fpy, smode, &vm_fault), color,
                                              F.preserve_transparency != 0u);
                }
            } else {
                color = F.has_bg_color ? F.bg_color : 0u;  // rasterizer.rs:277-282
                if (!F.ignore_bg_shader && F.bg_shader != RXC_BG_NONE) color = shade_background(F, px, py);
            }
            s_color[coff + RX_CPIX(k)] = color;
            if (PLANES && px < fw && py < fy1) {
                const size_t o = (size_t)(py - F.band_y0) * (size_t)fpitch + (size_t)(px - bx0);
                if (out.owner) out.owner[o] = owner;
                if (out.depth) out.depth[o] = st.x;
            }
        }

        // 2D pass in submission order (rasterizer.rs:501-553) over the thread's own pixels in s_color
        // (fast mode: the union of the 2D records' pixel boxes, staged per frame, keeps the tiles they cannot touch out of the loop)
        if (RX_K_FS(RX_FS_D2, F.d2_active && S.n_rec2d != 0u) && region_ok &&
            (GENERAL || (s_union[4] < tx1 && s_union[6] > tx0 && s_union[5] < ty1 && s_union[7] > ty0))) {
            const Tri2D* recs = Wk.tri2d + (size_t)f * Wk.tri2d_stride;
            const DFrameBatch2* fb2 = Wk.fb2 + (size_t)f * Wk.fb2_stride;
            const uint32_t n2 = GENERAL ? Wk.tile_count2[(size_t)f * Wk.tile_stride + tile] : S.n_rec2d;
            const uint32_t* list2 = GENERAL ? Wk.lists2 + (size_t)f * Wk.list2_stride + Wk.tile_base2[(size_t)f * Wk.tile_stride + tile] : nullptr;
#pragma unroll 1
            for (uint32_t base = 0; base < n2; base += 32) {
                const uint32_t i = base + lane;
                uint32_t r = 0u;
                bool hit = false;
                if (i < n2) {
                    r = GENERAL ? __ldg(list2 + i) : i;
                    const uint32_t bbx = recs[r].bbx, bby = recs[r].bby;
                    const int x0 = bbx & 0xFFFF, x1 = bbx >> 16, y0 = bby & 0xFFFF, y1 = bby >> 16;
                    hit = !(x0 >= rx1 || x1 <= rx0 || y0 >= ry1 || y1 <= ry0);
                }
                uint32_t mask = __ballot_sync(0xFFFFFFFFu, hit);
                while (mask) {
                    const int b = __ffs(mask) - 1;
                    mask &= mask - 1u;
                    const Tri2D& T = recs[__shfl_sync(0xFFFFFFFFu, r, b)];
#pragma unroll 1
                    for (int k = 0; k < 4; ++k) {
                        const int px = px0 + (k & 1) * RX_DX, py = py0 + ((k >> 1) << 2);
                        uint32_t* c = &s_color[coff + RX_CPIX(k)];
                        *c = apply_2d<VM>(S, F, lights, fb2, T, px, py, smode, *c, &vm_fault);
                    }
                }
            }
        }
        if (VM && vm_fault) atomicOr(&Wk.counters[f].overflow, 16u);  // a program hit a device limit
#if RX_BULK_STORE || RX_TMA_STORE
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the tile's pixels, written through the generic proxy, for the bulk-copy engine
#endif
        __syncthreads();

        // write back: RGBA8 rows of the tile.  Full tiles with 16 B aligned rows leave through the bulk-copy (TMA) engine,
        // one 128 B row per lane of warp 0 (cp.async.bulk shared -> global); the warp only waits until shared memory has
        // been read, the global writes complete asynchronously.
        uint8_t* frame_px = out.pixels + (size_t)f * out.frame_stride;
        const bool full_tile = (tx0 + RX_TILE_W <= fw) && (ty0 + RX_TILE_H <= fy1);
#if RX_TMA_STORE
        if (out.tma_store) {
            // ONE bulk tensor store per tile, issued by one thread: the TMA engine reads the swizzled 4 KB tile from shared
            // memory and writes the 32 rows; rows / columns beyond the buffer (partial tiles) are clipped by the engine
            if (tid == 0) {
                const uint32_t src = (uint32_t)__cvta_generic_to_shared(&s_color[coff]);
                asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                             ::"l"(&out.tmap), "r"(src), "r"(tx0 - bx0), "r"(ty0 - F.band_y0), "r"((int)f) : "memory");
            }
        } else
#endif
        if (out.vec_store && full_tile) {
#if RX_BULK_STORE
            if (warp == 0) {
                const uint32_t src = (uint32_t)__cvta_generic_to_shared(&s_color[coff + lane * RX_COLOR_STRIDE]);
                uint8_t* dst = frame_px + ((size_t)(ty0 - F.band_y0 + (int)lane) * (size_t)ppitch + (size_t)(tx0 - bx0)) * 4;
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], 128;" ::"l"(dst), "r"(src) : "memory");
            }
#else
            const int r = (int)tid >> 3, c4 = (int)tid & 7;   // 8 x 16 B per 32-pixel row
            const uint4 v = *reinterpret_cast<const uint4*>(&s_color[coff + RX_CIDX(r, c4 * 4)]);
            uint4* dst = reinterpret_cast<uint4*>(frame_px + ((size_t)(ty0 - F.band_y0 + r) * (size_t)ppitch + (size_t)(tx0 - bx0)) * 4) + c4;
            *dst = v;
#endif
        } else {
            for (int i = (int)tid; i < RX_TILE_W * RX_TILE_H; i += RX_TILE_THREADS) {
                const int r = i >> 5, c = i & 31;
                if (tx0 + c < fw && ty0 + r < fy1)
                    reinterpret_cast<uint32_t*>(frame_px)[(size_t)(ty0 - F.band_y0 + r) * (size_t)ppitch + (size_t)(tx0 - bx0 + c)] =
                        s_color[coff + RX_CIDX(r, c)];
            }
        }
#if RX_TMA_STORE
        // the next tile writes the other half: its previous store (two tiles ago) must have read shared memory
        if (tid == 0) {
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        }
        coff ^= (uint32_t)(RX_TILE_H * RX_TILE_W);
#endif
#if RX_BULK_STORE
        // the next tile writes the other half: its previous drain (two tiles ago) must have read shared memory; at most
        // this tile's group stays in flight
        // (every tile commits a group, an empty one when it was stored directly, so "all but the latest" means the other half)
        if (warp == 0) {
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        }
        coff ^= (uint32_t)(RX_TILE_H * RX_COLOR_STRIDE);
#endif
        __syncthreads();  // s_color (its other half), s_work and s_nsel are rewritten by the next tile
    }
#if RX_TMA_STORE
    if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // shared memory stays valid until the engine has read it
#endif
#undef RX_CIDX
#undef RX_CPIX
}

// ---------------------------------------------------------------------------------------------
// k_vm_execute (diagnostics): one thread per record runs a program outside the rasterizer
// ---------------------------------------------------------------------------------------------
#if !defined(__CUDACC_RTC__) || RXVM_JIT_DIAG   // (a JIT build of k_raster leaves it out)
__global__ void __launch_bounds__(128) k_vm_execute(VmDev vm, uint32_t program, uint32_t n, const float* __restrict__ in, float* __restrict__ out,
                                                    uint32_t* faults) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* r = in + (size_t)i * 18;
    VmIO io;
    vm_io_reset(io);
    io.uv = {r[0], r[1], r[2]}; io.color = {r[3], r[4], r[5]}; io.normal = {r[6], r[7], r[8]};
    io.hitpoint = {r[9], r[10], r[11]}; io.time = {r[12], r[13], r[14]}; io.opacity = {r[15], r[16], r[17]};
    if (!vm_run(vm, vm.programs[program], io)) atomicAdd(faults, 1u);
    float* o = out + (size_t)i * 24;
    const f3 v[8] = {io.uv, io.color, io.normal, io.roughness, io.metallic, io.emissive, io.opacity, io.bump};
    for (int k = 0; k < 8; ++k) { o[3 * k] = v[k].x; o[3 * k + 1] = v[k].y; o[3 * k + 2] = v[k].z; }
}
#endif

#if !defined(__CUDACC_RTC__) || RXVM_JIT_ORDERED   // (a JIT build of k_raster leaves it out)
// ---------------------------------------------------------------------------------------------
// k_raster_ordered : the reference's order, for scenes whose batch-shader programs can observe it.
//
// The reference shades FORWARD -- every fragment that passes the depth test is shaded at once, in submission order, triangle by
// triangle, row by row -- and keeps ONE Execution per API tile (tile_size x tile_size pixels) that is never reset
// (src/rasterizer.rs:310): what a program leaves in it (emissive, globals, the .yz of roughness ...) is seen by the fragments shaded
// after it in that tile, with or without a program of their own.  k_raster decides the owner of a pixel first
// and shades only the owner, so it cannot carry that state; this kernel can: one thread (a warp's lane 0) per API tile walks the tile exactly like
// rasterize()'s closure does (:310-553) -- the 3D records of the tile in submission order (the sorted lists of the 32x32 device tiles
// the API tile touches, merged), each over its pixel box row by row, depth test, surface-id skip, the fragment's program and
// lighting with the carried Execution, the write when the result is opaque; the opacity layer likewise; then the miss pass and the
// opacity blend; then the 2D records in submission order -- with the same per-fragment device functions as k_raster.  It is the slow
// path by construction (a thread per tile): rxc_set_vm_state_mode selects it, by default nothing does.
// ---------------------------------------------------------------------------------------------
struct OrderedScratch {   // per frame `stride` entries each (one per pixel)
    float* z; float* zop; uint32_t* cop; uint32_t* sid; uint32_t* some; uint32_t* own;
    size_t stride;
};
#define RX_ORDERED_MAX_LISTS 64   // device tiles one API tile may touch (tile_size <= 224)

// the sorted lists of several device tiles as one ascending sequence without repeats
struct ListMerge {
    const uint32_t* p[RX_ORDERED_MAX_LISTS];
    uint32_t n[RX_ORDERED_MAX_LISTS], i[RX_ORDERED_MAX_LISTS];
    int k;
    __device__ bool next(uint32_t* out) {
        uint32_t best = 0xFFFFFFFFu;
        for (int c = 0; c < k; ++c) if (i[c] < n[c]) best = min(best, p[c][i[c]]);
        if (best == 0xFFFFFFFFu) return false;   // (the padding of a sorted list is 0xFFFFFFFF too)
        for (int c = 0; c < k; ++c) while (i[c] < n[c] && p[c][i[c]] == best) ++i[c];
        *out = best;
        return true;
    }
};

__global__ void __launch_bounds__(128) k_raster_ordered(SceneDev S, Workspace Wk, RasterOut out, uint32_t n_frames, OrderedScratch Q) {
    // One API tile per WARP.  The tile's Execution lives in shared memory; everything that does not touch it is done by the 32 lanes
    // side by side -- the coverage and depth tests of a record's row (a record visits a pixel once, so the depth tests of one row do not
    // depend on each other), the exact derivation of the fragments' program inputs, and the lighting of what the programs returned --
    // and what does (the assignments to the Execution and the program run, in pixel order) is done one fragment at a time by the lane
    // that owns the pixel.  (A thread per tile, 32 tiles sharing a warp: 79 ms per 1080p frame of the batch-shader scene; a warp per
    // tile with one working lane: 36 ms.)
    __shared__ VmIO s_io[4];
    __shared__ VmPersist s_ps[4];
    pdl_enter();
    const uint32_t f = blockIdx.y;
    if (f >= n_frames) return;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const DFrame& F = Wk.frames[f];
    const int W = F.width, H = F.height, ts = max(1, (int)F.tile_size);
    const int atx = (W + ts - 1) / ts, aty = (H + ts - 1) / ts;
    const int at = (int)(blockIdx.x * (blockDim.x >> 5) + warp);
    if (at >= atx * aty) return;   // (the whole warp)
    const int ax0 = (at % atx) * ts, ay0 = (at / atx) * ts, ax1 = min(ax0 + ts, W), ay1 = min(ay0 + ts, H);
    const DLight* lights = Wk.lights + (size_t)f * Wk.lights_stride;
    const TriVis* vis = Wk.vis + (size_t)f * Wk.slot_stride;
    const TriShade* shade = Wk.shade + (size_t)f * Wk.slot_stride;
    const DFrameBatch* fbs = Wk.fb + (size_t)f * Wk.fb_stride;
    const uint32_t smode = F.sample_mode;
    float* zb = Q.z + (size_t)f * Q.stride; float* zop = Q.zop + (size_t)f * Q.stride;
    uint32_t* cop = Q.cop + (size_t)f * Q.stride; uint32_t* sid = Q.sid + (size_t)f * Q.stride;
    uint32_t* some = Q.some + (size_t)f * Q.stride; uint32_t* own = Q.own + (size_t)f * Q.stride;
    uint32_t* px_out = reinterpret_cast<uint32_t*>(out.pixels + (size_t)f * out.frame_stride);
    const int pitch = (int)out.pitch;
    uint32_t fault = 0u;
    const uint32_t FULL = 0xFFFFFFFFu;

    VmIO& io = s_io[warp];             // `let mut execution = Execution::new(0)` of this tile (:310)
    VmPersist& ps = s_ps[warp];
    if (lane == 0u) { vm_io_reset(io); ps.n_globals = 0u; ps.n_locals = 0u; }
    const int tw = ax1 - ax0, th = ay1 - ay0;
    for (int p = (int)lane; p < tw * th; p += 32) {
        const size_t i = (size_t)(ay0 + p / tw) * W + (ax0 + p % tw);
        zb[i] = 1.0f; zop[i] = 1.0f; some[i] = 0u; sid[i] = 0u; cop[i] = 0u; own[i] = RX_OWNER_NONE;   // :287-290
    }
    __syncwarp();

    // the device tiles under this API tile
    const int dx0 = ax0 / RX_TILE_W, dx1 = (ax1 - 1) / RX_TILE_W, dy0 = ay0 / RX_TILE_H, dy1 = (ay1 - 1) / RX_TILE_H;
    ListMerge M;   // (every lane walks the same lists: uniform)
    if (F.d3_active) {
        M.k = 0;
        for (int dy = dy0; dy <= dy1; ++dy)
            for (int dx = dx0; dx <= dx1; ++dx) {
                const size_t t = (size_t)f * Wk.tile_stride + (size_t)dy * F.tiles_x + dx;
                M.p[M.k] = Wk.lists + (size_t)f * Wk.list_stride + Wk.tile_base[t]; M.n[M.k] = Wk.tile_count[t]; M.i[M.k] = 0u;
                ++M.k;
            }
        uint32_t slot;
        while (M.next(&slot)) {
            const float4* q = reinterpret_cast<const float4*>(vis + slot);
            const float4 q0 = q[0], q1 = q[1], q2 = q[2], q3 = q[3], q4 = q[4], q5 = q[5];
            const uint32_t bbx = __float_as_uint(q5.y), bby = __float_as_uint(q5.z), meta = __float_as_uint(q5.w);
            const int x0 = max(ax0, (int)(bbx & 0xFFFFu)), x1 = min(ax1, (int)(bbx >> 16)), y0 = max(ay0, (int)(bby & 0xFFFFu)), y1 = min(ay1, (int)(bby >> 16));
            if (x0 >= x1 || y0 >= y1) continue;
            const DFrameBatch& FB = fbs[meta & RX_META_BATCH];
            const TriShade& sh = shade[slot];
            const float ea[3] = {q3.x, q3.y, q3.z}, eb[3] = {q3.w, q4.x, q4.y}, ec[3] = {q4.z, q4.w, q5.x};
            const bool opacity = (meta & RX_META_OPACITY) != 0u, has_profile = (FB.sd_flags & RX_SD_HAS_PROFILE) != 0u;
            const bool program = (FB.sd_flags & RX_SD_SHADER) && FB.sd_program >= 0 && (uint32_t)FB.sd_program < S.vm.n_programs &&
                                 S.vm.programs[FB.sd_program].n_words != 0u;
            for (int y = y0; y < y1; ++y)            // rasterizer.rs:1018-1020: rows, then columns
                for (int xs = x0; xs < x1; xs += 32) {
                    const int x = xs + (int)lane;
                    const float fpx = (float)x + 0.5f, fpy = (float)y + 0.5f;   // :1022
                    const size_t i = (size_t)y * W + min(x, x1 - 1);
                    bool pass = x < x1 && !((ea[0] * fpx + eb[0] * fpy) + ec[0] < 0.0f || (ea[1] * fpx + eb[1] * fpy) + ec[1] < 0.0f ||
                                            (ea[2] * fpx + eb[2] * fpy) + ec[2] < 0.0f);   // edge.rs:28-36
                    float al = 0.0f, be = 0.0f, z = 1.0f;
                    if (pass) {
                        z = fragment_depth(q0, q1, q2, meta, fpx, fpy, &al, &be);
                        if (opacity) pass = z < zop[i];                                                   // d3_rasterize_opacity, :1425-1690
                        else pass = !(has_profile && some[i] && sid[i] == FB.sd_profile) && z < zb[i];     // :1041-1047, :1051-1060
                    }
                    Frag3D g = {};
                    if (pass) g = vm_inputs_3d(S, F, FB, sh, al, be, z, fpx, fpy, smode, opacity);
                    const uint32_t mask = __ballot_sync(FULL, pass);
                    if (!mask) continue;
                    VmIO snap;   // the Execution as the rasterizer reads it back for THIS fragment
                    if (program) {
                        // in pixel order, one fragment at a time: the assignments of :1259-1298 (:1633-1661 in the opacity pass) and the program
                        for (uint32_t m = mask; m; m &= m - 1u) {
                            if (lane == (uint32_t)(__ffs(m) - 1)) {
                                io.color = texel_linear_exact(g.texel);
                                io.opacity.x = (float)(g.texel >> 24) / 255.0f;
                                io.normal = opacity ? f3{0.0f, 0.0f, 0.0f} : g.normal;
                                io.roughness.x = 0.5f; io.metallic.x = 0.0f;
                                io.uv.x = g.u / 4.0f; io.uv.y = g.v / 4.0f;
                                io.hitpoint = g.world;
                                io.time = {F.time, F.time, F.time};
                                if (!vm_run_t<true>(S.vm, S.vm.programs[FB.sd_program], io, &ps)) fault = 1u;
                                snap = io;
                            }
                            __syncwarp();
                        }
                    } else {
                        // no program: a fragment assigns color, opacity.x (and, opaque pass, normal, roughness.x, metallic.x; :1310-1316) and reads
                        // nothing another fragment of this row wrote -- the row's last fragment is what stays in the Execution
                        snap = io;
                        snap.color = texel_linear_exact(g.texel);
                        snap.opacity.x = (float)(g.texel >> 24) / 255.0f;
                        if (!opacity) { snap.normal = g.normal; snap.roughness.x = 0.5f; snap.metallic.x = 0.0f; }
                        __syncwarp();
                        if (lane == (uint32_t)(31 - __clz(mask))) {
                            io.color = snap.color; io.opacity.x = snap.opacity.x;
                            if (!opacity) { io.normal = snap.normal; io.roughness.x = 0.5f; io.metallic.x = 0.0f; }
                        }
                        __syncwarp();
                    }
                    if (pass) {
                        if (opacity) {   // :1670-1685: the layer's pixel, its depth, the surface id
                            cop[i] = pack_pixel(linear_to_srgb_exact(snap.color.x), linear_to_srgb_exact(snap.color.y), linear_to_srgb_exact(snap.color.z), snap.opacity.x);
                            zop[i] = z; sid[i] = FB.sd_profile; some[i] = has_profile ? 1u : 0u;
                        } else {
                            const uint32_t color = vm_light_3d(S, F, lights, FB, snap, g.world);   // :1319-1404, emissive of the carried Execution included
                            if ((color >> 24) == 255u) { px_out[(size_t)y * pitch + x] = color; zb[i] = z; own[i] = slot; }   // :1408-1412
                        }
                    }
                    __syncwarp();   // the next record's lanes read what these lanes wrote
                }
        }
    }

    // miss pass (:409-461), opacity blend (:464-495), or the 2D-only background (:277-307)
    __syncwarp();
    for (int p = (int)lane; p < tw * th; p += 32) {
        const int x = ax0 + p % tw, y = ay0 + p / tw;
        const size_t i = (size_t)y * W + x;
        uint32_t color;
        if (F.d3_active) {
            if (own[i] != RX_OWNER_NONE) color = px_out[(size_t)y * pitch + x];
            else { color = 0xFF000000u; if (F.has_sky | F.has_brush) color = miss_color(&F, x, y); }
            if (zop[i] < 1.0f && zb[i] > zop[i]) color = blend_opacity(cop[i], color, F.preserve_transparency != 0u);
        } else {
            color = F.has_bg_color ? F.bg_color : 0u;
            if (!F.ignore_bg_shader && F.bg_shader != RXC_BG_NONE) color = shade_background(F, x, y);
        }
        px_out[(size_t)y * pitch + x] = color;
        if (out.owner) out.owner[i] = own[i];
        if (out.depth) out.depth[i] = zb[i];
    }
    __syncwarp();

    // 2D batches in submission order (:501-553, :584-959) with the same Execution: records without a program do not touch it (:760)
    if (F.d2_active && S.n_rec2d != 0u) {
        const Tri2D* recs = Wk.tri2d + (size_t)f * Wk.tri2d_stride;
        const DFrameBatch2* fb2 = Wk.fb2 + (size_t)f * Wk.fb2_stride;
        M.k = 0;
        for (int dy = dy0; dy <= dy1; ++dy)
            for (int dx = dx0; dx <= dx1; ++dx) {
                const size_t t = (size_t)f * Wk.tile_stride + (size_t)dy * F.tiles_x + dx;
                M.p[M.k] = Wk.lists2 + (size_t)f * Wk.list2_stride + Wk.tile_base2[t]; M.n[M.k] = Wk.tile_count2[t]; M.i[M.k] = 0u;
                ++M.k;
            }
        uint32_t r;
        while (M.next(&r)) {
            const Tri2D& T = recs[r];
            const int x0 = max(ax0, (int)(T.bbx & 0xFFFFu)), x1 = min(ax1, (int)(T.bbx >> 16)), y0 = max(ay0, (int)(T.bby & 0xFFFFu)), y1 = min(ay1, (int)(T.bby >> 16));
            const bool program = T.kind == 0u && fb2[T.batch].program >= 0;
            for (int y = y0; y < y1; ++y)
                for (int xs = x0; xs < x1; xs += 32) {
                    const int x = xs + (int)lane;
                    uint32_t* c = px_out + (size_t)y * pitch + min(x, x1 - 1);
                    if (!program) {
                        if (x < x1) *c = apply_2d<true>(S, F, lights, fb2, T, x, y, smode, *c, &fault);
                    } else {
                        for (uint32_t m = __ballot_sync(FULL, x < x1); m; m &= m - 1u) {   // pixel order through the Execution
                            if (lane == (uint32_t)(__ffs(m) - 1)) *c = apply_2d<true>(S, F, lights, fb2, T, x, y, smode, *c, &fault, &io, &ps);
                            __syncwarp();
                        }
                    }
                    __syncwarp();
                }
        }
    }
    if (__any_sync(FULL, fault != 0u) && lane == 0u) atomicOr(&Wk.counters[f].overflow, 16u);
}

#endif

#ifdef __CUDACC_RTC__
}  // namespace
#else
// ---------------------------------------------------------------------------------------------
// k_list_sort (general mode): one CTA per tile sorts its list ascending (= submission order).  The
// allocation of a list is a power of two (k_tile_alloc), the tail is padded with 0xFFFFFFFF.
// ---------------------------------------------------------------------------------------------
#define RX_SORT_SMEM 4096
__global__ void __launch_bounds__(256) k_list_sort(Workspace Wk, int which) {
    __shared__ uint32_t sm[RX_SORT_SMEM];
    pdl_enter();
    const uint32_t f = blockIdx.y, t = blockIdx.x, tid = threadIdx.x;
    const uint32_t n = (which ? Wk.tile_count2 : Wk.tile_count)[(size_t)f * Wk.tile_stride + t];
    if (n < 2u) return;
    uint32_t* list = (which ? Wk.lists2 + (size_t)f * Wk.list2_stride : Wk.lists + (size_t)f * Wk.list_stride) +
                     (which ? Wk.tile_base2 : Wk.tile_base)[(size_t)f * Wk.tile_stride + t];
    const uint32_t P = 1u << (32 - __clz(n - 1u));
    uint32_t* d = P <= RX_SORT_SMEM ? sm : list;
    for (uint32_t i = tid; i < P; i += blockDim.x) {
        if (P <= RX_SORT_SMEM) sm[i] = i < n ? list[i] : 0xFFFFFFFFu;
        else if (i >= n) list[i] = 0xFFFFFFFFu;
    }
    __syncthreads();
    for (uint32_t k = 2; k <= P; k <<= 1) {
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            for (uint32_t i = tid; i < P; i += blockDim.x) {
                const uint32_t ixj = i ^ j;
                if (ixj > i) {
                    const uint32_t a = d[i], b = d[ixj];
                    if ((a > b) == ((i & k) == 0u)) { d[i] = b; d[ixj] = a; }
                }
            }
            __syncthreads();
        }
    }
    if (P <= RX_SORT_SMEM)
        for (uint32_t i = tid; i < n; i += blockDim.x) list[i] = sm[i];
}

// ---------------------------------------------------------------------------------------------
// k_selftest_div : rx_div_by against div.rn over the operand ranges make_tri admits
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t splitmix64(uint64_t& x) {
    uint64_t z = (x += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__device__ __forceinline__ float make_float(uint32_t sign, int exp2, uint32_t mant23) {
    return __uint_as_float((sign << 31) | ((uint32_t)(exp2 + 127) << 23) | (mant23 & 0x7FFFFFu));
}
__global__ void __launch_bounds__(256) k_selftest_div(uint64_t seed, uint32_t iters, unsigned long long* mismatches) {
    uint64_t st = seed ^ ((uint64_t)(blockIdx.x * blockDim.x + threadIdx.x) * 0xD1342543DE82EF95ull);
    uint32_t bad = 0;
    for (uint32_t it = 0; it < iters; ++it) {
        const uint64_t r0 = splitmix64(st), r1 = splitmix64(st);
        // divisor: |b| in [2^-20, 2^44), mantissa random or one of the hard patterns
        uint32_t mb = (uint32_t)r0 & 0x7FFFFFu;
        const uint32_t pat = (uint32_t)(r0 >> 23) & 7u;
        if (pat == 0) mb = 0x7FFFFFu; else if (pat == 1) mb = 0u; else if (pat == 2) mb = 0x7FFFFEu; else if (pat == 3) mb = 1u;
        const float b = make_float((uint32_t)(r0 >> 26) & 1u, -20 + (int)((r0 >> 27) % 64u), mb);
        const float rb = 1.0f / b;
        float a;
        const uint32_t kind = (uint32_t)(r1 >> 60);
        if (kind == 0) {
            a = 0.0f;
        } else if (kind < 6) {   // independent numerator, |a| in [2^-80, 2^62)
            a = make_float((uint32_t)(r1 >> 23) & 1u, -80 + (int)((r1 >> 24) % 142u), (uint32_t)r1);
        } else {                 // quotient near a representable value or near a rounding midpoint
            const float q = make_float((uint32_t)(r1 >> 23) & 1u, -30 + (int)((r1 >> 24) % 48u), (uint32_t)r1);
            const float qh = __uint_as_float(__float_as_uint(q) + 1u);
            a = (kind < 11) ? q * b : (float)(((double)q + (double)qh) * 0.5 * (double)b);
            a = __uint_as_float(__float_as_uint(a) + (((uint32_t)(r1 >> 40) % 5u)) - 2u);
            if (!(fabsf(a) >= 8.271806125530277e-25f && fabsf(a) < 4.611686018427388e18f)) a = 0.0f;
        }
        const float want = __fdiv_rn(a, b), got = rx_div_by(a, b, rb);
        if (__float_as_uint(want) != __float_as_uint(got) && !(want == 0.0f && got == 0.0f)) {
            ++bad;
            mismatches[1] = ((unsigned long long)__float_as_uint(a) << 32) | __float_as_uint(b);  // any failing pair
        }
    }
    if (bad) atomicAdd(mismatches, (unsigned long long)bad);
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// launch wrappers
// ---------------------------------------------------------------------------------------------
// Kernels of a frame's chain are launched with programmatic stream serialization (see pdl_enter): the launch may begin once
// every CTA of the kernel before it in the stream has started; the kernel's own griddepcontrol.wait orders the data.
// RXC_PDL=0 launches them the ordinary way (A/B measurements).
static bool pdl_enabled() {
    static const bool on = !(getenv("RXC_PDL") && atoi(getenv("RXC_PDL")) == 0);
    return on;
}
static void pdl_config(cudaLaunchConfig_t* cfg, cudaLaunchAttribute* at, dim3 grid, dim3 block, cudaStream_t st, unsigned cluster_x = 0) {
    *cfg = cudaLaunchConfig_t{};
    cfg->gridDim = grid; cfg->blockDim = block; cfg->stream = st;
    unsigned n = 0;
    if (cluster_x) { at[n].id = cudaLaunchAttributeClusterDimension; at[n].val.clusterDim.x = cluster_x; at[n].val.clusterDim.y = 1; at[n].val.clusterDim.z = 1; ++n; }
    // (not on the legacy default stream, whose implicit synchronisation with other streams is not what the attribute relaxes)
    if (pdl_enabled() && st != nullptr && st != cudaStreamLegacy) { at[n].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[n].val.programmaticStreamSerializationAllowed = 1; ++n; }
    cfg->attrs = at; cfg->numAttrs = n;
}
template <typename... KArgs, typename... Args>
static cudaError_t pdl_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg; cudaLaunchAttribute at[2];
    pdl_config(&cfg, at, grid, block, st);
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
static cudaError_t pdl_launch_ptr(const void* kernel, dim3 grid, dim3 block, cudaStream_t st, void** args) {
    cudaLaunchConfig_t cfg; cudaLaunchAttribute at[2];
    pdl_config(&cfg, at, grid, block, st);
    return cudaLaunchKernelExC(&cfg, kernel, args);
}
cudaError_t rxk_frame_setup(const SceneDev& S, const Workspace& W, uint32_t n_frames, uint32_t tiles_per_frame, cudaStream_t st) {
    const uint32_t zero_blocks = max(max(1u, min(64u, (tiles_per_frame + 255u) / 256u)), min(64u, (S.n_b3 + 63u) / 64u));
    dim3 grid(1 + S.n_b2 + zero_blocks, n_frames);
    return pdl_launch(k_frame_setup, dim3(grid), dim3(256), st, S, W, tiles_per_frame);
}
cudaError_t rxk_front_small(const SceneDev& S, const Workspace& W, uint32_t n_frames, uint32_t tiles_per_frame, cudaStream_t st) {
    return pdl_launch(k_front_small, dim3(n_frames), dim3(256), st, S, W, tiles_per_frame, 2u + S.n_b2);   // block 0, the 2D batches, one state/zeroing block
}
cudaError_t rxk_front_cluster(const SceneDev& S, const Workspace& W, uint32_t n_frames, uint32_t tiles_per_frame, uint32_t stop_phase, cudaStream_t st) {
    const uint32_t zero_blocks = max(1u, min(64u, (tiles_per_frame + 255u) / 256u));
    cudaLaunchConfig_t cfg; cudaLaunchAttribute at[2];
    pdl_config(&cfg, at, dim3(RX_FRONT_CLUSTER, n_frames), dim3(256), st, RX_FRONT_CLUSTER);
    return cudaLaunchKernelEx(&cfg, k_front_cluster, S, W, tiles_per_frame, 1u + S.n_b2 + zero_blocks, stop_phase);
}
cudaError_t rxk_tri_setup(const SceneDev& S, const Workspace& W, uint32_t n_frames, cudaStream_t st) {
    if (S.n_chunks == 0) return cudaSuccess;
    dim3 grid(S.n_chunks, n_frames);
    return pdl_launch(k_tri_setup, dim3(grid), dim3(RX_CHUNK_TRIS), st, S, W);
}
cudaError_t rxk_tri_setup_projected(const SceneDev& S, const Workspace& W, const ProjectedDev& P, cudaStream_t st) {
    const uint32_t n = max(max(S.n_tris, P.n_clipped), S.n_b3);
    if (n == 0) return cudaSuccess;
    return pdl_launch(k_tri_setup_projected, dim3((n + 255) / 256), dim3(256), st, S, W, P);
}

cudaError_t rxk_batch_finalize(const SceneDev& S, const Workspace& W, uint32_t n_frames, cudaStream_t st) {
    if (S.n_b3 == 0) return cudaSuccess;
    dim3 grid((S.n_b3 + 7) / 8, n_frames);
    return pdl_launch(k_batch_finalize, dim3(grid), dim3(256), st, S, W);
}
cudaError_t rxk_clip_emit(const SceneDev& S, const Workspace& W, uint32_t n_frames, int grid_x, cudaStream_t st) {
    if (S.n_tris == 0) return cudaSuccess;
    dim3 grid(grid_x, n_frames);
    return pdl_launch(k_clip_emit, dim3(grid), dim3(128), st, S, W);
}
cudaError_t rxk_bin_count(const SceneDev& S, const Workspace& W, uint32_t n_frames, int grid_x, cudaStream_t st) {
    if (S.n_tris == 0) return cudaSuccess;
    dim3 grid(grid_x, n_frames);
    return pdl_launch(k_bin_count, dim3(grid), dim3(256), st, S, W);
}
cudaError_t rxk_tile_alloc(const SceneDev& S, const Workspace& W, uint32_t n_frames, uint32_t tiles_per_frame, int which, int pow2,
                           cudaStream_t st) {
    (void)S;
    dim3 grid((tiles_per_frame + 255) / 256, n_frames);
    return pdl_launch(k_tile_alloc, dim3(grid), dim3(256), st, W, tiles_per_frame, which, pow2);
}
cudaError_t rxk_bin2d(const SceneDev& S, const Workspace& W, uint32_t n_frames, int fill, cudaStream_t st) {
    if (S.n_rec2d == 0) return cudaSuccess;
    dim3 grid((S.n_rec2d + 7) / 8, n_frames);
    return pdl_launch(k_bin2d, dim3(grid), dim3(256), st, S, W, fill);
}
cudaError_t rxk_bin_large(const SceneDev& S, const Workspace& W, uint32_t n_frames, int fill, int grid_x, cudaStream_t st) {
    if (S.n_tris == 0) return cudaSuccess;
    dim3 grid(grid_x, n_frames);
    return pdl_launch(k_bin_large, dim3(grid), dim3(256), st, S, W, fill);
}
cudaError_t rxk_list_sort(const SceneDev& S, const Workspace& W, uint32_t n_frames, uint32_t tiles_per_frame, int which, cudaStream_t st) {
    (void)S;
    dim3 grid(tiles_per_frame, n_frames);
    return pdl_launch(k_list_sort, dim3(grid), dim3(256), st, W, which);
}
cudaError_t rxk_bin_fill(const SceneDev& S, const Workspace& W, uint32_t n_frames, int grid_x, cudaStream_t st) {
    if (S.n_tris == 0) return cudaSuccess;
    dim3 grid(grid_x, n_frames);
    return pdl_launch(k_bin_fill, dim3(grid), dim3(256), st, S, W);
}
cudaError_t rxk_raster_ordered(const SceneDev& S, const Workspace& W, const RasterOut& out, uint32_t n_frames, uint32_t api_tiles, float* z, float* zop,
                               uint32_t* cop, uint32_t* sid, uint32_t* some, uint32_t* own, size_t stride, cudaStream_t st, void* jit_kernel) {
    OrderedScratch q = {z, zop, cop, sid, some, own, stride};
    dim3 grid((api_tiles + 3u) / 4u, n_frames);
    if (jit_kernel) {   // the same kernel with the scene's programs compiled (rx_jit.cu)
        SceneDev s = S; Workspace w = W; RasterOut o = out;
        void* args[] = {&s, &w, &o, &n_frames, &q};
        return pdl_launch_ptr((const void*)jit_kernel, grid, dim3(128), st, args);
    }
    return pdl_launch(k_raster_ordered, grid, dim3(128), st, S, W, out, n_frames, q);
}
// experiments: RXC_SMEM_CARVEOUT = preferred shared-memory carve-out of the raster kernels in percent (what is left is L1)
static void raster_carveout(const void* kernel) {
    static const int carve = getenv("RXC_SMEM_CARVEOUT") ? atoi(getenv("RXC_SMEM_CARVEOUT")) : -1;
    if (carve < 0) return;
    static std::set<const void*> done;
    if (done.insert(kernel).second) { cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, carve); cudaGetLastError(); }
}
int rxk_raster_mode(const SceneDev& S, const Workspace& W) {
    return (S.general && S.vm.n_programs) ? 2 : S.general ? 1 : S.n_tris >= W.small_min_tris ? 3 : 0;
}
cudaError_t rxk_raster(const SceneDev& S, const Workspace& W, const RasterOut& out, uint32_t n_frames, uint32_t tile0, uint32_t n_tiles,
                       uint32_t counter, int sample_mode, int grid_x, cudaStream_t st, void* jit_kernel) {
    const bool planes = out.owner || out.depth;
    if (jit_kernel) {
        raster_carveout(jit_kernel);
        // the same kernel recompiled for this scene (rx_jit.cu: its programs as straight-line code, its constants folded): same arguments
        SceneDev s = S; Workspace w = W; RasterOut o = out;
        void* args[] = {&s, &w, &o, &n_frames, &tile0, &n_tiles, &counter};
        return pdl_launch_ptr((const void*)jit_kernel, dim3(grid_x), dim3(RX_TILE_THREADS), st, args);
    }
#define RX_LAUNCH(SM, PL, MD) do { raster_carveout((const void*)k_raster<SM, PL, MD>); return pdl_launch(k_raster<SM, PL, MD>, dim3(grid_x), dim3(RX_TILE_THREADS), st, S, W, out, n_frames, tile0, n_tiles, counter); } while (0)
#define RX_LAUNCH2(SM, PL) do { if (S.general && S.vm.n_programs) RX_LAUNCH(SM, PL, 2); else if (S.general) RX_LAUNCH(SM, PL, 1); else if (S.n_tris >= W.small_min_tris) RX_LAUNCH(SM, PL, 3); else RX_LAUNCH(SM, PL, 0); } while (0)
    if (sample_mode == 0) { if (planes) RX_LAUNCH2(0, true); else RX_LAUNCH2(0, false); }
    else if (sample_mode == 1) { if (planes) RX_LAUNCH2(1, true); else RX_LAUNCH2(1, false); }
    else { if (planes) RX_LAUNCH2(2, true); else RX_LAUNCH2(2, false); }
#undef RX_LAUNCH2
#undef RX_LAUNCH
    return cudaGetLastError();
}
int rxk_raster_blocks_per_sm() {
    int n = 0, best = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_raster<0, false, 0>, RX_TILE_THREADS, 0) == cudaSuccess) best = n;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_raster<1, false, 0>, RX_TILE_THREADS, 0) == cudaSuccess && n < best) best = n;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_raster<1, false, 3>, RX_TILE_THREADS, 0) == cudaSuccess && n < best) best = n;
    return best < 1 ? 1 : best;
}
cudaError_t rxk_vm_execute(const SceneDev& S, uint32_t program, uint32_t n, const float* d_in, float* d_out, uint32_t* d_faults, cudaStream_t st, void* jit_kernel) {
    if (jit_kernel) {
        VmDev vm = S.vm;
        void* args[] = {&vm, &program, &n, &d_in, &d_out, &d_faults};
        return cudaLaunchKernel((const void*)jit_kernel, dim3((n + 127) / 128), dim3(128), args, 0, st);
    }
    k_vm_execute<<<(n + 127) / 128, 128, 0, st>>>(S.vm, program, n, d_in, d_out, d_faults);
    return cudaGetLastError();
}
cudaError_t rxk_selftest_div(uint64_t seed, uint32_t blocks, uint32_t iters, unsigned long long* d_mismatches, cudaStream_t st) {
    k_selftest_div<<<blocks, 256, 0, st>>>(seed, iters, d_mismatches);
    return cudaGetLastError();
}
#endif  // !__CUDACC_RTC__
