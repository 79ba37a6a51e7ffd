// rx_api.cu -- the C ABI of include/rxcuda.h: context, scene/texture residency, per-frame launch
// sequence.  Host logic only; the kernels are in rx_kernels.cu.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <exception>
#include <new>
#include <string>
#include <vector>

#include "rx_internal.h"
#include "rx_jit.h"
#include "rx_kernels.cuh"

namespace {

const char* const kKernelNames[RXC_N_KERNELS] = {"k_frame_setup", "k_tri_setup", "k_batch_finalize", "k_clip_emit", "k_bin_count",
                                                 "k_tile_alloc",  "k_bin_fill",  "k_raster",         "k_bin2d",     "k_list_sort",
                                                 "k_front_fused"};  // k_front_small or k_front_cluster

struct HChunk {  // host copy of what rxc_chunk carries besides its batches
    int32_t origin[2]; int32_t size;
    uint32_t sector_off, n_sectors;  // into rxc_ctx::h_sectors
    int32_t terrain_tex;             // index into the scene-texture list (h_dyn_tex) or -1
};

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct PendingEvent { int cls; cudaEvent_t a, b; };

}  // namespace

struct rxc_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    std::string err;

    // textures: static (assets.tile_list) then dynamic (scene.dynamic_textures)
    std::vector<uint8_t> h_static_arena;
    std::vector<DTex> h_static_tex;
    std::vector<DTile> h_static_tiles;
    std::vector<uint8_t> h_dyn_arena;
    std::vector<DTex> h_dyn_tex;
    std::vector<DTile> h_dyn_tiles;
    DevBuf d_arena, d_tex, d_tiles;
    bool textures_dirty = true;
    uint32_t n_dyn_tiles = 0, n_actor_tiles = 0;   // h_dyn_tiles = dynamic tiles, then entity/item tiles
    // chunks and the mapmini (occluded sectors, terrain textures, linedefs)
    std::vector<HChunk> h_chunks;
    std::vector<DSector> h_sectors;        // sectors of the scene's chunks
    std::vector<DSector> h_mm_sectors;     // mapmini.occluded_sectors
    std::vector<float> h_linedefs;         // 4 floats per mapmini linedef
    DevBuf d_sectors, d_chunkinfo, d_linedefs;
    DevBuf d_vm_code, d_vm_programs, d_vm_patdata, d_vm_patterns, d_vm_palette;   // Rusteria VM residency

    // scene
    std::vector<DBatch3> h_b3;
    std::vector<DBatch2> h_b2;
    // the flattened 3D geometry of the current scene, kept on the host so that rxc_update_scene can reuse the part of it that did not change
    std::vector<float> h_pos, h_uv, h_nrm;
    std::vector<uint32_t> h_idx, h_orphans;
    std::vector<DChunk> h_setup_chunks;
    bool h_geometry_valid = false;
    std::vector<uint32_t> owner_base;
    DevBuf d_pos, d_uv, d_nrm, d_idx, d_b3, d_chunks, d_orphans, d_pos2, d_uv2, d_idx2, d_b2, d_lights;
    SceneDev S = {};
    bool have_scene = false;
    bool lists_sized = false;   // a frame of this scene has been checked for tile-list overflow (the first call always is)

    // per-frame workspace
    Workspace W = {};
    DevBuf w_frames, w_fb, w_fb2, w_lights, w_counters, w_vis, w_shade, w_bins, w_ctot, w_cbase, w_clip, w_large, w_tcount,
        w_tbase, w_tfill, w_lists, w_tri2d, w_rcounter, w_tcount2, w_tbase2, w_tfill2, w_lists2;
    uint32_t ws_frames = 0, ws_tiles = 0;   // what the workspace is currently sized for
    uint32_t list_cap_per_frame = 0, list_cap_min = 0;
    uint32_t list2_cap_per_frame = 0, list2_cap_min = 0;
    DevBuf d_out_px, d_out_owner, d_out_depth;  // staging for host outputs
    DFrame* h_frames = nullptr;  size_t h_frames_cap = 0;      // pinned
    DCounters* h_counters = nullptr; size_t h_counters_cap = 0; // pinned
    int raster_blocks_per_sm = 1;
    int piece_mb = 8;             // host output: small frames are rendered and drained in groups of about this size
    int slice_mb = 4;             // host output: large frames are rendered and drained in slices of about this size (0 = whole frames)
    int tma_store = 1;            // RX_TMA_STORE builds: tiles leave through the tensor-map store (RXC_TMA_STORE=0 switches back to STG.128)
    int front_stop = 0;           // profiling aid: k_front_cluster leaves after this many phases (RXC_FRONT_STOP)
    int raster_groups = RX_RASTER_COUNTERS;   // k_raster: frames of a launch dealt out in up to this many groups (Workspace::raster_groups)
    bool raster_groups_forced = false;        // RXC_RASTER_GROUPS set: also for frames with many tiles per CTA
    int small_min_list = RX_SMALL_MIN_LIST, small_max_pix = RX_SMALL_MAX_PIX, small_min_tris = RX_SMALL_MIN_TRIS, small_gshift = RX_SMALL_GSHIFT;   // k_raster: thread-per-record pass of tile lists at least this long, for boxes up to this many pixels
    int front_cluster_max = 64;   // setup chunks up to which the front end runs as one cluster per frame (0 = never)
    // host-output pipelining: a copy stream and two staging halves so the D2H of one sub-group of
    // frames overlaps the kernels of the next
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_render[2] = {nullptr, nullptr}, ev_copy[2] = {nullptr, nullptr};
    cudaEvent_t ev_slice[16] = {};   // one per slice of a large frame

    // stats / profiling
    rxc_stats stats = {};
    bool profiling = false;
    std::vector<PendingEvent> pending;
    std::vector<cudaEvent_t> free_events;
    RxMgpu* mgpu = nullptr;       // rxc_mgpu_* state (rx_mgpu.cu)
    RxJit* jit = nullptr;         // kernels recompiled for the current scene (rx_jit.cu): batch shaders as straight-line code, scene constants folded
    int vm_jit = 1;               // RXC_VM_JIT: 0 off, 1 compile in the background, 2 compile synchronously
    int kernel_spec = 1;          // RXC_KERNEL_SPEC: 1 = the raster kernel is also recompiled with the scene's / frame's constants folded in
    std::string spec_scene;       // -D switches of the current scene's batches and textures (recomputed when spec_dirty), of its lights (set_lights)
    std::string spec_lights;
    bool spec_dirty = true, spec_vm_opacity = false;
    std::vector<uint32_t> vm_state_report;   // per program of the current scene: rxj_state_report with the batches' bindings
    int vm_state_mode = 2;        // rxc_set_vm_state_mode: 0 = a fresh Execution per fragment (k_raster), 1 = the reference's per-tile Execution
                                  // for every scene with programs (k_raster_ordered), 2 (default) = for the scenes whose report says it can be observed
    DevBuf d_ordered;             // scratch planes of k_raster_ordered
    uint64_t ordered_frames = 0;  // frames rendered by it
    bool spec_mismatch = false;   // a specialised kernel found a scene it was not compiled for (a bug): specialisation stays off
    std::string jit_note;         // last compiler log / load failure (diagnostics)
    uint32_t jit_translated = 0;  // programs of the current scene the translator accepted
    bool pj_active = false;       // rxc_rasterize_projected is running: the front end reads `pj` instead of the geometry
    ProjectedDev pj = {};
    DevBuf d_pj_pv, d_pj_uv, d_pj_nrm, d_pj_idx, d_pj_edges, d_pj_info, d_pj_bbox;
    uint32_t async_pending = 0;   // frames of the last asynchronous group whose counters (ctx->h_counters) nobody has looked at yet
};

namespace {

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess) {                                                                  \
            ctx->err = std::string(#call) + ": " + cudaGetErrorString(e__);                        \
            return RXC_ERR_CUDA;                                                                   \
        }                                                                                          \
    } while (0)

#if RX_TMA_STORE
// The pixel buffer of a launch as a 3-D tensor (x, y, frame) of 32-bit pixels, tiled 32 x 32 x 1 with the 128 B shared-memory
// swizzle k_raster writes its tile in.  cuTensorMapEncodeTiled is a driver entry point: resolved through the runtime, so
// the library keeps no link-time dependency on libcuda.
bool encode_pixel_tensor_map(CUtensorMap* map, uint8_t* d_pixels, uint32_t cols, uint32_t rows, uint32_t pitch_px, uint32_t n_frames, uint64_t frame_stride) {
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                 const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn encode = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess) encode = (EncodeFn)fn;
        else cudaGetLastError();
    }
    if (!encode) return false;
    if (n_frames > 1 && (frame_stride & 15)) return false;
    const cuuint64_t dims[3] = {cols, rows, n_frames};
    const cuuint64_t strides[2] = {(cuuint64_t)pitch_px * 4, n_frames > 1 ? frame_stride : (cuuint64_t)pitch_px * 4 * rows};
    const cuuint32_t box[3] = {RX_TILE_W, RX_TILE_H, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    return encode(map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, d_pixels, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
#endif

int32_t fail(rxc_ctx* ctx, int32_t code, const std::string& msg) {
    ctx->err = msg;
    return code;
}

int32_t reserve(rxc_ctx* ctx, DevBuf& b, size_t bytes) {
    if (bytes <= b.cap) return RXC_OK;
    CK(cudaStreamSynchronize(ctx->stream));
    if (b.p) CK(cudaFree(b.p));
    b.p = nullptr; b.cap = 0;
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&b.p, want);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(ctx, RXC_ERR_OOM, "cudaMalloc of " + std::to_string(want) + " bytes failed"); }
    b.cap = want;
    return RXC_OK;
}

int32_t upload(rxc_ctx* ctx, DevBuf& b, const void* src, size_t bytes) {
    int32_t st = reserve(ctx, b, std::max<size_t>(bytes, 16));
    if (st != RXC_OK) return st;
    if (bytes) {
        CK(cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));  // the source is only borrowed for the call
        ctx->stats.h2d_bytes += bytes;
    }
    return RXC_OK;
}

// the same when the first `from` bytes on the device are already what `src` holds (rxc_update_scene); a buffer that has to grow is uploaded whole
int32_t upload_from(rxc_ctx* ctx, DevBuf& b, const void* src, size_t bytes, size_t from) {
    if (from == 0 || from > bytes || std::max<size_t>(bytes, 16) > b.cap) return upload(ctx, b, src, bytes);
    if (bytes > from) {
        CK(cudaMemcpyAsync((uint8_t*)b.p + from, (const uint8_t*)src + from, bytes - from, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        ctx->stats.h2d_bytes += bytes - from;
    }
    return RXC_OK;
}

void free_buf(DevBuf& b) { if (b.p) cudaFree(b.p); b.p = nullptr; b.cap = 0; }

size_t idx_at(const void* indices, uint32_t index_bytes, size_t i) {
    return index_bytes == 8 ? (size_t)((const uint64_t*)indices)[i] : (size_t)((const uint32_t*)indices)[i];
}

int32_t build_tiles(rxc_ctx* ctx, const rxc_tile* tiles, uint32_t n, std::vector<uint8_t>& arena, std::vector<DTex>& tex,
                    std::vector<DTile>& out_tiles) {
    arena.clear(); tex.clear(); out_tiles.clear();
    for (uint32_t i = 0; i < n; ++i) {
        DTile t; t.first = (uint32_t)tex.size(); t.n_frames = tiles[i].n_textures;
        if (t.n_frames && !tiles[i].textures) return fail(ctx, RXC_ERR_INVALID, "tile without textures pointer");
        for (uint32_t j = 0; j < t.n_frames; ++j) {
            const rxc_texture& x = tiles[i].textures[j];
            if (!x.data || x.width == 0 || x.height == 0) return fail(ctx, RXC_ERR_INVALID, "empty texture");
            if (x.width > 65535u || x.height > 65535u) return fail(ctx, RXC_ERR_UNSUPPORTED, "textures larger than 65535 texels per side");
            DTex d; d.offset = arena.size(); d.width = x.width; d.height = x.height; d.pad = 0;
            size_t bytes = (size_t)x.width * x.height * 4;
            bool opaque = true;
            for (size_t k = 3; k < bytes; k += 4) if (x.data[k] != 255) { opaque = false; break; }
            d.all_opaque = opaque ? 1u : 0u;
            arena.insert(arena.end(), x.data, x.data + bytes);
            arena.resize((arena.size() + 255) & ~(size_t)255);
            tex.push_back(d);
        }
        out_tiles.push_back(t);
    }
    return RXC_OK;
}

// chunk table (+ the mapmini as the last entry), sectors and linedefs; `toff` = index of the first scene texture
int32_t upload_chunks(rxc_ctx* ctx, uint32_t toff) {
    std::vector<DChunkInfo> info(ctx->h_chunks.size() + 1);
    std::vector<DSector> sectors = ctx->h_sectors;
    for (size_t i = 0; i < ctx->h_chunks.size(); ++i) {
        const HChunk& c = ctx->h_chunks[i];
        DChunkInfo d = {};
        d.sector_off = c.sector_off; d.n_sectors = c.n_sectors;
        d.origin_x = c.origin[0]; d.origin_y = c.origin[1]; d.size = c.size;
        d.terrain_tex = c.terrain_tex < 0 ? 0xFFFFFFFFu : toff + (uint32_t)c.terrain_tex;
        info[i] = d;
    }
    DChunkInfo mm = {};
    mm.sector_off = (uint32_t)sectors.size(); mm.n_sectors = (uint32_t)ctx->h_mm_sectors.size(); mm.terrain_tex = 0xFFFFFFFFu;
    info.back() = mm;
    sectors.insert(sectors.end(), ctx->h_mm_sectors.begin(), ctx->h_mm_sectors.end());
    int32_t st;
    if ((st = upload(ctx, ctx->d_chunkinfo, info.data(), info.size() * sizeof(DChunkInfo))) != RXC_OK) return st;
    if ((st = upload(ctx, ctx->d_sectors, sectors.data(), sectors.size() * sizeof(DSector))) != RXC_OK) return st;
    if ((st = upload(ctx, ctx->d_linedefs, ctx->h_linedefs.data(), ctx->h_linedefs.size() * sizeof(float))) != RXC_OK) return st;
    ctx->S.chunk_info = ctx->d_chunkinfo.as<DChunkInfo>();
    ctx->S.sectors = ctx->d_sectors.as<DSector>();
    ctx->S.linedefs = ctx->d_linedefs.as<float4>();
    ctx->S.n_scene_chunks = (uint32_t)ctx->h_chunks.size();
    ctx->S.n_sectors = (uint32_t)sectors.size();
    ctx->S.n_linedefs = (uint32_t)(ctx->h_linedefs.size() / 4);
    return RXC_OK;
}

int32_t upload_textures(rxc_ctx* ctx) {
    std::vector<uint8_t> arena = ctx->h_static_arena;
    std::vector<DTex> tex = ctx->h_static_tex;
    std::vector<DTile> tiles = ctx->h_static_tiles;
    const size_t aoff = arena.size();
    const uint32_t toff = (uint32_t)tex.size();
    arena.insert(arena.end(), ctx->h_dyn_arena.begin(), ctx->h_dyn_arena.end());
    for (DTex d : ctx->h_dyn_tex) { d.offset += aoff; tex.push_back(d); }
    for (DTile t : ctx->h_dyn_tiles) { t.first += toff; tiles.push_back(t); }
    int32_t st;
    if (arena.size() >= ((size_t)1 << 34)) return fail(ctx, RXC_ERR_UNSUPPORTED, "more than 16 GiB of texels");
    if ((st = upload(ctx, ctx->d_arena, arena.data(), arena.size())) != RXC_OK) return st;
    if ((st = upload(ctx, ctx->d_tex, tex.data(), tex.size() * sizeof(DTex))) != RXC_OK) return st;
    if ((st = upload(ctx, ctx->d_tiles, tiles.data(), tiles.size() * sizeof(DTile))) != RXC_OK) return st;
    ctx->S.arena = ctx->d_arena.as<uint8_t>();
    ctx->S.tex = ctx->d_tex.as<DTex>();
    ctx->S.tiles = ctx->d_tiles.as<DTile>();
    ctx->S.n_static_tiles = (uint32_t)ctx->h_static_tiles.size();
    ctx->S.n_dynamic_tiles = ctx->n_dyn_tiles;
    ctx->S.n_actor_tiles = ctx->n_actor_tiles;
    ctx->textures_dirty = false;
    return upload_chunks(ctx, toff);
}

// What a kernel recompiled for this scene may treat as constants (k_raster's RX_SPEC_* switches): the
// shade-descriptor flag bits k_frame_setup derives from static batch state and that are the same on every 3D batch, and
// whether any fragment can be alpha-tested at all.  Mirrors d_frame_setup (rx_kernels.cu); the specialised kernel re-checks it.
std::string scene_signature(rxc_ctx* ctx, bool any_vm_opacity) {
    if (ctx->h_b3.empty()) return std::string();
    uint32_t all = 0xFFFFFFFFu, any = 0u, unknown = 0u;
    for (const DBatch3& B : ctx->h_b3) {
        uint32_t f = B.has_normals ? RX_SD_NORMALS : 0u;
        if (B.repeat_mode == RXC_REPEAT_REPEAT_XY || B.repeat_mode == RXC_REPEAT_REPEAT_X) f |= RX_SD_REPEAT_X;
        if (B.repeat_mode == RXC_REPEAT_REPEAT_XY || B.repeat_mode == RXC_REPEAT_REPEAT_Y) f |= RX_SD_REPEAT_Y;
        if (B.source_kind == RXC_SRC_STATIC_TILE || B.source_kind == RXC_SRC_DYNAMIC_TILE) f |= RX_SD_TEXTURED;   // validated: they resolve
        else if (B.source_kind == RXC_SRC_ENTITY_TILE || B.source_kind == RXC_SRC_ITEM_TILE) unknown |= RX_SD_TEXTURED;   // may not resolve
        else if (B.source_kind == RXC_SRC_TERRAIN) unknown |= RX_SD_TERRAIN;   // depends on the chunk's terrain texture
        all &= f; any |= f;
    }
    const uint32_t bits = RX_SD_TEXTURED | RX_SD_REPEAT_X | RX_SD_REPEAT_Y | RX_SD_NORMALS | RX_SD_TERRAIN;
    const uint32_t known = bits & ~unknown & ~(all ^ any);   // the same on every batch
    std::string d = "-DRX_SPEC_FLAGS_KNOWN=" + std::to_string(known) + "u -DRX_SPEC_FLAGS_VALUE=" + std::to_string(all & known) + "u";
    bool opaque = !any_vm_opacity;
    for (const DTex& t : ctx->h_static_tex) opaque = opaque && t.all_opaque;
    for (const DTex& t : ctx->h_dyn_tex) opaque = opaque && t.all_opaque;
    if (opaque) d += " -DRX_SPEC_NO_ALPHA=1";
    return d;
}

std::string lights_signature(const std::vector<DLight>& h) {
    std::string d;
    static const size_t cap = getenv("RXC_SPEC_MAX_LIGHTS") ? (size_t)atoi(getenv("RXC_SPEC_MAX_LIGHTS")) : 4;
    if (h.size() <= cap) d = "-DRX_SPEC_NLIGHTS=" + std::to_string(h.size());   // a constant trip count; more (or a count that keeps changing) stay a run-time loop
    bool one_type = !h.empty();
    for (const DLight& l : h) one_type = one_type && l.light_type == h[0].light_type;
    if (one_type) d += std::string(d.empty() ? "" : " ") + "-DRX_SPEC_LIGHT_TYPE=" + std::to_string(h[0].light_type);
    return d;
}

int32_t upload_lights(rxc_ctx* ctx, const rxc_light* lights, uint32_t n) {
    std::vector<DLight> h(n);
    for (uint32_t i = 0; i < n; ++i) {
        const rxc_light& l = lights[i];
        DLight d = {};
        d.light_type = l.light_type; d.emitting = l.emitting; d.from_linedef = l.from_linedef;
        d.px = l.position[0]; d.py = l.position[1]; d.pz = l.position[2]; d.intensity = l.intensity;
        d.cr = l.color[0]; d.cg = l.color[1]; d.cb = l.color[2];
        d.flicker_factor = l.flicker;  // raw flicker; k_frame_setup turns it into the per-frame factor
        d.start_distance = l.start_distance; d.end_distance = l.end_distance; d.cone_angle = l.cone_angle;
        d.width = l.width; d.height = l.height;
        d.dx = l.direction[0]; d.dy = l.direction[1]; d.dz = l.direction[2];
        d.nx = l.normal[0]; d.ny = l.normal[1]; d.nz = l.normal[2];
        if (l.light_type > RXC_LIGHT_DAYLIGHT) return fail(ctx, RXC_ERR_INVALID, "unknown light type");
        h[i] = d;
    }
    int32_t st = upload(ctx, ctx->d_lights, h.data(), h.size() * sizeof(DLight));
    if (st != RXC_OK) return st;
    ctx->S.lights = ctx->d_lights.as<DLight>();
    ctx->S.n_lights = n;
    ctx->spec_lights = lights_signature(h);
    return RXC_OK;
}

cudaEvent_t get_event(rxc_ctx* ctx) {
    if (!ctx->free_events.empty()) { cudaEvent_t e = ctx->free_events.back(); ctx->free_events.pop_back(); return e; }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}

void drain_events(rxc_ctx* ctx) {  // caller has synchronized the stream
    for (auto& p : ctx->pending) {
        float ms = 0.0f;
        if (cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) ctx->stats.kernel_ms[p.cls] += ms;
        ctx->free_events.push_back(p.a);
        ctx->free_events.push_back(p.b);
    }
    ctx->pending.clear();
}

struct LaunchScope {  // counts a launch and, when profiling, brackets it with events
    rxc_ctx* ctx; int cls; cudaEvent_t a = nullptr, b = nullptr;
    LaunchScope(rxc_ctx* c, int k) : ctx(c), cls(k) {
        ctx->stats.kernel_launches++; ctx->stats.launches[cls]++;
        if (ctx->profiling) { a = get_event(ctx); b = get_event(ctx); cudaEventRecord(a, ctx->stream); }
    }
    ~LaunchScope() {
        if (ctx->profiling) { cudaEventRecord(b, ctx->stream); ctx->pending.push_back({cls, a, b}); }
    }
};

bool is_device_pointer(const void* p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

int32_t ensure_workspace(rxc_ctx* ctx, uint32_t n_frames, uint32_t tiles_per_frame) {
    SceneDev& S = ctx->S;
    Workspace& W = ctx->W;
    const size_t T = S.n_tris, nf = n_frames;
    const uint32_t want_list = std::max<uint32_t>(ctx->list_cap_min, (uint32_t)std::min<size_t>(6 * T + tiles_per_frame + 1024, 0xFFFFFFF0u));
    const uint32_t want_list2 = !S.general ? 16u : std::max<uint32_t>(ctx->list2_cap_min, (uint32_t)std::min<size_t>(8 * (size_t)S.n_rec2d + tiles_per_frame + 1024, 0xFFFFFFF0u));
    if (n_frames <= ctx->ws_frames && tiles_per_frame <= ctx->ws_tiles && want_list <= ctx->list_cap_per_frame &&
        want_list2 <= ctx->list2_cap_per_frame)
        return RXC_OK;
    ctx->list_cap_per_frame = want_list;
    ctx->list2_cap_per_frame = want_list2;
    int32_t st;
#define RES(buf, bytes) if ((st = reserve(ctx, ctx->buf, (bytes))) != RXC_OK) return st
    RES(w_frames, nf * sizeof(DFrame));
    RES(w_fb, nf * std::max<size_t>(1, S.n_b3) * sizeof(DFrameBatch));
    RES(w_fb2, nf * std::max<size_t>(1, S.n_b2) * sizeof(DFrameBatch2));
    RES(w_lights, nf * std::max<size_t>(1, S.n_lights) * sizeof(DLight));
    RES(w_counters, 2 * nf * sizeof(DCounters));   // two blocks: the pipelined host path alternates them with the staging halves
    RES(w_vis, nf * std::max<size_t>(1, 3 * T) * sizeof(TriVis));
    RES(w_shade, nf * std::max<size_t>(1, 3 * T) * sizeof(TriShade));
    RES(w_bins, nf * std::max<size_t>(1, 3 * T) * sizeof(TriBin));
    RES(w_ctot, nf * std::max<size_t>(1, S.n_chunks) * 4);
    RES(w_cbase, nf * std::max<size_t>(1, S.n_chunks) * 4);
    RES(w_clip, nf * std::max<size_t>(1, T) * sizeof(DClip));
    RES(w_large, nf * std::max<size_t>(1, 3 * T) * 4);
    RES(w_tcount, nf * (size_t)tiles_per_frame * 4);
    RES(w_tbase, nf * (size_t)tiles_per_frame * 4);
    RES(w_tfill, nf * (size_t)tiles_per_frame * 4);
    RES(w_lists, nf * (size_t)want_list * 4);
    RES(w_tcount2, nf * (size_t)tiles_per_frame * 4);
    RES(w_tbase2, nf * (size_t)tiles_per_frame * 4);
    RES(w_tfill2, nf * (size_t)tiles_per_frame * 4);
    RES(w_lists2, nf * (size_t)want_list2 * 4);
    RES(w_tri2d, nf * std::max<size_t>(1, S.n_rec2d) * sizeof(Tri2D));
    RES(w_rcounter, RX_RASTER_COUNTERS * 4);
#undef RES
    W.frames = ctx->w_frames.as<DFrame>();
    W.fb = ctx->w_fb.as<DFrameBatch>(); W.fb_stride = std::max(1u, S.n_b3);
    W.fb2 = ctx->w_fb2.as<DFrameBatch2>(); W.fb2_stride = std::max(1u, S.n_b2);
    W.lights = ctx->w_lights.as<DLight>(); W.lights_stride = std::max(1u, S.n_lights);
    W.counters = ctx->w_counters.as<DCounters>();
    W.vis = ctx->w_vis.as<TriVis>(); W.slot_stride = (uint32_t)std::max<size_t>(1, 3 * T);
    W.shade = ctx->w_shade.as<TriShade>();
    W.bins = ctx->w_bins.as<TriBin>(); W.bins_stride = (uint32_t)std::max<size_t>(1, 3 * T);
    W.chunk_new_total = ctx->w_ctot.as<uint32_t>();
    W.chunk_new_base = ctx->w_cbase.as<uint32_t>(); W.chunk_stride = std::max(1u, S.n_chunks);
    W.clip = ctx->w_clip.as<DClip>(); W.clip_stride = (uint32_t)std::max<size_t>(1, T);
    W.small_min_list = (uint32_t)std::max(1, ctx->small_min_list); W.small_max_pix = (uint32_t)ctx->small_max_pix;
    W.small_gshift = (uint32_t)ctx->small_gshift;
    W.neg_zero = -0.0f;
    W.small_min_tris = ctx->small_min_list > 0 ? (uint32_t)std::max(0, ctx->small_min_tris) : 0xFFFFFFFFu;
    W.large = ctx->w_large.as<uint32_t>(); W.large_stride = (uint32_t)std::max<size_t>(1, 3 * T);
    W.tile_count = ctx->w_tcount.as<uint32_t>();
    W.tile_base = ctx->w_tbase.as<uint32_t>();
    W.tile_fill = ctx->w_tfill.as<uint32_t>(); W.tile_stride = tiles_per_frame;
    W.lists = ctx->w_lists.as<uint32_t>(); W.list_stride = want_list;
    W.tile_count2 = ctx->w_tcount2.as<uint32_t>();
    W.tile_base2 = ctx->w_tbase2.as<uint32_t>();
    W.tile_fill2 = ctx->w_tfill2.as<uint32_t>();
    W.lists2 = ctx->w_lists2.as<uint32_t>(); W.list2_stride = want_list2;
    W.tri2d = ctx->w_tri2d.as<Tri2D>(); W.tri2d_stride = std::max(1u, S.n_rec2d);
    W.raster_counter = ctx->w_rcounter.as<uint32_t>();
    W.raster_groups = 1u;   // set per launch (launch_group)
    ctx->ws_frames = n_frames;
    ctx->ws_tiles = tiles_per_frame;
    return RXC_OK;
}

size_t workspace_bytes_per_frame(const SceneDev& S, uint32_t tiles_per_frame) {
    const size_t T = S.n_tris;
    return sizeof(DFrame) + S.n_b3 * sizeof(DFrameBatch) + 3 * T * (sizeof(TriVis) + sizeof(TriShade) + sizeof(TriBin) + 4) +
           T * sizeof(DClip) + (size_t)tiles_per_frame * 24 + (6 * T + tiles_per_frame + 1024) * 4 + S.n_rec2d * (sizeof(Tri2D) + 32) + 4096;
}

uint32_t hash_u32_host(uint32_t seed) {  // rasterizer.rs:199-207
    uint32_t state = seed;
    state = (state ^ 61u) ^ (state >> 16);
    state = state + (state << 3);
    state ^= state >> 4;
    state = state * 0x27d4eb2du;
    state ^= state >> 15;
    return state;
}

int32_t fill_frame(rxc_ctx* ctx, const rxc_frame& f, DFrame* d) {
    if (f.width == 0 || f.height == 0 || f.tile_size == 0) return fail(ctx, RXC_ERR_INVALID, "width, height and tile_size must be non-zero");
    if (f.width > 16384 || f.height > 16384) return fail(ctx, RXC_ERR_UNSUPPORTED, "frames larger than 16384 pixels per side");
    if (f.sample_mode > RXC_SAMPLE_LINEAR || f.matvec_mode > RXC_MATVEC_PLAIN_ROWS || f.background_shader > RXC_BG_GRID)
        return fail(ctx, RXC_ERR_INVALID, "bad enum value in rxc_frame");
    uint32_t y0 = f.band_y0, y1 = f.band_y1;
    if (y0 == 0 && y1 == 0) y1 = f.height;
    if (y0 >= y1 || y1 > f.height) return fail(ctx, RXC_ERR_INVALID, "bad band");
    uint32_t x0 = f.band_x0, x1 = f.band_x1;
    if (x0 == 0 && x1 == 0) x1 = f.width;
    if (x0 >= x1 || x1 > f.width) return fail(ctx, RXC_ERR_INVALID, "bad column band");
    memset(d, 0, sizeof(*d));
    memcpy(d->view, f.view, 64); memcpy(d->proj, f.projection, 64);
    memcpy(d->inv_view, f.inverse_view, 64); memcpy(d->inv_proj, f.inverse_projection, 64);
    memcpy(d->mat2d, f.matrix2d, 36);
    d->has_mat2d = f.has_matrix2d ? 1u : 0u;
    d->cam[0] = f.inverse_view[12]; d->cam[1] = f.inverse_view[13]; d->cam[2] = f.inverse_view[14];  // rasterizer.rs:98-102
    d->width_f = (float)f.width; d->height_f = (float)f.height;
    d->width = (int32_t)f.width; d->height = (int32_t)f.height;
    d->band_y0 = (int32_t)y0; d->band_y1 = (int32_t)y1;
    d->band_x0 = (int32_t)x0; d->band_x1 = (int32_t)x1;
    d->tiles_x = (int32_t)((x1 - x0 + RX_TILE_W - 1) / RX_TILE_W);
    d->tiles_y = (int32_t)((y1 - y0 + RX_TILE_H - 1) / RX_TILE_H);
    d->tile_size = std::min<uint32_t>(f.tile_size, std::max(f.width, f.height));  // one tile either way
    d->sample_mode = f.sample_mode;
    d->has_bg_color = f.has_background_color ? 1u : 0u;
    memcpy(&d->bg_color, f.background_color, 4);
    d->bg_shader = f.background_shader;
    d->grid_size = f.grid_size; d->grid_subdiv = f.grid_subdivisions;
    d->grid_off[0] = f.grid_offset[0]; d->grid_off[1] = f.grid_offset[1];
    d->has_ambient = f.has_ambient ? 1u : 0u;
    memcpy(d->ambient, f.ambient, 16);
    d->hash_anim = hash_u32_host((uint32_t)f.animation_frame);  // rasterizer.rs:208
    d->d2_active = f.d2_active ? 1u : 0u; d->d3_active = f.d3_active ? 1u : 0u;
    d->ignore_bg_shader = f.ignore_background_shader ? 1u : 0u;
    d->preserve_transparency = f.preserve_transparency ? 1u : 0u;
    d->matvec_mode = f.matvec_mode;
    d->trans2d[0] = 0.0f; d->trans2d[1] = 0.0f; d->scale2d = 1.0f;  // rasterizer.rs:104-110
    if (f.has_matrix2d) { d->trans2d[0] = f.matrix2d[6]; d->trans2d[1] = f.matrix2d[7]; d->scale2d = f.matrix2d[0]; }
    d->animation_frame = f.animation_frame;
    d->time = f.time;
    if (f.has_sky && f.sky_clouds)
        return fail(ctx, RXC_ERR_UNSUPPORTED, "the Sky node's cloud layer needs noiselib's perlin_noise_2d (not vendored with the reference); pass sky_clouds = 0");
    if (f.has_sun && f.day_factor > 0.0f) {   // rasterizer.rs:1342-1347
        d->sun_radiance = std::fmax(f.day_factor, 0.0f);
        const float m = std::sqrt(f.sun_dir[0] * f.sun_dir[0] + f.sun_dir[1] * f.sun_dir[1] + f.sun_dir[2] * f.sun_dir[2]);
        for (int k = 0; k < 3; ++k) d->sun_l[k] = -f.sun_dir[k] / m;
    }
    d->has_sky = f.has_sky ? 1u : 0u;
    memcpy(d->sky, f.sky, sizeof(d->sky));
    d->preprojected = ctx->pj_active ? 1u : 0u;
    d->has_brush = f.has_brush_preview ? 1u : 0u;
    memcpy(d->brush_pos, f.brush_position, 12);
    d->brush_radius = f.brush_radius; d->brush_falloff = f.brush_falloff;
    {   // screen_to_world (rasterizer.rs:1707-1727) as one projective map of (px+.5, py+.5, z, 1), in double
        double IP[4][4], IV[4][4], G[4][4], N[4][4] = {{2.0 / f.width, 0, 0, -1.0}, {0, -2.0 / f.height, 0, 1.0}, {0, 0, 1, 0}, {0, 0, 0, 1}};
        for (int r = 0; r < 4; ++r)
            for (int c = 0; c < 4; ++c) { IP[r][c] = f.inverse_projection[c * 4 + r]; IV[r][c] = f.inverse_view[c * 4 + r]; }
        for (int r = 0; r < 4; ++r)
            for (int c = 0; c < 4; ++c) {
                double a = 0;
                for (int k = 0; k < 4; ++k) a += IV[r][k] * IP[k][c];
                G[r][c] = a;
            }
        for (int c = 0; c < 4; ++c) { G[3][c] = IP[3][c]; }  // the divisor is view_space.w
        for (int r = 0; r < 4; ++r)
            for (int c = 0; c < 4; ++c) {
                double a = 0;
                for (int k = 0; k < 4; ++k) a += G[r][k] * N[k][c];
                d->s2w[r * 4 + c] = (float)a;
            }
    }
    return RXC_OK;
}

// Static checks of one flat program (include/rxcuda.h): every word decodes, operands and jump targets stay inside.
int32_t validate_program(rxc_ctx* ctx, const rxc_program& p, const std::string& who) {
    if (p.n_words == 0) return RXC_OK;
    if (!p.code) return fail(ctx, RXC_ERR_INVALID, who + "null code");
    if (p.entry >= p.n_words) return fail(ctx, RXC_ERR_INVALID, who + "entry outside the code");
    if (p.shade_locals > 32 || p.n_globals > 16) return fail(ctx, RXC_ERR_UNSUPPORTED, who + "more than 32 locals or 16 globals");
    if (p.n_words >= (1u << 24)) return fail(ctx, RXC_ERR_UNSUPPORTED, who + "program too long");
    for (uint32_t pc = 0; pc < p.n_words;) {
        const uint32_t w = p.code[pc++], op = w & 0xFFu, a = w >> 8;
        if (op >= RXVM_N_OPS || op == RXVM_IF || op == RXVM_FOR) return fail(ctx, RXC_ERR_INVALID, who + "bad opcode (If/For must be lowered to jumps)");
        if (op == RXVM_ALLOC || op == RXVM_ITERATE || op == RXVM_SAVE) return fail(ctx, RXC_ERR_UNSUPPORTED, who + "texture baking ops (Alloc/Iterate/Save) in a shade program");
        if (op == RXVM_PUSH) { if (pc + 3 > p.n_words) return fail(ctx, RXC_ERR_INVALID, who + "truncated Push"); pc += 3; }
        else if (op == RXVM_FUNCTION_CALL) {
            if (pc >= p.n_words || p.code[pc] >= p.n_words) return fail(ctx, RXC_ERR_INVALID, who + "bad call target");
            if ((a >> 8) > 32) return fail(ctx, RXC_ERR_UNSUPPORTED, who + "more than 32 locals in a call");
            ++pc;
        } else if ((op == RXVM_JZ || op == RXVM_JMP) && a > p.n_words) return fail(ctx, RXC_ERR_INVALID, who + "jump outside the code");
        else if ((op == RXVM_LOAD_GLOBAL || op == RXVM_STORE_GLOBAL) && a >= p.n_globals) return fail(ctx, RXC_ERR_INDEX, who + "global index out of range (reference panics)");
    }
    if ((p.code[p.n_words - 1] & 0xFFu) != RXVM_END) return fail(ctx, RXC_ERR_INVALID, who + "code must end with End");
    return RXC_OK;
}

int32_t upload_vm(rxc_ctx* ctx, const rxc_scene* sc) {
    std::vector<uint32_t> code;
    std::vector<DProgram> progs(sc->n_shaders);
    int32_t st;
    for (uint32_t i = 0; i < sc->n_shaders; ++i) {
        const rxc_program& p = sc->shaders[i];
        if ((st = validate_program(ctx, p, "shader " + std::to_string(i) + ": ")) != RXC_OK) return st;
        DProgram d = {};
        d.code_off = (uint32_t)code.size(); d.n_words = p.n_words; d.entry = p.entry; d.shade_locals = p.shade_locals;
        d.n_globals = p.n_globals; d.sets_opacity = p.sets_opacity ? 1u : 0u;
        d.jit_index = 0xFFFFFFFFu;
        code.insert(code.end(), p.code, p.code + p.n_words);
        progs[i] = d;
    }
    // the programs as straight-line C++ for the JIT-compiled kernel variant (the interpreter runs what the translator declines)
    ctx->jit_translated = 0;
    if (ctx->vm_jit) {
        std::string generated;
        std::vector<uint32_t> jit_index;
        if (sc->n_shaders && rxj_generate(sc->shaders, sc->n_shaders, &generated, &jit_index)) {
            for (uint32_t i = 0; i < sc->n_shaders; ++i) { progs[i].jit_index = jit_index[i]; if (jit_index[i] != 0xFFFFFFFFu) ++ctx->jit_translated; }
        } else {
            generated.clear();
        }
        if (ctx->jit) rxj_set_programs(ctx->jit, generated, ctx->vm_jit);
        else if (!generated.empty() || ctx->kernel_spec) ctx->jit = rxj_create(generated, ctx->vm_jit);
    } else if (ctx->jit) {
        rxj_set_programs(ctx->jit, std::string(), 0);
    }
    ctx->spec_vm_opacity = false;
    for (uint32_t i = 0; i < sc->n_shaders; ++i) ctx->spec_vm_opacity = ctx->spec_vm_opacity || sc->shaders[i].sets_opacity;
    ctx->spec_dirty = true;
    std::vector<float> patdata;
    std::vector<DPattern> pats;
    auto add_patterns = [&](const rxc_pattern* list, uint32_t n) -> int32_t {
        for (uint32_t i = 0; i < n; ++i) {
            const rxc_pattern& q = list[i];
            if (!q.data || q.width == 0 || q.height == 0 || q.width > 32768 || q.height > 32768) return fail(ctx, RXC_ERR_INVALID, "bad pattern texture");
            DPattern d = {(uint32_t)(patdata.size() / 3), q.width, q.height, 0};
            patdata.insert(patdata.end(), q.data, q.data + (size_t)q.width * q.height * 3);
            pats.push_back(d);
        }
        return RXC_OK;
    };
    if ((st = add_patterns(sc->patterns, sc->n_patterns)) != RXC_OK) return st;
    if ((st = add_patterns(sc->patterns_normal, sc->n_patterns_normal)) != RXC_OK) return st;
    if ((st = upload(ctx, ctx->d_vm_code, code.data(), code.size() * 4)) != RXC_OK) return st;
    if ((st = upload(ctx, ctx->d_vm_programs, progs.data(), progs.size() * sizeof(DProgram))) != RXC_OK) return st;
    if ((st = upload(ctx, ctx->d_vm_patdata, patdata.data(), patdata.size() * 4)) != RXC_OK) return st;
    if ((st = upload(ctx, ctx->d_vm_patterns, pats.data(), pats.size() * sizeof(DPattern))) != RXC_OK) return st;
    if ((st = upload(ctx, ctx->d_vm_palette, sc->palette, (size_t)sc->n_palette * 16)) != RXC_OK) return st;
    VmDev& vm = ctx->S.vm;
    vm.code = ctx->d_vm_code.as<uint32_t>(); vm.programs = ctx->d_vm_programs.as<DProgram>();
    vm.pattern_data = ctx->d_vm_patdata.as<float>(); vm.patterns = ctx->d_vm_patterns.as<DPattern>();
    vm.palette = ctx->d_vm_palette.as<float4>();
    vm.n_programs = sc->n_shaders; vm.n_patterns = sc->n_patterns; vm.n_patterns_normal = sc->n_patterns_normal; vm.n_palette = sc->n_palette;
    return RXC_OK;
}

// batch.shader -> absolute program index (scene.shaders for batches without a chunk, else the chunk's range;
// rasterizer.rs:1281-1285), -1 when nothing runs
int32_t resolve_program(const rxc_scene* sc, int32_t shader, int32_t chunk) {
    if (shader < 0) return -1;
    if (chunk >= 0) {
        const rxc_chunk& c = sc->chunks[chunk];
        return (uint32_t)shader < c.n_shaders ? (int32_t)(c.shader_base + (uint32_t)shader) : -1;
    }
    return (uint32_t)shader < sc->n_scene_shaders ? shader : -1;
}

int32_t validate_sources(rxc_ctx* ctx) {
    // 3D: the reference indexes tile_list / dynamic_textures directly and panics (rasterizer.rs:1103,:1126);
    // entity/item tiles that do not resolve sample as transparent, a resolved tile without frames panics on `% 0`
    auto actor_ok = [&](uint32_t index) {
        if (index == 0xFFFFFFFFu) return true;
        const size_t k = (size_t)ctx->n_dyn_tiles + index;
        return index < ctx->n_actor_tiles && ctx->h_dyn_tiles[k].n_frames != 0;
    };
    for (size_t i = 0; i < ctx->h_b3.size(); ++i) {
        const DBatch3& b = ctx->h_b3[i];
        if (b.source_kind == RXC_SRC_STATIC_TILE) {
            if (b.source_index >= ctx->h_static_tiles.size() || ctx->h_static_tiles[b.source_index].n_frames == 0)
                return fail(ctx, RXC_ERR_INDEX, "3D batch " + std::to_string(i) + ": StaticTileIndex out of range (reference panics)");
        } else if (b.source_kind == RXC_SRC_DYNAMIC_TILE) {
            if (b.source_index >= ctx->n_dyn_tiles || ctx->h_dyn_tiles[b.source_index].n_frames == 0)
                return fail(ctx, RXC_ERR_INDEX, "3D batch " + std::to_string(i) + ": DynamicTileIndex out of range (reference panics)");
        } else if (b.source_kind == RXC_SRC_ENTITY_TILE || b.source_kind == RXC_SRC_ITEM_TILE) {
            if (!actor_ok(b.source_index)) return fail(ctx, RXC_ERR_INDEX, "3D batch " + std::to_string(i) + ": bad actor tile index");
        }
    }
    for (size_t i = 0; i < ctx->h_b2.size(); ++i) {
        const DBatch2& b = ctx->h_b2[i];
        if (b.source_kind == RXC_SRC_STATIC_TILE && b.source_index < ctx->h_static_tiles.size() && ctx->h_static_tiles[b.source_index].n_frames == 0)
            return fail(ctx, RXC_ERR_INDEX, "2D batch " + std::to_string(i) + ": tile without textures (reference panics on % 0)");
        if (b.source_kind == RXC_SRC_DYNAMIC_TILE && b.source_index < ctx->n_dyn_tiles && ctx->h_dyn_tiles[b.source_index].n_frames == 0)
            return fail(ctx, RXC_ERR_INDEX, "2D batch " + std::to_string(i) + ": tile without textures (reference panics on % 0)");
        if ((b.source_kind == RXC_SRC_ENTITY_TILE || b.source_kind == RXC_SRC_ITEM_TILE) && !actor_ok(b.source_index))
            return fail(ctx, RXC_ERR_INDEX, "2D batch " + std::to_string(i) + ": bad actor tile index");
    }
    return RXC_OK;
}

// Runs frames [first, first+n) of one group through the kernel sequence.
// `slices` > 1 rasterises every frame as that many horizontal slices (ranges of GPU tile rows), one k_raster launch
// each, and calls `after_slice(first_row, end_row)` (rows relative to the frame's band) once a slice has been
// enqueued, so that the caller can start draining it while the next slice renders.
template <class AfterSlice>
int32_t launch_group(rxc_ctx* ctx, const DFrame* h_frames, DCounters* h_counters, uint32_t n, uint8_t* d_pixels, uint64_t stride,
                     uint32_t pitch_px, uint32_t* d_owner, float* d_depth, uint32_t slices, AfterSlice after_slice) {
    SceneDev& S = ctx->S;
    const uint32_t tiles_per_frame = (uint32_t)h_frames[0].tiles_x * (uint32_t)h_frames[0].tiles_y;
    CK(cudaMemcpyAsync(ctx->W.frames, h_frames, n * sizeof(DFrame), cudaMemcpyHostToDevice, ctx->stream));
    ctx->stats.h2d_bytes += n * sizeof(DFrame);
    const int wide = ctx->sm_count * 8;
    auto grid_for = [&](size_t items, int per_block) { return (int)std::max<size_t>(1, std::min<size_t>((items + per_block - 1) / per_block, (size_t)wide)); };
    // tiny scenes (a couple of batches): the seven front-end launches cost more than their work, one CTA per frame
    // runs them back to back (measured: 22 us instead of ~70 us for the cube; from ~5 setup chunks on the serial
    // walk of one CTA loses against the parallel launches)
    const bool pj = ctx->pj_active;   // host-projected 3D batches: always the separate kernels
    const bool small_front = !pj && !S.general && S.n_chunks <= 2 && S.n_b2 <= 4 && tiles_per_frame <= 16384;
    // mid-sized scenes: one cluster of CTAs per frame, cluster barriers instead of kernel boundaries
    const bool cluster_front = !pj && !small_front && ctx->front_cluster_max != 0 && S.n_chunks <= (uint32_t)ctx->front_cluster_max &&
                               S.n_b2 <= 64 && tiles_per_frame <= 65536 && (!S.general || S.n_rec2d <= 512);   // long sorted 2D lists want k_list_sort's CTA per tile (834 records: no gain)
    if (small_front) { LaunchScope l(ctx, RXK_FRONT_SMALL); CK(rxk_front_small(S, ctx->W, n, tiles_per_frame, ctx->stream)); }
    else if (cluster_front) { LaunchScope l(ctx, RXK_FRONT_SMALL); CK(rxk_front_cluster(S, ctx->W, n, tiles_per_frame, (uint32_t)ctx->front_stop, ctx->stream)); }
    else { LaunchScope l(ctx, RXK_FRAME_SETUP); CK(rxk_frame_setup(S, ctx->W, n, tiles_per_frame, ctx->stream)); }
    if (S.n_tris && !small_front && !cluster_front) {
        if (pj) {
            LaunchScope l(ctx, RXK_TRI_SETUP); CK(rxk_tri_setup_projected(S, ctx->W, ctx->pj, ctx->stream));
        } else {
            { LaunchScope l(ctx, RXK_TRI_SETUP); CK(rxk_tri_setup(S, ctx->W, n, ctx->stream)); }
            { LaunchScope l(ctx, RXK_BATCH_FINALIZE); CK(rxk_batch_finalize(S, ctx->W, n, ctx->stream)); }
            { LaunchScope l(ctx, RXK_CLIP_EMIT); CK(rxk_clip_emit(S, ctx->W, n, grid_for(std::min<size_t>(S.n_tris, 65536), 128), ctx->stream)); }
        }
        { LaunchScope l(ctx, RXK_BIN_COUNT); CK(rxk_bin_count(S, ctx->W, n, grid_for((size_t)S.n_tris + S.n_tris / 8, 256), ctx->stream)); }
        if (S.general) { LaunchScope l(ctx, RXK_BIN_COUNT); CK(rxk_bin_large(S, ctx->W, n, 0, ctx->sm_count, ctx->stream)); }
        { LaunchScope l(ctx, RXK_TILE_ALLOC); CK(rxk_tile_alloc(S, ctx->W, n, tiles_per_frame, 0, S.general ? 1 : 0, ctx->stream)); }
        { LaunchScope l(ctx, RXK_BIN_FILL); CK(rxk_bin_fill(S, ctx->W, n, grid_for((size_t)S.n_tris + S.n_tris / 8, 256), ctx->stream)); }
        if (S.general) { LaunchScope l(ctx, RXK_BIN_FILL); CK(rxk_bin_large(S, ctx->W, n, 1, ctx->sm_count, ctx->stream)); }
        if (S.general) { LaunchScope l(ctx, RXK_LIST_SORT); CK(rxk_list_sort(S, ctx->W, n, tiles_per_frame, 0, ctx->stream)); }
    }
    if (S.general && S.n_rec2d && !cluster_front) {  // 2D records into sorted per-tile lists
        { LaunchScope l(ctx, RXK_BIN2D); CK(rxk_bin2d(S, ctx->W, n, 0, ctx->stream)); }
        { LaunchScope l(ctx, RXK_TILE_ALLOC); CK(rxk_tile_alloc(S, ctx->W, n, tiles_per_frame, 1, 1, ctx->stream)); }
        { LaunchScope l(ctx, RXK_BIN2D); CK(rxk_bin2d(S, ctx->W, n, 1, ctx->stream)); }
        { LaunchScope l(ctx, RXK_LIST_SORT); CK(rxk_list_sort(S, ctx->W, n, tiles_per_frame, 1, ctx->stream)); }
    }
    RasterOut out;
    out.pixels = d_pixels; out.frame_stride = stride; out.owner = d_owner; out.depth = d_depth;
    out.pitch = pitch_px ? pitch_px : (uint32_t)(h_frames[0].band_x1 - h_frames[0].band_x0);
    out.vec_store = (((uintptr_t)d_pixels & 15) == 0 && (stride & 15) == 0 && ((out.pitch * 4) & 15) == 0) ? 1u : 0u;
    out.tma_store = 0u;
    memset(&out.tmap, 0, sizeof(out.tmap));
#if RX_TMA_STORE
    if (ctx->tma_store && out.vec_store && !d_owner && !d_depth)
        out.tma_store = encode_pixel_tensor_map(&out.tmap, d_pixels, (uint32_t)(h_frames[0].band_x1 - h_frames[0].band_x0),
                                                (uint32_t)(h_frames[0].band_y1 - h_frames[0].band_y0), out.pitch, n, stride) ? 1u : 0u;
#endif
    int sample_mode = (int)h_frames[0].sample_mode;
    for (uint32_t i = 1; i < n; ++i) if ((int)h_frames[i].sample_mode != sample_mode) sample_mode = 2;
    const uint32_t tiles_x = (uint32_t)h_frames[0].tiles_x, tiles_y = (uint32_t)h_frames[0].tiles_y;
    const uint32_t rows_total = (uint32_t)(h_frames[0].band_y1 - h_frames[0].band_y0);
    slices = std::max(1u, std::min(std::min(slices, tiles_y), (uint32_t)RX_RASTER_COUNTERS));
    const uint32_t rows_per_slice = (tiles_y + slices - 1) / slices;   // in tile rows
    // the reference's order and per-tile Execution instead (k_raster_ordered), when asked for and the scene has programs to observe it
    bool ordered = false;
    if (ctx->vm_state_mode && S.general && S.vm.n_programs) {
        ordered = ctx->vm_state_mode == 1;
        for (uint32_t r : ctx->vm_state_report) ordered = ordered || r != 0u;   // can observe it, or could not be analysed
    }
    if (ordered) {
        // whole frames of one size, API tiles over at most 64 of the 32x32 device tiles whose lists the kernel merges; what it cannot
        // render is an error when the mode was forced (1) and goes to the fast kernel (a fresh Execution per fragment) in auto mode (2)
        const DFrame& F0 = h_frames[0];
        const char* why = nullptr;
        for (uint32_t i = 0; i < n && !why; ++i) {
            const DFrame& F = h_frames[i];
            if (F.band_x0 != 0 || F.band_y0 != 0 || F.band_x1 != F.width || F.band_y1 != F.height)
                why = "reference-order mode (rxc_set_vm_state_mode) renders whole frames only (an API tile cut by a band would lose the state of its other part)";
            else if (F.tile_size != F0.tile_size || F.width != F0.width || F.height != F0.height)
                why = "reference-order mode: the frames of a batch must share width, height and tile_size";
        }
        const uint32_t ts_probe = std::max<uint32_t>(1u, (uint32_t)F0.tile_size);
        const uint32_t sx = std::min<uint32_t>(ts_probe, (uint32_t)F0.width), sy = std::min<uint32_t>(ts_probe, (uint32_t)F0.height);
        if (!why && ((sx + 30u) / 32u + 1u) * ((sy + 30u) / 32u + 1u) > 64u)
            why = "reference-order mode: tile_size too large (an API tile may span at most 64 device tiles: up to 224 x 224 pixels)";
        if (why) {
            if (ctx->vm_state_mode == 1) return fail(ctx, RXC_ERR_UNSUPPORTED, why);
            ordered = false;
        }
    }
    if (ordered) {
        const DFrame& F0 = h_frames[0];
        const uint32_t ts = std::max<uint32_t>(1u, (uint32_t)F0.tile_size);
        const size_t px = (size_t)F0.width * (size_t)F0.height;
        int32_t st = reserve(ctx, ctx->d_ordered, (size_t)n * px * 24);
        if (st != RXC_OK) return st;
        uint8_t* base = ctx->d_ordered.as<uint8_t>();
        const size_t plane = (size_t)n * px * 4;
        const uint32_t api_tiles = ((uint32_t)F0.width + ts - 1) / ts * (((uint32_t)F0.height + ts - 1) / ts);
        void* ordered_kernel = nullptr;   // the kernel with the scene's programs compiled instead of interpreted, once it is there
        if (ctx->jit && ctx->jit_translated) {
            std::string note;
            ordered_kernel = rxj_kernel(ctx->jit, -2, false, 2, std::string(), &note);
            if (!note.empty()) ctx->jit_note = note;
        }
        { LaunchScope l(ctx, RXK_RASTER);
          CK(rxk_raster_ordered(S, ctx->W, out, n, api_tiles, (float*)base, (float*)(base + plane), (uint32_t*)(base + 2 * plane), (uint32_t*)(base + 3 * plane),
                                (uint32_t*)(base + 4 * plane), (uint32_t*)(base + 5 * plane), px, ctx->stream, ordered_kernel)); }
        const int32_t sa = after_slice(0u, rows_total);
        if (sa != RXC_OK) return sa;
        if (h_counters) CK(cudaMemcpyAsync(h_counters, ctx->W.counters, n * sizeof(DCounters), cudaMemcpyDeviceToHost, ctx->stream));
        ctx->stats.frames += n;
        ctx->ordered_frames += n;
        return RXC_OK;
    }
    // which k_raster: the library's generic instantiation, or the one recompiled for this scene and these frames (rx_jit.cu) once it is there
    void* jit_kernel = nullptr;
    const int raster_mode = rxk_raster_mode(S, ctx->W);
    const bool spec = ctx->kernel_spec && !ctx->spec_mismatch;
    if (spec && ctx->spec_dirty) { ctx->spec_scene = scene_signature(ctx, ctx->spec_vm_opacity); ctx->spec_dirty = false; }
    if (ctx->jit && (raster_mode == 2 || spec)) {
        std::string defines, note;
        if (spec) {
            // what all frames of this launch agree on (RX_FS_* of rx_kernels.cu)
            uint32_t known = 63u, value = 0u;
            auto bits = [&](const DFrame& F) {
                return (F.has_ambient ? 1u : 0u) | (F.sun_radiance > 0.0f ? 2u : 0u) | ((F.has_sky | F.has_brush) ? 4u : 0u) | (F.d3_active ? 8u : 0u) |
                       ((F.d2_active && S.n_rec2d != 0u) ? 16u : 0u) | (S.n_sectors ? 32u : 0u);
            };
            value = bits(h_frames[0]);
            for (uint32_t i = 1; i < n; ++i) known &= ~(bits(h_frames[i]) ^ value);
            defines = "-DRX_SPEC_ACTIVE=1" + (ctx->spec_scene.empty() ? "" : " " + ctx->spec_scene) + (ctx->spec_lights.empty() ? "" : " " + ctx->spec_lights) +
                      " -DRX_SPEC_FRAME_KNOWN=" + std::to_string(known) + "u -DRX_SPEC_FRAME_VALUE=" + std::to_string(value & known) + "u";
            if (const char* e = getenv("RXC_JIT_EXTRA")) defines += std::string(" ") + e;   // experiments: further -D switches for the recompiled kernel
        }
        jit_kernel = rxj_kernel(ctx->jit, sample_mode, d_owner || d_depth, raster_mode, defines, &note);
        if (!note.empty()) ctx->jit_note = note;
    }
    // frame groups of k_raster's work fetch: a whole-frame launch of several frames deals them out in runs of consecutive frames
    // where a frame gives a CTA only a few tiles (1080p: 2040 tiles over 592 CTAs; the 4096-frame sweep 0.1663 -> 0.1581 s); at 4K (13.7 tiles
    // per CTA and frame) one counter over all frames measured 0.5 % faster, so the groups stay off there
    const bool few_tiles = (size_t)tiles_per_frame <= (size_t)5 * (size_t)ctx->sm_count * (size_t)ctx->raster_blocks_per_sm;
    ctx->W.raster_groups = (slices <= 1u && n > 1u && ctx->raster_groups > 1 && (few_tiles || ctx->raster_groups_forced))
                               ? std::min<uint32_t>(std::min<uint32_t>(n, (uint32_t)ctx->raster_groups), (uint32_t)RX_RASTER_COUNTERS) : 1u;
    for (uint32_t k = 0, ty0 = 0; ty0 < tiles_y; ++k, ty0 += rows_per_slice) {
        const uint32_t ty1 = std::min(tiles_y, ty0 + rows_per_slice);
        const size_t slice_tiles = (size_t)(ty1 - ty0) * tiles_x;
        const int grid = (int)std::max<size_t>(1, std::min<size_t>((size_t)n * slice_tiles, (size_t)ctx->sm_count * ctx->raster_blocks_per_sm));
        { LaunchScope l(ctx, RXK_RASTER); CK(rxk_raster(S, ctx->W, out, n, ty0 * tiles_x, (uint32_t)slice_tiles, k, sample_mode, grid, ctx->stream, jit_kernel)); }
        const int32_t st = after_slice(ty0 * (uint32_t)RX_TILE_H, std::min(rows_total, ty1 * (uint32_t)RX_TILE_H));
        if (st != RXC_OK) return st;
    }
    // the counters follow the pixels: on the render stream, or (h_counters == nullptr) copied by the caller on its copy stream
    if (h_counters) CK(cudaMemcpyAsync(h_counters, ctx->W.counters, n * sizeof(DCounters), cudaMemcpyDeviceToHost, ctx->stream));
    ctx->stats.frames += n;
    return RXC_OK;
}

// checks the counters copied back by launch_group (stream must be synchronized); 1 = retry needed
int32_t check_group(rxc_ctx* ctx, const DCounters* h_counters, uint32_t n, bool* retry) {
    *retry = false;
    uint32_t ov = 0, need = 0, need2 = 0;
    for (uint32_t i = 0; i < n; ++i) {
        ov |= h_counters[i].overflow;
        need = std::max(need, h_counters[i].list_cursor);
        need2 = std::max(need2, h_counters[i].list_cursor2);
    }
    const DCounters& c = h_counters[n - 1];
    ctx->stats.last_binned_refs = c.list_cursor;
    ctx->stats.last_large_tris = c.n_large;
    ctx->stats.last_clipped_tris = c.n_new_slots;
    ctx->stats.last_visible_tris = c.n_visible;
    if (ov & 1u) {  // tile-list arena too small: grow to twice what the frame asked for and run it again
        ctx->list_cap_min = 2 * need + 1024;
        *retry = true;
    }
    if (ov & 8u) {  // 2D tile-list arena
        ctx->list2_cap_min = 2 * need2 + 1024;
        *retry = true;
    }
    if (ov & 6u) return fail(ctx, RXC_ERR_OOM, "internal list overflow (large/clip); this is a bug");
    if (ov & 32u) {
        ctx->spec_mismatch = true;   // never expected: the library's generic kernels from here on
        return fail(ctx, RXC_ERR_CUDA, "a scene-specialised raster kernel ran on a scene or frame it was not compiled for (internal error; specialisation is now off for this context, render the frame again)");
    }
    if (ov & 16u) return fail(ctx, RXC_ERR_UNSUPPORTED, "a batch shader exceeded a device VM limit (stack 32, call depth 8, 2^20 ops) or popped an empty stack");
    return RXC_OK;
}

// The counters of the last asynchronous group, if nobody has looked at them yet (the stream must be synchronized).
// Reported once: a tile-list overflow there means geometry was dropped from those frames.
int32_t check_async(rxc_ctx* ctx) {
    const uint32_t n = ctx->async_pending;
    ctx->async_pending = 0;
    if (!n || !ctx->h_counters) return RXC_OK;
    bool retry = false;
    const int32_t st = check_group(ctx, ctx->h_counters, n, &retry);
    if (st != RXC_OK) return st;
    if (retry) return fail(ctx, RXC_ERR_OOM, "a tile-list arena overflowed during an asynchronous call (its frames are incomplete); the arena has been grown, render them again");
    return RXC_OK;
}

int32_t rasterize_impl(rxc_ctx* ctx, const rxc_frame* frames, uint32_t n_frames, uint8_t* pixels, uint64_t stride, uint32_t* owner,
                       float* depth, bool sync, uint64_t pitch_bytes = 0) {
    if (!ctx) return RXC_ERR_INVALID;
    if (!frames || n_frames == 0 || !pixels) return fail(ctx, RXC_ERR_INVALID, "frames and pixels are required");
    if (!ctx->have_scene) return fail(ctx, RXC_ERR_INVALID, "rxc_set_scene has not been called");
    if ((owner || depth) && n_frames != 1) return fail(ctx, RXC_ERR_INVALID, "owner/depth planes are single-frame outputs");
    CK(cudaSetDevice(ctx->device));
    int32_t st;
    if (ctx->textures_dirty && (st = upload_textures(ctx)) != RXC_OK) return st;
    if ((st = validate_sources(ctx)) != RXC_OK) return st;

    // all frames of a batch share the geometry of the output
    const rxc_frame& f0 = frames[0];
    for (uint32_t i = 1; i < n_frames; ++i)
        if (frames[i].width != f0.width || frames[i].height != f0.height || frames[i].band_y0 != f0.band_y0 || frames[i].band_y1 != f0.band_y1 ||
            frames[i].band_x0 != f0.band_x0 || frames[i].band_x1 != f0.band_x1)
            return fail(ctx, RXC_ERR_INVALID, "all frames of a batch must share width/height/band");
    DFrame probe;
    if ((st = fill_frame(ctx, f0, &probe)) != RXC_OK) return st;
    const uint32_t rows = (uint32_t)(probe.band_y1 - probe.band_y0);
    const uint32_t cols = (uint32_t)(probe.band_x1 - probe.band_x0);
    const uint64_t frame_bytes = (uint64_t)cols * rows * 4;
    if (n_frames > 1 && stride < frame_bytes) return fail(ctx, RXC_ERR_INVALID, "frame_stride_bytes smaller than a frame");
    if (pitch_bytes && ((pitch_bytes & 3) || pitch_bytes < (uint64_t)cols * 4 || pitch_bytes > ((uint64_t)1 << 20)))
        return fail(ctx, RXC_ERR_INVALID, "row pitch must be a multiple of 4 bytes and at least one row of the rendered rectangle");
    const uint32_t pitch_px = (uint32_t)(pitch_bytes / 4);   // 0 = tight rows
    const uint32_t tiles_per_frame = (uint32_t)probe.tiles_x * (uint32_t)probe.tiles_y;

    // frames are processed in groups sized to a workspace budget
    const size_t per_frame = workspace_bytes_per_frame(ctx->S, tiles_per_frame);
    const size_t budget = (size_t)3 << 30;
    uint32_t group = (uint32_t)std::max<size_t>(1, std::min<size_t>(n_frames, budget / std::max<size_t>(1, per_frame)));
    group = std::min(group, 1024u);

    const bool dev_px = is_device_pointer(pixels);
    if (pitch_px && !dev_px) return fail(ctx, RXC_ERR_INVALID, "a row pitch needs a device pixel buffer");
    const bool dev_owner = owner && is_device_pointer(owner);
    const bool dev_depth = depth && is_device_pointer(depth);

    if (ctx->h_frames_cap < group) {
        CK(cudaStreamSynchronize(ctx->stream));
        if ((st = check_async(ctx)) != RXC_OK) return st;
        if (ctx->h_frames) cudaFreeHost(ctx->h_frames);
        if (ctx->h_counters) cudaFreeHost(ctx->h_counters);
        CK(cudaMallocHost((void**)&ctx->h_frames, group * sizeof(DFrame)));
        CK(cudaMallocHost((void**)&ctx->h_counters, group * sizeof(DCounters)));
        ctx->h_frames_cap = group;
    }

    // Host pixels: the frames are rendered in pieces of about 8 MB -- sub-groups of small frames, horizontal slices of
    // large ones (one front-end pass per frame, one k_raster launch per slice) -- into alternating halves of the
    // device staging buffer, while the copy stream drains the finished pieces over PCIe.  What stays exposed is the
    // first piece: front end + one slice instead of a whole frame (or group of frames).
    if (!dev_px && sync && !owner && !depth && (n_frames > 1 || frame_bytes >= ((uint64_t)2 << 20))) {
        const uint64_t piece = (uint64_t)ctx->piece_mb << 20;
        const uint32_t sub = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(group, piece / std::max<uint64_t>(1, frame_bytes)));
        const uint64_t slice_bytes = (uint64_t)ctx->slice_mb << 20;
        // slices only pay off from about 16 MB per frame on: every D2H copy carries ~10 us of fixed cost
        const uint32_t slices = (sub > 1 || slice_bytes == 0 || frame_bytes < ((uint64_t)16 << 20)) ? 1u
                                : (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(16, (frame_bytes + slice_bytes / 2) / slice_bytes));
        const uint64_t row_bytes = (uint64_t)cols * 4;
        if (ctx->h_frames_cap < n_frames) {
            CK(cudaStreamSynchronize(ctx->stream));
            if ((st = check_async(ctx)) != RXC_OK) return st;
            if (ctx->h_frames) cudaFreeHost(ctx->h_frames);
            if (ctx->h_counters) cudaFreeHost(ctx->h_counters);
            ctx->h_frames = nullptr; ctx->h_counters = nullptr; ctx->h_frames_cap = 0;
            CK(cudaMallocHost((void**)&ctx->h_frames, n_frames * sizeof(DFrame)));
            CK(cudaMallocHost((void**)&ctx->h_counters, n_frames * sizeof(DCounters)));
            ctx->h_frames_cap = n_frames;
        }
        if (!ctx->copy_stream) {
            CK(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
            for (int i = 0; i < 2; ++i) {
                CK(cudaEventCreateWithFlags(&ctx->ev_render[i], cudaEventDisableTiming));
                CK(cudaEventCreateWithFlags(&ctx->ev_copy[i], cudaEventDisableTiming));
            }
            for (int i = 0; i < 16; ++i) CK(cudaEventCreateWithFlags(&ctx->ev_slice[i], cudaEventDisableTiming));
        }
        CK(cudaStreamSynchronize(ctx->stream));  // the pinned frame block may still feed an earlier asynchronous call
        if ((st = check_async(ctx)) != RXC_OK) return st;
        for (uint32_t i = 0; i < n_frames; ++i)
            if ((st = fill_frame(ctx, frames[i], &ctx->h_frames[i])) != RXC_OK) return st;
        for (int attempt = 0; attempt < 4; ++attempt) {
            // the first piece is small (its render is exposed), the later ones grow to four times its size (about 32 MB):
            // they render while their predecessor drains, and every D2H copy carries a few microseconds of fixed cost
            const uint32_t sub_big = (uint32_t)std::max<uint64_t>(sub, std::min<uint64_t>(group, (4 * piece) / std::max<uint64_t>(1, frame_bytes)));
            if ((st = ensure_workspace(ctx, sub_big, tiles_per_frame)) != RXC_OK) return st;
            if ((st = reserve(ctx, ctx->d_out_px, (size_t)2 * sub_big * frame_bytes)) != RXC_OK) return st;
            uint32_t k = 0;
            for (uint32_t first = 0, n = 0; first < n_frames; first += n, ++k) {
                // doubling: a piece renders faster than it drains, so the next one may be twice as large without a bubble
                n = std::min(std::min<uint32_t>(sub_big, k < 16 ? sub << k : sub_big), n_frames - first);
                const int half = (int)(k & 1u);
                uint8_t* d_px = ctx->d_out_px.as<uint8_t>() + (size_t)half * sub_big * frame_bytes;
                if (k >= 2) CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_copy[half], 0));  // the half has been drained
                // only the first frame of a call is sliced: it is the one whose render is exposed; the later ones render
                // while their predecessor drains and leave as one copy (each D2H copy carries ~6 us of fixed cost)
                const uint32_t slices_now = k == 0 ? slices : 1u;
                uint32_t slice_no = 0;
                auto drain = [&](uint32_t row0, uint32_t row1) -> int32_t {
                    cudaEvent_t ev = slices_now > 1 ? ctx->ev_slice[slice_no++ & 15u] : ctx->ev_render[half];
                    CK(cudaEventRecord(ev, ctx->stream));
                    CK(cudaStreamWaitEvent(ctx->copy_stream, ev, 0));
                    if (slices_now > 1) {   // n == 1: rows [row0, row1) of the frame
                        CK(cudaMemcpyAsync(pixels + (uint64_t)first * stride + row0 * row_bytes, d_px + row0 * row_bytes, (size_t)(row1 - row0) * row_bytes,
                                           cudaMemcpyDeviceToHost, ctx->copy_stream));
                    } else if (stride == frame_bytes) {
                        CK(cudaMemcpyAsync(pixels + (uint64_t)first * stride, d_px, (size_t)n * frame_bytes, cudaMemcpyDeviceToHost, ctx->copy_stream));
                    } else {
                        CK(cudaMemcpy2DAsync(pixels + (uint64_t)first * stride, stride, d_px, frame_bytes, frame_bytes, n, cudaMemcpyDeviceToHost, ctx->copy_stream));
                    }
                    return RXC_OK;
                };
                // the counter block alternates with the staging half, and is read back behind the pixels on the copy stream
                // (a small D2H on the render stream would queue behind the pixel copies on the same DMA engine and stall it)
                ctx->W.counters = ctx->w_counters.as<DCounters>() + (size_t)half * ctx->ws_frames;
                st = launch_group(ctx, ctx->h_frames + first, nullptr, n, d_px, frame_bytes, 0u, nullptr, nullptr, slices_now, drain);
                DCounters* d_counters = ctx->W.counters;
                ctx->W.counters = ctx->w_counters.as<DCounters>();
                if (st != RXC_OK) return st;
                CK(cudaMemcpyAsync(ctx->h_counters + first, d_counters, n * sizeof(DCounters), cudaMemcpyDeviceToHost, ctx->copy_stream));
                CK(cudaEventRecord(ctx->ev_copy[half], ctx->copy_stream));
                ctx->stats.d2h_bytes += (uint64_t)n * frame_bytes;
            }
            CK(cudaStreamSynchronize(ctx->stream));
            CK(cudaStreamSynchronize(ctx->copy_stream));
            if (ctx->profiling) drain_events(ctx);
            bool retry = false;
            if ((st = check_group(ctx, ctx->h_counters, n_frames, &retry)) != RXC_OK) return st;
            if (!retry) { ctx->lists_sized = true; return RXC_OK; }  // else: a tile-list arena overflowed somewhere; it has been grown, run the call again
        }
        return fail(ctx, RXC_ERR_OOM, "tile-list arena kept overflowing");
    }

    for (uint32_t first = 0; first < n_frames; first += group) {
        const uint32_t n = std::min(group, n_frames - first);
        for (int attempt = 0; attempt < 4; ++attempt) {
            if ((st = ensure_workspace(ctx, n, tiles_per_frame)) != RXC_OK) return st;
            // the pinned frame and counter blocks are reused per group and per call: whatever was enqueued before (an
            // earlier asynchronous call included) must have consumed / filled them, and an asynchronous group's
            // counters are looked at before the next group overwrites them
            CK(cudaStreamSynchronize(ctx->stream));
            if ((st = check_async(ctx)) != RXC_OK) return st;
            for (uint32_t i = 0; i < n; ++i)
                if ((st = fill_frame(ctx, frames[first + i], &ctx->h_frames[i])) != RXC_OK) return st;
            uint8_t* d_px = pixels + (uint64_t)first * stride;
            uint64_t d_stride = stride;
            if (!dev_px) {
                if ((st = reserve(ctx, ctx->d_out_px, (size_t)n * frame_bytes)) != RXC_OK) return st;
                d_px = ctx->d_out_px.as<uint8_t>(); d_stride = frame_bytes;
            }
            uint32_t* d_ow = owner; float* d_dp = depth;
            if (owner && !dev_owner) { if ((st = reserve(ctx, ctx->d_out_owner, frame_bytes)) != RXC_OK) return st; d_ow = ctx->d_out_owner.as<uint32_t>(); }
            if (depth && !dev_depth) { if ((st = reserve(ctx, ctx->d_out_depth, frame_bytes)) != RXC_OK) return st; d_dp = ctx->d_out_depth.as<float>(); }
            if ((st = launch_group(ctx, ctx->h_frames, ctx->h_counters, n, d_px, d_stride, pitch_px, d_ow, d_dp, 1u, [](uint32_t, uint32_t) -> int32_t { return RXC_OK; })) != RXC_OK) return st;
            if (!sync && dev_px && ctx->lists_sized) { ctx->async_pending = n; break; }  // truly asynchronous: the counters are checked before they are reused, or at the next synchronize
            CK(cudaStreamSynchronize(ctx->stream));
            if (ctx->profiling) drain_events(ctx);
            bool retry = false;
            if ((st = check_group(ctx, ctx->h_counters, n, &retry)) != RXC_OK) return st;
            if (retry) continue;
            ctx->lists_sized = true;
            if (!dev_px) {
                if (stride == frame_bytes || n == 1) {
                    CK(cudaMemcpyAsync(pixels + (uint64_t)first * stride, d_px, (size_t)n * frame_bytes, cudaMemcpyDeviceToHost, ctx->stream));
                } else {
                    CK(cudaMemcpy2DAsync(pixels + (uint64_t)first * stride, stride, d_px, frame_bytes, frame_bytes, n, cudaMemcpyDeviceToHost, ctx->stream));
                }
                ctx->stats.d2h_bytes += (uint64_t)n * frame_bytes;
            }
            if (owner && !dev_owner) { CK(cudaMemcpyAsync(owner, d_ow, frame_bytes, cudaMemcpyDeviceToHost, ctx->stream)); ctx->stats.d2h_bytes += frame_bytes; }
            if (depth && !dev_depth) { CK(cudaMemcpyAsync(depth, d_dp, frame_bytes, cudaMemcpyDeviceToHost, ctx->stream)); ctx->stats.d2h_bytes += frame_bytes; }
            CK(cudaStreamSynchronize(ctx->stream));
            break;
        }
    }
    return RXC_OK;
}

// C++ exceptions (std::bad_alloc from the host-side staging vectors, std::length_error ...) must not cross the C ABI:
// a Rust or C host cannot catch them.  Every entry point runs inside this barrier and reports them as a status.
template <class F>
int32_t guarded(rxc_ctx* ctx, F&& f) noexcept {
    try {
        return f();
    } catch (const std::bad_alloc&) {
        if (ctx) { try { ctx->err = "out of host memory"; } catch (...) {} }
        return RXC_ERR_OOM;
    } catch (const std::exception& e) {
        if (ctx) { try { ctx->err = std::string("internal error: ") + e.what(); } catch (...) {} }
        return RXC_ERR_INVALID;
    } catch (...) {
        if (ctx) { try { ctx->err = "internal error"; } catch (...) {} }
        return RXC_ERR_INVALID;
    }
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// internal interface (rx_internal.h)
// ---------------------------------------------------------------------------------------------
cudaStream_t rxi_stream(rxc_ctx* ctx) { return ctx->stream; }
int rxi_device(rxc_ctx* ctx) { return ctx->device; }
int32_t rxi_fail(rxc_ctx* ctx, int32_t code, const std::string& msg) { return fail(ctx, code, msg); }
RxMgpu** rxi_mgpu_slot(rxc_ctx* ctx) { return &ctx->mgpu; }
void rxi_count_launch(rxc_ctx* ctx, uint32_t n) { ctx->stats.kernel_launches += n; }
int32_t rxi_rasterize_device(rxc_ctx* ctx, const rxc_frame* frames, uint32_t n_frames, uint8_t* d_pixels, uint64_t stride, uint64_t pitch_bytes) {
    return rasterize_impl(ctx, frames, n_frames, d_pixels, stride, nullptr, nullptr, false, pitch_bytes);
}

// ---------------------------------------------------------------------------------------------
// exported C ABI
// ---------------------------------------------------------------------------------------------
extern "C" {

uint32_t rxc_abi_version(void) { return RXC_ABI_VERSION; }

const char* rxc_kernel_name(uint32_t k) { return k < RXC_N_KERNELS ? kKernelNames[k] : ""; }

int32_t rxc_create(int32_t device, rxc_ctx** out) {
    return guarded(nullptr, [&]() -> int32_t {
    if (!out) return RXC_ERR_INVALID;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0 || device < 0 || device >= n) { cudaGetLastError(); return RXC_ERR_NO_DEVICE; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return RXC_ERR_CUDA;
    if (prop.major != 10) return RXC_ERR_NO_DEVICE;  // the library holds sm_100a code only
    if (cudaSetDevice(device) != cudaSuccess) return RXC_ERR_CUDA;
    rxc_ctx* ctx = new rxc_ctx();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return RXC_ERR_CUDA; }
    ctx->stream = ctx->own_stream;
    ctx->raster_blocks_per_sm = rxk_raster_blocks_per_sm();
    if (const char* e = getenv("RXC_RASTER_BLOCKS")) ctx->raster_blocks_per_sm = std::max(1, atoi(e));   // experiments (with RXC_JIT_EXTRA=-DRX_RASTER_MIN_BLOCKS=n)
    if (const char* e = getenv("RXC_PIECE_MB")) ctx->piece_mb = std::max(1, atoi(e));
    if (const char* e = getenv("RXC_SLICE_MB")) ctx->slice_mb = std::max(0, atoi(e));
    if (const char* e = getenv("RXC_TMA_STORE")) ctx->tma_store = atoi(e);
    if (const char* e = getenv("RXC_VM_JIT")) ctx->vm_jit = std::min(2, std::max(0, atoi(e)));
    if (const char* e = getenv("RXC_KERNEL_SPEC")) ctx->kernel_spec = atoi(e) != 0;
    if (const char* e = getenv("RXC_VM_STATE_MODE")) ctx->vm_state_mode = std::min(2, std::max(0, atoi(e)));
    if (const char* e = getenv("RXC_FRONT_STOP")) ctx->front_stop = std::max(0, atoi(e));
    if (const char* e = getenv("RXC_SMALL_MIN_LIST")) ctx->small_min_list = atoi(e);   // 0 = pass off
    if (const char* e = getenv("RXC_RASTER_GROUPS")) { ctx->raster_groups = atoi(e); ctx->raster_groups_forced = true; }   // 0 / 1 = one work counter over all frames of a launch; > 1: groups whatever the frame size
    if (const char* e = getenv("RXC_SMALL_GSHIFT")) ctx->small_gshift = std::min(5, std::max(0, atoi(e)));
    if (const char* e = getenv("RXC_SMALL_MIN_TRIS")) ctx->small_min_tris = atoi(e);
    if (const char* e = getenv("RXC_SMALL_MAX_PIX")) ctx->small_max_pix = std::min(1024, std::max(1, atoi(e)));
    if (const char* e = getenv("RXC_FRONT_CLUSTER_MAX")) ctx->front_cluster_max = atoi(e);  // tuning knob for experiments
    *out = ctx;
    return RXC_OK;
    });
}

void rxc_destroy(rxc_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    rxi_mgpu_destroy(ctx);
    if (ctx->jit) { rxj_destroy(ctx->jit); ctx->jit = nullptr; }
    DevBuf* bufs[] = {&ctx->d_ordered, &ctx->d_arena, &ctx->d_tex, &ctx->d_tiles, &ctx->d_pos, &ctx->d_uv, &ctx->d_nrm, &ctx->d_idx, &ctx->d_b3,
                      &ctx->d_chunks, &ctx->d_orphans, &ctx->d_pos2, &ctx->d_uv2, &ctx->d_idx2, &ctx->d_b2, &ctx->d_lights,
                      &ctx->w_frames, &ctx->w_fb, &ctx->w_fb2, &ctx->w_lights, &ctx->w_counters, &ctx->w_vis, &ctx->w_shade,
                      &ctx->w_bins, &ctx->w_ctot, &ctx->w_cbase, &ctx->w_clip, &ctx->w_large, &ctx->w_tcount, &ctx->w_tbase,
                      &ctx->w_tfill, &ctx->w_lists, &ctx->w_tri2d, &ctx->w_rcounter, &ctx->d_out_px, &ctx->d_out_owner, &ctx->d_out_depth,
                      &ctx->w_tcount2, &ctx->w_tbase2, &ctx->w_tfill2, &ctx->w_lists2, &ctx->d_sectors, &ctx->d_chunkinfo, &ctx->d_linedefs,
                      &ctx->d_vm_code, &ctx->d_vm_programs, &ctx->d_vm_patdata, &ctx->d_vm_patterns, &ctx->d_vm_palette,
                      &ctx->d_pj_pv, &ctx->d_pj_uv, &ctx->d_pj_nrm, &ctx->d_pj_idx, &ctx->d_pj_edges, &ctx->d_pj_info, &ctx->d_pj_bbox};
    for (DevBuf* b : bufs) free_buf(*b);
    if (ctx->h_frames) cudaFreeHost(ctx->h_frames);
    if (ctx->h_counters) cudaFreeHost(ctx->h_counters);
    for (auto& p : ctx->pending) { cudaEventDestroy(p.a); cudaEventDestroy(p.b); }
    for (auto e : ctx->free_events) cudaEventDestroy(e);
    if (ctx->copy_stream) {
        cudaStreamSynchronize(ctx->copy_stream);
        for (int i = 0; i < 2; ++i) { cudaEventDestroy(ctx->ev_render[i]); cudaEventDestroy(ctx->ev_copy[i]); }
        for (int i = 0; i < 16; ++i) if (ctx->ev_slice[i]) cudaEventDestroy(ctx->ev_slice[i]);
        cudaStreamDestroy(ctx->copy_stream);
    }
    cudaStreamDestroy(ctx->own_stream);
    delete ctx;
}

const char* rxc_last_error(const rxc_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int32_t rxc_set_stream(rxc_ctx* ctx, void* cuda_stream) {
    return guarded(ctx, [&]() -> int32_t {
    if (!ctx) return RXC_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
    return RXC_OK;
    });
}

int32_t rxc_set_assets(rxc_ctx* ctx, const rxc_tile* tiles, uint32_t n_tiles) {
    return guarded(ctx, [&]() -> int32_t {
    if (!ctx || (n_tiles && !tiles)) return RXC_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    int32_t st = build_tiles(ctx, tiles, n_tiles, ctx->h_static_arena, ctx->h_static_tex, ctx->h_static_tiles);
    if (st != RXC_OK) return st;
    ctx->textures_dirty = true;
    ctx->spec_dirty = true;
    return upload_textures(ctx);
    });
}

int32_t rxc_set_lights(rxc_ctx* ctx, const rxc_light* lights, uint32_t n_lights) {
    return guarded(ctx, [&]() -> int32_t {
    if (!ctx || (n_lights && !lights)) return RXC_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    const uint32_t before = ctx->S.n_lights;
    int32_t st = upload_lights(ctx, lights, n_lights);
    if (st != RXC_OK) return st;
    if (n_lights > before) ctx->ws_frames = 0;  // per-frame light slices must grow
    return RXC_OK;
    });
}

int32_t rxc_set_mapmini(rxc_ctx* ctx, const rxc_mapmini* mm) {
    return guarded(ctx, [&]() -> int32_t {
    if (!ctx) return RXC_ERR_INVALID;
    if (mm && ((mm->n_linedefs && !mm->linedefs) || (mm->n_occluded_sectors && !mm->occluded_sectors)))
        return fail(ctx, RXC_ERR_INVALID, "null array with non-zero count in rxc_mapmini");
    CK(cudaSetDevice(ctx->device));
    ctx->h_linedefs.clear(); ctx->h_mm_sectors.clear();
    if (mm) {
        for (uint32_t i = 0; i < mm->n_linedefs; ++i) {
            const rxc_linedef& l = mm->linedefs[i];
            ctx->h_linedefs.insert(ctx->h_linedefs.end(), {l.start[0], l.start[1], l.end[0], l.end[1]});
        }
        for (uint32_t i = 0; i < mm->n_occluded_sectors; ++i) {
            const rxc_sector& q = mm->occluded_sectors[i];
            ctx->h_mm_sectors.push_back(DSector{q.min[0], q.min[1], q.max[0], q.max[1], q.occlusion, {0, 0, 0}});
        }
    }
    return upload_chunks(ctx, (uint32_t)ctx->h_static_tex.size());
    });
}

namespace {
// rxc_set_scene (keep3d = 0) and rxc_update_scene: the first keep3d 3D batches have the geometry (vertex / uv / normal / index arrays)
// of the previous call -- their flattened copies, bounding boxes, orphan lists and setup chunks are reused and their pointers not read
int32_t set_scene_impl(rxc_ctx* ctx, const rxc_scene* sc, uint32_t keep3d) {
    if (!ctx || !sc) return RXC_ERR_INVALID;
    if ((sc->n_batches3d && !sc->batches3d) || (sc->n_batches2d && !sc->batches2d) || (sc->n_lights && !sc->lights) ||
        (sc->n_dynamic_textures && !sc->dynamic_textures) || (sc->n_chunks && !sc->chunks) || (sc->n_actor_tiles && !sc->actor_tiles) ||
        (sc->n_shaders && !sc->shaders) || (sc->n_patterns && !sc->patterns) || (sc->n_patterns_normal && !sc->patterns_normal) ||
        (sc->n_palette && !sc->palette) || sc->n_scene_shaders > sc->n_shaders)
        return fail(ctx, RXC_ERR_INVALID, "null array with non-zero count in rxc_scene");
    CK(cudaSetDevice(ctx->device));
    const std::vector<DBatch3> old_b3 = ctx->h_b3;   // geometry-derived fields of the batches that are kept
    if (keep3d) {
        if (!ctx->have_scene || !ctx->h_geometry_valid || keep3d > old_b3.size() || keep3d > sc->n_batches3d)
            return fail(ctx, RXC_ERR_INVALID, "rxc_update_scene: keep_batches3d exceeds the batches of the resident scene (or there is none)");
        for (uint32_t i = 0; i < keep3d; ++i)
            if (old_b3[i].n_verts != sc->batches3d[i].n_vertices || old_b3[i].n_tris != sc->batches3d[i].n_triangles ||
                (old_b3[i].has_normals != 0u) != (sc->batches3d[i].normals != nullptr))
                return fail(ctx, RXC_ERR_INVALID, "rxc_update_scene: 3D batch " + std::to_string(i) + " is not the batch of the resident scene (vertex / triangle count, normals)");
    }
    ctx->have_scene = false;
    ctx->h_geometry_valid = false;
    for (uint32_t i = 0; i < sc->n_chunks; ++i)
        if ((uint64_t)sc->chunks[i].shader_base + sc->chunks[i].n_shaders > sc->n_shaders)
            return fail(ctx, RXC_ERR_INDEX, "chunk " + std::to_string(i) + ": shader range outside rxc_scene.shaders");
    bool any_shader = false;

    // ---- validate + size
    size_t V = 0, T = 0, V2 = 0, T2 = 0;
    for (uint32_t i = 0; i < sc->n_batches3d; ++i) {
        const rxc_batch3d& b = sc->batches3d[i];
        const std::string who = "3D batch " + std::to_string(i) + ": ";
        if (b.source_kind > RXC_SRC_TERRAIN) return fail(ctx, RXC_ERR_INVALID, who + "bad source kind");
        if (b.pass > RXC_PASS_CHUNK_OPACITY) return fail(ctx, RXC_ERR_INVALID, who + "bad pass");
        if (b.chunk >= (int32_t)sc->n_chunks) return fail(ctx, RXC_ERR_INDEX, who + "chunk index out of range");
        if (b.source_kind == RXC_SRC_TERRAIN && b.chunk >= 0 && sc->chunks[b.chunk].terrain_texture && sc->chunks[b.chunk].size == 0)
            return fail(ctx, RXC_ERR_INDEX, who + "terrain chunk of size 0 (reference divides by chunk.size)");
        if (b.index_bytes != 4 && b.index_bytes != 8) return fail(ctx, RXC_ERR_INVALID, who + "index_bytes must be 4 or 8");
        if (b.cull_mode > RXC_CULL_BACK || b.repeat_mode > RXC_REPEAT_REPEAT_Y) return fail(ctx, RXC_ERR_INVALID, who + "bad enum value");
        if (i >= keep3d) {   // (a kept batch was checked when it was uploaded; its arrays are not read again)
            if ((b.n_vertices && (!b.vertices || !b.uvs)) || (b.n_triangles && !b.indices)) return fail(ctx, RXC_ERR_INVALID, who + "null geometry pointer");
            for (size_t k = 0; k < (size_t)b.n_triangles * 3; ++k)
                if (idx_at(b.indices, b.index_bytes, k) >= b.n_vertices) return fail(ctx, RXC_ERR_INDEX, who + "vertex index out of range (reference panics)");
        }
        V += b.n_vertices; T += b.n_triangles;
    }
    for (uint32_t i = 0; i < sc->n_batches2d; ++i) {
        const rxc_batch2d& b = sc->batches2d[i];
        const std::string who = "2D batch " + std::to_string(i) + ": ";
        if (b.source_kind > RXC_SRC_TERRAIN) return fail(ctx, RXC_ERR_INVALID, who + "bad source kind");
        if (b.mode > RXC_MODE_LINE_LOOP) return fail(ctx, RXC_ERR_INVALID, who + "bad primitive mode");
        if (b.chunk >= (int32_t)sc->n_chunks) return fail(ctx, RXC_ERR_INDEX, who + "chunk index out of range");
        if (b.source_kind == RXC_SRC_TERRAIN && b.chunk >= 0 && sc->chunks[b.chunk].terrain_texture && sc->chunks[b.chunk].size == 0)
            return fail(ctx, RXC_ERR_INDEX, who + "terrain chunk of size 0 (reference divides by chunk.size)");
        if (b.index_bytes != 4 && b.index_bytes != 8) return fail(ctx, RXC_ERR_INVALID, who + "index_bytes must be 4 or 8");
        if (b.repeat_mode > RXC_REPEAT_REPEAT_Y) return fail(ctx, RXC_ERR_INVALID, who + "bad enum value");
        if ((b.n_vertices && (!b.vertices || !b.uvs)) || (b.n_triangles && !b.indices)) return fail(ctx, RXC_ERR_INVALID, who + "null geometry pointer");
        if (b.mode == RXC_MODE_LINE_STRIP && b.n_vertices == 0) return fail(ctx, RXC_ERR_INDEX, who + "LineStrip without vertices (reference underflows)");
        // every index triple is read by Batch2D::project (batch2d.rs:405-423) whatever the mode
        for (size_t k = 0; k < (size_t)b.n_triangles * 3; ++k)
            if (idx_at(b.indices, b.index_bytes, k) >= b.n_vertices) return fail(ctx, RXC_ERR_INDEX, who + "vertex index out of range (reference panics)");
        V2 += b.n_vertices; T2 += b.n_triangles;
    }
    if (sc->n_batches3d >= RX_META_BATCH) return fail(ctx, RXC_ERR_UNSUPPORTED, "too many 3D batches");
    if (3 * T >= 0x7FFFFFFFull || V >= 0xFFFFFFFFull) return fail(ctx, RXC_ERR_UNSUPPORTED, "scene too large for 32-bit slots");

    // ---- flatten 3D
    // (the vectors live in the context: what the kept batches put there stays, the rest is rewritten)
    size_t keep_v = 0, keep_t = 0, keep_orphans = 0, keep_chunks = 0;
    if (keep3d) {
        const DBatch3& last = old_b3[keep3d - 1];
        keep_v = (size_t)last.v_off + last.n_verts; keep_t = (size_t)last.t_off + last.n_tris;
        keep_orphans = (size_t)last.orphan_off + last.n_orphans; keep_chunks = (size_t)last.chunk_first + last.n_chunks;
    }
    std::vector<float>&pos = ctx->h_pos, &uv = ctx->h_uv, &nrm = ctx->h_nrm;
    std::vector<uint32_t>&idx = ctx->h_idx, &orphans = ctx->h_orphans;
    std::vector<DChunk>& chunks = ctx->h_setup_chunks;
    pos.resize(V * 4); uv.resize(V * 2); nrm.resize(V * 3); idx.resize(T * 3);
    std::fill(nrm.begin() + keep_v * 3, nrm.end(), 0.0f);   // batches without normals read zeros
    orphans.resize(keep_orphans); chunks.resize(keep_chunks);
    ctx->h_b3.assign(sc->n_batches3d, DBatch3{});
    ctx->owner_base.assign(sc->n_batches3d, 0);
    size_t vo = 0, to = 0;
    bool any_opacity = false;
    std::vector<std::pair<uint32_t, const rxc_texture*>> baked;   // (3D batch, baked shader texture)
    for (uint32_t i = 0; i < sc->n_batches3d; ++i) {
        const rxc_batch3d& b = sc->batches3d[i];
        DBatch3& d = ctx->h_b3[i];
        d.v_off = (uint32_t)vo; d.n_verts = b.n_vertices; d.t_off = (uint32_t)to; d.n_tris = b.n_triangles;
        d.owner_base = (uint32_t)(3 * to);
        ctx->owner_base[i] = d.owner_base;
        d.cull_mode = b.cull_mode; d.repeat_mode = b.repeat_mode; d.source_kind = b.source_kind; d.source_index = b.source_index;
        memcpy(&d.source_pixel, b.source_pixel, 4);
        d.has_normals = b.normals ? 1u : 0u;
        d.chunk = b.chunk < 0 ? -1 : b.chunk;
        d.program = resolve_program(sc, b.shader, b.chunk);
        if (b.shader >= 0 && b.chunk >= 0 && b.pass != RXC_PASS_CHUNK_OPACITY) {
            // chunk.shader_textures[shader]: a baked texture replaces the texel and the program does not run (rasterizer.rs:1227-1262)
            const rxc_chunk& c = sc->chunks[b.chunk];
            if (c.shader_textures && (uint32_t)b.shader < c.n_shaders && c.shader_textures[b.shader]) {
                baked.push_back({i, c.shader_textures[b.shader]});
                d.program = -1;
            }
        }
        if (d.program >= 0 && sc->shaders[d.program].n_words) any_shader = true;
        d.profile_id = b.profile_id;
        d.bflags = (b.has_profile_id ? RX_BF_HAS_PROFILE : 0u) | (b.pass == RXC_PASS_CHUNK_OPACITY ? RX_BF_OPACITY : 0u);
        if (b.pass == RXC_PASS_CHUNK_OPACITY) any_opacity = true;
        memcpy(d.ambient, b.ambient_color, 12);
        memcpy(d.transform, b.transform, 64);
        if (i < keep3d) {   // geometry of the previous call: flattened arrays, bounding box, orphans and setup chunks stand
            const DBatch3& o = old_b3[i];
            d.bflags |= o.bflags & RX_BF_UNIT_W;
            memcpy(d.aabb_min, o.aabb_min, 12); memcpy(d.aabb_max, o.aabb_max, 12);
            d.orphan_off = o.orphan_off; d.n_orphans = o.n_orphans; d.chunk_first = o.chunk_first; d.n_chunks = o.n_chunks;
            vo += b.n_vertices; to += b.n_triangles;
            continue;
        }
        if (b.n_vertices) {
            memcpy(&pos[vo * 4], b.vertices, (size_t)b.n_vertices * 16);
            memcpy(&uv[vo * 2], b.uvs, (size_t)b.n_vertices * 8);
            if (b.normals) memcpy(&nrm[vo * 3], b.normals, (size_t)b.n_vertices * 12);
        }
        float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
        bool unit_w = b.n_vertices != 0;
        for (uint32_t v = 0; v < b.n_vertices; ++v) {
            for (int k = 0; k < 3; ++k) { mn[k] = std::fmin(mn[k], b.vertices[v * 4 + k]); mx[k] = std::fmax(mx[k], b.vertices[v * 4 + k]); }
            unit_w = unit_w && b.vertices[v * 4 + 3] == 1.0f && std::isfinite(b.vertices[v * 4]) && std::isfinite(b.vertices[v * 4 + 1]) && std::isfinite(b.vertices[v * 4 + 2]);
        }
        if (unit_w) d.bflags |= RX_BF_UNIT_W;   // the band reject of whole batches (k_frame_setup) relies on it
        memcpy(d.aabb_min, mn, 12); memcpy(d.aabb_max, mx, 12);
        std::vector<uint8_t> used(b.n_vertices, 0);
        for (size_t k = 0; k < (size_t)b.n_triangles * 3; ++k) {
            size_t v = idx_at(b.indices, b.index_bytes, k);
            used[v] = 1;
            idx[to * 3 + k] = (uint32_t)(vo + v);
        }
        d.orphan_off = (uint32_t)orphans.size();
        if (b.n_triangles)
            for (uint32_t v = 0; v < b.n_vertices; ++v) if (!used[v]) orphans.push_back((uint32_t)(vo + v));
        d.n_orphans = (uint32_t)orphans.size() - d.orphan_off;
        d.chunk_first = (uint32_t)chunks.size();
        for (uint32_t t = 0; t < b.n_triangles; t += RX_CHUNK_TRIS)
            chunks.push_back(DChunk{i, (uint32_t)to + t, std::min<uint32_t>(RX_CHUNK_TRIS, b.n_triangles - t), 0});
        d.n_chunks = (uint32_t)chunks.size() - d.chunk_first;
        vo += b.n_vertices; to += b.n_triangles;
    }
    // ---- flatten 2D
    std::vector<float> pos2(V2 * 2), uv2(V2 * 2);
    std::vector<uint32_t> idx2(T2 * 3);
    ctx->h_b2.assign(sc->n_batches2d, DBatch2{});
    size_t vo2 = 0, to2 = 0, ro2 = 0;
    for (uint32_t i = 0; i < sc->n_batches2d; ++i) {
        const rxc_batch2d& b = sc->batches2d[i];
        DBatch2& d = ctx->h_b2[i];
        d.v_off = (uint32_t)vo2; d.n_verts = b.n_vertices; d.t_off = (uint32_t)to2; d.n_tris = b.n_triangles;
        d.mode = b.mode; d.repeat_mode = b.repeat_mode; d.source_kind = b.source_kind; d.source_index = b.source_index;
        memcpy(&d.source_pixel, b.source_pixel, 4);
        d.receives_light = b.receives_light ? 1u : 0u;
        d.chunk = b.chunk < 0 ? -1 : b.chunk;
        d.program = resolve_program(sc, b.shader, b.chunk);
        if (d.program >= 0 && sc->shaders[d.program].n_words) any_shader = true;
        // records per frame (rasterizer.rs:604, :901-955): triangles, index pairs, or consecutive vertices
        d.n_recs = b.mode == RXC_MODE_LINE_STRIP ? b.n_vertices - 1 : b.mode == RXC_MODE_LINE_LOOP ? b.n_vertices : b.n_triangles;
        d.rec_off = (uint32_t)ro2;
        ro2 += d.n_recs;
        if (b.n_vertices) { memcpy(&pos2[vo2 * 2], b.vertices, (size_t)b.n_vertices * 8); memcpy(&uv2[vo2 * 2], b.uvs, (size_t)b.n_vertices * 8); }
        for (size_t k = 0; k < (size_t)b.n_triangles * 3; ++k) idx2[to2 * 3 + k] = (uint32_t)(vo2 + idx_at(b.indices, b.index_bytes, k));
        vo2 += b.n_vertices; to2 += b.n_triangles;
    }

    int32_t st;
    {   // scene textures: dynamic tiles, host-resolved entity/item tiles, then the chunks' terrain textures
        std::vector<rxc_tile> all(sc->dynamic_textures, sc->dynamic_textures + sc->n_dynamic_textures);
        all.insert(all.end(), sc->actor_tiles, sc->actor_tiles + sc->n_actor_tiles);
        for (size_t k = 0; k < baked.size(); ++k) {  // baked shader textures ride along as one-frame actor tiles
            all.push_back(rxc_tile{baked[k].second, 1u});
            DBatch3& d = ctx->h_b3[baked[k].first];
            d.source_kind = RXC_SRC_ENTITY_TILE; d.source_index = sc->n_actor_tiles + (uint32_t)k;
        }
        if ((st = build_tiles(ctx, all.data(), (uint32_t)all.size(), ctx->h_dyn_arena, ctx->h_dyn_tex, ctx->h_dyn_tiles)) != RXC_OK) return st;
        ctx->n_dyn_tiles = sc->n_dynamic_textures; ctx->n_actor_tiles = sc->n_actor_tiles + (uint32_t)baked.size();
        ctx->h_chunks.clear(); ctx->h_sectors.clear();
        for (uint32_t i = 0; i < sc->n_chunks; ++i) {
            const rxc_chunk& c = sc->chunks[i];
            if (c.n_occluded_sectors && !c.occluded_sectors) return fail(ctx, RXC_ERR_INVALID, "chunk with null occluded_sectors");
            HChunk h = {};
            h.origin[0] = c.origin[0]; h.origin[1] = c.origin[1]; h.size = c.size;
            h.sector_off = (uint32_t)ctx->h_sectors.size(); h.n_sectors = c.n_occluded_sectors;
            for (uint32_t k = 0; k < c.n_occluded_sectors; ++k) {
                const rxc_sector& q = c.occluded_sectors[k];
                ctx->h_sectors.push_back(DSector{q.min[0], q.min[1], q.max[0], q.max[1], q.occlusion, {0, 0, 0}});
            }
            h.terrain_tex = -1;
            if (c.terrain_texture) {
                const rxc_texture& x = *c.terrain_texture;
                if (!x.data || x.width == 0 || x.height == 0) return fail(ctx, RXC_ERR_INVALID, "empty terrain texture");
                if (x.width > 65535u || x.height > 65535u) return fail(ctx, RXC_ERR_UNSUPPORTED, "textures larger than 65535 texels per side");
                DTex d; d.offset = ctx->h_dyn_arena.size(); d.width = x.width; d.height = x.height; d.pad = 0;
                const size_t bytes = (size_t)x.width * x.height * 4;
                bool opaque = true;
                for (size_t k = 3; k < bytes; k += 4) if (x.data[k] != 255) { opaque = false; break; }
                d.all_opaque = opaque ? 1u : 0u;
                ctx->h_dyn_arena.insert(ctx->h_dyn_arena.end(), x.data, x.data + bytes);
                ctx->h_dyn_arena.resize((ctx->h_dyn_arena.size() + 255) & ~(size_t)255);
                h.terrain_tex = (int32_t)ctx->h_dyn_tex.size();
                ctx->h_dyn_tex.push_back(d);
            }
            ctx->h_chunks.push_back(h);
        }
    }
    ctx->textures_dirty = true;
    if ((st = upload_textures(ctx)) != RXC_OK) return st;
    if ((st = upload_from(ctx, ctx->d_pos, pos.data(), pos.size() * 4, keep_v * 16)) != RXC_OK) return st;
    if ((st = upload_from(ctx, ctx->d_uv, uv.data(), uv.size() * 4, keep_v * 8)) != RXC_OK) return st;
    if ((st = upload_from(ctx, ctx->d_nrm, nrm.data(), nrm.size() * 4, keep_v * 12)) != RXC_OK) return st;
    if ((st = upload_from(ctx, ctx->d_idx, idx.data(), idx.size() * 4, keep_t * 12)) != RXC_OK) return st;
    if ((st = upload(ctx, ctx->d_b3, ctx->h_b3.data(), ctx->h_b3.size() * sizeof(DBatch3))) != RXC_OK) return st;
    if ((st = upload_from(ctx, ctx->d_chunks, chunks.data(), chunks.size() * sizeof(DChunk), keep_chunks * sizeof(DChunk))) != RXC_OK) return st;
    if ((st = upload_from(ctx, ctx->d_orphans, orphans.data(), orphans.size() * 4, keep_orphans * 4)) != RXC_OK) return st;
    if ((st = upload(ctx, ctx->d_pos2, pos2.data(), pos2.size() * 4)) != RXC_OK) return st;
    if ((st = upload(ctx, ctx->d_uv2, uv2.data(), uv2.size() * 4)) != RXC_OK) return st;
    if ((st = upload(ctx, ctx->d_idx2, idx2.data(), idx2.size() * 4)) != RXC_OK) return st;
    if ((st = upload(ctx, ctx->d_b2, ctx->h_b2.data(), ctx->h_b2.size() * sizeof(DBatch2))) != RXC_OK) return st;
    if ((st = upload_lights(ctx, sc->lights, sc->n_lights)) != RXC_OK) return st;
    if ((st = upload_vm(ctx, sc)) != RXC_OK) return st;
    {   // which programs could tell the reference's per-tile Execution from the device's fresh one
        std::vector<uint8_t> usage(sc->n_shaders, 0);
        for (const DBatch3& b : ctx->h_b3) if (b.program >= 0 && (uint32_t)b.program < sc->n_shaders) usage[b.program] |= 1u;
        for (const DBatch2& b : ctx->h_b2) if (b.program >= 0 && (uint32_t)b.program < sc->n_shaders) usage[b.program] |= 2u;
        ctx->vm_state_report.assign(sc->n_shaders, 0u);
        if (sc->n_shaders) rxj_state_report(sc->shaders, sc->n_shaders, usage.data(), !ctx->h_b3.empty(), ctx->vm_state_report.data());
        for (uint32_t i = 0; i < sc->n_shaders; ++i) if (!usage[i]) ctx->vm_state_report[i] = 0u;   // bound to no batch: never runs
    }

    SceneDev& S = ctx->S;
    S.pos = ctx->d_pos.as<float4>(); S.uv = ctx->d_uv.as<float2>(); S.nrm = ctx->d_nrm.as<float>(); S.idx = ctx->d_idx.as<uint32_t>();
    S.b3 = ctx->d_b3.as<DBatch3>(); S.chunks = ctx->d_chunks.as<DChunk>(); S.orphans = ctx->d_orphans.as<uint32_t>();
    S.pos2 = ctx->d_pos2.as<float2>(); S.uv2 = ctx->d_uv2.as<float2>(); S.idx2 = ctx->d_idx2.as<uint32_t>(); S.b2 = ctx->d_b2.as<DBatch2>();
    S.n_b3 = sc->n_batches3d; S.n_b2 = sc->n_batches2d; S.n_chunks = (uint32_t)chunks.size();
    S.n_tris = (uint32_t)T; S.n_verts = (uint32_t)V; S.n_rec2d = (uint32_t)ro2;
    // ordered 3D lists when the opacity layer is in play, binned 2D lists when a warp ballot cannot hold the records
    // ... and whenever a batch carries a VM program (the interpreter is only compiled into the general kernels)
    S.general = (any_opacity || ro2 > 32 || any_shader) ? 1u : 0u;
    ctx->ws_frames = 0;  // workspace strides depend on the scene
    ctx->list_cap_min = 0;
    ctx->list2_cap_min = 0;
    ctx->lists_sized = false;
    ctx->have_scene = true;
    ctx->h_geometry_valid = true;
    return RXC_OK;
}
}  // namespace

int32_t rxc_set_scene(rxc_ctx* ctx, const rxc_scene* sc) {
    return guarded(ctx, [&]() -> int32_t { return set_scene_impl(ctx, sc, 0u); });
}

int32_t rxc_update_scene(rxc_ctx* ctx, const rxc_scene* sc, uint32_t keep_batches3d) {
    return guarded(ctx, [&]() -> int32_t { return set_scene_impl(ctx, sc, keep_batches3d); });
}

int32_t rxc_rasterize(rxc_ctx* ctx, const rxc_frame* frame, uint8_t* pixels, uint32_t* owner, float* depth) {
    return guarded(ctx, [&]() -> int32_t {
    return rasterize_impl(ctx, frame, 1, pixels, 0, owner, depth, true);
    });
}
int32_t rxc_rasterize_async(rxc_ctx* ctx, const rxc_frame* frame, uint8_t* pixels, uint32_t* owner, float* depth) {
    return guarded(ctx, [&]() -> int32_t {
    return rasterize_impl(ctx, frame, 1, pixels, 0, owner, depth, false);
    });
}
int32_t rxc_rasterize_batch(rxc_ctx* ctx, const rxc_frame* frames, uint32_t n_frames, uint8_t* pixels, uint64_t frame_stride_bytes) {
    return guarded(ctx, [&]() -> int32_t {
    return rasterize_impl(ctx, frames, n_frames, pixels, frame_stride_bytes, nullptr, nullptr, true);
    });
}
int32_t rxc_rasterize_batch_async(rxc_ctx* ctx, const rxc_frame* frames, uint32_t n_frames, uint8_t* pixels, uint64_t frame_stride_bytes) {
    return guarded(ctx, [&]() -> int32_t {
    return rasterize_impl(ctx, frames, n_frames, pixels, frame_stride_bytes, nullptr, nullptr, false);
    });
}

int32_t rxc_rasterize_projected(rxc_ctx* ctx, const rxc_frame* frame, const rxc_projected3d* batches, uint32_t n_batches, uint8_t* pixels,
                                uint32_t* owner, float* depth) {
    return guarded(ctx, [&]() -> int32_t {
    if (!ctx) return RXC_ERR_INVALID;
    if (!frame || !pixels || (n_batches && !batches)) return fail(ctx, RXC_ERR_INVALID, "frame, batches and pixels are required");
    if (!ctx->have_scene) return fail(ctx, RXC_ERR_INVALID, "rxc_set_scene has not been called");
    if (n_batches != ctx->h_b3.size()) return fail(ctx, RXC_ERR_INVALID, "rxc_rasterize_projected: one rxc_projected3d per 3D batch of the scene");
    CK(cudaSetDevice(ctx->device));
    // ---- validate + flatten (host memory is only borrowed for the call)
    size_t NP = 0, NC = 0;
    for (uint32_t b = 0; b < n_batches; ++b) {
        const rxc_projected3d& q = batches[b];
        const std::string who = "projected batch " + std::to_string(b) + ": ";
        if (!q.has_bounding_box) continue;   // skipped altogether, whatever else it holds
        if (q.index_bytes != 4 && q.index_bytes != 8) return fail(ctx, RXC_ERR_INVALID, who + "index_bytes must be 4 or 8");
        if ((q.n_projected && (!q.projected_vertices || !q.clipped_uvs)) || (q.n_clipped && (!q.clipped_indices || !q.edges || !q.visible)))
            return fail(ctx, RXC_ERR_INVALID, who + "null array with non-zero count");
        if (q.n_clipped > 3ull * ctx->h_b3[b].n_tris) return fail(ctx, RXC_ERR_INVALID, who + "more clipped triangles than 3 per input triangle (batch3d.rs:625-681 emits at most 2 per clipped one)");
        if ((ctx->h_b3[b].has_normals != 0) != (q.clipped_normals != nullptr) && q.n_projected)
            return fail(ctx, RXC_ERR_INVALID, who + "clipped_normals must be given exactly when the batch has normals");
        for (size_t k = 0; k < (size_t)q.n_clipped * 3; ++k)
            if (idx_at(q.clipped_indices, q.index_bytes, k) >= q.n_projected) return fail(ctx, RXC_ERR_INDEX, who + "clipped index out of range (reference panics)");
        NP += q.n_projected; NC += q.n_clipped;
    }
    std::vector<float> pv(NP * 4), uv(NP * 2), nrm(NP * 3, 0.0f), edges(NC * 9), bbox((size_t)n_batches * 5, 0.0f);
    std::vector<uint32_t> idx(NC * 3), info(NC * 2);
    size_t po = 0, co = 0;
    for (uint32_t b = 0; b < n_batches; ++b) {
        const rxc_projected3d& q = batches[b];
        if (!q.has_bounding_box) continue;
        bbox[(size_t)b * 5] = 1.0f;
        memcpy(&bbox[(size_t)b * 5 + 1], q.bounding_box, 16);
        if (q.n_projected) {
            memcpy(&pv[po * 4], q.projected_vertices, (size_t)q.n_projected * 16);
            memcpy(&uv[po * 2], q.clipped_uvs, (size_t)q.n_projected * 8);
            if (q.clipped_normals) memcpy(&nrm[po * 3], q.clipped_normals, (size_t)q.n_projected * 12);
        }
        if (q.n_clipped) memcpy(&edges[co * 9], q.edges, (size_t)q.n_clipped * 36);
        for (uint32_t i = 0; i < q.n_clipped; ++i) {
            for (int k = 0; k < 3; ++k) idx[(co + i) * 3 + k] = (uint32_t)(po + idx_at(q.clipped_indices, q.index_bytes, (size_t)i * 3 + k));
            info[(co + i) * 2] = b | (q.visible[i] ? 0x80000000u : 0u);
            info[(co + i) * 2 + 1] = ctx->owner_base[b] + i;
        }
        po += q.n_projected; co += q.n_clipped;
    }
    int32_t st;
    if ((st = upload(ctx, ctx->d_pj_pv, pv.data(), pv.size() * 4)) != RXC_OK) return st;
    if ((st = upload(ctx, ctx->d_pj_uv, uv.data(), uv.size() * 4)) != RXC_OK) return st;
    if ((st = upload(ctx, ctx->d_pj_nrm, nrm.data(), nrm.size() * 4)) != RXC_OK) return st;
    if ((st = upload(ctx, ctx->d_pj_idx, idx.data(), idx.size() * 4)) != RXC_OK) return st;
    if ((st = upload(ctx, ctx->d_pj_edges, edges.data(), edges.size() * 4)) != RXC_OK) return st;
    if ((st = upload(ctx, ctx->d_pj_info, info.data(), info.size() * 4)) != RXC_OK) return st;
    if ((st = upload(ctx, ctx->d_pj_bbox, bbox.data(), bbox.size() * 4)) != RXC_OK) return st;
    ctx->pj.pv = ctx->d_pj_pv.as<float4>(); ctx->pj.uv = ctx->d_pj_uv.as<float2>(); ctx->pj.nrm = ctx->d_pj_nrm.as<float>();
    ctx->pj.idx = ctx->d_pj_idx.as<uint32_t>(); ctx->pj.edges = ctx->d_pj_edges.as<float>(); ctx->pj.info = ctx->d_pj_info.as<uint32_t>();
    ctx->pj.bbox = ctx->d_pj_bbox.as<float>(); ctx->pj.n_clipped = (uint32_t)NC;
    ctx->pj_active = true;
    st = rasterize_impl(ctx, frame, 1, pixels, 0, owner, depth, true);
    ctx->pj_active = false;
    return st;
    });
}

int32_t rxc_synchronize(rxc_ctx* ctx) {
    return guarded(ctx, [&]() -> int32_t {
    if (!ctx) return RXC_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    if (ctx->profiling) drain_events(ctx);
    return check_async(ctx);
    });
}

int32_t rxc_pin_host(rxc_ctx* ctx, void* ptr, uint64_t bytes) {
    return guarded(ctx, [&]() -> int32_t {
    if (!ctx) return RXC_ERR_INVALID;
    if (!ptr || bytes == 0) return fail(ctx, RXC_ERR_INVALID, "rxc_pin_host: null pointer or empty range");
    CK(cudaSetDevice(ctx->device));
    cudaError_t e = cudaHostRegister(ptr, (size_t)bytes, cudaHostRegisterDefault);
    if (e == cudaErrorHostMemoryAlreadyRegistered) { cudaGetLastError(); return RXC_OK; }
    if (e != cudaSuccess) { cudaGetLastError(); return fail(ctx, RXC_ERR_CUDA, std::string("cudaHostRegister: ") + cudaGetErrorString(e)); }
    return RXC_OK;
    });
}

int32_t rxc_unpin_host(rxc_ctx* ctx, void* ptr) {
    return guarded(ctx, [&]() -> int32_t {
    if (!ctx || !ptr) return RXC_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));   // nothing may still be draining into it
    if (ctx->copy_stream) CK(cudaStreamSynchronize(ctx->copy_stream));
    cudaError_t e = cudaHostUnregister(ptr);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(ctx, RXC_ERR_CUDA, std::string("cudaHostUnregister: ") + cudaGetErrorString(e)); }
    return RXC_OK;
    });
}

int32_t rxc_owner_base(const rxc_ctx* ctx, uint32_t batch, uint32_t* base) {
    return guarded(nullptr, [&]() -> int32_t {
    if (!ctx || !base || batch >= ctx->owner_base.size()) return RXC_ERR_INVALID;
    *base = ctx->owner_base[batch];
    return RXC_OK;
    });
}

int32_t rxc_selftest_div(rxc_ctx* ctx, uint64_t seed, uint64_t n_pairs, uint64_t* mismatches, uint32_t* bad_pair) {
    return guarded(ctx, [&]() -> int32_t {
    if (!ctx || !mismatches) return RXC_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    unsigned long long* d = nullptr;
    CK(cudaMalloc((void**)&d, 2 * sizeof(unsigned long long)));
    CK(cudaMemsetAsync(d, 0, 2 * sizeof(unsigned long long), ctx->stream));
    const uint32_t blocks = (uint32_t)ctx->sm_count * 8u;
    const uint32_t iters = (uint32_t)std::max<uint64_t>(1, n_pairs / ((uint64_t)blocks * 256u));
    cudaError_t e = rxk_selftest_div(seed, blocks, iters, d, ctx->stream);
    unsigned long long h[2] = {0, 0};
    if (e == cudaSuccess) e = cudaMemcpyAsync(h, d, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d);
    if (e != cudaSuccess) { ctx->err = std::string("rxc_selftest_div: ") + cudaGetErrorString(e); return RXC_ERR_CUDA; }
    *mismatches = h[0];
    if (bad_pair) { bad_pair[0] = (uint32_t)(h[1] >> 32); bad_pair[1] = (uint32_t)h[1]; }
    return RXC_OK;
    });
}

int32_t rxc_vm_execute(rxc_ctx* ctx, uint32_t program, uint32_t n, const float* in, float* out, uint32_t* faults) {
    return guarded(ctx, [&]() -> int32_t {
    if (!ctx || !in || !out) return RXC_ERR_INVALID;
    if (!ctx->have_scene || program >= ctx->S.vm.n_programs) return fail(ctx, RXC_ERR_INDEX, "rxc_vm_execute: no such program in the current scene");
    if (n == 0) return RXC_OK;
    CK(cudaSetDevice(ctx->device));
    float *d_in = nullptr, *d_out = nullptr;
    uint32_t* d_f = nullptr;
    cudaError_t e = cudaMalloc((void**)&d_in, (size_t)n * 18 * 4);
    if (e == cudaSuccess) e = cudaMalloc((void**)&d_out, (size_t)n * 24 * 4);
    if (e == cudaSuccess) e = cudaMalloc((void**)&d_f, 4);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_in, in, (size_t)n * 18 * 4, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(d_f, 0, 4, ctx->stream);
    void* jit_kernel = nullptr;
    if (ctx->jit) {
        std::string note;
        jit_kernel = ctx->jit_translated ? rxj_kernel(ctx->jit, -1, false, 2, std::string(), &note) : nullptr;
        if (!note.empty()) ctx->jit_note = note;
    }
    if (e == cudaSuccess) { ctx->stats.kernel_launches++; e = rxk_vm_execute(ctx->S, program, n, d_in, d_out, d_f, ctx->stream, jit_kernel); }
    uint32_t h_f = 0;
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, d_out, (size_t)n * 24 * 4, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(&h_f, d_f, 4, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_in); cudaFree(d_out); cudaFree(d_f);
    if (e != cudaSuccess) { ctx->err = std::string("rxc_vm_execute: ") + cudaGetErrorString(e); return RXC_ERR_CUDA; }
    if (faults) *faults = h_f;
    return RXC_OK;
    });
}

int64_t rxc_vm_translate(const rxc_program* programs, uint32_t n_programs, char* source, uint64_t cap, uint32_t* jit_index) {
    try {
        if (n_programs && !programs) return RXC_ERR_INVALID;
        std::string src;
        std::vector<uint32_t> idx;
        rxj_generate(programs, n_programs, &src, &idx);
        if (jit_index) for (uint32_t i = 0; i < n_programs; ++i) jit_index[i] = idx[i];
        if (source && cap) { const size_t n = std::min<size_t>(src.size(), (size_t)cap - 1); memcpy(source, src.data(), n); source[n] = 0; }
        return (int64_t)src.size();
    } catch (...) {
        return RXC_ERR_OOM;
    }
}

int64_t rxc_vm_jit_compile(const rxc_program* programs, uint32_t n_programs, int32_t sample_mode, int32_t planes, char* log, uint32_t log_cap) {
    try {
        if ((n_programs && !programs) || sample_mode < -2 || sample_mode > 2) return RXC_ERR_INVALID;
        std::string src, msg;
        std::vector<uint32_t> idx;
        if (!rxj_generate(programs, n_programs, &src, &idx)) return 0;
        const size_t bytes = rxj_compile_offline(src, sample_mode, planes != 0, 2, std::string(), &msg);
        if (log && log_cap) { const size_t n = std::min<size_t>(msg.size(), log_cap - 1); memcpy(log, msg.data(), n); log[n] = 0; }
        return bytes ? (int64_t)bytes : (int64_t)RXC_ERR_UNSUPPORTED;
    } catch (...) {
        return RXC_ERR_OOM;
    }
}

int32_t rxc_vm_state_report(const rxc_program* programs, uint32_t n_programs, const uint8_t* usage, int32_t scene_has_3d, uint32_t* report) {
    try {
        if ((n_programs && (!programs || !report))) return RXC_ERR_INVALID;
        if (n_programs) rxj_state_report(programs, n_programs, usage, scene_has_3d != 0, report);
        return RXC_OK;
    } catch (...) {
        return RXC_ERR_OOM;
    }
}

int32_t rxc_vm_scene_state_report(rxc_ctx* ctx, uint32_t* report, uint32_t cap, uint32_t* n_programs) {
    return guarded(ctx, [&]() -> int32_t {
    if (!ctx) return RXC_ERR_INVALID;
    if (n_programs) *n_programs = (uint32_t)ctx->vm_state_report.size();
    for (uint32_t i = 0; report && i < cap && i < ctx->vm_state_report.size(); ++i) report[i] = ctx->vm_state_report[i];
    return RXC_OK;
    });
}

int32_t rxc_set_vm_state_mode(rxc_ctx* ctx, int32_t mode) {
    return guarded(ctx, [&]() -> int32_t {
    if (!ctx || mode < 0 || mode > 2) return RXC_ERR_INVALID;
    ctx->vm_state_mode = mode;
    return RXC_OK;
    });
}

int32_t rxc_get_vm_state_mode(rxc_ctx* ctx, int32_t* mode, uint64_t* ordered_frames) {
    return guarded(ctx, [&]() -> int32_t {
    if (!ctx) return RXC_ERR_INVALID;
    if (mode) *mode = ctx->vm_state_mode;
    if (ordered_frames) *ordered_frames = ctx->ordered_frames;
    return RXC_OK;
    });
}

int32_t rxc_set_vm_jit(rxc_ctx* ctx, int32_t mode) {
    return guarded(ctx, [&]() -> int32_t {
    if (!ctx || mode < 0 || mode > 2) return RXC_ERR_INVALID;
    ctx->vm_jit = mode;
    return RXC_OK;
    });
}

int32_t rxc_vm_jit_info(rxc_ctx* ctx, uint32_t* n_translated, uint32_t* kernels_compiled, uint32_t* pending, uint64_t* jit_launches, char* log, uint32_t log_cap) {
    return guarded(ctx, [&]() -> int32_t {
    if (!ctx) return RXC_ERR_INVALID;
    uint64_t compiled = 0, used = 0;
    rxj_stats(ctx->jit, &compiled, &used);
    if (n_translated) *n_translated = ctx->jit_translated;
    if (kernels_compiled) *kernels_compiled = (uint32_t)compiled;
    if (pending) *pending = rxj_idle(ctx->jit) ? 0u : 1u;
    if (jit_launches) *jit_launches = used;
    if (log && log_cap) { const size_t n = std::min<size_t>(ctx->jit_note.size(), log_cap - 1); memcpy(log, ctx->jit_note.data(), n); log[n] = 0; }
    return RXC_OK;
    });
}

int32_t rxc_set_profiling(rxc_ctx* ctx, int32_t enabled) {
    return guarded(ctx, [&]() -> int32_t {
    if (!ctx) return RXC_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    drain_events(ctx);
    ctx->profiling = enabled != 0;
    return RXC_OK;
    });
}

int32_t rxc_get_stats(rxc_ctx* ctx, rxc_stats* out) {
    return guarded(ctx, [&]() -> int32_t {
    if (!ctx || !out) return RXC_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    if (ctx->profiling) { CK(cudaStreamSynchronize(ctx->stream)); drain_events(ctx); }
    *out = ctx->stats;
    return RXC_OK;
    });
}

int32_t rxc_reset_stats(rxc_ctx* ctx) {
    return guarded(ctx, [&]() -> int32_t {
    if (!ctx) return RXC_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    drain_events(ctx);
    ctx->stats = rxc_stats{};
    return RXC_OK;
    });
}

}  // extern "C"
