// rx_kernels.cuh -- launch interface between the host API (rx_api.cu) and the kernels (rx_kernels.cu)
#pragma once
#include <cuda.h>   // CUtensorMap (type only: the driver entry point is resolved at run time)

#include "rx_device.cuh"

#ifndef RX_TMA_STORE
#define RX_TMA_STORE 0   // 1: finished tiles leave shared memory as ONE 2-D tensor-map bulk store (cp.async.bulk.tensor, UTMASTG)
#endif

struct TriBin {           // 16 B per triangle: what the binning kernels stream
    uint32_t bbx, bby;    // raw pixel bbox from setup; final (scissored) bbox after k_bin_count
    uint32_t slot;        // owner id / record slot
    uint32_t batch;       // 3D batch index | large<<31 (set by k_bin_count)
};

// Scene-resident device data (set by rxc_set_assets / rxc_set_scene)
struct SceneDev {
    const float4* pos;          // [V]   positions
    const float2* uv;           // [V]
    const float* nrm;           // [V*3] (unused entries for batches without normals)
    const uint32_t* idx;        // [T*3] global vertex ids
    const DBatch3* b3;
    const DChunk* chunks;
    const uint32_t* orphans;    // global vertex ids
    const float2* pos2;         // 2D
    const float2* uv2;
    const uint32_t* idx2;
    const DBatch2* b2;
    const DLight* lights;       // scene lights (flicker factor not yet applied)
    const DTex* tex;
    const DTile* tiles;         // static tiles first, then dynamic tiles
    const uint8_t* arena;
    const DSector* sectors;     // occluded sectors of every chunk, then of the mapmini
    const DChunkInfo* chunk_info; // [n_scene_chunks + 1]; the last entry is the mapmini
    const float4* linedefs;     // (start.xy, end.xy) of mapmini.linedefs
    uint32_t n_scene_chunks, n_linedefs, n_actor_tiles;
    uint32_t n_sectors;         // total entries of `sectors` (0: every occlusion is 1)
    uint32_t general;           // ordered 3D lists (opacity batches) and binned 2D lists (many 2D records)
    uint32_t n_b3, n_b2, n_chunks, n_lights;
    uint32_t n_tris, n_verts;   // 3D totals
    uint32_t n_rec2d;           // 2D records per frame
    uint32_t n_static_tiles, n_dynamic_tiles;
    VmDev vm;                   // Rusteria VM programs, pattern bank and palette
};

// Per-frame workspace: every array holds `n_frames` slices of the given stride (in elements)
#define RX_SMALL_MIN_LIST 64    // tiles whose binned list is at least this long take the pass
#define RX_SMALL_MIN_TRIS 16384 // scenes with fewer triangles run the plain fast-path kernel (no tile list can be long enough to matter)
#define RX_SMALL_GSHIFT 3       // 8 lanes share a record's pixels
#define RX_SMALL_MAX_PIX 256    // pixel count of the clipped box up to which a record is "small"
struct Workspace {
    DFrame* frames;
    DFrameBatch* fb;        uint32_t fb_stride;
    DFrameBatch2* fb2;      uint32_t fb2_stride;
    DLight* lights;         uint32_t lights_stride;
    DCounters* counters;
    TriVis* vis;            uint32_t slot_stride;   // 3 * n_tris
    TriShade* shade;
    TriBin* bins;           uint32_t bins_stride;   // n_tris + new_cap
    uint32_t* chunk_new_total;
    uint32_t* chunk_new_base; uint32_t chunk_stride;
    DClip* clip;            uint32_t clip_stride;
    uint32_t* large;        uint32_t large_stride;
    uint32_t* tile_count;
    uint32_t* tile_base;
    uint32_t* tile_fill;    uint32_t tile_stride;
    uint32_t* lists;        uint32_t list_stride;
    uint32_t* tile_count2;  // 2D record lists per tile (general mode), same tile_stride
    uint32_t* tile_base2;
    uint32_t* tile_fill2;
    uint32_t* lists2;       uint32_t list2_stride;
    Tri2D* tri2d;           uint32_t tri2d_stride;
    uint32_t* raster_counter;  // RX_RASTER_COUNTERS work counters, one per k_raster launch of a call (zeroed by the frame setup)
    uint32_t raster_groups;    // > 1: the frames of a launch are dealt out in this many groups, counters [0, groups) (k_raster's work fetch)
    uint32_t small_min_list, small_max_pix;  // k_raster's thread-per-record pass of long tile lists
    uint32_t small_gshift;                   // log2 of the lanes that share one record in the pass
    uint32_t small_min_tris;                 // scenes with at least this many triangles run the k_raster variant that has the pass (0xFFFFFFFF = never)
    float neg_zero;                          // -0.0f, as a run-time value the compiler cannot see through (rx_mul2 in rx_device.cuh)
};

// rxc_rasterize_projected: the outputs of the host's own Scene::project (batch3d.rs:482-740), flattened over the batches
struct ProjectedDev {
    const float4* pv;        // [NP] projected_vertices (sx, sy, z, w)
    const float2* uv;        // [NP] clipped_uvs
    const float* nrm;        // [NP*3] clipped_normals (zeros for batches without normals)
    const uint32_t* idx;     // [NC*3] clipped_indices, global into pv
    const float* edges;      // [NC*9] Edges a[3], b[3], c[3]
    const uint32_t* info;    // [NC*2] batch | visible << 31, record slot (= owner id)
    const float* bbox;       // [n_b3*5] has_bounding_box, Rect x, y, width, height
    uint32_t n_clipped;      // NC
};

struct RasterOut {
    uint8_t* pixels;       // frame f at pixels + f*frame_stride
    uint64_t frame_stride; // bytes
    uint32_t* owner;       // optional, only for single-frame calls
    float* depth;
    uint32_t vec_store;    // rows are 16 B aligned -> 128-bit stores
    uint32_t pitch;        // pixels per row of the pixel buffer (the rendered rectangle's width unless a band is written in place)
    uint32_t tma_store;    // tmap describes the pixel buffer as (x, y, frame) of u32: tiles are stored through it (RX_TMA_STORE builds)
    alignas(64) CUtensorMap tmap;
};

enum {
    RXK_FRAME_SETUP = 0,
    RXK_TRI_SETUP = 1,
    RXK_BATCH_FINALIZE = 2,
    RXK_CLIP_EMIT = 3,
    RXK_BIN_COUNT = 4,
    RXK_TILE_ALLOC = 5,
    RXK_BIN_FILL = 6,
    RXK_RASTER = 7,
    RXK_BIN2D = 8,
    RXK_LIST_SORT = 9,
    RXK_FRONT_SMALL = 10
};

#define RX_FRONT_CLUSTER 8       // CTAs of the cluster that runs a mid-sized scene's front end
#define RX_RASTER_COUNTERS 64    // work counters of k_raster launches per call

#ifndef __CUDACC_RTC__   // host-side launch interface (the JIT recompiles the device code only)
// Each returns the cudaError_t of the launch.  `n_frames` frames are processed by one launch.
cudaError_t rxk_frame_setup(const SceneDev& S, const Workspace& W, uint32_t n_frames, uint32_t tiles_per_frame, cudaStream_t st);
// the whole front-end (frame setup .. bin fill) of a small non-general scene in one launch, one CTA per frame
cudaError_t rxk_front_small(const SceneDev& S, const Workspace& W, uint32_t n_frames, uint32_t tiles_per_frame, cudaStream_t st);
// the same for mid-sized scenes: a cluster of RX_FRONT_CLUSTER CTAs per frame shares every phase, cluster barriers between them
// stop_phase: 0 = all phases; n = return after the n-th cluster barrier (profiling aid only, the frame is then incomplete)
cudaError_t rxk_front_cluster(const SceneDev& S, const Workspace& W, uint32_t n_frames, uint32_t tiles_per_frame, uint32_t stop_phase, cudaStream_t st);
cudaError_t rxk_tri_setup(const SceneDev& S, const Workspace& W, uint32_t n_frames, cudaStream_t st);
// frame 0 only: records, bins and batch scissors from host-projected data instead of k_tri_setup / k_batch_finalize / k_clip_emit
cudaError_t rxk_tri_setup_projected(const SceneDev& S, const Workspace& W, const ProjectedDev& P, cudaStream_t st);
cudaError_t rxk_batch_finalize(const SceneDev& S, const Workspace& W, uint32_t n_frames, cudaStream_t st);
cudaError_t rxk_clip_emit(const SceneDev& S, const Workspace& W, uint32_t n_frames, int grid, cudaStream_t st);
cudaError_t rxk_bin_count(const SceneDev& S, const Workspace& W, uint32_t n_frames, int grid, cudaStream_t st);
// which: 0 = 3D lists, 1 = 2D lists; pow2: round every allocation up to a power of two (lists that get sorted)
cudaError_t rxk_tile_alloc(const SceneDev& S, const Workspace& W, uint32_t n_frames, uint32_t tiles_per_frame, int which, int pow2,
                           cudaStream_t st);
cudaError_t rxk_bin_large(const SceneDev& S, const Workspace& W, uint32_t n_frames, int fill, int grid, cudaStream_t st);
cudaError_t rxk_bin2d(const SceneDev& S, const Workspace& W, uint32_t n_frames, int fill, cudaStream_t st);
cudaError_t rxk_list_sort(const SceneDev& S, const Workspace& W, uint32_t n_frames, uint32_t tiles_per_frame, int which, cudaStream_t st);
cudaError_t rxk_bin_fill(const SceneDev& S, const Workspace& W, uint32_t n_frames, int grid, cudaStream_t st);
// sample_mode: RXC_SAMPLE_* when every frame of the launch uses it, 2 = mixed (read per frame).  The launch covers the
// tiles [tile0, tile0 + n_tiles) of every frame (row-major tile index: a range of tile rows is a horizontal slice of
// the frame) and hands them out through work counter `counter` (< RX_RASTER_COUNTERS).
cudaError_t rxk_raster(const SceneDev& S, const Workspace& W, const RasterOut& out, uint32_t n_frames, uint32_t tile0, uint32_t n_tiles,
                       uint32_t counter, int sample_mode, int grid, cudaStream_t st, void* jit_kernel = nullptr);
// the reference-order kernel (k_raster_ordered: one thread per API tile, the tile's Execution carried from fragment to fragment):
// general-mode scenes, whole frames only; the six scratch planes hold `stride` (= width * height) entries per frame
cudaError_t rxk_raster_ordered(const SceneDev& S, const Workspace& W, const RasterOut& out, uint32_t n_frames, uint32_t api_tiles, float* z, float* zop,
                               uint32_t* cop, uint32_t* sid, uint32_t* some, uint32_t* own, size_t stride, cudaStream_t st, void* jit_kernel = nullptr);
int rxk_raster_mode(const SceneDev& S, const Workspace& W);   // the MODE template argument rxk_raster picks (0 fast, 1 general, 2 general + VM, 3 fast + small-triangle pass)
int rxk_raster_blocks_per_sm();
// diagnostics: program `program` of S.vm on n records (18 floats in, 24 floats out each)
cudaError_t rxk_vm_execute(const SceneDev& S, uint32_t program, uint32_t n, const float* d_in, float* d_out, uint32_t* d_faults, cudaStream_t st, void* jit_kernel = nullptr);
// diagnostics: rx_div_by vs div.rn on blocks*256*iters random operand pairs; adds the mismatch count
cudaError_t rxk_selftest_div(uint64_t seed, uint32_t blocks, uint32_t iters, unsigned long long* d_mismatches, cudaStream_t st);
#endif  // !__CUDACC_RTC__
