/*
 * rxcuda.h -- C ABI of the B200-native replacement for Rusterix's CPU tile rasterizer.
 *
 * The reference has no FFI: the boundary it replaces is the Rust method
 *   Rasterizer::setup(m2d, view, proj)[.ambient()/.sample_mode()/...]
 *       .rasterize(&mut scene, pixels, width, height, tile_size, &assets)
 * (reference src/rasterizer.rs:92-152 and :185-193).  A Rust host marshals its
 * Scene / Batch3D / Batch2D / Tile / Texture / CompiledLight values into the POD structs
 * below (all #[repr(C)]-mirrorable: plain pointers, sizes, fixed arrays) and calls these
 * entry points; see INTEGRATION.md for the cc/build.rs + extern "C" stub.
 *
 * Conventions
 *  - every function returns RXC_OK (0) or a negative rxc_status; nothing unwinds;
 *    rxc_last_error(ctx) gives a human readable string owned by the context.
 *  - host memory passed to rxc_set_* is only borrowed for the duration of the call.
 *  - matrices are column-major (vek `Mat4.cols`, reference src/rasterizer.rs:99-101):
 *    m[c*4+r] is row r, column c.  Mat3 likewise m[c*3+r].
 *  - a context is bound to one GPU and one CUDA stream; it is Send but not Sync
 *    (same as `&mut Rasterizer` + `&mut Scene` in the reference).
 *  - there is no CPU fallback: without a usable sm_100 device rxc_create fails.
 */
#ifndef RXCUDA_H
#define RXCUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RXC_ABI_VERSION 6u

typedef struct rxc_ctx rxc_ctx;

typedef enum rxc_status {
    RXC_OK = 0,
    RXC_ERR_INVALID = -1,     /* bad argument (null pointer, zero size, short buffer ...)     */
    RXC_ERR_CUDA = -2,        /* a CUDA runtime call failed; see rxc_last_error                */
    RXC_ERR_UNSUPPORTED = -3, /* feature of the reference path not (yet) on the device path    */
    RXC_ERR_OOM = -4,         /* device or pinned allocation failed                            */
    RXC_ERR_INDEX = -5,       /* index out of range where the reference would panic            */
    RXC_ERR_NO_DEVICE = -6    /* no sm_100-class GPU visible                                   */
} rxc_status;

/* PrimitiveMode, reference src/batch/mod.rs:5-15 */
enum { RXC_MODE_TRIANGLES = 0, RXC_MODE_LINES = 1, RXC_MODE_LINE_STRIP = 2, RXC_MODE_LINE_LOOP = 3 };
/* CullMode, reference src/batch/mod.rs:18-26 */
enum { RXC_CULL_OFF = 0, RXC_CULL_FRONT = 1, RXC_CULL_BACK = 2 };
/* RepeatMode, reference src/texture.rs:15-25 */
enum { RXC_REPEAT_CLAMP_XY = 0, RXC_REPEAT_REPEAT_XY = 1, RXC_REPEAT_REPEAT_X = 2, RXC_REPEAT_REPEAT_Y = 3 };
/* SampleMode, reference src/texture.rs:6-12 (a field of Rasterizer, not of the batch) */
enum { RXC_SAMPLE_NEAREST = 0, RXC_SAMPLE_LINEAR = 1 };
/* LightType, reference src/map/light.rs:7-14 */
enum {
    RXC_LIGHT_POINT = 0,
    RXC_LIGHT_AMBIENT = 1,
    RXC_LIGHT_AMBIENT_DAYLIGHT = 2,
    RXC_LIGHT_SPOT = 3,
    RXC_LIGHT_AREA = 4,
    RXC_LIGHT_DAYLIGHT = 5
};
/* PixelSource, reference src/map/pixelsource.rs:23-37.  Kinds the rasterizer treats alike
 * (Off, TileId, MaterialId, Sequence, Color, ShapeFXGraphId) collapse to RXC_SRC_OTHER:
 * opaque black in 3D (src/rasterizer.rs:1221), transparent in 2D (:757). */
enum {
    RXC_SRC_OTHER = 0,
    RXC_SRC_STATIC_TILE = 1,  /* index into the tiles given to rxc_set_assets (assets.tile_list) */
    RXC_SRC_DYNAMIC_TILE = 2, /* index into rxc_scene.dynamic_textures                           */
    RXC_SRC_PIXEL = 3,        /* constant RGBA8                                                  */
    RXC_SRC_ENTITY_TILE = 4,  /* EntityTile(id, i): the host resolves assets.entity_tiles[id].get_index(i)
                                 to an index into rxc_scene.actor_tiles; 0xFFFFFFFF = not found, which
                                 samples as [0,0,0,0] (src/rasterizer.rs:1130-1152, :712-733)    */
    RXC_SRC_ITEM_TILE = 5,    /* ItemTile(id, i): same, through assets.item_tiles                */
    RXC_SRC_TERRAIN = 6       /* chunk.sample_terrain_texture of the batch's chunk (src/chunk.rs:135-151);
                                 without a chunk: [255,0,0,255] in 3D, transparent in 2D          */
};
/* Background shaders, reference src/shader/vgradient.rs, src/shader/grid.rs */
enum { RXC_BG_NONE = 0, RXC_BG_VGRAY_GRADIENT = 1, RXC_BG_GRID = 2 };
/* Which list of Scene a 3D batch came from; batches are passed in SUBMISSION ORDER
 * (reference src/rasterizer.rs:314-405) and the tag only names the origin. */
enum { RXC_PASS_STATIC = 0, RXC_PASS_DYNAMIC = 1, RXC_PASS_OVERLAY = 2, RXC_PASS_CHUNK = 3, RXC_PASS_CHUNK_OPACITY = 4 };
/* Mat*Vec rounding convention (vek 0.17 is not vendored in the reference):
 *  FMA_COLUMNS: r = col0*x; r = fma(col1,y,r); r = fma(col2,z,r); r = fma(col3,w,r)
 *  PLAIN_ROWS : r_i = ((m_i0*x + m_i1*y) + m_i2*z) + m_i3*w  with every op rounded */
enum { RXC_MATVEC_FMA_COLUMNS = 0, RXC_MATVEC_PLAIN_ROWS = 1 };

/* One animation frame of a Tile: reference src/texture.rs:46-54 (data/width/height). */
typedef struct rxc_texture {
    const uint8_t* data; /* RGBA8 row-major, width*height*4 bytes */
    uint32_t width;
    uint32_t height;
} rxc_texture;

/* reference src/map/tile.rs:83-96: the frame used is animation_frame % n_textures */
typedef struct rxc_tile {
    const rxc_texture* textures;
    uint32_t n_textures;
} rxc_tile;

/* reference src/map/light.rs:457-477 (CompiledLight) */
typedef struct rxc_light {
    uint32_t light_type;
    float position[3];
    float color[3];
    float intensity;
    uint32_t emitting;
    float start_distance;
    float end_distance;
    float flicker;
    float direction[3];
    float cone_angle;
    float normal[3];
    float width;
    float height;
    uint32_t from_linedef;
} rxc_light;

/* reference src/batch/batch3d.rs:15-78 (inputs only; the projection caches stay on device) */
typedef struct rxc_batch3d {
    const float* vertices;  /* n_vertices * [x,y,z,w]                                   */
    const float* uvs;       /* n_vertices * [u,v]                                       */
    const float* normals;   /* n_vertices * [x,y,z], or NULL when batch.normals is empty */
    const void* indices;    /* n_triangles * 3 indices, index_bytes wide each           */
    uint32_t n_vertices;
    uint32_t n_triangles;
    uint32_t index_bytes;   /* 4 (u32) or 8 (Rust usize triples, 24 B per triangle)     */
    uint32_t mode;          /* RXC_MODE_*; 3D batches are always rasterized as triangles */
    uint32_t repeat_mode;   /* RXC_REPEAT_* */
    uint32_t cull_mode;     /* RXC_CULL_*   */
    uint32_t source_kind;   /* RXC_SRC_*    */
    uint32_t source_index;  /* tile index for STATIC_TILE / DYNAMIC_TILE                */
    uint8_t source_pixel[4];/* RGBA for RXC_SRC_PIXEL                                   */
    uint32_t receives_light;
    float ambient_color[3];
    uint32_t has_profile_id;
    uint32_t profile_id;
    int32_t shader;         /* batch.shader: -1 = None, else an index into scene.shaders, or into the
                               shaders of the batch's chunk (rxc_chunk.shader_base) when chunk >= 0
                               (src/rasterizer.rs:1226-1293); an index past the list runs no program */
    uint32_t pass;          /* RXC_PASS_*; CHUNK_OPACITY batches go through d3_rasterize_opacity
                               (src/rasterizer.rs:1425-1690), everything else through d3_rasterize */
    float transform[16];    /* transform_3d, column-major                                */
    int32_t chunk;          /* index into rxc_scene.chunks of the Chunk the batch belongs to, -1 = none
                               (the `chunk: Option<&Chunk>` argument, src/rasterizer.rs:973)      */
} rxc_batch3d;

/* reference src/batch/batch2d.rs:10-53 */
typedef struct rxc_batch2d {
    const float* vertices; /* n_vertices * [x,y] */
    const float* uvs;      /* n_vertices * [u,v] */
    const void* indices;   /* n_triangles * 3    */
    uint32_t n_vertices;
    uint32_t n_triangles;
    uint32_t index_bytes;
    uint32_t mode;
    uint32_t repeat_mode;
    uint32_t source_kind;
    uint32_t source_index;
    uint8_t source_pixel[4];
    uint32_t receives_light;
    int32_t shader;
    int32_t chunk;         /* as in rxc_batch3d */
} rxc_batch2d;

/* Rusteria VM programs (reference rusteria/src/node/nodeop.rs:12-103, program.rs:7-29, execution.rs:109-768).
 * The host lowers the NodeOp tree of Program.user_functions[shade_index] and of every function it can call
 * to a stream of 32-bit words: word = opcode | (a << 8).  Opcodes 0..89 are the NodeOp variants in declaration
 * order (RXVM_LOAD_GLOBAL = 0 ... RXVM_PALETTE_INDEX = 89); the lowering adds RXVM_JZ..RXVM_END:
 *   LoadGlobal/StoreGlobal/LoadLocal/StoreLocal   a = index
 *   GetComponents   a = n | c0<<3 | c1<<5 | c2<<7 with the swizzle entries > 2 already dropped; n = 7 pushes 0
 *   SetComponents   a = n | c0<<3 | c1<<5 | c2<<7, n = swizzle length if 1..3 else 0, c = 3 is ignored
 *   Push            followed by three words: the f32 bits of x, y, z
 *   FunctionCall    a = arity | total_locals<<8, followed by one word: the offset of the callee
 *   If(t, e)        Jz L1; t...; Jmp L2; L1: e...; L2:        (Jz pops and jumps when value.x == 0.0)
 *   For(i,c,n,b)    Mark; i; Trunc; L: c; Jz E; Trunc; b; Trunc; n; Trunc; Jmp L; E: Unmark
 *                   (Mark remembers the stack height, Trunc is stack.truncate(base), execution.rs:259-285)
 *   every function body ends with End; Alloc / Iterate / Save (texture baking) are rejected.
 * Device limits: value stack 32, 32 locals per frame, 16 globals, call depth 8, loop depth 8, 2^20 ops per call. */
enum {
    RXVM_LOAD_GLOBAL = 0, RXVM_STORE_GLOBAL, RXVM_LOAD_LOCAL, RXVM_STORE_LOCAL, RXVM_SWAP, RXVM_GET_COMPONENTS,
    RXVM_SET_COMPONENTS, RXVM_IF, RXVM_FOR, RXVM_PUSH, RXVM_FUNCTION_CALL, RXVM_RETURN, RXVM_DUP, RXVM_CLEAR, RXVM_PACK2,
    RXVM_PACK3, RXVM_ADD, RXVM_SUB, RXVM_MUL, RXVM_DIV, RXVM_LENGTH, RXVM_LENGTH2, RXVM_LENGTH3, RXVM_ABS, RXVM_SIN,
    RXVM_SIN1, RXVM_SIN2, RXVM_COS, RXVM_COS1, RXVM_COS2, RXVM_TAN, RXVM_ATAN, RXVM_ATAN2, RXVM_ROTATE2D, RXVM_DOT,
    RXVM_DOT2, RXVM_DOT3, RXVM_CROSS, RXVM_NORMALIZE, RXVM_FLOOR, RXVM_CEIL, RXVM_ROUND, RXVM_FRACT, RXVM_MOD,
    RXVM_DEGREES, RXVM_RADIANS, RXVM_MIN, RXVM_MAX, RXVM_MIX, RXVM_SMOOTHSTEP, RXVM_STEP, RXVM_CLAMP, RXVM_SQRT,
    RXVM_POW, RXVM_LOG, RXVM_PRINT, RXVM_EQ, RXVM_NE, RXVM_LT, RXVM_LE, RXVM_GT, RXVM_GE, RXVM_AND, RXVM_OR, RXVM_NOT,
    RXVM_NEG, RXVM_UV, RXVM_SET_UV, RXVM_NORMAL, RXVM_SET_NORMAL, RXVM_HITPOINT, RXVM_TIME, RXVM_SAMPLE,
    RXVM_SAMPLE_NORMAL, RXVM_COLOR, RXVM_SET_COLOR, RXVM_ROUGHNESS, RXVM_SET_ROUGHNESS, RXVM_METALLIC, RXVM_SET_METALLIC,
    RXVM_EMISSIVE, RXVM_SET_EMISSIVE, RXVM_OPACITY, RXVM_SET_OPACITY, RXVM_BUMP, RXVM_SET_BUMP, RXVM_ALLOC, RXVM_ITERATE,
    RXVM_SAVE, RXVM_PALETTE_INDEX,
    RXVM_JZ, RXVM_JMP, RXVM_MARK, RXVM_TRUNC, RXVM_UNMARK, RXVM_END, RXVM_N_OPS
};
typedef struct rxc_program {
    const uint32_t* code;   /* NULL / 0 words: Program.shade_index is None, nothing runs             */
    uint32_t n_words;
    uint32_t entry;         /* word offset of shade()                                                */
    uint32_t shade_locals;  /* Program.shade_locals                                                  */
    uint32_t n_globals;     /* Program.globals                                                       */
    uint32_t sets_opacity;  /* Program::shader_supports_opacity() (program.rs:44-55)                 */
} rxc_program;
/* a pattern texture of the VM's `sample` op: TexStorage of Value = Vec3<f32> (rusteria/src/textures/mod.rs:11-24),
 * nearest lookup with wrap-around (:131-146) */
typedef struct rxc_pattern {
    const float* data;      /* width*height * [x,y,z] */
    uint32_t width;
    uint32_t height;
} rxc_pattern;

/* (BBox, occlusion) entries of chunk.occluded_sectors / mapmini.occluded_sectors
 * (src/chunk.rs:41, src/map/mini.rs:33, src/map/bbox.rs:35-40: contains() is inclusive). */
typedef struct rxc_sector {
    float min[2];
    float max[2];
    float occlusion;
} rxc_sector;

/* reference src/chunk.rs:23-57: the members the rasterizer reads besides the batches (which the
 * host flattens into batches3d / batches2d) and the lights (appended to the light list). */
typedef struct rxc_chunk {
    int32_t origin[2];
    int32_t size;
    const rxc_sector* occluded_sectors; /* searched in order, first hit wins (src/chunk.rs:154-161) */
    uint32_t n_occluded_sectors;
    const rxc_texture* terrain_texture; /* NULL = None */
    uint32_t shader_base;               /* chunk.shaders[i] is rxc_scene.shaders[shader_base + i]     */
    uint32_t n_shaders;
    const rxc_texture* const* shader_textures; /* chunk.shader_textures (n_shaders entries, NULL = None) or NULL:
                                           a baked texture replaces the batch's texel and its program does
                                           not run in the opaque 3D pass (src/rasterizer.rs:1227-1262)     */
} rxc_chunk;

/* CompiledLinedef start/end (src/map/mini.rs:88-95: a light is blocked when the segment
 * pixel->light crosses any of them) */
typedef struct rxc_linedef {
    float start[2];
    float end[2];
} rxc_linedef;

/* Rasterizer.mapmini (src/rasterizer.rs:71, src/map/mini.rs:22-38): line-of-sight and sector
 * occlusion for batches that do not belong to a chunk. */
typedef struct rxc_mapmini {
    const rxc_linedef* linedefs;
    uint32_t n_linedefs;
    const rxc_sector* occluded_sectors;
    uint32_t n_occluded_sectors;
} rxc_mapmini;

/* reference src/scene.rs:8-50.  batches3d: per chunk (batches3d_opacity..., batches3d...,
 * terrain_batch3d), then d3_static, d3_dynamic, d3_overlay; batches2d: per chunk (batches2d...,
 * terrain_batch2d), then d2_static, d2_dynamic (src/rasterizer.rs:314-405, :501-553).  lights = scene.lights followed by
 * scene.dynamic_lights AFTER the per-call chunk-light append (src/rasterizer.rs:219-223). */
typedef struct rxc_scene {
    const rxc_batch3d* batches3d;
    uint32_t n_batches3d;
    const rxc_batch2d* batches2d;
    uint32_t n_batches2d;
    const rxc_light* lights;
    uint32_t n_lights;
    const rxc_tile* dynamic_textures;
    uint32_t n_dynamic_textures;
    const rxc_chunk* chunks;          /* scene.chunks.values() in the host's iteration order */
    uint32_t n_chunks;
    const rxc_tile* actor_tiles;      /* tiles RXC_SRC_ENTITY_TILE / RXC_SRC_ITEM_TILE index     */
    uint32_t n_actor_tiles;
    const rxc_program* shaders;       /* scene.shaders, then the shaders of every chunk (rxc_chunk.shader_base) */
    uint32_t n_shaders;
    uint32_t n_scene_shaders;         /* scene.shaders.len(): the range batches without a chunk index     */
    const rxc_pattern* patterns;      /* rusteria patterns() bank (textures/patterns.rs:88-102), host computed */
    uint32_t n_patterns;
    const rxc_pattern* patterns_normal;
    uint32_t n_patterns_normal;
    const float* palette;             /* assets.palette.colors as n_palette * [present, r, g, b]          */
    uint32_t n_palette;
} rxc_scene;

/* Everything Rasterizer::setup + the builder methods + rasterize()'s scalar arguments carry
 * (reference src/rasterizer.rs:35-88, :92-193). */
typedef struct rxc_frame {
    float view[16];
    float projection[16];
    float inverse_view[16];       /* host computes with vek so the bits match its own      */
    float inverse_projection[16];
    uint32_t has_matrix2d;
    float matrix2d[9];            /* projection_matrix_2d                                   */
    uint32_t width;
    uint32_t height;
    uint32_t tile_size;           /* the API tile size; only used to reproduce the per-tile
                                     batch-bbox reject exactly (src/rasterizer.rs:978-983)   */
    uint32_t sample_mode;
    uint32_t has_background_color;
    uint8_t background_color[4];
    uint32_t background_shader;   /* RXC_BG_* (scene.background)                            */
    float grid_size;              /* GridShader parameters                                  */
    float grid_subdivisions;
    float grid_offset[2];
    uint32_t has_ambient;
    float ambient[4];
    uint64_t animation_frame;     /* scene.animation_frame                                  */
    float time;
    float hour;
    uint32_t d2_active;           /* RenderMode */
    uint32_t d3_active;
    uint32_t ignore_background_shader;
    uint32_t preserve_transparency;
    uint32_t matvec_mode;         /* RXC_MATVEC_* */
    uint32_t band_y0;             /* multi-GPU band split: render rows [band_y0, band_y1);   */
    uint32_t band_y1;             /* both 0 = whole frame. `pixels` then holds only the band */
    uint32_t band_x0;             /* ... and columns [band_x0, band_x1) (both 0 = every column): the output buffer */
    uint32_t band_x1;             /* holds the rectangle, rows of (band_x1 - band_x0) pixels; band_x0 a multiple of 32 */
    /* Render graph (src/rasterizer.rs:227-253, :419-461; src/shapestack/shapefx.rs:935-1223).  The graph stays
     * on the host: it runs collect_nodes_from / render_setup / render_ambient_color (the latter lands in
     * `ambient`) and passes what the per-pixel code reads. */
    uint32_t has_sun;             /* self.sun_dir.is_some(): directional sun in the lighting block (:1342-1361) */
    float sun_dir[3];
    float day_factor;
    uint32_t has_sky;             /* a Sky node is among render_miss: render_miss_d3 (shapefx.rs:1122-1223) colours
                                     the pixels no geometry covered; the last Sky node wins, so one is passed */
    float sky[6][4];              /* its `precomputed`: (sun_dir, day_factor), haze, day_horizon, day_zenith,
                                     night_horizon, night_zenith (shapefx.rs:1016-1055)                       */
    uint32_t sky_clouds;          /* 1 = the node's cloud layer (shapefx.rs:1172-1219), which needs noiselib 0.2.4's
                                     perlin_noise_2d (not vendored with the reference): RXC_ERR_UNSUPPORTED.
                                     0 = the host asks for the sky without the cloud layer                  */
    uint32_t has_brush_preview;   /* self.brush_preview (src/rasterizer.rs:13-17, :434-456)                 */
    float brush_position[3];
    float brush_radius;
    float brush_falloff;
} rxc_frame;

/* Counters filled by rxc_get_stats; times are device times from CUDA events (profiling on). */
#define RXC_N_KERNELS 11
typedef struct rxc_stats {
    uint64_t frames;                       /* frames rasterized since create/reset           */
    uint64_t kernel_launches;              /* kernels launched since create/reset            */
    uint64_t launches[RXC_N_KERNELS];      /* per kernel class (rxc_kernel_name)             */
    double kernel_ms[RXC_N_KERNELS];       /* accumulated, only while profiling is enabled   */
    uint64_t h2d_bytes;                    /* bytes copied host->device since reset          */
    uint64_t d2h_bytes;
    uint32_t last_binned_refs;             /* triangle references in the last frame's bins   */
    uint32_t last_large_tris;              /* triangles on the last frame's large list       */
    uint32_t last_clipped_tris;            /* near-clip output triangles of the last frame   */
    uint32_t last_visible_tris;
} rxc_stats;

uint32_t rxc_abi_version(void);
int32_t rxc_create(int32_t device, rxc_ctx** out);
void rxc_destroy(rxc_ctx* ctx);
const char* rxc_last_error(const rxc_ctx* ctx);

/* Launch on `cuda_stream` (a cudaStream_t) instead of the context's own stream; NULL restores it.
 * The kernels of a frame follow each other by programmatic dependent launch (the next kernel's blocks become resident early and
 * wait for their predecessor's completion on the device); work the caller enqueues on the stream before or after a call is
 * ordered as usual.  RXC_PDL=0 in the environment launches them the ordinary way. */
int32_t rxc_set_stream(rxc_ctx* ctx, void* cuda_stream);

/* assets.tile_list (reference src/server/assets.rs:19): uploaded once, kept on device. */
int32_t rxc_set_assets(rxc_ctx* ctx, const rxc_tile* tiles, uint32_t n_tiles);
/* Geometry, lights and dynamic textures of a Scene: uploaded once, kept on device. */
int32_t rxc_set_scene(rxc_ctx* ctx, const rxc_scene* scene);
/* Replace only the light list (lights animate per frame in examples/cube.rs:72-73). */
/* The same for a scene that differs from the resident one only behind its first `keep_batches3d` 3D batches (an engine's frame
 * loop: the world stays, the entities' dynamic batches change -- the reference re-projects `&mut scene` every call, src/scene.rs:154-200).
 * `scene` describes the WHOLE scene as for rxc_set_scene; of the kept batches only the per-batch state (source, shader, transform,
 * ambient, modes ...) is read -- their vertex / uv / normal / index arrays are not touched (the pointers may dangle), their flattened
 * copies, bounding boxes and setup chunks stay on the device and only the rest is validated, flattened and uploaded.  A kept batch
 * must have the vertex and triangle count it had (else RXC_ERR_INVALID).  Everything else (2D batches, textures, chunks, lights,
 * programs) is taken from `scene` as usual. */
int32_t rxc_update_scene(rxc_ctx* ctx, const rxc_scene* scene, uint32_t keep_batches3d);
int32_t rxc_set_lights(rxc_ctx* ctx, const rxc_light* lights, uint32_t n_lights);
/* Rasterizer.mapmini; NULL or all-empty = MapMini::default() (everything visible, occlusion 1). */
int32_t rxc_set_mapmini(rxc_ctx* ctx, const rxc_mapmini* mapmini);

/* The replacement for Rasterizer::rasterize.  `pixels` receives width*rows*4 bytes RGBA8
 * (rows = height, or the band height); it may be host (pageable or pinned) or device memory.
 * `owner` (u32 per pixel: submission ordinal of the triangle owning the pixel after the 3D
 * passes, 0xFFFFFFFF = none) and `depth` (f32 z-buffer after the 3D passes) are optional
 * parity outputs and may be NULL.  Synchronous: outputs are valid on return. */
int32_t rxc_rasterize(rxc_ctx* ctx, const rxc_frame* frame, uint8_t* pixels, uint32_t* owner, float* depth);
/* Same, but only enqueues on the stream; `pixels` must be device or pinned memory. */
int32_t rxc_rasterize_async(rxc_ctx* ctx, const rxc_frame* frame, uint8_t* pixels, uint32_t* owner, float* depth);
/* Camera sweep: n_frames frames of the current scene in one launch sequence; frame i is
 * written at pixels + i*frame_stride_bytes.  All frames must share width/height/band. */
int32_t rxc_rasterize_batch(rxc_ctx* ctx, const rxc_frame* frames, uint32_t n_frames, uint8_t* pixels,
                            uint64_t frame_stride_bytes);
int32_t rxc_rasterize_batch_async(rxc_ctx* ctx, const rxc_frame* frames, uint32_t n_frames, uint8_t* pixels,
                                  uint64_t frame_stride_bytes);
int32_t rxc_synchronize(rxc_ctx* ctx);

/* Pre-projected entry (the escape hatch for third-party arithmetic).  Coverage at triangle borders
 * depends on the exact bits of the projected vertices, i.e. on how vek 0.17 rounds Mat4 * Vec4 -- a crate that is not part
 * of the reference's sources.  A Rust host can run its own `Scene::project` (src/scene.rs:154-200, i.e.
 * Batch3D::clip_and_project, src/batch/batch3d.rs:482-740) and hand over what that leaves in every Batch3D; the device then
 * skips its own view transform / cull / near clip / projection and builds its triangle records from these values verbatim
 * (edge equations and visibility included), so that ownership is decided by the host's own vertex bits.
 * One entry per 3D batch of the current scene, in the order of rxc_scene.batches3d.  Owner ids stay comparable: triangle i
 * of clipped_indices of batch b is rxc_owner_base(b) + i (n_clipped <= 3 * n_triangles of the batch). */
typedef struct rxc_projected3d {
    const float* projected_vertices; /* n_projected * [screen x, screen y, z_ndc, w_clip]  (batch3d.rs:689-700)      */
    const float* clipped_uvs;        /* n_projected * [u, v]                                (batch3d.rs:602-607)      */
    const float* clipped_normals;    /* n_projected * [x, y, z], or NULL when the batch has no normals                */
    uint32_t n_projected;
    uint32_t n_clipped;
    const void* clipped_indices;     /* n_clipped * 3 indices into projected_vertices, index_bytes wide each          */
    uint32_t index_bytes;            /* 4 or 8 (Rust usize triples)                                                   */
    uint32_t has_bounding_box;       /* bounding_box.is_some(); None = the batch is skipped (rasterizer.rs:976-977)   */
    const float* edges;              /* n_clipped * 9: Edges.a[3], Edges.b[3], Edges.c[3]   (src/edge.rs:3-8)         */
    const uint8_t* visible;          /* n_clipped: Edges.visible                                                      */
    float bounding_box[4];           /* Rect x, y, width, height                            (batch3d.rs:762-767)      */
} rxc_projected3d;
/* rxc_rasterize with the 3D front end replaced by the host's projection results (2D batches are projected on the device
 * as usual).  Single frame, synchronous. */
int32_t rxc_rasterize_projected(rxc_ctx* ctx, const rxc_frame* frame, const rxc_projected3d* batches, uint32_t n_batches,
                                uint8_t* pixels, uint32_t* owner, float* depth);

/* Page-locks (cudaHostRegister) / releases a host pixel buffer the caller owns -- a Rust `Vec<u8>` that is reused for
 * every frame, say -- so that the frames drain over PCIe by DMA at the pinned rate (about 2.5x the pageable one) and
 * overlap with rendering.  Optional: rxc_rasterize works with pageable memory too.  Unpin before freeing the buffer. */
int32_t rxc_pin_host(rxc_ctx* ctx, void* ptr, uint64_t bytes);
int32_t rxc_unpin_host(rxc_ctx* ctx, void* ptr);

/* Global submission ordinal of triangle 0 of 3D batch `batch` (owner ids are
 * base + index into the batch's clipped_indices; capacity 3*n_triangles per batch). */
int32_t rxc_owner_base(const rxc_ctx* ctx, uint32_t batch, uint32_t* base);

/* Diagnostics: the raster kernel divides barycentric numerators by the triangle area with a
 * residual-corrected multiply by the correctly rounded reciprocal instead of div.rn (same bits,
 * fewer instructions).  This runs that routine against div.rn on about n_pairs operand pairs drawn
 * from the admitted ranges (hard mantissa patterns, quotients at rounding midpoints) on the device
 * and returns how many results differ (must be 0); bad_pair (optional, 2 words) receives the bits
 * of one failing (numerator, divisor). */
int32_t rxc_selftest_div(rxc_ctx* ctx, uint64_t seed, uint64_t n_pairs, uint64_t* mismatches, uint32_t* bad_pair);

/* Diagnostics: runs shader `program` of the current scene once per record on the device, outside the
 * rasterizer.  in: n * 18 floats (uv, color, normal, hitpoint, time, opacity as Vec3 each); out: n * 24 floats
 * (uv, color, normal, roughness, metallic, emissive, opacity, bump).  Roughness starts at 0.5, the rest of
 * Execution at zero (execution.rs:58-77).  faults (optional) counts records that hit a device limit. */
int32_t rxc_vm_execute(rxc_ctx* ctx, uint32_t program, uint32_t n, const float* in, float* out, uint32_t* faults);

/* ---- multi-GPU delivery to rank 0 ------------------------------------------------------------------
 * The path shards by frame (camera sweeps) and by screen band (one large frame) with no exchange step; what is left
 * is getting the finished pixels into ONE buffer.  In the reference that is the serial compose of the tile buffers
 * into `pixels` (src/rasterizer.rs:560-579); here it is one process and one context per GPU, and
 *   - rank 0 owns the delivery buffer (rxc_mgpu_target), every other rank maps it (cudaIpc: NVLink / NVSwitch peer
 *     access) and rxc_mgpu_rasterize makes the raster kernel's tile write-back store straight into the mapping --
 *     no gather follows the render;
 *   - rxc_mgpu_deliver marks the step complete: ranks > 0 raise a flag in rank 0's memory behind their kernels,
 *     rank 0's stream waits for all flags; rxc_mgpu_release is the reverse hand-shake (rank 0 is done with the
 *     buffer, the other ranks' streams wait for that before they overwrite it);
 *   - where a peer mapping cannot be had, ranks render into local staging and the regions move as ONE ncclGroup of
 *     send/recv pairs on a second stream (RXC_MGPU_NCCL).
 * NCCL (libnccl.so.2, resolved at run time) carries the set-up and the fallback; the host only has to get the 128
 * bytes of rxc_mgpu_unique_id from rank 0 to the other ranks (MPI, a socket, a file, torch.distributed ...).
 * rxc_mgpu_init, rxc_mgpu_target and rxc_mgpu_shutdown are collective: every rank calls them, in the same order
 * (rxc_destroy alone tears the rank's state down without waiting for anyone). */
#define RXC_MGPU_ID_BYTES 128
enum { RXC_MGPU_LOCAL = 0,  /* this rank writes the delivery buffer directly (rank 0; world == 1)        */
       RXC_MGPU_PEER = 1,   /* this rank writes rank 0's buffer through its peer mapping                 */
       RXC_MGPU_NCCL = 2 }; /* no peer mapping: local staging + grouped ncclSend/ncclRecv                */
/* What one rank contributes to a step: `rows` rows of `row_bytes` bytes, `pitch_bytes` apart (0 = contiguous), the
 * first at `offset` of the delivery buffer.  Only read in RXC_MGPU_NCCL mode (every rank passes the same list). */
typedef struct rxc_mgpu_region {
    uint32_t rank;
    uint32_t rows;
    uint64_t offset;
    uint64_t row_bytes;
    uint64_t pitch_bytes;
} rxc_mgpu_region;
int32_t rxc_mgpu_unique_id(uint8_t* id /* RXC_MGPU_ID_BYTES, an ncclUniqueId */);
int32_t rxc_mgpu_init(rxc_ctx* ctx, const uint8_t* id, uint32_t rank, uint32_t world);
int32_t rxc_mgpu_shutdown(rxc_ctx* ctx);
/* (Re)allocates the delivery buffer of `bytes` bytes on rank 0 and maps it on the other ranks.  rank0_ptr (optional)
 * receives its device address on rank 0 (NULL elsewhere), mode (optional) this rank's RXC_MGPU_* mode. */
int32_t rxc_mgpu_target(rxc_ctx* ctx, uint64_t bytes, void** rank0_ptr, uint32_t* mode);
/* rxc_rasterize_batch_async into the delivery buffer: frame i at offset_bytes + i * frame_stride_bytes.  pitch_bytes
 * = 0: rows of the rendered rectangle are contiguous; else the distance between rows, so that a band (rxc_frame.band_*)
 * lands in place inside a full frame (offset_bytes = (band_y0 * width + band_x0) * 4, pitch_bytes = width * 4). */
int32_t rxc_mgpu_rasterize(rxc_ctx* ctx, const rxc_frame* frames, uint32_t n_frames, uint64_t offset_bytes,
                           uint64_t frame_stride_bytes, uint64_t pitch_bytes);
/* Every rank, once per step, after its rxc_mgpu_rasterize calls.  When rank 0's stream has passed it, the step's
 * pixels of every rank are in the delivery buffer.  Asynchronous (stream-ordered) on every rank. */
int32_t rxc_mgpu_deliver(rxc_ctx* ctx, const rxc_mgpu_region* regions, uint32_t n_regions);
/* Every rank, once per step (or once per ring slot): rank 0 is done reading, the others may overwrite. */
int32_t rxc_mgpu_release(rxc_ctx* ctx);
/* mode of this rank, deliveries issued so far, and (rank 0, synchronizes) how many waits timed out on a dead peer */
int32_t rxc_mgpu_status(rxc_ctx* ctx, uint32_t* mode, uint32_t* deliveries, uint32_t* timeouts);

/* Batch shaders without the interpreter: rxc_set_scene translates the programs it can verify (static
 * stack heights, no run-time fault possible) to straight-line C++ and the VM variant of the raster kernel is recompiled with
 * NVRTC.  The compiler runs in a child process (`rxjitc`, installed next to the library; libnvrtc.so.12 is resolved there, at
 * run time) driven by a worker thread, results are cached on disk ($RXC_JIT_CACHE, else $XDG_CACHE_HOME/rusterix_b200); frames
 * use the interpreter until the kernel is ready and whenever rxjitc / NVRTC is missing.  Exiting or unloading the library
 * while a compilation is running kills that process.
 * Environment: RXC_VM_JIT = 0 (off) / 1 (background, default) / 2 (compile synchronously); rxc_set_vm_jit changes it for
 * the scenes set afterwards.
 * rxc_vm_translate (no context, no GPU): the generated source for a program table; jit_index[i] = i when program i was
 * accepted, 0xFFFFFFFF when it stays with the interpreter.  Returns the length of the source (written up to `cap` bytes).
 * rxc_vm_jit_compile (no context, no GPU): runs the NVRTC compilation of k_raster<sample_mode, planes, VM> (sample_mode -1: of the
 * rxc_vm_execute kernel, -2: of the reference-order kernel) for a program
 * table and returns the size of the cubin (0 = no program accepted, RXC_ERR_UNSUPPORTED = NVRTC missing or the compilation
 * failed; the compiler's messages in `log`).
 * rxc_vm_jit_info: programs translated for the current scene, kernels loaded so far, whether a requested kernel is still
 * being compiled (pending: a kernel is requested by the first frame that needs it), raster launches that used one,
 * and the last compiler / loader message (up to log_cap bytes). */
int64_t rxc_vm_translate(const rxc_program* programs, uint32_t n_programs, char* source, uint64_t cap, uint32_t* jit_index);
int64_t rxc_vm_jit_compile(const rxc_program* programs, uint32_t n_programs, int32_t sample_mode, int32_t planes, char* log, uint32_t log_cap);
int32_t rxc_set_vm_jit(rxc_ctx* ctx, int32_t mode);
/* The one documented deviation of batch shaders: the reference keeps ONE Execution per screen tile and never
 * resets it (src/rasterizer.rs:310), so state a fragment's program leaves behind (emissive, globals, the .yz of roughness ...) is
 * seen by the next fragment of the tile; the device starts every fragment from Execution::new().  A static analysis says which
 * programs can observe the difference: report[i] = 0 cannot (device == reference on this program), 1 can (writes `emissive`, reads a
 * channel the rasterizer does not fully reset and some program writes, reads a global / local before writing it), 2 not analysable.
 * rxc_vm_state_report: for a program table; usage[i] bit 0 = bound to a 3D batch, bit 1 = to a 2D batch (NULL: both).
 * rxc_vm_scene_state_report: for the current scene with its real bindings (programs no batch uses report 0). */
/* rxc_set_vm_state_mode: how batch-shader scenes are rendered.  0 = a fresh Execution per fragment, the fast tiled kernel, always;
 * 1 = the reference's own order and ONE never-reset Execution per screen tile of `tile_size` pixels (src/rasterizer.rs:310) for every
 * scene with programs: forward shading in submission order, one GPU warp per tile (k_raster_ordered) -- the reference's frame for
 * state-dependent programs too, at a fraction of the speed; 2 (the default; environment RXC_VM_STATE_MODE) = that kernel only for
 * the scenes whose report (below) flags a program (1 or 2), the fast one otherwise: every frame equals the reference's and only the scenes
 * that need it pay for it.  The kernel renders whole frames (no band) with API tiles up to 224 x 224 pixels: what it
 * cannot render is RXC_ERR_UNSUPPORTED in mode 1 and goes to the fast kernel in mode 2.
 * rxc_get_vm_state_mode: the mode and how many frames the reference-order kernel has rendered. */
int32_t rxc_set_vm_state_mode(rxc_ctx* ctx, int32_t mode);
int32_t rxc_get_vm_state_mode(rxc_ctx* ctx, int32_t* mode, uint64_t* ordered_frames);
int32_t rxc_vm_state_report(const rxc_program* programs, uint32_t n_programs, const uint8_t* usage, int32_t scene_has_3d, uint32_t* report);
int32_t rxc_vm_scene_state_report(rxc_ctx* ctx, uint32_t* report, uint32_t cap, uint32_t* n_programs);
int32_t rxc_vm_jit_info(rxc_ctx* ctx, uint32_t* n_translated, uint32_t* kernels_compiled, uint32_t* pending, uint64_t* jit_launches, char* log, uint32_t log_cap);

int32_t rxc_set_profiling(rxc_ctx* ctx, int32_t enabled);
int32_t rxc_get_stats(rxc_ctx* ctx, rxc_stats* out);
int32_t rxc_reset_stats(rxc_ctx* ctx);
const char* rxc_kernel_name(uint32_t kernel_class);

#ifdef __cplusplus
}
#endif
#endif /* RXCUDA_H */
