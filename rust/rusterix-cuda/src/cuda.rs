//! `CudaRasterizer` + `Rasterizer::rasterize_cuda`: the drop-in for `Rasterizer::rasterize` (src/rasterizer.rs:185-193)
//! on top of the C ABI of include/rxcuda.h (crate `rusterix-cuda-sys`).  Compiled inside the `rusterix` crate (see
//! README.md).  Host memory is only borrowed for the duration of each `rxc_set_*` call, so the marshalling below
//! builds short-lived arrays of PODs that point INTO the scene's own Vecs -- nothing is copied on the host.
use crate::cuda_lower::{lower_program, FlatProgram};
use crate::{Assets, Batch2D, Batch3D, Chunk, CompiledLight, CullMode, LightType, MapMini, PixelSource, PrimitiveMode, Rasterizer, RepeatMode,
            SampleMode, Scene, Texture, Tile};
use rusterix_cuda_sys::*;
use std::ffi::CStr;
use std::os::raw::c_void;

pub struct CudaRasterizer {
    ctx: *mut rxc_ctx,
    assets_key: (usize, usize),          // (tile_list.as_ptr(), len): re-upload when the list was rebuilt
    scene_key: u64,                      // structural hash of the scene (batch pointers / lengths)
    lights_key: u64,
    geometry_keys: Vec<(usize, usize, usize, usize)>,   // per 3D batch in submission order: (vertices, indices, counts) at the last upload
    pinned: Vec<(*mut u8, usize)>,
}
unsafe impl Send for CudaRasterizer {}   // one context per (thread, GPU): Send, not Sync -- like &mut Rasterizer

impl CudaRasterizer {
    pub fn new(device: i32) -> Result<Self, String> {
        let mut ctx: *mut rxc_ctx = std::ptr::null_mut();
        let st = unsafe { rxc_create(device, &mut ctx) };
        if st != RXC_OK { return Err(format!("rxc_create failed with status {st} (no sm_100 GPU visible?)")); }
        if unsafe { rxc_abi_version() } != RXC_ABI_VERSION { unsafe { rxc_destroy(ctx) }; return Err("librxcuda ABI version mismatch".into()); }
        Ok(Self { ctx, assets_key: (0, 0), scene_key: 0, lights_key: 0, geometry_keys: Vec::new(), pinned: Vec::new() })
    }

    pub fn last_error(&self) -> String { unsafe { CStr::from_ptr(rxc_last_error(self.ctx)) }.to_string_lossy().into_owned() }

    fn check(&self, st: i32, what: &str) { if st != RXC_OK { panic!("{what}: {}", self.last_error()); } }   // the reference's error style is panic

    /// Page-locks a pixel buffer the first time it is seen, so that frames drain by DMA (rxc_pin_host).
    pub fn pin_once(&mut self, pixels: &mut [u8]) {
        let key = (pixels.as_mut_ptr(), pixels.len());
        if !self.pinned.contains(&key) {
            if unsafe { rxc_pin_host(self.ctx, key.0 as *mut c_void, key.1 as u64) } == RXC_OK { self.pinned.push(key); }
        }
    }

    pub fn sync_assets(&mut self, assets: &Assets) {
        let key = (assets.tile_list.as_ptr() as usize, assets.tile_list.len());
        if key == self.assets_key { return; }
        let (tiles, _keep) = marshal_tiles(assets.tile_list.iter());
        let st = unsafe { rxc_set_assets(self.ctx, tiles.as_ptr(), tiles.len() as u32) };
        self.check(st, "rxc_set_assets");
        self.assets_key = key;
        self.scene_key = 0;
    }

    pub fn sync_mapmini(&mut self, mm: &MapMini) {
        let linedefs: Vec<rxc_linedef> = mm.linedefs.iter().chain(mm.dynamic_linedefs.iter())
            .map(|l| rxc_linedef { start: [l.start.x, l.start.y], end: [l.end.x, l.end.y] }).collect();
        let sectors: Vec<rxc_sector> = mm.occluded_sectors.iter().map(sector).collect();
        let m = rxc_mapmini { linedefs: linedefs.as_ptr(), n_linedefs: linedefs.len() as u32, occluded_sectors: sectors.as_ptr(),
                              n_occluded_sectors: sectors.len() as u32 };
        let st = unsafe { rxc_set_mapmini(self.ctx, &m) };
        self.check(st, "rxc_set_mapmini");
    }

    /// Flattens the scene in the reference's SUBMISSION ORDER (src/rasterizer.rs:314-405, :501-553) and uploads it
    /// when its structure changed; when only the lights moved, rxc_set_lights.  Returns false when a batch shader
    /// cannot run on the device (the caller then uses the CPU `rasterize`).
    pub fn sync_scene(&mut self, scene: &Scene, assets: &Assets) -> bool {
        let lights: Vec<rxc_light> = scene.lights.iter().chain(scene.dynamic_lights.iter()).map(light).collect();
        let lkey = hash_bytes(unsafe { std::slice::from_raw_parts(lights.as_ptr() as *const u8, lights.len() * std::mem::size_of::<rxc_light>()) });
        let skey = scene_key(scene);
        if skey == self.scene_key {
            if lkey != self.lights_key {
                let st = unsafe { rxc_set_lights(self.ctx, lights.as_ptr(), lights.len() as u32) };
                self.check(st, "rxc_set_lights");
                self.lights_key = lkey;
            }
            return true;
        }
        // ---- programs: scene.shaders, then every chunk's (rxc_chunk.shader_base)
        let mut flat: Vec<FlatProgram> = Vec::new();
        for p in &scene.shaders { match lower_program(p) { Ok(f) => flat.push(f), Err(_) => return false } }
        let n_scene_shaders = flat.len() as u32;
        let mut chunk_base = Vec::new();
        for chunk in scene.chunks.values() {
            chunk_base.push(flat.len() as u32);
            for p in &chunk.shaders { match lower_program(p) { Ok(f) => flat.push(f), Err(_) => return false } }
        }
        let programs: Vec<rxc_program> = flat.iter().map(|f| rxc_program {
            code: if f.words.is_empty() { std::ptr::null() } else { f.words.as_ptr() }, n_words: f.words.len() as u32, entry: f.entry,
            shade_locals: f.shade_locals, n_globals: f.n_globals, sets_opacity: f.sets_opacity as u32 }).collect();
        // ---- batches
        let mut actors = ActorTiles { assets, tiles: Vec::new() };
        let mut b3: Vec<rxc_batch3d> = Vec::new();
        let mut b2: Vec<rxc_batch2d> = Vec::new();
        for (ci, chunk) in scene.chunks.values().enumerate() {
            for b in &chunk.batches3d_opacity { b3.push(batch3d(b, RXC_PASS_CHUNK_OPACITY, ci as i32, &mut actors)); }
            for b in &chunk.batches3d { b3.push(batch3d(b, RXC_PASS_CHUNK, ci as i32, &mut actors)); }
            if let Some(b) = &chunk.terrain_batch3d { b3.push(batch3d(b, RXC_PASS_CHUNK, ci as i32, &mut actors)); }
            for b in &chunk.batches2d { b2.push(batch2d(b, ci as i32, &mut actors)); }
            if let Some(b) = &chunk.terrain_batch2d { b2.push(batch2d(b, ci as i32, &mut actors)); }
        }
        for b in &scene.d3_static { b3.push(batch3d(b, RXC_PASS_STATIC, -1, &mut actors)); }
        for b in &scene.d3_dynamic { b3.push(batch3d(b, RXC_PASS_DYNAMIC, -1, &mut actors)); }
        for b in &scene.d3_overlay { b3.push(batch3d(b, RXC_PASS_OVERLAY, -1, &mut actors)); }
        for b in &scene.d2_static { b2.push(batch2d(b, -1, &mut actors)); }
        for b in &scene.d2_dynamic { b2.push(batch2d(b, -1, &mut actors)); }
        // ---- chunks
        let sector_lists: Vec<Vec<rxc_sector>> = scene.chunks.values().map(|c| c.occluded_sectors.iter().map(sector).collect()).collect();
        let terrain: Vec<Option<rxc_texture>> = scene.chunks.values().map(|c| c.terrain_texture.as_ref().map(texture)).collect();
        let baked: Vec<Vec<Option<rxc_texture>>> = scene.chunks.values().map(|c| c.shader_textures.iter().map(|t| t.as_ref().map(texture)).collect()).collect();
        let baked_ptrs: Vec<Vec<*const rxc_texture>> = baked.iter().map(|v| v.iter().map(|t| t.as_ref().map_or(std::ptr::null(), |t| t as *const _)).collect()).collect();
        let chunks: Vec<rxc_chunk> = scene.chunks.values().enumerate().map(|(i, c)| chunk(c, &sector_lists[i], &terrain[i], &baked_ptrs[i], chunk_base[i])).collect();
        // ---- textures, VM banks, palette
        let (dyn_tiles, _k1) = marshal_tiles(scene.dynamic_textures.iter());
        let (actor_tiles, _k2) = marshal_tiles(actors.tiles.iter().copied());
        let pats: Vec<rxc_pattern> = rusteria::textures::patterns::patterns().iter().map(pattern).collect();
        let pats_n: Vec<rxc_pattern> = rusteria::textures::patterns::patterns_normal().iter().map(pattern).collect();
        let palette: Vec<[f32; 4]> = assets.palette.colors.iter().map(|c| match c { Some(c) => { let v = c.to_vec3(); [1.0, v.x, v.y, v.z] } None => [0.0; 4] }).collect();
        let s = rxc_scene {
            batches3d: b3.as_ptr(), n_batches3d: b3.len() as u32, batches2d: b2.as_ptr(), n_batches2d: b2.len() as u32,
            lights: lights.as_ptr(), n_lights: lights.len() as u32,
            dynamic_textures: dyn_tiles.as_ptr(), n_dynamic_textures: dyn_tiles.len() as u32,
            chunks: chunks.as_ptr(), n_chunks: chunks.len() as u32,
            actor_tiles: actor_tiles.as_ptr(), n_actor_tiles: actor_tiles.len() as u32,
            shaders: programs.as_ptr(), n_shaders: programs.len() as u32, n_scene_shaders,
            patterns: pats.as_ptr(), n_patterns: pats.len() as u32, patterns_normal: pats_n.as_ptr(), n_patterns_normal: pats_n.len() as u32,
            palette: palette.as_ptr() as *const f32, n_palette: palette.len() as u32,
        };
        // An engine's frame loop replaces the dynamic batches and keeps the world: the leading 3D batches whose arrays are the ones
        // already resident (same pointers and lengths as at the last upload) are not validated, flattened or uploaded again.
        let gkeys: Vec<(usize, usize, usize, usize)> = b3.iter().map(|b| (b.vertices as usize, b.indices as usize, b.n_vertices as usize, b.n_triangles as usize)).collect();
        let keep = if self.scene_key != 0 { gkeys.iter().zip(self.geometry_keys.iter()).take_while(|(a, b)| a == b).count() } else { 0 };
        let mut st = if keep > 0 { unsafe { rxc_update_scene(self.ctx, &s, keep as u32) } } else { RXC_ERR_INVALID };
        if st == RXC_ERR_INVALID { st = unsafe { rxc_set_scene(self.ctx, &s) }; }   // nothing to keep, or the library no longer holds that prefix
        if st == RXC_ERR_UNSUPPORTED { return false; }
        self.check(st, "rxc_set_scene");
        self.geometry_keys = gkeys;
        self.scene_key = skey;
        self.lights_key = lkey;
        true
    }
}

impl Drop for CudaRasterizer {
    fn drop(&mut self) {
        for (p, _) in &self.pinned { unsafe { rxc_unpin_host(self.ctx, *p as *mut c_void) }; }
        unsafe { rxc_destroy(self.ctx) };
    }
}

impl Rasterizer {
    /// Same arguments and result as `rasterize`; the work happens on the GPU owned by `cuda`.  Falls back to the CPU
    /// path itself when the library reports RXC_ERR_UNSUPPORTED (the library never does).
    pub fn rasterize_cuda(&mut self, cuda: &mut CudaRasterizer, scene: &mut Scene, pixels: &mut [u8], width: usize, height: usize,
                          tile_size: usize, assets: &Assets) {
        assert!(pixels.len() >= width * height * 4);              // the reference panics on a short slice (:572)
        self.width = width as f32;                                // :194-195
        self.height = height as f32;
        let chunk_lights: Vec<CompiledLight> = scene.chunks.values().flat_map(|c| c.lights.iter().cloned()).collect();
        scene.dynamic_lights.extend(chunk_lights);                // :219-223, a host-side effect the reference has too
        // render graph (:227-253): stays on the host
        self.render_hit = self.render_graph.collect_nodes_from(0, 0);
        self.render_miss = self.render_graph.collect_nodes_from(0, 1);
        let mut sky: Option<[[f32; 4]; 6]> = None;
        for node in self.render_miss.clone() {
            let n = &mut self.render_graph.nodes[node as usize];
            if let Some((sun_dir, day_factor)) = n.render_setup(self.hour) {
                self.sun_dir = Some(sun_dir);
                self.day_factor = day_factor;
                let mut block = [[0.0f32; 4]; 6];
                for (i, v) in n.precomputed.iter().take(6).enumerate() { block[i] = v.into_array(); }
                sky = Some(block);                                // the last Sky node wins, like in render_miss_d3
            }
        }
        for node in self.render_miss.clone() {
            if let Some(a) = self.render_graph.nodes[node as usize].render_ambient_color(self.hour) { self.ambient_color = Some(a); }
        }
        cuda.pin_once(pixels);
        cuda.sync_assets(assets);
        if !cuda.sync_scene(scene, assets) { return self.rasterize(scene, pixels, width, height, tile_size, assets); }
        cuda.sync_mapmini(&self.mapmini);
        let (bg_kind, grid) = background_kind(scene);
        let brush = self.brush_preview.as_ref();
        let f = rxc_frame {
            view: self.view_matrix.into_col_array(),
            projection: self.projection_matrix.into_col_array(),
            inverse_view: self.inverse_view_matrix.into_col_array(),       // vek's own inverses: bit-identical to the CPU path
            inverse_projection: self.inverse_projection_matrix.into_col_array(),
            has_matrix2d: self.projection_matrix_2d.is_some() as u32,
            matrix2d: self.projection_matrix_2d.map(|m| m.into_col_array()).unwrap_or([0.0; 9]),
            width: width as u32, height: height as u32, tile_size: tile_size as u32,
            sample_mode: match self.sample_mode { SampleMode::Nearest => RXC_SAMPLE_NEAREST, SampleMode::Linear => RXC_SAMPLE_LINEAR },
            has_background_color: self.background_color.is_some() as u32,
            background_color: self.background_color.unwrap_or([0; 4]),
            background_shader: bg_kind, grid_size: grid[0], grid_subdivisions: grid[1], grid_offset: [grid[2], grid[3]],
            has_ambient: self.ambient_color.is_some() as u32,
            ambient: self.ambient_color.map(|a| a.into_array()).unwrap_or([0.0; 4]),
            animation_frame: scene.animation_frame as u64,
            time: self.time, hour: self.hour,
            d2_active: self.render_mode.d2_active as u32, d3_active: self.render_mode.d3_active as u32,
            ignore_background_shader: self.render_mode.ignore_background_shader as u32,
            preserve_transparency: self.preserve_transparency as u32,
            matvec_mode: RXC_MATVEC_FMA_COLUMNS,
            band_y0: 0, band_y1: 0, band_x0: 0, band_x1: 0,
            has_sun: self.sun_dir.is_some() as u32,
            sun_dir: self.sun_dir.map(|d| d.into_array()).unwrap_or([0.0; 3]),
            day_factor: self.day_factor,
            has_sky: sky.is_some() as u32, sky: sky.unwrap_or([[0.0; 4]; 6]),
            sky_clouds: 1,        // the reference's Sky node always draws its cloud layer: the library answers UNSUPPORTED and the CPU path runs
            has_brush_preview: brush.is_some() as u32,
            brush_position: brush.map(|b| b.position.into_array()).unwrap_or([0.0; 3]),
            brush_radius: brush.map_or(0.0, |b| b.radius), brush_falloff: brush.map_or(0.0, |b| b.falloff),
        };
        let st = unsafe { rxc_rasterize(cuda.ctx, &f, pixels.as_mut_ptr(), std::ptr::null_mut(), std::ptr::null_mut()) };
        if st == RXC_ERR_UNSUPPORTED { return self.rasterize(scene, pixels, width, height, tile_size, assets); }
        cuda.check(st, "rxc_rasterize");
    }
}

// ---- marshalling helpers ------------------------------------------------------------------------------------------
fn texture(t: &Texture) -> rxc_texture { rxc_texture { data: t.data.as_ptr(), width: t.width as u32, height: t.height as u32 } }

fn marshal_tiles<'a>(tiles: impl Iterator<Item = &'a Tile>) -> (Vec<rxc_tile>, Vec<Vec<rxc_texture>>) {
    let keep: Vec<Vec<rxc_texture>> = tiles.map(|t| t.textures.iter().map(texture).collect()).collect();
    let out = keep.iter().map(|v| rxc_tile { textures: v.as_ptr(), n_textures: v.len() as u32 }).collect();
    (out, keep)
}

fn sector(s: &(crate::BBox, f32)) -> rxc_sector { rxc_sector { min: [s.0.min.x, s.0.min.y], max: [s.0.max.x, s.0.max.y], occlusion: s.1 } }

fn pattern(p: &rusteria::textures::TexStorage) -> rxc_pattern {
    rxc_pattern { data: p.data.as_ptr() as *const f32, width: p.width as u32, height: p.height as u32 }   // Value = Vec3<f32>, repr(C)
}

fn light(l: &CompiledLight) -> rxc_light {
    rxc_light {
        light_type: match l.light_type { LightType::Point => RXC_LIGHT_POINT, LightType::Ambient => RXC_LIGHT_AMBIENT, LightType::AmbientDaylight => RXC_LIGHT_AMBIENT_DAYLIGHT,
                                         LightType::Spot => RXC_LIGHT_SPOT, LightType::Area => RXC_LIGHT_AREA, LightType::Daylight => RXC_LIGHT_DAYLIGHT },
        position: l.position.into_array(), color: l.color, intensity: l.intensity, emitting: l.emitting as u32,
        start_distance: l.start_distance, end_distance: l.end_distance, flicker: l.flicker, direction: l.direction.into_array(),
        cone_angle: l.cone_angle, normal: l.normal.into_array(), width: l.width, height: l.height, from_linedef: l.from_linedef as u32,
    }
}

struct ActorTiles<'a> { assets: &'a Assets, tiles: Vec<&'a Tile> }
impl<'a> ActorTiles<'a> {
    /// assets.entity_tiles[id].get_index(index) (src/rasterizer.rs:1130-1177) -> index into rxc_scene.actor_tiles
    fn resolve(&mut self, entity: bool, id: u32, index: u32) -> u32 {
        let table = if entity { &self.assets.entity_tiles } else { &self.assets.item_tiles };
        let Some((_, tile)) = table.get(&id).and_then(|seq| seq.get_index(index as usize)) else { return 0xFFFF_FFFF };
        if tile.textures.is_empty() { return 0xFFFF_FFFF; }
        if let Some(k) = self.tiles.iter().position(|t| std::ptr::eq(*t, tile)) { return k as u32; }
        self.tiles.push(tile);
        (self.tiles.len() - 1) as u32
    }
}

fn source(src: &PixelSource, actors: &mut ActorTiles) -> (u32, u32, [u8; 4]) {
    match src {
        PixelSource::StaticTileIndex(i) => (RXC_SRC_STATIC_TILE, *i as u32, [0; 4]),
        PixelSource::DynamicTileIndex(i) => (RXC_SRC_DYNAMIC_TILE, *i as u32, [0; 4]),
        PixelSource::Pixel(p) => (RXC_SRC_PIXEL, 0, *p),
        PixelSource::EntityTile(id, i) => (RXC_SRC_ENTITY_TILE, actors.resolve(true, *id, *i), [0; 4]),
        PixelSource::ItemTile(id, i) => (RXC_SRC_ITEM_TILE, actors.resolve(false, *id, *i), [0; 4]),
        PixelSource::Terrain => (RXC_SRC_TERRAIN, 0, [0; 4]),
        _ => (RXC_SRC_OTHER, 0, [0; 4]),     // Off / TileId / MaterialId / Sequence / Color / ShapeFXGraphId (:1221, :757)
    }
}

fn mode(m: &PrimitiveMode) -> u32 {
    match m { PrimitiveMode::Lines => RXC_MODE_LINES, PrimitiveMode::LineStrip => RXC_MODE_LINE_STRIP, PrimitiveMode::LineLoop => RXC_MODE_LINE_LOOP, _ => RXC_MODE_TRIANGLES }
}
fn repeat(m: &RepeatMode) -> u32 {
    match m { RepeatMode::ClampXY => RXC_REPEAT_CLAMP_XY, RepeatMode::RepeatXY => RXC_REPEAT_REPEAT_XY, RepeatMode::RepeatX => RXC_REPEAT_REPEAT_X, RepeatMode::RepeatY => RXC_REPEAT_REPEAT_Y }
}

fn batch3d(b: &Batch3D, pass: u32, chunk: i32, actors: &mut ActorTiles) -> rxc_batch3d {
    let (source_kind, source_index, source_pixel) = source(&b.source, actors);
    rxc_batch3d {
        vertices: b.vertices.as_ptr() as *const f32, uvs: b.uvs.as_ptr() as *const f32,
        normals: if b.normals.is_empty() { std::ptr::null() } else { b.normals.as_ptr() as *const f32 },   // vek Vec3 is repr(C)
        indices: b.indices.as_ptr() as *const c_void, n_vertices: b.vertices.len() as u32, n_triangles: b.indices.len() as u32,
        index_bytes: std::mem::size_of::<usize>() as u32,          // (usize, usize, usize) triples, 24 B per triangle
        mode: mode(&b.mode), repeat_mode: repeat(&b.repeat_mode),
        cull_mode: match b.cull_mode { CullMode::Off => RXC_CULL_OFF, CullMode::Front => RXC_CULL_FRONT, CullMode::Back => RXC_CULL_BACK },
        source_kind, source_index, source_pixel, receives_light: b.receives_light as u32, ambient_color: b.ambient_color.into_array(),
        has_profile_id: b.profile_id.is_some() as u32, profile_id: b.profile_id.unwrap_or(0),
        shader: b.shader.map_or(-1, |s| s as i32), pass, transform: b.transform_3d.into_col_array(), chunk,
    }
}

fn batch2d(b: &Batch2D, chunk: i32, actors: &mut ActorTiles) -> rxc_batch2d {
    let (source_kind, source_index, source_pixel) = source(&b.source, actors);
    rxc_batch2d {
        vertices: b.vertices.as_ptr() as *const f32, uvs: b.uvs.as_ptr() as *const f32, indices: b.indices.as_ptr() as *const c_void,
        n_vertices: b.vertices.len() as u32, n_triangles: b.indices.len() as u32, index_bytes: std::mem::size_of::<usize>() as u32,
        mode: mode(&b.mode), repeat_mode: repeat(&b.repeat_mode), source_kind, source_index, source_pixel,
        receives_light: b.receives_light as u32, shader: b.shader.map_or(-1, |s| s as i32), chunk,
    }
}

fn chunk(c: &Chunk, sectors: &[rxc_sector], terrain: &Option<rxc_texture>, baked: &[*const rxc_texture], shader_base: u32) -> rxc_chunk {
    rxc_chunk {
        origin: [c.origin.x, c.origin.y], size: c.size, occluded_sectors: sectors.as_ptr(), n_occluded_sectors: sectors.len() as u32,
        terrain_texture: terrain.as_ref().map_or(std::ptr::null(), |t| t as *const _), shader_base, n_shaders: c.shaders.len() as u32,
        shader_textures: if baked.iter().any(|p| !p.is_null()) { baked.as_ptr() } else { std::ptr::null() },
    }
}

/// scene.background: the two shaders of the crate are recognised through `Shader::name()`-less downcasting being
/// unavailable on `dyn Shader`, so the maintainer adds `fn kind(&self) -> u32 { 0 }` to the trait (1 for
/// VGrayGradientShader, 2 for GridShader, whose parameters it returns through `fn params(&self) -> [f32; 4]`).
fn background_kind(scene: &Scene) -> (u32, [f32; 4]) {
    match &scene.background { Some(s) => (s.kind(), s.params()), None => (RXC_BG_NONE, [0.0; 4]) }
}

fn hash_bytes(b: &[u8]) -> u64 { b.iter().fold(0xcbf29ce484222325u64, |h, x| (h ^ *x as u64).wrapping_mul(0x100000001b3)) }

/// Which Vecs the scene currently holds (pointer + length of every geometry array, in submission order): the reference
/// re-projects `&mut scene` on every call, so anything that moved or grew must be seen.  In-place edits of vertex data
/// need `CudaRasterizer::invalidate()` (or a generation counter on Batch3D, the maintainer's choice).
fn scene_key(scene: &Scene) -> u64 {
    let mut h = 0xcbf29ce484222325u64;
    let mut mix = |p: usize, n: usize| { for x in [p as u64, n as u64] { h = (h ^ x).wrapping_mul(0x100000001b3); } };
    let mut b3 = |b: &Batch3D| { mix(b.vertices.as_ptr() as usize, b.vertices.len()); mix(b.indices.as_ptr() as usize, b.indices.len()); };
    for c in scene.chunks.values() { for b in c.batches3d_opacity.iter().chain(c.batches3d.iter()).chain(c.terrain_batch3d.iter()) { b3(b); } }
    for b in scene.d3_static.iter().chain(scene.d3_dynamic.iter()).chain(scene.d3_overlay.iter()) { b3(b); }
    for c in scene.chunks.values() { for b in c.batches2d.iter().chain(c.terrain_batch2d.iter()) { mix(b.vertices.as_ptr() as usize, b.vertices.len()); } }
    for b in scene.d2_static.iter().chain(scene.d2_dynamic.iter()) { mix(b.vertices.as_ptr() as usize, b.vertices.len()); }
    mix(scene.dynamic_textures.as_ptr() as usize, scene.dynamic_textures.len());
    mix(scene.shaders.as_ptr() as usize, scene.shaders.len());
    h | 1
}

impl CudaRasterizer {
    /// Forget what is resident (call after editing vertex or texture data in place).
    pub fn invalidate(&mut self) { self.scene_key = 0; self.assets_key = (0, 0); self.geometry_keys.clear(); }

    /// How the library recompiles its raster kernel for the scene (NVRTC): 0 = never (generic kernels, batch shaders
    /// interpreted), 1 = in the background (default), 2 = before the first frame that needs the kernel (offline renderers, benchmarks).
    /// Takes effect with the next scene upload.
    pub fn set_kernel_jit(&mut self, mode: i32) {
        let st = unsafe { rxc_set_vm_jit(self.ctx, mode) };
        self.check(st, "rxc_set_vm_jit");
        self.scene_key = 0;
    }

    /// 0 = a fresh `Execution` per fragment (fast, default); 1 = the reference's order and never-reset per-tile `Execution` for every
    /// scene with shader programs; 2 = for the scenes whose `shader_state_report` flags a program (every frame then equals the reference's).
    pub fn set_shader_state_mode(&mut self, mode: i32) {
        let st = unsafe { rxc_set_vm_state_mode(self.ctx, mode) };
        self.check(st, "rxc_set_vm_state_mode");
    }

    /// Per shader program of the resident scene: 0 = device and reference agree by construction, 1 = the program can observe the
    /// reference's never-reset per-tile `Execution` (src/rasterizer.rs:310), 2 = not analysable.
    pub fn shader_state_report(&self) -> Vec<u32> {
        let mut n: u32 = 0;
        unsafe { rxc_vm_scene_state_report(self.ctx, std::ptr::null_mut(), 0, &mut n) };
        let mut out = vec![0u32; n as usize];
        unsafe { rxc_vm_scene_state_report(self.ctx, out.as_mut_ptr(), n, &mut n) };
        out
    }
}
