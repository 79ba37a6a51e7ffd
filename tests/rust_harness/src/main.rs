// Reads scene dumps (tools/rust_harness_export.py, format RXH1), renders each with the reference's own
// `Rasterizer::setup(..).rasterize(..)` and writes <case>.rxo: the pixels, what `Scene::project` left in every
// Batch3D (projected_vertices, clipped_indices, edge visibility, bounding_box) and vek's answers to the
// Mat4*Vec4 / Mat4*Mat4 probes of the dump.  tests/test_rust_reference.py compares those with the CPU oracle.
use rusterix::prelude::*;
use std::io::{Read, Write};
use vek::{Mat4, Vec3, Vec4};

struct In<'a>(&'a [u8], usize);
impl<'a> In<'a> {
    fn u32(&mut self) -> u32 { let v = u32::from_le_bytes(self.0[self.1..self.1 + 4].try_into().unwrap()); self.1 += 4; v }
    fn f32(&mut self) -> f32 { f32::from_bits(self.u32()) }
    fn f32s(&mut self, n: usize) -> Vec<f32> { (0..n).map(|_| self.f32()).collect() }
    fn bytes(&mut self, n: usize) -> Vec<u8> { let v = self.0[self.1..self.1 + n].to_vec(); self.1 += n; v }
    fn mat4(&mut self) -> Mat4<f32> { let m = self.f32s(16); Mat4::from_col_array(m.try_into().unwrap()) }
}
fn put_u32(o: &mut Vec<u8>, v: u32) { o.extend_from_slice(&v.to_le_bytes()); }
fn put_f32(o: &mut Vec<u8>, v: f32) { o.extend_from_slice(&v.to_bits().to_le_bytes()); }

fn source(kind: u32, index: u32, px: [u8; 4]) -> PixelSource {
    match kind { 1 => PixelSource::StaticTileIndex(index as u16), 2 => PixelSource::DynamicTileIndex(index as u16),
                 3 => PixelSource::Pixel(px), _ => PixelSource::Off }
}

fn run(path: &str) {
    let mut data = Vec::new();
    std::fs::File::open(path).unwrap().read_to_end(&mut data).unwrap();
    let mut r = In(&data, 0);
    assert_eq!(&r.bytes(4), b"RXH1");
    let (w, h, ts, sample) = (r.u32() as usize, r.u32() as usize, r.u32() as usize, r.u32());
    let has_ambient = r.u32();
    let ambient = r.f32s(4);
    let (view, proj) = (r.mat4(), r.mat4());
    let tiles: Vec<Tile> = (0..r.u32()).map(|_| { let (tw, th) = (r.u32() as usize, r.u32() as usize);
        Tile::from_texture(Texture::new(r.bytes(tw * th * 4), tw, th)) }).collect();
    let lights: Vec<_> = (0..r.u32()).map(|_| { let ty = r.u32(); let v = r.f32s(19); let (emit, lin) = (r.u32(), r.u32());
        let mut l = Light::new(match ty { 1 => LightType::Ambient, 2 => LightType::AmbientDaylight, 3 => LightType::Spot,
                                          4 => LightType::Area, 5 => LightType::Daylight, _ => LightType::Point }).compile();
        l.position = Vec3::new(v[0], v[1], v[2]); l.color = [v[3], v[4], v[5]]; l.intensity = v[6];
        l.start_distance = v[7]; l.end_distance = v[8]; l.flicker = v[9]; l.direction = Vec3::new(v[10], v[11], v[12]);
        l.cone_angle = v[13]; l.normal = Vec3::new(v[14], v[15], v[16]); l.width = v[17]; l.height = v[18];
        l.emitting = emit != 0; l.from_linedef = lin != 0; l }).collect();
    let mut d3 = Vec::new();
    for _ in 0..r.u32() {
        let (nv, nt, repeat, cull, kind, index) = (r.u32() as usize, r.u32() as usize, r.u32(), r.u32(), r.u32(), r.u32());
        let px: [u8; 4] = r.bytes(4).try_into().unwrap();
        let (recv, has_n) = (r.u32(), r.u32());
        let transform = r.mat4();
        let verts: Vec<[f32; 4]> = (0..nv).map(|_| r.f32s(4).try_into().unwrap()).collect();
        let idx: Vec<(usize, usize, usize)> = (0..nt).map(|_| (r.u32() as usize, r.u32() as usize, r.u32() as usize)).collect();
        let uvs: Vec<[f32; 2]> = (0..nv).map(|_| r.f32s(2).try_into().unwrap()).collect();
        let mut b = Batch3D::new(verts, idx, uvs).source(source(kind, index, px)).transform(transform).receives_light(recv != 0)
            .repeat_mode(match repeat { 1 => RepeatMode::RepeatXY, 2 => RepeatMode::RepeatX, 3 => RepeatMode::RepeatY, _ => RepeatMode::ClampXY })
            .cull_mode(match cull { 1 => CullMode::Front, 2 => CullMode::Back, _ => CullMode::Off });
        if has_n != 0 { b.normals = (0..nv).map(|_| { let n = r.f32s(3); Vec3::new(n[0], n[1], n[2]) }).collect(); }
        d3.push(b);
    }
    let mut d2 = Vec::new();
    for _ in 0..r.u32() {
        let (nv, nt, kind, index) = (r.u32() as usize, r.u32() as usize, r.u32(), r.u32());
        let px: [u8; 4] = r.bytes(4).try_into().unwrap();
        let recv = r.u32();
        let verts: Vec<[f32; 2]> = (0..nv).map(|_| r.f32s(2).try_into().unwrap()).collect();
        let idx: Vec<(usize, usize, usize)> = (0..nt).map(|_| (r.u32() as usize, r.u32() as usize, r.u32() as usize)).collect();
        let uvs: Vec<[f32; 2]> = (0..nv).map(|_| r.f32s(2).try_into().unwrap()).collect();
        d2.push(Batch2D::new(verts, idx, uvs).source(source(kind, index, px)).receives_light(recv != 0));
    }
    let mut scene = Scene::from_static(d2, d3).lights(lights);
    let assets = Assets::default().textures(tiles);
    let mut pixels = vec![0u8; w * h * 4];
    let mut rast = Rasterizer::setup(None, view, proj).sample_mode(if sample == 1 { SampleMode::Linear } else { SampleMode::Nearest });
    if has_ambient != 0 { rast = rast.ambient(Vec4::new(ambient[0], ambient[1], ambient[2], ambient[3])); }
    rast.rasterize(&mut scene, &mut pixels[..], w, h, ts, &assets);

    let mut o = Vec::new();
    o.extend_from_slice(b"RXO1");
    put_u32(&mut o, w as u32); put_u32(&mut o, h as u32);
    o.extend_from_slice(&pixels);
    put_u32(&mut o, scene.d3_static.len() as u32);
    for b in &scene.d3_static {
        match &b.bounding_box { Some(bb) => { put_u32(&mut o, 1); for v in [bb.x, bb.y, bb.width, bb.height] { put_f32(&mut o, v); } }
                                None => { put_u32(&mut o, 0); for _ in 0..4 { put_f32(&mut o, 0.0); } } }
        put_u32(&mut o, b.projected_vertices.len() as u32);
        for p in &b.projected_vertices { for v in p { put_f32(&mut o, *v); } }
        put_u32(&mut o, b.clipped_indices.len() as u32);
        for t in &b.clipped_indices { put_u32(&mut o, t.0 as u32); put_u32(&mut o, t.1 as u32); put_u32(&mut o, t.2 as u32); }
        for e in &b.edges { o.push(e.visible as u8); }
    }
    // vek probes: n x (Mat4, Vec4) -> Mat4 * Vec4, then n x (Mat4, Mat4) -> Mat4 * Mat4, then n x Mat4 -> inverted()
    let n = r.u32();
    put_u32(&mut o, n);
    for _ in 0..n { let m = r.mat4(); let v = r.f32s(4); let q = m * Vec4::new(v[0], v[1], v[2], v[3]); for c in [q.x, q.y, q.z, q.w] { put_f32(&mut o, c); } }
    for _ in 0..n { let (a, b) = (r.mat4(), r.mat4()); for c in (a * b).into_col_array() { put_f32(&mut o, c); } }
    for _ in 0..n { let a = r.mat4(); for c in a.inverted().into_col_array() { put_f32(&mut o, c); } }
    let out = format!("{}.rxo", path.trim_end_matches(".rxh"));
    std::fs::File::create(&out).unwrap().write_all(&o).unwrap();
    println!("{} -> {}", path, out);
}

fn main() {
    for p in std::env::args().skip(1) { run(&p); }
}
