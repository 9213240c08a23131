//! src/backend.rs — add to the `shimmer` crate (jalberse/RayTracingInOneWeekendInRust) to route `Renderer::render`
//! through libshimmer_b200.  Nothing a user of the crate touches changes: the Hittable / Material / Texture traits,
//! every constructor, `Camera::new`, `Renderer::from_aspect_ratio`, the clap CLI and the PPM on stdout stay.
//!
//! How it works: each trait gains ONE required method, `record`, through which an implementor describes itself to
//! a `Recorder` (one C call per reference constructor, include/shimmer_b200.h).  Trait objects never cross the FFI
//! boundary; a user-defined implementor that cannot describe itself returns `Err(Unsupported)` and `render` reports
//! it (the backend has no CPU fallback; the crate may keep its rayon loop behind a cargo feature).
//!
//! NOTE: written against the reference sources without a Rust toolchain at hand (none in the build image) — compile
//! errors are possible; the C ABI underneath is what the test-suite exercises (through ctypes).
use std::collections::HashMap;
use std::sync::Arc;

use glam::Vec3;
use shimmer_b200::{Camera as CameraPod, Error, HittableId, MaterialId, RenderParams, Result, Scene, TextureId};

/// Wraps the scene handle and remembers which `Arc`s were already recorded (shared materials / textures / instanced
/// hittables are described once; the address of the Arc's payload is the identity).
pub struct Recorder {
    pub scene: Scene,
    textures: HashMap<usize, TextureId>,
    materials: HashMap<usize, MaterialId>,
    hittables: HashMap<usize, HittableId>,
    /// Marble::new draws its Perlin seed from thread_rng (marble.rs:14); the backend wants it explicit
    pub next_perlin_seed: u32,
    /// the `predictors: Option<..>` argument of render was `Some`: BVHs built by `with_predictor` get a device table
    pub with_predictors: bool,
}

fn key<T: ?Sized>(a: &Arc<T>) -> usize {
    Arc::as_ptr(a) as *const () as usize
}
pub fn unsupported(what: &str) -> Error {
    Error { code: shimmer_b200::sys::SHIM_ERR_UNSUPPORTED, message: format!("{what} cannot run on the device backend") }
}

impl Recorder {
    pub fn new(with_predictors: bool) -> Result<Recorder> {
        Ok(Recorder { scene: Scene::new()?, textures: HashMap::new(), materials: HashMap::new(), hittables: HashMap::new(),
                      next_perlin_seed: rand::random(), with_predictors })
    }
    pub fn texture(&mut self, t: &Arc<dyn crate::textures::texture::Texture>) -> Result<TextureId> {
        if let Some(id) = self.textures.get(&key(t)) { return Ok(*id); }
        let id = t.record(self)?;
        self.textures.insert(key(t), id);
        Ok(id)
    }
    pub fn material(&mut self, m: &Arc<dyn crate::materials::material::Material>) -> Result<MaterialId> {
        if let Some(id) = self.materials.get(&key(m)) { return Ok(*id); }
        let id = m.record(self)?;
        self.materials.insert(key(m), id);
        Ok(id)
    }
    pub fn hittable(&mut self, h: &Arc<dyn crate::hittable::Hittable>) -> Result<HittableId> {
        if let Some(id) = self.hittables.get(&key(h)) { return Ok(*id); }
        let id = h.record(self)?;
        self.hittables.insert(key(h), id);
        Ok(id)
    }
}
fn v3(v: Vec3) -> [f32; 3] { [v.x, v.y, v.z] }

// ------------------------------------------------------------------------------------------------- trait additions
// textures/texture.rs:      pub trait Texture: Send + Sync { fn value(..) -> Vec3;  fn record(&self, r: &mut Recorder) -> Result<TextureId>; }
// materials/material.rs:    pub trait Material: Send + Sync { fn scatter(..); fn emit(..); fn record(&self, r: &mut Recorder) -> Result<MaterialId>; }
// hittable.rs:              pub trait Hittable: Send + Sync { fn hit(..); fn bounding_box(..); fn record(&self, r: &mut Recorder) -> Result<HittableId>; }

// ------------------------------------------------------------------------------------------------- textures/*.rs
// impl Texture for SolidColor  (solid_color.rs:5-18)
//     fn record(&self, r: &mut Recorder) -> Result<TextureId> { r.scene.texture_solid(self.color.x, self.color.y, self.color.z) }
// impl Texture for Checker     (checker.rs:7-24)
//     fn record(&self, r: &mut Recorder) -> Result<TextureId> {
//         let (e, o) = (r.texture(&self.even)?, r.texture(&self.odd)?);
//         r.scene.texture_checker(self.scale, e, o)
//     }
// impl Texture for Marble      (marble.rs:7-20; keep the seed given to Perlin::new in a new field `seed: u32`)
//     fn record(&self, r: &mut Recorder) -> Result<TextureId> { r.scene.texture_marble(self.scale, self.seed) }
// impl Texture for ImageTexture (image_texture.rs:8-17)
//     fn record(&self, r: &mut Recorder) -> Result<TextureId> { r.scene.texture_image(self.image.as_raw(), self.image.width(), self.image.height()) }

// ------------------------------------------------------------------------------------------------- materials/*.rs
// impl Material for Lambertian   fn record(..) { let t = r.texture(&self.albedo)?; r.scene.lambertian(t) }
// impl Material for Metal        fn record(..) { r.scene.metal(v3(self.albedo), self.fuzz) }            // fuzz already clamped by Metal::new
// impl Material for Dialectric   fn record(..) { r.scene.dielectric(self.index_of_refraction) }
// impl Material for DiffuseLight fn record(..) { let t = r.texture(&self.emission_texture)?; r.scene.diffuse_light(t) }
// impl Material for Isotropic    fn record(..) { let t = r.texture(&self.albedo)?; r.scene.isotropic(t) }

// ------------------------------------------------------------------------------------------------- geometry/*.rs
// impl Hittable for Sphere        fn record(..) { let m = r.material(&self.material)?; r.scene.sphere(v3(self.center), self.radius, m) }
// impl Hittable for MovingSphere  fn record(..) { let m = r.material(&self.material)?;
//                                                 r.scene.moving_sphere(v3(self.center_start), v3(self.center_end), self.time_start, self.time_end, self.radius, m) }
// impl Hittable for XyRect        fn record(..) { let m = r.material(&self.material)?; r.scene.xy_rect(self.x0, self.x1, self.y0, self.y1, self.z, m) }
// impl Hittable for XzRect        fn record(..) { let m = r.material(&self.material)?; r.scene.xz_rect(self.x0, self.x1, self.z0, self.z1, self.y, m) }
// impl Hittable for YzRect        fn record(..) { let m = r.material(&self.material)?; r.scene.yz_rect(self.y0, self.y1, self.z0, self.z1, self.x, m) }
// impl Hittable for Tri           fn record(..) { let m = r.material(&self.material)?; r.scene.tri(v3(self.p0), v3(self.p1), v3(self.p2), m) }
// impl Hittable for Cube          (keep the material Arc given to Cube::new in a new field)
//                                 fn record(..) { let m = r.material(&self.material)?; r.scene.cube(v3(self.min_point), v3(self.max_point), m) }
// impl Hittable for Translate     fn record(..) { let h = r.hittable(&self.hittable)?; r.scene.translate(h, v3(self.displacement)) }
// impl Hittable for RotateY       (keep `degrees` in a new field)   fn record(..) { let h = r.hittable(&self.hittable)?; r.scene.rotate_y(h, self.degrees) }
// impl Hittable for ConstantMedium (keep `density` and the albedo texture Arc in new fields)
//                                 fn record(..) { let b = r.hittable(&self.boundary)?; let t = r.texture(&self.albedo)?; r.scene.constant_medium(b, self.density, t) }
// impl Hittable for HittableList  fn record(..) { let l = r.scene.list()?; for o in &self.objects { let h = r.hittable(o)?; r.scene.list_add(l, h)?; } Ok(l) }

/// bvh.rs — the crate keeps its own tree (random axis, median split, post-order indices, bvh.rs:249-333) and uploads
/// the topology, so the leaf indices HRPP stores are the reference's own.
pub fn record_bvh(nodes_left: &[BvhChild], nodes_right: &[BvhChild], root_index: usize, r: &mut Recorder, has_predictor: bool) -> Result<HittableId> {
    let enc = |c: &BvhChild, r: &mut Recorder| -> Result<i32> {
        Ok(match c {
            BvhChild::Index(i) => *i as i32,
            BvhChild::Hittable(h) => !r.hittable(h)?.0, // ~id
        })
    };
    let mut left = Vec::with_capacity(nodes_left.len());
    let mut right = Vec::with_capacity(nodes_right.len());
    for (l, rr) in nodes_left.iter().zip(nodes_right) {
        left.push(enc(l, r)?);
        right.push(enc(rr, r)?);
    }
    // time_0 / time_1 as given to Bvh::new (every scene in main.rs passes 0.0, 1.0)
    r.scene.bvh_from_nodes(&left, &right, root_index, 0.0, 1.0, has_predictor && r.with_predictors)
}
/// Mirror of the private `enum Child` (bvh.rs:31-34), exposed to this module.
pub enum BvhChild {
    Index(usize),
    Hittable(Arc<dyn crate::hittable::Hittable>),
}

// ------------------------------------------------------------------------------------------------- camera.rs
// Camera::new keeps its nine arguments next to the derived fields (camera.rs:44-54):
//     pub fn to_pod(&self) -> CameraPod { CameraPod { look_from: v3(self.look_from), look_at: v3(self.look_at), view_up: v3(self.view_up),
//         vertical_fov: self.vertical_fov, aspect_ratio: self.aspect_ratio, aperture: self.aperture, focus_dist: self.focus_dist,
//         time_start: self.time_start, time_end: self.time_end } }

// ------------------------------------------------------------------------------------------------- renderer.rs
/// The body of `Renderer::render` (renderer.rs:42-105): same signature, same PPM on stdout, same `io::Result`.
pub fn render_on_device(
    image_width: usize, image_height: usize, camera: &CameraPod, world: &crate::hittable::HittableList, background: Vec3,
    samples_per_pixel: u32, max_depth: u32, tile_width: usize, tile_height: usize, predictors_enabled: bool,
) -> std::io::Result<()> {
    let mut rec = Recorder::new(predictors_enabled)?;
    for obj in &world.objects {
        let id = rec.hittable(obj)?;
        rec.scene.world_add(id)?;            // list order is the tie-break order of HittableList::hit (hittable.rs:100-118)
    }
    rec.scene.commit()?;                     // flatten to SoA + one H2D copy
    let p = RenderParams {
        width: image_width as i32, height: image_height as i32, samples_per_pixel: samples_per_pixel as i32, max_depth: max_depth as i32,
        tile_width: tile_width as i32, tile_height: tile_height as i32, background: v3(background), seed: rand::random(),
        sample_begin: 0, sample_count: 0, tile_rank: 0, tile_world: 0,
        flags: if predictors_enabled { shimmer_b200::sys::SHIM_RENDER_PREDICTORS } else { 0 }, pool_paths: 0,
    };
    let mut fb = shimmer_b200::HostFramebuffer::new(image_width, image_height)?;   // ImageColors, page-locked
    let stats = rec.scene.render(camera, &p, &mut fb)?;
    eprintln!("{} rays, {} samples, {:.3} ms on the device", stats.rays, stats.samples, stats.device_ms);
    if predictors_enabled {                 // hrpp.rs:85-130 prints these on Drop
        eprintln!("HRPP true positives {} false positives {} no prediction {}", stats.hrpp_true_positive, stats.hrpp_false_positive, stats.hrpp_no_prediction);
    }
    shimmer_b200::write_ppm(&fb, image_width, image_height, None)?;                 // identical P3 text (renderer.rs:107-127)
    Ok(())
}
