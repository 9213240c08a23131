//! src/backend.rs — added to the `shimmer` crate (jalberse/RayTracingInOneWeekendInRust) by shimmer-b200.patch: routes
//! `Renderer::render` through libshimmer_b200.  Nothing a user of the crate touches changes: the Hittable / Material /
//! Texture traits, every constructor, `Camera::new`, `Renderer::from_aspect_ratio`, the clap CLI and the PPM on
//! stdout stay.
//!
//! Each trait gains ONE provided method, `record`, through which an implementor describes itself to a `Recorder`
//! (one C call per reference constructor, include/shimmer_b200.h); the patch overrides it for every implementor the
//! crate ships.  Trait objects never cross the FFI boundary; a user-defined implementor keeps the default, which
//! returns `Err(unsupported)`, and `render` reports it (the backend has no CPU fallback; SHIMMER_CPU=1 keeps the
//! crate's rayon loop).
//!
//! NOTE: written against the reference sources without a Rust toolchain at hand (none in the build image): the patch
//! is checked to APPLY (tests/test_rust_binding.py), it has never been compiled.
use std::collections::{HashMap, HashSet};
use std::sync::{Arc, Mutex};

use ahash::AHashMap;
use glam::Vec3;
use shimmer_b200::{Error, HittableId, MaterialId, RenderParams, Result, Scene, TextureId};

use crate::bvh::BvhId;
use crate::camera::Camera;
use crate::hittable::{Hittable, HittableList};
use crate::hrpp::Predictor;
use crate::materials::material::Material;
use crate::textures::texture::Texture;

/// Wraps the scene handle and remembers which `Arc`s were already recorded (shared materials / textures / instanced
/// hittables are described once; the address of the Arc's payload is the identity).
pub struct Recorder {
    pub scene: Scene,
    textures: HashMap<usize, TextureId>,
    materials: HashMap<usize, MaterialId>,
    hittables: HashMap<usize, HittableId>,
    /// ids of the BVHs that were built by `Bvh::with_predictor` (the keys of render's `predictors` argument)
    pub predictor_bvhs: HashSet<BvhId>,
}

fn key<T: ?Sized>(a: &Arc<T>) -> usize {
    Arc::as_ptr(a) as *const () as usize
}

pub fn unsupported(what: &str) -> Error {
    Error { code: shimmer_b200::sys::SHIM_ERR_UNSUPPORTED, message: format!("{what} cannot run on the device backend") }
}

pub fn v3(v: Vec3) -> [f32; 3] {
    [v.x, v.y, v.z]
}

impl Recorder {
    pub fn new(predictor_bvhs: HashSet<BvhId>) -> Result<Recorder> {
        Ok(Recorder { scene: Scene::new()?, textures: HashMap::new(), materials: HashMap::new(), hittables: HashMap::new(), predictor_bvhs })
    }
    pub fn texture(&mut self, t: &Arc<dyn Texture>) -> Result<TextureId> {
        if let Some(id) = self.textures.get(&key(t)) {
            return Ok(*id);
        }
        let id = t.record(self)?;
        self.textures.insert(key(t), id);
        Ok(id)
    }
    pub fn material(&mut self, m: &Arc<dyn Material>) -> Result<MaterialId> {
        if let Some(id) = self.materials.get(&key(m)) {
            return Ok(*id);
        }
        let id = m.record(self)?;
        self.materials.insert(key(m), id);
        Ok(id)
    }
    pub fn hittable(&mut self, h: &Arc<dyn Hittable>) -> Result<HittableId> {
        if let Some(id) = self.hittables.get(&key(h)) {
            return Ok(*id);
        }
        let id = h.record(self)?;
        self.hittables.insert(key(h), id);
        Ok(id)
    }
}

/// The body of `Renderer::render` (renderer.rs:42-105): same arguments, same PPM on stdout, same `io::Result`.
pub fn render_on_device(
    image_width: usize,
    image_height: usize,
    camera: &Camera,
    world: &HittableList,
    background: Vec3,
    samples_per_pixel: u32,
    max_depth: u32,
    tile_width: usize,
    tile_height: usize,
    predictors: &Option<AHashMap<BvhId, Mutex<Predictor>>>,
) -> std::io::Result<()> {
    let predictor_bvhs: HashSet<BvhId> = match predictors {
        Some(map) => map.keys().copied().collect(),
        None => HashSet::new(),
    };
    let with_predictors = !predictor_bvhs.is_empty();
    let mut rec = Recorder::new(predictor_bvhs)?;
    for obj in &world.objects {
        let id = rec.hittable(obj)?;
        rec.scene.world_add(id)?; // list order is the tie-break order of HittableList::hit (hittable.rs:100-118)
    }
    rec.scene.use_reference_bvh_on_device(with_predictors)?; // HRPP stores the crate's own leaf indices
    rec.scene.commit()?; // flatten to SoA + one H2D copy
    let p = RenderParams {
        width: image_width as i32,
        height: image_height as i32,
        samples_per_pixel: samples_per_pixel as i32,
        max_depth: max_depth as i32,
        tile_width: tile_width as i32,
        tile_height: tile_height as i32,
        background: v3(background),
        seed: rand::random(),
        sample_begin: 0,
        sample_count: 0,
        tile_rank: 0,
        tile_world: 0,
        flags: if with_predictors { shimmer_b200::sys::SHIM_RENDER_PREDICTORS } else { 0 },
        pool_paths: 0,
    };
    let mut fb = shimmer_b200::HostFramebuffer::new(image_width, image_height)?; // ImageColors, page-locked
    let stats = rec.scene.render(&camera.to_pod(), &p, &mut fb)?;
    eprintln!("{} rays, {} samples, {:.3} ms on the device", stats.rays, stats.samples, stats.device_ms);
    if with_predictors {
        // hrpp.rs:85-130 prints these on Drop
        eprintln!(
            "HRPP true positives {} false positives {} no prediction {}",
            stats.hrpp_true_positive, stats.hrpp_false_positive, stats.hrpp_no_prediction
        );
    }
    shimmer_b200::write_ppm(&fb, image_width, image_height, None)?; // identical P3 text (renderer.rs:107-127)
    Ok(())
}
