// bindings/rust/b200.rs - the Rust side of include/rt_b200.h, as a maintainer of rust-tracing would add it (src/b200.rs).
//
// NOT COMPILED HERE: the build image has no cargo / rustc. What can be checked without a compiler is checked:
// tools/check_rust_abi.py parses every #[repr(C)] struct and every extern "C" declaration of this file and compares them
// with include/rt_b200.h field for field / parameter for parameter (tests/test_rust_binding.py runs it on every CPU run).
// The same entry points are exercised through the ctypes binding (rust-tracing_b200/_abi.py) by every GPU test.
// INTEGRATION.md says where each piece goes in the crate.
//
// How the crate's objects reach the device: the crate's Hittable / Material / Texture objects are Arc<dyn Trait> with private
// fields and no visitor, so each trait gets ONE new method, `flatten`, in which an object describes itself to a SceneBuilder
// and returns the id it was given (children first). Arcs shared by many objects (one ground material under 2400 quads,
// main.rs:511-533) are described once: SceneBuilder remembers the ids by Arc address.
#![allow(non_camel_case_types, dead_code)]

use std::collections::HashMap;
use std::os::raw::{c_char, c_double, c_float, c_int, c_void};
use std::sync::Arc;

use crate::bvh::{BVHNode, Node};
use crate::camera::Camera;
use crate::common::FP;
use crate::constant_medium::ConstantMedium;
use crate::hittable::{Hittable, HittableList, RotateY, Translate};
use crate::material::{Dielectric, DiffuseLight, Isotropic, Lambertian, Material, Metal};
use crate::quad::Quad;
use crate::sphere::Sphere;
use crate::texture::{CheckerTexture, ImageTexture, NoiseTexture, SolidColor, Texture};
use crate::vec3::{Color, Vec3};

// ------------------------------------------------------------------------------------------------ C structures
#[repr(C)] pub struct rt_builder { _p: [u8; 0] }
#[repr(C)] pub struct rt_context { _p: [u8; 0] }
#[repr(C)] pub struct rt_scene { _p: [u8; 0] }

#[repr(C)]
#[derive(Clone, Copy)]
pub struct rt_texture_desc {
    pub kind: i32,
    pub a: i32,
    pub b: i32,
    pub _pad: i32,
    pub color: [c_double; 3],
    pub scale: c_double,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct rt_material_desc {
    pub kind: i32,
    pub tex: i32,
    pub albedo: [c_double; 3],
    pub param: c_double,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct rt_hittable_desc {
    pub kind: i32,
    pub mat: i32,
    pub child: i32,
    pub count: i32,
    pub flags: u32,
    pub _pad: i32,
    pub v0: [c_double; 3],
    pub v1: [c_double; 3],
    pub v2: [c_double; 3],
    pub v3: [c_double; 3],
    pub n: [c_double; 3],
    pub s0: c_double,
    pub s1: c_double,
    pub bbox: [c_double; 6],
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct rt_bvh_node_desc {
    pub bbox: [c_double; 6],
    pub left: i32,
    pub right: i32,
    pub object: i32,
    pub axis: i32,
}

#[repr(C)]
pub struct rt_perlin_desc {
    pub ranvec: [[c_double; 3]; 256],
    pub perm_x: [i32; 256],
    pub perm_y: [i32; 256],
    pub perm_z: [i32; 256],
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct rt_image_desc {
    pub width: i32,
    pub height: i32,
    pub rgb8: *const u8,
}

#[repr(C)]
pub struct rt_scene_desc {
    pub abi_version: i32,
    pub world: i32,
    pub n_textures: i32,
    pub n_materials: i32,
    pub n_hittables: i32,
    pub n_list_items: i32,
    pub n_bvh_nodes: i32,
    pub n_perlins: i32,
    pub n_images: i32,
    pub _pad: i32,
    pub textures: *const rt_texture_desc,
    pub materials: *const rt_material_desc,
    pub hittables: *const rt_hittable_desc,
    pub list_items: *const i32,
    pub bvh_nodes: *const rt_bvh_node_desc,
    pub perlins: *const rt_perlin_desc,
    pub images: *const rt_image_desc,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct rt_camera_desc {
    pub image_width: i64,
    pub image_height: i64,
    pub samples_per_pixel: i32,
    pub max_depth: i32,
    pub background: [c_double; 3],
    pub center: [c_double; 3],
    pub pixel00_loc: [c_double; 3],
    pub pixel_delta_u: [c_double; 3],
    pub pixel_delta_v: [c_double; 3],
    pub defocus_angle: c_double,
    pub defocus_disk_u: [c_double; 3],
    pub defocus_disk_v: [c_double; 3],
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct rt_render_stats {
    pub paths: u64,
    pub segments: u64,
    pub kernel_launches: u64,
    pub last_kernel_ms: c_float,
}

pub const RT_B200_ABI_VERSION: i32 = 2;
pub const RT_OK: c_int = 0;

// ------------------------------------------------------------------------------------------------ C functions
#[link(name = "rt_b200")]
extern "C" {
    pub fn rt_last_error() -> *const c_char;
    pub fn rt_abi_version() -> c_int;
    pub fn rt_builder_create(bvh_seed: u64, out: *mut *mut rt_builder) -> c_int;
    pub fn rt_builder_destroy(b: *mut rt_builder);
    pub fn rt_tex_solid(b: *mut rt_builder, r: c_double, g: c_double, bl: c_double) -> c_int;
    pub fn rt_tex_checker_inv(b: *mut rt_builder, inv_scale: c_double, even_tex: c_int, odd_tex: c_int) -> c_int;
    pub fn rt_tex_image(b: *mut rt_builder, width: c_int, height: c_int, rgb8: *const u8) -> c_int;
    pub fn rt_tex_noise_tables(b: *mut rt_builder, scale: c_double, ranvec: *const c_double, perm_x: *const i32,
                               perm_y: *const i32, perm_z: *const i32) -> c_int;
    pub fn rt_mat_lambertian(b: *mut rt_builder, albedo_tex: c_int) -> c_int;
    pub fn rt_mat_metal(b: *mut rt_builder, albedo: *const c_double, fuzz: c_double) -> c_int;
    pub fn rt_mat_dielectric(b: *mut rt_builder, ir: c_double) -> c_int;
    pub fn rt_mat_diffuse_light(b: *mut rt_builder, emit_tex: c_int) -> c_int;
    pub fn rt_mat_isotropic(b: *mut rt_builder, albedo_tex: c_int) -> c_int;
    pub fn rt_hit_sphere(b: *mut rt_builder, center: *const c_double, radius: c_double, mat: c_int) -> c_int;
    pub fn rt_hit_moving_sphere(b: *mut rt_builder, center: *const c_double, target: *const c_double, radius: c_double,
                                mat: c_int) -> c_int;
    pub fn rt_hit_quad(b: *mut rt_builder, q: *const c_double, u: *const c_double, v: *const c_double, mat: c_int) -> c_int;
    pub fn rt_hit_list(b: *mut rt_builder, ids: *const c_int, n: c_int) -> c_int;
    pub fn rt_hit_translate(b: *mut rt_builder, object: c_int, offset: *const c_double) -> c_int;
    pub fn rt_hit_rotate_y_sincos(b: *mut rt_builder, object: c_int, sin_theta: c_double, cos_theta: c_double) -> c_int;
    pub fn rt_hit_constant_medium_nid(b: *mut rt_builder, boundary: c_int, neg_inv_density: c_double, albedo_tex: c_int) -> c_int;
    pub fn rt_hit_bvh_nodes(b: *mut rt_builder, nodes: *const rt_bvh_node_desc, n: c_int) -> c_int;
    pub fn rt_hit_bvh_device(b: *mut rt_builder, ctx: *mut rt_context, ids: *const c_int, n: c_int) -> c_int;
    pub fn rt_jpeg_decode(ctx: *mut rt_context, jpeg: *const u8, n_bytes: usize, width: *mut c_int, height: *mut c_int,
                          host_rgb8: *mut u8, capacity: usize) -> c_int;
    pub fn rt_builder_finish(b: *mut rt_builder, world: c_int, out: *mut rt_scene_desc) -> c_int;
    pub fn rt_context_create(device_id: c_int, out: *mut *mut rt_context) -> c_int;
    pub fn rt_context_destroy(ctx: *mut rt_context);
    pub fn rt_scene_upload(ctx: *mut rt_context, desc: *const rt_scene_desc, out: *mut *mut rt_scene) -> c_int;
    pub fn rt_scene_destroy(scene: *mut rt_scene);
    pub fn rt_render(ctx: *mut rt_context, scene: *const rt_scene, cam: *const rt_camera_desc, sample_begin: i64,
                     sample_count: i64, seed: u64, host_sum_rgba: *mut c_float) -> c_int;
    pub fn rt_render_multi(ctxs: *const *mut rt_context, scenes: *const *const rt_scene, n_devices: c_int,
                           cam: *const rt_camera_desc, sample_begin: i64, sample_count: i64, seed: u64,
                           weights: *const c_double, host_sum_rgba: *mut c_float, shares_out: *mut i64) -> c_int;
    pub fn rt_render_accumulate(ctx: *mut rt_context, scene: *const rt_scene, cam: *const rt_camera_desc, sample_begin: i64,
                                sample_count: i64, seed: u64, d_sum_rgba: *mut c_void, stream: *mut c_void) -> c_int;
    pub fn rt_render_rgb8(ctx: *mut rt_context, scene: *const rt_scene, cam: *const rt_camera_desc, sample_begin: i64,
                          sample_count: i64, seed: u64, host_rgb8: *mut u8) -> c_int;
    pub fn rt_finalize_rgb8(ctx: *mut rt_context, d_sum_rgba: *const c_void, n_pixels: i64, spp: c_double,
                            host_rgb8: *mut u8, stream: *mut c_void) -> c_int;
    pub fn rt_render_get_stats(ctx: *mut rt_context, out: *mut rt_render_stats) -> c_int;
}

pub fn check(rc: c_int) -> c_int {
    if rc < 0 {
        let msg = unsafe { std::ffi::CStr::from_ptr(rt_last_error()) }.to_string_lossy().into_owned();
        panic!("rt_b200 error {}: {}", rc, msg); // the crate panics on every failure it meets (texture.rs:78, renderer.rs:59,72)
    }
    rc
}

fn v3(v: &Vec3) -> [c_double; 3] { [v.x, v.y, v.z] }

// ------------------------------------------------------------------------------------------------ the builder
pub struct SceneBuilder {
    pub raw: *mut rt_builder,
    seen: HashMap<usize, c_int>, // Arc address -> id already handed out
}

impl SceneBuilder {
    pub fn new() -> Self {
        let mut raw = std::ptr::null_mut();
        check(unsafe { rt_builder_create(0, &mut raw) }); // the seed only feeds rt_hit_bvh / rt_tex_noise, which this file does not call
        Self { raw, seen: HashMap::new() }
    }
    fn key<T: ?Sized>(a: &Arc<T>) -> usize { Arc::as_ptr(a) as *const () as usize }
    pub fn texture(&mut self, t: &Arc<dyn Texture>) -> c_int {
        let k = Self::key(t);
        if let Some(id) = self.seen.get(&k) { return *id; }
        let id = t.flatten(self);
        self.seen.insert(k, id);
        id
    }
    pub fn material(&mut self, m: &Arc<dyn Material>) -> c_int {
        let k = Self::key(m);
        if let Some(id) = self.seen.get(&k) { return *id; }
        let id = m.flatten(self);
        self.seen.insert(k, id);
        id
    }
    pub fn hittable(&mut self, h: &Arc<dyn Hittable>) -> c_int {
        let k = Self::key(h);
        if let Some(id) = self.seen.get(&k) { return *id; }
        let id = h.flatten(self);
        self.seen.insert(k, id);
        id
    }
}
impl Drop for SceneBuilder {
    fn drop(&mut self) { unsafe { rt_builder_destroy(self.raw) } }
}

// ------------------------------------------------------------------------------------------------ trait additions
// texture.rs:12-14     pub trait Texture:  Sync + Send { fn value(..) -> Color;  fn flatten(&self, b: &mut SceneBuilder) -> c_int; }
// material.rs:11-16    pub trait Material: Sync + Send { fn scatter(..); fn emitted(..); fn flatten(&self, b: &mut SceneBuilder) -> c_int; }
// hittable.rs:45-48    pub trait Hittable: Sync + Send { fn hit(..); fn bounding_box(&self) -> AABB; fn flatten(&self, b: &mut SceneBuilder) -> c_int; }
// The bodies below are those methods; each goes into the `impl Trait for Type` block the crate already has (private fields
// are visible there).

// ---- texture.rs
impl SolidColor {                                        // texture.rs:17-19
    fn flatten(&self, b: &mut SceneBuilder) -> c_int {
        check(unsafe { rt_tex_solid(b.raw, self.color.x, self.color.y, self.color.z) })
    }
}
impl CheckerTexture {                                    // texture.rs:38-42
    fn flatten(&self, b: &mut SceneBuilder) -> c_int {
        let (even, odd) = (b.texture(&self.even), b.texture(&self.odd));
        check(unsafe { rt_tex_checker_inv(b.raw, self.inv_scale, even, odd) })   // the stored 1/scale, bit for bit (texture.rs:46)
    }
}
impl ImageTexture {                                      // texture.rs:72-74: the decoded image, tightly packed RGB8, row 0 = top
    fn flatten(&self, b: &mut SceneBuilder) -> c_int {
        let rgb = self.image.to_rgb8();
        check(unsafe { rt_tex_image(b.raw, rgb.width() as c_int, rgb.height() as c_int, rgb.as_raw().as_ptr()) }) // copied by the library
    }
}
impl NoiseTexture {                                      // texture.rs:95-98 + perlin.rs:8-13: the tables THIS run drew
    fn flatten(&self, b: &mut SceneBuilder) -> c_int {
        let p = &self.noise;
        let ranvec: Vec<c_double> = p.ranvec.iter().flat_map(|v| [v.x, v.y, v.z]).collect();
        check(unsafe { rt_tex_noise_tables(b.raw, self.scale, ranvec.as_ptr(), p.perm_x.as_ptr(), p.perm_y.as_ptr(), p.perm_z.as_ptr()) })
    }
}

// ---- material.rs
impl Lambertian {                                        // material.rs:18-20
    fn flatten(&self, b: &mut SceneBuilder) -> c_int {
        let t = b.texture(&self.albedo);
        check(unsafe { rt_mat_lambertian(b.raw, t) })
    }
}
impl Metal {                                             // material.rs:44-47 (fuzz is not clamped)
    fn flatten(&self, b: &mut SceneBuilder) -> c_int {
        check(unsafe { rt_mat_metal(b.raw, v3(&self.albedo).as_ptr(), self.fuzz) })
    }
}
impl Dielectric {                                        // material.rs:66-68
    fn flatten(&self, b: &mut SceneBuilder) -> c_int { check(unsafe { rt_mat_dielectric(b.raw, self.ir) }) }
}
impl DiffuseLight {                                      // material.rs:106-108
    fn flatten(&self, b: &mut SceneBuilder) -> c_int {
        let t = b.texture(&self.emit);
        check(unsafe { rt_mat_diffuse_light(b.raw, t) })
    }
}
impl Isotropic {                                         // material.rs:124-126
    fn flatten(&self, b: &mut SceneBuilder) -> c_int {
        let t = b.texture(&self.albedo);
        check(unsafe { rt_mat_isotropic(b.raw, t) })
    }
}

// ---- sphere.rs / quad.rs / hittable.rs / constant_medium.rs / bvh.rs
impl Sphere {                                            // sphere.rs:13-20
    fn flatten(&self, b: &mut SceneBuilder) -> c_int {
        let mat = b.material(&self.material);
        let c = v3(&self.center);
        if self.is_moving {
            let target = v3(&(self.center + self.center_vec)); // Sphere::with_target stores target - center (sphere.rs:34-45)
            check(unsafe { rt_hit_moving_sphere(b.raw, c.as_ptr(), target.as_ptr(), self.radius, mat) })
        } else {
            check(unsafe { rt_hit_sphere(b.raw, c.as_ptr(), self.radius, mat) })
        }
    }
}
impl Quad {                                              // quad.rs:11-20: q, u, v; the library re-derives w, d, normal, bbox (quad.rs:23-43)
    fn flatten(&self, b: &mut SceneBuilder) -> c_int {
        let mat = b.material(&self.mat);
        check(unsafe { rt_hit_quad(b.raw, v3(&self.q).as_ptr(), v3(&self.u).as_ptr(), v3(&self.v).as_ptr(), mat) })
    }
}
impl HittableList {                                      // hittable.rs:51-54; a list made by Quad::cube is recognised by the library
    fn flatten(&self, b: &mut SceneBuilder) -> c_int {
        let ids: Vec<c_int> = self.objects.iter().map(|o| b.hittable(o)).collect();
        check(unsafe { rt_hit_list(b.raw, ids.as_ptr(), ids.len() as c_int) })
    }
}
impl Translate {                                         // hittable.rs:81-85
    fn flatten(&self, b: &mut SceneBuilder) -> c_int {
        let child = b.hittable(&self.object);
        check(unsafe { rt_hit_translate(b.raw, child, v3(&self.offset).as_ptr()) })
    }
}
impl RotateY {                                           // hittable.rs:113-118: the angle is gone, sin / cos are kept - and handed over as they are
    fn flatten(&self, b: &mut SceneBuilder) -> c_int {
        let child = b.hittable(&self.object);
        check(unsafe { rt_hit_rotate_y_sincos(b.raw, child, self.sin_theta, self.cos_theta) })
    }
}
impl ConstantMedium {                                    // constant_medium.rs:14-18: the phase function is always Isotropic(albedo)
    fn flatten(&self, b: &mut SceneBuilder) -> c_int {
        let boundary = b.hittable(&self.boundary);
        let albedo = self.phase_function.albedo_texture(b); // one-line accessor on Isotropic: b.texture(&self.albedo)
        check(unsafe { rt_hit_constant_medium_nid(b.raw, boundary, self.neg_inv_density, albedo) })
    }
}
impl BVHNode {                                           // bvh.rs:12-19: the tree this run built, node for node, in pre-order
    fn flatten(&self, b: &mut SceneBuilder) -> c_int {
        fn walk(n: &(Node, crate::aabb::AABB), b: &mut SceneBuilder, out: &mut Vec<rt_bvh_node_desc>) -> i32 {
            let me = out.len();
            let bb = &n.1;
            out.push(rt_bvh_node_desc { bbox: [bb.x.min, bb.x.max, bb.y.min, bb.y.max, bb.z.min, bb.z.max],
                                        left: -1, right: -1, object: -1, axis: -1 });
            match &n.0 {
                Node::Leaf(obj) => { out[me].object = b.hittable(obj); }
                Node::Branch(l, r) => {
                    let li = walk(l, b, out);
                    let ri = walk(r, b, out);
                    out[me].left = li;
                    out[me].right = ri;
                }
            }
            me as i32
        }
        let mut nodes = Vec::new();
        walk(&self.root, b, &mut nodes);
        check(unsafe { rt_hit_bvh_nodes(b.raw, nodes.as_ptr(), nodes.len() as c_int) })
    }
}

// ---- camera.rs
impl Camera {                                            // camera.rs:38-51, field for field
    pub fn to_desc(&self) -> rt_camera_desc {
        rt_camera_desc {
            image_width: self.image_width as i64,
            image_height: self.image_height as i64,
            samples_per_pixel: self.samples_per_pixel,
            max_depth: self.max_depth,
            background: v3(&self.background),
            center: v3(&self.center),
            pixel00_loc: v3(&self.pixel00_loc),
            pixel_delta_u: v3(&self.pixel_delta_u),
            pixel_delta_v: v3(&self.pixel_delta_v),
            defocus_angle: self.defocus_angle,
            defocus_disk_u: v3(&self.defocus_disk_u),
            defocus_disk_v: v3(&self.defocus_disk_v),
        }
    }
}

// ------------------------------------------------------------------------------------------------ renderer.rs
/// Replaces the rayon loop of renderer.rs:26-49: per-pixel SUM over samples_per_pixel, row-major, W*H entries.
/// `devices`: GPU ids to use; more than one shards the sample range and reduces over NVLink (rt_render_multi).
pub fn render_sums(camera: &Camera, world: &Arc<dyn Hittable>, devices: &[c_int]) -> Vec<Color> {
    let mut b = SceneBuilder::new();
    let world_id = b.hittable(world);
    let mut desc: rt_scene_desc = unsafe { std::mem::zeroed() };
    check(unsafe { rt_builder_finish(b.raw, world_id, &mut desc) });
    assert_eq!(desc.abi_version, RT_B200_ABI_VERSION);
    let cam = camera.to_desc();
    let (w, h) = (camera.image_width, camera.image_height);
    let mut sums = vec![0f32; w * h * 4]; // x, y, z = SUM over spp (renderer.rs:39), w = sample count
    let mut ctxs: Vec<*mut rt_context> = Vec::new();
    let mut scenes: Vec<*const rt_scene> = Vec::new();
    for d in devices {
        let (mut c, mut s) = (std::ptr::null_mut(), std::ptr::null_mut());
        check(unsafe { rt_context_create(*d, &mut c) });
        check(unsafe { rt_scene_upload(c, &desc, &mut s) });
        ctxs.push(c);
        scenes.push(s as *const rt_scene);
    }
    check(unsafe {
        rt_render_multi(ctxs.as_ptr(), scenes.as_ptr(), ctxs.len() as c_int, &cam, 0, camera.samples_per_pixel as i64, 0,
                        std::ptr::null(), sums.as_mut_ptr(), std::ptr::null_mut())
    });
    for (c, s) in ctxs.iter().zip(scenes.iter()) {
        unsafe { rt_scene_destroy(*s as *mut rt_scene); rt_context_destroy(*c); }
    }
    sums.chunks_exact(4).map(|p| Color::new(p[0] as FP, p[1] as FP, p[2] as FP)).collect()
}

/* renderer.rs, inside render(): the one edit to the crate's control flow.

    let start = Instant::now();
    let raw_pixels: Vec<Color> = crate::b200::render_sums(&camera, &world, &[0]);   // was: (0..width*height).into_par_iter() ... .collect()
    println!("Render time: {:.2?}", start.elapsed());
    // unchanged from here: c / samples_per_pixel, color_to_rgb, PngEncoder (renderer.rs:53-74)
*/
