// bindings/rust/b200.rs — Rust side of the C ABI in include/rt_b200.h, as a maintainer of rust-tracing would add it.
//
// NOT COMPILED OR TESTED HERE: the build image has no cargo/rustc. It is shipped as source (SURVEY.md 8(f) rank 3);
// the same entry points are exercised through the ctypes binding (rust-tracing_b200/_abi.py) by every GPU test.
// See INTEGRATION.md for where each piece goes in the crate.

// src/b200.rs — new file
use std::os::raw::{c_char, c_double, c_int, c_void};

#[repr(C)] pub struct RtBuilder { _p: [u8; 0] }
#[repr(C)] pub struct RtContext { _p: [u8; 0] }
#[repr(C)] pub struct RtScene   { _p: [u8; 0] }
#[repr(C)] #[derive(Default)] pub struct RtSceneDesc { /* field for field: rt_scene_desc, include/rt_b200.h */ }
#[repr(C)] pub struct RtCameraDesc {                  // rt_camera_desc == Camera (camera.rs:38-51)
    pub image_width: i64, pub image_height: i64, pub samples_per_pixel: i32, pub max_depth: i32,
    pub background: [c_double; 3], pub center: [c_double; 3], pub pixel00_loc: [c_double; 3],
    pub pixel_delta_u: [c_double; 3], pub pixel_delta_v: [c_double; 3], pub defocus_angle: c_double,
    pub defocus_disk_u: [c_double; 3], pub defocus_disk_v: [c_double; 3],
}

#[link(name = "rt_b200")]
extern "C" {
    pub fn rt_last_error() -> *const c_char;
    pub fn rt_builder_create(bvh_seed: u64, out: *mut *mut RtBuilder) -> c_int;
    pub fn rt_builder_destroy(b: *mut RtBuilder);
    pub fn rt_tex_solid(b: *mut RtBuilder, r: c_double, g: c_double, bl: c_double) -> c_int;          // texture.rs:21
    pub fn rt_tex_checker(b: *mut RtBuilder, scale: c_double, even: c_int, odd: c_int) -> c_int;      // texture.rs:44
    pub fn rt_tex_image(b: *mut RtBuilder, w: c_int, h: c_int, rgb8: *const u8) -> c_int;             // texture.rs:76
    pub fn rt_tex_noise(b: *mut RtBuilder, scale: c_double, perlin_seed: u64) -> c_int;               // texture.rs:100
    pub fn rt_mat_lambertian(b: *mut RtBuilder, tex: c_int) -> c_int;                                 // material.rs:22
    pub fn rt_mat_metal(b: *mut RtBuilder, albedo: *const c_double, fuzz: c_double) -> c_int;         // material.rs:49
    pub fn rt_mat_dielectric(b: *mut RtBuilder, ir: c_double) -> c_int;                               // material.rs:70
    pub fn rt_mat_diffuse_light(b: *mut RtBuilder, tex: c_int) -> c_int;                              // material.rs:110
    pub fn rt_mat_isotropic(b: *mut RtBuilder, tex: c_int) -> c_int;                                  // material.rs:128
    pub fn rt_hit_sphere(b: *mut RtBuilder, c: *const c_double, r: c_double, mat: c_int) -> c_int;    // sphere.rs:23
    pub fn rt_hit_moving_sphere(b: *mut RtBuilder, c: *const c_double, target: *const c_double, r: c_double, mat: c_int) -> c_int; // sphere.rs:34
    pub fn rt_hit_quad(b: *mut RtBuilder, q: *const c_double, u: *const c_double, v: *const c_double, mat: c_int) -> c_int;        // quad.rs:23
    pub fn rt_hit_cube(b: *mut RtBuilder, a: *const c_double, bb: *const c_double, mat: c_int) -> c_int;                            // quad.rs:45
    pub fn rt_hit_list(b: *mut RtBuilder, ids: *const c_int, n: c_int) -> c_int;                      // hittable.rs:56
    pub fn rt_hit_translate(b: *mut RtBuilder, obj: c_int, offset: *const c_double) -> c_int;         // hittable.rs:87
    pub fn rt_hit_rotate_y(b: *mut RtBuilder, obj: c_int, angle_deg: c_double) -> c_int;              // hittable.rs:120
    pub fn rt_hit_constant_medium(b: *mut RtBuilder, boundary: c_int, density: c_double, tex: c_int) -> c_int; // constant_medium.rs:21
    pub fn rt_hit_bvh(b: *mut RtBuilder, ids: *const c_int, n: c_int) -> c_int;                       // bvh.rs:25
    pub fn rt_builder_finish(b: *mut RtBuilder, world: c_int, out: *mut RtSceneDesc) -> c_int;
    pub fn rt_context_create(device: c_int, out: *mut *mut RtContext) -> c_int;
    pub fn rt_context_destroy(c: *mut RtContext);
    pub fn rt_scene_upload(c: *mut RtContext, d: *const RtSceneDesc, out: *mut *mut RtScene) -> c_int;
    pub fn rt_scene_destroy(s: *mut RtScene);
    pub fn rt_render(c: *mut RtContext, s: *const RtScene, cam: *const RtCameraDesc, sample_begin: i64,
                     sample_count: i64, seed: u64, host_sum_rgba: *mut f32) -> c_int;                 // renderer.rs:26-49
}

pub struct SceneBuilder { pub raw: *mut RtBuilder }

// one extra method per trait (hittable.rs:45-48, material.rs:11-16, texture.rs:12-14)
pub trait FlattenTexture  { fn flatten(&self, b: &mut SceneBuilder) -> c_int; }
pub trait FlattenMaterial { fn flatten(&self, b: &mut SceneBuilder) -> c_int; }
pub trait FlattenHittable { fn flatten(&self, b: &mut SceneBuilder) -> c_int; }

// examples — the other impls follow the same pattern
impl FlattenHittable for Sphere {            // sphere.rs:12-19
    fn flatten(&self, b: &mut SceneBuilder) -> c_int {
        let mat = self.material.flatten(b);
        let c = [self.center.x, self.center.y, self.center.z];
        if self.is_moving {
            let t = self.center + self.center_vec;
            unsafe { rt_hit_moving_sphere(b.raw, c.as_ptr(), [t.x, t.y, t.z].as_ptr(), self.radius, mat) }
        } else {
            unsafe { rt_hit_sphere(b.raw, c.as_ptr(), self.radius, mat) }
        }
    }
}
impl FlattenHittable for BVHNode {           // bvh.rs:11-19: re-described from the leaf objects in insertion order;
    fn flatten(&self, b: &mut SceneBuilder) -> c_int {   // the library rebuilds the tree with the same split rule
        let ids: Vec<c_int> = self.leaves_in_insertion_order().iter().map(|o| o.flatten(b)).collect();
        unsafe { rt_hit_bvh(b.raw, ids.as_ptr(), ids.len() as c_int) }
    }
}

// ---- renderer.rs: replacement of the rayon loop (renderer.rs:26-49) --------------------------------------------
/*
// renderer.rs, inside render(): replaces the (0..width*height).into_par_iter() ... .collect()
let raw_pixels: Vec<Color> = {
    let mut b = SceneBuilder::new(/*bvh_seed*/ 2);
    let world_id = world.flatten(&mut b);
    let mut desc = RtSceneDesc::default();
    check(unsafe { rt_builder_finish(b.raw, world_id, &mut desc) });
    let (mut ctx, mut scene) = (std::ptr::null_mut(), std::ptr::null_mut());
    check(unsafe { rt_context_create(0, &mut ctx) });
    check(unsafe { rt_scene_upload(ctx, &desc, &mut scene) });
    let cam = camera.to_desc();                          // copies the 12 fields of camera.rs:38-51
    let mut sums = vec![0f32; width * height * 4];       // x,y,z = SUM over spp (renderer.rs:39), w = sample count
    check(unsafe { rt_render(ctx, scene, &cam, 0, spp as i64, /*seed*/ 0, sums.as_mut_ptr()) });
    unsafe { rt_scene_destroy(scene); rt_context_destroy(ctx); }
    sums.chunks_exact(4).map(|p| Color::new(p[0] as FP, p[1] as FP, p[2] as FP)).collect()
};
// unchanged from here: println!("Render time ..."), c / spp, color_to_rgb, PngEncoder (renderer.rs:51-74)
*/
