/*
 * rt_b200.h — C-ABI boundary of the B200-native path-tracing hot path.
 *
 * The reference (Husenap/rust-tracing) has no FFI; its seam is
 *     pub fn render(camera: Arc<Camera>, world: Arc<dyn Hittable>, output_file_name: String)
 * (renderer.rs:12) and, inside it, the parallel sample loop renderer.rs:26-49.
 * Everything in this header replaces that loop (and the host-side flattening
 * that a device needs before it):
 *
 *   rt_builder_*      the crate's constructors, one for one, producing ids into
 *                     a flat description instead of Arc<dyn Trait> objects
 *                     (sphere.rs:23,34  quad.rs:23,45  hittable.rs:56,87,120
 *                      constant_medium.rs:21,28  bvh.rs:22,25  material.rs:22,49,70,110,128
 *                      texture.rs:21,27,44,51,76,100)
 *   rt_camera_new     Camera::new (camera.rs:54-110)
 *   rt_scene_builtin  the nine CLI scenes (main.rs:56-639, indices main.rs:645-656)
 *   rt_scene_upload   flatten -> device layout (no reference equivalent: the
 *                     reference shares the object graph through Arc)
 *   rt_render*        the parallel loop renderer.rs:26-49 + ray_color :139-155;
 *                     output is the per-pixel SUM over the sample range, as in
 *                     renderer.rs:39,46 (the caller divides by spp, :57)
 *   rt_render_multi   the same loop sharded over the GPUs of one process, one ncclReduce to the first
 *   rt_finalize_rgb8  color_to_rgb (color.rs:12-19) applied to sum/spp (renderer.rs:55-58)
 *   rt_hit_batch      Hittable::hit (hittable.rs:45-48) on a ray batch — parity entry point
 *   rt_texture_batch  Texture::value (texture.rs:12-14) on a batch — parity entry point
 *   rt_get_ray_batch  Camera::get_ray (camera.rs:112-126) — parity entry point
 *   rt_scatter_batch  Material::emitted / scatter (material.rs:26-138) on a batch — parity entry point
 *
 * Conventions: extern "C", POD only, no exceptions cross the boundary. Every
 * function returns 0 on success or a negative rt_status; rt_last_error() gives
 * the message of the calling thread's last failure. Builder functions that
 * create something return its id (>= 0) or a negative rt_status.
 * All reals in descriptions are f64 because the reference computes in f64
 * (common.rs:1, FP = f64); the device library narrows to f32 at upload.
 * Host memory passed in is owned by the caller and copied; device memory is
 * owned by the handle that allocated it. One context per GPU; a context is
 * used by one host thread at a time. A scene keeps its context alive: destroying
 * the context first is allowed, its memory goes when the last scene does.
 * The library reads no environment variables.
 */
#ifndef RT_B200_H
#define RT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RT_B200_ABI_VERSION 2

typedef enum rt_status {
    RT_OK = 0,
    RT_ERR_INVALID_ARGUMENT = -1,
    RT_ERR_OUT_OF_RANGE = -2,
    RT_ERR_UNSUPPORTED = -3,   /* scene nesting the device layout cannot express */
    RT_ERR_CUDA = -4,
    RT_ERR_NO_DEVICE = -5,
    RT_ERR_OUT_OF_MEMORY = -6,
    RT_ERR_INTERNAL = -7
} rt_status;

/* ---- closed sets (impl blocks of Texture / Material / Hittable) ---- */

typedef enum rt_texture_kind {
    RT_TEX_SOLID = 0,    /* texture.rs:32-36  */
    RT_TEX_CHECKER = 1,  /* texture.rs:59-70  */
    RT_TEX_IMAGE = 2,    /* texture.rs:82-93  */
    RT_TEX_NOISE = 3     /* texture.rs:107-111 */
} rt_texture_kind;

typedef enum rt_material_kind {
    RT_MAT_LAMBERTIAN = 0,     /* material.rs:26-42  */
    RT_MAT_METAL = 1,          /* material.rs:53-64  */
    RT_MAT_DIELECTRIC = 2,     /* material.rs:80-104 */
    RT_MAT_DIFFUSE_LIGHT = 3,  /* material.rs:114-122 */
    RT_MAT_ISOTROPIC = 4       /* material.rs:132-138 */
} rt_material_kind;

typedef enum rt_hittable_kind {
    RT_HIT_SPHERE = 0,           /* sphere.rs:58   */
    RT_HIT_QUAD = 1,             /* quad.rs:96     */
    RT_HIT_LIST = 2,             /* hittable.rs:61 */
    RT_HIT_TRANSLATE = 3,        /* hittable.rs:96 */
    RT_HIT_ROTATE_Y = 4,         /* hittable.rs:159 */
    RT_HIT_CONSTANT_MEDIUM = 5,  /* constant_medium.rs:33 */
    RT_HIT_BVH = 6               /* bvh.rs:115     */
} rt_hittable_kind;

/* rt_hittable_desc.flags */
#define RT_FLAG_MOVING 1u     /* sphere: is_moving (sphere.rs:17) */
#define RT_FLAG_CUBE_LIST 2u  /* list: made by Quad::cube (quad.rs:45-93); v0 = min corner, v1 = max corner */

typedef struct rt_texture_desc {
    int32_t kind;      /* rt_texture_kind */
    int32_t a;         /* checker: even texture id | image: image id | noise: perlin id */
    int32_t b;         /* checker: odd texture id */
    int32_t _pad;
    double color[3];   /* solid: colour */
    double scale;      /* checker: inv_scale (= 1/scale, texture.rs:46) | noise: scale */
} rt_texture_desc;

typedef struct rt_material_desc {
    int32_t kind;      /* rt_material_kind */
    int32_t tex;       /* lambertian albedo | diffuse_light emit | isotropic albedo; -1 otherwise */
    double albedo[3];  /* metal */
    double param;      /* metal: fuzz | dielectric: ir */
} rt_material_desc;

typedef struct rt_hittable_desc {
    int32_t kind;      /* rt_hittable_kind */
    int32_t mat;       /* sphere, quad: material id | constant_medium: phase material id; else -1 */
    int32_t child;     /* translate, rotate_y, constant_medium: child hittable id
                          list: first index into list_items | bvh: root index into bvh_nodes */
    int32_t count;     /* list: number of items | bvh: number of nodes (2*leaves-1) */
    uint32_t flags;
    int32_t _pad;
    double v0[3];      /* sphere: center | quad: q | translate: offset | cube list: min */
    double v1[3];      /* sphere: center_vec | quad: u | cube list: max */
    double v2[3];      /* quad: v */
    double v3[3];      /* quad: w = n/|n|^2 (quad.rs:27) */
    double n[3];       /* quad: unit normal (quad.rs:25) */
    double s0;         /* sphere: radius | quad: d (quad.rs:26) | rotate_y: sin_theta | medium: neg_inv_density */
    double s1;         /* rotate_y: cos_theta */
    double bbox[6];    /* x.min x.max y.min y.max z.min z.max — Hittable::bounding_box() */
} rt_hittable_desc;

/* One node of a median-split BVH (bvh.rs:31-66), stored in pre-order. */
typedef struct rt_bvh_node_desc {
    double bbox[6];
    int32_t left;      /* node index (absolute into bvh_nodes), -1 for a leaf */
    int32_t right;
    int32_t object;    /* leaf: hittable id; branch: -1 */
    int32_t axis;      /* the axis this node drew (bvh.rs:32) — drawn for leaves too */
} rt_bvh_node_desc;

typedef struct rt_perlin_desc {   /* perlin.rs:9-14 */
    double ranvec[256][3];
    int32_t perm_x[256];
    int32_t perm_y[256];
    int32_t perm_z[256];
} rt_perlin_desc;

typedef struct rt_image_desc {    /* decoded ImageTexture (texture.rs:72-80): tightly packed RGB8, row 0 = top */
    int32_t width;
    int32_t height;
    const uint8_t* rgb8;
} rt_image_desc;

typedef struct rt_scene_desc {
    int32_t abi_version;
    int32_t world;     /* hittable id handed to render() (main.rs:659-665: always a BVH there) */
    int32_t n_textures, n_materials, n_hittables, n_list_items, n_bvh_nodes, n_perlins, n_images;
    int32_t _pad;
    const rt_texture_desc* textures;
    const rt_material_desc* materials;
    const rt_hittable_desc* hittables;
    const int32_t* list_items;
    const rt_bvh_node_desc* bvh_nodes;
    const rt_perlin_desc* perlins;
    const rt_image_desc* images;
} rt_scene_desc;

typedef struct rt_camera_settings {   /* CameraSettings, camera.rs:8-20; defaults :21-37 */
    double aspect_ratio;
    int64_t image_width;
    int32_t samples_per_pixel;
    int32_t max_depth;
    double vfov;
    double look_from[3];
    double look_at[3];
    double vup[3];
    double defocus_angle;
    double focus_dist;
    double background[3];
} rt_camera_settings;

typedef struct rt_camera_desc {       /* Camera, camera.rs:38-51 */
    int64_t image_width;
    int64_t image_height;
    int32_t samples_per_pixel;
    int32_t max_depth;                /* <= 0: every sample is black and counts, as ray_color returns (renderer.rs:140-142) */
    double background[3];
    double center[3];
    double pixel00_loc[3];
    double pixel_delta_u[3];
    double pixel_delta_v[3];
    double defocus_angle;
    double defocus_disk_u[3];
    double defocus_disk_v[3];
} rt_camera_desc;

typedef struct rt_ray_desc {          /* Ray, ray.rs:7-11 */
    double origin[3];
    double direction[3];
    double time;
} rt_ray_desc;

typedef struct rt_hit_desc {          /* HitRecord, hittable.rs:11-19 (+ ids instead of &dyn Material) */
    double t;
    double p[3];
    double normal[3];
    double u, v;
    int32_t hit;         /* 0 = None */
    int32_t front_face;
    int32_t prim_id;     /* hittable id of the Sphere / Quad / ConstantMedium that produced the record */
    int32_t mat_id;
} rt_hit_desc;

/* ------------------------------------------------------------------ errors */
const char* rt_last_error(void);
int rt_abi_version(void);

/* ------------------------------------------------------- scene construction */
typedef struct rt_builder rt_builder;

int rt_builder_create(uint64_t bvh_seed, rt_builder** out);
void rt_builder_destroy(rt_builder* b);

int rt_tex_solid(rt_builder* b, double r, double g, double bl);                           /* SolidColor::new     texture.rs:21 */
int rt_tex_checker(rt_builder* b, double scale, int even_tex, int odd_tex);               /* CheckerTexture::new texture.rs:44 */
int rt_tex_image(rt_builder* b, int width, int height, const uint8_t* rgb8);              /* ImageTexture::new   texture.rs:76 (decoded) */
int rt_tex_noise(rt_builder* b, double scale, uint64_t perlin_seed);                      /* NoiseTexture::new   texture.rs:100 */
/* ... with the tables the host's own Perlin::new drew (perlin.rs:9-25): ranvec 256 x 3 doubles, perm_* 256 x int32 */
int rt_tex_noise_tables(rt_builder* b, double scale, const double* ranvec, const int32_t* perm_x,
                        const int32_t* perm_y, const int32_t* perm_z);

int rt_mat_lambertian(rt_builder* b, int albedo_tex);                                     /* material.rs:22  */
int rt_mat_metal(rt_builder* b, const double albedo[3], double fuzz);                     /* material.rs:49  */
int rt_mat_dielectric(rt_builder* b, double ir);                                          /* material.rs:70  */
int rt_mat_diffuse_light(rt_builder* b, int emit_tex);                                    /* material.rs:110 */
int rt_mat_isotropic(rt_builder* b, int albedo_tex);                                      /* material.rs:128 */

int rt_hit_sphere(rt_builder* b, const double center[3], double radius, int mat);         /* Sphere::new        sphere.rs:23 */
int rt_hit_moving_sphere(rt_builder* b, const double center[3], const double target[3],
                         double radius, int mat);                                          /* .with_target       sphere.rs:34 */
int rt_hit_quad(rt_builder* b, const double q[3], const double u[3], const double v[3], int mat); /* Quad::new  quad.rs:23 */
int rt_hit_cube(rt_builder* b, const double a[3], const double bb[3], int mat);           /* Quad::cube         quad.rs:45 */
int rt_hit_list(rt_builder* b, const int* ids, int n);                                    /* HittableList + add hittable.rs:56
                                                                                             (six quads that are what Quad::cube
                                                                                             makes are recognised as a cube list) */
int rt_hit_translate(rt_builder* b, int object, const double offset[3]);                  /* Translate::new     hittable.rs:87 */
int rt_hit_rotate_y(rt_builder* b, int object, double angle_degrees);                     /* RotateY::new       hittable.rs:120 */
int rt_hit_constant_medium(rt_builder* b, int boundary, double density, int albedo_tex);  /* ConstantMedium::new constant_medium.rs:21 */
int rt_hit_bvh(rt_builder* b, const int* ids, int n);                                     /* BVHNode::new_from_objects bvh.rs:25 */
/* The same hittable from a tree the HOST already built (the crate's own BVHNode, bvh.rs:12-19): nodes in pre-order, root
 * first, left / right as indices into `nodes`, object = leaf hittable id or -1. The device walks exactly that tree. */
int rt_hit_bvh_nodes(rt_builder* b, const rt_bvh_node_desc* nodes, int n);

/* The same objects from their STORED state (a host that flattens objects it already built must hand over inv_scale,
 * sin / cos and neg_inv_density bit for bit, not values that round-trip through 1/x, atan2 or -1/x). */
int rt_tex_checker_inv(rt_builder* b, double inv_scale, int even_tex, int odd_tex);        /* texture.rs:38-42 */
int rt_hit_rotate_y_sincos(rt_builder* b, int object, double sin_theta, double cos_theta); /* hittable.rs:113-118 */
int rt_hit_constant_medium_nid(rt_builder* b, int boundary, double neg_inv_density, int albedo_tex); /* constant_medium.rs:14-18 */

/* Fill *out with pointers into builder-owned memory (valid until rt_builder_destroy). */
int rt_builder_finish(rt_builder* b, int world, rt_scene_desc* out);

/* Camera::new (camera.rs:54-110). */
int rt_camera_new(const rt_camera_settings* settings, rt_camera_desc* out);
/* CameraSettings::default() (camera.rs:21-37). */
void rt_camera_settings_default(rt_camera_settings* out);

/* The nine scenes of main.rs:56-639 with seeded layout RNG. image rgb8 for scenes 2 and 8
 * (earth) is passed in decoded; width/spp/depth <= 0 keep the scene's hard-coded value.
 * The world is wrapped in a BVH exactly as main.rs:659 does. */
typedef struct rt_scene_request {
    int32_t scene;            /* 0..8 (main.rs:47); anything else -> 0 (main.rs:655) */
    int32_t image_width;      /* override, <=0 = reference value */
    int32_t samples_per_pixel;
    int32_t max_depth;
    uint64_t scene_seed;
    uint64_t bvh_seed;
    uint64_t perlin_seed;
    int32_t earth_width, earth_height;
    const uint8_t* earth_rgb8;
} rt_scene_request;
int rt_scene_builtin(const rt_scene_request* req, rt_builder** builder_out,
                     rt_scene_desc* scene_out, rt_camera_settings* settings_out);

/* ------------------------------------------------------------ device side */
typedef struct rt_context rt_context;
typedef struct rt_scene rt_scene;

int rt_context_create(int device_id, rt_context** out);
void rt_context_destroy(rt_context* ctx);
int rt_device_info(rt_context* ctx, int* sm_count, int* sm_clock_khz, size_t* total_mem);

int rt_scene_upload(rt_context* ctx, const rt_scene_desc* desc, rt_scene** out);
void rt_scene_destroy(rt_scene* scene);

/* Host work moved to the GPU, results identical to the host forms:
 * rt_hit_bvh_device   rt_hit_bvh (BVHNode::node_from_list, bvh.rs:31-66) with the per-level sorting on the device - same
 *                     seeded axis stream, same node array, bit for bit;
 * rt_bvh_build_device the build itself on caller-supplied boxes (n x 6 doubles: x.min x.max y.min ...) and axis draws
 *                     (rt_bvh_axis_draws(n) of them, in the order node_from_list makes them); nodes_out receives 2n - 1
 *                     nodes in pre-order whose leaf `object` is the index into the input; returns the node count;
 * rt_jpeg_decode      ImageTexture::new's decode (texture.rs:76-80): JPEG bytes -> tightly packed RGB8 in host memory, ready
 *                     for rt_tex_image; host_rgb8 == NULL only queries the size. The entropy (Huffman) decode is one serial
 *                     bit stream and runs on the host; dequantisation, the inverse DCT, chroma upsampling and YCbCr -> RGB
 *                     run on the device with libjpeg's integer arithmetic (islow IDCT, fancy upsampling), so the bytes
 *                     equal libjpeg-turbo's / PIL's - the decode every other input of this library and of the oracle came
 *                     from. Baseline Huffman streams, 8 bit, grey or three components at 4:4:4 / 4:2:2 / 4:2:0; anything
 *                     else is RT_ERR_UNSUPPORTED;
 * rt_jpeg_entropy_decode  the host half alone (no GPU): frame geometry, quantisation tables and, when coef != NULL, the
 *                     quantised coefficients (int16, natural order, 64 per block, blocks row-major per component plane);
 * rt_jpeg_decode_nvjpeg   the same contract through NVIDIA's nvJPEG (libnvjpeg.so.12 is loaded on first use,
 *                     RT_ERR_UNSUPPORTED without it). A library decoder with its own IDCT and plain chroma replication:
 *                     NOT byte-identical to libjpeg; tests/test_gpu_jpeg.py states the measured distance. */
typedef struct rt_jpeg_info {
    int32_t width, height, components;
    int32_t h_samp[3], v_samp[3];
    int32_t blocks_w[3], blocks_h[3];     /* blocks stored per component (padded to whole MCUs) */
    int32_t adobe_rgb;                    /* Adobe APP14 transform 0: components are R, G, B */
    uint16_t quant[3][64];                /* per component, natural order */
    int64_t coef_offset[3];               /* first coefficient of the component, in int16 units */
    int64_t coef_count;
} rt_jpeg_info;
int rt_hit_bvh_device(rt_builder* b, rt_context* ctx, const int* ids, int n);
int rt_bvh_axis_draws(int n);
int rt_bvh_build_device(rt_context* ctx, const double* bboxes, int n, const int32_t* axes, rt_bvh_node_desc* nodes_out,
                        int32_t* order_out);
int rt_jpeg_decode(rt_context* ctx, const uint8_t* jpeg, size_t n_bytes, int* width, int* height, uint8_t* host_rgb8,
                   size_t capacity);
int rt_jpeg_entropy_decode(const uint8_t* jpeg, size_t n_bytes, rt_jpeg_info* info, int16_t* coef, size_t capacity);
int rt_jpeg_decode_nvjpeg(rt_context* ctx, const uint8_t* jpeg, size_t n_bytes, int* width, int* height, uint8_t* host_rgb8,
                          size_t capacity);

/* Switches of the flattening, for tests and A/B runs (0 = the product's layout). Every combination renders the same
 * image: they only change how the device walks the scene. */
#define RT_LAYOUT_NO_PRUNE 1u           /* keep every cull box of the reference's BVH (no prune_stream) */
#define RT_LAYOUT_NO_BOX_PRIMITIVES 2u  /* Quad::cube lists stay six quads */
#define RT_LAYOUT_NO_HOIST 4u           /* media stay at their BVH position */
#define RT_LAYOUT_OPS_IN_GLOBAL 8u      /* the render kernel reads the op stream from global memory, not from shared memory */
#define RT_LAYOUT_GENERIC_KERNEL 16u    /* the render kernel's generic instantiation (every feature compiled in), not the one specialised on this scene */
int rt_scene_upload_ex(rt_context* ctx, const rt_scene_desc* desc, uint32_t layout_flags, rt_scene** out);

/* What rt_scene_upload would build for this description, computed on the host (no GPU needed): sizes of the device
 * layout and how many ops of each kind the flattened traversal stream holds. Fails with the same status as
 * rt_scene_upload for descriptions the device layout cannot express (RT_ERR_UNSUPPORTED) or that are malformed. */
typedef struct rt_layout_info {
    int32_t n_words;            /* float4 words of the op stream (world program) */
    int32_t n_inner, n_sphere, n_quad, n_box, n_xform, n_medium_in_stream, n_medium_hoisted;
    int32_t n_precise_spheres;  /* spheres tested in f64 (radius > 200) */
    int32_t n_bvh;              /* BVH hittables reachable from the world */
    int64_t device_bytes;       /* op stream + materials + textures + Perlin tables + image texels (float4) */
} rt_layout_info;
int rt_scene_layout(const rt_scene_desc* desc, uint32_t layout_flags, rt_layout_info* out);

/* The flattened traversal stream itself, as rt_scene_upload would place it in HBM (host dry run, no GPU): float4
 * words (4 floats each; headers and links are integers stored bit-for-bit), the world program in [0, n_world_words),
 * hoisted media bodies behind it. words may be NULL to query *n_total_words. media_ops receives the word indices of
 * the hoisted media (capacity 8), *first_link the link of op 0 (byte offset | op class << 28). Lets a host-side check
 * walk exactly what the kernels walk (tests/opstream.py). */
int rt_scene_ops_export(const rt_scene_desc* desc, uint32_t layout_flags, float* words, int64_t capacity_words, int64_t* n_total_words,
                        int32_t* n_world_words, int32_t* media_ops, int32_t* n_media, uint32_t* first_link);

/* Render samples [sample_begin, sample_begin+sample_count) of every pixel and ADD the
 * per-pixel sums into a device float4 buffer (x,y,z = radiance sum, w = sample count),
 * W*H elements row-major (pos = j*W + i, renderer.rs:32-33), 16-byte aligned. Asynchronous on
 * `stream` (a cudaStream_t, NULL = default stream). The sample range must lie in [0, 2^32) (the
 * RNG key takes the sample index as 32 bits). Launches on different streams of one context may be
 * in flight together (each takes its own work counter; up to 64 at a time); calls into one context
 * still come from one host thread at a time. */
int rt_render_accumulate(rt_context* ctx, const rt_scene* scene, const rt_camera_desc* cam,
                         int64_t sample_begin, int64_t sample_count, uint64_t seed,
                         void* d_sum_rgba, void* stream);

/* Drop-in for renderer.rs:26-49: zero a framebuffer, render the range, copy the sums
 * to host memory (W*H*4 floats), synchronous. */
int rt_render(rt_context* ctx, const rt_scene* scene, const rt_camera_desc* cam,
              int64_t sample_begin, int64_t sample_count, uint64_t seed, float* host_sum_rgba);

/* color_to_rgb(sum/spp) for every pixel (color.rs:12-19, renderer.rs:55-58): device float4
 * sums -> host RGB8, W*H*3 bytes. spp <= 0: divide by the w channel instead. Runs on `stream`
 * (pass the stream the framebuffer was rendered on: the kernel is then ordered after those
 * launches) and returns when the bytes are in host memory. */
int rt_finalize_rgb8(rt_context* ctx, const void* d_sum_rgba, int64_t n_pixels, double spp,
                     uint8_t* host_rgb8, void* stream);

/* renderer.rs:26-58 end to end on the device: render the range into the context's framebuffer, then
 * color_to_rgb(sum / sample_count) there; only the W*H*3 RGB8 bytes cross to the host (what `-o name` writes). */
int rt_render_rgb8(rt_context* ctx, const rt_scene* scene, const rt_camera_desc* cam,
                   int64_t sample_begin, int64_t sample_count, uint64_t seed, uint8_t* host_rgb8);

/* The multi-GPU form of rt_render for a host that is not Python: one context and one uploaded copy of the
 * scene per GPU of this process, all driven from the calling thread. The sample range is split across the devices
 * (weights[i] > 0: in proportion, e.g. to each GPU's measured paths/s; weights == NULL: equally), every device accumulates
 * its share into its own SUM framebuffer, the partial framebuffers are summed onto the first device with ONE ncclReduce
 * over NVLink (communicators from ncclCommInitAll, created on first use and cached; libnccl.so.2 is loaded at that
 * moment - RT_ERR_UNSUPPORTED if the host has none), and the result is copied to host memory. With the keyed RNG the
 * image is the one a single GPU renders, up to f32 summation order. shares_out (optional, n_devices entries) receives
 * the sample count each device rendered. */
int rt_render_multi(rt_context* const* ctxs, const rt_scene* const* scenes, int n_devices, const rt_camera_desc* cam,
                    int64_t sample_begin, int64_t sample_count, uint64_t seed, const double* weights,
                    float* host_sum_rgba, int64_t* shares_out);

/* Statistics of the last rt_render_accumulate on this context; waits for the device to go idle first. */
typedef struct rt_render_stats {
    uint64_t paths;
    uint64_t segments;      /* world.hit calls (one per bounce) */
    uint64_t kernel_launches;
    float last_kernel_ms;   /* 0 unless timing was enabled */
} rt_render_stats;
int rt_render_get_stats(rt_context* ctx, rt_render_stats* out);

/* Instrumented run of the same kernel: counts the ops the device traversal executes (box tests, sphere tests,
 * quad tests, shades by material, ...) over the given sample range. counters[] receives the counts (returns how
 * many), *names_csv their comma-separated names. Used to state the device's own algorithmic flops per path. */
int rt_render_count_ops(rt_context* ctx, const rt_scene* scene, const rt_camera_desc* cam, int64_t sample_begin,
                        int64_t sample_count, uint64_t seed, uint64_t* counters, int capacity, const char** names_csv);

/* Parity entry points (host in, host out, synchronous). */
int rt_hit_batch(rt_context* ctx, const rt_scene* scene, const rt_ray_desc* rays, int64_t n,
                 double t_min, double t_max, uint64_t seed, rt_hit_desc* out);
int rt_texture_batch(rt_context* ctx, const rt_scene* scene, int tex, const double* uvp /* n*5: u v px py pz */,
                     int64_t n, double* rgb_out /* n*3 */);
int rt_get_ray_batch(rt_context* ctx, const rt_camera_desc* cam, const int64_t* pixel_index,
                     const int64_t* sample_index, int64_t n, uint64_t seed, rt_ray_desc* out);

/* Material::emitted + Material::scatter (material.rs:11-16,26-138) on a batch of (incoming ray, hit record) pairs -
 * parity entry point for the tagged scatter switch. The device's keyed RNG is (seed, pixel[k], sample[k], segment); the
 * CPU oracle driven by the same key draws the same numbers. out[k].scattered = 0: absorbed (Metal below the surface,
 * DiffuseLight); ray_out / attenuation are then left zero. */
typedef struct rt_scatter_desc {
    rt_ray_desc ray_out;      /* scattered ray (origin = hit point, direction not normalised, time carried over) */
    double attenuation[3];
    double emitted[3];        /* Material::emitted(u, v, p): non-zero for DiffuseLight only */
    int32_t scattered;
    int32_t _pad;
} rt_scatter_desc;
int rt_scatter_batch(rt_context* ctx, const rt_scene* scene, const rt_ray_desc* rays_in, const rt_hit_desc* hits, int64_t n,
                     uint64_t seed, const uint32_t* pixel, const uint32_t* sample, uint32_t segment, rt_scatter_desc* out);

/* Topology of the device's flattened traversal order for one BVH hittable: for every
 * node in pre-order, its leaf object id or -1. n_out receives the count. */
int rt_bvh_export(const rt_scene* scene, int bvh_hittable, int32_t* object_of_node, int32_t capacity,
                  int32_t* n_out);

/* Sustained FP32 FMA throughput of this GPU in TFLOP/s (FFMA loop kernel, CUDA-event timed). */
int rt_measure_fp32_peak(rt_context* ctx, double* tflops_out);

#ifdef __cplusplus
}
#endif
#endif /* RT_B200_H */
