// Kernels + device half of the C ABI (include/rt_b200.h).
//
// K1 render_kernel_mk (render_mk.cuh): persistent megakernel for the parallel loop of renderer.rs:26-49, one CTA per SM.
//                    The op stream is staged in shared memory by a bulk (TMA) copy; one warp owns a pool of (8x4 pixel
//                    tile) x (sample chunk) paths; a lane whose path ends takes the next path of the pool in the same
//                    iteration (ballot/popc ranking); lanes regroup by op class through a warp vote. Radiance sums go to
//                    a float4 framebuffer, one vector reduction per path.
// K2 hit_kernel      Hittable::hit on a ray batch (parity).
// K3 finalize_kernel color_to_rgb(sum/spp) (color.rs:12-19, renderer.rs:55-58).
// K4 texture_kernel / get_ray_kernel / scatter_kernel (parity), expand_image_kernel (upload), fma_peak_kernel.
#include "rt_kernels.cuh"

#include "../host/host_common.h"

#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <map>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

using namespace rtdev;
using rt_host::fail;

namespace {

constexpr int kTileW = 8, kTileH = 4;
constexpr int kMaxPerlinShared = 4;

struct RenderParams {
    DevScene scene;
    DevCamera cam;
    uint64_t seed;
    int64_t sample_begin;
    int sample_count;
    int chunk;            // samples per pool
    int tiles_x, tiles_y;
    int n_chunks;
    float4* sum;          // W*H float4
    unsigned int* work_counter;
    unsigned long long* stats;   // [0] paths, [1] segments, [2..] op counters (counting build only)
    uint32_t first_link;         // link of op 0 (dev_scene.h)
    uint32_t ops_bytes;          // bytes of the op stream staged in shared memory (0: read from global memory)
    int shade_min;               // the shade class may win the vote once this many lanes wait for it
    int slab_fast;               // lanes in the slab class that skip the full vote
    int sphere_reps, quad_reps;  // consecutive sphere / quad ops per vote
    // render_q.cuh (paths decoupled from lanes)
    unsigned long long* path_counter;   // next global path number
    unsigned long long total_paths;     // sample_count x tiles x 32
    int xchg_min;                // finished segments in a warp that trigger an exchange
    int slab_drop;               // a box-test round ends once this many lanes have left the class
    int min_trav;                // a traversal warp with fewer slots than this looks at the SHADE queue
};

// op counters of the instrumented kernel (rt_render_count_ops): what the device traversal actually executes
enum Counter { K_PATHS = 0, K_SEGMENTS, K_SLAB, K_BOX, K_BOX_HIT, K_SPHERE, K_SPHERE_MOVING, K_SPHERE_PRECISE, K_SPHERE_HIT,
               K_QUAD, K_QUAD_HIT, K_XFORM_ENTER, K_MEDIUM, K_MEDIUM_HIT, K_LAMBERTIAN, K_METAL, K_DIELECTRIC, K_ISOTROPIC,
               K_LIGHT, K_TEX_NOISE, K_TEX_IMAGE, K_TEX_CHECKER, K_FINALIZE_XFORM, K_VOTES, K_LANE_OPS, K_NUM };
const char* const kCounterNames =
    "paths,segments,slab,box,box_hit,sphere,sphere_moving,sphere_precise,sphere_hit,quad,quad_hit,xform_enter,medium,"
    "medium_hit,lambertian,metal,dielectric,isotropic,light,tex_noise,tex_image,tex_checker,finalize_xform,votes,lane_ops";

__device__ __forceinline__ void red_add_f4(float4* addr, float x, float y, float z, float w) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}

__device__ __forceinline__ void stage_perlin(const DevScene& S, float4* sh_vec, uint8_t* sh_perm) {
    const int n = min(S.n_perlin, kMaxPerlinShared);
    for (int k = threadIdx.x; k < n * 256; k += blockDim.x) sh_vec[k] = S.perlin_vec[k];
    for (int k = threadIdx.x; k < n * 768; k += blockDim.x) sh_perm[k] = S.perlin_perm[k];
    __syncthreads();
}

__device__ __forceinline__ void set_ops_base(OpsGlobal& o, const float4* g) { o.base = g; }
__device__ __forceinline__ void set_ops_base(OpsShared&, const float4*) {}

#include "render_mk.cuh"
#ifdef RT_B200_DEV
#include "render_q.cuh"   // measured alternative (paths decoupled from lanes through shared-memory queues): A/B build only
#endif

struct DevRayIn { float ox, oy, oz, dx, dy, dz, time, pad; };
struct DevHitOut { float t, px, py, pz, nx, ny, nz, u, v; int hit, front_face, prim, mat; };

__global__ void hit_kernel(DevScene S, const DevRayIn* rays, int64_t n, float tmin, float tmax, uint64_t seed, DevHitOut* out) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const DevRayIn r = rays[k];
    Ray ray;
    ray.o = f3(r.ox, r.oy, r.oz); ray.d = f3(r.dx, r.dy, r.dz); ray.time = r.time;
    Best best;
    const uint4 key = path_key(seed, (uint32_t)k, 0u);
    const OpsGlobal ops{S.ops};
    traverse<true>(S, ops, 0, S.n_words, ray, tmin, tmax, best, -1, key, 0u);
    DevHitOut o;
    memset(&o, 0, sizeof(o));
    o.prim = -1; o.mat = -1;
    if (best.op >= 0) {
        HitRec h;
        finalize_hit(S, ops, ray, best, h);
        if (h.uv_lazy) sphere_uv(h.sn, &h.u, &h.v);
        o.hit = 1; o.t = h.t;
        o.px = h.p.x; o.py = h.p.y; o.pz = h.p.z;
        o.nx = h.normal.x; o.ny = h.normal.y; o.nz = h.normal.z;
        o.u = h.u; o.v = h.v;
        o.front_face = h.front_face ? 1 : 0;
        o.prim = h.prim; o.mat = h.mat;
    }
    out[k] = o;
}

__global__ void texture_kernel(DevScene S, int tex, const float* uvp, int64_t n, float* rgb) {
    float4* sh_vec = dyn_smem;
    uint8_t* sh_perm = reinterpret_cast<uint8_t*>(dyn_smem + kMaxPerlinShared * 256);
    stage_perlin(S, sh_vec, sh_perm);
    PerlinShared P{sh_vec, sh_perm, min(S.n_perlin, kMaxPerlinShared)};
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const float3 c = texture_value(S, P, tex, f3(uvp[k * 5 + 2], uvp[k * 5 + 3], uvp[k * 5 + 4]), uvp[k * 5], uvp[k * 5 + 1], false,
                                   f3(0.0f, 0.0f, 0.0f));
    rgb[k * 3] = c.x; rgb[k * 3 + 1] = c.y; rgb[k * 3 + 2] = c.z;
}

__global__ void get_ray_kernel(DevCamera C, const int64_t* pixel, const int64_t* sample, int64_t n, uint64_t seed, DevRayIn* out) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const uint4 key = path_key(seed, (uint32_t)pixel[k], (uint32_t)sample[k]);
    const Ray r = camera_ray(C, (int)(pixel[k] % C.width), (int)(pixel[k] / C.width), key);
    DevRayIn o;
    o.ox = r.o.x; o.oy = r.o.y; o.oz = r.o.z; o.dx = r.d.x; o.dy = r.d.y; o.dz = r.d.z; o.time = r.time; o.pad = 0.0f;
    out[k] = o;
}

struct DevScatterIn { float ox, oy, oz, dx, dy, dz, time; float px, py, pz, nx, ny, nz, t, u, v; int front_face, mat; uint32_t pixel, sample; };
struct DevScatterOut { float ox, oy, oz, dx, dy, dz, time; float ar, ag, ab, er, eg, eb; int scattered; };

// Material::emitted + Material::scatter through the very shade() the render kernel calls (parity).
__global__ void scatter_kernel(DevScene S, const DevScatterIn* in, int64_t n, uint64_t seed, uint32_t seg, DevScatterOut* out) {
    float4* sh_vec = dyn_smem;
    uint8_t* sh_perm = reinterpret_cast<uint8_t*>(dyn_smem + min(S.n_perlin, kMaxPerlinShared) * 256);
    stage_perlin(S, sh_vec, sh_perm);
    PerlinShared P{sh_vec, sh_perm, min(S.n_perlin, kMaxPerlinShared)};
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const DevScatterIn r = in[k];
    Ray ray;
    ray.o = f3(r.ox, r.oy, r.oz); ray.d = f3(r.dx, r.dy, r.dz); ray.time = r.time;
    HitRec h;
    h.p = f3(r.px, r.py, r.pz); h.normal = f3(r.nx, r.ny, r.nz);
    h.t = r.t; h.u = r.u; h.v = r.v; h.mat = r.mat; h.prim = -1; h.origin = -1;
    h.front_face = r.front_face != 0; h.uv_lazy = false; h.sn = f3(0.0f, 0.0f, 0.0f);
    float3 L = f3(0.0f, 0.0f, 0.0f), T = f3(1.0f, 1.0f, 1.0f);
    const uint4 key = path_key(seed, r.pixel, r.sample);
    const bool alive = shade(S, P, ray, h, key, seg, L, T);
    DevScatterOut o;
    memset(&o, 0, sizeof(o));
    o.scattered = alive ? 1 : 0;
    o.er = L.x; o.eg = L.y; o.eb = L.z;
    if (alive) {
        o.ox = ray.o.x; o.oy = ray.o.y; o.oz = ray.o.z; o.dx = ray.d.x; o.dy = ray.d.y; o.dz = ray.d.z; o.time = ray.time;
        o.ar = T.x; o.ag = T.y; o.ab = T.z;
    }
    out[k] = o;
}

// color_to_rgb(sum/spp): x^(1/2.2), clamp [0, 0.999], *256 -> u8; NaN -> 0 (Rust saturating cast)
__global__ void finalize_kernel(const float4* sum, int64_t n, float inv_spp, int use_w, uint8_t* rgb) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const float4 s = sum[k];
    const float sc = use_w ? 1.0f / s.w : inv_spp;
    const float c[3] = {s.x * sc, s.y * sc, s.z * sc};
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        float g = powf(c[j], 1.0f / 2.2f);
        g = fminf(fmaxf(g, 0.0f), 0.999f);   // fmaxf(NaN, 0) = 0
        rgb[k * 3 + j] = (uint8_t)(256.0f * g);
    }
}

// max_depth <= 0: ray_color returns black before it looks at the world (renderer.rs:140-142), so every sample adds
// (0, 0, 0) and counts as one
__global__ void black_samples_kernel(float4* sum, int64_t n, float count) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) sum[k].w += count;
}

__global__ void expand_image_kernel(const uint8_t* rgb8, int64_t n, const float* lut, float4* out) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    out[k] = make_float4(lut[rgb8[k * 3]], lut[rgb8[k * 3 + 1]], lut[rgb8[k * 3 + 2]], 0.0f);
}

__global__ void fma_peak_kernel(float* out, int iters) {
    float a0 = threadIdx.x * 1e-3f, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f, a4 = a0 + 4.f, a5 = a0 + 5.f, a6 = a0 + 6.f, a7 = a0 + 7.f;
    const float b = 0.999f, c = 1e-3f;
    for (int i = 0; i < iters; ++i) {
        a0 = fmaf(a0, b, c); a1 = fmaf(a1, b, c); a2 = fmaf(a2, b, c); a3 = fmaf(a3, b, c);
        a4 = fmaf(a4, b, c); a5 = fmaf(a5, b, c); a6 = fmaf(a6, b, c); a7 = fmaf(a7, b, c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

int cuda_fail(cudaError_t e, const char* what) {
    return fail(e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver ? RT_ERR_NO_DEVICE
                : e == cudaErrorMemoryAllocation                            ? RT_ERR_OUT_OF_MEMORY
                                                                              : RT_ERR_CUDA,
                std::string(what) + ": " + cudaGetErrorString(e));
}
#define CU(call)                                            \
    do {                                                    \
        cudaError_t e__ = (call);                           \
        if (e__ != cudaSuccess) return cuda_fail(e__, #call); \
    } while (0)

size_t perlin_smem_bytes() { return (size_t)kMaxPerlinShared * (256 * sizeof(float4) + 768); }

DevCamera make_dev_camera(const rt_camera_desc& c) {
    DevCamera d;
    d.width = (int)c.image_width;
    d.height = (int)c.image_height;
    d.max_depth = c.max_depth;
    auto f = [](const double* p) { return make_float3((float)p[0], (float)p[1], (float)p[2]); };
    d.background = f(c.background);
    d.center = f(c.center);
    d.rel00 = make_float3((float)(c.pixel00_loc[0] - c.center[0]), (float)(c.pixel00_loc[1] - c.center[1]),
                          (float)(c.pixel00_loc[2] - c.center[2]));
    d.du = f(c.pixel_delta_u);
    d.dv = f(c.pixel_delta_v);
    d.disk_u = f(c.defocus_disk_u);
    d.disk_v = f(c.defocus_disk_v);
    d.defocus = !(c.defocus_angle <= 0.0) ? 1 : 0;   // camera.rs:117
    return d;
}

}  // namespace

// FEAT_* bits of a compiled scene: walks the world program op by op (payload words can alias a header only by accident
// of their bits) and the hoisted medium bodies.
static unsigned scene_features(const CompiledScene& cs) {
    unsigned feat = 0;
    auto hdr_at = [&](int i) { uint32_t h; std::memcpy(&h, &cs.ops[i].w, 4); return h; };
    auto bits = [&](float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; };
    auto medium = [&](int i) {
        const uint32_t flags = hdr_flags(hdr_at(i));
        if ((int)flags == MEDIUM_BOUNDARY_PROGRAM) feat |= FEAT_RARE;
        if ((int)flags == MEDIUM_BOUNDARY_XBOX) feat |= FEAT_XBOX;
        if ((int)flags == MEDIUM_BOUNDARY_SPHERE && ((bits(cs.ops[i + 2].w) >> 24) & FLAG_PRECISE)) feat |= FEAT_PRECISE;
    };
    for (int i = 0; i < cs.n_world_words;) {
        const uint32_t hdr = hdr_at(i), kind = hdr_kind(hdr), flags = hdr_flags(hdr);
        if (kind == OP_BOX || kind == OP_XFORM_ENTER) feat |= FEAT_FOLD;
        if (kind == OP_SPHERE && (flags & FLAG_PRECISE)) feat |= FEAT_PRECISE;
        if (kind == OP_INNER_REF) feat |= FEAT_RARE;
        if (kind == OP_MEDIUM) { feat |= FEAT_RARE; medium(i); }
        i += op_words(kind, flags);
    }
    for (int32_t m : cs.hoisted_media) medium(m);
    return feat;
}

typedef void (*render_fn)(const RenderParams);
// The render kernel is specialised on what the scene holds (FEAT_*, rt_kernels.cuh): the instantiation that renders a scene
// carries no code the scene cannot reach (the kernel is bound by instruction fetch: profiles/r2_k1_icache.md), for both
// homes of the op stream (shared memory; global memory for a stream that does not fit). FEAT_ALL is the generic form
// (RT_LAYOUT_GENERIC_KERNEL asks for it); the instrumented build exists in that form only.
template <bool OPS_SMEM, unsigned... F>
static render_fn mk_specialised(unsigned feat, std::integer_sequence<unsigned, F...>) {
    static const render_fn table[] = {render_kernel_mk<false, OPS_SMEM, F>...};
    return table[feat];
}
static render_fn mk_kernel(bool counting, bool ops_smem, unsigned feat) {
    if (counting) return ops_smem ? render_kernel_mk<true, true, FEAT_ALL> : render_kernel_mk<true, false, FEAT_ALL>;
    if (!ops_smem) return mk_specialised<false>(feat & FEAT_ALL, std::make_integer_sequence<unsigned, FEAT_ALL + 1u>{});
    return mk_specialised<true>(feat & FEAT_ALL, std::make_integer_sequence<unsigned, FEAT_ALL + 1u>{});
}

#ifdef RT_B200_DEV
static render_fn q_kernel(bool ops_smem) { return ops_smem ? render_kernel_q<true> : render_kernel_q<false>; }
#endif

// rt_render_accumulate may be called on several streams of one context: every launch takes its own work counter and
// statistics block from a ring, zeroed on the launching stream, so launches in flight never share them.
constexpr int kLaunchSlots = 64;
constexpr size_t kStagingBytes = 8u << 20;   // two pinned buffers of this size carry host -> device uploads

struct rt_context {
    int device = 0;
    int sm_count = 0;
    int clock_khz = 0;
    size_t total_mem = 0;
    size_t smem_optin = 0;       // largest dynamic shared memory a CTA may ask for
    // vote parameters of the render kernel (render_mk.cuh) and layout options; the product never reads the environment,
    // the -DRT_B200_DEV build of the library (A/B work) takes them from RT_B200_* variables
    int shade_min = 24, slab_fast = 6, sphere_reps = 2, quad_reps = 8;   // measured plateaus: profiles/r2_sweep*.log
    bool hoist_media = true, prune_boxes = true, box_primitives = true, ops_in_smem = true;
    int chunk = 8;
    bool queue_kernel = false;                 // A/B build only: render_q.cuh instead of render_mk.cuh (RT_B200_KERNEL=q)
    int xchg_min = 8, slab_drop = 8, min_trav = 24;
    unsigned long long* d_path_counters = nullptr;   // kLaunchSlots path counters (render_q.cuh)
    unsigned int* d_counters = nullptr;        // kLaunchSlots work counters
    unsigned long long* d_stats = nullptr;     // kLaunchSlots x K_NUM
    int last_slot = 0;
    float4* d_fb = nullptr;                    // rt_render's framebuffer
    size_t fb_pixels = 0;
    uint8_t* d_rgb8 = nullptr;                 // rt_finalize_rgb8's output
    size_t rgb8_bytes = 0;
    void* h_jpeg = nullptr;                    // rt_jpeg_decode: pinned coefficients and the device block, kept between calls
    size_t h_jpeg_bytes = 0;
    unsigned char* d_jpeg = nullptr;
    size_t d_jpeg_bytes = 0;
    void* h_stage[2] = {nullptr, nullptr};     // pinned staging
    cudaEvent_t stage_done[2] = {nullptr, nullptr};
    std::vector<std::pair<void*, size_t>> free_blocks;   // device blocks released by rt_scene_destroy, reused by the next upload
    uint64_t launches = 0;
    int live_scenes = 0;
    bool destroyed = false;                    // rt_context_destroy was called while scenes were alive: the last scene frees it
};

struct rt_scene {
    rt_context* ctx = nullptr;
    int device = 0;
    DevScene dev{};
    std::vector<std::pair<void*, size_t>> allocations;
    CompiledScene compiled;
    uint32_t ops_bytes = 0;
    bool ops_in_global = false;
    bool generic_kernel = false;     // RT_LAYOUT_GENERIC_KERNEL
    unsigned features = 0;       // FEAT_* the stream needs: picks the render kernel's instantiation
};

static void context_free(rt_context* c) {
    cudaSetDevice(c->device);
    cudaFree(c->d_counters);
    cudaFree(c->d_path_counters);
    cudaFree(c->d_jpeg);
    if (c->h_jpeg) cudaFreeHost(c->h_jpeg);
    cudaFree(c->d_stats);
    cudaFree(c->d_fb);
    cudaFree(c->d_rgb8);
    for (auto& b : c->free_blocks) cudaFree(b.first);
    for (void* h : c->h_stage) if (h) cudaFreeHost(h);
    for (cudaEvent_t e : c->stage_done) if (e) cudaEventDestroy(e);
    delete c;
}

// Device memory for scene data comes from the context's free list first: a host that uploads a scene per frame (or per
// bench step) gets the same blocks back instead of paying cudaMalloc / cudaFree of hundreds of megabytes every time.
static cudaError_t ctx_alloc(rt_context* c, size_t bytes, void** out) {
    const size_t want = bytes ? (bytes + 255) / 256 * 256 : 256;
    size_t best = c->free_blocks.size();
    for (size_t k = 0; k < c->free_blocks.size(); ++k)
        if (c->free_blocks[k].second >= want && c->free_blocks[k].second <= want + want / 8 &&
            (best == c->free_blocks.size() || c->free_blocks[k].second < c->free_blocks[best].second)) best = k;
    if (best < c->free_blocks.size()) {
        *out = c->free_blocks[best].first;
        c->free_blocks.erase(c->free_blocks.begin() + best);
        return cudaSuccess;
    }
    cudaError_t e = cudaMalloc(out, want);
    if (e == cudaErrorMemoryAllocation) {      // give the cache back and retry once
        cudaGetLastError();
        for (auto& b : c->free_blocks) cudaFree(b.first);
        c->free_blocks.clear();
        e = cudaMalloc(out, want);
    }
    return e;
}

// Host -> device through the two pinned staging buffers (the caller's memory is pageable as far as we know).
static cudaError_t staged_upload(rt_context* c, void* dst, const void* src, size_t bytes) {
    const char* s = static_cast<const char*>(src);
    char* d = static_cast<char*>(dst);
    int k = 0;
    for (size_t off = 0; off < bytes; off += kStagingBytes, k ^= 1) {
        const size_t n = std::min(kStagingBytes, bytes - off);
        cudaError_t e = cudaEventSynchronize(c->stage_done[k]);
        if (e != cudaSuccess) return e;
        std::memcpy(c->h_stage[k], s + off, n);
        e = cudaMemcpyAsync(d + off, c->h_stage[k], n, cudaMemcpyHostToDevice, 0);
        if (e != cudaSuccess) return e;
        e = cudaEventRecord(c->stage_done[k], 0);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

#include "bvh_build.cuh"
#include "jpeg_kernels.cuh"
#include "../host/jpeg_entropy.h"

// ---- nvJPEG through dlopen (like NCCL: a host that hands over decoded pixels needs no libnvjpeg)
#include <nvjpeg.h>
namespace {
struct NvjpegApi {
    void* handle = nullptr;
    bool tried = false;
    nvjpegStatus_t (*CreateSimple)(nvjpegHandle_t*) = nullptr;
    nvjpegStatus_t (*Destroy)(nvjpegHandle_t) = nullptr;
    nvjpegStatus_t (*JpegStateCreate)(nvjpegHandle_t, nvjpegJpegState_t*) = nullptr;
    nvjpegStatus_t (*JpegStateDestroy)(nvjpegJpegState_t) = nullptr;
    nvjpegStatus_t (*GetImageInfo)(nvjpegHandle_t, const unsigned char*, size_t, int*, nvjpegChromaSubsampling_t*, int*, int*) = nullptr;
    nvjpegStatus_t (*Decode)(nvjpegHandle_t, nvjpegJpegState_t, const unsigned char*, size_t, nvjpegOutputFormat_t, nvjpegImage_t*, cudaStream_t) = nullptr;
};
NvjpegApi g_nvjpeg;
int load_nvjpeg() {
    if (g_nvjpeg.Decode) return RT_OK;
    if (!g_nvjpeg.tried) {
        g_nvjpeg.tried = true;
        for (const char* name : {"libnvjpeg.so.12", "libnvjpeg.so"}) {
            g_nvjpeg.handle = dlopen(name, RTLD_NOW | RTLD_LOCAL);
            if (g_nvjpeg.handle) break;
        }
        if (g_nvjpeg.handle) {
            void* h = g_nvjpeg.handle;
            g_nvjpeg.CreateSimple = reinterpret_cast<decltype(g_nvjpeg.CreateSimple)>(dlsym(h, "nvjpegCreateSimple"));
            g_nvjpeg.Destroy = reinterpret_cast<decltype(g_nvjpeg.Destroy)>(dlsym(h, "nvjpegDestroy"));
            g_nvjpeg.JpegStateCreate = reinterpret_cast<decltype(g_nvjpeg.JpegStateCreate)>(dlsym(h, "nvjpegJpegStateCreate"));
            g_nvjpeg.JpegStateDestroy = reinterpret_cast<decltype(g_nvjpeg.JpegStateDestroy)>(dlsym(h, "nvjpegJpegStateDestroy"));
            g_nvjpeg.GetImageInfo = reinterpret_cast<decltype(g_nvjpeg.GetImageInfo)>(dlsym(h, "nvjpegGetImageInfo"));
            g_nvjpeg.Decode = reinterpret_cast<decltype(g_nvjpeg.Decode)>(dlsym(h, "nvjpegDecode"));
        }
    }
    if (!g_nvjpeg.CreateSimple || !g_nvjpeg.Destroy || !g_nvjpeg.JpegStateCreate || !g_nvjpeg.JpegStateDestroy || !g_nvjpeg.GetImageInfo || !g_nvjpeg.Decode)
        return fail(RT_ERR_UNSUPPORTED, "rt_jpeg_decode: libnvjpeg.so.12 could not be loaded");
    return RT_OK;
}
}  // namespace

extern "C" {

// ImageTexture::new's decode (texture.rs:76-80): JPEG bytes -> tightly packed RGB8, row 0 = top, in host memory (then
// rt_tex_image as usual). width / height are always filled in; host_rgb8 may be NULL to query them. Host: markers and the
// Huffman bit stream (jpeg_entropy.cpp); device: everything per block and per pixel (jpeg_kernels.cuh).
int rt_jpeg_decode(rt_context* c, const uint8_t* jpeg, size_t n_bytes, int* width, int* height, uint8_t* host_rgb8, size_t capacity) {
    if (!c || !jpeg || !width || !height || n_bytes == 0) return fail(RT_ERR_INVALID_ARGUMENT, "rt_jpeg_decode: null argument");
    rt_host::JpegFrame f;
    std::string err;
    int rc = rt_host::jpeg_entropy_decode(jpeg, n_bytes, &f, nullptr, 0, &err);
    if (rc < 0) return fail(rc, "rt_jpeg_decode: " + err);
    *width = f.width;
    *height = f.height;
    if (!host_rgb8) return RT_OK;
    const size_t need = (size_t)f.width * (size_t)f.height * 3;
    if (capacity < need) return fail(RT_ERR_OUT_OF_RANGE, "rt_jpeg_decode: capacity too small");
    int mode = 0;
    if (f.ncomp == 3) {
        const rt_host::JpegComponent &Y = f.comp[0], &B = f.comp[1], &R = f.comp[2];
        if (B.h != R.h || B.v != R.v || B.h != 1 || B.v != 1 || Y.h != f.hmax || Y.v != f.vmax)
            return fail(RT_ERR_UNSUPPORTED, "rt_jpeg_decode: chroma layouts other than 4:4:4, 4:2:2 and 4:2:0 are not supported");
        if (Y.h == 1 && Y.v == 1) mode = 1;
        else if (Y.h == 2 && Y.v == 1) mode = 2;
        else if (Y.h == 2 && Y.v == 2) mode = 3;
        else return fail(RT_ERR_UNSUPPORTED, "rt_jpeg_decode: chroma layouts other than 4:4:4, 4:2:2 and 4:2:0 are not supported");
    }
    CU(cudaSetDevice(c->device));
    // pinned coefficients (the Huffman decoder writes them where the DMA engine reads them), one device block for
    // coefficients + sample planes + RGB; both are kept in the context and grown on demand
    struct { int16_t* h; unsigned char* d; } g{nullptr, nullptr};
    const size_t coef_bytes = f.coef_count * sizeof(int16_t);
    if (c->h_jpeg_bytes < coef_bytes) {
        if (c->h_jpeg) cudaFreeHost(c->h_jpeg);
        c->h_jpeg = nullptr; c->h_jpeg_bytes = 0;
        CU(cudaMallocHost(&c->h_jpeg, coef_bytes));
        c->h_jpeg_bytes = coef_bytes;
    }
    g.h = static_cast<int16_t*>(c->h_jpeg);
    rc = rt_host::jpeg_entropy_decode(jpeg, n_bytes, &f, g.h, f.coef_count, &err);
    if (rc < 0) return fail(rc, "rt_jpeg_decode: " + err);
    size_t plane_off[3], off = (f.coef_count * sizeof(int16_t) + 255) & ~(size_t)255;
    for (int k = 0; k < f.ncomp; ++k) {
        plane_off[k] = off;
        off += ((size_t)f.comp[k].blocks_w * 8 * f.comp[k].blocks_h * 8 + 255) & ~(size_t)255;
    }
    const size_t rgb_off = off;
    off += need;
    if (c->d_jpeg_bytes < off) {
        cudaFree(c->d_jpeg);
        c->d_jpeg = nullptr; c->d_jpeg_bytes = 0;
        CU(cudaMalloc(&c->d_jpeg, off));
        c->d_jpeg_bytes = off;
    }
    g.d = c->d_jpeg;
    CU(cudaMemcpyAsync(g.d, g.h, coef_bytes, cudaMemcpyHostToDevice, 0));
    uint16_t quant[3][64];
    std::memset(quant, 0, sizeof(quant));
    JpegColourArgs A;
    std::memset(&A, 0, sizeof(A));
    JpegPlane* planes[3] = {&A.Y, &A.Cb, &A.Cr};
    for (int k = 0; k < f.ncomp; ++k) {
        const rt_host::JpegComponent& C = f.comp[k];
        std::memcpy(quant[k], f.quant[C.tq], sizeof(quant[k]));
        JpegPlane& P = *planes[k];
        P.coef = reinterpret_cast<const int16_t*>(g.d) + C.coef_offset;
        P.samples = g.d + plane_off[k];
        P.blocks_w = C.blocks_w; P.blocks_h = C.blocks_h; P.pitch = C.blocks_w * 8;
        P.ds_w = C.ds_w; P.ds_h = C.ds_h; P.h = C.h; P.v = C.v;
    }
    CU(cudaMemcpyToSymbolAsync(c_jpeg_quant, quant, sizeof(quant), 0, cudaMemcpyHostToDevice, 0));
    JpegIdctArgs I;
    std::memset(&I, 0, sizeof(I));
    int ctas = 0;
    for (int k = 0; k < 3; ++k) {
        I.first_cta[k] = ctas;
        if (k < f.ncomp) { I.plane[k] = *planes[k]; ctas += (planes[k]->blocks_w * planes[k]->blocks_h + 127) / 128; }
    }
    I.first_cta[3] = ctas;
    jpeg_idct_kernel<<<ctas, 128>>>(I);
    A.mode = mode; A.rgb_passthrough = f.adobe_rgb ? 1 : 0; A.width = f.width; A.height = f.height;
    // chroma_window reads whole 32-bit words of a chroma row: fine for every plane (pitch = blocks x 8), and the fast form
    // needs rows of whole 8-pixel groups and more than two chroma samples (jdsample.c: fancy upsampling only then)
    const bool fast = (f.width % 8) == 0 && (f.ncomp == 1 || A.Cb.ds_w > 2);
    if (fast) jpeg_colour8_kernel<<<dim3((unsigned)((f.width / 8 + 255) / 256), (unsigned)f.height), 256>>>(A, g.d + rgb_off);
    else jpeg_colour_kernel<<<dim3((unsigned)((f.width + 1023) / 1024), (unsigned)f.height), 256>>>(A, g.d + rgb_off);
    CU(cudaGetLastError());
    CU(cudaMemcpy(host_rgb8, g.d + rgb_off, need, cudaMemcpyDeviceToHost));
    return RT_OK;
}

// The same contract as rt_jpeg_decode through the nvJPEG library (not byte-identical to libjpeg: see rt_b200.h).
int rt_jpeg_decode_nvjpeg(rt_context* c, const uint8_t* jpeg, size_t n_bytes, int* width, int* height, uint8_t* host_rgb8, size_t capacity) {
    if (!c || !jpeg || !width || !height || n_bytes == 0) return fail(RT_ERR_INVALID_ARGUMENT, "rt_jpeg_decode_nvjpeg: null argument");
    int rc = load_nvjpeg();
    if (rc < 0) return rc;
    CU(cudaSetDevice(c->device));
    nvjpegHandle_t h = nullptr;
    nvjpegJpegState_t st = nullptr;
    if (g_nvjpeg.CreateSimple(&h) != NVJPEG_STATUS_SUCCESS) return fail(RT_ERR_CUDA, "nvjpegCreateSimple failed");
    struct Guard { nvjpegHandle_t h; nvjpegJpegState_t* st; unsigned char* d = nullptr;
                   ~Guard() { if (d) cudaFree(d); if (*st) g_nvjpeg.JpegStateDestroy(*st); g_nvjpeg.Destroy(h); } } guard{h, &st};
    int comps = 0, ws[NVJPEG_MAX_COMPONENT] = {0}, hs[NVJPEG_MAX_COMPONENT] = {0};
    nvjpegChromaSubsampling_t sub;
    if (g_nvjpeg.GetImageInfo(h, jpeg, n_bytes, &comps, &sub, ws, hs) != NVJPEG_STATUS_SUCCESS)
        return fail(RT_ERR_INVALID_ARGUMENT, "rt_jpeg_decode_nvjpeg: not a JPEG stream nvJPEG can read");
    *width = ws[0];
    *height = hs[0];
    if (!host_rgb8) return RT_OK;
    const size_t need = (size_t)ws[0] * (size_t)hs[0] * 3;
    if (capacity < need) return fail(RT_ERR_OUT_OF_RANGE, "rt_jpeg_decode_nvjpeg: capacity too small");
    if (g_nvjpeg.JpegStateCreate(h, &st) != NVJPEG_STATUS_SUCCESS) return fail(RT_ERR_CUDA, "nvjpegJpegStateCreate failed");
    CU(cudaMalloc(&guard.d, need));
    nvjpegImage_t img;
    std::memset(&img, 0, sizeof(img));
    img.channel[0] = guard.d;
    img.pitch[0] = (size_t)ws[0] * 3;
    if (g_nvjpeg.Decode(h, st, jpeg, n_bytes, NVJPEG_OUTPUT_RGBI, &img, 0) != NVJPEG_STATUS_SUCCESS)
        return fail(RT_ERR_CUDA, "nvjpegDecode failed");
    CU(cudaMemcpy(host_rgb8, guard.d, need, cudaMemcpyDeviceToHost));
    return RT_OK;
}

int rt_context_create(int device_id, rt_context** out) {
    if (!out) return fail(RT_ERR_INVALID_ARGUMENT, "rt_context_create: out is null");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) return fail(RT_ERR_NO_DEVICE, std::string("rt_context_create: no CUDA device (") + cudaGetErrorString(e) + "); there is no CPU fallback");
    if (device_id < 0 || device_id >= n) return fail(RT_ERR_OUT_OF_RANGE, "rt_context_create: device id out of range");
    CU(cudaSetDevice(device_id));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device_id));
    rt_context* c = new rt_context;
    struct Guard {                 // every early return below releases what was built so far
        rt_context* c;
        ~Guard() { if (c) context_free(c); }
    } guard{c};
    c->device = device_id;
    c->sm_count = prop.multiProcessorCount;
    c->total_mem = prop.totalGlobalMem;
    c->smem_optin = prop.sharedMemPerBlockOptin;
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, device_id);
    c->clock_khz = khz;
#ifdef RT_B200_DEV
    // development switches of the A/B build; the product library is compiled without them and reads no environment
    if (const char* v = std::getenv("RT_B200_SHADE_MIN")) c->shade_min = std::max(1, std::atoi(v));
    if (const char* v = std::getenv("RT_B200_SLAB_FAST")) c->slab_fast = std::max(1, std::atoi(v));
    if (const char* v = std::getenv("RT_B200_SPHERE_REPS")) c->sphere_reps = std::max(1, std::atoi(v));
    if (const char* v = std::getenv("RT_B200_QUAD_REPS")) c->quad_reps = std::max(1, std::atoi(v));
    if (const char* v = std::getenv("RT_B200_NO_BOX")) c->box_primitives = std::atoi(v) == 0;
    if (const char* v = std::getenv("RT_B200_NO_PRUNE")) c->prune_boxes = std::atoi(v) == 0;
    if (const char* v = std::getenv("RT_B200_NO_HOIST")) c->hoist_media = std::atoi(v) == 0;
    if (const char* v = std::getenv("RT_B200_OPS_GLOBAL")) c->ops_in_smem = std::atoi(v) == 0;
    if (const char* v = std::getenv("RT_B200_CHUNK")) c->chunk = std::max(1, std::atoi(v));
    if (const char* v = std::getenv("RT_B200_KERNEL")) c->queue_kernel = std::string(v) == "q";
    if (const char* v = std::getenv("RT_B200_XCHG_MIN")) c->xchg_min = std::max(1, std::atoi(v));
    if (const char* v = std::getenv("RT_B200_SLAB_DROP")) c->slab_drop = std::max(1, std::atoi(v));
    if (const char* v = std::getenv("RT_B200_MIN_TRAV")) c->min_trav = std::max(1, std::atoi(v));
#endif
    for (int k = 0; k < 4; ++k)
        for (unsigned feat = 0; feat <= FEAT_ALL; ++feat)
            CU(cudaFuncSetAttribute(mk_kernel((k & 1) != 0, (k & 2) != 0, feat), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->smem_optin));
#ifdef RT_B200_DEV
    for (int in_smem = 0; in_smem < 2; ++in_smem)
        CU(cudaFuncSetAttribute(q_kernel(in_smem != 0), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->smem_optin));
#endif
    CU(cudaMalloc(&c->d_counters, kLaunchSlots * sizeof(unsigned int)));
    CU(cudaMalloc(&c->d_path_counters, kLaunchSlots * sizeof(unsigned long long)));
    CU(cudaMalloc(&c->d_stats, (size_t)kLaunchSlots * K_NUM * sizeof(unsigned long long)));
    CU(cudaMemset(c->d_stats, 0, (size_t)kLaunchSlots * K_NUM * sizeof(unsigned long long)));
    for (int k = 0; k < 2; ++k) {
        CU(cudaMallocHost(&c->h_stage[k], kStagingBytes));
        CU(cudaEventCreateWithFlags(&c->stage_done[k], cudaEventDisableTiming));
    }
    guard.c = nullptr;
    *out = c;
    return RT_OK;
}

void rt_context_destroy(rt_context* c) {
    if (!c) return;
    if (c->live_scenes > 0) { c->destroyed = true; return; }   // scenes still point here: the last rt_scene_destroy frees it
    context_free(c);
}

int rt_device_info(rt_context* c, int* sm_count, int* sm_clock_khz, size_t* total_mem) {
    if (!c) return fail(RT_ERR_INVALID_ARGUMENT, "rt_device_info: context is null");
    if (sm_count) *sm_count = c->sm_count;
    if (sm_clock_khz) *sm_clock_khz = c->clock_khz;
    if (total_mem) *total_mem = c->total_mem;
    return RT_OK;
}

static int upload_vec(rt_scene* s, const void* src, size_t bytes, void** dst) {
    *dst = nullptr;
    cudaError_t e = ctx_alloc(s->ctx, bytes, dst);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc(scene data)");
    s->allocations.push_back({*dst, bytes ? (bytes + 255) / 256 * 256 : 256});
    if (bytes) {
        e = staged_upload(s->ctx, *dst, src, bytes);
        if (e != cudaSuccess) return cuda_fail(e, "scene upload");
    }
    return RT_OK;
}

static CompileOptions compile_options(const rt_context* c, uint32_t layout_flags) {
    CompileOptions o;
    if (c) { o.box_primitives = c->box_primitives; o.hoist_media = c->hoist_media; o.prune_boxes = c->prune_boxes; }
    if (layout_flags & RT_LAYOUT_NO_PRUNE) o.prune_boxes = false;
    if (layout_flags & RT_LAYOUT_NO_BOX_PRIMITIVES) o.box_primitives = false;
    if (layout_flags & RT_LAYOUT_NO_HOIST) o.hoist_media = false;
#ifdef RT_B200_DEV
    if (const char* e = std::getenv("RT_B200_COST_SPHERE")) o.cost_sphere = std::atof(e);
    if (const char* e = std::getenv("RT_B200_COST_QUAD")) o.cost_quad = std::atof(e);
    if (const char* e = std::getenv("RT_B200_COST_BOX")) o.cost_box = std::atof(e);
#endif
    return o;
}

int rt_scene_upload(rt_context* c, const rt_scene_desc* desc, rt_scene** out) { return rt_scene_upload_ex(c, desc, 0u, out); }

int rt_scene_upload_ex(rt_context* c, const rt_scene_desc* desc, uint32_t layout_flags, rt_scene** out) {
    if (!c || !desc || !out) return fail(RT_ERR_INVALID_ARGUMENT, "rt_scene_upload: null argument");
    CU(cudaSetDevice(c->device));
    rt_scene* s = new rt_scene;
    s->ctx = c;
    s->device = c->device;
    c->live_scenes++;
    const char* err = nullptr;
    s->ops_in_global = (layout_flags & RT_LAYOUT_OPS_IN_GLOBAL) != 0;
    s->generic_kernel = (layout_flags & RT_LAYOUT_GENERIC_KERNEL) != 0;
    int rc = compile_scene(desc, compile_options(c, layout_flags), &s->compiled, &err);
    if (rc < 0) { rt_scene_destroy(s); return fail(rc, err ? err : "compile_scene failed"); }
    const CompiledScene& cs = s->compiled;
    void* p = nullptr;
#define UP(vec, field, type)                                                                   \
    rc = upload_vec(s, (vec).data(), (vec).size() * sizeof((vec)[0]), &p);                     \
    if (rc < 0) { rt_scene_destroy(s); return rc; }                                            \
    s->dev.field = static_cast<type>(p);
    UP(cs.ops, ops, const float4*)
    UP(cs.materials, mats, const float4*)
    UP(cs.textures, texs, const float4*)
    UP(cs.perlin_vec, perlin_vec, const float4*)
    UP(cs.perlin_perm, perlin_perm, const uint8_t*)
    UP(cs.precise, precise, const double4*)
#undef UP
    s->ops_bytes = (uint32_t)(cs.ops.size() * sizeof(F4));
    s->features = scene_features(cs);
    s->dev.n_words = cs.n_world_words;
    s->dev.n_media = (int)cs.hoisted_media.size();
    for (int k = 0; k < s->dev.n_media; ++k) s->dev.media_op[k] = cs.hoisted_media[k];
    s->dev.n_perlin = cs.n_perlin;
    // images: upload RGB8, expand on the device to linear float4 through a 256-entry LUT computed in f64
    std::vector<DevImage> imgs((size_t)desc->n_images);
    if (desc->n_images > 0) {
        float lut[256];
        for (int k = 0; k < 256; ++k) lut[k] = (float)std::pow((double)k / 255.0, 2.2);   // color.rs:8-10,21-27
        float* d_lut = nullptr;
        rc = upload_vec(s, lut, sizeof(lut), (void**)&d_lut);
        if (rc < 0) { rt_scene_destroy(s); return rc; }
        for (int k = 0; k < desc->n_images; ++k) {
            const rt_image_desc& im = desc->images[k];
            if (im.width <= 0 || im.height <= 0 || !im.rgb8) { rt_scene_destroy(s); return fail(RT_ERR_INVALID_ARGUMENT, "rt_scene_upload: empty image"); }
            const int64_t n = (int64_t)im.width * im.height;
            uint8_t* d_rgb = nullptr;
            float4* d_tex = nullptr;
            rc = upload_vec(s, im.rgb8, (size_t)n * 3, (void**)&d_rgb);      // returns to the context's free list with the scene
            if (rc < 0) { rt_scene_destroy(s); return rc; }
            cudaError_t e = ctx_alloc(c, (size_t)n * sizeof(float4), (void**)&d_tex);
            if (e != cudaSuccess) { rt_scene_destroy(s); return cuda_fail(e, "cudaMalloc(image texels)"); }
            s->allocations.push_back({d_tex, ((size_t)n * sizeof(float4) + 255) / 256 * 256});
            expand_image_kernel<<<(unsigned)((n + 255) / 256), 256>>>(d_rgb, n, d_lut, d_tex);
            e = cudaGetLastError();
            if (e != cudaSuccess) { rt_scene_destroy(s); return cuda_fail(e, "image upload"); }
            imgs[k].texels = d_tex;
            imgs[k].width = im.width;
            imgs[k].height = im.height;
        }
    }
    rc = upload_vec(s, imgs.data(), imgs.size() * sizeof(DevImage), &p);
    if (rc < 0) { rt_scene_destroy(s); return rc; }
    s->dev.images = static_cast<const DevImage*>(p);
    cudaError_t e = cudaDeviceSynchronize();    // the caller may free its host arrays when this returns
    if (e != cudaSuccess) { rt_scene_destroy(s); return cuda_fail(e, "rt_scene_upload"); }
    *out = s;
    return RT_OK;
}

int rt_scene_layout(const rt_scene_desc* desc, uint32_t layout_flags, rt_layout_info* out) {
    if (!desc || !out) return fail(RT_ERR_INVALID_ARGUMENT, "rt_scene_layout: null argument");
    CompiledScene cs;
    const char* err = nullptr;
    const int rc = compile_scene(desc, compile_options(nullptr, layout_flags), &cs, &err);
    if (rc < 0) return fail(rc, err ? err : "compile_scene failed");
    std::memset(out, 0, sizeof(*out));
    out->n_words = cs.n_world_words;
    for (int i = 0; i < cs.n_world_words;) {   // words of other ops can alias a header only by accident of their bits: walk op by op
        uint32_t hdr;
        std::memcpy(&hdr, &cs.ops[i].w, 4);
        const uint32_t kind = hdr_kind(hdr), flags = hdr_flags(hdr);
        switch (kind) {
            case OP_INNER: case OP_INNER_REF: out->n_inner++; break;
            case OP_SPHERE: out->n_sphere++; break;
            case OP_QUAD: out->n_quad++; break;
            case OP_XFORM_ENTER: out->n_xform++; break;
            case OP_XFORM_EXIT: break;
            case OP_MEDIUM: out->n_medium_in_stream++; break;
            case OP_BOX: out->n_box++; break;
            default: return fail(RT_ERR_INTERNAL, "rt_scene_layout: bad op in stream");
        }
        i += op_words(kind, flags);
    }
    out->n_medium_hoisted = (int32_t)cs.hoisted_media.size();
    out->n_precise_spheres = (int32_t)cs.precise.size() / 2;
    out->n_bvh = (int32_t)cs.bvh_hittable_ids.size();
    int64_t bytes = (int64_t)cs.ops.size() * 16 + (int64_t)(cs.materials.size() + cs.textures.size() + cs.perlin_vec.size()) * 16 +
                    (int64_t)cs.perlin_perm.size() + (int64_t)cs.precise.size() * 32;
    for (int k = 0; k < desc->n_images; ++k) bytes += (int64_t)desc->images[k].width * desc->images[k].height * 16;
    out->device_bytes = bytes;
    return RT_OK;
}

int rt_scene_ops_export(const rt_scene_desc* desc, uint32_t layout_flags, float* words, int64_t capacity_words, int64_t* n_total_words,
                        int32_t* n_world_words, int32_t* media_ops, int32_t* n_media, uint32_t* first_link) {
    if (!desc || !n_total_words) return fail(RT_ERR_INVALID_ARGUMENT, "rt_scene_ops_export: null argument");
    CompiledScene cs;
    const char* err = nullptr;
    const int rc = compile_scene(desc, compile_options(nullptr, layout_flags), &cs, &err);
    if (rc < 0) return fail(rc, err ? err : "compile_scene failed");
    *n_total_words = (int64_t)cs.ops.size();
    if (n_world_words) *n_world_words = cs.n_world_words;
    if (n_media) *n_media = (int32_t)cs.hoisted_media.size();
    if (media_ops) for (size_t k = 0; k < cs.hoisted_media.size() && k < (size_t)kMaxHoistedMedia; ++k) media_ops[k] = cs.hoisted_media[k];
    if (first_link) *first_link = cs.first_link;
    if (words) {
        if (capacity_words < (int64_t)cs.ops.size()) return fail(RT_ERR_OUT_OF_RANGE, "rt_scene_ops_export: capacity too small");
        std::memcpy(words, cs.ops.data(), cs.ops.size() * sizeof(F4));
    }
    return RT_OK;
}

void rt_scene_destroy(rt_scene* s) {
    if (!s) return;
    cudaSetDevice(s->device);
    rt_context* c = s->ctx;
    if (c) {
        cudaDeviceSynchronize();                  // nothing in flight may still read these blocks when they are handed out again
        size_t cached = 0;
        for (auto& b : c->free_blocks) cached += b.second;
        for (auto& b : s->allocations) {
            if (!c->destroyed && cached + b.second <= (size_t)2 << 30) { c->free_blocks.push_back(b); cached += b.second; }
            else cudaFree(b.first);
        }
        c->live_scenes--;
        if (c->destroyed && c->live_scenes == 0) context_free(c);
    } else {
        for (auto& b : s->allocations) cudaFree(b.first);
    }
    delete s;
}

static int launch_render(rt_context* c, const rt_scene* s, const rt_camera_desc* cam, int64_t sample_begin,
                         int64_t sample_count, uint64_t seed, void* d_sum_rgba, cudaStream_t stream, bool counting) {
    RenderParams prm;
    prm.scene = s->dev;
    prm.cam = make_dev_camera(*cam);
    prm.seed = seed;
    prm.sample_begin = sample_begin;
    prm.sample_count = (int)sample_count;
    if (cam->max_depth <= 0 && !counting) {
        const int64_t n_pix = (int64_t)cam->image_width * cam->image_height;
        black_samples_kernel<<<(unsigned)((n_pix + 255) / 256), 256, 0, stream>>>(static_cast<float4*>(d_sum_rgba), n_pix, (float)sample_count);
        CU(cudaGetLastError());
        return RT_OK;
    }
    prm.tiles_x = (prm.cam.width + kTileW - 1) / kTileW;
    prm.tiles_y = (prm.cam.height + kTileH - 1) / kTileH;
    // pool = tile x chunk samples; keep >= 64 pools per resident warp so the tail of the launch (warps running dry
    // while others still hold a pool) stays near 1%: short renders get smaller chunks than the default of 8
    const int64_t n_tiles = (int64_t)prm.tiles_x * prm.tiles_y;
    const int64_t resident_warps = (int64_t)c->sm_count * (kRenderThreads / 32);
    int chunk = c->chunk;
    while (chunk > 1 && n_tiles * ((sample_count + chunk - 1) / chunk) < resident_warps * 64) chunk >>= 1;
    prm.chunk = chunk;
    prm.n_chunks = (int)((sample_count + chunk - 1) / chunk);
    if ((uint64_t)n_tiles * (uint64_t)prm.n_chunks >= 0xffffffffull) return fail(RT_ERR_INVALID_ARGUMENT, "rt_render_accumulate: too many work items; split the sample range");
    prm.sum = static_cast<float4*>(d_sum_rgba);
    const int slot = (int)(c->launches % kLaunchSlots);
    prm.work_counter = c->d_counters + slot;
    prm.stats = c->d_stats + (size_t)slot * K_NUM;
    prm.first_link = s->compiled.first_link;
    prm.shade_min = c->shade_min;
    prm.slab_fast = c->slab_fast;
    prm.sphere_reps = c->sphere_reps;
    prm.quad_reps = c->quad_reps;
    prm.path_counter = nullptr;
    prm.total_paths = 0;
    prm.xchg_min = prm.slab_drop = prm.min_trav = 0;
    CU(cudaMemsetAsync(prm.stats, 0, K_NUM * sizeof(unsigned long long), stream));
#ifdef RT_B200_DEV
    if (c->queue_kernel && !counting) {
        // render_q.cuh: paths numbered globally, 32 per (tile, sample)
        if ((uint64_t)n_tiles * (uint64_t)sample_count >= 0xffffffffull) return fail(RT_ERR_INVALID_ARGUMENT, "rt_render_accumulate: too many work items; split the sample range");
        prm.path_counter = c->d_path_counters + slot;
        prm.total_paths = (unsigned long long)n_tiles * (unsigned long long)sample_count * 32ull;
        prm.xchg_min = c->xchg_min;
        prm.slab_drop = c->slab_drop;
        prm.min_trav = c->min_trav;
        QSmem lay = q_smem_layout(s->ops_bytes, s->dev.n_perlin);
        const bool in_smem = c->ops_in_smem && !s->ops_in_global && lay.total <= c->smem_optin;
        if (!in_smem) lay = q_smem_layout(0, s->dev.n_perlin);
        if (lay.total > c->smem_optin) return fail(RT_ERR_INTERNAL, "render kernel: path slots do not fit in shared memory");
        prm.ops_bytes = in_smem ? s->ops_bytes : 0u;
        CU(cudaMemsetAsync(prm.path_counter, 0, sizeof(unsigned long long), stream));
        q_kernel(in_smem)<<<c->sm_count, kQThreads, lay.total, stream>>>(prm);
    } else
#endif
    {
        // the op stream rides in shared memory when it fits beside the per-thread path state and the Perlin tables
        MkSmem lay = mk_smem_layout(s->ops_bytes, s->dev.n_perlin);
        const bool in_smem = c->ops_in_smem && !s->ops_in_global && lay.total <= c->smem_optin;
        if (!in_smem) lay = mk_smem_layout(0, s->dev.n_perlin);
        if (lay.total > c->smem_optin) return fail(RT_ERR_INTERNAL, "render kernel: per-thread state does not fit in shared memory");
        prm.ops_bytes = in_smem ? s->ops_bytes : 0u;
        CU(cudaMemsetAsync(prm.work_counter, 0, sizeof(unsigned int), stream));
        render_fn fn = mk_kernel(counting, in_smem, s->generic_kernel ? FEAT_ALL : (s->features | (prm.cam.defocus ? FEAT_DEFOCUS : 0u)));
        fn<<<c->sm_count, kRenderThreads, lay.total, stream>>>(prm);
    }
    CU(cudaGetLastError());
    c->last_slot = slot;
    c->launches += 1;
    return RT_OK;
}

static int check_render_args(rt_context* c, const rt_scene* s, const rt_camera_desc* cam, int64_t sample_begin, int64_t sample_count,
                             const void* buf, const char* who) {
    if (!c || !s || !cam || !buf) return fail(RT_ERR_INVALID_ARGUMENT, std::string(who) + ": null argument");
    if (s->ctx != c) return fail(RT_ERR_INVALID_ARGUMENT, std::string(who) + ": the scene was uploaded through another context");
    if (sample_count < 0 || sample_count > 0x7fffffff) return fail(RT_ERR_INVALID_ARGUMENT, std::string(who) + ": bad sample_count");
    // the RNG key takes the sample index as 32 bits (path_key): a range past 2^32 would silently reuse keys
    if (sample_begin < 0 || sample_begin + sample_count > (int64_t)1 << 32)
        return fail(RT_ERR_INVALID_ARGUMENT, std::string(who) + ": sample range must lie in [0, 2^32)");
    if (cam->image_width <= 0 || cam->image_height <= 0 || cam->image_width * cam->image_height > 0x7fffffff)
        return fail(RT_ERR_INVALID_ARGUMENT, std::string(who) + ": bad image size");
    return RT_OK;
}

int rt_render_accumulate(rt_context* c, const rt_scene* s, const rt_camera_desc* cam, int64_t sample_begin,
                         int64_t sample_count, uint64_t seed, void* d_sum_rgba, void* stream_) {
    int rc = check_render_args(c, s, cam, sample_begin, sample_count, d_sum_rgba, "rt_render_accumulate");
    if (rc < 0) return rc;
    if ((reinterpret_cast<uintptr_t>(d_sum_rgba) & 15u) != 0)   // red.global.add.v4.f32
        return fail(RT_ERR_INVALID_ARGUMENT, "rt_render_accumulate: the framebuffer must be 16-byte aligned");
    if (sample_count == 0) return RT_OK;
    CU(cudaSetDevice(c->device));
    return launch_render(c, s, cam, sample_begin, sample_count, seed, d_sum_rgba, static_cast<cudaStream_t>(stream_), false);
}

int rt_render_count_ops(rt_context* c, const rt_scene* s, const rt_camera_desc* cam, int64_t sample_begin, int64_t sample_count,
                        uint64_t seed, uint64_t* counters, int capacity, const char** names_csv) {
    int rc = check_render_args(c, s, cam, sample_begin, sample_count, counters, "rt_render_count_ops");
    if (rc < 0) return rc;
    if (capacity < (int)K_NUM) return fail(RT_ERR_OUT_OF_RANGE, "rt_render_count_ops: capacity too small");
    if (cam->max_depth <= 0) return fail(RT_ERR_INVALID_ARGUMENT, "rt_render_count_ops: nothing to count at max_depth <= 0");
    if (names_csv) *names_csv = kCounterNames;
    CU(cudaSetDevice(c->device));
    const size_t n = (size_t)cam->image_width * (size_t)cam->image_height;
    float4* scratch = nullptr;
    CU(cudaMalloc(&scratch, n * sizeof(float4)));
    cudaMemset(scratch, 0, n * sizeof(float4));
    unsigned long long h[K_NUM] = {0};
    cudaError_t e = cudaSuccess;
    if (sample_count > 0) {
        rc = launch_render(c, s, cam, sample_begin, sample_count, seed, scratch, nullptr, true);
        if (rc == RT_OK) e = cudaMemcpy(h, c->d_stats + (size_t)c->last_slot * K_NUM, sizeof(h), cudaMemcpyDeviceToHost);
    }
    cudaFree(scratch);
    if (rc < 0) return rc;
    if (e != cudaSuccess) return cuda_fail(e, "rt_render_count_ops");
    for (int k = 0; k < (int)K_NUM; ++k) counters[k] = h[k];
    return (int)K_NUM;
}

int rt_render_get_stats(rt_context* c, rt_render_stats* out) {
    if (!c || !out) return fail(RT_ERR_INVALID_ARGUMENT, "rt_render_get_stats: null argument");
    CU(cudaSetDevice(c->device));
    CU(cudaDeviceSynchronize());    // the last launch may sit on a non-blocking stream: wait for the whole device
    unsigned long long h[2] = {0, 0};
    CU(cudaMemcpy(h, c->d_stats + (size_t)c->last_slot * K_NUM, sizeof(h), cudaMemcpyDeviceToHost));
    out->paths = h[0];
    out->segments = h[1];
    out->kernel_launches = c->launches;
    out->last_kernel_ms = 0.0f;
    return RT_OK;
}

static int render_into_context_fb(rt_context* c, const rt_scene* s, const rt_camera_desc* cam, int64_t sample_begin,
                                  int64_t sample_count, uint64_t seed, const char* who) {
    if (!c || !s || !cam) return fail(RT_ERR_INVALID_ARGUMENT, std::string(who) + ": null argument");
    if (cam->image_width <= 0 || cam->image_height <= 0) return fail(RT_ERR_INVALID_ARGUMENT, std::string(who) + ": bad image size");
    CU(cudaSetDevice(c->device));
    const size_t n = (size_t)cam->image_width * (size_t)cam->image_height;
    if (c->fb_pixels < n) {
        cudaFree(c->d_fb);
        c->d_fb = nullptr;
        c->fb_pixels = 0;
        CU(cudaMalloc(&c->d_fb, n * sizeof(float4)));
        c->fb_pixels = n;
    }
    CU(cudaMemsetAsync(c->d_fb, 0, n * sizeof(float4), 0));
    return rt_render_accumulate(c, s, cam, sample_begin, sample_count, seed, c->d_fb, nullptr);
}

int rt_render(rt_context* c, const rt_scene* s, const rt_camera_desc* cam, int64_t sample_begin, int64_t sample_count,
              uint64_t seed, float* host_sum_rgba) {
    if (!host_sum_rgba) return fail(RT_ERR_INVALID_ARGUMENT, "rt_render: null argument");
    const int rc = render_into_context_fb(c, s, cam, sample_begin, sample_count, seed, "rt_render");
    if (rc < 0) return rc;
    CU(cudaMemcpy(host_sum_rgba, c->d_fb, (size_t)cam->image_width * (size_t)cam->image_height * sizeof(float4), cudaMemcpyDeviceToHost));
    return RT_OK;
}

int rt_render_rgb8(rt_context* c, const rt_scene* s, const rt_camera_desc* cam, int64_t sample_begin, int64_t sample_count,
                   uint64_t seed, uint8_t* host_rgb8) {
    if (!host_rgb8) return fail(RT_ERR_INVALID_ARGUMENT, "rt_render_rgb8: null argument");
    const int rc = render_into_context_fb(c, s, cam, sample_begin, sample_count, seed, "rt_render_rgb8");
    if (rc < 0) return rc;
    return rt_finalize_rgb8(c, c->d_fb, cam->image_width * cam->image_height, (double)sample_count, host_rgb8, nullptr);
}

// ---- multi-GPU in one process: NCCL through dlopen (no link-time dependency: a single-GPU host needs no libnccl)
namespace {
struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*Reduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool tried = false;
};
NcclApi g_nccl;
std::map<std::vector<int>, std::vector<ncclComm_t>> g_comms;   // device list -> communicators (kept for the life of the process)

int load_nccl() {
    if (g_nccl.CommInitAll) return RT_OK;
    if (!g_nccl.tried) {
        g_nccl.tried = true;
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
            g_nccl.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (g_nccl.handle) break;
        }
        if (g_nccl.handle) {
            g_nccl.CommInitAll = reinterpret_cast<decltype(g_nccl.CommInitAll)>(dlsym(g_nccl.handle, "ncclCommInitAll"));
            g_nccl.Reduce = reinterpret_cast<decltype(g_nccl.Reduce)>(dlsym(g_nccl.handle, "ncclReduce"));
            g_nccl.GroupStart = reinterpret_cast<decltype(g_nccl.GroupStart)>(dlsym(g_nccl.handle, "ncclGroupStart"));
            g_nccl.GroupEnd = reinterpret_cast<decltype(g_nccl.GroupEnd)>(dlsym(g_nccl.handle, "ncclGroupEnd"));
            g_nccl.GetErrorString = reinterpret_cast<decltype(g_nccl.GetErrorString)>(dlsym(g_nccl.handle, "ncclGetErrorString"));
        }
    }
    if (!g_nccl.CommInitAll || !g_nccl.Reduce || !g_nccl.GroupStart || !g_nccl.GroupEnd)
        return fail(RT_ERR_UNSUPPORTED, "rt_render_multi: libnccl.so.2 could not be loaded; more than one GPU needs NCCL");
    return RT_OK;
}
int nccl_fail(ncclResult_t r, const char* what) {
    return fail(RT_ERR_CUDA, std::string(what) + ": " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "NCCL error"));
}
}  // namespace

int rt_render_multi(rt_context* const* ctxs, const rt_scene* const* scenes, int n, const rt_camera_desc* cam, int64_t sample_begin,
                    int64_t sample_count, uint64_t seed, const double* weights, float* host_sum_rgba, int64_t* shares_out) {
    if (!ctxs || !scenes || !cam || !host_sum_rgba || n <= 0) return fail(RT_ERR_INVALID_ARGUMENT, "rt_render_multi: null argument");
    if (sample_count < 0) return fail(RT_ERR_INVALID_ARGUMENT, "rt_render_multi: bad sample_count");
    std::vector<int> devices;
    for (int i = 0; i < n; ++i) {
        if (!ctxs[i] || !scenes[i]) return fail(RT_ERR_INVALID_ARGUMENT, "rt_render_multi: null context / scene");
        if (scenes[i]->ctx != ctxs[i]) return fail(RT_ERR_INVALID_ARGUMENT, "rt_render_multi: scenes[i] must be uploaded through ctxs[i]");
        for (int d : devices) if (d == ctxs[i]->device) return fail(RT_ERR_INVALID_ARGUMENT, "rt_render_multi: one context per device");
        devices.push_back(ctxs[i]->device);
    }
    // the split: contiguous sample ranges, largest remainders take the leftover samples (distributed.py does the same)
    std::vector<int64_t> share((size_t)n, 0);
    {
        double total = 0.0;
        for (int i = 0; i < n; ++i) {
            const double w = weights ? weights[i] : 1.0;
            if (!(w >= 0.0)) return fail(RT_ERR_INVALID_ARGUMENT, "rt_render_multi: weights must be >= 0");
            total += w;
        }
        if (!(total > 0.0)) return fail(RT_ERR_INVALID_ARGUMENT, "rt_render_multi: weights sum to zero");
        std::vector<std::pair<double, int>> rem;
        int64_t given = 0;
        for (int i = 0; i < n; ++i) {
            const double exact = (double)sample_count * (weights ? weights[i] : 1.0) / total;
            share[i] = (int64_t)exact;
            given += share[i];
            rem.push_back({exact - (double)share[i], -i});
        }
        std::sort(rem.begin(), rem.end(), [](const auto& a, const auto& b) { return a > b; });
        for (int64_t k = 0; k < sample_count - given; ++k) share[(size_t)(-rem[(size_t)k % rem.size()].second)]++;
    }
    if (shares_out) for (int i = 0; i < n; ++i) shares_out[i] = share[i];
    if (n == 1) return rt_render(ctxs[0], scenes[0], cam, sample_begin, sample_count, seed, host_sum_rgba);
    int rc = load_nccl();
    if (rc < 0) return rc;
    auto it = g_comms.find(devices);
    if (it == g_comms.end()) {
        std::vector<ncclComm_t> comms((size_t)n);
        const ncclResult_t r = g_nccl.CommInitAll(comms.data(), n, devices.data());
        if (r != ncclSuccess) return nccl_fail(r, "ncclCommInitAll");
        it = g_comms.emplace(devices, comms).first;
    }
    const size_t npx = (size_t)cam->image_width * (size_t)cam->image_height;
    int64_t begin = sample_begin;
    for (int i = 0; i < n; ++i) {          // every device starts on its share; the launches are asynchronous
        rt_context* c = ctxs[i];
        CU(cudaSetDevice(c->device));
        if (c->fb_pixels < npx) {
            cudaFree(c->d_fb);
            c->d_fb = nullptr;
            c->fb_pixels = 0;
            CU(cudaMalloc(&c->d_fb, npx * sizeof(float4)));
            c->fb_pixels = npx;
        }
        CU(cudaMemsetAsync(c->d_fb, 0, npx * sizeof(float4), 0));
        rc = rt_render_accumulate(c, scenes[i], cam, begin, share[i], seed, c->d_fb, nullptr);
        if (rc < 0) return rc;
        begin += share[i];
    }
    ncclResult_t r = g_nccl.GroupStart();   // partial sums -> the first device, over NVLink
    for (int i = 0; i < n && r == ncclSuccess; ++i) {
        cudaSetDevice(ctxs[i]->device);
        r = g_nccl.Reduce(ctxs[i]->d_fb, ctxs[0]->d_fb, npx * 4, ncclFloat, ncclSum, 0, it->second[(size_t)i], 0);
    }
    const ncclResult_t r2 = g_nccl.GroupEnd();
    if (r != ncclSuccess) return nccl_fail(r, "ncclReduce");
    if (r2 != ncclSuccess) return nccl_fail(r2, "ncclGroupEnd");
    for (int i = n - 1; i >= 1; --i) {
        CU(cudaSetDevice(ctxs[i]->device));
        CU(cudaStreamSynchronize(0));
    }
    CU(cudaSetDevice(ctxs[0]->device));
    CU(cudaMemcpy(host_sum_rgba, ctxs[0]->d_fb, npx * sizeof(float4), cudaMemcpyDeviceToHost));
    return RT_OK;
}

int rt_finalize_rgb8(rt_context* c, const void* d_sum_rgba, int64_t n_pixels, double spp, uint8_t* host_rgb8, void* stream_) {
    if (!c || !d_sum_rgba || !host_rgb8) return fail(RT_ERR_INVALID_ARGUMENT, "rt_finalize_rgb8: null argument");
    if (n_pixels <= 0) return RT_OK;
    CU(cudaSetDevice(c->device));
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const size_t bytes = (size_t)n_pixels * 3;
    if (c->rgb8_bytes < bytes) {
        cudaFree(c->d_rgb8);
        c->d_rgb8 = nullptr;
        c->rgb8_bytes = 0;
        CU(cudaMalloc(&c->d_rgb8, bytes));
        c->rgb8_bytes = bytes;
    }
    // on the caller's stream: ordered after the rt_render_accumulate calls that filled the framebuffer there
    finalize_kernel<<<(unsigned)((n_pixels + 255) / 256), 256, 0, stream>>>(static_cast<const float4*>(d_sum_rgba), n_pixels,
                                                                          spp > 0 ? (float)(1.0 / spp) : 0.0f, spp > 0 ? 0 : 1, c->d_rgb8);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(host_rgb8, c->d_rgb8, bytes, cudaMemcpyDeviceToHost, stream));
    CU(cudaStreamSynchronize(stream));
    return RT_OK;
}

int rt_hit_batch(rt_context* c, const rt_scene* s, const rt_ray_desc* rays, int64_t n, double t_min, double t_max,
                 uint64_t seed, rt_hit_desc* out) {
    if (!c || !s) return fail(RT_ERR_INVALID_ARGUMENT, "rt_hit_batch: null argument");
    if (n < 0) return fail(RT_ERR_INVALID_ARGUMENT, "rt_hit_batch: negative count");
    if (n == 0) return RT_OK;
    if (!rays || !out) return fail(RT_ERR_INVALID_ARGUMENT, "rt_hit_batch: null buffer");
    CU(cudaSetDevice(c->device));
    std::vector<DevRayIn> h_in((size_t)n);
    for (int64_t k = 0; k < n; ++k) {
        DevRayIn& r = h_in[k];
        r.ox = (float)rays[k].origin[0]; r.oy = (float)rays[k].origin[1]; r.oz = (float)rays[k].origin[2];
        r.dx = (float)rays[k].direction[0]; r.dy = (float)rays[k].direction[1]; r.dz = (float)rays[k].direction[2];
        r.time = (float)rays[k].time; r.pad = 0.0f;
    }
    DevRayIn* d_in = nullptr;
    DevHitOut* d_out = nullptr;
    CU(cudaMalloc(&d_in, (size_t)n * sizeof(DevRayIn)));
    cudaError_t e = cudaMalloc(&d_out, (size_t)n * sizeof(DevHitOut));
    if (e != cudaSuccess) { cudaFree(d_in); return cuda_fail(e, "cudaMalloc(hits)"); }
    std::vector<DevHitOut> h_out((size_t)n);
    e = cudaMemcpy(d_in, h_in.data(), (size_t)n * sizeof(DevRayIn), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        hit_kernel<<<(unsigned)((n + 127) / 128), 128>>>(s->dev, d_in, n, (float)t_min, (float)t_max, seed, d_out);
        e = cudaMemcpy(h_out.data(), d_out, (size_t)n * sizeof(DevHitOut), cudaMemcpyDeviceToHost);
    }
    cudaFree(d_in);
    cudaFree(d_out);
    if (e != cudaSuccess) return cuda_fail(e, "rt_hit_batch");
    for (int64_t k = 0; k < n; ++k) {
        const DevHitOut& h = h_out[k];
        rt_hit_desc& o = out[k];
        std::memset(&o, 0, sizeof(o));
        o.hit = h.hit; o.front_face = h.front_face; o.prim_id = h.prim; o.mat_id = h.mat;
        o.t = h.t; o.p[0] = h.px; o.p[1] = h.py; o.p[2] = h.pz;
        o.normal[0] = h.nx; o.normal[1] = h.ny; o.normal[2] = h.nz;
        o.u = h.u; o.v = h.v;
    }
    return RT_OK;
}

int rt_texture_batch(rt_context* c, const rt_scene* s, int tex, const double* uvp, int64_t n, double* rgb_out) {
    if (!c || !s) return fail(RT_ERR_INVALID_ARGUMENT, "rt_texture_batch: null argument");
    if (tex < 0 || (size_t)tex * 2 >= s->compiled.textures.size()) return fail(RT_ERR_OUT_OF_RANGE, "rt_texture_batch: unknown texture id");
    if (n <= 0) return n == 0 ? RT_OK : fail(RT_ERR_INVALID_ARGUMENT, "rt_texture_batch: negative count");
    if (!uvp || !rgb_out) return fail(RT_ERR_INVALID_ARGUMENT, "rt_texture_batch: null buffer");
    CU(cudaSetDevice(c->device));
    std::vector<float> h_in((size_t)n * 5), h_out((size_t)n * 3);
    for (int64_t k = 0; k < n * 5; ++k) h_in[k] = (float)uvp[k];
    float *d_in = nullptr, *d_out = nullptr;
    CU(cudaMalloc(&d_in, h_in.size() * sizeof(float)));
    cudaError_t e = cudaMalloc(&d_out, h_out.size() * sizeof(float));
    if (e != cudaSuccess) { cudaFree(d_in); return cuda_fail(e, "cudaMalloc"); }
    e = cudaMemcpy(d_in, h_in.data(), h_in.size() * sizeof(float), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        cudaFuncSetAttribute(texture_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)perlin_smem_bytes());
        texture_kernel<<<(unsigned)((n + 127) / 128), 128, perlin_smem_bytes()>>>(s->dev, tex, d_in, n, d_out);
        e = cudaMemcpy(h_out.data(), d_out, h_out.size() * sizeof(float), cudaMemcpyDeviceToHost);
    }
    cudaFree(d_in);
    cudaFree(d_out);
    if (e != cudaSuccess) return cuda_fail(e, "rt_texture_batch");
    for (int64_t k = 0; k < n * 3; ++k) rgb_out[k] = h_out[k];
    return RT_OK;
}

int rt_get_ray_batch(rt_context* c, const rt_camera_desc* cam, const int64_t* pixel_index, const int64_t* sample_index,
                     int64_t n, uint64_t seed, rt_ray_desc* out) {
    if (!c || !cam) return fail(RT_ERR_INVALID_ARGUMENT, "rt_get_ray_batch: null argument");
    if (n <= 0) return n == 0 ? RT_OK : fail(RT_ERR_INVALID_ARGUMENT, "rt_get_ray_batch: negative count");
    if (!pixel_index || !sample_index || !out) return fail(RT_ERR_INVALID_ARGUMENT, "rt_get_ray_batch: null buffer");
    CU(cudaSetDevice(c->device));
    int64_t *d_pix = nullptr, *d_smp = nullptr;
    DevRayIn* d_out = nullptr;
    CU(cudaMalloc(&d_pix, (size_t)n * 8));
    cudaError_t e = cudaMalloc(&d_smp, (size_t)n * 8);
    if (e == cudaSuccess) e = cudaMalloc(&d_out, (size_t)n * sizeof(DevRayIn));
    std::vector<DevRayIn> h((size_t)n);
    if (e == cudaSuccess) e = cudaMemcpy(d_pix, pixel_index, (size_t)n * 8, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_smp, sample_index, (size_t)n * 8, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        get_ray_kernel<<<(unsigned)((n + 127) / 128), 128>>>(make_dev_camera(*cam), d_pix, d_smp, n, seed, d_out);
        e = cudaMemcpy(h.data(), d_out, (size_t)n * sizeof(DevRayIn), cudaMemcpyDeviceToHost);
    }
    cudaFree(d_pix); cudaFree(d_smp); cudaFree(d_out);
    if (e != cudaSuccess) return cuda_fail(e, "rt_get_ray_batch");
    for (int64_t k = 0; k < n; ++k) {
        out[k].origin[0] = h[k].ox; out[k].origin[1] = h[k].oy; out[k].origin[2] = h[k].oz;
        out[k].direction[0] = h[k].dx; out[k].direction[1] = h[k].dy; out[k].direction[2] = h[k].dz;
        out[k].time = h[k].time;
    }
    return RT_OK;
}

int rt_scatter_batch(rt_context* c, const rt_scene* s, const rt_ray_desc* rays_in, const rt_hit_desc* hits, int64_t n, uint64_t seed,
                     const uint32_t* pixel, const uint32_t* sample, uint32_t segment, rt_scatter_desc* out) {
    if (!c || !s) return fail(RT_ERR_INVALID_ARGUMENT, "rt_scatter_batch: null argument");
    if (n <= 0) return n == 0 ? RT_OK : fail(RT_ERR_INVALID_ARGUMENT, "rt_scatter_batch: negative count");
    if (!rays_in || !hits || !pixel || !sample || !out) return fail(RT_ERR_INVALID_ARGUMENT, "rt_scatter_batch: null buffer");
    const int n_mats = (int)(s->compiled.materials.size() / 2);
    std::vector<DevScatterIn> h_in((size_t)n);
    for (int64_t k = 0; k < n; ++k) {
        if (hits[k].mat_id < 0 || hits[k].mat_id >= n_mats) return fail(RT_ERR_OUT_OF_RANGE, "rt_scatter_batch: hit record names an unknown material");
        DevScatterIn& r = h_in[k];
        r.ox = (float)rays_in[k].origin[0]; r.oy = (float)rays_in[k].origin[1]; r.oz = (float)rays_in[k].origin[2];
        r.dx = (float)rays_in[k].direction[0]; r.dy = (float)rays_in[k].direction[1]; r.dz = (float)rays_in[k].direction[2];
        r.time = (float)rays_in[k].time;
        r.px = (float)hits[k].p[0]; r.py = (float)hits[k].p[1]; r.pz = (float)hits[k].p[2];
        r.nx = (float)hits[k].normal[0]; r.ny = (float)hits[k].normal[1]; r.nz = (float)hits[k].normal[2];
        r.t = (float)hits[k].t; r.u = (float)hits[k].u; r.v = (float)hits[k].v;
        r.front_face = hits[k].front_face; r.mat = hits[k].mat_id;
        r.pixel = pixel[k]; r.sample = sample[k];
    }
    CU(cudaSetDevice(c->device));
    DevScatterIn* d_in = nullptr;
    DevScatterOut* d_out = nullptr;
    CU(cudaMalloc(&d_in, (size_t)n * sizeof(DevScatterIn)));
    cudaError_t e = cudaMalloc(&d_out, (size_t)n * sizeof(DevScatterOut));
    if (e != cudaSuccess) { cudaFree(d_in); return cuda_fail(e, "cudaMalloc(scatter)"); }
    std::vector<DevScatterOut> h_out((size_t)n);
    e = cudaMemcpy(d_in, h_in.data(), (size_t)n * sizeof(DevScatterIn), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        cudaFuncSetAttribute(scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)perlin_smem_bytes());
        scatter_kernel<<<(unsigned)((n + 127) / 128), 128, perlin_smem_bytes()>>>(s->dev, d_in, n, seed, segment, d_out);
        e = cudaMemcpy(h_out.data(), d_out, (size_t)n * sizeof(DevScatterOut), cudaMemcpyDeviceToHost);
    }
    cudaFree(d_in);
    cudaFree(d_out);
    if (e != cudaSuccess) return cuda_fail(e, "rt_scatter_batch");
    for (int64_t k = 0; k < n; ++k) {
        const DevScatterOut& h = h_out[k];
        rt_scatter_desc& o = out[k];
        std::memset(&o, 0, sizeof(o));
        o.scattered = h.scattered;
        o.ray_out.origin[0] = h.ox; o.ray_out.origin[1] = h.oy; o.ray_out.origin[2] = h.oz;
        o.ray_out.direction[0] = h.dx; o.ray_out.direction[1] = h.dy; o.ray_out.direction[2] = h.dz;
        o.ray_out.time = h.time;
        o.attenuation[0] = h.ar; o.attenuation[1] = h.ag; o.attenuation[2] = h.ab;
        o.emitted[0] = h.er; o.emitted[1] = h.eg; o.emitted[2] = h.eb;
    }
    return RT_OK;
}

int rt_bvh_export(const rt_scene* s, int bvh_hittable, int32_t* object_of_node, int32_t capacity, int32_t* n_out) {
    if (!s || !n_out) return fail(RT_ERR_INVALID_ARGUMENT, "rt_bvh_export: null argument");
    const CompiledScene& cs = s->compiled;
    for (size_t k = 0; k < cs.bvh_hittable_ids.size(); ++k) {
        if (cs.bvh_hittable_ids[k] != bvh_hittable) continue;
        const auto& pre = cs.bvh_preorder_objects[k];
        *n_out = (int32_t)pre.size();
        if (object_of_node) {
            if (capacity < (int32_t)pre.size()) return fail(RT_ERR_OUT_OF_RANGE, "rt_bvh_export: capacity too small");
            std::memcpy(object_of_node, pre.data(), pre.size() * sizeof(int32_t));
        }
        return RT_OK;
    }
    return fail(RT_ERR_OUT_OF_RANGE, "rt_bvh_export: that hittable is not a BVH reachable from the world");
}

int rt_measure_fp32_peak(rt_context* c, double* tflops_out) {
    if (!c || !tflops_out) return fail(RT_ERR_INVALID_ARGUMENT, "rt_measure_fp32_peak: null argument");
    CU(cudaSetDevice(c->device));
    const int threads = 256, blocks = c->sm_count * 8, iters = 1 << 16;
    float* d = nullptr;
    CU(cudaMalloc(&d, (size_t)threads * blocks * sizeof(float)));
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    fma_peak_kernel<<<blocks, threads>>>(d, 1024);   // warm-up
    float best_ms = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(a);
        fma_peak_kernel<<<blocks, threads>>>(d, iters);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms = 0.0f;
        cudaEventElapsedTime(&ms, a, b);
        if (ms < best_ms) best_ms = ms;
    }
    cudaEventDestroy(a); cudaEventDestroy(b);
    cudaError_t e = cudaDeviceSynchronize();
    cudaFree(d);
    if (e != cudaSuccess) return cuda_fail(e, "rt_measure_fp32_peak");
    const double flops = 2.0 * 8.0 * (double)iters * threads * blocks;
    *tflops_out = flops / (best_ms * 1e-3) / 1e12;
    return RT_OK;
}

}  // extern "C"
