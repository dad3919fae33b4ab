// Kernels + device half of the C ABI (include/rt_b200.h).
//
// K1 render_kernel_v3 (render_v3.cuh): persistent megakernel for the parallel loop of renderer.rs:26-49. One
//                    warp owns a pool of (8x4 pixel tile) x (sample chunk) paths; a lane whose path ends takes
//                    the next path of the pool in the same iteration (ballot/popc ranking); lanes regroup by op
//                    class through a warp vote. Radiance sums go to a float4 framebuffer, one vector reduction
//                    per path. Template parameter FOLD: cube primitives share the box-test instruction stream,
//                    picked per scene (kFoldBoxMin). render_kernel below is the first, whole-segment-per-iteration
//                    form (A/B only); wf_shade_kernel / wf_extend_kernel (render_v4.cuh) are the same loop as a
//                    wavefront over a pool of in-flight paths (RT_B200_KERNEL=4: measured, slower, not the product).
// K2 hit_kernel      Hittable::hit on a ray batch (parity).
// K3 finalize_kernel color_to_rgb(sum/spp) (color.rs:12-19, renderer.rs:55-58).
// K4 texture_kernel / get_ray_kernel (parity), expand_image_kernel (upload), fma_peak_kernel.
#include "rt_kernels.cuh"

#include "../host/host_common.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

using namespace rtdev;
using rt_host::fail;

namespace {

#ifdef RT_OPT_BLOCK
constexpr int kBlockThreads = RT_OPT_BLOCK;
#else
constexpr int kBlockThreads = 128;
#endif
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
    int first_class;             // OpClass of op 0
    int shade_min;               // the shade class may win the vote once this many lanes wait for it
    int slab_fast;               // v3: lanes in the slab class that skip the full vote
    int slab_reps, sphere_reps;  // v3: consecutive ops a class may run per vote (slab: compile-time kSlabReps)
    int slab_exit;               // v3: the slab repetitions stop (and the warp votes again) once fewer lanes than this remain in the class
    int sphere_min, box_min, quad_min;   // v3: quorum at which a minority class runs ahead of the slab class
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

extern __shared__ float4 dyn_smem[];

__global__ void __launch_bounds__(kBlockThreads) render_kernel(const RenderParams prm) {
    float4* sh_vec = dyn_smem;
    uint8_t* sh_perm = reinterpret_cast<uint8_t*>(dyn_smem + kMaxPerlinShared * 256);
    stage_perlin(prm.scene, sh_vec, sh_perm);
    PerlinShared P{sh_vec, sh_perm};
    const DevScene& S = prm.scene;
    const DevCamera& C = prm.cam;

    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt_mask = (1u << lane) - 1u;
    const unsigned n_tiles = (unsigned)(prm.tiles_x * prm.tiles_y);
    const unsigned n_items = n_tiles * (unsigned)prm.n_chunks;

    // warp-uniform pool state
    int pool_next = 0, pool_size = 0;
    int tile_x0 = 0, tile_y0 = 0, tile_w = 1, tile_n = 1;
    int64_t pool_sample0 = 0;
    bool no_more = false;

    // per-lane path state
    bool active = false;
    Ray ray;
    float3 L, T;
    int depth = 0, origin = -1, pix = 0;
    uint4 key;
    unsigned long long n_paths = 0, n_segments = 0;

    for (;;) {
        const unsigned need = __ballot_sync(0xffffffffu, !active);
        if (need) {
            if (pool_next >= pool_size && !no_more) {
                unsigned item = 0;
                if (lane == 0) item = atomicAdd(prm.work_counter, 1u);
                item = __shfl_sync(0xffffffffu, item, 0);
                if (item >= n_items) {
                    no_more = true;
                } else {
                    const unsigned chunk_idx = item / n_tiles, tile = item % n_tiles;   // chunk-major: concurrent warps spread over tiles
                    tile_x0 = (int)(tile % (unsigned)prm.tiles_x) * kTileW;
                    tile_y0 = (int)(tile / (unsigned)prm.tiles_x) * kTileH;
                    tile_w = min(kTileW, C.width - tile_x0);
                    const int tile_h = min(kTileH, C.height - tile_y0);
                    tile_n = tile_w * tile_h;
                    const int s0 = (int)chunk_idx * prm.chunk;
                    const int ns = min(prm.chunk, prm.sample_count - s0);
                    pool_sample0 = prm.sample_begin + s0;
                    pool_size = tile_n * ns;
                    pool_next = 0;
                }
            }
            if (!active) {
                const int idx = pool_next + __popc(need & lt_mask);
                if (idx < pool_size) {
                    const int pv = idx % tile_n, sv = idx / tile_n;
                    const int px = tile_x0 + pv % tile_w, py = tile_y0 + pv / tile_w;
                    pix = py * C.width + px;                                         // renderer.rs:32-33
                    key = path_key(prm.seed, (uint32_t)pix, (uint32_t)(pool_sample0 + sv));
                    ray = camera_ray(C, px, py, key);
                    L = f3(0.0f, 0.0f, 0.0f);
                    T = f3(1.0f, 1.0f, 1.0f);
                    depth = 0;
                    origin = -1;
                    active = true;
                    ++n_paths;
                }
            }
            pool_next = min(pool_size, pool_next + __popc(need));
        }
        if (__ballot_sync(0xffffffffu, active) == 0u) {
            if (no_more) break;
            continue;
        }
        if (active) {
            // one bounce of ray_color (renderer.rs:139-155), iteratively
            ++n_segments;
            Best best;
            traverse<true>(S, 0, S.n_words, ray, 0.001f, __int_as_float(0x7f800000), best, origin, key, (uint32_t)depth);
            bool alive;
            if (best.op < 0) {
                L = L + T * C.background;                                            // renderer.rs:152-153
                alive = false;
            } else {
                HitRec h;
                finalize_hit(S, ray, best, h);
                alive = shade(S, P, ray, h, key, (uint32_t)depth, L, T);
                origin = h.origin;
                ++depth;
                if (depth >= C.max_depth) alive = false;                             // depth <= 0 returns black (:140-142)
            }
            if (!alive) {
                red_add_f4(prm.sum + pix, L.x, L.y, L.z, 1.0f);                      // avg_color += new_color (:39)
                active = false;
            }
        }
    }
    // stats: one atomic per warp
    for (int off = 16; off > 0; off >>= 1) {
        n_paths += __shfl_down_sync(0xffffffffu, n_paths, off);
        n_segments += __shfl_down_sync(0xffffffffu, n_segments, off);
    }
    if (lane == 0) {
        atomicAdd(prm.stats + 0, n_paths);
        atomicAdd(prm.stats + 1, n_segments);
    }
}

#include "render_v3.cuh"
#include "render_v4.cuh"

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
    traverse<true>(S, 0, S.n_words, ray, tmin, tmax, best, -1, key, 0u);
    DevHitOut o;
    memset(&o, 0, sizeof(o));
    o.prim = -1; o.mat = -1;
    if (best.op >= 0) {
        HitRec h;
        finalize_hit(S, ray, best, h);
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
    PerlinShared P{sh_vec, sh_perm};
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

typedef void (*render_fn)(const RenderParams);
static render_fn v3_kernel(bool counting, int min_blocks, bool fold) {
    if (counting) return render_kernel_v3<true, 1, false>;
    if (fold) {
        switch (min_blocks) {
            case 5: return render_kernel_v3<false, 5, true>;
            case 4: return render_kernel_v3<false, 4, true>;
            default: return render_kernel_v3<false, 6, true>;
        }
    }
    switch (min_blocks) {
        case 8: return render_kernel_v3<false, 8, false>;
        case 7: return render_kernel_v3<false, 7, false>;
        case 3: return render_kernel_v3<false, 3, false>;
        case 2: return render_kernel_v3<false, 2, false>;
        case 6: return render_kernel_v3<false, 6, false>;
        case 5: return render_kernel_v3<false, 5, false>;
        default: return render_kernel_v3<false, 4, false>;
    }
}
constexpr int kFoldBoxMin = 64;   // scenes with at least this many cube primitives run the FOLD form of the kernel (render_v3.cuh)

typedef void (*wf_extend_fn)(const WfParams);
typedef void (*wf_shade_fn)(const WfParams, const int);
static wf_extend_fn wf_extend(int min_blocks) {
    switch (min_blocks) {
        case 12: return wf_extend_kernel<12>;
        case 10: return wf_extend_kernel<10>;
        case 6: return wf_extend_kernel<6>;
        case 5: return wf_extend_kernel<5>;
        case 4: return wf_extend_kernel<4>;
        default: return wf_extend_kernel<8>;
    }
}
static wf_shade_fn wf_shade(int min_blocks) {
    switch (min_blocks) {
        case 8: return wf_shade_kernel<8>;
        case 6: return wf_shade_kernel<6>;
        case 3: return wf_shade_kernel<3>;
        default: return wf_shade_kernel<4>;
    }
}
constexpr int kWfBatch = 16;   // iterations enqueued between two looks at the live flag

struct rt_context {
    // wavefront renderer (render_v4.cuh): pool of in-flight paths and its bookkeeping
    WavePool pool{};
    size_t pool_capacity = 0;
    int pool_slots = 1 << 20;    // 5 x 16 B per slot = 80 MB: stays in the 126 MB L2 between the two kernels
    int wf_extend_blocks = 8, wf_shade_blocks = 4, wf_fetch_min = 6, wf_slab_fast = 14;
    unsigned long long* d_path_counter = nullptr;
    unsigned int* d_slot_cursor = nullptr;
    unsigned int* d_live = nullptr;          // 2 x kWfBatch flags
    unsigned int* h_live = nullptr;          // pinned, 2 flags
    ulonglong2* d_reserve = nullptr;
    size_t reserve_warps = 0;
    cudaEvent_t wf_event[2] = {nullptr, nullptr};
    bool timing = false;                     // RT_B200_TIMING: CUDA events around every launch (profiling runs only)
    double ms_extend = 0.0, ms_shade = 0.0;  // of the last render, timing mode only
    uint64_t wf_iterations = 0;
    int device = 0;
    int sm_count = 0;
    int clock_khz = 0;
    size_t total_mem = 0;
    int blocks_per_sm = 1;       // of the selected production kernel
    int variant = 3;             // 4: wavefront (render_v4.cuh); 3: megakernel (render_v3.cuh); 1: whole-segment loop (A/B only)
    int min_blocks = 6;          // occupancy variant: resident 128-thread blocks per SM the kernel is compiled for
    bool hoist_media = true;
    int shade_min = 24;
    int slab_fast = 14, slab_reps = 8, sphere_reps = 2;
    int slab_exit = 1, sphere_min = 33, box_min = 33, quad_min = 33;
    bool box_class = false;
    int fold_box = -1;           // -1: per scene (kFoldBoxMin); 0 / 1 forced (RT_B200_FOLD_BOX, A/B runs)
    bool prune_boxes = true;
    bool box_primitives = true;
    unsigned int* d_counter = nullptr;
    unsigned long long* d_stats = nullptr;
    float4* d_fb = nullptr;
    size_t fb_pixels = 0;
    rt_render_stats last{};
    uint64_t launches = 0;
};

struct rt_scene {
    rt_context* ctx = nullptr;
    int n_box = 0;               // OP_BOX primitives in the world program
    DevScene dev{};
    std::vector<void*> allocations;
    CompiledScene compiled;
};

extern "C" {

int rt_context_create(int device_id, rt_context** out) {
    if (!out) return fail(RT_ERR_INVALID_ARGUMENT, "rt_context_create: out is null");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) return fail(RT_ERR_NO_DEVICE, std::string("rt_context_create: no CUDA device (") + cudaGetErrorString(e) + "); there is no CPU fallback");
    if (device_id < 0 || device_id >= n) return fail(RT_ERR_OUT_OF_RANGE, "rt_context_create: device id out of range");
    CU(cudaSetDevice(device_id));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device_id));
    rt_context* c = new rt_context;
    c->device = device_id;
    c->sm_count = prop.multiProcessorCount;
    c->total_mem = prop.totalGlobalMem;
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, device_id);
    c->clock_khz = khz;
    // development switches (A/B runs; the defaults are the product)
    if (const char* e = std::getenv("RT_B200_KERNEL")) { const int v = std::atoi(e); c->variant = v == 1 ? 1 : v == 4 ? 4 : 3; }
    if (const char* e = std::getenv("RT_B200_POOL")) c->pool_slots = std::max(1024, std::atoi(e)) / 32 * 32;
    if (const char* e = std::getenv("RT_B200_WF_EXTEND_BLOCKS")) c->wf_extend_blocks = std::atoi(e);
    if (const char* e = std::getenv("RT_B200_WF_SHADE_BLOCKS")) c->wf_shade_blocks = std::atoi(e);
    if (const char* e = std::getenv("RT_B200_WF_FETCH_MIN")) c->wf_fetch_min = std::max(1, std::atoi(e));
    if (const char* e = std::getenv("RT_B200_WF_SLAB_FAST")) c->wf_slab_fast = std::max(1, std::atoi(e));
    if (const char* e = std::getenv("RT_B200_TIMING")) c->timing = std::atoi(e) != 0;
    if (const char* e = std::getenv("RT_B200_SHADE_MIN")) c->shade_min = std::max(1, std::atoi(e));
    if (const char* e = std::getenv("RT_B200_SLAB_FAST")) c->slab_fast = std::max(1, std::atoi(e));
    if (const char* e = std::getenv("RT_B200_SLAB_REPS")) c->slab_reps = std::max(1, std::atoi(e));
    if (const char* e = std::getenv("RT_B200_SPHERE_REPS")) c->sphere_reps = std::max(1, std::atoi(e));
    if (const char* e = std::getenv("RT_B200_NO_BOX")) c->box_primitives = std::atoi(e) == 0;
    if (const char* e = std::getenv("RT_B200_SLAB_EXIT")) c->slab_exit = std::max(1, std::atoi(e));
    if (const char* e = std::getenv("RT_B200_SPHERE_MIN")) c->sphere_min = std::max(1, std::atoi(e));
    if (const char* e = std::getenv("RT_B200_BOX_MIN")) c->box_min = std::max(1, std::atoi(e));
    if (const char* e = std::getenv("RT_B200_QUAD_MIN")) c->quad_min = std::max(1, std::atoi(e));
    if (const char* e = std::getenv("RT_B200_BOX_CLASS")) c->box_class = std::atoi(e) != 0;
    if (const char* e = std::getenv("RT_B200_NO_PRUNE")) c->prune_boxes = std::atoi(e) == 0;
    if (const char* e = std::getenv("RT_B200_FOLD_BOX")) c->fold_box = std::atoi(e) != 0 ? 1 : 0;
    if (const char* e = std::getenv("RT_B200_NO_HOIST")) c->hoist_media = std::atoi(e) == 0;
    if (const char* e = std::getenv("RT_B200_MIN_BLOCKS")) { const int v = std::atoi(e); c->min_blocks = v >= 8 ? 8 : v >= 2 ? v : 4; }
    CU(cudaFuncSetAttribute(render_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)perlin_smem_bytes()));
    for (int mb : {2, 3, 4, 5, 6, 7, 8})
        CU(cudaFuncSetAttribute(v3_kernel(false, mb, false), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)v3_smem_bytes(kMaxPerlinShared)));
    for (int mb : {4, 5, 6})
        CU(cudaFuncSetAttribute(v3_kernel(false, mb, true), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)v3_smem_bytes(kMaxPerlinShared)));
    CU(cudaFuncSetAttribute(v3_kernel(true, 1, false), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)v3_smem_bytes(kMaxPerlinShared)));
    CU(cudaMalloc(&c->d_path_counter, sizeof(unsigned long long)));
    CU(cudaMalloc(&c->d_slot_cursor, sizeof(unsigned int)));
    CU(cudaMalloc(&c->d_live, 2 * kWfBatch * sizeof(unsigned int)));
    CU(cudaMallocHost(&c->h_live, 2 * sizeof(unsigned int)));
    CU(cudaEventCreateWithFlags(&c->wf_event[0], cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&c->wf_event[1], cudaEventDisableTiming));
    int bps = 0;
    if (c->variant == 1) CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, render_kernel, kBlockThreads, perlin_smem_bytes()));
    else CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, v3_kernel(false, c->min_blocks, false), kBlockThreads, v3_smem_bytes(1)));
    c->blocks_per_sm = bps > 0 ? bps : 1;
    CU(cudaMalloc(&c->d_counter, sizeof(unsigned int)));
    CU(cudaMalloc(&c->d_stats, K_NUM * sizeof(unsigned long long)));
    CU(cudaMemset(c->d_stats, 0, K_NUM * sizeof(unsigned long long)));
    *out = c;
    return RT_OK;
}

void rt_context_destroy(rt_context* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaFree(c->d_counter);
    cudaFree(c->d_stats);
    cudaFree(c->d_fb);
    cudaFree(c->pool.ray0); cudaFree(c->pool.ray1); cudaFree(c->pool.hit); cudaFree(c->pool.st0); cudaFree(c->pool.st1);
    cudaFree(c->d_path_counter); cudaFree(c->d_slot_cursor); cudaFree(c->d_live); cudaFree(c->d_reserve);
    if (c->h_live) cudaFreeHost(c->h_live);
    for (cudaEvent_t e : c->wf_event) if (e) cudaEventDestroy(e);
    delete c;
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
    const size_t alloc = bytes ? bytes : 16;
    CU(cudaMalloc(dst, alloc));
    s->allocations.push_back(*dst);
    if (bytes) CU(cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice));
    return RT_OK;
}

// host dry runs compile with the product's defaults; RT_B200_NO_PRUNE=1 (development) shows the stream before box pruning
static CompileOptions dry_run_options() {
    CompileOptions o;
    if (const char* e = std::getenv("RT_B200_NO_PRUNE")) o.prune_boxes = std::atoi(e) == 0;
    if (const char* e = std::getenv("RT_B200_COST_SPHERE")) o.cost_sphere = std::atof(e);
    if (const char* e = std::getenv("RT_B200_COST_QUAD")) o.cost_quad = std::atof(e);
    if (const char* e = std::getenv("RT_B200_COST_BOX")) o.cost_box = std::atof(e);
    return o;
}

int rt_scene_upload(rt_context* c, const rt_scene_desc* desc, rt_scene** out) {
    if (!c || !desc || !out) return fail(RT_ERR_INVALID_ARGUMENT, "rt_scene_upload: null argument");
    CU(cudaSetDevice(c->device));
    rt_scene* s = new rt_scene;
    s->ctx = c;
    const char* err = nullptr;
    CompileOptions copt = dry_run_options();
    copt.box_primitives = c->box_primitives;
    copt.hoist_media = c->hoist_media;
    copt.box_class = c->box_class && c->variant == 3;
    copt.prune_boxes = c->prune_boxes && copt.prune_boxes;
    int rc = compile_scene(desc, copt, &s->compiled, &err);
    if (rc < 0) { delete s; return fail(rc, err ? err : "compile_scene failed"); }
    if (s->compiled.n_perlin > kMaxPerlinShared) { delete s; return fail(RT_ERR_UNSUPPORTED, "more than 4 NoiseTexture tables in one scene"); }
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
    s->dev.n_words = cs.n_world_words;
    for (int i = 0; i < cs.n_world_words; ++i) {   // words of other ops can alias a header only by accident of their bits:
        uint32_t hdr;                               // walk op by op
        std::memcpy(&hdr, &cs.ops[i].w, 4);
        const uint32_t kind = hdr & 15u, flags = (hdr >> 4) & 15u;
        if (kind == OP_BOX) s->n_box++;
        i += (kind == OP_QUAD || kind == OP_XFORM_ENTER) ? 3 : kind == OP_BOX ? 2 : kind == OP_SPHERE ? ((flags & FLAG_MOVING) ? 2 : 1)
             : kind == OP_MEDIUM ? ((int)flags == MEDIUM_BOUNDARY_XBOX ? 4 : 2) : 1;
    }
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
            cudaError_t e = cudaMalloc(&d_rgb, (size_t)n * 3);
            if (e != cudaSuccess) { rt_scene_destroy(s); return cuda_fail(e, "cudaMalloc(image rgb8)"); }
            e = cudaMalloc(&d_tex, (size_t)n * sizeof(float4));
            if (e != cudaSuccess) { cudaFree(d_rgb); rt_scene_destroy(s); return cuda_fail(e, "cudaMalloc(image texels)"); }
            s->allocations.push_back(d_tex);
            e = cudaMemcpy(d_rgb, im.rgb8, (size_t)n * 3, cudaMemcpyHostToDevice);
            if (e == cudaSuccess) {
                expand_image_kernel<<<(unsigned)((n + 255) / 256), 256>>>(d_rgb, n, d_lut, d_tex);
                e = cudaDeviceSynchronize();
            }
            cudaFree(d_rgb);
            if (e != cudaSuccess) { rt_scene_destroy(s); return cuda_fail(e, "image upload"); }
            imgs[k].texels = d_tex;
            imgs[k].width = im.width;
            imgs[k].height = im.height;
        }
    }
    rc = upload_vec(s, imgs.data(), imgs.size() * sizeof(DevImage), &p);
    if (rc < 0) { rt_scene_destroy(s); return rc; }
    s->dev.images = static_cast<const DevImage*>(p);
    *out = s;
    return RT_OK;
}

int rt_scene_layout(const rt_scene_desc* desc, rt_layout_info* out) {
    if (!desc || !out) return fail(RT_ERR_INVALID_ARGUMENT, "rt_scene_layout: null argument");
    CompiledScene cs;
    const char* err = nullptr;
    const int rc = compile_scene(desc, dry_run_options(), &cs, &err);
    if (rc < 0) return fail(rc, err ? err : "compile_scene failed");
    std::memset(out, 0, sizeof(*out));
    out->n_words = cs.n_world_words;
    const int n_all = (int)cs.ops.size() - 2;   // without the two padding words
    for (int i = 0; i < cs.n_world_words;) {
        uint32_t hdr;
        std::memcpy(&hdr, &cs.ops[i].w, 4);
        const uint32_t kind = hdr & 15u, flags = (hdr >> 4) & 15u;
        switch (kind) {
            case OP_INNER: case OP_INNER_REF: out->n_inner++; i += 2; break;
            case OP_SPHERE: out->n_sphere++; i += (flags & FLAG_MOVING) ? 3 : 2; break;
            case OP_QUAD: out->n_quad++; i += 4; break;
            case OP_XFORM_ENTER: out->n_xform++; i += 4; break;
            case OP_XFORM_EXIT: i += 2; break;
            case OP_MEDIUM: out->n_medium_in_stream++; i += (int)flags == MEDIUM_BOUNDARY_XBOX ? 5 : 3; break;
            case OP_BOX: out->n_box++; i += 3; break;
            default: return fail(RT_ERR_INTERNAL, "rt_scene_layout: bad op in stream");
        }
    }
    out->n_medium_hoisted = (int32_t)cs.hoisted_media.size();
    out->n_precise_spheres = (int32_t)cs.precise.size() / 2;
    out->n_bvh = (int32_t)cs.bvh_hittable_ids.size();
    int64_t bytes = (int64_t)(n_all + 2) * 16 + (int64_t)(cs.materials.size() + cs.textures.size() + cs.perlin_vec.size()) * 16 +
                    (int64_t)cs.perlin_perm.size() + (int64_t)cs.precise.size() * 32;
    for (int k = 0; k < desc->n_images; ++k) bytes += (int64_t)desc->images[k].width * desc->images[k].height * 16;
    out->device_bytes = bytes;
    return RT_OK;
}

int rt_scene_ops_export(const rt_scene_desc* desc, float* words, int64_t capacity_words, int64_t* n_total_words,
                        int32_t* n_world_words, int32_t* media_ops, int32_t* n_media, int32_t* first_class) {
    if (!desc || !n_total_words) return fail(RT_ERR_INVALID_ARGUMENT, "rt_scene_ops_export: null argument");
    CompiledScene cs;
    const char* err = nullptr;
    const int rc = compile_scene(desc, dry_run_options(), &cs, &err);
    if (rc < 0) return fail(rc, err ? err : "compile_scene failed");
    *n_total_words = (int64_t)cs.ops.size();
    if (n_world_words) *n_world_words = cs.n_world_words;
    if (n_media) *n_media = (int32_t)cs.hoisted_media.size();
    if (media_ops) for (size_t k = 0; k < cs.hoisted_media.size() && k < (size_t)kMaxHoistedMedia; ++k) media_ops[k] = cs.hoisted_media[k];
    if (first_class) *first_class = (int32_t)cs.first_class;
    if (words) {
        if (capacity_words < (int64_t)cs.ops.size()) return fail(RT_ERR_OUT_OF_RANGE, "rt_scene_ops_export: capacity too small");
        std::memcpy(words, cs.ops.data(), cs.ops.size() * sizeof(F4));
    }
    return RT_OK;
}

void rt_scene_destroy(rt_scene* s) {
    if (!s) return;
    if (s->ctx) cudaSetDevice(s->ctx->device);
    for (void* p : s->allocations) cudaFree(p);
    delete s;
}

// Wavefront render (render_v4.cuh): alternate wf_shade_kernel / wf_extend_kernel over the pool until no slot carries
// a ray. Iterations are enqueued in batches of kWfBatch; the live flag of batch b is looked at after batch b + 1 has
// been enqueued, so the stream never runs dry while the host decides. Returns with the stream drained.
static size_t wf_shade_smem(int n_perlin) {
    const int np = n_perlin < kMaxPerlinShared ? n_perlin : kMaxPerlinShared;
    return (size_t)np * (256 * sizeof(float4) + 768);
}

static int wf_ensure_pool(rt_context* c, size_t slots) {
    if (c->pool_capacity >= slots) return RT_OK;
    cudaFree(c->pool.ray0); cudaFree(c->pool.ray1); cudaFree(c->pool.hit); cudaFree(c->pool.st0); cudaFree(c->pool.st1);
    c->pool = WavePool{};
    c->pool_capacity = 0;
    CU(cudaMalloc(&c->pool.ray0, slots * sizeof(float4)));
    CU(cudaMalloc(&c->pool.ray1, slots * sizeof(float4)));
    CU(cudaMalloc(&c->pool.hit, slots * sizeof(float4)));
    CU(cudaMalloc(&c->pool.st0, slots * sizeof(float4)));
    CU(cudaMalloc(&c->pool.st1, slots * sizeof(float4)));
    c->pool_capacity = slots;
    return RT_OK;
}

static int launch_render_v4(rt_context* c, const rt_scene* s, const rt_camera_desc* cam, int64_t sample_begin,
                            int64_t sample_count, uint64_t seed, void* d_sum_rgba, cudaStream_t stream) {
    WfParams prm;
    prm.scene = s->dev;
    prm.cam = make_dev_camera(*cam);
    prm.seed = seed;
    prm.sample_begin = sample_begin;
    prm.n_pixels = (uint32_t)((int64_t)prm.cam.width * prm.cam.height);
    prm.n_paths = (unsigned long long)prm.n_pixels * (unsigned long long)sample_count;
    prm.tiled = (prm.cam.width % kTileW == 0 && prm.cam.height % kTileH == 0) ? 1 : 0;
    prm.tiles_x = prm.cam.width / kTileW;
    size_t slots = (size_t)c->pool_slots;
    const unsigned long long rounded = (prm.n_paths + 31ull) / 32ull * 32ull;
    if (rounded < (unsigned long long)slots) slots = (size_t)rounded;
    int rc = wf_ensure_pool(c, slots);
    if (rc < 0) return rc;
    prm.pool = c->pool;
    prm.pool.n_slots = (int)slots;
    prm.sum = static_cast<float4*>(d_sum_rgba);
    prm.path_counter = c->d_path_counter;
    prm.slot_cursor = c->d_slot_cursor;
    prm.stats = c->d_stats;
    prm.first_class = (int)s->compiled.first_class;
    prm.fetch_min = c->wf_fetch_min;
    prm.slab_fast = c->wf_slab_fast;
    prm.sphere_reps = c->sphere_reps;
    prm.slab_exit = c->slab_exit;
    prm.sphere_min = c->sphere_min;

    wf_extend_fn extend = wf_extend(c->wf_extend_blocks);
    wf_shade_fn shade_k = wf_shade(c->wf_shade_blocks);
    const size_t smem = wf_shade_smem(s->dev.n_perlin);
    int bps_e = 1, bps_s = 1;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps_e, extend, kBlockThreads, 0));
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps_s, shade_k, kWfShadeThreads, smem));
    const size_t slot_warps = slots / 32;
    size_t grid_e = (size_t)c->sm_count * (size_t)std::max(1, bps_e);
    size_t grid_s = (size_t)c->sm_count * (size_t)std::max(1, bps_s);
    grid_e = std::max<size_t>(1, std::min(grid_e, (slot_warps + kBlockThreads / 32 - 1) / (kBlockThreads / 32)));
    grid_s = std::max<size_t>(1, std::min(grid_s, (slot_warps + kWfShadeThreads / 32 - 1) / (kWfShadeThreads / 32)));
    const size_t shade_warps = grid_s * (kWfShadeThreads / 32);
    if (c->reserve_warps < shade_warps) {
        cudaFree(c->d_reserve);
        c->d_reserve = nullptr;
        c->reserve_warps = 0;
        CU(cudaMalloc(&c->d_reserve, shade_warps * sizeof(ulonglong2)));
        c->reserve_warps = shade_warps;
    }
    prm.reserve = c->d_reserve;

    CU(cudaMemsetAsync(c->d_path_counter, 0, sizeof(unsigned long long), stream));
    CU(cudaMemsetAsync(c->d_slot_cursor, 0, sizeof(unsigned int), stream));
    CU(cudaMemsetAsync(c->d_reserve, 0, shade_warps * sizeof(ulonglong2), stream));
    CU(cudaMemsetAsync(c->d_stats, 0, K_NUM * sizeof(unsigned long long), stream));
    wf_init_kernel<<<(unsigned)((slots + 255) / 256), 256, 0, stream>>>(prm.pool.hit, (int)slots);
    c->launches += 1;
    c->ms_extend = c->ms_shade = 0.0;
    c->wf_iterations = 0;

    struct EventBag {               // timing mode: begin / middle / end of every iteration; destroyed on every exit path
        std::vector<cudaEvent_t> v;
        ~EventBag() { for (cudaEvent_t e : v) cudaEventDestroy(e); }
    } bag;
    std::vector<cudaEvent_t>& tev = bag.v;
    for (int b = 0;; ++b) {
        unsigned int* live = c->d_live + (b & 1) * kWfBatch;
        prm.live_flag = live;
        CU(cudaMemsetAsync(live, 0, kWfBatch * sizeof(unsigned int), stream));
        for (int k = 0; k < kWfBatch; ++k) {
            cudaEvent_t e0 = nullptr, e1 = nullptr, e2 = nullptr;
            if (c->timing) {
                cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2);
                tev.push_back(e0); tev.push_back(e1); tev.push_back(e2);
                cudaEventRecord(e0, stream);
            }
            shade_k<<<(unsigned)grid_s, kWfShadeThreads, smem, stream>>>(prm, k);
            if (c->timing) cudaEventRecord(e1, stream);
            extend<<<(unsigned)grid_e, kBlockThreads, 0, stream>>>(prm);
            if (c->timing) cudaEventRecord(e2, stream);
        }
        CU(cudaGetLastError());
        c->launches += 2 * kWfBatch;
        c->wf_iterations += kWfBatch;
        CU(cudaMemcpyAsync(c->h_live + (b & 1), live + (kWfBatch - 1), sizeof(unsigned int), cudaMemcpyDeviceToHost, stream));
        CU(cudaEventRecord(c->wf_event[b & 1], stream));
        if (b >= 1) {
            CU(cudaEventSynchronize(c->wf_event[(b - 1) & 1]));
            if (c->h_live[(b - 1) & 1] == 0u) break;
        }
    }
    CU(cudaStreamSynchronize(stream));
    if (c->timing) {
        for (size_t k = 0; k + 2 < tev.size(); k += 3) {
            float a = 0.0f, b2 = 0.0f;
            cudaEventElapsedTime(&a, tev[k], tev[k + 1]);
            cudaEventElapsedTime(&b2, tev[k + 1], tev[k + 2]);
            c->ms_shade += a;
            c->ms_extend += b2;
        }
    }
    return RT_OK;
}

static int launch_render(rt_context* c, const rt_scene* s, const rt_camera_desc* cam, int64_t sample_begin,
                         int64_t sample_count, uint64_t seed, void* d_sum_rgba, cudaStream_t stream, bool counting) {
    if (c->variant == 4 && !counting) return launch_render_v4(c, s, cam, sample_begin, sample_count, seed, d_sum_rgba, stream);
    RenderParams prm;
    prm.scene = s->dev;
    prm.cam = make_dev_camera(*cam);
    prm.seed = seed;
    prm.sample_begin = sample_begin;
    prm.sample_count = (int)sample_count;
    prm.tiles_x = (prm.cam.width + kTileW - 1) / kTileW;
    prm.tiles_y = (prm.cam.height + kTileH - 1) / kTileH;
    // pool = tile x chunk samples; keep >= 64 pools per resident warp so the tail of the launch (warps running dry
    // while others still hold a pool) stays near 1%: short renders get small chunks, the 10000-spp bench keeps 32
    const int64_t n_tiles = (int64_t)prm.tiles_x * prm.tiles_y;
    const int64_t resident_warps = (int64_t)c->sm_count * c->blocks_per_sm * (kBlockThreads / 32);
    int chunk = 32;
    if (const char* e = std::getenv("RT_B200_CHUNK")) chunk = std::max(1, std::atoi(e));
    while (chunk > 1 && n_tiles * ((sample_count + chunk - 1) / chunk) < resident_warps * 64) chunk >>= 1;
    prm.chunk = chunk;
    prm.n_chunks = (int)((sample_count + chunk - 1) / chunk);
    if ((uint64_t)n_tiles * (uint64_t)prm.n_chunks >= 0xffffffffull) return fail(RT_ERR_INVALID_ARGUMENT, "rt_render_accumulate: too many work items; split the sample range");
    prm.sum = static_cast<float4*>(d_sum_rgba);
    prm.work_counter = c->d_counter;
    prm.stats = c->d_stats;
    prm.first_class = (int)s->compiled.first_class;
    prm.shade_min = c->shade_min;
    prm.slab_fast = c->slab_fast;
    prm.slab_reps = c->slab_reps;
    prm.sphere_reps = c->sphere_reps;
    prm.slab_exit = c->slab_exit;
    prm.sphere_min = c->sphere_min;
    prm.box_min = c->box_min;
    prm.quad_min = c->quad_min;
    CU(cudaMemsetAsync(c->d_counter, 0, sizeof(unsigned int), stream));
    CU(cudaMemsetAsync(c->d_stats, 0, K_NUM * sizeof(unsigned long long), stream));
    int grid = c->sm_count * c->blocks_per_sm;
    if (counting || c->variant == 3) {
        const size_t smem = v3_smem_bytes(s->dev.n_perlin);
        const bool fold = c->fold_box < 0 ? s->n_box >= kFoldBoxMin : c->fold_box != 0;
        render_fn fn = v3_kernel(counting, c->min_blocks, fold);
        int bps = 1;
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, fn, kBlockThreads, smem));
        grid = c->sm_count * (bps > 0 ? bps : 1);
        fn<<<grid, kBlockThreads, smem, stream>>>(prm);
    } else {
        render_kernel<<<grid, kBlockThreads, perlin_smem_bytes(), stream>>>(prm);
    }
    CU(cudaGetLastError());
    c->launches += 1;
    return RT_OK;
}

static int check_render_args(rt_context* c, const rt_scene* s, const rt_camera_desc* cam, int64_t sample_count, const void* buf,
                             const char* who) {
    if (!c || !s || !cam || !buf) return fail(RT_ERR_INVALID_ARGUMENT, std::string(who) + ": null argument");
    if (sample_count < 0 || sample_count > 0x7fffffff) return fail(RT_ERR_INVALID_ARGUMENT, std::string(who) + ": bad sample_count");
    if (cam->image_width <= 0 || cam->image_height <= 0 || cam->image_width * cam->image_height > 0x7fffffff)
        return fail(RT_ERR_INVALID_ARGUMENT, std::string(who) + ": bad image size");
    if (cam->max_depth <= 0)   // ray_color returns black at depth <= 0 (renderer.rs:140-142); the CLI scenes never ask for it
        return fail(RT_ERR_INVALID_ARGUMENT, std::string(who) + ": max_depth must be positive");
    return RT_OK;
}

int rt_render_accumulate(rt_context* c, const rt_scene* s, const rt_camera_desc* cam, int64_t sample_begin,
                         int64_t sample_count, uint64_t seed, void* d_sum_rgba, void* stream_) {
    int rc = check_render_args(c, s, cam, sample_count, d_sum_rgba, "rt_render_accumulate");
    if (rc < 0) return rc;
    if (sample_count == 0) return RT_OK;
    CU(cudaSetDevice(c->device));
    return launch_render(c, s, cam, sample_begin, sample_count, seed, d_sum_rgba, static_cast<cudaStream_t>(stream_), false);
}

int rt_render_count_ops(rt_context* c, const rt_scene* s, const rt_camera_desc* cam, int64_t sample_begin, int64_t sample_count,
                        uint64_t seed, uint64_t* counters, int capacity, const char** names_csv) {
    int rc = check_render_args(c, s, cam, sample_count, counters, "rt_render_count_ops");
    if (rc < 0) return rc;
    if (capacity < (int)K_NUM) return fail(RT_ERR_OUT_OF_RANGE, "rt_render_count_ops: capacity too small");
    if (names_csv) *names_csv = kCounterNames;
    CU(cudaSetDevice(c->device));
    const size_t n = (size_t)cam->image_width * (size_t)cam->image_height;
    float4* scratch = nullptr;
    CU(cudaMalloc(&scratch, n * sizeof(float4)));
    cudaMemset(scratch, 0, n * sizeof(float4));
    rc = sample_count > 0 ? launch_render(c, s, cam, sample_begin, sample_count, seed, scratch, nullptr, true) : RT_OK;
    unsigned long long h[K_NUM] = {0};
    cudaError_t e = cudaMemcpy(h, c->d_stats, sizeof(h), cudaMemcpyDeviceToHost);
    cudaFree(scratch);
    if (rc < 0) return rc;
    if (e != cudaSuccess) return cuda_fail(e, "rt_render_count_ops");
    for (int k = 0; k < (int)K_NUM; ++k) counters[k] = h[k];
    return (int)K_NUM;
}

int rt_render_get_stats(rt_context* c, rt_render_stats* out) {
    if (!c || !out) return fail(RT_ERR_INVALID_ARGUMENT, "rt_render_get_stats: null argument");
    CU(cudaSetDevice(c->device));
    unsigned long long h[2] = {0, 0};
    CU(cudaMemcpy(h, c->d_stats, sizeof(h), cudaMemcpyDeviceToHost));
    out->paths = h[0];
    out->segments = h[1];
    out->kernel_launches = c->launches;
    out->last_kernel_ms = 0.0f;
    return RT_OK;
}

int rt_render_get_kernel_times(rt_context* c, double* ms_shade, double* ms_extend, uint64_t* iterations) {
    if (!c) return fail(RT_ERR_INVALID_ARGUMENT, "rt_render_get_kernel_times: context is null");
    if (ms_shade) *ms_shade = c->ms_shade;
    if (ms_extend) *ms_extend = c->ms_extend;
    if (iterations) *iterations = c->wf_iterations;
    return RT_OK;
}

int rt_render(rt_context* c, const rt_scene* s, const rt_camera_desc* cam, int64_t sample_begin, int64_t sample_count,
              uint64_t seed, float* host_sum_rgba) {
    if (!c || !s || !cam || !host_sum_rgba) return fail(RT_ERR_INVALID_ARGUMENT, "rt_render: null argument");
    if (cam->image_width <= 0 || cam->image_height <= 0) return fail(RT_ERR_INVALID_ARGUMENT, "rt_render: bad image size");
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
    int rc = rt_render_accumulate(c, s, cam, sample_begin, sample_count, seed, c->d_fb, nullptr);
    if (rc < 0) return rc;
    CU(cudaMemcpy(host_sum_rgba, c->d_fb, n * sizeof(float4), cudaMemcpyDeviceToHost));
    return RT_OK;
}

int rt_finalize_rgb8(rt_context* c, const void* d_sum_rgba, int64_t n_pixels, double spp, uint8_t* host_rgb8) {
    if (!c || !d_sum_rgba || !host_rgb8) return fail(RT_ERR_INVALID_ARGUMENT, "rt_finalize_rgb8: null argument");
    if (n_pixels <= 0) return RT_OK;
    CU(cudaSetDevice(c->device));
    uint8_t* d_rgb = nullptr;
    CU(cudaMalloc(&d_rgb, (size_t)n_pixels * 3));
    finalize_kernel<<<(unsigned)((n_pixels + 255) / 256), 256>>>(static_cast<const float4*>(d_sum_rgba), n_pixels,
                                                               spp > 0 ? (float)(1.0 / spp) : 0.0f, spp > 0 ? 0 : 1, d_rgb);
    cudaError_t e = cudaMemcpy(host_rgb8, d_rgb, (size_t)n_pixels * 3, cudaMemcpyDeviceToHost);
    cudaFree(d_rgb);
    if (e != cudaSuccess) return cuda_fail(e, "rt_finalize_rgb8");
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
