// K1: the render loop (renderer.rs:26-49,139-155) as ONE persistent CTA per SM whose lanes walk the op stream out of
// shared memory, regrouping by op class through a warp vote.
//
// Every lane is always somewhere in {a traversal op of some class, waiting to shade / get a new path}. Each
// iteration the warp counts its lanes per class, runs only the most populated class (lanes of other classes wait,
// which costs no issue slots), and lanes whose op finished move on to their next op's class - known from the link
// they follow, before the op's words arrive (dev_scene.h). Lanes regroup by what they are about to execute instead
// of idling behind the longest traversal or the rarest op kind of the warp.
//
// What the round-2 profile changed (profiles/r2_k1_region_breakdown.md: the round-1 kernel spent 48% of its issue
// slots in the box-test loop at 65 instructions per repetition, 35 of them on the half-rate ALU pipe, which at 55%
// busy was the kernel's busiest unit; 13% of the stream fetches missed L1):
//  * the op stream lives in SHARED memory, staged once per CTA by a bulk (TMA) copy: a fetch is one LDS.128 with an
//    immediate offset, no address arithmetic, no L1 tag lookups, no L2 round trips. One CTA of 24 warps per SM shares
//    the copy. Streams that do not fit fall back to the same kernel reading global memory (template parameter);
//  * cull boxes are centre / half extent and the per-ray constants are 1/d and -o/d: the slab test is twelve
//    FMA-pipe instructions and two 3-input min/max, with no per-node sign selects;
//  * the cursor is a link (byte offset | class << 28) and the header of a cull box IS its fall-through increment: the
//    successor of a box test is one add and one select, "am I at a box" one unsigned compare;
//  * the repetition loop ends with one ballot + popc + compare (no repetition counter), and its body has no divergent
//    branch in the common case: EVERY lane runs the box arithmetic on the op words it holds (lanes of other classes keep
//    their link, the fetch is a predicated LDS), so a repetition is one straight run of ~37 instructions instead of a
//    branch around the body with its reconvergence pair (+5%);
//  * FOLD (scenes with cube primitives / instances): such an op whose box the ray MISSES is finished by the same arithmetic
//    (its second word ends with the link to follow), only the accepting side enters the rare-kind branch: 21% of the
//    repetitions instead of 70% (+6% final_scene; three more instructions per repetition, so it is a template parameter);
//  * the loop's top is the full vote; after another class has run the box-test loop is entered directly when enough lanes
//    stand at boxes, and it hands its last lane count to the next vote (no second ballot, no jump table: +4..7%);
//  * the kernel is specialised on what the scene holds (FEAT_*): f64 spheres, in-stream media / boundary programs and
//    reference boxes are compiled out of the instantiation that renders a scene without them (5 424 -> 2 928 instructions
//    for final_scene, +6%; the kernel is bound by instruction fetch, profiles/r2_k1_icache.md);
//  * shading, path regeneration and segment start are ONE out-of-line function (shade_phase): its register needs no
//    longer decide what the traversal loop may keep in registers (the inlined form spilled the loop's own per-ray
//    constants and reloaded them in every repetition). Everything the two sides share lives in shared memory;
//  * the code this kernel EXECUTES on final_scene (2 100 of its 6 000 instructions) sat at the edge of the 32 KB instruction
//    cache: divisions, reciprocals and square roots are single special-function instructions now (rt_kernels.cuh frcp /
//    fdiv / fsqrt) instead of the compiler's range-fix-up sequences - 9% fewer instructions, +23% Mpaths/s
//    (profiles/r2_k1_icache.md). Keep that in mind before inlining anything else into the shade phase.
//
// Included by rt_cuda.cu.
#pragma once

#ifdef RT_OPT_THREADS
constexpr int kRenderThreads = RT_OPT_THREADS;
#else
constexpr int kRenderThreads = 896;   // 28 warps: one CTA per SM (72 registers per thread; 768 / 80: -4..6%, profiles/r2_ab15*)
#endif
constexpr int kRenderWarps = kRenderThreads / 32;

// per-thread path state in shared memory, one SoA column per thread: field f of thread t at cold[f * kRenderThreads + t]
enum ColdField {
    F_WO = 0,      // world ray origin (3)
    F_WD = 3,      // world ray direction (3)
    F_L = 6,       // radiance gathered so far (3)
    F_TP = 9,      // throughput (3)
    F_SAMPLE = 12, F_TIME = 13, F_PIX = 14, F_DEPTH = 15,
    F_O = 16,      // origin of the CURRENT ray: the world ray, or the local ray inside an instance (3)
    F_D = 19,      // its direction (3)
    F_A = 22,      // |d|^2 of the current ray and its reciprocal (sphere tests)
    F_INVA = 23,
    F_XF = 24,     // word index of the active OP_XFORM_ENTER, -1 = world space
    F_ORIGIN = 25, // origin code of the current ray (rt_kernels.cuh)
    F_BEST_T = 26, F_BEST_OP = 27, F_BEST_XF = 28,   // the closest hit so far, parked here across the out-of-line phases
    kColdFields = 29
};
// per-warp pool of (tile x sample chunk) paths, and the warp's statistics
enum PoolField { W_NEXT = 0, W_SIZE, W_X0, W_Y0, W_TILE_W, W_TILE_N, W_SAMPLE0, W_NO_MORE, W_PATHS, W_SEGS, kPoolFields };

// shared memory: [0, 8) the mbarrier of the bulk copy, [kSmemParams, kSmemOps) a copy of RenderParams, then the op stream
static_assert(sizeof(RenderParams) <= 512, "RenderParams outgrew its shared-memory slot");
static_assert(sizeof(RenderParams) % 4 == 0, "RenderParams is copied word by word");

struct MkSmem { size_t vec_off, perm_off, cold_off, pool_off, total; };
inline MkSmem mk_smem_layout(size_t ops_bytes, int n_perlin) {   // ops_bytes = 0: the stream stays in global memory
    const int np = n_perlin < kMaxPerlinShared ? n_perlin : kMaxPerlinShared;
    MkSmem L;
    size_t off = kSmemOps + ops_bytes;     // ops_bytes is a multiple of 16
    L.vec_off = off;  off += (size_t)np * 256 * sizeof(float4);
    L.perm_off = off; off += (size_t)np * 768;
    L.cold_off = off; off += (size_t)kColdFields * kRenderThreads * sizeof(float);
    L.pool_off = off; off += (size_t)kPoolFields * kRenderWarps * sizeof(int);
    L.total = off;
    return L;
}

#define COLD(f) cold[(f) * kRenderThreads + tid]
#define COLD_U(f) reinterpret_cast<uint32_t*>(cold)[(f) * kRenderThreads + tid]
#define COLD_I(f) reinterpret_cast<int*>(cold)[(f) * kRenderThreads + tid]

// classes a lane without a traversal op can be in (besides CLS_SHADE = holds a finished segment)
constexpr uint32_t CLS_NEED = 6;   // holds no path and wants one; CLS_IDLE: holds none and there is none left

// Shade the finished segments (renderer.rs:144-153), hand out new paths, start the next segments: hoisted media,
// per-ray set-up into shared memory. Called by all 32 lanes of the warp; lanes that are mid-traversal pass through.
// Returns the lane's new link.
template <bool COUNT, class Ops, unsigned FEAT>
__device__ __noinline__ uint32_t shade_phase(uint32_t link, unsigned* cnt) {
    unsigned char* smem = reinterpret_cast<unsigned char*>(dyn_smem);
    const RenderParams& prm = *reinterpret_cast<const RenderParams*>(smem + kSmemParams);
    const DevScene& S = prm.scene;
    const DevCamera& C = prm.cam;
    const int np = min(S.n_perlin, kMaxPerlinShared);
    float4* sh_vec = reinterpret_cast<float4*>(smem + kSmemOps + prm.ops_bytes);
    uint8_t* sh_perm = reinterpret_cast<uint8_t*>(sh_vec + np * 256);
    float* cold = reinterpret_cast<float*>(sh_perm + np * 768);
    int* pool = reinterpret_cast<int*>(cold + kColdFields * kRenderThreads) + (threadIdx.x >> 5) * kPoolFields;
    const PerlinShared P{sh_vec, sh_perm, min(S.n_perlin, kMaxPerlinShared)};
    Ops ops;
    set_ops_base(ops, S.ops);
    const int tid = threadIdx.x;
    const unsigned lane = threadIdx.x & 31u;
    const float tmin = 0.001f;                     // renderer.rs:144
    const float inf = __int_as_float(0x7f800000);

    bool start = false;                 // this lane begins a new segment below (one shared copy of that code)
    bool has_path = (link >> 28) < CLS_IDLE;
    uint4 key = make_uint4(0, 0, 0, 0);
    uint32_t depth = 0;
    if ((link >> 28) == CLS_SHADE) {
        Ray ray;
        ray.o = f3(COLD(F_WO), COLD(F_WO + 1), COLD(F_WO + 2));
        ray.d = f3(COLD(F_WD), COLD(F_WD + 1), COLD(F_WD + 2));
        ray.time = COLD(F_TIME);
        float3 L = f3(COLD(F_L), COLD(F_L + 1), COLD(F_L + 2)), Tp = f3(COLD(F_TP), COLD(F_TP + 1), COLD(F_TP + 2));
        key = path_key(prm.seed, COLD_U(F_PIX), COLD_U(F_SAMPLE));
        depth = COLD_U(F_DEPTH);
        bool alive;
        const int best_op = COLD_I(F_BEST_OP), best_xf = COLD_I(F_BEST_XF);
        if (best_op < 0) {
            L = L + Tp * C.background;                                  // renderer.rs:152-153
            alive = false;
        } else {
            HitRec h;
            Best b;
            b.t = COLD(F_BEST_T); b.op = best_op; b.xf = best_xf;
            finalize_hit<FEAT>(S, ops, ray, b, h);
            if (COUNT) {
                if (best_xf >= 0) cnt[K_FINALIZE_XFORM]++;
                const float4 m0 = __ldg(S.mats + 2 * h.mat);
                const int mk = fbits(m0.x);
                cnt[mk == RT_MAT_LAMBERTIAN ? K_LAMBERTIAN : mk == RT_MAT_METAL ? K_METAL : mk == RT_MAT_DIELECTRIC ? K_DIELECTRIC
                    : mk == RT_MAT_ISOTROPIC ? K_ISOTROPIC : K_LIGHT]++;
                if (mk == RT_MAT_LAMBERTIAN || mk == RT_MAT_ISOTROPIC || mk == RT_MAT_DIFFUSE_LIGHT) {
                    int tx = fbits(m0.y);
                    for (int g = 0; g < 16; ++g) {
                        const float4 t0 = __ldg(S.texs + 2 * tx);
                        const int tk = fbits(t0.x);
                        if (tk == RT_TEX_CHECKER) {
                            cnt[K_TEX_CHECKER]++;
                            const int x = (int)floorf(t0.w * h.p.x), y = (int)floorf(t0.w * h.p.y), z = (int)floorf(t0.w * h.p.z);
                            tx = ((x + y + z) % 2 == 0) ? fbits(t0.y) : fbits(t0.z);
                            continue;
                        }
                        if (tk == RT_TEX_NOISE) cnt[K_TEX_NOISE]++;
                        if (tk == RT_TEX_IMAGE) cnt[K_TEX_IMAGE]++;
                        break;
                    }
                }
            }
            alive = shade(S, P, ray, h, key, depth, L, Tp);
            COLD_I(F_ORIGIN) = h.origin;
            ++depth;
            if ((int)depth >= C.max_depth) alive = false;               // renderer.rs:140-142
        }
        if (alive) {
            COLD(F_WO) = ray.o.x; COLD(F_WO + 1) = ray.o.y; COLD(F_WO + 2) = ray.o.z;
            COLD(F_WD) = ray.d.x; COLD(F_WD + 1) = ray.d.y; COLD(F_WD + 2) = ray.d.z;
            COLD(F_L) = L.x; COLD(F_L + 1) = L.y; COLD(F_L + 2) = L.z;
            COLD(F_TP) = Tp.x; COLD(F_TP + 1) = Tp.y; COLD(F_TP + 2) = Tp.z;
            COLD_U(F_DEPTH) = depth;
            start = true;
        } else {
            red_add_f4(prm.sum + COLD_U(F_PIX), L.x, L.y, L.z, 1.0f);   // avg_color += new_color (renderer.rs:39)
            has_path = false;
        }
    }
    const unsigned need = __ballot_sync(0xffffffffu, !has_path);
    if (need) {
        const unsigned n_tiles = (unsigned)(prm.tiles_x * prm.tiles_y);
        if (pool[W_NEXT] >= pool[W_SIZE] && !pool[W_NO_MORE]) {       // warp-uniform: the pool is drained, take the next one
            unsigned item = 0;
            if (lane == 0) item = atomicAdd(prm.work_counter, 1u);
            item = __shfl_sync(0xffffffffu, item, 0);
            if (lane == 0) {
                if (item >= n_tiles * (unsigned)prm.n_chunks) {
                    pool[W_NO_MORE] = 1;
                } else {
                    const unsigned chunk_idx = item / n_tiles, tile = item - chunk_idx * n_tiles;   // chunk-major: concurrent warps spread over tiles
                    const unsigned ty = tile / (unsigned)prm.tiles_x;
                    const int x0 = (int)(tile - ty * (unsigned)prm.tiles_x) * kTileW, y0 = (int)ty * kTileH;
                    const int tw = min(kTileW, C.width - x0), th = min(kTileH, C.height - y0);
                    const int s0 = (int)chunk_idx * prm.chunk;
                    pool[W_X0] = x0; pool[W_Y0] = y0; pool[W_TILE_W] = tw; pool[W_TILE_N] = tw * th;
                    pool[W_SAMPLE0] = s0;
                    pool[W_SIZE] = tw * th * min(prm.chunk, prm.sample_count - s0);
                    pool[W_NEXT] = 0;
                }
            }
            __syncwarp();
        }
        const int pool_next = pool[W_NEXT], pool_size = pool[W_SIZE];
        if (!has_path) {
            const int idx = pool_next + __popc(need & ((1u << lane) - 1u));
            if (idx < pool_size) {
                const int tile_n = pool[W_TILE_N], tile_w = pool[W_TILE_W];
                int pv, sv, tx_, ty_;
                if (tile_n == kTileW * kTileH) { pv = idx & 31; sv = idx >> 5; tx_ = pv & 7; ty_ = pv >> 3; }   // full 8x4 tile
                else { sv = idx / tile_n; pv = idx - sv * tile_n; ty_ = pv / tile_w; tx_ = pv - ty_ * tile_w; }
                const int px = pool[W_X0] + tx_, py = pool[W_Y0] + ty_;
                const int pix = py * C.width + px;                       // renderer.rs:32-33
                const uint32_t sample = (uint32_t)(prm.sample_begin + pool[W_SAMPLE0] + sv);
                key = path_key(prm.seed, (uint32_t)pix, sample);
                const Ray ray = camera_ray<(FEAT & FEAT_DEFOCUS) != 0u>(C, px, py, key);
                COLD(F_WO) = ray.o.x; COLD(F_WO + 1) = ray.o.y; COLD(F_WO + 2) = ray.o.z;
                COLD(F_WD) = ray.d.x; COLD(F_WD + 1) = ray.d.y; COLD(F_WD + 2) = ray.d.z;
                COLD(F_L) = 0.0f; COLD(F_L + 1) = 0.0f; COLD(F_L + 2) = 0.0f;
                COLD(F_TP) = 1.0f; COLD(F_TP + 1) = 1.0f; COLD(F_TP + 2) = 1.0f;
                COLD_U(F_SAMPLE) = sample;
                COLD(F_TIME) = ray.time;
                COLD_U(F_PIX) = (uint32_t)pix;
                COLD_U(F_DEPTH) = 0u;
                COLD_I(F_ORIGIN) = -1;
                depth = 0u;
                has_path = true;
                start = true;
            } else {
                link = (pool[W_NO_MORE] ? (uint32_t)CLS_IDLE : CLS_NEED) << 28;   // pool drained: ask again next round
            }
        }
        const unsigned fresh = __ballot_sync(0xffffffffu, start && depth == 0u);
        __syncwarp();
        if (lane == 0) { pool[W_NEXT] = min(pool_size, pool_next + __popc(need)); pool[W_PATHS] += __popc(fresh); }
    }
    const unsigned starting = __ballot_sync(0xffffffffu, start);
    if (lane == 0) pool[W_SEGS] += __popc(starting);
    __syncwarp();
    if (start) {   // world.hit(ray, [0.001, inf)) begins: hoisted media first, then the op stream from word 0
        const float3 so = f3(COLD(F_WO), COLD(F_WO + 1), COLD(F_WO + 2));
        const float3 sd = f3(COLD(F_WD), COLD(F_WD + 1), COLD(F_WD + 2));
        const float a = dot(sd, sd), inv_a = frcp(a);
        COLD(F_O) = so.x; COLD(F_O + 1) = so.y; COLD(F_O + 2) = so.z;
        COLD(F_D) = sd.x; COLD(F_D + 1) = sd.y; COLD(F_D + 2) = sd.z;
        COLD(F_A) = a; COLD(F_INVA) = inv_a;
        COLD_I(F_XF) = -1;
        Best b;
        b.t = inf; b.op = -1; b.xf = -1;
        media_prepass<FEAT>(S, ops, so, sd, a, inv_a, COLD(F_TIME), tmin, key, depth, b);
        COLD(F_BEST_T) = b.t; COLD_I(F_BEST_OP) = b.op; COLD_I(F_BEST_XF) = -1;
        if (COUNT) cnt[K_MEDIUM] += S.n_media;
        link = prm.first_link;
    }
    return link;
}

// ConstantMedium::hit for a medium that sits in the stream (inside an instance, or with a generic boundary): rare, and
// heavy on registers (free-flight draw, logarithm, boundary programs), so out of line like shade_phase. Updates the
// parked closest hit; returns the lane's next link.
template <class Ops>
__device__ __noinline__ uint32_t medium_phase(uint32_t link) {
    unsigned char* smem = reinterpret_cast<unsigned char*>(dyn_smem);
    const RenderParams& prm = *reinterpret_cast<const RenderParams*>(smem + kSmemParams);
    const DevScene& S = prm.scene;
    const int np = min(S.n_perlin, kMaxPerlinShared);
    float* cold = reinterpret_cast<float*>(smem + kSmemOps + prm.ops_bytes + (size_t)np * (256 * sizeof(float4) + 768));
    const int tid = threadIdx.x;
    Ops ops;
    set_ops_base(ops, S.ops);
    const uint32_t at = link & kLinkMask;
    const float4 w0 = ops(at), w1 = ops(at + 16u);
    const uint4 key = path_key(prm.seed, COLD_U(F_PIX), COLD_U(F_SAMPLE));
    float t;
    int next_word;
    if (medium_test(S, ops, at, w0, w1, f3(COLD(F_O), COLD(F_O + 1), COLD(F_O + 2)), f3(COLD(F_D), COLD(F_D + 1), COLD(F_D + 2)),
                    COLD(F_A), COLD(F_INVA), COLD(F_TIME), 0.001f, COLD(F_BEST_T), key, COLD_U(F_DEPTH), &t, &next_word)) {
        COLD(F_BEST_T) = t; COLD_I(F_BEST_OP) = (int)(at >> 4); COLD_I(F_BEST_XF) = COLD_I(F_XF);
    }
    return ((uint32_t)next_word << 4) | ((uint32_t)fbits(w0.w) & 0xf0000000u);
}

template <bool COUNT, bool OPS_SMEM, unsigned FEAT>
__global__ void __launch_bounds__(kRenderThreads, 1) render_kernel_mk(const RenderParams prm_in) {
    constexpr bool FOLD = (FEAT & FEAT_FOLD) != 0u;
    unsigned char* smem = reinterpret_cast<unsigned char*>(dyn_smem);
    const uint32_t smem_addr = (uint32_t)__cvta_generic_to_shared(smem);
    const int tid = threadIdx.x;
    typedef typename std::conditional<OPS_SMEM, OpsShared, OpsGlobal>::type Ops;

    // ---- stage: the launch parameters, the op stream (one bulk copy per CTA, completion on an mbarrier), the Perlin tables
    for (int k = tid; k < (int)(sizeof(RenderParams) / 4); k += kRenderThreads)
        reinterpret_cast<uint32_t*>(smem + kSmemParams)[k] = reinterpret_cast<const uint32_t*>(&prm_in)[k];
    const uint32_t ops_bytes = prm_in.ops_bytes;
    if (OPS_SMEM) {
        if (tid == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (tid == 0) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr), "r"(ops_bytes) : "memory");
            const char* src = reinterpret_cast<const char*>(prm_in.scene.ops);
            for (uint32_t off = 0; off < ops_bytes; off += 32768u) {
                const uint32_t n = min(32768u, ops_bytes - off);
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(smem_addr + kSmemOps + off), "l"(src + off), "r"(n), "r"(smem_addr) : "memory");
            }
        }
    }
    const int np = min(prm_in.scene.n_perlin, kMaxPerlinShared);
    float4* sh_vec = reinterpret_cast<float4*>(smem + kSmemOps + ops_bytes);
    uint8_t* sh_perm = reinterpret_cast<uint8_t*>(sh_vec + np * 256);
    float* cold = reinterpret_cast<float*>(sh_perm + np * 768);
    int* pool = reinterpret_cast<int*>(cold + kColdFields * kRenderThreads) + (tid >> 5) * kPoolFields;
    if ((tid & 31) < kPoolFields) pool[tid & 31] = 0;
    stage_perlin(prm_in.scene, sh_vec, sh_perm);      // ends with __syncthreads()
    if (OPS_SMEM) {
        uint32_t done;
        do {
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                         : "=r"(done) : "r"(smem_addr), "r"(0) : "memory");
        } while (!done);
    }
    Ops ops;
    set_ops_base(ops, prm_in.scene.ops);
    const DevScene& S = prm_in.scene;
    const float tmin = 0.001f;                     // renderer.rs:144
    const float inf = __int_as_float(0x7f800000);
    const int slab_fast = prm_in.slab_fast, shade_min = prm_in.shade_min, sphere_reps = prm_in.sphere_reps, quad_reps = prm_in.quad_reps;

    // hot per-lane state: the cursor, the per-ray constants of the slab test, the closest hit so far
    uint32_t link = CLS_NEED << 28;                // no path yet: wants one
    float3 inv = f3(0.0f, 0.0f, 0.0f), oi = inv;
    float eps = 0.0f;
    float best_t = inf;
    int best_op = -1, best_xf = -1;
    float4 w0 = make_float4(0, 0, 0, 0), w1 = w0;  // first two words of the lane's next op
    unsigned cnt[COUNT ? K_NUM : 1];               // op counters of the instrumented build (paths / segments: per warp, in the pool)
    for (int k = 0; k < (COUNT ? (int)K_NUM : 1); ++k) cnt[k] = 0;
#define CNT(k) do { if (COUNT) cnt[COUNT ? (k) : 0]++; } while (0)
#define FETCH_NEXT() do { const uint32_t at__ = link & kLinkMask; w0 = ops(at__); w1 = ops(at__ + 16u); } while (0)
#define CUR_O() f3(COLD(F_O), COLD(F_O + 1), COLD(F_O + 2))
#define CUR_D() f3(COLD(F_D), COLD(F_D + 1), COLD(F_D + 2))
#define SET_RAY(o_, d_) do {                                                                       \
        const RaySetup R__ = ray_setup(o_, d_);                                                    \
        inv = R__.inv; oi = R__.oi; eps = R__.eps;                                                 \
        COLD(F_O) = (o_).x; COLD(F_O + 1) = (o_).y; COLD(F_O + 2) = (o_).z;                        \
        COLD(F_D) = (d_).x; COLD(F_D + 1) = (d_).y; COLD(F_D + 2) = (d_).z;                        \
        COLD(F_A) = R__.a; COLD(F_INVA) = R__.inv_a;                                               \
    } while (0)

    // The loop's top is the full vote. The box-test loop is entered either as the vote's pick (a minority round: it lasts
    // until half of its lanes have left the class) or straight after another class has run, when at least slab_fast lanes
    // stand at box-headed ops (no second vote, no dispatch). It leaves with its last lane count, which the next vote uses.
    unsigned n_slab = 0u;
    for (;;) {
        // ---- the vote: lanes per class, five 6-bit counters in one REDUX (CLS_NEED counts as CLS_SHADE, CLS_IDLE as nothing)
        unsigned pick = CLS_SLAB;
        {
            unsigned cls = link >> 28;
            if (cls == CLS_NEED) cls = CLS_SHADE;
            const unsigned tot = __reduce_add_sync(0xffffffffu, cls < CLS_IDLE ? (1u << (6 * cls)) : 0u);
            if (tot == 0u) break;
            const unsigned c_sph = (tot >> 6) & 63u, c_quad = (tot >> 12) & 63u, c_med = (tot >> 18) & 63u, c_shade = (tot >> 24) & 63u;
            // shading runs once enough lanes wait for it (or nothing else can run); otherwise the most populated class
            if (c_shade >= (unsigned)shade_min) pick = CLS_SHADE;
            else {
                unsigned best_n = n_slab;
                if (c_sph > best_n) { pick = CLS_SPHERE; best_n = c_sph; }
                if (c_quad > best_n) { pick = CLS_QUAD; best_n = c_quad; }
                if (c_med > best_n) { pick = CLS_MEDIUM; best_n = c_med; }
                if (best_n == 0u) pick = CLS_SHADE;
            }
            if (COUNT && (tid & 31) == 0) cnt[K_VOTES]++;
        }
        if (COUNT) cnt[K_LANE_OPS] += ((link >> 28) == pick);
        unsigned stay = max(1u, (n_slab + 1u) >> 1);
        if (pick != CLS_SLAB) {
            if (pick == CLS_SPHERE) {
                // the ray is read from shared memory once per visit of the class, not in every repetition (+2.5% random_balls,
                // +1.3% Perlin spheres, +0.4% final_scene, profiles/r2_ab21*)
                const float3 so_ = CUR_O(), sd_ = CUR_D();
                const float sa_ = COLD(F_A), sinva_ = COLD(F_INVA), stime_ = COLD(F_TIME);
                const int sorigin_ = COLD_I(F_ORIGIN);
#pragma unroll 1
                for (int rep = 0; rep < sphere_reps; ++rep) {
                    if ((link >> 28) == CLS_SPHERE) {
                        const uint32_t hdr = (uint32_t)fbits(w0.w);
                        CNT(K_SPHERE);
                        if (COUNT) { if ((hdr >> 12) & FLAG_MOVING) cnt[K_SPHERE_MOVING]++; if ((hdr >> 12) & FLAG_PRECISE) cnt[K_SPHERE_PRECISE]++; }
                        float t;
                        if (sphere_test<FEAT>(S, ops, link, w0, w1, so_, sd_, sa_, sinva_, stime_, tmin, best_t, sorigin_, &t)) {
                            best_t = t; best_op = (int)((link & kLinkMask) >> 4); best_xf = COLD_I(F_XF);
                            CNT(K_SPHERE_HIT);
                        }
                        link = (link & kLinkMask) + (hdr & kHdrFallThrough);
                        FETCH_NEXT();
                    }
                    if (!__any_sync(0xffffffffu, (link >> 28) == CLS_SPHERE)) break;
                }
            } else if (pick == CLS_QUAD) {
                // the ray does not change inside the class: read it from shared memory once, not in every repetition
                // (+2.6% Cornell box, +4.3% Cornell smoke, +5.9% quads, profiles/r2_ab21*)
                const float3 qo = CUR_O(), qd = CUR_D();
                const int qorigin = COLD_I(F_ORIGIN);
#pragma unroll 1
                for (int rep = 0; rep < quad_reps; ++rep) {      // lists of quads sit next to each other in the stream
                    if ((link >> 28) == CLS_QUAD) {
                        const uint32_t hdr = (uint32_t)fbits(w0.w);
                        CNT(K_QUAD);
                        float t;
                        if (quad_test(ops, link, w0, w1, qo, qd, tmin, best_t, qorigin, &t)) {
                            best_t = t; best_op = (int)((link & kLinkMask) >> 4); best_xf = COLD_I(F_XF);
                            CNT(K_QUAD_HIT);
                        }
                        link = (link & kLinkMask) + (hdr & kHdrFallThrough);
                        FETCH_NEXT();
                    }
                    if (!__any_sync(0xffffffffu, (link >> 28) == CLS_QUAD)) break;
                }
            } else if (pick == CLS_MEDIUM) {
                // EVERY lane parks its closest hit (as in the shade branch): all of them reload it below, and a lane of
                // another class that reloaded a stale value would forget the hits it has found in this segment
                COLD(F_BEST_T) = best_t; COLD_I(F_BEST_OP) = best_op; COLD_I(F_BEST_XF) = best_xf;
                if ((link >> 28) == CLS_MEDIUM) {     // a medium that could not be hoisted (inside an instance / generic boundary)
                    CNT(K_MEDIUM);
                    if (FEAT & FEAT_RARE) link = medium_phase<Ops>(link);
                }
                const RaySetup R = ray_setup(CUR_O(), CUR_D());      // nothing live across the call (see the shade branch)
                inv = R.inv; oi = R.oi; eps = R.eps;
                best_t = COLD(F_BEST_T); best_op = COLD_I(F_BEST_OP); best_xf = COLD_I(F_BEST_XF);
                FETCH_NEXT();
            } else {
                // ---- shade / regenerate / start segments, out of line. Nothing of the traversal state stays in registers
                // across the call: the closest hit is parked in shared memory, the per-ray constants are re-derived from the
                // current ray, the op words are fetched again (once per ~25 box tests of every lane: cheap, and it frees the
                // register allocation of the box-test loop from the needs of the shading code).
                COLD(F_BEST_T) = best_t; COLD_I(F_BEST_OP) = best_op; COLD_I(F_BEST_XF) = best_xf;
                link = shade_phase<COUNT, Ops, FEAT>(link, cnt);
                const RaySetup R = ray_setup(CUR_O(), CUR_D());
                inv = R.inv; oi = R.oi; eps = R.eps;
                best_t = COLD(F_BEST_T); best_op = COLD_I(F_BEST_OP); best_xf = COLD_I(F_BEST_XF);
                FETCH_NEXT();
            }
            n_slab = __popc(__ballot_sync(0xffffffffu, link < kSlabLimit));
            if (n_slab < (unsigned)slab_fast) continue;
            stay = (unsigned)slab_fast;
            if (COUNT) cnt[K_LANE_OPS] += (link < kSlabLimit);
        }
        // ---- cull boxes, cube primitives, instance enter / exit: the box-test loop
#pragma unroll 1
        do {
            // every lane runs the box arithmetic on whatever op words it holds (lanes of other classes: harmless, their
            // link and words are kept), so the common case has no divergent branch at all
            const bool in_class = link < kSlabLimit;
            float te, tx;
            slab_ch(w0, w1, inv, oi, &te, &tx);
            const uint32_t hdr = (uint32_t)fbits(w0.w);
            const bool pass = cull_pass(te, tx, tmin, best_t, eps);
            uint32_t nl = pass ? link + hdr : (uint32_t)fbits(w1.w);      // OP_INNER: AABB::hit (aabb.rs:64-84)
            // FOLD (scenes with cube primitives or instances): a cube or an instance whose box the ray misses is finished
            // by the arithmetic above - their second word ends with the link to follow then - and only the accepting
            // side takes the rare branch. Three more instructions per repetition, so scenes without such ops go without.
            const bool rare = (FEAT & (FEAT_FOLD | FEAT_RARE)) != 0u &&      // a stream without either holds nothing but OP_INNER here
                              in_class && (hdr & kHdrNotInner) != 0u && (!FOLD || pass || (hdr & kHdrAlwaysRare) != 0u);
            if (rare) {
                const uint32_t kind = (hdr >> 8) & 15u;
                const uint32_t ft = link + (hdr & kHdrFallThrough);
                nl = ft;
                if (kind == OP_BOX) {
                    CNT(K_BOX);
                    float t;
                    bool win;
                    const int origin = COLD_I(F_ORIGIN);
                    if (!starts_on(origin, link)) {
                        win = box_accept(te, tx, tmin, best_t, &t);
                    } else {
                        const int face = origin & 7;
                        const float ia = (face == 1 || face == 3) ? inv.x : (face >= 4 ? inv.y : inv.z);
                        const bool max_side = face == 0 || face == 1 || face == 4;
                        win = (ia > 0.0f) != max_side &&
                              box_test_from_face(ops(link + 32u), ops(link + 48u), CUR_O(), inv, tmin, best_t, face, &t);
                    }
                    if (win) { best_t = t; best_op = (int)(link >> 4); best_xf = COLD_I(F_XF); CNT(K_BOX_HIT); }
                } else if (kind == OP_XFORM_ENTER) {
                    CNT(K_SLAB);
                    if (!pass) {
                        nl = (uint32_t)fbits(w1.w);
                    } else {
                        CNT(K_XFORM_ENTER);
                        const float4 w2 = ops(link + 32u), w3 = ops(link + 48u);
                        const float3 lo_ = xform_point(f3(COLD(F_WO), COLD(F_WO + 1), COLD(F_WO + 2)), w2, w3);
                        const float3 ld_ = xform_dir(f3(COLD(F_WD), COLD(F_WD + 1), COLD(F_WD + 2)), w2, w3);
                        SET_RAY(lo_, ld_);
                        COLD_I(F_XF) = (int)(link >> 4);
                    }
                } else if (kind == OP_XFORM_EXIT) {
                    const int parent = fbits(w0.x);
                    float3 lo_ = f3(COLD(F_WO), COLD(F_WO + 1), COLD(F_WO + 2));
                    float3 ld_ = f3(COLD(F_WD), COLD(F_WD + 1), COLD(F_WD + 2));
                    if (parent >= 0) {
                        const float4 p2 = ops(((uint32_t)parent << 4) + 32u), p3 = ops(((uint32_t)parent << 4) + 48u);
                        lo_ = xform_point(lo_, p2, p3);
                        ld_ = xform_dir(ld_, p2, p3);
                    }
                    SET_RAY(lo_, ld_);
                    COLD_I(F_XF) = parent;
                } else {
                    CNT(K_SLAB);
                    if (FEAT & FEAT_RARE) nl = aabb_hit_reference(w0, w1, CUR_O(), inv, tmin, best_t) ? ft : (uint32_t)fbits(w1.w);
                }
            } else if (COUNT && in_class) {   // OP_INNER, or a cube / instance finished by the fold
                CNT((hdr & kHdrNotInner) != 0u && ((hdr >> 8) & 15u) == OP_BOX ? K_BOX : K_SLAB);
            }
            link = in_class ? nl : link;
            if (OPS_SMEM) {   // predicated loads (the compiler would branch around them): lanes of other classes keep their words
                asm volatile("{\n.reg .pred p;\nsetp.ne.u32 p, %9, 0;\n@p ld.shared.v4.f32 {%0, %1, %2, %3}, [%8];\n"
                             "@p ld.shared.v4.f32 {%4, %5, %6, %7}, [%8+16];\n}"
                             : "+f"(w0.x), "+f"(w0.y), "+f"(w0.z), "+f"(w0.w), "+f"(w1.x), "+f"(w1.y), "+f"(w1.z), "+f"(w1.w)
                             : "r"(smem_addr + kSmemOps + (link & kLinkMask)), "r"((uint32_t)in_class));
            } else if (in_class) {
                FETCH_NEXT();
            }
            n_slab = __popc(__ballot_sync(0xffffffffu, link < kSlabLimit));
        } while (n_slab >= stay);
    }
#undef CNT
#undef FETCH_NEXT
#undef CUR_O
#undef CUR_D
#undef SET_RAY
    if (COUNT) {
        for (int k = 2; k < (int)K_NUM; ++k) {   // one atomic per warp and counter
            unsigned long long v = cnt[k];
            for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
            if ((tid & 31) == 0 && v) atomicAdd(prm_in.stats + k, v);
        }
    }
    __syncwarp();
    if ((tid & 31) == 0) {
        atomicAdd(prm_in.stats + K_PATHS, (unsigned long long)pool[W_PATHS]);
        atomicAdd(prm_in.stats + K_SEGMENTS, (unsigned long long)pool[W_SEGS]);
    }
}
#undef COLD
#undef COLD_U
#undef COLD_I
