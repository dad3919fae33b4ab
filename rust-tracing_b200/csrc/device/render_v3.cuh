// K1: the render loop (renderer.rs:26-49,139-155) as a per-lane state machine with warp-level class voting.
//
// Every lane is always somewhere in {a traversal op of some class, waiting to shade / get a new path}. Each
// iteration the warp counts its lanes per class, runs only the most populated class (lanes of other classes wait,
// which costs no issue slots), and lanes whose op finished move on to their next op's class - known from the header
// bits before the op's words arrive (dev_scene.h). Lanes regroup by what they are about to execute instead of
// idling behind the longest traversal or the rarest op kind of the warp. This is the third form of the kernel:
//
// what changed against the second (ncu: 120 registers -> 4 warps per scheduler, 'wait' + 'branch_resolving' stalls
// dominate, issue slots 52% busy):
//  * cold per-path state (world-space ray, radiance sum, throughput, RNG key, pixel, depth, time) lives in
//    shared memory, one SoA column per thread, and is touched only when a segment is shaded; the traversal
//    loop keeps just the cursor (local ray, reciprocal, best hit, op index, class) in registers, so the kernel
//    fits 64 registers and 8 blocks (32 warps) per SM without the compiler spilling inside the loop;
//  * OP_INNER, by far the most frequent op, has a hand-trimmed fast path; the other slab-class ops share it;
//  * the vote takes a one-ballot fast path while the slab class holds enough lanes;
//  * the op stream carries two padding words so the next op's words are fetched unconditionally.
//
// Since then: the vote takes quorums for minority classes and an exit threshold for the slab repetitions as runtime
// parameters (their defaults reproduce the rule above: a sweep found nothing better, profiles/r1_ab18*), and the FOLD
// template parameter lets OP_BOX share OP_INNER's arithmetic (picked per scene by launch_render).
//
// Included by rt_cuda.cu.
#pragma once


#ifdef RT_OPT_SLAB_REPS
constexpr int kSlabReps = RT_OPT_SLAB_REPS;
#else
constexpr int kSlabReps = 8;       // consecutive slab-class ops per vote (8 and 16 measured equal, 4 slower)
#endif
constexpr int kColdFields = 16;   // world o(3) d(3), L(3), Tp(3), sample index, time, pixel, depth (the RNG key is re-derived from pixel, sample)

inline size_t v3_smem_bytes(int n_perlin) {
    const int np = n_perlin < kMaxPerlinShared ? n_perlin : kMaxPerlinShared;
    return (size_t)np * (256 + 48) * sizeof(float4) + (size_t)kColdFields * kBlockThreads * sizeof(float);
}

template <bool COUNT, int MIN_BLOCKS, bool FOLD>
__global__ void __launch_bounds__(kBlockThreads, MIN_BLOCKS) render_kernel_v3(const RenderParams prm) {
    // dynamic shared memory: [n_perlin x 256 float4 gradients][n_perlin x 768 B permutations][cold state]
    const int np = min(prm.scene.n_perlin, kMaxPerlinShared);
    float4* sh_vec = dyn_smem;
    uint8_t* sh_perm = reinterpret_cast<uint8_t*>(dyn_smem + np * 256);
    float* cold = reinterpret_cast<float*>(dyn_smem + np * (256 + 48));
    stage_perlin(prm.scene, sh_vec, sh_perm);
    PerlinShared P{sh_vec, sh_perm};
    const DevScene& S = prm.scene;
    const DevCamera& C = prm.cam;
    const float4* __restrict__ ops = S.ops;
    const float tmin = 0.001f;                     // renderer.rs:144
    const float inf = __int_as_float(0x7f800000);
    const int tid = threadIdx.x;
#define COLD(f) cold[(f) * kBlockThreads + tid]
#define COLD_U(f) reinterpret_cast<uint32_t*>(cold)[(f) * kBlockThreads + tid]

    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt_mask = (1u << lane) - 1u;
    const unsigned n_tiles = (unsigned)(prm.tiles_x * prm.tiles_y);
    const unsigned n_items = n_tiles * (unsigned)prm.n_chunks;

    // warp-uniform pool of (tile x sample chunk) paths
    int pool_next = 0, pool_size = 0;
    int tile_x0 = 0, tile_y0 = 0, tile_w = 1, tile_n = 1;
    int pool_sample0 = 0;
    bool no_more = false;

    // hot per-lane state
    Trav T;
    T.i = 0; T.cur_xf = -1; T.best.op = -1; T.best.xf = -1; T.best.t = inf;
    T.o = T.d = T.inv = T.so = T.sd = f3(0.0f, 0.0f, 0.0f);
    int origin = -1;
    bool has_path = false;
    uint32_t cls = CLS_SHADE;                      // no path yet: wants one
    float4 w0 = make_float4(0, 0, 0, 0), w1 = w0;  // first two words of the lane's next op
    unsigned cnt[COUNT ? K_NUM : 2];
    for (int k = 0; k < (COUNT ? (int)K_NUM : 2); ++k) cnt[k] = 0;
#define CNT(k) do { if (COUNT || (k) < 2) cnt[(COUNT || (k) < 2) ? (k) : 0]++; } while (0)
#define FETCH_NEXT() do { w0 = __ldg(ops + T.i); w1 = __ldg(ops + T.i + 1); } while (0)
#define V3_INNER_STEP(HDR) do {                                                                                   \
        const float ax = (w0.x - T.o.x) * T.inv.x, bx = (w1.x - T.o.x) * T.inv.x;                                 \
        const float ay = (w0.y - T.o.y) * T.inv.y, by = (w1.y - T.o.y) * T.inv.y;                                 \
        const float az = (w0.z - T.o.z) * T.inv.z, bz = (w1.z - T.o.z) * T.inv.z;                                 \
        const bool sx = T.inv.x < 0.0f, sy = T.inv.y < 0.0f, sz = T.inv.z < 0.0f;                                 \
        const float te = fmaxf(fmaxf(fmaxf(sx ? bx : ax, sy ? by : ay), sz ? bz : az), tmin);                     \
        const float tx = fminf(fminf(fminf(sx ? ax : bx, sy ? ay : by), sz ? az : bz), T.best.t);                 \
        const bool hit = te <= tx * 1.0000012f;     /* te >= tmin > 0, so a negative tx can never pass */         \
        T.i = hit ? T.i + 2 : fbits(w1.w);                                                                        \
        cls = ((HDR) >> (hit ? 8 : 11)) & 7u;                                                                     \
    } while (0)

    for (;;) {
        unsigned pick;
        {
            const unsigned n_slab = __popc(__ballot_sync(0xffffffffu, cls == CLS_SLAB));
            if (n_slab >= (unsigned)prm.slab_fast) {
                pick = CLS_SLAB;
            } else {
                // lanes per class: five 6-bit counters in one REDUX, plus a ballot for the box class
                const unsigned tot = __reduce_add_sync(0xffffffffu, cls < CLS_IDLE ? (1u << (6 * cls)) : 0u);
                const unsigned c_box = __popc(__ballot_sync(0xffffffffu, cls == CLS_BOX));
                if (tot == 0u && c_box == 0u) break;
                const unsigned c_sph = (tot >> 6) & 63u, c_quad = (tot >> 12) & 63u, c_med = (tot >> 18) & 63u, c_shade = (tot >> 24) & 63u;
                // a class that has gathered its quorum runs before the slab class gets the warp back; otherwise the
                // most populated class runs (shading only with its quorum, or when nothing else can run)
                if (c_shade >= (unsigned)prm.shade_min) pick = CLS_SHADE;
                else if (c_sph >= (unsigned)prm.sphere_min) pick = CLS_SPHERE;
                else if (c_box >= (unsigned)prm.box_min) pick = CLS_BOX;
                else if (c_quad >= (unsigned)prm.quad_min) pick = CLS_QUAD;
                else {
                    unsigned best_n = n_slab;
                    pick = CLS_SLAB;
                    if (c_sph > best_n) { pick = CLS_SPHERE; best_n = c_sph; }
                    if (c_box > best_n) { pick = CLS_BOX; best_n = c_box; }
                    if (c_quad > best_n) { pick = CLS_QUAD; best_n = c_quad; }
                    if (c_med > best_n) { pick = CLS_MEDIUM; best_n = c_med; }
                    if (best_n == 0u) pick = CLS_SHADE;
                }
            }
            if (COUNT) { if (lane == 0) cnt[K_VOTES]++; cnt[K_LANE_OPS] += (cls == pick); }
        }

        if (pick == CLS_SLAB) {
#pragma unroll 1
                        for (int rep = 0; rep < kSlabReps; ++rep) {
                if (cls == CLS_SLAB) {
                    const uint32_t hdr = (uint32_t)fbits(w0.w);
                    const uint32_t kind = hdr & 15u;
                    bool refetch = true;
                    if (COUNT) { cnt[kind == OP_BOX ? K_BOX : K_SLAB]++; if (kind == OP_XFORM_ENTER) cnt[K_XFORM_ENTER]++; }
                    // FOLD: OP_BOX shares the slab arithmetic of OP_INNER (same lanes, same instructions) and differs only
                    // in what it does with the unclamped entry / exit parameters. It costs every inner node ~5 more
                    // instructions and saves the divergent OP_BOX path, which ran at 2.3 lanes and was 15% of all issued
                    // instructions on final_scene once the stream was pruned: +3.6% there, -2..-3.5% on scenes with few or
                    // no cubes (profiles/r1_ab23_fold_box.log), so the host picks it per scene (launch_render).
                    const bool box_fast = FOLD && kind == OP_BOX && !((origin >> 3) == T.i && origin >= 0);
                    if (kind == OP_INNER || box_fast) {
                        if (FOLD) {
                            const float ax = (w0.x - T.o.x) * T.inv.x, bx = (w1.x - T.o.x) * T.inv.x;
                            const float ay = (w0.y - T.o.y) * T.inv.y, by = (w1.y - T.o.y) * T.inv.y;
                            const float az = (w0.z - T.o.z) * T.inv.z, bz = (w1.z - T.o.z) * T.inv.z;
                            const bool sx = T.inv.x < 0.0f, sy = T.inv.y < 0.0f, sz = T.inv.z < 0.0f;
                            const float te_raw = fmaxf(fmaxf(sx ? bx : ax, sy ? by : ay), sz ? bz : az);
                            const float tx_raw = fminf(fminf(sx ? ax : bx, sy ? ay : by), sz ? az : bz);
                            if (box_fast) {      // Quad::cube as one slab primitive, see op_box()
                                float t = te_raw;
                                if (!(tmin <= t && t <= T.best.t)) t = tx_raw;
                                if (te_raw <= tx_raw && tmin <= t && t <= T.best.t) {
                                    if (COUNT) cnt[K_BOX_HIT]++;
                                    T.best.t = t; T.best.op = T.i; T.best.xf = T.cur_xf;
                                }
                                T.i += 3;
                                cls = (hdr >> 8) & 7u;
                            } else {             // AABB::hit (aabb.rs:64-84), tight slab form
                                const float te = fmaxf(te_raw, tmin), tx = fminf(tx_raw, T.best.t);
                                const bool hit = te <= tx * 1.0000012f;
                                T.i = hit ? T.i + 2 : fbits(w1.w);
                                cls = (hdr >> (hit ? 8 : 11)) & 7u;
                            }
                        } else {
                            // AABB::hit (aabb.rs:64-84), tight slab form; see slab_interval() for the NaN / sign rules
                            V3_INNER_STEP(hdr);
#ifdef RT_OPT_SLAB_DOUBLE
                            // a second inner node in the same repetition for the lanes that are at one again: saves the
                            // loop control and the class test between the two (measured: -7% final_scene, +2..5% random_balls)
                            FETCH_NEXT();
                            const uint32_t hdr2 = (uint32_t)fbits(w0.w);
                            if (cls == CLS_SLAB && (hdr2 & 15u) == OP_INNER) {
                                if (COUNT) cnt[K_SLAB]++;
                                V3_INNER_STEP(hdr2);
                            } else {
                                refetch = false;
                            }
#endif
                        }
                    } else {
                        const float tb = T.best.t;   // BOX, XFORM_ENTER / EXIT, INNER_REF (the world ray stays in COLD)
                        cls = op_slab_class(S, T, w0, w1, tmin, origin, [&](float3& wo, float3& wd) {
                            wo = f3(COLD(0), COLD(1), COLD(2));
                            wd = f3(COLD(3), COLD(4), COLD(5));
                        });
                        if (COUNT && T.best.t != tb) cnt[K_BOX_HIT]++;
                    }
                    if (refetch) FETCH_NEXT();
                }
                if (__popc(__ballot_sync(0xffffffffu, cls == CLS_SLAB)) < (unsigned)prm.slab_exit) break;   // too few left: vote again
            }
        } else if (pick == CLS_BOX) {
            if (cls == CLS_BOX) {        // OP_BOX as its own class (CompileOptions::box_class)
                if (COUNT) cnt[K_BOX]++;
                const float tb = T.best.t;
                cls = op_slab_class(S, T, w0, w1, tmin, origin, [&](float3& wo, float3& wd) {
                    wo = f3(COLD(0), COLD(1), COLD(2));
                    wd = f3(COLD(3), COLD(4), COLD(5));
                });
                if (COUNT && T.best.t != tb) cnt[K_BOX_HIT]++;
                FETCH_NEXT();
            }
        } else if (pick == CLS_SPHERE) {
#pragma unroll 1
            for (int rep = 0; rep < prm.sphere_reps; ++rep) {
                if (cls == CLS_SPHERE) {
                    const uint32_t hdr = (uint32_t)fbits(w0.w);
                    CNT(K_SPHERE);
                    if (COUNT) { if ((hdr >> 4) & FLAG_MOVING) cnt[K_SPHERE_MOVING]++; if ((hdr >> 4) & FLAG_PRECISE) cnt[K_SPHERE_PRECISE]++; }
                    const float tb = T.best.t;
                    op_sphere(S, T, w0, w1, COLD(13), tmin, origin);
                    if (COUNT && T.best.t != tb) cnt[K_SPHERE_HIT]++;
                    cls = (hdr >> 8) & 7u;
                    FETCH_NEXT();
                }
                if (!__any_sync(0xffffffffu, cls == CLS_SPHERE)) break;
            }
        } else if (pick == CLS_QUAD) {
            if (cls == CLS_QUAD) {
                const uint32_t hdr = (uint32_t)fbits(w0.w);
                CNT(K_QUAD);
                const float tb = T.best.t;
                op_quad(S, T, w0, w1, tmin, origin);
                if (COUNT && T.best.t != tb) cnt[K_QUAD_HIT]++;
                cls = (hdr >> 8) & 7u;
                FETCH_NEXT();
            }
        } else if (pick == CLS_MEDIUM) {
            if (cls == CLS_MEDIUM) {     // a medium that could not be hoisted (inside an instance / generic boundary)
                const uint32_t hdr = (uint32_t)fbits(w0.w);
                CNT(K_MEDIUM);
                const float tb = T.best.t;
                const uint4 key = path_key(prm.seed, COLD_U(14), COLD_U(12));
                op_medium(S, T, w0, w1, COLD(13), tmin, key, COLD_U(15));
                if (COUNT && T.best.t != tb) cnt[K_MEDIUM_HIT]++;
                cls = (hdr >> 8) & 7u;
                FETCH_NEXT();
            }
        } else {
            // ---- shade the finished segment (renderer.rs:144-153), hand out new paths, start the next segments ----
            bool start = false;                 // this lane begins a new segment below (one shared copy of that code)
            uint4 key = make_uint4(0, 0, 0, 0);
            uint32_t depth = 0;
            if (cls == CLS_SHADE && has_path) {
                Ray ray;
                ray.o = f3(COLD(0), COLD(1), COLD(2));
                ray.d = f3(COLD(3), COLD(4), COLD(5));
                ray.time = COLD(13);
                float3 L = f3(COLD(6), COLD(7), COLD(8)), Tp = f3(COLD(9), COLD(10), COLD(11));
                key = path_key(prm.seed, COLD_U(14), COLD_U(12));
                depth = COLD_U(15);
                bool alive;
                if (T.best.op < 0) {
                    L = L + Tp * C.background;                                  // renderer.rs:152-153
                    alive = false;
                } else {
                    HitRec h;
                    finalize_hit(S, ray, T.best, h);
                    if (COUNT) {
                        if (T.best.xf >= 0) cnt[K_FINALIZE_XFORM]++;
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
                    origin = h.origin;
                    ++depth;
                    if ((int)depth >= C.max_depth) alive = false;               // renderer.rs:140-142
                }
                if (alive) {
                    COLD(0) = ray.o.x; COLD(1) = ray.o.y; COLD(2) = ray.o.z;
                    COLD(3) = ray.d.x; COLD(4) = ray.d.y; COLD(5) = ray.d.z;
                    COLD(6) = L.x; COLD(7) = L.y; COLD(8) = L.z;
                    COLD(9) = Tp.x; COLD(10) = Tp.y; COLD(11) = Tp.z;
                    COLD_U(15) = depth;
                    start = true;
                } else {
                    red_add_f4(prm.sum + COLD_U(14), L.x, L.y, L.z, 1.0f);      // avg_color += new_color (renderer.rs:39)
                    has_path = false;
                }
            }
            const unsigned need = __ballot_sync(0xffffffffu, !has_path);
            if (need) {
                if (pool_next >= pool_size && !no_more) {
                    unsigned item = 0;
                    if (lane == 0) item = atomicAdd(prm.work_counter, 1u);
                    item = __shfl_sync(0xffffffffu, item, 0);
                    if (item >= n_items) {
                        no_more = true;
                    } else {
                        const unsigned chunk_idx = item / n_tiles, tile = item - chunk_idx * n_tiles;   // chunk-major: concurrent warps spread over tiles
                        const unsigned ty = tile / (unsigned)prm.tiles_x;
                        tile_x0 = (int)(tile - ty * (unsigned)prm.tiles_x) * kTileW;
                        tile_y0 = (int)ty * kTileH;
                        tile_w = min(kTileW, C.width - tile_x0);
                        const int tile_h = min(kTileH, C.height - tile_y0);
                        tile_n = tile_w * tile_h;
                        const int s0 = (int)chunk_idx * prm.chunk;
                        const int ns = min(prm.chunk, prm.sample_count - s0);
                        pool_sample0 = s0;
                        pool_size = tile_n * ns;
                        pool_next = 0;
                    }
                }
                if (!has_path) {
                    const int idx = pool_next + __popc(need & lt_mask);
                    if (idx < pool_size) {
                        int pv, sv, tx_, ty_;
                        if (tile_n == kTileW * kTileH) { pv = idx & 31; sv = idx >> 5; tx_ = pv & 7; ty_ = pv >> 3; }   // full 8x4 tile
                        else { sv = idx / tile_n; pv = idx - sv * tile_n; ty_ = pv / tile_w; tx_ = pv - ty_ * tile_w; }
                        const int px = tile_x0 + tx_, py = tile_y0 + ty_;
                        const int pix = py * C.width + px;                       // renderer.rs:32-33
                        const uint32_t sample = (uint32_t)(prm.sample_begin + pool_sample0 + sv);
                        key = path_key(prm.seed, (uint32_t)pix, sample);
                        const Ray ray = camera_ray(C, px, py, key);
                        COLD(0) = ray.o.x; COLD(1) = ray.o.y; COLD(2) = ray.o.z;
                        COLD(3) = ray.d.x; COLD(4) = ray.d.y; COLD(5) = ray.d.z;
                        COLD(6) = 0.0f; COLD(7) = 0.0f; COLD(8) = 0.0f;
                        COLD(9) = 1.0f; COLD(10) = 1.0f; COLD(11) = 1.0f;
                        COLD_U(12) = sample;
                        COLD(13) = ray.time;
                        COLD_U(14) = (uint32_t)pix;
                        COLD_U(15) = 0u;
                        depth = 0u;
                        origin = -1;
                        has_path = true;
                        start = true;
                        CNT(K_PATHS);
                    } else {
                        cls = no_more ? (uint32_t)CLS_IDLE : (uint32_t)CLS_SHADE;   // pool drained: ask again next round
                    }
                }
                pool_next = min(pool_size, pool_next + __popc(need));
            }
            if (start) {   // world.hit(ray, [0.001, inf)) begins: hoisted media first, then the op stream from word 0
                Ray ray;
                ray.o = f3(COLD(0), COLD(1), COLD(2));
                ray.d = f3(COLD(3), COLD(4), COLD(5));
                ray.time = COLD(13);
                trav_begin(T, ray, 0, inf);
                media_prepass(S, T, ray.time, tmin, key, depth);
                if (COUNT) cnt[K_MEDIUM] += S.n_media;
                CNT(K_SEGMENTS);
                cls = (uint32_t)prm.first_class;
                FETCH_NEXT();
            }
        }
    }
#undef CNT
#undef FETCH_NEXT
#undef V3_INNER_STEP
#undef COLD
#undef COLD_U
    for (int k = 0; k < (COUNT ? (int)K_NUM : 2); ++k) {   // one atomic per warp and counter
        unsigned long long v = cnt[k];
        for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
        if (lane == 0 && v) atomicAdd(prm.stats + k, v);
    }
}
