// K1 v2: the render loop (renderer.rs:26-49,139-155) as a per-lane state machine with warp-level class voting.
//
// Every lane is always somewhere in {a traversal op of some class, waiting to shade / get a new path}. Each
// iteration the warp counts its lanes per class with one REDUX, runs only the most populated class (lanes of
// other classes wait, which costs no issue slots), and lanes whose op finished move on to their next op's class —
// known from the header bits before the op's words arrive (dev_scene.h). Lanes therefore regroup by what they
// are about to execute instead of idling behind the longest traversal or the rarest op kind of the warp.
//
// Included by rt_cuda.cu after RenderParams / Counter / red_add_f4 / stage_perlin are defined.
#pragma once

constexpr int kSlabReps = 8;

template <bool COUNT, int MIN_BLOCKS>
__global__ void __launch_bounds__(kBlockThreads, MIN_BLOCKS) render_kernel_v2(const RenderParams prm) {
    float4* sh_vec = dyn_smem;
    uint8_t* sh_perm = reinterpret_cast<uint8_t*>(dyn_smem + kMaxPerlinShared * 256);
    stage_perlin(prm.scene, sh_vec, sh_perm);
    PerlinShared P{sh_vec, sh_perm};
    const DevScene& S = prm.scene;
    const DevCamera& C = prm.cam;
    const float4* __restrict__ ops = S.ops;
    const int end = S.n_words;
    const float tmin = 0.001f;                     // renderer.rs:144
    const float inf = __int_as_float(0x7f800000);

    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt_mask = (1u << lane) - 1u;
    const unsigned n_tiles = (unsigned)(prm.tiles_x * prm.tiles_y);
    const unsigned n_items = n_tiles * (unsigned)prm.n_chunks;

    // warp-uniform pool of (tile x sample chunk) paths
    int pool_next = 0, pool_size = 0;
    int tile_x0 = 0, tile_y0 = 0, tile_w = 1, tile_n = 1;
    int64_t pool_sample0 = 0;
    bool no_more = false;

    // per-lane state
    Trav T;
    T.i = end; T.cur_xf = -1; T.best.op = -1; T.best.xf = -1; T.best.t = inf;
    T.o = T.d = T.inv = T.so = T.sd = f3(0.0f, 0.0f, 0.0f);
    float time = 0.0f;
    float3 L = f3(0.0f, 0.0f, 0.0f), Tp = f3(1.0f, 1.0f, 1.0f);
    int depth = 0, origin = -1, pix = 0;
    uint4 key = make_uint4(0, 0, 0, 0);
    bool has_path = false;
    uint32_t cls = CLS_SHADE;                      // no path yet: wants one
    float4 w0 = make_float4(0, 0, 0, 0), w1 = w0;  // first two words of the lane's next op
    unsigned cnt[COUNT ? K_NUM : 2];               // [K_PATHS], [K_SEGMENTS] always; the rest in the counting build
    for (int k = 0; k < (COUNT ? (int)K_NUM : 2); ++k) cnt[k] = 0;
#define CNT(k) do { if (COUNT || (k) < 2) cnt[(COUNT || (k) < 2) ? (k) : 0]++; } while (0)
#define FETCH_NEXT() do { if (T.i < end) { w0 = __ldg(ops + T.i); w1 = __ldg(ops + T.i + 1); } } while (0)

    for (;;) {
        const unsigned tot = __reduce_add_sync(0xffffffffu, cls < CLS_IDLE ? (1u << (6 * cls)) : 0u);
        if (tot == 0u) break;
        const unsigned c_slab = tot & 63u, c_sph = (tot >> 6) & 63u, c_quad = (tot >> 12) & 63u, c_med = (tot >> 18) & 63u,
                       c_shade = (tot >> 24) & 63u;
        unsigned pick = CLS_SLAB, best_n = c_slab;
        if (c_sph > best_n) { pick = CLS_SPHERE; best_n = c_sph; }
        if (c_quad > best_n) { pick = CLS_QUAD; best_n = c_quad; }
        if (c_med > best_n) { pick = CLS_MEDIUM; best_n = c_med; }
        if ((c_shade > best_n && c_shade >= (unsigned)prm.shade_min) || best_n == 0u) { pick = CLS_SHADE; best_n = c_shade; }
        if (COUNT) { if (lane == 0) cnt[K_VOTES]++; cnt[K_LANE_OPS] += (cls == pick); }

        if (pick == CLS_SLAB) {
#pragma unroll 1
            for (int rep = 0; rep < kSlabReps; ++rep) {
                if (cls == CLS_SLAB) {
                    if (COUNT) { const uint32_t kd = (uint32_t)fbits(w0.w) & 15u; cnt[kd == OP_BOX ? K_BOX : K_SLAB]++; if (kd == OP_XFORM_ENTER) cnt[K_XFORM_ENTER]++; }
                    const float tb = T.best.t;
                    cls = op_slab_class(S, T, w0, w1, tmin, origin);
                    if (COUNT && T.best.t != tb) cnt[K_BOX_HIT]++;
                    FETCH_NEXT();
                }
                if (!__any_sync(0xffffffffu, cls == CLS_SLAB)) break;
            }
        } else if (pick == CLS_SPHERE) {
#pragma unroll 1
            for (int rep = 0; rep < 2; ++rep) {
                if (cls == CLS_SPHERE) {
                    const uint32_t hdr = (uint32_t)fbits(w0.w);
                    CNT(K_SPHERE);
                    if (COUNT) { if ((hdr >> 4) & FLAG_MOVING) cnt[K_SPHERE_MOVING]++; if ((hdr >> 4) & FLAG_PRECISE) cnt[K_SPHERE_PRECISE]++; }
                    const float tb = T.best.t;
                    op_sphere(S, T, w0, w1, time, tmin, origin);
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
            if (cls == CLS_MEDIUM) {
                const uint32_t hdr = (uint32_t)fbits(w0.w);
                CNT(K_MEDIUM);
                const float tb = T.best.t;
                op_medium(S, T, w0, w1, time, tmin, key, (uint32_t)depth);
                if (COUNT && T.best.t != tb) cnt[K_MEDIUM_HIT]++;
                cls = (hdr >> 8) & 7u;
                FETCH_NEXT();
            }
        } else {
            // ---- shade the finished segment (renderer.rs:144-153), then hand out new paths ----
            if (cls == CLS_SHADE && has_path) {
                Ray ray; ray.o = T.so; ray.d = T.sd; ray.time = time;          // world-space ray of the segment
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
                    alive = shade(S, P, ray, h, key, (uint32_t)depth, L, Tp);
                    origin = h.origin;
                    ++depth;
                    if (depth >= C.max_depth) alive = false;                    // renderer.rs:140-142
                }
                if (alive) {
                    trav_begin(T, ray, 0, inf);
                    media_prepass(S, T, time, tmin, key, (uint32_t)depth);
                    if (COUNT) cnt[K_MEDIUM] += S.n_media;
                    CNT(K_SEGMENTS);
                    cls = (uint32_t)prm.first_class;
                    w0 = __ldg(ops); w1 = __ldg(ops + 1);
                } else {
                    red_add_f4(prm.sum + pix, L.x, L.y, L.z, 1.0f);             // avg_color += new_color (renderer.rs:39)
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
                if (!has_path) {
                    const int idx = pool_next + __popc(need & lt_mask);
                    if (idx < pool_size) {
                        const int pv = idx % tile_n, sv = idx / tile_n;
                        const int px = tile_x0 + pv % tile_w, py = tile_y0 + pv / tile_w;
                        pix = py * C.width + px;                                 // renderer.rs:32-33
                        key = path_key(prm.seed, (uint32_t)pix, (uint32_t)(pool_sample0 + sv));
                        const Ray ray = camera_ray(C, px, py, key);
                        time = ray.time;
                        L = f3(0.0f, 0.0f, 0.0f);
                        Tp = f3(1.0f, 1.0f, 1.0f);
                        depth = 0;
                        origin = -1;
                        has_path = true;
                        trav_begin(T, ray, 0, inf);
                        media_prepass(S, T, time, tmin, key, 0u);
                        if (COUNT) cnt[K_MEDIUM] += S.n_media;
                        CNT(K_PATHS); CNT(K_SEGMENTS);
                        cls = (uint32_t)prm.first_class;
                        w0 = __ldg(ops); w1 = __ldg(ops + 1);
                    } else {
                        cls = no_more ? (uint32_t)CLS_IDLE : (uint32_t)CLS_SHADE;   // pool drained: ask again next round
                    }
                }
                pool_next = min(pool_size, pool_next + __popc(need));
            }
        }
    }
#undef CNT
#undef FETCH_NEXT
    for (int k = 0; k < (COUNT ? (int)K_NUM : 2); ++k) {   // one atomic per warp and counter
        unsigned long long v = cnt[k];
        for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
        if (lane == 0 && v) atomicAdd(prm.stats + k, v);
    }
}
