// K1 as a wavefront: the loop of renderer.rs:26-49,139-155 split into two kernels over a pool of in-flight paths
// that lives in HBM / L2 (structure of arrays of float4, one column per path slot).
//
//   wf_shade_kernel   one lane per slot: finish the segment the extend kernel traced (renderer.rs:144-153: background,
//                     emitted, scatter), add finished paths to the SUM framebuffer (renderer.rs:39), hand the slot a new
//                     camera path (camera.rs:112-137) when its path ended, evaluate the hoisted media for the next
//                     segment, and write the ray back. Every lane of a warp shades together.
//   wf_extend_kernel  world.hit(ray, [0.001, inf)) (bvh.rs:90-113 and below): persistent warps, each LANE takes the
//                     next ray of the pool the moment its own traversal ends (ballot/popc ranking inside a window of
//                     slots the warp claimed with one atomic). Lanes still regroup by op class through the warp vote
//                     of render_v3.cuh, but nobody waits for shading any more: the only parked lanes are those of
//                     minority op classes and the few waiting for the next (cheap) fetch round.
//
// Why: in the megakernel (render_v3.cuh) a lane whose segment ended waits until 24 lanes can shade together, so the
// slab loop - 56% of all issued instructions - ran with 10 of 32 lanes (ncu source view of capture r1_e). Here the
// per-ray state a lane must load to start a segment is three float4 words and the result it leaves is one.
//
// The host loop (launch_render_v4 in rt_cuda.cu) alternates the two kernels until no slot carries a ray. Per-path
// results are those of the megakernel bit for bit (same keyed RNG, same device functions); only the order in which
// finished paths are added to the framebuffer differs.
//
// Included by rt_cuda.cu.
#pragma once

constexpr int kWfWindow = 256;      // slots a warp of the extend kernel claims per atomic
constexpr int kWfReserve = 128;     // path indices a warp of the shade kernel claims per atomic
constexpr int kWfShadeThreads = 128;
constexpr int WF_EMPTY = -3;        // hit.op sentinels: slot never held a path / slot retired (no paths left); -1 = miss
constexpr int WF_DEAD = -2;

struct WavePool {
    float4* ray0;   // {o.xyz, time}
    float4* ray1;   // {d.xyz, depth}
    float4* hit;    // {t, op, xf, origin code}: written by shade (hoisted media / sentinels), updated by extend
    float4* st0;    // {L.xyz, pixel}
    float4* st1;    // {throughput.xyz, sample}
    int n_slots;    // multiple of 32
};

struct WfParams {
    DevScene scene;
    DevCamera cam;
    uint64_t seed;
    int64_t sample_begin;
    unsigned long long n_paths;          // W*H*sample_count
    uint32_t n_pixels;
    int tiled, tiles_x;                  // path index -> pixel through 8x4 tiles when the image divides evenly
    WavePool pool;
    float4* sum;
    unsigned long long* path_counter;    // next unclaimed path index
    unsigned int* slot_cursor;           // extend kernel: next unclaimed slot (reset by the shade kernel)
    ulonglong2* reserve;                 // per shade warp: claimed but unused path indices [x, y), kept across launches
    unsigned int* live_flag;             // [iteration in batch]: some slot carries a ray after this shade pass
    unsigned long long* stats;           // [0] paths started, [1] segments
    int first_class, fetch_min, slab_fast, sphere_reps, slab_exit, sphere_min;
};

__device__ __forceinline__ float4 f4i(float x, int y, int z, int w) {
    return make_float4(x, __int_as_float(y), __int_as_float(z), __int_as_float(w));
}

__global__ void wf_init_kernel(float4* hit, int n) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) hit[k] = f4i(0.0f, WF_EMPTY, -1, -1);
}

template <int MIN_BLOCKS>
__global__ void __launch_bounds__(kWfShadeThreads, MIN_BLOCKS) wf_shade_kernel(const WfParams prm, const int iter) {
    const int np = min(prm.scene.n_perlin, kMaxPerlinShared);
    float4* sh_vec = dyn_smem;
    uint8_t* sh_perm = reinterpret_cast<uint8_t*>(dyn_smem + np * 256);
    __shared__ unsigned int sh_live, sh_started;
    if (threadIdx.x == 0) { sh_live = 0u; sh_started = 0u; }
    stage_perlin(prm.scene, sh_vec, sh_perm);          // ends with __syncthreads()
    PerlinShared P{sh_vec, sh_perm};
    const DevScene& S = prm.scene;
    const DevCamera& C = prm.cam;
    const WavePool& pool = prm.pool;
    const float tmin = 0.001f;                     // renderer.rs:144
    const float inf = __int_as_float(0x7f800000);
    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt_mask = (1u << lane) - 1u;
    const int warps_per_block = kWfShadeThreads / 32;
    const int gw = blockIdx.x * warps_per_block + (int)(threadIdx.x >> 5);
    const int n_warps = gridDim.x * warps_per_block;

    if (blockIdx.x == 0 && threadIdx.x == 0) *prm.slot_cursor = 0u;   // the next extend pass starts from slot 0

    const ulonglong2 res0 = prm.reserve[gw];
    unsigned long long res_next = res0.x, res_end = res0.y;
    unsigned n_live = 0u, n_started = 0u;

    for (int base = gw * 32; base < pool.n_slots; base += n_warps * 32) {
        const int slot = base + (int)lane;
        const float4 hq = pool.hit[slot];
        const int hop = fbits(hq.y);
        bool has_ray = false;
        bool need_new = hop == WF_EMPTY;
        Ray ray;
        ray.o = ray.d = f3(0.0f, 0.0f, 0.0f);
        ray.time = 0.0f;
        float3 L = f3(0.0f, 0.0f, 0.0f), Tp = f3(1.0f, 1.0f, 1.0f);
        uint32_t pixel = 0u, sample = 0u, depth = 0u;
        int origin = -1;
        uint4 key = make_uint4(0, 0, 0, 0);
        if (hop >= -1) {
            // ---- the slot's segment was traced: ray_color's body (renderer.rs:144-153) ----
            const float4 r0 = pool.ray0[slot], r1 = pool.ray1[slot], s0 = pool.st0[slot], s1 = pool.st1[slot];
            ray.o = f3(r0); ray.time = r0.w;
            ray.d = f3(r1); depth = (uint32_t)fbits(r1.w);
            L = f3(s0); pixel = (uint32_t)fbits(s0.w);
            Tp = f3(s1); sample = (uint32_t)fbits(s1.w);
            key = path_key(prm.seed, pixel, sample);
            bool alive;
            if (hop < 0) {
                L = L + Tp * C.background;                                  // renderer.rs:152-153
                alive = false;
            } else {
                Best best;
                best.t = hq.x; best.op = hop; best.xf = fbits(hq.z);
                HitRec h;
                finalize_hit(S, ray, best, h);
                alive = shade(S, P, ray, h, key, depth, L, Tp);
                origin = h.origin;
                ++depth;
                if ((int)depth >= C.max_depth) alive = false;               // renderer.rs:140-142
            }
            if (alive) {
                has_ray = true;
            } else {
                red_add_f4(prm.sum + pixel, L.x, L.y, L.z, 1.0f);           // avg_color += new_color (renderer.rs:39)
                need_new = true;
            }
        }
        // ---- slots whose path ended (or that never had one) take the next path indices ----
        unsigned pending = __ballot_sync(0xffffffffu, need_new);
        bool got = false;
        unsigned long long path_idx = 0ull;
        while (pending) {
            if (res_next >= res_end) {
                unsigned long long b = 0ull;
                if (lane == 0) b = atomicAdd(prm.path_counter, (unsigned long long)kWfReserve);
                b = __shfl_sync(0xffffffffu, b, 0);
                if (b >= prm.n_paths) break;                                // no paths left, for good
                res_next = b;
                res_end = min(b + (unsigned long long)kWfReserve, prm.n_paths);
            }
            const unsigned long long left = res_end - res_next;
            const unsigned avail = left < 32ull ? (unsigned)left : 32u;
            const bool mine = need_new && !got;
            const unsigned rank = __popc(pending & lt_mask);
            if (mine && rank < avail) { path_idx = res_next + rank; got = true; }
            const unsigned n_pending = __popc(pending);
            res_next += n_pending < avail ? n_pending : avail;
            pending = __ballot_sync(0xffffffffu, need_new && !got);
        }
        if (need_new && got) {
            const unsigned long long sample_rel = path_idx / prm.n_pixels;
            const uint32_t q = (uint32_t)(path_idx - sample_rel * prm.n_pixels);
            int px, py;
            if (prm.tiled) {
                const uint32_t tile = q >> 5, pv = q & 31u;
                const uint32_t ty = tile / (uint32_t)prm.tiles_x, tx = tile - ty * (uint32_t)prm.tiles_x;
                px = (int)(tx * kTileW + (pv & 7u));
                py = (int)(ty * kTileH + (pv >> 3));
            } else {
                py = (int)(q / (uint32_t)C.width);
                px = (int)(q - (uint32_t)py * (uint32_t)C.width);
            }
            pixel = (uint32_t)(py * C.width + px);                          // renderer.rs:32-33
            sample = (uint32_t)(prm.sample_begin + (int64_t)sample_rel);
            key = path_key(prm.seed, pixel, sample);
            ray = camera_ray(C, px, py, key);
            L = f3(0.0f, 0.0f, 0.0f);
            Tp = f3(1.0f, 1.0f, 1.0f);
            depth = 0u;
            origin = -1;
            has_ray = true;
            ++n_started;
        } else if (need_new) {
            pool.hit[slot] = f4i(0.0f, WF_DEAD, -1, -1);
        }
        if (has_ray) {   // world.hit(ray, [0.001, inf)) begins: hoisted media here, the op stream in the extend kernel
            Trav T;
            trav_begin(T, ray, 0, inf);
            media_prepass(S, T, ray.time, tmin, key, depth);
            pool.ray0[slot] = make_float4(ray.o.x, ray.o.y, ray.o.z, ray.time);
            pool.ray1[slot] = make_float4(ray.d.x, ray.d.y, ray.d.z, __int_as_float((int)depth));
            pool.hit[slot] = f4i(T.best.t, T.best.op, -1, origin);
            pool.st0[slot] = make_float4(L.x, L.y, L.z, __int_as_float((int)pixel));
            pool.st1[slot] = make_float4(Tp.x, Tp.y, Tp.z, __int_as_float((int)sample));
            ++n_live;
        }
    }
    if (lane == 0) prm.reserve[gw] = make_ulonglong2(res_next, res_end);
    for (int off = 16; off > 0; off >>= 1) {
        n_live += __shfl_down_sync(0xffffffffu, n_live, off);
        n_started += __shfl_down_sync(0xffffffffu, n_started, off);
    }
    if (lane == 0) {
        if (n_live) atomicAdd(&sh_live, n_live);
        if (n_started) atomicAdd(&sh_started, n_started);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (sh_live) { prm.live_flag[iter] = 1u; atomicAdd(prm.stats + 1, (unsigned long long)sh_live); }
        if (sh_started) atomicAdd(prm.stats + 0, (unsigned long long)sh_started);
    }
}

template <int MIN_BLOCKS>
__global__ void __launch_bounds__(kBlockThreads, MIN_BLOCKS) wf_extend_kernel(const WfParams prm) {
    const DevScene& S = prm.scene;
    const WavePool& pool = prm.pool;
    const float4* __restrict__ ops = S.ops;
    const float tmin = 0.001f;
    const float inf = __int_as_float(0x7f800000);
    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt_mask = (1u << lane) - 1u;
    const unsigned n_slots = (unsigned)pool.n_slots;

    unsigned win_next = 0u, win_end = 0u;          // warp-uniform window of claimed slots
    bool no_more = false;

    Trav T;
    T.i = 0; T.cur_xf = -1; T.best.op = -1; T.best.xf = -1; T.best.t = inf;
    T.o = T.d = T.inv = T.so = T.sd = f3(0.0f, 0.0f, 0.0f);
    int origin = -1, slot = -1;
    float time = 0.0f;
    uint32_t cls = CLS_SHADE;                      // "needs a ray"
    float4 w0 = make_float4(0, 0, 0, 0), w1 = w0;
#define FETCH_NEXT() do { w0 = __ldg(ops + T.i); w1 = __ldg(ops + T.i + 1); } while (0)

    for (;;) {
        unsigned pick;
        if (__popc(__ballot_sync(0xffffffffu, cls == CLS_SHADE)) >= (unsigned)prm.fetch_min) {
            pick = CLS_SHADE;
        } else {
            const unsigned n_slab = __popc(__ballot_sync(0xffffffffu, cls == CLS_SLAB));
            if (n_slab >= (unsigned)prm.slab_fast) {
                pick = CLS_SLAB;
            } else {
                const unsigned tot = __reduce_add_sync(0xffffffffu, cls < CLS_IDLE ? (1u << (6 * cls)) : 0u);
                if (tot == 0u) break;
                const unsigned c_sph = (tot >> 6) & 63u, c_quad = (tot >> 12) & 63u, c_med = (tot >> 18) & 63u;
                unsigned best_n = n_slab;
                pick = CLS_SLAB;
                if (c_sph > best_n) { pick = CLS_SPHERE; best_n = c_sph; }
                if (c_quad > best_n) { pick = CLS_QUAD; best_n = c_quad; }
                if (c_med > best_n) { pick = CLS_MEDIUM; best_n = c_med; }
                if (best_n == 0u) pick = CLS_SHADE;
                if (c_sph >= (unsigned)prm.sphere_min) pick = CLS_SPHERE;
            }
        }

        if (pick == CLS_SLAB) {
#pragma unroll 1
            for (int rep = 0; rep < kSlabReps; ++rep) {
                if (cls == CLS_SLAB) {
                    const uint32_t hdr = (uint32_t)fbits(w0.w);
                    const uint32_t kind = hdr & 15u;
                    if (kind == OP_INNER) {
                        // AABB::hit (aabb.rs:64-84), tight slab form; see slab_interval() for the NaN / sign rules
                        const float ax = (w0.x - T.o.x) * T.inv.x, bx = (w1.x - T.o.x) * T.inv.x;
                        const float ay = (w0.y - T.o.y) * T.inv.y, by = (w1.y - T.o.y) * T.inv.y;
                        const float az = (w0.z - T.o.z) * T.inv.z, bz = (w1.z - T.o.z) * T.inv.z;
                        const bool sx = T.inv.x < 0.0f, sy = T.inv.y < 0.0f, sz = T.inv.z < 0.0f;
                        const float te = fmaxf(fmaxf(fmaxf(sx ? bx : ax, sy ? by : ay), sz ? bz : az), tmin);
                        const float tx = fminf(fminf(fminf(sx ? ax : bx, sy ? ay : by), sz ? az : bz), T.best.t);
                        const bool hit = te <= tx * 1.0000012f;     // te >= tmin > 0, so a negative tx can never pass
                        T.i = hit ? T.i + 2 : fbits(w1.w);
                        cls = (hdr >> (hit ? 8 : 11)) & 7u;
                    } else {     // BOX, XFORM_ENTER / EXIT, INNER_REF (the world ray stays in the pool)
                        cls = op_slab_class(S, T, w0, w1, tmin, origin, [&](float3& wo, float3& wd) {
                            wo = f3(__ldg(pool.ray0 + slot));
                            wd = f3(__ldg(pool.ray1 + slot));
                        });
                    }
                    FETCH_NEXT();
                }
                if (__popc(__ballot_sync(0xffffffffu, cls == CLS_SLAB)) < (unsigned)prm.slab_exit) break;
            }
        } else if (pick == CLS_SPHERE) {
#pragma unroll 1
            for (int rep = 0; rep < prm.sphere_reps; ++rep) {
                if (cls == CLS_SPHERE) {
                    const uint32_t hdr = (uint32_t)fbits(w0.w);
                    op_sphere(S, T, w0, w1, time, tmin, origin);
                    cls = (hdr >> 8) & 7u;
                    FETCH_NEXT();
                }
                if (!__any_sync(0xffffffffu, cls == CLS_SPHERE)) break;
            }
        } else if (pick == CLS_QUAD) {
            if (cls == CLS_QUAD) {
                const uint32_t hdr = (uint32_t)fbits(w0.w);
                op_quad(S, T, w0, w1, tmin, origin);
                cls = (hdr >> 8) & 7u;
                FETCH_NEXT();
            }
        } else if (pick == CLS_MEDIUM) {
            if (cls == CLS_MEDIUM) {     // a medium that could not be hoisted (inside an instance / generic boundary)
                const uint32_t hdr = (uint32_t)fbits(w0.w);
                const uint32_t pixel = (uint32_t)fbits(__ldg(pool.st0 + slot).w), sample = (uint32_t)fbits(__ldg(pool.st1 + slot).w);
                const uint32_t depth = (uint32_t)fbits(__ldg(pool.ray1 + slot).w);
                op_medium(S, T, w0, w1, time, tmin, path_key(prm.seed, pixel, sample), depth);
                cls = (hdr >> 8) & 7u;
                FETCH_NEXT();
            }
        } else {
            // ---- lanes whose traversal ended leave their hit and take the next rays of the pool ----
            const bool waiting = cls == CLS_SHADE;
            if (waiting && slot >= 0) {
                pool.hit[slot] = f4i(T.best.t, T.best.op, T.best.xf, origin);
                slot = -1;
            }
            const unsigned need = __ballot_sync(0xffffffffu, waiting);
            if (need) {
                if (win_next >= win_end && !no_more) {
                    unsigned b = 0u;
                    if (lane == 0) b = atomicAdd(prm.slot_cursor, (unsigned)kWfWindow);
                    b = __shfl_sync(0xffffffffu, b, 0);
                    if (b >= n_slots) {
                        no_more = true;
                    } else {
                        win_next = b;
                        win_end = min(b + (unsigned)kWfWindow, n_slots);
                    }
                }
                if (waiting) {
                    const unsigned idx = win_next + __popc(need & lt_mask);
                    if (idx < win_end) {
                        const float4 hq = pool.hit[idx];
                        if (fbits(hq.y) != WF_DEAD) {
                            const float4 r0 = __ldg(pool.ray0 + idx), r1 = __ldg(pool.ray1 + idx);
                            slot = (int)idx;
                            T.o = f3(r0); time = r0.w;
                            T.d = f3(r1);
                            T.inv = safe_inv(T.d);
                            T.cur_xf = -1;
                            T.i = 0;
                            T.best.t = hq.x; T.best.op = fbits(hq.y); T.best.xf = -1;
                            origin = fbits(hq.w);
                            cls = (uint32_t)prm.first_class;
                            FETCH_NEXT();
                        }
                    } else if (no_more) {
                        cls = CLS_IDLE;
                    }
                }
                win_next = min(win_end, win_next + (unsigned)__popc(need));
            }
        }
    }
#undef FETCH_NEXT
}
