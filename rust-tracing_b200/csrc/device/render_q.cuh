// K1, second form: the render loop (renderer.rs:26-49,139-155) with the paths of a CTA DECOUPLED from its lanes.
//
// render_mk.cuh ties 32 paths to the 32 lanes of a warp for their whole life. A lane whose segment is finished waits
// (idle in every box-test repetition the warp runs meanwhile) until enough lanes of ITS warp wait for shading too:
// profiles/r2_k1_region_breakdown.md - the box-test loop runs at 12.4 of 32 lanes, shading at 8-17.
// A first decoupled form that only took finished segments out of the warps (two queues, traversal / shade) left the loop at
// 12.2 lanes: what a lane waits for most is not shading but its own op class to win the warp's vote
// (profiles/r2_q_queue_forms.md). So here NOTHING waits inside a warp. A CTA owns kQSlots path slots in shared
// memory, twice as many as it has lanes, and one queue per op class (dev_scene.h: box tests, spheres, quads, media) plus
// one for finished segments (SHADE; slots without a path wait there too). A slot is always in exactly one queue or held
// by exactly one lane. Warps take roles by what the queues hold:
//   box-test role   the box-test loop of render_mk.cuh over the 32 slots the lanes hold. A lane whose cursor leaves the
//                   class parks its slot in the queue of the class it arrived at and takes the next waiting slot, together
//                   with the other such lanes of the warp ("exchange"); the warp leaves the role when the queue runs dry;
//   leaf roles      one batch: every lane takes a slot from the queue, runs its sphere / quad / medium test(s), parks it in
//                   the queue of the class that follows;
//   shade role      one batch: hit record, material, new ray, hoisted media, a new path for an ended one -> box-test queue.
//
// Queues are LANE-AFFINE: slot s = row * 32 + lane belongs to lane column s & 31 and is only ever held by that lane (of any
// warp). Per queue and lane one 32-bit mask of waiting rows: push = atomicOr, pop = find a set bit + atomicAnd. No rings,
// no leader election, no warp-wide broadcast; and slot state, laid out field-major, is read and written without bank
// conflicts by construction (the bank of field f of slot s is s & 31 = the lane).
//
// Paths are numbered globally: path p = (sample, tile, pixel in tile), 32 consecutive numbers = one 8x4 tile at one sample
// index; a shade batch claims as many numbers as it has ended paths with one atomicAdd. Results are those of
// render_mk.cuh path for path (same keyed draws, same ops); only the order of the framebuffer additions differs.
//
// Included by rt_cuda.cu after render_mk.cuh (shares its ColdField layout, RenderParams, staging helpers).
#pragma once

#ifdef RT_OPT_QTHREADS
constexpr int kQThreads = RT_OPT_QTHREADS;
#else
constexpr int kQThreads = 512;
#endif
constexpr int kQRows = 32;                   // slots per lane column: one bit each in the column's queue masks
constexpr int kQSlots = kQRows * 32;
constexpr int QF_LINK = kColdFields;         // the cursor of a parked slot / CLS_SHADE / CLS_NEED
constexpr int kQFields = kColdFields + 1;
constexpr int kQQueues = 5;                  // queue index = op class (CLS_SLAB .. CLS_MEDIUM), CLS_SHADE = finished segments

struct QSmem { size_t vec_off, perm_off, state_off, mask_off, total; };
inline QSmem q_smem_layout(size_t ops_bytes, int n_perlin) {   // ops_bytes = 0: the stream stays in global memory
    const int np = n_perlin < kMaxPerlinShared ? n_perlin : kMaxPerlinShared;
    QSmem L;
    size_t off = kSmemOps + ops_bytes;
    L.vec_off = off;   off += (size_t)np * 256 * sizeof(float4);
    L.perm_off = off;  off += (size_t)np * 768;
    L.state_off = off; off += (size_t)kQFields * kQSlots * sizeof(float);
    L.mask_off = off;  off += (size_t)(kQQueues * 32 + 4) * sizeof(uint32_t);   // masks, then [0] = live slots
    L.total = off;
    return L;
}

#define ST(f) st[(f) * kQSlots + slot]
#define ST_U(f) reinterpret_cast<uint32_t*>(st)[(f) * kQSlots + slot]
#define ST_I(f) reinterpret_cast<int*>(st)[(f) * kQSlots + slot]

// One waiting row of this lane's column, or -1. `rot` spreads the choice so no row starves.
__device__ __forceinline__ int q_pop(uint32_t* m, unsigned rot) {
    uint32_t v = *reinterpret_cast<volatile uint32_t*>(m);
    while (v) {
        const uint32_t r = __funnelshift_r(v, v, rot & 31u);
        const uint32_t b = (uint32_t)(__ffs((int)r) - 1 + (int)(rot & 31u)) & 31u;
        const uint32_t old = atomicAnd(m, ~(1u << b));
        if (old & (1u << b)) { __threadfence_block(); return (int)b; }
        v = old & ~(1u << b);
    }
    return -1;
}
__device__ __forceinline__ void q_push(uint32_t* m, int row) {
    __threadfence_block();               // the slot's state before its bit
    atomicOr(m, 1u << row);
}

struct QShared {
    const RenderParams* prm;
    float* st;
    uint32_t* masks;        // [queue * 32 + lane], then live counter at [kQQueues * 32]
    float4* sh_vec;
    uint8_t* sh_perm;
};
__device__ __forceinline__ QShared q_shared() {
    unsigned char* smem = reinterpret_cast<unsigned char*>(dyn_smem);
    QShared q;
    q.prm = reinterpret_cast<const RenderParams*>(smem + kSmemParams);
    const int np = min(q.prm->scene.n_perlin, kMaxPerlinShared);
    q.sh_vec = reinterpret_cast<float4*>(smem + kSmemOps + q.prm->ops_bytes);
    q.sh_perm = reinterpret_cast<uint8_t*>(q.sh_vec + np * 256);
    q.st = reinterpret_cast<float*>(q.sh_perm + np * 768);
    q.masks = reinterpret_cast<uint32_t*>(q.st + kQFields * kQSlots);
    return q;
}

// One batch of the shade role: every lane takes a slot of its column from SHADE (if there is one), finishes its segment
// (renderer.rs:144-153), replaces an ended path by the next global path number, starts the next segment (hoisted media,
// per-ray set-up) and hands the slot to TRAV. Slots for which no path is left leave the game (live count).
template <class Ops>
__device__ __noinline__ unsigned q_shade_batch(unsigned rot) {   // returns (segments started << 16) | paths started, of this lane
    const QShared Q = q_shared();
    const RenderParams& prm = *Q.prm;
    const DevScene& S = prm.scene;
    const DevCamera& C = prm.cam;
    float* st = Q.st;
    const PerlinShared P{Q.sh_vec, Q.sh_perm, min(S.n_perlin, kMaxPerlinShared)};
    Ops ops;
    set_ops_base(ops, S.ops);
    const unsigned lane = threadIdx.x & 31u;
    const float tmin = 0.001f;                     // renderer.rs:144
    const float inf = __int_as_float(0x7f800000);

    const int row = q_pop(Q.masks + CLS_SHADE * 32 + lane, rot);
    const int slot = max(row, 0) * 32 + (int)lane;
    const bool has_slot = row >= 0;
    bool start = false;
    bool want = false;                  // holds a slot and needs a new path
    unsigned counts = 0;
    uint4 key = make_uint4(0, 0, 0, 0);
    uint32_t depth = 0;
    if (has_slot) {
        want = true;
        if ((ST_U(QF_LINK) >> 28) == CLS_SHADE) {
            Ray ray;
            ray.o = f3(ST(F_WO), ST(F_WO + 1), ST(F_WO + 2));
            ray.d = f3(ST(F_WD), ST(F_WD + 1), ST(F_WD + 2));
            ray.time = ST(F_TIME);
            float3 L = f3(ST(F_L), ST(F_L + 1), ST(F_L + 2)), Tp = f3(ST(F_TP), ST(F_TP + 1), ST(F_TP + 2));
            key = path_key(prm.seed, ST_U(F_PIX), ST_U(F_SAMPLE));
            depth = ST_U(F_DEPTH);
            bool alive;
            const int best_op = ST_I(F_BEST_OP), best_xf = ST_I(F_BEST_XF);
            if (best_op < 0) {
                L = L + Tp * C.background;                                  // renderer.rs:152-153
                alive = false;
            } else {
                HitRec h;
                Best b;
                b.t = ST(F_BEST_T); b.op = best_op; b.xf = best_xf;
                finalize_hit(S, ops, ray, b, h);
                alive = shade(S, P, ray, h, key, depth, L, Tp);
                ST_I(F_ORIGIN) = h.origin;
                ++depth;
                if ((int)depth >= C.max_depth) alive = false;               // renderer.rs:140-142
            }
            if (alive) {
                ST(F_WO) = ray.o.x; ST(F_WO + 1) = ray.o.y; ST(F_WO + 2) = ray.o.z;
                ST(F_WD) = ray.d.x; ST(F_WD + 1) = ray.d.y; ST(F_WD + 2) = ray.d.z;
                ST(F_L) = L.x; ST(F_L + 1) = L.y; ST(F_L + 2) = L.z;
                ST(F_TP) = Tp.x; ST(F_TP + 1) = Tp.y; ST(F_TP + 2) = Tp.z;
                ST_U(F_DEPTH) = depth;
                start = true;
                want = false;
            } else {
                red_add_f4(prm.sum + ST_U(F_PIX), L.x, L.y, L.z, 1.0f);     // avg_color += new_color (renderer.rs:39)
            }
        }
    }
    // new paths for the slots whose path ended (or that never had one): path number -> (sample, tile, pixel)
    bool dead = false;
    for (;;) {
        const unsigned need = __ballot_sync(0xffffffffu, want);
        if (!need) break;
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(prm.path_counter, (unsigned long long)__popc(need));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (want) {
            const unsigned long long p = base + (unsigned)__popc(need & ((1u << lane) - 1u));
            if (p >= prm.total_paths) {
                want = false;
                dead = true;
            } else {
                const uint32_t unit = (uint32_t)(p >> 5), pv = (uint32_t)p & 31u;
                const uint32_t n_tiles = (uint32_t)(prm.tiles_x * prm.tiles_y);
                const uint32_t sample_idx = unit / n_tiles, tile = unit - sample_idx * n_tiles;   // sample-major: one sweep over the image per sample
                const uint32_t ty = tile / (uint32_t)prm.tiles_x;
                const int px = (int)(tile - ty * (uint32_t)prm.tiles_x) * kTileW + (int)(pv & 7u), py = (int)ty * kTileH + (int)(pv >> 3);
                if (px < C.width && py < C.height) {
                    const int pix = py * C.width + px;                       // renderer.rs:32-33
                    const uint32_t sample = (uint32_t)(prm.sample_begin + sample_idx);
                    key = path_key(prm.seed, (uint32_t)pix, sample);
                    const Ray ray = camera_ray(C, px, py, key);
                    ST(F_WO) = ray.o.x; ST(F_WO + 1) = ray.o.y; ST(F_WO + 2) = ray.o.z;
                    ST(F_WD) = ray.d.x; ST(F_WD + 1) = ray.d.y; ST(F_WD + 2) = ray.d.z;
                    ST(F_L) = 0.0f; ST(F_L + 1) = 0.0f; ST(F_L + 2) = 0.0f;
                    ST(F_TP) = 1.0f; ST(F_TP + 1) = 1.0f; ST(F_TP + 2) = 1.0f;
                    ST_U(F_SAMPLE) = sample;
                    ST(F_TIME) = ray.time;
                    ST_U(F_PIX) = (uint32_t)pix;
                    ST_U(F_DEPTH) = 0u;
                    ST_I(F_ORIGIN) = -1;
                    depth = 0u;
                    start = true;
                    want = false;
                    counts += 1u;
                }   // else: a pixel beyond the image edge (ragged tile) - ask again
            }
        }
    }
    const unsigned died = __ballot_sync(0xffffffffu, dead);
    if (died && lane == 0) atomicSub(Q.masks + kQQueues * 32, (unsigned)__popc(died));
    if (start) {   // world.hit(ray, [0.001, inf)) begins: hoisted media first, then the op stream from word 0
        const float3 so = f3(ST(F_WO), ST(F_WO + 1), ST(F_WO + 2));
        const float3 sd = f3(ST(F_WD), ST(F_WD + 1), ST(F_WD + 2));
        const float a = dot(sd, sd), inv_a = frcp(a);
        ST(F_O) = so.x; ST(F_O + 1) = so.y; ST(F_O + 2) = so.z;
        ST(F_D) = sd.x; ST(F_D + 1) = sd.y; ST(F_D + 2) = sd.z;
        ST(F_A) = a; ST(F_INVA) = inv_a;
        ST_I(F_XF) = -1;
        Best b;
        b.t = inf; b.op = -1; b.xf = -1;
        media_prepass(S, ops, so, sd, a, inv_a, ST(F_TIME), tmin, key, depth, b);
        ST(F_BEST_T) = b.t; ST_I(F_BEST_OP) = b.op; ST_I(F_BEST_XF) = -1;
        ST_U(QF_LINK) = prm.first_link;
        counts += 1u << 16;
        q_push(Q.masks + (prm.first_link >> 28) * 32 + lane, row);
    }
    return counts;
}

// ConstantMedium::hit for a medium that sits in the stream (render_mk.cuh medium_phase, slot-indexed).
template <class Ops>
__device__ __noinline__ uint32_t q_medium_phase(uint32_t link, int slot) {
    const QShared Q = q_shared();
    const RenderParams& prm = *Q.prm;
    const DevScene& S = prm.scene;
    float* st = Q.st;
    Ops ops;
    set_ops_base(ops, S.ops);
    const uint32_t at = link & kLinkMask;
    const float4 w0 = ops(at), w1 = ops(at + 16u);
    const uint4 key = path_key(prm.seed, ST_U(F_PIX), ST_U(F_SAMPLE));
    float t;
    int next_word;
    if (medium_test(S, ops, at, w0, w1, f3(ST(F_O), ST(F_O + 1), ST(F_O + 2)), f3(ST(F_D), ST(F_D + 1), ST(F_D + 2)),
                    ST(F_A), ST(F_INVA), ST(F_TIME), 0.001f, ST(F_BEST_T), key, ST_U(F_DEPTH), &t, &next_word)) {
        ST(F_BEST_T) = t; ST_I(F_BEST_OP) = (int)(at >> 4); ST_I(F_BEST_XF) = ST_I(F_XF);
    }
    return ((uint32_t)next_word << 4) | ((uint32_t)fbits(w0.w) & 0xf0000000u);
}

// The op stream in shared memory through an explicit 32-bit shared address computed once per function: inside the
// out-of-line role functions the compiler re-derived the shared window base (S2UR SR_CgaCtaId + ULEA) in every repetition
// of the box-test loop - its most sampled instruction (profiles/r2_q_queue_forms.md).
struct OpsSharedAddr {
    uint32_t base;
    __device__ __forceinline__ float4 operator()(uint32_t byte_off) const {
        float4 v;
        asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(base + byte_off));
        return v;
    }
};
__device__ __forceinline__ void set_ops_base(OpsSharedAddr& o, const float4*) { o.base = (uint32_t)__cvta_generic_to_shared(dyn_smem) + kSmemOps; }

// The cursor, the closest hit so far: what a slot needs besides its ray to be picked up by another lane.
#define PARK() do { ST(F_BEST_T) = best_t; ST_I(F_BEST_OP) = best_op; ST_I(F_BEST_XF) = best_xf; ST_U(QF_LINK) = link; } while (0)
#define FETCH_NEXT() do { const uint32_t at__ = link & kLinkMask; w0 = ops(at__); w1 = ops(at__ + 16u); } while (0)
#define CUR_O() f3(ST(F_O), ST(F_O + 1), ST(F_O + 2))
#define CUR_D() f3(ST(F_D), ST(F_D + 1), ST(F_D + 2))

// lanes of this warp whose column has a slot waiting in queue q
__device__ __forceinline__ unsigned q_avail(const uint32_t* masks, int q, unsigned lane) {
    return __popc(__ballot_sync(0xffffffffu, *reinterpret_cast<const volatile uint32_t*>(masks + q * 32 + lane) != 0u));
}

// One batch of a leaf role (CLS = CLS_SPHERE or CLS_QUAD): Sphere::hit / Quad::hit for up to `reps` consecutive ops of the class.
template <uint32_t CLS, class Ops>
__device__ __noinline__ void q_leaf_batch(unsigned rot, int reps) {
    const QShared Q = q_shared();
    const DevScene& S = Q.prm->scene;
    float* st = Q.st;
    Ops ops;
    set_ops_base(ops, S.ops);
    const unsigned lane = threadIdx.x & 31u;
    const float tmin = 0.001f;
    const int row = q_pop(Q.masks + CLS * 32 + lane, rot);
    if (row < 0) return;
    const int slot = row * 32 + (int)lane;
    uint32_t link = ST_U(QF_LINK);
    float best_t = ST(F_BEST_T);
    int best_op = ST_I(F_BEST_OP), best_xf = ST_I(F_BEST_XF);
    const float3 o = CUR_O(), d = CUR_D();
    const int origin = ST_I(F_ORIGIN);
#pragma unroll 1
    for (int rep = 0; rep < reps && (link >> 28) == CLS; ++rep) {
        const uint32_t at = link & kLinkMask;
        const float4 w0 = ops(at), w1 = ops(at + 16u);
        const uint32_t hdr = (uint32_t)fbits(w0.w);
        float t;
        bool win;
        if (CLS == CLS_SPHERE) win = sphere_test(S, ops, link, w0, w1, o, d, ST(F_A), ST(F_INVA), ST(F_TIME), tmin, best_t, origin, &t);
        else win = quad_test(ops, link, w0, w1, o, d, tmin, best_t, origin, &t);
        if (win) { best_t = t; best_op = (int)(at >> 4); best_xf = ST_I(F_XF); }
        link = at + (hdr & kHdrFallThrough);
    }
    PARK();
    q_push(Q.masks + (link >> 28) * 32 + lane, row);
}

// One batch of the medium role: ConstantMedium::hit for media that sit in the stream.
template <class Ops>
__device__ __noinline__ void q_medium_batch(unsigned rot) {
    const QShared Q = q_shared();
    float* st = Q.st;
    const unsigned lane = threadIdx.x & 31u;
    const int row = q_pop(Q.masks + CLS_MEDIUM * 32 + lane, rot);
    if (row < 0) return;
    const int slot = row * 32 + (int)lane;
    const uint32_t link = q_medium_phase<Ops>(ST_U(QF_LINK), slot);
    ST_U(QF_LINK) = link;
    q_push(Q.masks + (link >> 28) * 32 + lane, row);
}

// The box-test role: cull boxes, cube primitives, instance enter / exit (render_mk.cuh's loop) over the slots the lanes hold,
// exchanging slots whose cursor left the class for waiting ones. Returns when the queue runs dry.
template <class Ops>
__device__ __noinline__ void q_slab_worker(unsigned rot) {
    const QShared Q = q_shared();
    const RenderParams& prm = *Q.prm;
    float* st = Q.st;
    uint32_t* const masks = Q.masks;
    Ops ops;
    set_ops_base(ops, prm.scene.ops);
    const unsigned lane = threadIdx.x & 31u;
    const float tmin = 0.001f;                     // renderer.rs:144
    const int slab_drop = prm.slab_drop, min_slab = prm.min_trav;
    uint32_t* const my_slab = masks + CLS_SLAB * 32 + lane;
    int row = -1, slot = (int)lane;
    uint32_t link = CLS_NEED << 28;                // holds no slot
    float3 inv = f3(0.0f, 0.0f, 0.0f), oi = inv;
    float eps = 0.0f, best_t = 0.0f;
    int best_op = -1, best_xf = -1;
    float4 w0 = make_float4(0, 0, 0, 0), w1 = w0;
#define SET_RAY(o_, d_) do {                                                                       \
        const RaySetup R__ = ray_setup(o_, d_);                                                    \
        inv = R__.inv; oi = R__.oi; eps = R__.eps;                                                 \
        ST(F_O) = (o_).x; ST(F_O + 1) = (o_).y; ST(F_O + 2) = (o_).z;                              \
        ST(F_D) = (d_).x; ST(F_D + 1) = (d_).y; ST(F_D + 2) = (d_).z;                              \
        ST(F_A) = R__.a; ST(F_INVA) = R__.inv_a;                                                   \
    } while (0)
    for (;;) {
        // ---- exchange: cursors that left the class park in the queue of the class they arrived at; free lanes take a slot
        if (row >= 0 && link >= kSlabLimit) {
            PARK();
            q_push(masks + (link >> 28) * 32 + lane, row);
            row = -1;
        }
        if (row < 0) {
            row = q_pop(my_slab, rot);
            if (row >= 0) {
                slot = row * 32 + (int)lane;
                link = ST_U(QF_LINK);
                best_t = ST(F_BEST_T); best_op = ST_I(F_BEST_OP); best_xf = ST_I(F_BEST_XF);
                const RaySetup R = ray_setup(CUR_O(), CUR_D());
                inv = R.inv; oi = R.oi; eps = R.eps;
                FETCH_NEXT();
            } else {
                link = CLS_NEED << 28;
            }
        }
        ++rot;
        const unsigned n_act = __popc(__ballot_sync(0xffffffffu, row >= 0));
        if (n_act < (unsigned)min_slab) {
            // the queue runs dry for this warp: leave if another queue would fill more lanes
            unsigned other = 0;
#pragma unroll
            for (int q = 1; q < kQQueues; ++q) other = max(other, q_avail(masks, q, lane));
            if (n_act == 0u || other > n_act) {
                if (row >= 0) { PARK(); q_push(my_slab, row); }
                return;
            }
        }
        const unsigned stay = (unsigned)max(1, (int)n_act - slab_drop);
        unsigned n_slab;
#pragma unroll 1
        do {
            if (link < kSlabLimit) {
                float te, tx;
                slab_ch(w0, w1, inv, oi, &te, &tx);
                const uint32_t hdr = (uint32_t)fbits(w0.w);
                if ((hdr & kHdrNotInner) == 0u) {          // OP_INNER: AABB::hit (aabb.rs:64-84)
                    link = cull_pass(te, tx, tmin, best_t, eps) ? link + hdr : (uint32_t)fbits(w1.w);
                } else {
                    const uint32_t kind = (hdr >> 8) & 15u;
                    const uint32_t ft = link + (hdr & kHdrFallThrough);
                    if (kind == OP_BOX) {                  // Quad::cube as one slab primitive (box_accept)
                        float t;
                        bool win;
                        const int origin = ST_I(F_ORIGIN);
                        if (!starts_on(origin, link)) {
                            win = box_accept(te, tx, tmin, best_t, &t);
                        } else {                           // the ray starts on a face of this very box (render_mk.cuh)
                            const int face = origin & 7;   // 0 +z, 1 +x, 2 -z, 3 -x, 4 +y, 5 -y
                            const float ia = (face == 1 || face == 3) ? inv.x : (face >= 4 ? inv.y : inv.z);
                            const bool max_side = face == 0 || face == 1 || face == 4;
                            win = (ia > 0.0f) != max_side &&
                                  box_test_from_face(ops(link + 32u), ops(link + 48u), CUR_O(), inv, tmin, best_t, face, &t);
                        }
                        if (win) { best_t = t; best_op = (int)(link >> 4); best_xf = ST_I(F_XF); }
                        link = ft;
                    } else if (kind == OP_XFORM_ENTER) {   // Translate::hit / RotateY::hit (hittable.rs:96-111,159-193)
                        if (!cull_pass(te, tx, tmin, best_t, eps)) {
                            link = (uint32_t)fbits(w1.w);
                        } else {
                            const float4 w2 = ops(link + 32u), w3 = ops(link + 48u);
                            const float3 lo_ = xform_point(f3(ST(F_WO), ST(F_WO + 1), ST(F_WO + 2)), w2, w3);
                            const float3 ld_ = xform_dir(f3(ST(F_WD), ST(F_WD + 1), ST(F_WD + 2)), w2, w3);
                            SET_RAY(lo_, ld_);             // the op holds the composed world -> local transform
                            ST_I(F_XF) = (int)(link >> 4);
                            link = ft;
                        }
                    } else if (kind == OP_XFORM_EXIT) {    // back in the enclosing space
                        const int parent = fbits(w0.x);
                        float3 lo_ = f3(ST(F_WO), ST(F_WO + 1), ST(F_WO + 2));
                        float3 ld_ = f3(ST(F_WD), ST(F_WD + 1), ST(F_WD + 2));
                        if (parent >= 0) {
                            const float4 p2 = ops(((uint32_t)parent << 4) + 32u), p3 = ops(((uint32_t)parent << 4) + 48u);
                            lo_ = xform_point(lo_, p2, p3);
                            ld_ = xform_dir(ld_, p2, p3);
                        }
                        SET_RAY(lo_, ld_);
                        ST_I(F_XF) = parent;
                        link = ft;
                    } else {                               // OP_INNER_REF: the reference's box, the reference's test
                        link = aabb_hit_reference(w0, w1, CUR_O(), inv, tmin, best_t) ? ft : (uint32_t)fbits(w1.w);
                    }
                }
                FETCH_NEXT();
            }
            n_slab = __popc(__ballot_sync(0xffffffffu, link < kSlabLimit));
        } while (n_slab >= stay);
    }
#undef SET_RAY
}

template <bool OPS_SMEM>
__global__ void __launch_bounds__(kQThreads, 1) render_kernel_q(const RenderParams prm_in) {
    unsigned char* smem = reinterpret_cast<unsigned char*>(dyn_smem);
    const uint32_t smem_addr = (uint32_t)__cvta_generic_to_shared(smem);
    const int tid = threadIdx.x;
    const unsigned lane = (unsigned)tid & 31u;
    typedef typename std::conditional<OPS_SMEM, OpsSharedAddr, OpsGlobal>::type Ops;

    // ---- stage: the launch parameters, the op stream (one bulk copy per CTA, completion on an mbarrier), the Perlin tables
    for (int k = tid; k < (int)(sizeof(RenderParams) / 4); k += kQThreads)
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
    float* st = reinterpret_cast<float*>(sh_perm + np * 768);
    uint32_t* masks = reinterpret_cast<uint32_t*>(st + kQFields * kQSlots);
    // every slot starts without a path, waiting in SHADE for its first one
    for (int k = tid; k < kQSlots; k += kQThreads) reinterpret_cast<uint32_t*>(st)[QF_LINK * kQSlots + k] = CLS_NEED << 28;
    for (int k = tid; k < kQQueues * 32; k += kQThreads) masks[k] = (k >> 5) == (int)CLS_SHADE ? 0xffffffffu : 0u;
    if (tid == 0) masks[kQQueues * 32] = (uint32_t)kQSlots;
    stage_perlin(prm_in.scene, sh_vec, sh_perm);      // ends with __syncthreads()
    if (OPS_SMEM) {
        uint32_t done;
        do {
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                         : "=r"(done) : "r"(smem_addr), "r"(0) : "memory");
        } while (!done);
    }
    const int sphere_reps = prm_in.sphere_reps, quad_reps = prm_in.quad_reps;
    volatile uint32_t* const live = masks + kQQueues * 32;
    unsigned rot = (unsigned)(tid >> 5) * 5u + blockIdx.x;
    unsigned n_paths = 0, n_segs = 0;

    for (;;) {
        // ---- the role: the queue that would fill most lanes of this warp
        unsigned best_n = 0;
        int role = -1;
#pragma unroll
        for (int q = 0; q < kQQueues; ++q) {
            const unsigned n = q_avail(masks, q, lane);
            if (n >= best_n && n > 0u) { best_n = n; role = q; }   // ties: the later class (leaf and shade batches feed the box-test queue)
        }
        rot += 7u;
        if (role == (int)CLS_SLAB) q_slab_worker<Ops>(rot);
        else if (role == (int)CLS_SHADE) { const unsigned c = q_shade_batch<Ops>(rot); n_paths += c & 0xffffu; n_segs += c >> 16; }
        else if (role == (int)CLS_SPHERE) q_leaf_batch<CLS_SPHERE, Ops>(rot, sphere_reps);
        else if (role == (int)CLS_QUAD) q_leaf_batch<CLS_QUAD, Ops>(rot, quad_reps);
        else if (role == (int)CLS_MEDIUM) q_medium_batch<Ops>(rot);
        else {
            if (*live == 0u) break;
            __nanosleep(100);
        }
    }
    for (int off = 16; off > 0; off >>= 1) {
        n_paths += __shfl_down_sync(0xffffffffu, n_paths, off);
        n_segs += __shfl_down_sync(0xffffffffu, n_segs, off);
    }
    if (lane == 0) {
        atomicAdd(prm_in.stats + K_PATHS, (unsigned long long)n_paths);
        atomicAdd(prm_in.stats + K_SEGMENTS, (unsigned long long)n_segs);
    }
}
#undef PARK
#undef FETCH_NEXT
#undef CUR_O
#undef CUR_D
#undef ST
#undef ST_U
#undef ST_I
