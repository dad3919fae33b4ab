// Device code of the ray_color hot path (renderer.rs:26-49,139-155 and everything it calls),
// written for sm_100a. f32 arithmetic except where a primitive is flagged FLAG_PRECISE.
//
//   sphere_test / quad_test / box_accept / medium_test / slab_ch
//                     Sphere / Quad / Quad::cube / ConstantMedium / AABB (sphere.rs:59-89, quad.rs:97-133,
//                     constant_medium.rs:34-70, aabb.rs:64-84), shared by every kernel
//   traverse<>()      world.hit(ray, ray_t) as one stackless loop over the threaded op stream (dev_scene.h;
//                     bvh.rs:90-113, hittable.rs:61-188) - the parity kernels' form; the render kernel (render_mk.cuh)
//                     walks the same stream with the same tests, regrouping its lanes by op class
//   finalize_hit()    HitRecord::new + uv (hittable.rs:22-37, sphere.rs:48-52) for the winner only
//   texture_value()   texture.rs:32-111 + perlin.rs:27-100 (tables in shared memory)
//   shade()           Material::emitted / scatter (material.rs:26-138)
//   camera_ray()      Camera::get_ray (camera.rs:112-137)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../../include/rt_b200.h"
#include "dev_scene.h"

namespace rtdev {

struct DevImage {
    const float4* texels;  // pre-linearised (byte/255)^2.2 (color.rs:21-27), row 0 = top; one LDG.128 per lookup
    int width, height;
};

struct DevScene {
    const float4* ops;
    int n_words;              // end of the world program (hoisted media bodies live beyond it)
    int n_media;              // hoisted media, evaluated at the start of every segment
    int media_op[kMaxHoistedMedia];
    const float4* mats;
    const float4* texs;
    const float4* perlin_vec;
    const uint8_t* perlin_perm;
    int n_perlin;
    const double4* precise;
    const DevImage* images;
};

struct DevCamera {
    int width, height;
    int max_depth;
    float3 background;
    float3 center;
    float3 rel00;      // pixel00_loc - center (computed in f64 on the host: small, so f32 keeps sub-pixel accuracy)
    float3 du, dv;
    float3 disk_u, disk_v;
    int defocus;
};

// ------------------------------------------------------------------ float3 helpers
__device__ __forceinline__ float3 f3(float x, float y, float z) { return make_float3(x, y, z); }
__device__ __forceinline__ float3 f3(float4 v) { return make_float3(v.x, v.y, v.z); }
__device__ __forceinline__ float3 operator+(float3 a, float3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ float3 operator-(float3 a, float3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ float3 operator-(float3 a) { return f3(-a.x, -a.y, -a.z); }
__device__ __forceinline__ float3 operator*(float3 a, float3 b) { return f3(a.x * b.x, a.y * b.y, a.z * b.z); }
__device__ __forceinline__ float3 operator*(float3 a, float s) { return f3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ float3 operator*(float s, float3 a) { return a * s; }
__device__ __forceinline__ float dot(float3 a, float3 b) { return fmaf(a.x, b.x, fmaf(a.y, b.y, a.z * b.z)); }
__device__ __forceinline__ float3 fma3(float s, float3 a, float3 b) { return f3(fmaf(s, a.x, b.x), fmaf(s, a.y, b.y), fmaf(s, a.z, b.z)); }

// Division, reciprocal and square root as ONE special-function instruction (+ one multiply for the division): the
// `.approx.ftz` forms, 1-2 ulp like the `-prec-div=false -prec-sqrt=false` sequences the compiler emits for `/` and sqrtf,
// but without their range fix-ups (operands beyond 2^126 or below 2^-126: 7-9 instructions per site). Geometry of this
// path never gets there - a denormal direction component counts as zero, i.e. a ray parallel to that slab - and with ~35
// executed division sites the fix-ups alone were 12% of the render kernel's active instruction footprint, which sits at the
// edge of the 32 KB instruction cache (profiles/r2_k1_icache.md).
#ifndef RT_OPT_NO_FASTDIV
__device__ __forceinline__ float frcp(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float fdiv(float a, float b) { float r; asm("div.approx.ftz.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float fsqrt(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float frsqrt(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
// sin and cos of 2 pi u, u in [0, 1): the special-function unit on the argument folded to [-pi, pi) (absolute error
// 2^-20.9 there), sign restored: sin(2 pi u) = -sin(2 pi u - pi)
__device__ __forceinline__ void fsincos_turn(float u, float* s, float* c) {
    const float a = fmaf(u, 6.283185307179586f, -3.141592653589793f);
    float sn, cs;
    asm("sin.approx.ftz.f32 %0, %1;" : "=f"(sn) : "f"(a));
    asm("cos.approx.ftz.f32 %0, %1;" : "=f"(cs) : "f"(a));
    *s = -sn;
    *c = -cs;
}
// sin(x) for |x| up to a few hundred (NoiseTexture: scale * p.z + 10 * turbulence): argument folded to [-pi, pi] with a
// two-term 2 pi (error ~1e-7 |x| / 2 pi), then the special-function unit (2^-20.9 there) - instead of sinf's range
// reduction with its slow path through local memory.
#ifndef RT_OPT_LIBM_SIN_LOG
__device__ __forceinline__ float fsin(float x) {
    const float k = rintf(x * 0.15915494309189535f);
    const float r = fmaf(k, 1.7484555e-7f, fmaf(k, -6.2831854820251465f, x));    // 2 pi = 6.2831854820251465 - 1.7484555e-7
    float s;
    asm("sin.approx.ftz.f32 %0, %1;" : "=f"(s) : "f"(r));
    return s;
}
#ifdef RT_OPT_APPROX_LOG
__device__ __forceinline__ float flog(float x) { float r; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r * 0.6931471805599453f; }
#else
// ln(u) of a medium's free-flight draw stays logf: lg2.approx is good to 2^-22 ABSOLUTE near 1, and the fog of final_scene
// multiplies it by 1 / density = 10^4 - four times the stated bound on a scatter point's t (measured: +1% Mpaths/s, dropped)
__device__ __forceinline__ float flog(float x) { return logf(x); }
#endif
#else
__device__ __forceinline__ float fsin(float x) { return sinf(x); }
__device__ __forceinline__ float flog(float x) { return logf(x); }
#endif
#else
__device__ __forceinline__ float fsin(float x) { return sinf(x); }
__device__ __forceinline__ float flog(float x) { return logf(x); }
__device__ __forceinline__ float frsqrt(float x) { return rsqrtf(x); }
__device__ __forceinline__ void fsincos_turn(float u, float* s, float* c) { sincospif(2.0f * u, s, c); }
__device__ __forceinline__ float frcp(float x) { return 1.0f / x; }
__device__ __forceinline__ float fdiv(float a, float b) { return a / b; }
__device__ __forceinline__ float fsqrt(float x) { return sqrtf(x); }
#endif
__device__ __forceinline__ int fbits(float f) { return __float_as_int(f); }

__device__ __forceinline__ float3 normalize3(float3 a) { return a * frsqrt(dot(a, a)); }

// ------------------------------------------------------------------ keyed RNG (shared spec with the oracle)
// pcg4d (Jarzynski & Olano, JCGT 9(3) 2020). path key = pcg4d(pixel, sample, seed_lo, seed_hi);
// draw(purpose) = pcg4d(key.x, key.y, key.z + segment, key.w + purpose); u01 = (x >> 8) * 2^-24.
#ifdef RT_OPT_PCG_NOINLINE
__device__ __noinline__ uint4 pcg4d(uint4 v) {
#else
__device__ __forceinline__ uint4 pcg4d(uint4 v) {
#endif
    v.x = v.x * 1664525u + 1013904223u; v.y = v.y * 1664525u + 1013904223u;
    v.z = v.z * 1664525u + 1013904223u; v.w = v.w * 1664525u + 1013904223u;
    v.x += v.y * v.w; v.y += v.z * v.x; v.z += v.x * v.y; v.w += v.y * v.z;
    v.x ^= v.x >> 16; v.y ^= v.y >> 16; v.z ^= v.z >> 16; v.w ^= v.w >> 16;
    v.x += v.y * v.w; v.y += v.z * v.x; v.z += v.x * v.y; v.w += v.y * v.z;
    return v;
}
constexpr uint32_t P_CAMERA = 0, P_CAMERA_DISK = 1, P_SCATTER = 2, P_MEDIUM = 16;
__device__ __forceinline__ uint4 path_key(uint64_t seed, uint32_t pixel, uint32_t sample) {
    return pcg4d(make_uint4(pixel, sample, (uint32_t)seed, (uint32_t)(seed >> 32)));
}
__device__ __forceinline__ uint4 draw(uint4 key, uint32_t seg, uint32_t purpose) {
    return pcg4d(make_uint4(key.x, key.y, key.z + seg, key.w + purpose));
}
__device__ __forceinline__ float u01(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }

// loop-free samplers with the distributions of vec3.rs:54-65 (rejection loops in the reference)
__device__ __forceinline__ float3 unit_vector(float u0, float u1) {
    const float z = 1.0f - 2.0f * u0;
    const float r = fsqrt(fmaxf(0.0f, 1.0f - z * z));
    float s, c;
    fsincos_turn(u1, &s, &c);
    return f3(r * c, r * s, z);
}

// ------------------------------------------------------------------ traversal
struct Ray {
    float3 o, d;
    float time;
};

// Features of a scene the render kernel is specialised on (render_mk.cuh): code a scene cannot reach is not compiled into
// the instantiation that renders it. The parity kernels and the generic instantiation carry everything.
constexpr unsigned FEAT_FOLD = 1u;      // cube primitives or instances in the stream: FOLD form of the box-test loop
constexpr unsigned FEAT_PRECISE = 2u;   // an f64 sphere (radius > 200) as a primitive or as a medium boundary
constexpr unsigned FEAT_RARE = 4u;      // media in the stream / boundary programs / OP_INNER_REF
constexpr unsigned FEAT_XBOX = 8u;      // a medium bounded by a (rotated, translated) cube
constexpr unsigned FEAT_DEFOCUS = 16u;  // the camera has a defocus disk (a property of the launch, not of the scene)
constexpr unsigned FEAT_ALL = 31u;

struct Best {
    float t;
    int op;   // word index of the winning primitive / medium op, -1 = none
    int xf;   // word index of the enclosing OP_XFORM_ENTER, -1 = world space
};

// Where the op stream is read from. Offsets are BYTE offsets (the low 28 bits of a link, dev_scene.h).
struct OpsGlobal {   // parity kernels, and render launches whose stream does not fit in shared memory
    const float4* base;
    __device__ __forceinline__ float4 operator()(uint32_t byte_off) const {
        return __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const char*>(base) + byte_off));
    }
};
// The render kernel: the whole stream staged in shared memory by one bulk (TMA) copy per CTA, kSmemOps bytes into the
// dynamic shared window (render_mk.cuh: [0, 8) the copy's mbarrier, [16, 528) the launch parameters). Addressed off the
// symbol so a fetch is one LDS.128 with an immediate offset.
extern __shared__ float4 dyn_smem[];
constexpr uint32_t kSmemParams = 16;
constexpr uint32_t kSmemOps = 16 + 512;
struct OpsShared {
    __device__ __forceinline__ float4 operator()(uint32_t byte_off) const {
        return *reinterpret_cast<const float4*>(reinterpret_cast<const unsigned char*>(dyn_smem) + kSmemOps + byte_off);
    }
};

// Per-ray constants of the centre / half-extent slab test (dev_scene.h, CULL BOXES).
struct RaySetup {
    float3 inv, oi;         // 1/d, -o/d
    float eps;              // 2^-21 max_k |o_k / d_k| over the axes with a finite quotient
    float a, inv_a;         // |d|^2 and its reciprocal (sphere tests)
};
__device__ __forceinline__ float3 safe_inv(float3 d) { return f3(frcp(d.x), frcp(d.y), frcp(d.z)); }
__device__ __forceinline__ RaySetup ray_setup(float3 o, float3 d) {
    RaySetup R;
    R.inv = safe_inv(d);
    R.oi = f3(-o.x * R.inv.x, -o.y * R.inv.y, -o.z * R.inv.z);
    const float inf = __int_as_float(0x7f800000);
    const float ex = fabsf(R.oi.x), ey = fabsf(R.oi.y), ez = fabsf(R.oi.z);   // NaN (0 * inf) and inf axes constrain nothing
    R.eps = 4.76837158e-7f * fmaxf(fmaxf(ex < inf ? ex : 0.0f, ey < inf ? ey : 0.0f), ez < inf ? ez : 0.0f);
    R.a = dot(d, d);
    R.inv_a = frcp(R.a);
    return R;
}

// origin code of a ray that starts on a surface: (byte offset of the op) | box face; -1 = none. Byte offsets are
// multiples of 16, so the face lives in the low nibble and "does this ray start on the op at `link`" is one masked xor.
__device__ __forceinline__ int origin_code(int op_word, int face) { return (op_word << 4) | face; }
__device__ __forceinline__ bool starts_on(int origin, uint32_t link) { return (((uint32_t)origin ^ link) & 0x0ffffff0u) == 0u; }

// Entry / exit parameters of the ray against the cull box {c = w0.xyz, h = w1.xyz}, not clamped to any interval.
// fminf/fmaxf drop a NaN operand like f64::min/max do (aabb.rs:76-77): an axis the ray is parallel to (inv = inf,
// tc = inf - inf) constrains nothing here - the classic form would still test the origin against that slab; boxes only
// cull, so passing a few more is allowed.
__device__ __forceinline__ void slab_ch(float4 w0, float4 w1, const float3& inv, const float3& oi, float* t_enter, float* t_exit) {
    const float tcx = fmaf(w0.x, inv.x, oi.x), tcy = fmaf(w0.y, inv.y, oi.y), tcz = fmaf(w0.z, inv.z, oi.z);
    const float thx = fabsf(w1.x * inv.x), thy = fabsf(w1.y * inv.y), thz = fabsf(w1.z * inv.z);   // |.| is an operand modifier
    *t_enter = fmaxf(fmaxf(tcx - thx, tcy - thy), tcz - thz);
    *t_exit = fminf(fminf(tcx + thx, tcy + thy), tcz + thz);
}
// AABB::hit (aabb.rs:64-84) for a padded cull box: tight slab test with the reciprocal hoisted per ray (permitted
// substitution, SURVEY.md §8(a)-Q: it only culls more, it never changes which hits exist).
constexpr float kSlabSlack = 1.0000012f;
__device__ __forceinline__ bool cull_pass(float te, float tx, float tmin, float tmax, float eps) {
    return fmaxf(te, tmin) <= fmaf(fminf(tx, tmax), kSlabSlack, eps);
}

// Entry / exit against exact corners {lo = w0.xyz, hi = w1.xyz} in the classic form (no cancellation): box primitives
// whose ray starts on one of their faces, medium boundaries, hit records.
__device__ __forceinline__ void slab_interval(float4 w0, float4 w1, float3 o, float3 inv, float* t_enter, float* t_exit) {
    const float ax = (w0.x - o.x) * inv.x, bx = (w1.x - o.x) * inv.x;
    const float ay = (w0.y - o.y) * inv.y, by = (w1.y - o.y) * inv.y;
    const float az = (w0.z - o.z) * inv.z, bz = (w1.z - o.z) * inv.z;
    const bool sx = inv.x < 0.0f, sy = inv.y < 0.0f, sz = inv.z < 0.0f;
    *t_enter = fmaxf(fmaxf(sx ? bx : ax, sy ? by : ay), sz ? bz : az);
    *t_exit = fminf(fminf(sx ? ax : bx, sy ? ay : by), sz ? az : bz);
}

// AABB::hit exactly as aabb.rs:64-84 states it: every axis is tested against the ORIGINAL interval, the interval is
// never narrowed between axes. Used only for OP_INNER_REF nodes (dev_scene.h), where the outcome of this very test
// decides which part of a quad the reference can see.
__device__ __forceinline__ bool aabb_hit_reference(float4 w0, float4 w1, float3 o, float3 inv, float tmin, float tmax) {
    const float ax = (w0.x - o.x) * inv.x, bx = (w1.x - o.x) * inv.x;
    const float ay = (w0.y - o.y) * inv.y, by = (w1.y - o.y) * inv.y;
    const float az = (w0.z - o.z) * inv.z, bz = (w1.z - o.z) * inv.z;
    const bool sx = inv.x < 0.0f, sy = inv.y < 0.0f, sz = inv.z < 0.0f;
    const bool mx = fminf(sx ? ax : bx, tmax) <= fmaxf(sx ? bx : ax, tmin);   // t_max <= t_min -> miss (aabb.rs:79-81)
    const bool my = fminf(sy ? ay : by, tmax) <= fmaxf(sy ? by : ay, tmin);
    const bool mz = fminf(sz ? az : bz, tmax) <= fmaxf(sz ? bz : az, tmin);
    return !(mx || my || mz);
}

// local = R(x - a) + b with R = rotate-y as in hittable.rs:164-168.
// Written with explicit round-to-nearest intrinsics, which the compiler never contracts or re-associates: the traversal
// (entering an instance) and finalize_hit() (recomputing the local ray of the winner) then produce the SAME bits, so a
// plane parameter computed at the two sites cannot differ by a contraction.
__device__ __forceinline__ float3 xform_point(float3 x, float4 w2, float4 w3) {
    const float qx = __fsub_rn(x.x, w2.x), qy = __fsub_rn(x.y, w2.y), qz = __fsub_rn(x.z, w2.z);
    const float s = w2.w, c = w3.w;
    return f3(__fadd_rn(__fmaf_rn(c, qx, -__fmul_rn(s, qz)), w3.x), __fadd_rn(qy, w3.y),
              __fadd_rn(__fmaf_rn(s, qx, __fmul_rn(c, qz)), w3.z));
}
__device__ __forceinline__ float3 xform_dir(float3 v, float4 w2, float4 w3) {
    const float s = w2.w, c = w3.w;
    return f3(__fmaf_rn(c, v.x, -__fmul_rn(s, v.z)), v.y, __fmaf_rn(s, v.x, __fmul_rn(c, v.z)));
}
// inverse rotation (hittable.rs:173-179)
__device__ __forceinline__ float3 xform_dir_back(float3 v, float4 w2, float4 w3) {
    const float s = w2.w, c = w3.w;
    return f3(__fmaf_rn(c, v.x, __fmul_rn(s, v.z)), v.y, __fmaf_rn(-s, v.x, __fmul_rn(c, v.z)));
}

// Sphere::hit roots (sphere.rs:59-83): false on a negative discriminant, else near/far roots.
// `self_origin`: the ray starts on this very sphere, so the root that is analytically 0 is dropped (NaN)
// (in f64 the reference rejects it through ray_t.min = 0.001; in f32 its rounding noise can exceed that).
// a = |d|^2 and inv_a come from the ray set-up.
__device__ __forceinline__ bool sphere_roots_f32(float3 oc, float3 d, float a, float inv_a, float r, bool self_origin,
                                                 float* r1, float* r2) {
    const float hb = dot(oc, d);
    // discriminant/a = r^2 - |oc - (hb/a) d|^2 : no cancellation between hb^2 and a*c for distant origins
    const float3 l = fma3(-hb * inv_a, d, oc);
    const float disc = fmaf(r, r, -dot(l, l));
    if (disc < 0.0f) return false;
    const float sq = fsqrt(disc * a);
    const float cc = fmaf(-r, r, dot(oc, oc));
    const float nan = __int_as_float(0x7fc00000);
    float near_root, far_root;
    if (hb > 0.0f) {            // both roots via q to avoid -hb + sq cancellation
        const float q = -hb - sq;
        near_root = q * inv_a;
        far_root = self_origin ? nan : fdiv(cc, q);
    } else {
        const float q = -hb + sq;
        far_root = q * inv_a;
        near_root = self_origin ? nan : fdiv(cc, q);
    }
    *r1 = near_root;
    *r2 = far_root;
    return true;
}
__device__ __noinline__ bool sphere_roots_f64(float3 o, float3 d, float time, const double4* pr, bool moving,
                                              bool self_origin, float* r1, float* r2) {
    const double4 c = pr[0];
    double cx = c.x, cy = c.y, cz = c.z;
    if (moving) { const double4 v = pr[1]; cx += v.x * (double)time; cy += v.y * (double)time; cz += v.z * (double)time; }
    const double ox = (double)o.x - cx, oy = (double)o.y - cy, oz = (double)o.z - cz;
    const double dx = d.x, dy = d.y, dz = d.z;
    const double a = dx * dx + dy * dy + dz * dz;
    const double hb = ox * dx + oy * dy + oz * dz;
    const double cc = ox * ox + oy * oy + oz * oz - c.w * c.w;
    const double disc = hb * hb - a * cc;
    if (disc < 0.0) return false;
    const double sq = sqrt(disc);
    double n = (-hb - sq) / a, f = (-hb + sq) / a;
    const double nan = __longlong_as_double(0x7ff8000000000000ll);
    if (self_origin) { if (hb > 0.0) f = nan; else n = nan; }
    *r1 = (float)n;
    *r2 = (float)f;
    return true;
}

// ---- primitive tests. Each returns true and the accepted parameter when the op wins over the current best ----

// Sphere::hit (sphere.rs:59-89). w2 (center_vec) is fetched only for a moving sphere.
template <unsigned FEAT = FEAT_ALL, class Ops>
__device__ __forceinline__ bool sphere_test(const DevScene& S, const Ops& ops, uint32_t link, float4 w0, float4 w1, float3 o, float3 d,
                                            float a, float inv_a, float time, float tmin, float tmax, int origin, float* t_out) {
    const uint32_t flags = ((uint32_t)fbits(w0.w) >> 12) & 15u;
    const bool self_origin = starts_on(origin, link);
    float r1, r2;
    bool ok;
    if ((FEAT & FEAT_PRECISE) && (flags & FLAG_PRECISE)) {
        ok = sphere_roots_f64(o, d, time, S.precise + 2 * fbits(w1.w), flags & FLAG_MOVING, self_origin, &r1, &r2);
    } else {
        float3 c = f3(w0);
        if (flags & FLAG_MOVING) c = fma3(time, f3(ops((link & kLinkMask) + 32u)), c);   // sphere.rs:53-55
        ok = sphere_roots_f32(o - c, d, a, inv_a, w1.x, self_origin, &r1, &r2);
    }
    if (!ok) return false;
    float root = r1;  // ray_t.surrounds: open interval (sphere.rs:78-83)
    if (!(tmin < root && root < tmax)) root = r2;
    *t_out = root;
    return tmin < root && root < tmax;
}

// Quad::hit (quad.rs:97-133).
template <class Ops>
__device__ __forceinline__ bool quad_test(const Ops& ops, uint32_t link, float4 w0, float4 w1, float3 o, float3 d, float tmin,
                                          float tmax, int origin, float* t_out) {
    const float3 n = f3(w0);
    const float denom = dot(n, d);
    const uint32_t at = link & kLinkMask;
    const float4 w3 = ops(at + 48u);
    if (fabsf(denom) < 1e-8f || starts_on(origin, link)) return false;   // quad.rs:110-112; a ray cannot re-hit the plane it starts on
    const float t = fdiv(w3.x - dot(n, o), denom);
    if (!(tmin <= t && t <= tmax)) return false;                          // ray_t.contains: closed (quad.rs:115)
    const float4 w2 = ops(at + 32u);
    const float3 p = fma3(t, d, o);
    const float alpha = dot(f3(w1), p) + w1.w;
    const float beta = dot(f3(w2), p) + w2.w;
    *t_out = t;
    return !(alpha < 0.0f || alpha > 1.0f || beta < 0.0f || beta > 1.0f);
}

// Quad::cube's six quads (quad.rs:45-93) as one slab test: the nearest face hit inside [tmin, tmax] is the entry
// plane if it lies in the interval, else the exit plane (HittableList::hit keeps the closest, closed interval).
// te / tx come from slab_ch on the UNPADDED centre / half extent: good to ~2^-23 (|o| + |c|) / |d_k|, which is what
// decides hit / miss at the box's edges and the order of two candidates; the hit record takes its t from the exact
// corners (finalize_hit). A ray that starts on a face of this box goes through the exact form below instead.
__device__ __forceinline__ bool box_accept(float te, float tx, float tmin, float tmax, float* t_out) {
    float t = te;
    if (!(tmin <= t && t <= tmax)) t = tx;
    *t_out = t;
    return te <= tx && tmin <= t && t <= tmax;
}
// lo = w2.xyz, hi = w3.xyz (exact corners)
__device__ __forceinline__ bool box_test_from_face(float4 lo, float4 hi, float3 o, float3 inv, float tmin, float tmax, int face,
                                                   float* t_out) {
    float te, tx;
    slab_interval(lo, hi, o, inv, &te, &tx);
    if (!(te <= tx)) return false;
    // the plane the ray starts on cannot be hit again: 0 +z, 1 +x, 2 -z, 3 -x, 4 +y, 5 -y
    const int axis = (face == 1 || face == 3) ? 0 : (face >= 4 ? 1 : 2);
    const bool max_side = face == 0 || face == 1 || face == 4;
    const float plane = axis == 0 ? (max_side ? hi.x : lo.x) : axis == 1 ? (max_side ? hi.y : lo.y) : (max_side ? hi.z : lo.z);
    const float oa = axis == 0 ? o.x : axis == 1 ? o.y : o.z;
    const float ia = axis == 0 ? inv.x : axis == 1 ? inv.y : inv.z;
    const float t_self = (plane - oa) * ia;
    const float nan = __int_as_float(0x7fc00000);
    if (te == t_self) te = nan;
    if (tx == t_self) tx = nan;
    float t = te;
    if (!(tmin <= t && t <= tmax)) t = tx;
    *t_out = t;
    return tmin <= t && t <= tmax;
}

template <class Ops>
__device__ float boundary_closest_t(const DevScene& S, const Ops& ops, int begin, int end, const Ray& ray, float tmin, float tmax);

// ConstantMedium::hit (constant_medium.rs:34-70); the medium's box was tested by the preceding OP_INNER.
// `at` = byte offset of the op. Returns true and the scatter parameter when the medium wins; *next_word = the op after it.
template <unsigned FEAT = FEAT_ALL, class Ops>
__device__ __forceinline__ bool medium_test(const DevScene& S, const Ops& ops, uint32_t at, float4 w0, float4 w1, float3 o, float3 d,
                                            float a, float inv_a, float time, float tmin, float tmax, uint4 key, uint32_t seg,
                                            float* t_out, int* next_word) {
    const int bkind = (int)(((uint32_t)fbits(w0.w) >> 12) & 15u);
    const float4 w2 = ops(at + 32u);
    float t1 = 0.0f, t2 = 0.0f;
    bool ok = false;
    if (bkind == MEDIUM_BOUNDARY_SPHERE) {
        const uint32_t aux = (uint32_t)fbits(w2.w);
        const bool moving = (aux >> 24) & FLAG_MOVING;
        if ((FEAT & FEAT_PRECISE) && ((aux >> 24) & FLAG_PRECISE)) {
            ok = sphere_roots_f64(o, d, time, S.precise + 2 * (aux & 0xffffffu), moving, false, &t1, &t2);
        } else {
            float3 c = f3(w1);
            if (moving) c = fma3(time, f3(w2), c);
            ok = sphere_roots_f32(o - c, d, a, inv_a, w1.w, false, &t1, &t2);
        }
        // hit1 over the universe takes the near root; hit2 needs a root > hit1.t + 0.0001
        ok = ok && (t2 > t1 + 0.0001f);
        *next_word = (int)(at >> 4) + 3;
    } else if ((FEAT & FEAT_XBOX) && bkind == MEDIUM_BOUNDARY_XBOX) {
        // both boundary hits of a (rotated, translated) cube from one slab test in the cube's frame
        const float4 lo = ops(at + 48u), hi = ops(at + 64u);
        const float3 lo_ = xform_point(o, w1, w2), ld_ = xform_dir(d, w1, w2);
        slab_interval(lo, hi, lo_, safe_inv(ld_), &t1, &t2);
        const float inf = __int_as_float(0x7f800000);
        ok = (t1 <= t2) && (t2 >= t1 + 0.0001f) && fabsf(t1) < inf && fabsf(t2) < inf;   // hit2: closed interval from hit1.t + 0.0001
        *next_word = (int)(at >> 4) + 5;
    } else if (FEAT & FEAT_RARE) {
        Ray lr; lr.o = o; lr.d = d; lr.time = time;
        const float inf = __int_as_float(0x7f800000);
        t1 = boundary_closest_t(S, ops, fbits(w1.x), fbits(w1.y), lr, -inf, inf);
        ok = (t1 == t1);
        if (ok) { t2 = boundary_closest_t(S, ops, fbits(w1.x), fbits(w1.y), lr, t1 + 0.0001f, inf); ok = (t2 == t2); }
        *next_word = fbits(w1.y);
    }
    if (!ok) return false;
    t1 = fmaxf(t1, tmin);
    t2 = fminf(t2, tmax);
    if (!(t1 < t2)) return false;
    t1 = fmaxf(t1, 0.0f);
    const float ray_length = fsqrt(a);
    const float inside = (t2 - t1) * ray_length;
    const float u = u01(draw(key, seg, P_MEDIUM + (uint32_t)fbits(w0.z)).x);
    const float hit_distance = w0.x * flog(u);   // drawn only on this branch (constant_medium.rs:48)
    *t_out = t1 + fdiv(hit_distance, ray_length);
    return hit_distance <= inside;
}

// Hoisted (world-space) media: evaluated before the traversal of every segment; each may lower best.t.
template <unsigned FEAT = FEAT_ALL, class Ops>
__device__ __forceinline__ void media_prepass(const DevScene& S, const Ops& ops, float3 o, float3 d, float a, float inv_a, float time,
                                              float tmin, uint4 key, uint32_t seg, Best& best) {
    for (int m = 0; m < S.n_media; ++m) {
        const uint32_t at = (uint32_t)S.media_op[m] << 4;
        float t;
        int next;
        if (medium_test<FEAT>(S, ops, at, ops(at), ops(at + 16u), o, d, a, inv_a, time, tmin, best.t, key, seg, &t, &next)) {
            best.t = t; best.op = S.media_op[m]; best.xf = -1;
        }
    }
}

// Generic loop (parity kernels, medium boundary programs): every lane walks its own ray to the end. WORLD = false: t only,
// no media. `begin` / `end` are word indices.
template <bool WORLD, class Ops>
__device__ __forceinline__ void traverse(const DevScene& S, const Ops& ops, int begin, int end, const Ray& ray, float tmin, float tmax,
                                         Best& best, int origin, uint4 key, uint32_t seg) {
    float3 o = ray.o, d = ray.d;
    RaySetup R = ray_setup(o, d);
    int cur_xf = -1;
    best.t = tmax; best.op = -1; best.xf = -1;
    if (WORLD) media_prepass(S, ops, o, d, R.a, R.inv_a, ray.time, tmin, key, seg, best);
    uint32_t at = (uint32_t)begin << 4;
    const uint32_t stop = (uint32_t)end << 4;
    while (at < stop) {
        const float4 w0 = ops(at), w1 = ops(at + 16u);
        const uint32_t hdr = (uint32_t)fbits(w0.w);
        const uint32_t kind = (hdr >> 8) & 15u;
        uint32_t next = at + (hdr & 0xffu);
        float t;
        if (kind == OP_INNER || kind == OP_BOX || kind == OP_XFORM_ENTER) {
            float te, tx;
            slab_ch(w0, w1, R.inv, R.oi, &te, &tx);
            if (kind == OP_BOX) {
                bool win;
                if (starts_on(origin, at)) win = box_test_from_face(ops(at + 32u), ops(at + 48u), o, R.inv, tmin, best.t, origin & 7, &t);
                else win = box_accept(te, tx, tmin, best.t, &t);
                if (win) { best.t = t; best.op = (int)(at >> 4); best.xf = cur_xf; }
            } else if (!cull_pass(te, tx, tmin, best.t, R.eps)) {
                next = (uint32_t)fbits(w1.w) & kLinkMask;
            } else if (kind == OP_XFORM_ENTER) {   // Translate::hit / RotateY::hit (hittable.rs:96-111,159-193)
                const float4 w2 = ops(at + 32u), w3 = ops(at + 48u);
                o = xform_point(ray.o, w2, w3);    // the op holds the composed world -> local transform
                d = xform_dir(ray.d, w2, w3);
                R = ray_setup(o, d);
                cur_xf = (int)(at >> 4);
            }
        } else if (kind == OP_SPHERE) {
            if (sphere_test(S, ops, at, w0, w1, o, d, R.a, R.inv_a, ray.time, tmin, best.t, origin, &t)) { best.t = t; best.op = (int)(at >> 4); best.xf = cur_xf; }
        } else if (kind == OP_QUAD) {
            if (quad_test(ops, at, w0, w1, o, d, tmin, best.t, origin, &t)) { best.t = t; best.op = (int)(at >> 4); best.xf = cur_xf; }
        } else if (kind == OP_INNER_REF) {
            if (!aabb_hit_reference(w0, w1, o, R.inv, tmin, best.t)) next = (uint32_t)fbits(w1.w) & kLinkMask;
        } else if (kind == OP_XFORM_EXIT) {    // back in the enclosing space: the world ray, or the world ray through the parent
            const int parent = fbits(w0.x);
            o = ray.o; d = ray.d;
            if (parent >= 0) {
                const float4 p2 = ops(((uint32_t)parent << 4) + 32u), p3 = ops(((uint32_t)parent << 4) + 48u);
                o = xform_point(ray.o, p2, p3);
                d = xform_dir(ray.d, p2, p3);
            }
            R = ray_setup(o, d);
            cur_xf = parent;
        } else if (WORLD) {                    // OP_MEDIUM in the stream
            int nw;
            if (medium_test(S, ops, at, w0, w1, o, d, R.a, R.inv_a, ray.time, tmin, best.t, key, seg, &t, &nw)) { best.t = t; best.op = (int)(at >> 4); best.xf = cur_xf; }
            next = (uint32_t)nw << 4;
        } else {
            next = stop;   // a medium inside a boundary program is rejected at upload
        }
        at = next;
    }
}

// Closest t of a medium's boundary program (t only; constant_medium.rs:35-39 needs nothing else).
template <class Ops>
__device__ __noinline__ float boundary_closest_t(const DevScene& S, const Ops& ops, int begin, int end, const Ray& ray, float tmin, float tmax) {
    Best b;
    traverse<false>(S, ops, begin, end, ray, tmin, tmax, b, -1, make_uint4(0, 0, 0, 0), 0u);
    return b.op >= 0 ? b.t : __int_as_float(0x7fc00000);
}

// ------------------------------------------------------------------ hit record of the winner
struct HitRec {
    float3 p, normal;
    float t, u, v;
    int mat, prim;
    int origin;         // origin code for the ray that leaves this hit (-1 for media)
    bool front_face;
    bool uv_lazy;       // sphere: u,v derived from `sn` only if a texture asks (sphere.rs:87 computes it always)
    float3 sn;          // sphere outward normal in the sphere's own space
};

__device__ __forceinline__ void sphere_uv(float3 n, float* u, float* v) {   // sphere.rs:48-52
    const float PI = 3.14159265358979323846f;
    const float theta = acosf(-n.y);
    const float phi = atan2f(-n.z, n.x) + PI;
    *u = phi * (0.5f / PI);
    *v = theta * (1.0f / PI);
}

// hit point in f64 relative to the centre: keeps the normal of a huge sphere accurate
__device__ __noinline__ float3 precise_sphere_normal(const double4* pr, bool moving, float3 o, float3 d, float t, float time) {
    const double4 c = pr[0];
    double cx = c.x, cy = c.y, cz = c.z;
    if (moving) { const double4 cv = pr[1]; cx += cv.x * (double)time; cy += cv.y * (double)time; cz += cv.z * (double)time; }
    const double inv_r = 1.0 / c.w;
    return f3((float)((((double)o.x - cx) + (double)t * (double)d.x) * inv_r),
              (float)((((double)o.y - cy) + (double)t * (double)d.y) * inv_r),
              (float)((((double)o.z - cz) + (double)t * (double)d.z) * inv_r));
}

template <unsigned FEAT = FEAT_ALL, class Ops>
__device__ __forceinline__ void finalize_hit(const DevScene& S, const Ops& ops, const Ray& ray, const Best& best, HitRec& h) {
    float3 o = ray.o, d = ray.d;
    float4 x2, x3;
    if (best.xf >= 0) {
        x2 = ops(((uint32_t)best.xf << 4) + 32u); x3 = ops(((uint32_t)best.xf << 4) + 48u);
        o = xform_point(o, x2, x3);
        d = xform_dir(d, x2, x3);
    }
    const uint32_t at = (uint32_t)best.op << 4;
    const float4 w0 = ops(at), w1 = ops(at + 16u);
    const uint32_t hdr = (uint32_t)fbits(w0.w);
    const uint32_t kind = (hdr >> 8) & 15u;
    float t = best.t;
    h.uv_lazy = false;
    h.u = 0.0f; h.v = 0.0f;
    h.origin = origin_code(best.op, 0);
    float3 outward;
    if (kind == OP_SPHERE) {
        const uint32_t flags = (hdr >> 12) & 15u;
        const float3 pl = fma3(t, d, o);
        if ((FEAT & FEAT_PRECISE) && (flags & FLAG_PRECISE)) {
            outward = precise_sphere_normal(S.precise + 2 * fbits(w1.w), flags & FLAG_MOVING, o, d, t, ray.time);
        } else {
            float3 c = f3(w0);
            if (flags & FLAG_MOVING) c = fma3(ray.time, f3(ops(at + 32u)), c);
            outward = (pl - c) * frcp(w1.x);   // (p - center) / radius, reciprocal-multiply (vec3.rs:244-249)
        }
        h.mat = fbits(w1.y);
        h.prim = fbits(w1.z);
        h.uv_lazy = true;
        h.sn = outward;
    } else if (kind == OP_QUAD) {
        const float4 w2 = ops(at + 32u), w3 = ops(at + 48u);
        const float3 pl = fma3(t, d, o);
        outward = f3(w0);
        h.u = dot(f3(w1), pl) + w1.w;
        h.v = dot(f3(w2), pl) + w2.w;
        h.mat = fbits(w3.y);
        h.prim = fbits(w3.z);
    } else if (kind == OP_BOX) {
        // The traversal's t for a box is good to a few ulp of the coordinates (box_accept). The record's t is the
        // parameter of the face plane nearest to it, from the exact corners in the cancellation-free form; on an
        // edge the face that comes later in the list wins (closed interval, hittable.rs:66-71).
        const float4 lo = ops(at + 32u), hi = ops(at + 48u);
        const float3 inv = safe_inv(d);
        const float tx0 = (lo.x - o.x) * inv.x, tx1 = (hi.x - o.x) * inv.x;
        const float ty0 = (lo.y - o.y) * inv.y, ty1 = (hi.y - o.y) * inv.y;
        const float tz0 = (lo.z - o.z) * inv.z, tz1 = (hi.z - o.z) * inv.z;
        int face = 0;
        float miss = __int_as_float(0x7f800000), tf = t;
        { const float m = fabsf(tz1 - t); if (m <= miss) { miss = m; face = 0; tf = tz1; } }
        { const float m = fabsf(tx1 - t); if (m <= miss) { miss = m; face = 1; tf = tx1; } }
        { const float m = fabsf(tz0 - t); if (m <= miss) { miss = m; face = 2; tf = tz0; } }
        { const float m = fabsf(tx0 - t); if (m <= miss) { miss = m; face = 3; tf = tx0; } }
        { const float m = fabsf(ty1 - t); if (m <= miss) { miss = m; face = 4; tf = ty1; } }
        { const float m = fabsf(ty0 - t); if (m <= miss) { miss = m; face = 5; tf = ty0; } }
        t = tf;
        const float3 pl = fma3(t, d, o);
        const float ex = frcp(hi.x - lo.x), ey = frcp(hi.y - lo.y), ez = frcp(hi.z - lo.z);
        const float ax = (pl.x - lo.x) * ex, ay = (pl.y - lo.y) * ey, az = (pl.z - lo.z) * ez;   // 0..1 along +x,+y,+z
        outward = f3(0.0f, 0.0f, 0.0f);
        switch (face) {   // (u, v) = (alpha, beta) of the face's quad (quad.rs:55-90)
            case 0: outward.z = 1.0f;  h.u = ax;        h.v = ay; break;
            case 1: outward.x = 1.0f;  h.u = 1.0f - az; h.v = ay; break;
            case 2: outward.z = -1.0f; h.u = 1.0f - ax; h.v = ay; break;
            case 3: outward.x = -1.0f; h.u = az;        h.v = ay; break;
            case 4: outward.y = 1.0f;  h.u = ax;        h.v = 1.0f - az; break;
            default: outward.y = -1.0f; h.u = ax;       h.v = az; break;
        }
        h.mat = fbits(hi.w);
        h.prim = fbits(lo.w) + face;
        h.origin = origin_code(best.op, face);
    } else {  // OP_MEDIUM: HitRecord::new(r.at(t), phase, t, r, r.direction) (constant_medium.rs:52-58)
        outward = d;
        h.mat = fbits(w0.y);
        h.prim = fbits(w0.z);
        h.origin = -1;
    }
    h.t = t;
    h.front_face = dot(d, outward) < 0.0f;                  // hittable.rs:23
    float3 n = h.front_face ? outward : -outward;
    if (best.xf >= 0) n = xform_dir_back(n, x2, x3);        // hittable.rs:176-179 (Translate leaves the normal alone)
    h.normal = n;
    h.p = fma3(t, ray.d, ray.o);                            // t is preserved by the instance transforms
}

// ------------------------------------------------------------------ textures
struct PerlinShared {
    const float4* vec;      // shared memory: the first n_shared tables of the scene
    const uint8_t* perm;    // shared memory
    int n_shared;           // tables beyond this many are read from global memory (every NoiseTexture::new owns a table,
                            // texture.rs:100; four fit beside the render kernel's other shared data)
};

// perlin.rs:27-50,81-100. rv: 256 gradients, perm: perm_x | perm_y | perm_z of this table (shared or global memory).
__device__ __forceinline__ float perlin_noise(const float4* rv, const uint8_t* px, float3 p) {
    const float fx = floorf(p.x), fy = floorf(p.y), fz = floorf(p.z);
    const int i = (int)fx, j = (int)fy, k = (int)fz;
    const float u = p.x - fx, v = p.y - fy, w = p.z - fz;
    const float uu = u * u * (3.0f - 2.0f * u);
    const float vv = v * v * (3.0f - 2.0f * v);
    const float ww = w * w * (3.0f - 2.0f * w);
    const uint8_t* py = px + 256;
    const uint8_t* pz = px + 512;
    const int xi0 = px[i & 255], xi1 = px[(i + 1) & 255];
    const int yi[2] = {py[j & 255], py[(j + 1) & 255]};
    const int zi[2] = {pz[k & 255], pz[(k + 1) & 255]};
    float acc = 0.0f;
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b)
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const float4 g = rv[(a ? xi1 : xi0) ^ yi[b] ^ zi[c]];
                const float wa = a ? uu : 1.0f - uu, wb = b ? vv : 1.0f - vv, wc = c ? ww : 1.0f - ww;
                acc += wa * wb * wc * (g.x * (u - (float)a) + g.y * (v - (float)b) + g.z * (w - (float)c));
            }
    return acc;
}

// perlin.rs:52-64, depth 7. Two out-of-line copies, one per address space of the tables; a scene with up to four
// NoiseTextures only ever runs the shared-memory one.
__device__ __noinline__ float perlin_turbulence_shared(const float4* rv, const uint8_t* perm, float3 p) {
    float acc = 0.0f, w = 1.0f;
#pragma unroll 1
    for (int o = 0; o < 7; ++o) {
        acc = fmaf(w, perlin_noise(rv, perm, p), acc);
        w *= 0.5f;
        p = p * 2.0f;
    }
    return fabsf(acc);
}
__device__ __noinline__ float perlin_turbulence_global(const float4* __restrict__ rv, const uint8_t* __restrict__ perm, float3 p) {
    float acc = 0.0f, w = 1.0f;
#pragma unroll 1
    for (int o = 0; o < 7; ++o) {
        acc = fmaf(w, perlin_noise(rv, perm, p), acc);
        w *= 0.5f;
        p = p * 2.0f;
    }
    return fabsf(acc);
}
__device__ __forceinline__ float perlin_turbulence(const DevScene& S, const PerlinShared& P, int table, float3 p) {
    if (table < P.n_shared) return perlin_turbulence_shared(P.vec + table * 256, P.perm + table * 768, p);
    return perlin_turbulence_global(S.perlin_vec + table * 256, S.perlin_perm + table * 768, p);
}

// Texture::value (texture.rs:12-14). One out-of-line copy: it is called once per shaded hit, and keeping it (and
// the Perlin code behind it) out of the render loop's body keeps the loop's instruction footprint small.
__device__ __noinline__ float3 texture_value(const DevScene& S, const PerlinShared P, int tex, float3 p, float u, float v,
                                             bool uv_lazy, float3 sn) {
    const float4* __restrict__ T = S.texs;
    for (int guard = 0; guard < 16; ++guard) {
        const float4 t0 = __ldg(T + 2 * tex);
        const int kind = fbits(t0.x);
        if (kind == RT_TEX_SOLID) {
            return f3(__ldg(T + 2 * tex + 1));
        } else if (kind == RT_TEX_CHECKER) {   // texture.rs:59-70
            const int x = (int)floorf(t0.w * p.x), y = (int)floorf(t0.w * p.y), z = (int)floorf(t0.w * p.z);
            tex = ((x + y + z) % 2 == 0) ? fbits(t0.y) : fbits(t0.z);
        } else if (kind == RT_TEX_IMAGE) {     // texture.rs:82-93
            if (uv_lazy) { sphere_uv(sn, &u, &v); uv_lazy = false; }
            const DevImage im = S.images[fbits(t0.y)];
            const float uc = fminf(fmaxf(u, 0.0f), 1.0f);
            const float vc = 1.0f - fminf(fmaxf(v, 0.0f), 1.0f);
            const uint32_t i = (uint32_t)(uc * (float)(im.width - 1));
            const uint32_t j = (uint32_t)(vc * (float)(im.height - 1));
            return f3(__ldg(im.texels + (size_t)j * im.width + i));
        } else {                                // texture.rs:107-111
            const float s = fsin(t0.w * p.z + 10.0f * perlin_turbulence(S, P, fbits(t0.y), p)) * 0.5f + 0.5f;
            return f3(s, s, s);
        }
    }
    return f3(0.0f, 0.0f, 0.0f);
}

// ------------------------------------------------------------------ materials
__device__ __forceinline__ float3 reflect3(float3 v, float3 n) { return v - (2.0f * dot(v, n)) * n; }   // vec3.rs:91-93

// Material::emitted + Material::scatter (material.rs:26-138). Returns true if the path continues; updates ray and
// throughput, adds emission to L. Written with a single texture call site.
__device__ __forceinline__ bool shade(const DevScene& S, const PerlinShared& P, Ray& ray, HitRec& h, uint4 key, uint32_t seg,
                                      float3& L, float3& T) {
    const float4 m0 = __ldg(S.mats + 2 * h.mat);
    const int kind = fbits(m0.x);
    float3 dir = h.normal, att = f3(1.0f, 1.0f, 1.0f);
    bool scattered = true;
    if (kind == RT_MAT_DIELECTRIC) {          // material.rs:81-103
        const float ratio = h.front_face ? frcp(m0.z) : m0.z;
        const float3 unit = normalize3(ray.d);
        const float cos_theta = fminf(dot(-unit, h.normal), 1.0f);
        const float sin_theta = fsqrt(1.0f - cos_theta * cos_theta);
        float r0 = fdiv(1.0f - ratio, 1.0f + ratio);
        r0 = r0 * r0;
        const float x = 1.0f - cos_theta;
        const float refl = r0 + (1.0f - r0) * (x * x * x * x * x);           // material.rs:74-78
        if (ratio * sin_theta > 1.0f || refl > u01(draw(key, seg, P_SCATTER).w)) {
            dir = reflect3(unit, h.normal);
        } else {                                                               // vec3.rs:96-101
            const float3 perp = ratio * (unit + cos_theta * h.normal);
            dir = perp + (-fsqrt(fabsf(1.0f - dot(perp, perp)))) * h.normal;
        }
    } else if (kind != RT_MAT_DIFFUSE_LIGHT) {
        const uint4 r = draw(key, seg, P_SCATTER);
        const float3 uv = unit_vector(u01(r.x), u01(r.y));
        if (kind == RT_MAT_METAL) {           // material.rs:54-63
            dir = reflect3(normalize3(ray.d), h.normal) + (m0.z * cbrtf(u01(r.z))) * uv;
            scattered = dot(dir, h.normal) > 0.0f;
            att = f3(__ldg(S.mats + 2 * h.mat + 1));
        } else if (kind == RT_MAT_LAMBERTIAN) {   // material.rs:27-41
            dir = h.normal + uv;
            if (fabsf(dir.x) < 1e-8f && fabsf(dir.y) < 1e-8f && fabsf(dir.z) < 1e-8f) dir = h.normal;
        } else {                              // isotropic: material.rs:132-138
            dir = uv;
        }
    }
    if (kind == RT_MAT_LAMBERTIAN || kind == RT_MAT_ISOTROPIC || kind == RT_MAT_DIFFUSE_LIGHT) {
        const int tex = fbits(m0.y);
        const float4 t0 = __ldg(S.texs + 2 * tex);
        // SolidColor (texture.rs:32-36) is by far the most common texture: answer it here, call out for the rest
        const float3 c = fbits(t0.x) == RT_TEX_SOLID ? f3(__ldg(S.texs + 2 * tex + 1))
                                                     : texture_value(S, P, tex, h.p, h.u, h.v, h.uv_lazy, h.sn);
        if (kind == RT_MAT_DIFFUSE_LIGHT) {   // emitted (both faces), no scatter: material.rs:114-122
            L = L + T * c;
            return false;
        }
        att = c;
    }
    if (!scattered) return false;
    T = T * att;
    ray.o = h.p;
    ray.d = dir;
    return true;
}

// ------------------------------------------------------------------ camera
template <bool DEFOCUS = true>
__device__ __forceinline__ Ray camera_ray(const DevCamera& C, int px, int py, uint4 key) {   // camera.rs:112-137
    const uint4 r = draw(key, 0u, P_CAMERA);
    const float sx = (float)px + (-0.5f + u01(r.x));
    const float sy = (float)py + (-0.5f + u01(r.y));
    float3 rel = fma3(sy, C.dv, fma3(sx, C.du, C.rel00));   // pixel_sample - center
    Ray ray;
    ray.o = C.center;
    if (DEFOCUS && C.defocus) {
        const uint4 e = draw(key, 0u, P_CAMERA_DISK);
        const float rad = sqrtf(u01(e.x));
        float s, c;
        fsincos_turn(u01(e.y), &s, &c);
        const float3 off = (rad * c) * C.disk_u + (rad * s) * C.disk_v;      // camera.rs:128-131
        ray.o = C.center + off;
        rel = rel - off;
    }
    ray.d = rel;
    ray.time = u01(r.z);
    return ray;
}

}  // namespace rtdev
