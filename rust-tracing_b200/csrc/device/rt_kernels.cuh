// Device code of the ray_color hot path (renderer.rs:26-49,139-155 and everything it calls),
// written for sm_100a. f32 arithmetic except where a primitive is flagged FLAG_PRECISE.
//
//   traverse<>()      world.hit(ray, ray_t): BVHNode / HittableList / Translate / RotateY /
//                     Sphere / Quad / ConstantMedium (bvh.rs:90-113, hittable.rs:61-188,
//                     sphere.rs:59-89, quad.rs:97-133, constant_medium.rs:34-70) as one stackless
//                     loop over the threaded op stream (dev_scene.h)
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
__device__ __forceinline__ float3 normalize3(float3 a) { return a * rsqrtf(dot(a, a)); }
__device__ __forceinline__ int fbits(float f) { return __float_as_int(f); }

// ------------------------------------------------------------------ keyed RNG (shared spec with the oracle)
// pcg4d (Jarzynski & Olano, JCGT 9(3) 2020). path key = pcg4d(pixel, sample, seed_lo, seed_hi);
// draw(purpose) = pcg4d(key.x, key.y, key.z + segment, key.w + purpose); u01 = (x >> 8) * 2^-24.
__device__ __forceinline__ uint4 pcg4d(uint4 v) {
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
    const float r = sqrtf(fmaxf(0.0f, 1.0f - z * z));
    float s, c;
    sincospif(2.0f * u1, &s, &c);
    return f3(r * c, r * s, z);
}

// ------------------------------------------------------------------ traversal
struct Ray {
    float3 o, d;
    float time;
};

struct Best {
    float t;
    int op;   // word index of the winning primitive / medium op, -1 = none
    int xf;   // word index of the enclosing OP_XFORM_ENTER, -1 = world space
};

// Per-lane traversal cursor over the threaded op stream.
struct Trav {
    float3 o, d, inv;   // current (possibly instance-local) ray and its reciprocal direction
    float3 so, sd;      // the world ray: every OP_XFORM_ENTER holds the composed world -> local transform, nested or not
    int cur_xf;         // word index of the active OP_XFORM_ENTER, -1 = none
    int i;              // word index of the next op
    Best best;
};

// origin code of a ray that starts on a surface: (op word index << 3) | box face; -1 = none.
__device__ __forceinline__ int origin_code(int op, int face) { return (op << 3) | face; }

__device__ __forceinline__ float3 safe_inv(float3 d) { return f3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z); }

// Entry / exit parameters of the ray against the box {w0.xyz, w1.xyz}, not clamped to any interval.
// Near/far planes are chosen by the sign of the reciprocal direction (as aabb.rs:73-75 swaps on inv_d < 0), and
// fminf/fmaxf drop a NaN operand like f64::min/max do: a 0*inf product (origin exactly on a slab plane of a ray
// parallel to it) leaves that axis unconstrained, which is what the reference computes.
__device__ __forceinline__ void slab_interval(float4 w0, float4 w1, float3 o, float3 inv, float* t_enter, float* t_exit) {
    const float ax = (w0.x - o.x) * inv.x, bx = (w1.x - o.x) * inv.x;
    const float ay = (w0.y - o.y) * inv.y, by = (w1.y - o.y) * inv.y;
    const float az = (w0.z - o.z) * inv.z, bz = (w1.z - o.z) * inv.z;
    const bool sx = inv.x < 0.0f, sy = inv.y < 0.0f, sz = inv.z < 0.0f;
    *t_enter = fmaxf(fmaxf(sx ? bx : ax, sy ? by : ay), sz ? bz : az);
    *t_exit = fminf(fminf(sx ? ax : bx, sy ? ay : by), sz ? az : bz);
}

// AABB::hit (aabb.rs:64-84) as a tight slab test with the reciprocal hoisted per ray (permitted
// substitution, SURVEY.md §8(a)-Q: it only culls more, it never changes which hits exist). The exit
// side is inflated by a few ulp so f32 rounding can never cull a box the ray grazes.
__device__ __forceinline__ bool slab(float4 w0, float4 w1, float3 o, float3 inv, float tmin, float tmax) {
    float te, tx;
    slab_interval(w0, w1, o, inv, &te, &tx);
    te = fmaxf(te, tmin);
    tx = fminf(tx, tmax);
    return te <= tx * 1.0000012f + 1e-30f || te <= tx;
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
// (entering an instance) and finalize_hit() (recomputing the local ray of the winner) must produce the SAME bits, because
// a cube's face is recognised by t == plane parameter exactly. With plain operators the two inline sites could get
// different FMA contractions: a 1-ulp difference then mis-identifies the face, the self-intersection guard of the next
// segment looks at the wrong plane, and the path re-hits its own surface until max_depth (seen once the instance code
// was restructured: a handful of 90 000 Cornell paths trapped; the wavefront kernel, compiled separately, was unaffected).
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
__device__ __forceinline__ bool sphere_roots_f32(float3 oc, float3 d, float r, bool self_origin, float* r1, float* r2) {
    const float a = dot(d, d);
    const float hb = dot(oc, d);
    const float inv_a = 1.0f / a;
    // discriminant/a = r^2 - |oc - (hb/a) d|^2 : no cancellation between hb^2 and a*c for distant origins
    const float3 l = fma3(-hb * inv_a, d, oc);
    const float disc = fmaf(r, r, -dot(l, l));
    if (disc < 0.0f) return false;
    const float sq = sqrtf(disc * a);
    const float cc = fmaf(-r, r, dot(oc, oc));
    const float nan = __int_as_float(0x7fc00000);
    float near_root, far_root;
    if (hb > 0.0f) {            // both roots via q to avoid -hb + sq cancellation
        const float q = -hb - sq;
        near_root = q * inv_a;
        far_root = self_origin ? nan : cc / q;
    } else {
        const float q = -hb + sq;
        far_root = q * inv_a;
        near_root = self_origin ? nan : cc / q;
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

// ---- one op each; every function advances T.i past the op (or along its skip link) ----

__device__ __forceinline__ void op_inner(Trav& T, float4 w0, float4 w1, float tmin) {
    T.i = slab(w0, w1, T.o, T.inv, tmin, T.best.t) ? T.i + 2 : fbits(w1.w);
}
__device__ __forceinline__ void op_inner_ref(Trav& T, float4 w0, float4 w1, float tmin) {
    T.i = aabb_hit_reference(w0, w1, T.o, T.inv, tmin, T.best.t) ? T.i + 2 : fbits(w1.w);
}

__device__ __forceinline__ void op_sphere(const DevScene& S, Trav& T, float4 w0, float4 w1, float time, float tmin, int origin) {
    const uint32_t flags = ((uint32_t)fbits(w0.w) >> 4) & 15u;
    const bool moving = flags & FLAG_MOVING;
    const bool self_origin = (origin >> 3) == T.i && origin >= 0;
    float r1, r2;
    bool ok;
    if (flags & FLAG_PRECISE) {
        ok = sphere_roots_f64(T.o, T.d, time, S.precise + 2 * fbits(w1.w), moving, self_origin, &r1, &r2);
    } else {
        float3 c = f3(w0);
        if (moving) c = fma3(time, f3(__ldg(S.ops + T.i + 2)), c);  // sphere.rs:53-55
        ok = sphere_roots_f32(T.o - c, T.d, w1.x, self_origin, &r1, &r2);
    }
    if (ok) {
        float root = r1;  // ray_t.surrounds: open interval (sphere.rs:78-83)
        if (!(tmin < root && root < T.best.t)) root = r2;
        if (tmin < root && root < T.best.t) { T.best.t = root; T.best.op = T.i; T.best.xf = T.cur_xf; }
    }
    T.i += moving ? 3 : 2;
}

__device__ __forceinline__ void op_quad(const DevScene& S, Trav& T, float4 w0, float4 w1, float tmin, int origin) {
    const float3 n = f3(w0);
    const float denom = dot(n, T.d);
    const float4 w3 = __ldg(S.ops + T.i + 3);
    const bool self_origin = (origin >> 3) == T.i && origin >= 0;    // a ray cannot re-hit the plane it starts on
    if (!(fabsf(denom) < 1e-8f) && !self_origin) {                   // quad.rs:110-112
        const float t = (w3.x - dot(n, T.o)) / denom;
        if (tmin <= t && t <= T.best.t) {                             // ray_t.contains: closed (quad.rs:115)
            const float4 w2 = __ldg(S.ops + T.i + 2);
            const float3 p = fma3(t, T.d, T.o);
            const float alpha = dot(f3(w1), p) + w1.w;
            const float beta = dot(f3(w2), p) + w2.w;
            if (!(alpha < 0.0f || alpha > 1.0f || beta < 0.0f || beta > 1.0f)) { T.best.t = t; T.best.op = T.i; T.best.xf = T.cur_xf; }
        }
    }
    T.i += 4;
}

// Quad::cube's six quads (quad.rs:45-93) as one slab test: the nearest face hit inside [tmin, best.t] is the
// entry plane if it lies in the interval, else the exit plane (HittableList::hit keeps the closest, closed interval).
__device__ __forceinline__ void op_box(Trav& T, float4 w0, float4 w1, float tmin, int origin) {
    float te, tx;
    slab_interval(w0, w1, T.o, T.inv, &te, &tx);
    if (te <= tx) {
        if ((origin >> 3) == T.i && origin >= 0) {   // ray starts on a face of this box: that plane cannot be hit again
            const int face = origin & 7;             // 0 +z, 1 +x, 2 -z, 3 -x, 4 +y, 5 -y
            const int axis = (face == 1 || face == 3) ? 0 : (face >= 4 ? 1 : 2);
            const bool max_side = face == 0 || face == 1 || face == 4;
            const float plane = axis == 0 ? (max_side ? w1.x : w0.x) : axis == 1 ? (max_side ? w1.y : w0.y) : (max_side ? w1.z : w0.z);
            const float oa = axis == 0 ? T.o.x : axis == 1 ? T.o.y : T.o.z;
            const float ia = axis == 0 ? T.inv.x : axis == 1 ? T.inv.y : T.inv.z;
            const float t_self = (plane - oa) * ia;
            const float nan = __int_as_float(0x7fc00000);
            if (te == t_self) te = nan;
            if (tx == t_self) tx = nan;
        }
        float t = te;
        if (!(tmin <= t && t <= T.best.t)) t = tx;
        if (tmin <= t && t <= T.best.t) { T.best.t = t; T.best.op = T.i; T.best.xf = T.cur_xf; }
    }
    T.i += 3;
}

// Translate::hit / RotateY::hit (hittable.rs:96-111,159-193). The box is tested with the current ray (the space the
// instance sits in); the transform stored in the op is the composition of all enclosing instances, applied to the
// WORLD ray, so an instance nested in another instance's subtree needs no stack of saved rays.
__device__ __forceinline__ void op_xform_enter(const DevScene& S, Trav& T, float4 w0, float4 w1, float tmin) {
    if (slab(w0, w1, T.o, T.inv, tmin, T.best.t)) {
        const float4 w2 = __ldg(S.ops + T.i + 2), w3 = __ldg(S.ops + T.i + 3);
        T.o = xform_point(T.so, w2, w3);     // hittable.rs:98,164-168
        T.d = xform_dir(T.sd, w2, w3);
        T.inv = safe_inv(T.d);
        T.cur_xf = T.i;
        T.i += 4;
    } else {
        T.i = fbits(w1.w);
    }
}

// Back in the enclosing space: the world ray, or the world ray through the parent instance named by the exit op.
__device__ __forceinline__ void xform_restore(const DevScene& S, Trav& T, float3 wo, float3 wd, int parent) {
    T.o = wo; T.d = wd;
    if (parent >= 0) {
        const float4 p2 = __ldg(S.ops + parent + 2), p3 = __ldg(S.ops + parent + 3);
        T.o = xform_point(wo, p2, p3);
        T.d = xform_dir(wd, p2, p3);
    }
    T.inv = safe_inv(T.d);
    T.cur_xf = parent;
}
__device__ __forceinline__ void op_xform_exit(const DevScene& S, Trav& T, float4 w0) {
    xform_restore(S, T, T.so, T.sd, fbits(w0.x));
    T.i += 2;
}

// INNER / BOX / XFORM_ENTER / XFORM_EXIT in one body that shares the slab arithmetic (the render kernel's
// "slab class"). Returns the class of the lane's next op, read from the header's successor bits.
// `world(o, d)` fetches the world ray (the render kernels keep it out of registers); it is called only where an
// instance is entered from inside another one or left.
template <class WorldRay>
__device__ __forceinline__ uint32_t op_slab_class(const DevScene& S, Trav& T, float4 w0, float4 w1, float tmin, int origin, WorldRay world) {
    const uint32_t hdr = (uint32_t)fbits(w0.w);
    const uint32_t kind = hdr & 15u;
    const uint32_t ft = (hdr >> 8) & 7u, sk = (hdr >> 11) & 7u;
    if (kind == OP_XFORM_EXIT) {
        float3 wo, wd;
        world(wo, wd);
        xform_restore(S, T, wo, wd, fbits(w0.x));
        T.i += 2;
        return ft;
    }
    if (kind == OP_INNER_REF) {
        const bool pass = aabb_hit_reference(w0, w1, T.o, T.inv, tmin, T.best.t);
        T.i = pass ? T.i + 2 : fbits(w1.w);
        return pass ? ft : sk;
    }
    float te, tx;
    slab_interval(w0, w1, T.o, T.inv, &te, &tx);
    if (kind == OP_BOX) {
        if ((origin >> 3) == T.i && origin >= 0) {   // rare: the ray starts on a face of this box
            op_box(T, w0, w1, tmin, origin);
            return ft;
        }
        float t = te;
        if (!(tmin <= t && t <= T.best.t)) t = tx;
        if (te <= tx && tmin <= t && t <= T.best.t) { T.best.t = t; T.best.op = T.i; T.best.xf = T.cur_xf; }
        T.i += 3;
        return ft;
    }
    const float ce = fmaxf(te, tmin), cx = fminf(tx, T.best.t);
    const bool hit = ce <= cx * 1.0000012f + 1e-30f || ce <= cx;
    if (!hit) { T.i = fbits(w1.w); return sk; }
    if (kind == OP_INNER) { T.i += 2; return ft; }
    const float4 w2 = __ldg(S.ops + T.i + 2), w3 = __ldg(S.ops + T.i + 3);   // OP_XFORM_ENTER
    float3 wo = T.o, wd = T.d;
    if (T.cur_xf >= 0) world(wo, wd);      // nested: the op holds the composed world -> local transform
    T.o = xform_point(wo, w2, w3);
    T.d = xform_dir(wd, w2, w3);
    T.inv = safe_inv(T.d);
    T.cur_xf = T.i;
    T.i += 4;
    return ft;
}

__device__ float boundary_closest_t(const DevScene& S, int begin, int end, const Ray& ray, float tmin, float tmax);

// ConstantMedium::hit (constant_medium.rs:34-70); the medium's box was tested by the preceding OP_INNER.
__device__ __forceinline__ void op_medium(const DevScene& S, Trav& T, float4 w0, float4 w1, float time, float tmin,
                                          uint4 key, uint32_t seg) {
    const int bkind = (int)(((uint32_t)fbits(w0.w) >> 4) & 15u);
    const float4 w2 = __ldg(S.ops + T.i + 2);
    float t1, t2;
    bool ok;
    int next;
    if (bkind == MEDIUM_BOUNDARY_SPHERE) {
        const uint32_t aux = (uint32_t)fbits(w2.w);
        const bool moving = (aux >> 24) & FLAG_MOVING;
        if ((aux >> 24) & FLAG_PRECISE) {
            ok = sphere_roots_f64(T.o, T.d, time, S.precise + 2 * (aux & 0xffffffu), moving, false, &t1, &t2);
        } else {
            float3 c = f3(w1);
            if (moving) c = fma3(time, f3(w2), c);
            ok = sphere_roots_f32(T.o - c, T.d, w1.w, false, &t1, &t2);
        }
        // hit1 over the universe takes the near root; hit2 needs a root > hit1.t + 0.0001
        ok = ok && (t2 > t1 + 0.0001f);
        next = T.i + 3;
    } else if (bkind == MEDIUM_BOUNDARY_XBOX) {
        // both boundary hits of a (rotated, translated) cube from one slab test in the cube's frame
        const float4 lo = __ldg(S.ops + T.i + 3), hi = __ldg(S.ops + T.i + 4);
        const float3 lo_ = xform_point(T.o, w1, w2), ld_ = xform_dir(T.d, w1, w2);
        slab_interval(lo, hi, lo_, safe_inv(ld_), &t1, &t2);
        const float inf = __int_as_float(0x7f800000);
        ok = (t1 <= t2) && (t2 >= t1 + 0.0001f) && fabsf(t1) < inf && fabsf(t2) < inf;   // hit2: closed interval from hit1.t + 0.0001
        next = T.i + 5;
    } else {
        Ray lr; lr.o = T.o; lr.d = T.d; lr.time = time;
        const float inf = __int_as_float(0x7f800000);
        t1 = boundary_closest_t(S, fbits(w1.x), fbits(w1.y), lr, -inf, inf);
        ok = (t1 == t1);
        if (ok) { t2 = boundary_closest_t(S, fbits(w1.x), fbits(w1.y), lr, t1 + 0.0001f, inf); ok = (t2 == t2); }
        next = fbits(w1.y);
    }
    if (ok) {
        t1 = fmaxf(t1, tmin);
        t2 = fminf(t2, T.best.t);
        if (t1 < t2) {
            t1 = fmaxf(t1, 0.0f);
            const float ray_length = sqrtf(dot(T.d, T.d));
            const float inside = (t2 - t1) * ray_length;
            const float u = u01(draw(key, seg, P_MEDIUM + (uint32_t)fbits(w0.z)).x);
            const float hit_distance = w0.x * logf(u);   // drawn only on this branch (constant_medium.rs:48)
            if (hit_distance <= inside) { T.best.t = t1 + hit_distance / ray_length; T.best.op = T.i; T.best.xf = T.cur_xf; }
        }
    }
    T.i = next;
}

__device__ __forceinline__ void trav_begin(Trav& T, const Ray& ray, int begin, float tmax) {
    T.o = ray.o; T.d = ray.d;
    T.inv = safe_inv(ray.d);
    T.so = ray.o; T.sd = ray.d;
    T.cur_xf = -1;
    T.i = begin;
    T.best.t = tmax; T.best.op = -1; T.best.xf = -1;
}

// Hoisted (world-space) media: evaluated before the traversal of every segment; each may lower best.t.
__device__ __forceinline__ void media_prepass(const DevScene& S, Trav& T, float time, float tmin, uint4 key, uint32_t seg) {
    const int begin = T.i;
    for (int m = 0; m < S.n_media; ++m) {
        T.i = S.media_op[m];
        op_medium(S, T, __ldg(S.ops + T.i), __ldg(S.ops + T.i + 1), time, tmin, key, seg);
    }
    T.i = begin;
}

// Generic loop (parity kernels, medium boundary programs). WORLD = false: t only, no media.
template <bool WORLD>
__device__ __forceinline__ void traverse(const DevScene& S, int begin, int end, const Ray& ray, float tmin, float tmax,
                                         Best& best, int origin, uint4 key, uint32_t seg) {
    Trav T;
    trav_begin(T, ray, begin, tmax);
    if (WORLD) media_prepass(S, T, ray.time, tmin, key, seg);
    const float4* __restrict__ ops = S.ops;
    while (T.i < end) {
        const float4 w0 = __ldg(ops + T.i);
        const float4 w1 = __ldg(ops + T.i + 1);
        const uint32_t kind = (uint32_t)fbits(w0.w) & 15u;
        if (kind == OP_INNER) op_inner(T, w0, w1, tmin);
        else if (kind == OP_INNER_REF) op_inner_ref(T, w0, w1, tmin);
        else if (kind == OP_SPHERE) op_sphere(S, T, w0, w1, ray.time, tmin, origin);
        else if (kind == OP_BOX) op_box(T, w0, w1, tmin, origin);
        else if (kind == OP_QUAD) op_quad(S, T, w0, w1, tmin, origin);
        else if (kind == OP_XFORM_ENTER) op_xform_enter(S, T, w0, w1, tmin);
        else if (kind == OP_XFORM_EXIT) op_xform_exit(S, T, w0);
        else if (WORLD) op_medium(S, T, w0, w1, ray.time, tmin, key, seg);
        else T.i = end;   // a medium inside a boundary program is rejected at upload
    }
    best = T.best;
}

// Closest t of a medium's boundary program (t only; constant_medium.rs:35-39 needs nothing else).
__device__ __noinline__ float boundary_closest_t(const DevScene& S, int begin, int end, const Ray& ray, float tmin, float tmax) {
    Best b;
    traverse<false>(S, begin, end, ray, tmin, tmax, b, -1, make_uint4(0, 0, 0, 0), 0u);
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
    *u = phi / (2.0f * PI);
    *v = theta / PI;
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

__device__ __forceinline__ void finalize_hit(const DevScene& S, const Ray& ray, const Best& best, HitRec& h) {
    const float4* __restrict__ ops = S.ops;
    float3 o = ray.o, d = ray.d;
    float4 x2, x3;
    if (best.xf >= 0) {
        x2 = __ldg(ops + best.xf + 2); x3 = __ldg(ops + best.xf + 3);
        o = xform_point(o, x2, x3);
        d = xform_dir(d, x2, x3);
    }
    const float4 w0 = __ldg(ops + best.op), w1 = __ldg(ops + best.op + 1);
    const uint32_t hdr = (uint32_t)fbits(w0.w);
    const uint32_t kind = hdr & 15u;
    const float t = best.t;
    h.t = t;
    h.uv_lazy = false;
    h.u = 0.0f; h.v = 0.0f;
    h.origin = origin_code(best.op, 0);
    float3 outward;
    if (kind == OP_SPHERE) {
        const uint32_t flags = (hdr >> 4) & 15u;
        const float3 pl = fma3(t, d, o);
        if (flags & FLAG_PRECISE) {
            outward = precise_sphere_normal(S.precise + 2 * fbits(w1.w), flags & FLAG_MOVING, o, d, t, ray.time);
        } else {
            float3 c = f3(w0);
            if (flags & FLAG_MOVING) c = fma3(ray.time, f3(__ldg(ops + best.op + 2)), c);
            outward = (pl - c) * (1.0f / w1.x);   // (p - center) / radius, reciprocal-multiply (vec3.rs:244-249)
        }
        h.mat = fbits(w1.y);
        h.prim = fbits(w1.z);
        h.uv_lazy = true;
        h.sn = outward;
    } else if (kind == OP_QUAD) {
        const float4 w2 = __ldg(ops + best.op + 2), w3 = __ldg(ops + best.op + 3);
        const float3 pl = fma3(t, d, o);
        outward = f3(w0);
        h.u = dot(f3(w1), pl) + w1.w;
        h.v = dot(f3(w2), pl) + w2.w;
        h.mat = fbits(w3.y);
        h.prim = fbits(w3.z);
    } else if (kind == OP_BOX) {
        // which face produced t: the plane whose parameter equals t exactly (t was taken from these very values);
        // on an edge the face that comes later in the list wins (closed interval, hittable.rs:66-71)
        const float3 inv = safe_inv(d);
        const float tx0 = (w0.x - o.x) * inv.x, tx1 = (w1.x - o.x) * inv.x;
        const float ty0 = (w0.y - o.y) * inv.y, ty1 = (w1.y - o.y) * inv.y;
        const float tz0 = (w0.z - o.z) * inv.z, tz1 = (w1.z - o.z) * inv.z;
        // (nearest rather than equal: t was taken from these very expressions during the traversal and is normally
        // bit-equal to one of them, but the face must not hinge on two inline sites rounding identically)
        int face = 0;
        float miss = __int_as_float(0x7f800000);
        { const float m = fabsf(tz1 - t); if (m <= miss) { miss = m; face = 0; } }
        { const float m = fabsf(tx1 - t); if (m <= miss) { miss = m; face = 1; } }
        { const float m = fabsf(tz0 - t); if (m <= miss) { miss = m; face = 2; } }
        { const float m = fabsf(tx0 - t); if (m <= miss) { miss = m; face = 3; } }
        { const float m = fabsf(ty1 - t); if (m <= miss) { miss = m; face = 4; } }
        { const float m = fabsf(ty0 - t); if (m <= miss) { miss = m; face = 5; } }
        const float3 pl = fma3(t, d, o);
        const float ex = 1.0f / (w1.x - w0.x), ey = 1.0f / (w1.y - w0.y), ez = 1.0f / (w1.z - w0.z);
        const float ax = (pl.x - w0.x) * ex, ay = (pl.y - w0.y) * ey, az = (pl.z - w0.z) * ez;   // 0..1 along +x,+y,+z
        outward = f3(0.0f, 0.0f, 0.0f);
        switch (face) {   // (u, v) = (alpha, beta) of the face's quad (quad.rs:55-90)
            case 0: outward.z = 1.0f;  h.u = ax;        h.v = ay; break;
            case 1: outward.x = 1.0f;  h.u = 1.0f - az; h.v = ay; break;
            case 2: outward.z = -1.0f; h.u = 1.0f - ax; h.v = ay; break;
            case 3: outward.x = -1.0f; h.u = az;        h.v = ay; break;
            case 4: outward.y = 1.0f;  h.u = ax;        h.v = 1.0f - az; break;
            default: outward.y = -1.0f; h.u = ax;       h.v = az; break;
        }
        h.mat = fbits(w1.w);
        h.prim = fbits(__ldg(ops + best.op + 2).x) + face;
        h.origin = origin_code(best.op, face);
    } else {  // OP_MEDIUM: HitRecord::new(r.at(t), phase, t, r, r.direction) (constant_medium.rs:52-58)
        outward = d;
        h.mat = fbits(w0.y);
        h.prim = fbits(w0.z);
        h.origin = -1;
    }
    h.front_face = dot(d, outward) < 0.0f;                  // hittable.rs:23
    float3 n = h.front_face ? outward : -outward;
    if (best.xf >= 0) n = xform_dir_back(n, x2, x3);        // hittable.rs:176-179 (Translate leaves the normal alone)
    h.normal = n;
    h.p = fma3(t, ray.d, ray.o);                            // t is preserved by the instance transforms
}

// ------------------------------------------------------------------ textures
struct PerlinShared {
    const float4* vec;      // shared memory
    const uint8_t* perm;    // shared memory
};

__device__ __forceinline__ float perlin_noise(const PerlinShared& P, int table, float3 p) {   // perlin.rs:27-50,81-100
    const float fx = floorf(p.x), fy = floorf(p.y), fz = floorf(p.z);
    const int i = (int)fx, j = (int)fy, k = (int)fz;
    const float u = p.x - fx, v = p.y - fy, w = p.z - fz;
    const float uu = u * u * (3.0f - 2.0f * u);
    const float vv = v * v * (3.0f - 2.0f * v);
    const float ww = w * w * (3.0f - 2.0f * w);
    const uint8_t* px = P.perm + table * 768;
    const uint8_t* py = px + 256;
    const uint8_t* pz = px + 512;
    const float4* rv = P.vec + table * 256;
    const int xi[2] = {px[i & 255], px[(i + 1) & 255]};
    const int yi[2] = {py[j & 255], py[(j + 1) & 255]};
    const int zi[2] = {pz[k & 255], pz[(k + 1) & 255]};
    float acc = 0.0f;
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b)
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const float4 g = rv[xi[a] ^ yi[b] ^ zi[c]];
                const float wa = a ? uu : 1.0f - uu, wb = b ? vv : 1.0f - vv, wc = c ? ww : 1.0f - ww;
                acc += wa * wb * wc * (g.x * (u - (float)a) + g.y * (v - (float)b) + g.z * (w - (float)c));
            }
    return acc;
}

__device__ __noinline__ float perlin_turbulence(const PerlinShared& P, int table, float3 p) {   // perlin.rs:52-64, depth 7
    float acc = 0.0f, w = 1.0f;
#pragma unroll 1
    for (int o = 0; o < 7; ++o) {
        acc = fmaf(w, perlin_noise(P, table, p), acc);
        w *= 0.5f;
        p = p * 2.0f;
    }
    return fabsf(acc);
}

// Texture::value (texture.rs:12-14). One out-of-line copy: it is called once per shaded hit, and keeping it (and
// the Perlin code behind it) out of the render loop's body keeps the loop's instruction footprint small.
__device__ __noinline__ float3 texture_value(const DevScene& S, const PerlinShared& P, int tex, float3 p, float u, float v,
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
            const float s = sinf(t0.w * p.z + 10.0f * perlin_turbulence(P, fbits(t0.y), p)) * 0.5f + 0.5f;
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
        const float ratio = h.front_face ? 1.0f / m0.z : m0.z;
        const float3 unit = normalize3(ray.d);
        const float cos_theta = fminf(dot(-unit, h.normal), 1.0f);
        const float sin_theta = sqrtf(1.0f - cos_theta * cos_theta);
        float r0 = (1.0f - ratio) / (1.0f + ratio);
        r0 = r0 * r0;
        const float x = 1.0f - cos_theta;
        const float refl = r0 + (1.0f - r0) * (x * x * x * x * x);           // material.rs:74-78
        if (ratio * sin_theta > 1.0f || refl > u01(draw(key, seg, P_SCATTER).w)) {
            dir = reflect3(unit, h.normal);
        } else {                                                               // vec3.rs:96-101
            const float3 perp = ratio * (unit + cos_theta * h.normal);
            dir = perp + (-sqrtf(fabsf(1.0f - dot(perp, perp)))) * h.normal;
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
__device__ __forceinline__ Ray camera_ray(const DevCamera& C, int px, int py, uint4 key) {   // camera.rs:112-137
    const uint4 r = draw(key, 0u, P_CAMERA);
    const float sx = (float)px + (-0.5f + u01(r.x));
    const float sy = (float)py + (-0.5f + u01(r.y));
    float3 rel = fma3(sy, C.dv, fma3(sx, C.du, C.rel00));   // pixel_sample - center
    Ray ray;
    ray.o = C.center;
    if (C.defocus) {
        const uint4 e = draw(key, 0u, P_CAMERA_DISK);
        const float rad = sqrtf(u01(e.x));
        float s, c;
        sincospif(2.0f * u01(e.y), &s, &c);
        const float3 off = (rad * c) * C.disk_u + (rad * s) * C.disk_v;      // camera.rs:128-131
        ray.o = C.center + off;
        rel = rel - off;
    }
    ray.d = rel;
    ray.time = u01(r.z);
    return ray;
}

}  // namespace rtdev
