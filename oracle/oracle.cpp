// oracle.cpp — CPU f64 restatement of rust-tracing's ray_color hot path.
//
// TEST INFRASTRUCTURE ONLY. Nothing in the product (rust-tracing_b200/, include/) may import,
// link or execute this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// `--impl reference` legs use it, and only as the checker / reported CPU baseline.
//
// PARITY STATUS: the Rust reference cannot be compiled in this image (no cargo/rustc, crates not
// vendored) and it ships no tests, golden vectors or seeds. This restatement
// is therefore pinned against (i) the analytic known-answer values SURVEY.md §8(c) derives from
// the reference formulas and (ii) statistical closed forms (tests/test_oracle_*.py). Bit-level
// parity with the Rust binary is UNPINNED (its RNG is unseeded; nothing exists to pin against).
//
// Each function cites the reference file:line it restates. Arithmetic is IEEE f64 in the
// reference's operation order (compile with -ffp-contract=off, no -ffast-math). Two sampling
// modes exist:
//   mode 0 "keyed"    — the counter-based RNG and loop-free samplers the device uses (same
//                       distributions as the reference's rejection loops), so device f32 and
//                       oracle f64 follow the same paths up to rounding;
//   mode 1 "faithful" — a sequential per-pixel generator consumed in the reference's call order
//                       with the reference's rejection loops (vec3.rs:54-61,77-88).
// tests/test_oracle_modes.py checks that the two modes converge to the same image.
#include "../include/rt_b200.h"

#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <thread>
#include <vector>
#include <algorithm>

namespace {

typedef double FP;                                   // common.rs:1
const FP PI = 3.14159265358979323846;                // common.rs:3
const FP INF = std::numeric_limits<FP>::infinity();

// ------------------------------------------------------------------ vec3.rs
struct Vec3 {
    FP x, y, z;
    Vec3() : x(0), y(0), z(0) {}
    Vec3(FP a, FP b, FP c) : x(a), y(b), z(c) {}
    explicit Vec3(const FP* p) : x(p[0]), y(p[1]), z(p[2]) {}
    FP operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }  // vec3.rs:261-271
};
inline Vec3 operator+(Vec3 a, Vec3 b) { return Vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline Vec3 operator-(Vec3 a, Vec3 b) { return Vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline Vec3 operator-(Vec3 a) { return Vec3(-a.x, -a.y, -a.z); }
inline Vec3 operator*(Vec3 a, Vec3 b) { return Vec3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline Vec3 operator*(Vec3 a, FP s) { return Vec3(a.x * s, a.y * s, a.z * s); }
inline Vec3 operator*(FP s, Vec3 a) { return a * s; }                      // vec3.rs:216-222
inline Vec3 operator/(Vec3 a, FP s) { return a * (1.0 / s); }              // vec3.rs:244-249
inline FP dot(Vec3 a, Vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }  // vec3.rs:103-106
inline FP length_squared(Vec3 a) { return dot(a, a); }
inline FP length(Vec3 a) { return std::sqrt(dot(a, a)); }
inline Vec3 normalize(Vec3 a) { return a * (1.0 / length(a)); }           // vec3.rs:119-131
inline Vec3 cross(Vec3 a, Vec3 b) {                                        // vec3.rs:133-139
    return Vec3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
inline bool near_zero(Vec3 a) {                                            // vec3.rs:113-116
    const FP EPS = 1e-8;
    return std::fabs(a.x) < EPS && std::fabs(a.y) < EPS && std::fabs(a.z) < EPS;
}
inline Vec3 reflect(Vec3 v, Vec3 n) { return v - (2.0 * dot(v, n)) * n; }  // vec3.rs:91-93
inline Vec3 refract(Vec3 uv, Vec3 n, FP etai_over_etat) {                  // vec3.rs:96-101
    const FP cos_theta = std::fmin(dot(-uv, n), 1.0);
    const Vec3 r_out_perp = etai_over_etat * (uv + cos_theta * n);
    const Vec3 r_out_parallel = (-std::sqrt(std::fabs(1.0 - length_squared(r_out_perp)))) * n;
    return r_out_perp + r_out_parallel;
}

// ------------------------------------------------------- ray.rs, interval.rs
struct Ray {
    Vec3 origin, direction;
    FP time;
    Vec3 at(FP t) const { return origin + t * direction; }  // ray.rs:30-32
};
struct Interval {
    FP min, max;
    bool contains(FP x) const { return min <= x && x <= max; }   // interval.rs:40-42
    bool surrounds(FP x) const { return min < x && x < max; }    // interval.rs:43-45
};
inline FP clampf(FP x, FP lo, FP hi) { return x < lo ? lo : (x > hi ? hi : x); }  // f64::clamp (NaN stays NaN)

// ---------------------------------------------------------------- counters
struct Counters {
    uint64_t paths, segments, node_tests, sphere_tests, sphere_accepts, moving_sphere_tests, quad_parallel,
        quad_t_reject, quad_ab_reject, quad_accepts, translate_in, translate_hit, rotate_in, rotate_hit,
        medium_tests, medium_scatters, get_ray, get_ray_defocus, lambertian, metal, dielectric, isotropic,
        emitted, tex_solid, tex_checker, tex_image, tex_noise, depth_exhausted, escaped;
};
const int kNumCounters = sizeof(Counters) / sizeof(uint64_t);

// ------------------------------------------------------------------ sampling
// Keyed generator: pcg4d (Jarzynski & Olano, "Hash Functions for GPU Rendering", JCGT 9(3), 2020)
// applied twice — once to derive a path key from (pixel, sample, seed), once per draw.
struct U4 { uint32_t x, y, z, w; };
inline U4 pcg4d(U4 v) {
    v.x = v.x * 1664525u + 1013904223u; v.y = v.y * 1664525u + 1013904223u;
    v.z = v.z * 1664525u + 1013904223u; v.w = v.w * 1664525u + 1013904223u;
    v.x += v.y * v.w; v.y += v.z * v.x; v.z += v.x * v.y; v.w += v.y * v.z;
    v.x ^= v.x >> 16; v.y ^= v.y >> 16; v.z ^= v.z >> 16; v.w ^= v.w >> 16;
    v.x += v.y * v.w; v.y += v.z * v.x; v.z += v.x * v.y; v.w += v.y * v.z;
    return v;
}
inline FP u01(uint32_t x) { return (FP)(x >> 8) * (1.0 / 16777216.0); }  // 24-bit uniform in [0,1)

enum Purpose : uint32_t { P_CAMERA = 0, P_CAMERA_DISK = 1, P_SCATTER = 2, P_MEDIUM = 16 };

struct Xoshiro {  // faithful-mode sequential generator (stands in for thread_rng)
    uint64_t s[4];
    void seed(uint64_t z) {
        for (int i = 0; i < 4; ++i) {
            z += 0x9E3779B97F4A7C15ull;
            uint64_t x = z;
            x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
            x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
            s[i] = x ^ (x >> 31);
        }
    }
    static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
    uint64_t next() {
        const uint64_t r = rotl(s[0] + s[3], 23) + s[0];
        const uint64_t t = s[1] << 17;
        s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
        return r;
    }
    FP random() { return (FP)(next() >> 11) * (1.0 / 9007199254740992.0); }
};

struct Sampler {
    int mode;        // 0 keyed, 1 faithful
    U4 key;          // keyed: path key
    uint32_t seg;    // keyed: bounce index of the segment being traced / shaded
    Xoshiro seq;     // faithful

    void begin_path(uint64_t seed, uint32_t pixel, uint32_t sample) {
        key = pcg4d(U4{pixel, sample, (uint32_t)seed, (uint32_t)(seed >> 32)});
        seg = 0;
    }
    U4 draw(uint32_t purpose) const { return pcg4d(U4{key.x, key.y, key.z + seg, key.w + purpose}); }

    FP range(FP lo, FP hi) { return lo + (hi - lo) * seq.random(); }

    // camera.rs:133-137 pixel_sample_square's two draws, then :123 ray_time (keyed: same draw call)
    void camera(FP* px, FP* py, FP* time, bool defocus, FP* dx, FP* dy) {
        if (mode == 0) {
            const U4 d = draw(P_CAMERA);
            *px = u01(d.x); *py = u01(d.y); *time = u01(d.z);
            if (defocus) {  // uniform point in the unit disk, loop-free
                const U4 e = draw(P_CAMERA_DISK);
                const FP r = std::sqrt(u01(e.x)), phi = 2.0 * PI * u01(e.y);
                *dx = r * std::cos(phi); *dy = r * std::sin(phi);
            }
        } else {  // reference order: px, py, [disk rejection pairs], time
            *px = seq.random(); *py = seq.random();
            if (defocus) {
                for (;;) {  // vec3.rs:77-88
                    const FP a = range(-1.0, 1.0), b = range(-1.0, 1.0);
                    if (a * a + b * b + 0.0 * 0.0 < 1.0) { *dx = a; *dy = b; break; }
                }
            }
            *time = seq.random();
        }
    }
    Vec3 in_unit_sphere_faithful() {  // vec3.rs:54-61
        for (;;) {
            const FP a = range(-1.0, 1.0), b = range(-1.0, 1.0), c = range(-1.0, 1.0);
            const Vec3 p(a, b, c);
            if (length_squared(p) < 1.0) return p;
        }
    }
    static Vec3 unit_from(FP u0, FP u1) {  // uniform on the unit sphere
        const FP z = 1.0 - 2.0 * u0;
        const FP r = std::sqrt(std::fmax(0.0, 1.0 - z * z));
        const FP phi = 2.0 * PI * u1;
        return Vec3(r * std::cos(phi), r * std::sin(phi), z);
    }
    Vec3 random_unit_vector() {  // vec3.rs:63-65
        if (mode == 0) { const U4 d = draw(P_SCATTER); return unit_from(u01(d.x), u01(d.y)); }
        return normalize(in_unit_sphere_faithful());
    }
    Vec3 random_in_unit_sphere() {  // vec3.rs:54-61
        if (mode == 0) {
            const U4 d = draw(P_SCATTER);
            return unit_from(u01(d.x), u01(d.y)) * std::cbrt(u01(d.z));
        }
        return in_unit_sphere_faithful();
    }
    FP dielectric_u() {  // material.rs:94
        if (mode == 0) return u01(draw(P_SCATTER).w);
        return seq.random();
    }
    FP medium_u(int medium_id) {  // constant_medium.rs:48
        if (mode == 0) return u01(draw(P_MEDIUM + (uint32_t)medium_id).x);
        return seq.random();
    }
};

// -------------------------------------------------------------------- scene
struct Scene {
    const rt_scene_desc* d;
};

struct HitRecord {  // hittable.rs:11-19
    Vec3 p, normal;
    int mat;
    FP t, u, v;
    bool front_face;
    int prim;
};

// HitRecord::new (hittable.rs:22-37)
inline HitRecord make_hit(Vec3 p, int mat, FP t, const Ray& r, Vec3 outward_normal, int prim) {
    HitRecord h;
    h.front_face = dot(r.direction, outward_normal) < 0.0;
    h.p = p;
    h.normal = h.front_face ? outward_normal : -outward_normal;
    h.mat = mat;
    h.t = t;
    h.u = 0.0; h.v = 0.0;
    h.prim = prim;
    return h;
}

// AABB::hit (aabb.rs:64-84): each axis is tested against the ORIGINAL ray_t (no narrowing).
inline bool aabb_hit(const FP* box, const Ray& r, const Interval& ray_t) {
    for (int a = 0; a < 3; ++a) {
        const FP inv_d = 1.0 / r.direction[a];
        const FP orig = r.origin[a];
        FP t0 = (box[2 * a] - orig) * inv_d;
        FP t1 = (box[2 * a + 1] - orig) * inv_d;
        if (inv_d < 0.0) { const FP tmp = t0; t0 = t1; t1 = tmp; }
        const FP t_min = std::fmax(t0, ray_t.min);  // f64::max ignores a NaN operand, as fmax does
        const FP t_max = std::fmin(t1, ray_t.max);
        if (t_max <= t_min) return false;
    }
    return true;
}

struct Tracer {
    Scene sc;
    Sampler* rng;
    Counters* cnt;

    bool hit(int id, const Ray& r, const Interval& ray_t, HitRecord* out);
    bool bvh_hit(int node, const Ray& r, const Interval& ray_t, HitRecord* out);
};

// get_sphere_uv (sphere.rs:48-52)
inline void sphere_uv(Vec3 n, FP* u, FP* v) {
    const FP theta = std::acos(-n.y);
    const FP phi = std::atan2(-n.z, n.x) + PI;
    *u = phi / (2.0 * PI);
    *v = theta / PI;
}

// (Node, AABB)::hit (bvh.rs:90-113): box test, then left-first with interval narrowing.
bool Tracer::bvh_hit(int node, const Ray& r, const Interval& ray_t, HitRecord* out) {
    const rt_bvh_node_desc& n = sc.d->bvh_nodes[node];
    cnt->node_tests++;
    if (!aabb_hit(n.bbox, r, ray_t)) return false;
    if (n.object >= 0) return hit(n.object, r, ray_t, out);
    HitRecord hit_left;
    if (bvh_hit(n.left, r, ray_t, &hit_left)) {
        HitRecord hit_right;
        if (bvh_hit(n.right, r, Interval{ray_t.min, hit_left.t}, &hit_right)) { *out = hit_right; return true; }
        *out = hit_left;
        return true;
    }
    return bvh_hit(n.right, r, ray_t, out);
}

bool Tracer::hit(int id, const Ray& r, const Interval& ray_t, HitRecord* out) {
    const rt_hittable_desc& h = sc.d->hittables[id];
    switch (h.kind) {
        case RT_HIT_SPHERE: {  // sphere.rs:59-89
            cnt->sphere_tests++;
            const bool moving = (h.flags & RT_FLAG_MOVING) != 0;
            if (moving) cnt->moving_sphere_tests++;
            const Vec3 center = moving ? Vec3(h.v0) + Vec3(h.v1) * r.time : Vec3(h.v0);  // sphere.rs:53-55
            const FP radius = h.s0;
            const Vec3 oc = r.origin - center;
            const FP a = length_squared(r.direction);
            const FP half_b = dot(oc, r.direction);
            const FP c = length_squared(oc) - radius * radius;
            const FP discriminant = half_b * half_b - a * c;
            if (discriminant < 0.0) return false;
            const FP sqrtd = std::sqrt(discriminant);
            FP root = (-half_b - sqrtd) / a;
            if (!ray_t.surrounds(root)) {
                root = (-half_b + sqrtd) / a;
                if (!ray_t.surrounds(root)) return false;
            }
            const Vec3 p = r.at(root);
            const Vec3 outward_normal = (p - center) / radius;
            FP u, v;
            sphere_uv(outward_normal, &u, &v);
            *out = make_hit(p, h.mat, root, r, outward_normal, id);
            out->u = u; out->v = v;
            cnt->sphere_accepts++;
            return true;
        }
        case RT_HIT_QUAD: {  // quad.rs:97-133
            const Vec3 normal(h.n), q(h.v0), uu(h.v1), vv(h.v2), w(h.v3);
            const FP denom = dot(normal, r.direction);
            if (std::fabs(denom) < 1e-8) { cnt->quad_parallel++; return false; }
            const FP t = (h.s0 - dot(normal, r.origin)) / denom;
            if (!ray_t.contains(t)) { cnt->quad_t_reject++; return false; }
            const Vec3 intersection = r.at(t);
            const Vec3 planar_hit_point = intersection - q;
            const FP alpha = dot(w, cross(planar_hit_point, vv));
            const FP beta = dot(w, cross(uu, planar_hit_point));
            if (alpha < 0.0 || alpha > 1.0 || beta < 0.0 || beta > 1.0) { cnt->quad_ab_reject++; return false; }
            *out = make_hit(intersection, h.mat, t, r, normal, id);
            out->u = alpha; out->v = beta;
            cnt->quad_accepts++;
            return true;
        }
        case RT_HIT_LIST: {  // hittable.rs:61-79
            FP closest_so_far = ray_t.max;
            bool hit_anything = false;
            for (int i = 0; i < h.count; ++i) {
                HitRecord rec;
                if (hit(sc.d->list_items[h.child + i], r, Interval{ray_t.min, closest_so_far}, &rec)) {
                    closest_so_far = rec.t;
                    *out = rec;
                    hit_anything = true;
                }
            }
            return hit_anything;
        }
        case RT_HIT_TRANSLATE: {  // hittable.rs:96-111
            cnt->translate_in++;
            const Vec3 offset(h.v0);
            const Ray offset_r{r.origin - offset, r.direction, r.time};
            if (hit(h.child, offset_r, ray_t, out)) {
                out->p = out->p + offset;
                cnt->translate_hit++;
                return true;
            }
            return false;
        }
        case RT_HIT_ROTATE_Y: {  // hittable.rs:159-188
            cnt->rotate_in++;
            const FP sin_theta = h.s0, cos_theta = h.s1;
            Vec3 origin = r.origin, direction = r.direction;
            origin.x = cos_theta * r.origin.x - sin_theta * r.origin.z;
            origin.z = sin_theta * r.origin.x + cos_theta * r.origin.z;
            direction.x = cos_theta * r.direction.x - sin_theta * r.direction.z;
            direction.z = sin_theta * r.direction.x + cos_theta * r.direction.z;
            const Ray rotated_r{origin, direction, r.time};
            if (hit(h.child, rotated_r, ray_t, out)) {
                Vec3 p = out->p;
                p.x = cos_theta * out->p.x + sin_theta * out->p.z;
                p.z = -sin_theta * out->p.x + cos_theta * out->p.z;
                Vec3 normal = out->normal;
                normal.x = cos_theta * out->normal.x + sin_theta * out->normal.z;
                normal.z = -sin_theta * out->normal.x + cos_theta * out->normal.z;
                out->p = p;
                out->normal = normal;
                cnt->rotate_hit++;
                return true;
            }
            return false;
        }
        case RT_HIT_CONSTANT_MEDIUM: {  // constant_medium.rs:34-70
            cnt->medium_tests++;
            HitRecord hit1, hit2;
            if (!hit(h.child, r, Interval{-INF, INF}, &hit1)) return false;
            if (!hit(h.child, r, Interval{hit1.t + 0.0001, INF}, &hit2)) return false;
            hit1.t = std::fmax(hit1.t, ray_t.min);
            hit2.t = std::fmin(hit2.t, ray_t.max);
            if (!(hit1.t < hit2.t)) return false;
            hit1.t = std::fmax(hit1.t, 0.0);
            const FP ray_length = length(r.direction);
            const FP distance_inside_boundary = (hit2.t - hit1.t) * ray_length;
            const FP hit_distance = h.s0 * std::log(rng->medium_u(id));  // drawn only on this branch
            if (hit_distance <= distance_inside_boundary) {
                const FP t = hit1.t + hit_distance / ray_length;
                *out = make_hit(r.at(t), h.mat, t, r, r.direction, id);
                cnt->medium_scatters++;
                return true;
            }
            return false;
        }
        case RT_HIT_BVH:  // bvh.rs:115-118
            return bvh_hit(h.child, r, ray_t, out);
    }
    return false;
}

// -------------------------------------------------------- perlin.rs, texture.rs, color.rs
inline int32_t as_i32(FP x) {  // Rust `as i32`: saturating, NaN -> 0
    if (x != x) return 0;
    if (x >= 2147483647.0) return 2147483647;
    if (x <= -2147483648.0) return (int32_t)(-2147483647 - 1);
    return (int32_t)x;
}
inline uint32_t as_u32(FP x) {  // Rust `as u32`
    if (x != x || x <= 0.0) return 0;
    if (x >= 4294967295.0) return 4294967295u;
    return (uint32_t)x;
}

FP perlin_noise(const rt_perlin_desc& pn, Vec3 p) {  // perlin.rs:27-50 + :81-100
    const int32_t i = as_i32(std::floor(p.x)), j = as_i32(std::floor(p.y)), k = as_i32(std::floor(p.z));
    const FP u = p.x - (FP)i, v = p.y - (FP)j, w = p.z - (FP)k;
    Vec3 c[2][2][2];
    for (int di = 0; di < 2; ++di)
        for (int dj = 0; dj < 2; ++dj)
            for (int dk = 0; dk < 2; ++dk) {
                const int idx = pn.perm_x[(uint32_t)(i + di) & 255] ^ pn.perm_y[(uint32_t)(j + dj) & 255] ^
                                pn.perm_z[(uint32_t)(k + dk) & 255];
                c[di][dj][dk] = Vec3(pn.ranvec[idx]);
            }
    const FP uu = u * u * (3.0 - 2.0 * u);
    const FP vv = v * v * (3.0 - 2.0 * v);
    const FP ww = w * w * (3.0 - 2.0 * w);
    FP acc = 0.0;
    for (int a = 0; a < 2; ++a)
        for (int b = 0; b < 2; ++b)
            for (int d = 0; d < 2; ++d) {
                const Vec3 weight_v(u - (FP)a, v - (FP)b, w - (FP)d);
                acc += ((FP)a * uu + (FP)(1 - a) * (1.0 - uu)) * ((FP)b * vv + (FP)(1 - b) * (1.0 - vv)) *
                       ((FP)d * ww + (FP)(1 - d) * (1.0 - ww)) * dot(c[a][b][d], weight_v);
            }
    return acc;
}

FP perlin_turbulence(const rt_perlin_desc& pn, Vec3 p, int depth) {  // perlin.rs:52-64
    FP acc = 0.0, w = 1.0;
    for (int o = 0; o < depth; ++o) {
        acc += w * perlin_noise(pn, p);
        w *= 0.5;
        p = p * 2.0;
    }
    return std::fabs(acc);
}

inline FP gamma_to_linear(FP g) { return std::pow(g, 2.2); }        // color.rs:8-10
inline FP linear_to_gamma(FP l) { return std::pow(l, 1.0 / 2.2); }  // color.rs:4-6

Vec3 texture_value(const rt_scene_desc* d, int tex, FP u, FP v, Vec3 p, Counters* cnt) {
    const rt_texture_desc& t = d->textures[tex];
    switch (t.kind) {
        case RT_TEX_SOLID:  // texture.rs:32-36
            cnt->tex_solid++;
            return Vec3(t.color);
        case RT_TEX_CHECKER: {  // texture.rs:59-70
            cnt->tex_checker++;
            const int32_t x = as_i32(std::floor(t.scale * p.x));
            const int32_t y = as_i32(std::floor(t.scale * p.y));
            const int32_t z = as_i32(std::floor(t.scale * p.z));
            const int32_t sum = (int32_t)((uint32_t)x + (uint32_t)y + (uint32_t)z);  // release-mode wrapping add
            return texture_value(d, (sum % 2 == 0) ? t.a : t.b, u, v, p, cnt);
        }
        case RT_TEX_IMAGE: {  // texture.rs:82-93
            cnt->tex_image++;
            const rt_image_desc& im = d->images[t.a];
            const FP uc = clampf(u, 0.0, 1.0);
            const FP vc = 1.0 - clampf(v, 0.0, 1.0);
            const uint32_t i = as_u32(uc * (FP)(im.width - 1));
            const uint32_t j = as_u32(vc * (FP)(im.height - 1));
            const uint8_t* px = im.rgb8 + ((size_t)j * im.width + i) * 3;
            return Vec3(gamma_to_linear((FP)px[0] / 255.0), gamma_to_linear((FP)px[1] / 255.0),
                        gamma_to_linear((FP)px[2] / 255.0));  // color.rs:21-27
        }
        case RT_TEX_NOISE: {  // texture.rs:107-111
            cnt->tex_noise++;
            const FP s = std::sin(t.scale * p.z + 10.0 * perlin_turbulence(d->perlins[t.a], p, 7)) * 0.5 + 0.5;
            return Vec3(s, s, s);
        }
    }
    return Vec3();
}

// ---------------------------------------------------------------- material.rs
inline FP reflectance(FP cosine, FP ref_idx) {  // material.rs:74-78
    FP r0 = (1.0 - ref_idx) / (1.0 + ref_idx);
    r0 = r0 * r0;
    return r0 + (1.0 - r0) * std::pow(1.0 - cosine, 5.0);
}

Vec3 emitted(const rt_scene_desc* d, int mat, FP u, FP v, Vec3 p, Counters* cnt) {  // material.rs:13-15,119-121
    const rt_material_desc& m = d->materials[mat];
    if (m.kind == RT_MAT_DIFFUSE_LIGHT) { cnt->emitted++; return texture_value(d, m.tex, u, v, p, cnt); }
    return Vec3();
}

bool scatter(const rt_scene_desc* d, const Ray& ray, const HitRecord& hit, Sampler* rng, Counters* cnt,
             Ray* scattered, Vec3* attenuation) {
    const rt_material_desc& m = d->materials[hit.mat];
    switch (m.kind) {
        case RT_MAT_LAMBERTIAN: {  // material.rs:27-41
            cnt->lambertian++;
            const Vec3 scatter_direction = hit.normal + rng->random_unit_vector();
            *scattered = Ray{hit.p, near_zero(scatter_direction) ? hit.normal : scatter_direction, ray.time};
            *attenuation = texture_value(d, m.tex, hit.u, hit.v, hit.p, cnt);
            return true;
        }
        case RT_MAT_METAL: {  // material.rs:54-63
            cnt->metal++;
            const Vec3 reflected = reflect(normalize(ray.direction), hit.normal) + m.param * rng->random_in_unit_sphere();
            if (dot(reflected, hit.normal) > 0.0) {
                *scattered = Ray{hit.p, reflected, ray.time};
                *attenuation = Vec3(m.albedo);
                return true;
            }
            return false;
        }
        case RT_MAT_DIELECTRIC: {  // material.rs:81-103
            cnt->dielectric++;
            const FP refraction_ratio = hit.front_face ? 1.0 / m.param : m.param;
            const Vec3 unit_direction = normalize(ray.direction);
            const FP cos_theta = std::fmin(dot(-unit_direction, hit.normal), 1.0);
            const FP sin_theta = std::sqrt(1.0 - cos_theta * cos_theta);
            Vec3 direction;
            if (refraction_ratio * sin_theta > 1.0 || reflectance(cos_theta, refraction_ratio) > rng->dielectric_u())
                direction = reflect(unit_direction, hit.normal);
            else
                direction = refract(unit_direction, hit.normal, refraction_ratio);
            *scattered = Ray{hit.p, direction, ray.time};
            *attenuation = Vec3(1.0, 1.0, 1.0);
            return true;
        }
        case RT_MAT_DIFFUSE_LIGHT:  // material.rs:115-117
            return false;
        case RT_MAT_ISOTROPIC: {  // material.rs:132-138
            cnt->isotropic++;
            *scattered = Ray{hit.p, rng->random_unit_vector(), ray.time};
            *attenuation = texture_value(d, m.tex, hit.u, hit.v, hit.p, cnt);
            return true;
        }
    }
    return false;
}

// ------------------------------------------------------------------ camera.rs
Ray get_ray(const rt_camera_desc& c, int64_t i, int64_t j, Sampler* rng, Counters* cnt) {  // camera.rs:112-126
    cnt->get_ray++;
    const Vec3 pixel00(c.pixel00_loc), du(c.pixel_delta_u), dv(c.pixel_delta_v), center(c.center);
    const Vec3 pixel_center = pixel00 + ((FP)i * du) + ((FP)j * dv);
    const bool defocus = !(c.defocus_angle <= 0.0);
    if (defocus) cnt->get_ray_defocus++;
    FP ux, uy, time, dx = 0.0, dy = 0.0;
    rng->camera(&ux, &uy, &time, defocus, &dx, &dy);
    const FP px = -0.5 + ux, py = -0.5 + uy;                 // camera.rs:133-137
    const Vec3 pixel_sample = pixel_center + (px * du + py * dv);
    const Vec3 ray_origin = defocus ? center + dx * Vec3(c.defocus_disk_u) + dy * Vec3(c.defocus_disk_v) : center;  // :128-131
    return Ray{ray_origin, pixel_sample - ray_origin, time};
}

// ---------------------------------------------------------------- renderer.rs
Vec3 ray_color(const Ray& ray, int depth, int max_depth, Vec3 background, Tracer* tr) {  // renderer.rs:139-155
    if (depth <= 0) { tr->cnt->depth_exhausted++; return Vec3(); }
    tr->rng->seg = (uint32_t)(max_depth - depth);
    tr->cnt->segments++;
    HitRecord hit;
    if (tr->hit(tr->sc.d->world, ray, Interval{0.001, INF}, &hit)) {
        const Vec3 color_from_emission = emitted(tr->sc.d, hit.mat, hit.u, hit.v, hit.p, tr->cnt);
        Ray scattered;
        Vec3 attenuation;
        if (scatter(tr->sc.d, ray, hit, tr->rng, tr->cnt, &scattered, &attenuation))
            return color_from_emission + attenuation * ray_color(scattered, depth - 1, max_depth, background, tr);
        return color_from_emission;
    }
    tr->cnt->escaped++;
    return background;
}

void add_counters(Counters* dst, const Counters& src) {
    uint64_t* a = reinterpret_cast<uint64_t*>(dst);
    const uint64_t* b = reinterpret_cast<const uint64_t*>(&src);
    for (int i = 0; i < kNumCounters; ++i) a[i] += b[i];
}

void fill_hit_desc(bool ok, const HitRecord& h, rt_hit_desc* o) {
    std::memset(o, 0, sizeof(*o));
    o->hit = ok ? 1 : 0;
    o->prim_id = -1;
    o->mat_id = -1;
    if (!ok) return;
    o->t = h.t;
    o->p[0] = h.p.x; o->p[1] = h.p.y; o->p[2] = h.p.z;
    o->normal[0] = h.normal.x; o->normal[1] = h.normal.y; o->normal[2] = h.normal.z;
    o->u = h.u; o->v = h.v;
    o->front_face = h.front_face ? 1 : 0;
    o->prim_id = h.prim;
    o->mat_id = h.mat;
}

// Independent restatement of BVHNode::node_from_list (bvh.rs:31-66) on bare bounding boxes.
struct BvhOut {
    std::vector<int32_t> left, right, object;
    std::vector<double> bbox;
};
int bvh_build_rec(const double* boxes, std::vector<int32_t>& objs, int lo, int span, const int* axes, int n_axes,
                  int* axis_pos, BvhOut* out) {
    if (*axis_pos >= n_axes) return -1;
    const int axis = axes[(*axis_pos)++];
    const int me = (int)out->object.size();
    out->left.push_back(-1); out->right.push_back(-1); out->object.push_back(-1);
    out->bbox.resize(out->bbox.size() + 6);
    auto mn = [&](int32_t id) { return boxes[(size_t)id * 6 + 2 * axis]; };
    auto leaf = [&](int32_t id) {
        const int k = (int)out->object.size();
        out->left.push_back(-1); out->right.push_back(-1); out->object.push_back(id);
        out->bbox.insert(out->bbox.end(), boxes + (size_t)id * 6, boxes + (size_t)id * 6 + 6);
        return k;
    };
    auto unite = [&](int a, int b) {
        for (int c = 0; c < 3; ++c) {
            out->bbox[(size_t)me * 6 + 2 * c] = std::fmin(out->bbox[(size_t)a * 6 + 2 * c], out->bbox[(size_t)b * 6 + 2 * c]);
            out->bbox[(size_t)me * 6 + 2 * c + 1] = std::fmax(out->bbox[(size_t)a * 6 + 2 * c + 1], out->bbox[(size_t)b * 6 + 2 * c + 1]);
        }
    };
    if (span == 1) {
        out->object[me] = objs[lo];
        std::memcpy(&out->bbox[(size_t)me * 6], boxes + (size_t)objs[lo] * 6, 6 * sizeof(double));
    } else if (span == 2) {
        int32_t l = objs[lo], r = objs[lo + 1];
        if (!(mn(l) < mn(r))) std::swap(l, r);
        const int li = leaf(l), ri = leaf(r);
        out->left[me] = li; out->right[me] = ri;
        unite(li, ri);
    } else {
        std::stable_sort(objs.begin() + lo, objs.begin() + lo + span, [&](int32_t a, int32_t b) { return mn(a) < mn(b); });
        const int li = bvh_build_rec(boxes, objs, lo, span / 2, axes, n_axes, axis_pos, out);
        if (li < 0) return -1;
        const int ri = bvh_build_rec(boxes, objs, lo + span / 2, span - span / 2, axes, n_axes, axis_pos, out);
        if (ri < 0) return -1;
        out->left[me] = li; out->right[me] = ri;
        unite(li, ri);
    }
    return me;
}

}  // namespace

extern "C" {

int oracle_num_counters(void) { return kNumCounters; }
const char* oracle_counter_names(void) {
    return "paths,segments,node_tests,sphere_tests,sphere_accepts,moving_sphere_tests,quad_parallel,quad_t_reject,"
           "quad_ab_reject,quad_accepts,translate_in,translate_hit,rotate_in,rotate_hit,medium_tests,medium_scatters,"
           "get_ray,get_ray_defocus,lambertian,metal,dielectric,isotropic,emitted,tex_solid,tex_checker,tex_image,"
           "tex_noise,depth_exhausted,escaped";
}

// renderer.rs:26-49: per-pixel SUM over the sample range. sum_rgb: W*H*3 doubles, row-major.
// sumsq_lum (optional): per-pixel sum of squared luminance of the samples (for variance bounds).
int oracle_render(const rt_scene_desc* scene, const rt_camera_desc* cam, int64_t sample_begin, int64_t sample_count,
                  uint64_t seed, int mode, int threads, double* sum_rgb, double* sumsq_lum, uint64_t* counters_out) {
    if (!scene || !cam || !sum_rgb) return -1;
    const int64_t W = cam->image_width, H = cam->image_height, N = W * H;
    if (threads <= 0) threads = (int)std::thread::hardware_concurrency();
    if (threads <= 0) threads = 1;
    std::atomic<int64_t> next(0);
    const int64_t chunk = 64;  // pixels per work item (rayon-like dynamic scheduling)
    std::vector<Counters> per_thread(threads);
    for (auto& c : per_thread) std::memset(&c, 0, sizeof(c));
    auto worker = [&](int tid) {
        Counters& cnt = per_thread[tid];
        Sampler rng;
        rng.mode = mode;
        Tracer tr{Scene{scene}, &rng, &cnt};
        const Vec3 background(cam->background);
        for (;;) {
            const int64_t start = next.fetch_add(chunk);
            if (start >= N) break;
            const int64_t end = std::min(N, start + chunk);
            for (int64_t pos = start; pos < end; ++pos) {
                const int64_t i = pos % W, j = pos / W;  // renderer.rs:32-33
                Vec3 avg_color;
                double sq = 0.0;
                if (mode == 1) rng.seq.seed(seed * 0x9E3779B97F4A7C15ull + (uint64_t)pos * 0xD1B54A32D192ED03ull + (uint64_t)sample_begin);
                for (int64_t s = sample_begin; s < sample_begin + sample_count; ++s) {  // renderer.rs:35-40
                    rng.begin_path(seed, (uint32_t)pos, (uint32_t)s);
                    cnt.paths++;
                    const Ray r = get_ray(*cam, i, j, &rng, &cnt);
                    const Vec3 c = ray_color(r, cam->max_depth, cam->max_depth, background, &tr);
                    avg_color = avg_color + c;
                    const double lum = 0.2126 * c.x + 0.7152 * c.y + 0.0722 * c.z;
                    sq += lum * lum;
                }
                sum_rgb[pos * 3 + 0] = avg_color.x;
                sum_rgb[pos * 3 + 1] = avg_color.y;
                sum_rgb[pos * 3 + 2] = avg_color.z;
                if (sumsq_lum) sumsq_lum[pos] = sq;
            }
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; ++t) pool.emplace_back(worker, t);
    worker(0);
    for (auto& t : pool) t.join();
    if (counters_out) {
        Counters total;
        std::memset(&total, 0, sizeof(total));
        for (auto& c : per_thread) add_counters(&total, c);
        std::memcpy(counters_out, &total, sizeof(total));
    }
    return 0;
}

// Hittable::hit on a ray batch; medium draws are keyed by (seed, ray index).
int oracle_hit_batch(const rt_scene_desc* scene, const rt_ray_desc* rays, int64_t n, double t_min, double t_max,
                     uint64_t seed, rt_hit_desc* out) {
    if (!scene || (n > 0 && (!rays || !out))) return -1;
    Counters cnt;
    std::memset(&cnt, 0, sizeof(cnt));
    Sampler rng;
    rng.mode = 0;
    Tracer tr{Scene{scene}, &rng, &cnt};
    for (int64_t k = 0; k < n; ++k) {
        rng.begin_path(seed, (uint32_t)k, 0u);
        const Ray r{Vec3(rays[k].origin), Vec3(rays[k].direction), rays[k].time};
        HitRecord h;
        const bool ok = tr.hit(scene->world, r, Interval{t_min, t_max}, &h);
        fill_hit_desc(ok, h, &out[k]);
    }
    return 0;
}

int oracle_texture_batch(const rt_scene_desc* scene, int tex, const double* uvp, int64_t n, double* rgb_out) {
    if (!scene || tex < 0 || tex >= scene->n_textures) return -1;
    Counters cnt;
    std::memset(&cnt, 0, sizeof(cnt));
    for (int64_t k = 0; k < n; ++k) {
        const Vec3 c = texture_value(scene, tex, uvp[k * 5], uvp[k * 5 + 1], Vec3(uvp + k * 5 + 2), &cnt);
        rgb_out[k * 3] = c.x; rgb_out[k * 3 + 1] = c.y; rgb_out[k * 3 + 2] = c.z;
    }
    return 0;
}

int oracle_get_ray_batch(const rt_camera_desc* cam, const int64_t* pixel_index, const int64_t* sample_index, int64_t n,
                         uint64_t seed, rt_ray_desc* out) {
    if (!cam) return -1;
    Counters cnt;
    std::memset(&cnt, 0, sizeof(cnt));
    Sampler rng;
    rng.mode = 0;
    for (int64_t k = 0; k < n; ++k) {
        rng.begin_path(seed, (uint32_t)pixel_index[k], (uint32_t)sample_index[k]);
        const Ray r = get_ray(*cam, pixel_index[k] % cam->image_width, pixel_index[k] / cam->image_width, &rng, &cnt);
        out[k].origin[0] = r.origin.x; out[k].origin[1] = r.origin.y; out[k].origin[2] = r.origin.z;
        out[k].direction[0] = r.direction.x; out[k].direction[1] = r.direction.y; out[k].direction[2] = r.direction.z;
        out[k].time = r.time;
    }
    return 0;
}

// One material scatter for a given hit record (keyed draws at segment `seg` of path (pixel, sample)).
int oracle_scatter(const rt_scene_desc* scene, const rt_ray_desc* ray_in, const rt_hit_desc* hit, uint64_t seed,
                   uint32_t pixel, uint32_t sample, uint32_t seg, int mode, rt_ray_desc* scattered,
                   double* attenuation, double* emission) {
    if (!scene || !ray_in || !hit) return -1;
    Counters cnt;
    std::memset(&cnt, 0, sizeof(cnt));
    Sampler rng;
    rng.mode = mode;
    rng.begin_path(seed, pixel, sample);
    rng.seg = seg;
    rng.seq.seed(seed ^ ((uint64_t)pixel << 32) ^ sample);
    HitRecord h;
    h.p = Vec3(hit->p); h.normal = Vec3(hit->normal); h.mat = hit->mat_id; h.t = hit->t; h.u = hit->u; h.v = hit->v;
    h.front_face = hit->front_face != 0; h.prim = hit->prim_id;
    const Ray r{Vec3(ray_in->origin), Vec3(ray_in->direction), ray_in->time};
    const Vec3 e = emitted(scene, h.mat, h.u, h.v, h.p, &cnt);
    if (emission) { emission[0] = e.x; emission[1] = e.y; emission[2] = e.z; }
    Ray s;
    Vec3 a;
    const bool ok = scatter(scene, r, h, &rng, &cnt, &s, &a);
    if (ok && scattered) {
        scattered->origin[0] = s.origin.x; scattered->origin[1] = s.origin.y; scattered->origin[2] = s.origin.z;
        scattered->direction[0] = s.direction.x; scattered->direction[1] = s.direction.y; scattered->direction[2] = s.direction.z;
        scattered->time = s.time;
    }
    if (ok && attenuation) { attenuation[0] = a.x; attenuation[1] = a.y; attenuation[2] = a.z; }
    return ok ? 1 : 0;
}

// Camera::new (camera.rs:54-110), restated independently of the product's rt_camera_new.
int oracle_camera_new(const rt_camera_settings* s, rt_camera_desc* c) {
    if (!s || !c) return -1;
    std::memset(c, 0, sizeof(*c));
    const int64_t image_height = (int64_t)((FP)s->image_width / s->aspect_ratio);
    const FP theta = s->vfov * PI / 180.0;
    const FP h = std::tan(theta / 2.0);
    const FP viewport_height = 2.0 * h * s->focus_dist;
    const FP viewport_width = viewport_height * ((FP)s->image_width / (FP)image_height);
    const Vec3 look_from(s->look_from), look_at(s->look_at), vup(s->vup);
    const Vec3 w = normalize(look_from - look_at);
    const Vec3 u = normalize(cross(vup, w));
    const Vec3 v = cross(w, u);
    const Vec3 viewport_u = viewport_width * u;
    const Vec3 viewport_v = -viewport_height * v;
    const Vec3 center = look_from;
    const Vec3 pixel_delta_u = viewport_u / (FP)s->image_width;
    const Vec3 pixel_delta_v = viewport_v / (FP)image_height;
    const Vec3 viewport_upper_left = center - s->focus_dist * w - viewport_u * 0.5 - viewport_v * 0.5;
    const Vec3 pixel00_loc = viewport_upper_left + 0.5 * (pixel_delta_u + pixel_delta_v);
    const FP defocus_radius = s->focus_dist * std::tan((s->defocus_angle / 2.0) * PI / 180.0);
    const Vec3 ddu = u * defocus_radius, ddv = v * defocus_radius;
    c->image_width = s->image_width;
    c->image_height = image_height;
    c->samples_per_pixel = s->samples_per_pixel;
    c->max_depth = s->max_depth;
    for (int k = 0; k < 3; ++k) c->background[k] = s->background[k];
    auto st = [](Vec3 a, double* o) { o[0] = a.x; o[1] = a.y; o[2] = a.z; };
    st(center, c->center); st(pixel00_loc, c->pixel00_loc); st(pixel_delta_u, c->pixel_delta_u);
    st(pixel_delta_v, c->pixel_delta_v); st(ddu, c->defocus_disk_u); st(ddv, c->defocus_disk_v);
    c->defocus_angle = s->defocus_angle;
    return 0;
}

// BVHNode::node_from_list on bare boxes. axes = the axis draws in call order. Outputs sized 2n-1.
int oracle_bvh_build(const double* boxes, int n, const int* axes, int n_axes, int32_t* left, int32_t* right,
                     int32_t* object, double* node_bbox, int32_t* n_nodes, int32_t* axes_used) {
    if (!boxes || n <= 0 || !axes) return -1;
    std::vector<int32_t> objs(n);
    for (int i = 0; i < n; ++i) objs[i] = i;
    BvhOut out;
    int pos = 0;
    if (bvh_build_rec(boxes, objs, 0, n, axes, n_axes, &pos, &out) < 0) return -2;
    const int m = (int)out.object.size();
    for (int i = 0; i < m; ++i) { left[i] = out.left[i]; right[i] = out.right[i]; object[i] = out.object[i]; }
    std::memcpy(node_bbox, out.bbox.data(), (size_t)m * 6 * sizeof(double));
    *n_nodes = m;
    if (axes_used) *axes_used = pos;
    return 0;
}

// Recompute every hittable's bounding box (and the quad / rotate derived fields) from its own
// parameters with the reference's formulas; report the largest absolute deviation from the
// description. 0.0 means bit-identical.
int oracle_validate_scene(const rt_scene_desc* d, double* max_bbox_diff, double* max_derived_diff) {
    if (!d) return -1;
    std::vector<double> bb((size_t)d->n_hittables * 6, 0.0);
    double worst_b = 0.0, worst_d = 0.0;
    auto from_points = [](Vec3 a, Vec3 b, double* o) {
        o[0] = std::fmin(a.x, b.x); o[1] = std::fmax(a.x, b.x); o[2] = std::fmin(a.y, b.y);
        o[3] = std::fmax(a.y, b.y); o[4] = std::fmin(a.z, b.z); o[5] = std::fmax(a.z, b.z);
    };
    auto unite = [](const double* a, const double* b, double* o) {
        for (int c = 0; c < 3; ++c) { o[2 * c] = std::fmin(a[2 * c], b[2 * c]); o[2 * c + 1] = std::fmax(a[2 * c + 1], b[2 * c + 1]); }
    };
    for (int id = 0; id < d->n_hittables; ++id) {
        const rt_hittable_desc& h = d->hittables[id];
        double* o = &bb[(size_t)id * 6];
        switch (h.kind) {
            case RT_HIT_SPHERE: {  // sphere.rs:23-46
                const Vec3 c(h.v0), rvec(h.s0, h.s0, h.s0);
                from_points(c - rvec, c + rvec, o);
                if (h.flags & RT_FLAG_MOVING) {
                    const Vec3 target = c + Vec3(h.v1);  // center_vec = target - center; exact when representable
                    double b2[6], u[6];
                    from_points(target - rvec, target + rvec, b2);
                    unite(o, b2, u);
                    std::memcpy(o, u, sizeof(u));
                }
                break;
            }
            case RT_HIT_QUAD: {  // quad.rs:23-43
                const Vec3 q(h.v0), u(h.v1), v(h.v2);
                const Vec3 n = cross(u, v);
                const Vec3 normal = normalize(n);
                const FP dd = dot(normal, q);
                const Vec3 w = n / length_squared(n);
                worst_d = std::fmax(worst_d, std::fabs(dd - h.s0));
                for (int k = 0; k < 3; ++k) {
                    worst_d = std::fmax(worst_d, std::fabs(normal[k] - h.n[k]));
                    worst_d = std::fmax(worst_d, std::fabs(w[k] - h.v3[k]));
                }
                from_points(q, q + u + v, o);
                for (int c = 0; c < 3; ++c)  // AABB::pad, aabb.rs:35-53
                    if (o[2 * c + 1] - o[2 * c] < 0.0001) { o[2 * c] = o[2 * c] - 0.0001 * 0.5; o[2 * c + 1] = o[2 * c + 1] + 0.0001 * 0.5; }
                break;
            }
            case RT_HIT_LIST: {  // hittable.rs:50-59 (Default bbox = [0,0]^3)
                for (int i = 0; i < h.count; ++i) {
                    double u[6];
                    unite(o, &bb[(size_t)d->list_items[h.child + i] * 6], u);
                    std::memcpy(o, u, sizeof(u));
                }
                break;
            }
            case RT_HIT_TRANSLATE: {  // hittable.rs:87-94
                const double* cb = &bb[(size_t)h.child * 6];
                for (int c = 0; c < 3; ++c) { o[2 * c] = cb[2 * c] + h.v0[c]; o[2 * c + 1] = cb[2 * c + 1] + h.v0[c]; }
                break;
            }
            case RT_HIT_ROTATE_Y: {  // hittable.rs:120-157
                const double* cb = &bb[(size_t)h.child * 6];
                const FP sin_theta = h.s0, cos_theta = h.s1;
                Vec3 mn(INF, INF, INF), mx(-INF, -INF, -INF);
                for (int i = 0; i < 2; ++i) for (int j = 0; j < 2; ++j) for (int k = 0; k < 2; ++k) {
                    const FP x = (FP)i * cb[1] + (1.0 - (FP)i) * cb[0];
                    const FP y = (FP)j * cb[3] + (1.0 - (FP)j) * cb[2];
                    const FP z = (FP)k * cb[5] + (1.0 - (FP)k) * cb[4];
                    const Vec3 t(cos_theta * x + sin_theta * z, y, -sin_theta * x + cos_theta * z);
                    mn = Vec3(std::fmin(mn.x, t.x), std::fmin(mn.y, t.y), std::fmin(mn.z, t.z));
                    mx = Vec3(std::fmax(mx.x, t.x), std::fmax(mx.y, t.y), std::fmax(mx.z, t.z));
                }
                from_points(mn, mx, o);
                worst_d = std::fmax(worst_d, std::fabs(sin_theta * sin_theta + cos_theta * cos_theta - 1.0));
                break;
            }
            case RT_HIT_CONSTANT_MEDIUM:  // constant_medium.rs:73-75
                std::memcpy(o, &bb[(size_t)h.child * 6], 6 * sizeof(double));
                break;
            case RT_HIT_BVH:  // bvh.rs:120-122: union of all leaves == root box
                std::memcpy(o, d->bvh_nodes[h.child].bbox, 6 * sizeof(double));
                break;
        }
        for (int k = 0; k < 6; ++k) worst_b = std::fmax(worst_b, std::fabs(o[k] - h.bbox[k]));
    }
    // every BVH node: leaf box == object's box, branch box == union of children
    for (int i = 0; i < d->n_bvh_nodes; ++i) {
        const rt_bvh_node_desc& n = d->bvh_nodes[i];
        double e[6];
        if (n.object >= 0) std::memcpy(e, &bb[(size_t)n.object * 6], sizeof(e));
        else unite(d->bvh_nodes[n.left].bbox, d->bvh_nodes[n.right].bbox, e);
        for (int k = 0; k < 6; ++k) worst_b = std::fmax(worst_b, std::fabs(e[k] - n.bbox[k]));
    }
    if (max_bbox_diff) *max_bbox_diff = worst_b;
    if (max_derived_diff) *max_derived_diff = worst_d;
    return 0;
}

// ---- scalar known-answer helpers (tests/test_oracle_kat.py) ----
void oracle_sphere_uv(const double n[3], double* u, double* v) { sphere_uv(Vec3(n), u, v); }
double oracle_reflectance(double cosine, double ref_idx) { return reflectance(cosine, ref_idx); }
void oracle_refract(const double uv[3], const double n[3], double eta, double out[3]) {
    const Vec3 r = refract(Vec3(uv), Vec3(n), eta);
    out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
void oracle_reflect(const double v[3], const double n[3], double out[3]) {
    const Vec3 r = reflect(Vec3(v), Vec3(n));
    out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
int oracle_aabb_hit(const double box[6], const rt_ray_desc* r, double t_min, double t_max) {
    return aabb_hit(box, Ray{Vec3(r->origin), Vec3(r->direction), r->time}, Interval{t_min, t_max}) ? 1 : 0;
}
void oracle_rgb_to_color(uint8_t r, uint8_t g, uint8_t b, double out[3]) {  // color.rs:21-27
    out[0] = gamma_to_linear((FP)r / 255.0); out[1] = gamma_to_linear((FP)g / 255.0); out[2] = gamma_to_linear((FP)b / 255.0);
}
void oracle_color_to_rgb(const double rgb[3], uint8_t out[3]) {  // color.rs:12-19
    for (int k = 0; k < 3; ++k) {
        const FP g = linear_to_gamma(rgb[k]);
        const FP x = 256.0 * clampf(g, 0.0, 0.999);
        out[k] = (x != x) ? 0 : (uint8_t)x;  // Rust `as u8`: NaN -> 0, saturating
    }
}
// renderer.rs:55-58: color_to_rgb(sum / spp) per pixel.
void oracle_finalize_rgb8(const double* sum_rgb, int64_t n_pixels, double spp, uint8_t* out) {
    for (int64_t i = 0; i < n_pixels; ++i) {
        const Vec3 c = Vec3(sum_rgb + i * 3) / spp;
        const double v[3] = {c.x, c.y, c.z};
        oracle_color_to_rgb(v, out + i * 3);
    }
}
double oracle_perlin_noise(const rt_perlin_desc* pn, const double p[3]) { return perlin_noise(*pn, Vec3(p)); }
double oracle_perlin_turbulence(const rt_perlin_desc* pn, const double p[3], int depth) { return perlin_turbulence(*pn, Vec3(p), depth); }
// Samplers, for distribution tests: kind 0 = unit vector, 1 = in unit sphere, 2 = in unit disk (z=0).
void oracle_sample(int kind, int mode, uint64_t seed, int64_t n, double* out) {
    Sampler rng;
    rng.mode = mode;
    rng.seq.seed(seed);
    for (int64_t k = 0; k < n; ++k) {
        rng.begin_path(seed, (uint32_t)k, (uint32_t)(k >> 32));
        Vec3 v;
        if (kind == 0) v = rng.random_unit_vector();
        else if (kind == 1) v = rng.random_in_unit_sphere();
        else { FP px, py, tm, dx = 0, dy = 0; rng.camera(&px, &py, &tm, true, &dx, &dy); v = Vec3(dx, dy, 0.0); }
        out[k * 3] = v.x; out[k * 3 + 1] = v.y; out[k * 3 + 2] = v.z;
    }
}
void oracle_pcg4d(const uint32_t in[4], uint32_t out[4]) {
    const U4 r = pcg4d(U4{in[0], in[1], in[2], in[3]});
    out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
}

}  // extern "C"
