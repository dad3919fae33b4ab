// Host side of the drop-in boundary: the crate's constructors re-expressed as a flat,
// id-based scene description (include/rt_b200.h). Nothing here touches the GPU.
//
// Every bounding box and derived field is computed with the reference's own formulas so
// that the median-split BVH (bvh.rs:31-66), which sorts on bbox minima, gets the same
// topology the crate would build from the same objects and axis draws.
#include "../../../include/rt_b200.h"
#include "host_common.h"
#include "host_rng.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <memory>
#include <string>
#include <vector>

namespace rt_host {

thread_local std::string g_last_error;

int fail(int code, const std::string& msg) {
    g_last_error = msg;
    return code;
}

}  // namespace rt_host

using rt_host::fail;

struct rt_builder {
    std::vector<rt_texture_desc> textures;
    std::vector<rt_material_desc> materials;
    std::vector<rt_hittable_desc> hittables;
    std::vector<int32_t> list_items;
    std::vector<rt_bvh_node_desc> bvh_nodes;
    std::vector<rt_perlin_desc> perlins;
    std::vector<rt_image_desc> images;
    std::vector<std::unique_ptr<std::vector<uint8_t>>> image_data;
    rt_host::HostRng axis_rng;
    explicit rt_builder(uint64_t bvh_seed) : axis_rng(bvh_seed) {}
};

namespace {

const double kPi = 3.14159265358979323846;  // common.rs:3

// ---- Interval / AABB arithmetic (interval.rs, aabb.rs) on the packed bbox[6] ----

// AABB::new_from_points (aabb.rs:20-26)
void bbox_from_points(const double a[3], const double b[3], double out[6]) {
    for (int c = 0; c < 3; ++c) {
        out[2 * c] = std::fmin(a[c], b[c]);
        out[2 * c + 1] = std::fmax(a[c], b[c]);
    }
}

// AABB::new_from_aabbs (aabb.rs:27-33) via Interval::new_from_intervals (interval.rs:23-28)
void bbox_union(const double a[6], const double b[6], double out[6]) {
    for (int c = 0; c < 3; ++c) {
        out[2 * c] = std::fmin(a[2 * c], b[2 * c]);
        out[2 * c + 1] = std::fmax(a[2 * c + 1], b[2 * c + 1]);
    }
}

// AABB::pad (aabb.rs:35-53) with Interval::expand (interval.rs:29-34)
void bbox_pad(double box[6]) {
    const double delta = 0.0001;
    for (int c = 0; c < 3; ++c) {
        if (box[2 * c + 1] - box[2 * c] < delta) {
            box[2 * c] = box[2 * c] - delta * 0.5;
            box[2 * c + 1] = box[2 * c + 1] + delta * 0.5;
        }
    }
}

bool valid_id(int id, size_t n) { return id >= 0 && (size_t)id < n; }

rt_hittable_desc blank_hittable(int kind) {
    rt_hittable_desc h;
    std::memset(&h, 0, sizeof(h));
    h.kind = kind;
    h.mat = -1;
    h.child = -1;
    return h;
}

// Perlin::new (perlin.rs:16-25): 256 Vec3::random_range(-1,1) draws (x, y, z order,
// vec3.rs:46-52), then three Fisher-Yates permutations (perlin.rs:66-79).
void perlin_new(uint64_t seed, rt_perlin_desc* out) {
    rt_host::HostRng rng(seed);
    for (int i = 0; i < 256; ++i)
        for (int c = 0; c < 3; ++c) out->ranvec[i][c] = rng.range(-1.0, 1.0);
    int32_t* perms[3] = {out->perm_x, out->perm_y, out->perm_z};
    for (int k = 0; k < 3; ++k) {
        int32_t* p = perms[k];
        for (int i = 0; i < 256; ++i) p[i] = i;
        for (int i = 255; i >= 1; --i) {
            const int target = rng.range_inclusive(0, i);
            std::swap(p[i], p[target]);
        }
    }
}

// BVHNode::node_from_list (bvh.rs:31-66). `objs` is the slice being split (sorted in place,
// as the reference sorts `objects`), nodes are appended in pre-order. Returns the node index.
int32_t bvh_node_from_list(rt_builder* b, int32_t* objs, int span) {
    const int axis = b->axis_rng.range_inclusive(0, 2);  // drawn before the span test (bvh.rs:32)
    const int32_t me = (int32_t)b->bvh_nodes.size();
    b->bvh_nodes.emplace_back();
    auto min_of = [&](int32_t id) { return b->hittables[id].bbox[2 * axis]; };
    auto make_leaf = [&](int32_t id, int ax) {
        rt_bvh_node_desc n;
        std::memcpy(n.bbox, b->hittables[id].bbox, sizeof(n.bbox));
        n.left = n.right = -1;
        n.object = id;
        n.axis = ax;
        return n;
    };
    if (span == 1) {
        b->bvh_nodes[me] = make_leaf(objs[0], axis);
    } else if (span == 2) {
        int32_t left = objs[0], right = objs[1];
        // comparator(left,right) != Less  <=>  !(left.min < right.min)  (bvh.rs:48,68-74)
        if (!(min_of(left) < min_of(right))) std::swap(left, right);
        const int32_t li = (int32_t)b->bvh_nodes.size();
        b->bvh_nodes.push_back(make_leaf(left, -1));   // built directly: no axis draw (bvh.rs:51-56)
        const int32_t ri = (int32_t)b->bvh_nodes.size();
        b->bvh_nodes.push_back(make_leaf(right, -1));
        rt_bvh_node_desc n;
        bbox_union(b->hittables[left].bbox, b->hittables[right].bbox, n.bbox);
        n.left = li;
        n.right = ri;
        n.object = -1;
        n.axis = axis;
        b->bvh_nodes[me] = n;
    } else {
        // sort_unstable_by with a comparator that never returns Equal (bvh.rs:59,68-74): the order
        // of ties is unspecified in the reference; here ties keep insertion order (stable, strict <).
        std::stable_sort(objs, objs + span, [&](int32_t x, int32_t y) { return min_of(x) < min_of(y); });
        const int half = span / 2;
        const int32_t li = bvh_node_from_list(b, objs, half);
        const int32_t ri = bvh_node_from_list(b, objs + half, span - half);
        rt_bvh_node_desc n;
        bbox_union(b->bvh_nodes[li].bbox, b->bvh_nodes[ri].bbox, n.bbox);
        n.left = li;
        n.right = ri;
        n.object = -1;
        n.axis = axis;
        b->bvh_nodes[me] = n;
    }
    return me;
}

}  // namespace

extern "C" {

const char* rt_last_error(void) { return rt_host::g_last_error.c_str(); }
int rt_abi_version(void) { return RT_B200_ABI_VERSION; }

int rt_builder_create(uint64_t bvh_seed, rt_builder** out) {
    if (!out) return fail(RT_ERR_INVALID_ARGUMENT, "rt_builder_create: out is null");
    *out = new (std::nothrow) rt_builder(bvh_seed);
    if (!*out) return fail(RT_ERR_OUT_OF_MEMORY, "rt_builder_create: allocation failed");
    return RT_OK;
}

void rt_builder_destroy(rt_builder* b) { delete b; }

// ------------------------------------------------------------------ textures
int rt_tex_solid(rt_builder* b, double r, double g, double bl) {
    if (!b) return fail(RT_ERR_INVALID_ARGUMENT, "rt_tex_solid: builder is null");
    rt_texture_desc t{};
    t.kind = RT_TEX_SOLID;
    t.a = t.b = -1;
    t.color[0] = r; t.color[1] = g; t.color[2] = bl;
    b->textures.push_back(t);
    return (int)b->textures.size() - 1;
}

int rt_tex_checker(rt_builder* b, double scale, int even_tex, int odd_tex) {
    return rt_tex_checker_inv(b, 1.0 / scale, even_tex, odd_tex);  // texture.rs:46
}

// The *_inv / *_sincos / *_nid constructors take the object's STORED state instead of the constructor's argument: a host
// that flattens objects it already built must hand over inv_scale, sin / cos, neg_inv_density bit for bit,
// not values that round-trip through 1/x, atan2 or -1/x.
int rt_tex_checker_inv(rt_builder* b, double inv_scale, int even_tex, int odd_tex) {
    if (!b) return fail(RT_ERR_INVALID_ARGUMENT, "rt_tex_checker: builder is null");
    if (!valid_id(even_tex, b->textures.size()) || !valid_id(odd_tex, b->textures.size()))
        return fail(RT_ERR_OUT_OF_RANGE, "rt_tex_checker: unknown texture id");
    rt_texture_desc t{};
    t.kind = RT_TEX_CHECKER;
    t.a = even_tex;
    t.b = odd_tex;
    t.scale = inv_scale;
    b->textures.push_back(t);
    return (int)b->textures.size() - 1;
}

int rt_tex_image(rt_builder* b, int width, int height, const uint8_t* rgb8) {
    if (!b || !rgb8) return fail(RT_ERR_INVALID_ARGUMENT, "rt_tex_image: null argument");
    if (width <= 0 || height <= 0) return fail(RT_ERR_INVALID_ARGUMENT, "rt_tex_image: empty image");
    auto data = std::make_unique<std::vector<uint8_t>>(rgb8, rgb8 + (size_t)width * height * 3);
    rt_image_desc im{};
    im.width = width;
    im.height = height;
    im.rgb8 = data->data();
    b->image_data.push_back(std::move(data));
    b->images.push_back(im);
    rt_texture_desc t{};
    t.kind = RT_TEX_IMAGE;
    t.a = (int)b->images.size() - 1;
    t.b = -1;
    b->textures.push_back(t);
    return (int)b->textures.size() - 1;
}

int rt_tex_noise(rt_builder* b, double scale, uint64_t perlin_seed) {
    if (!b) return fail(RT_ERR_INVALID_ARGUMENT, "rt_tex_noise: builder is null");
    b->perlins.emplace_back();
    perlin_new(perlin_seed, &b->perlins.back());
    rt_texture_desc t{};
    t.kind = RT_TEX_NOISE;
    t.a = (int)b->perlins.size() - 1;
    t.b = -1;
    t.scale = scale;
    b->textures.push_back(t);
    return (int)b->textures.size() - 1;
}

// NoiseTexture with the tables the HOST drew (the crate's Perlin::new uses its unseeded thread_rng, perlin.rs:16-25): a
// drop-in must shade with exactly those. ranvec: 256 x 3, perm_*: 256 entries in [0, 256).
int rt_tex_noise_tables(rt_builder* b, double scale, const double* ranvec, const int32_t* perm_x, const int32_t* perm_y,
                        const int32_t* perm_z) {
    if (!b || !ranvec || !perm_x || !perm_y || !perm_z) return fail(RT_ERR_INVALID_ARGUMENT, "rt_tex_noise_tables: null argument");
    for (int k = 0; k < 256; ++k)
        if ((uint32_t)perm_x[k] > 255u || (uint32_t)perm_y[k] > 255u || (uint32_t)perm_z[k] > 255u)
            return fail(RT_ERR_OUT_OF_RANGE, "rt_tex_noise_tables: permutation entries must lie in [0, 256)");
    b->perlins.emplace_back();
    rt_perlin_desc& p = b->perlins.back();
    for (int k = 0; k < 256; ++k) {
        for (int c = 0; c < 3; ++c) p.ranvec[k][c] = ranvec[3 * k + c];
        p.perm_x[k] = perm_x[k]; p.perm_y[k] = perm_y[k]; p.perm_z[k] = perm_z[k];
    }
    rt_texture_desc t{};
    t.kind = RT_TEX_NOISE;
    t.a = (int)b->perlins.size() - 1;
    t.b = -1;
    t.scale = scale;
    b->textures.push_back(t);
    return (int)b->textures.size() - 1;
}

// ----------------------------------------------------------------- materials
static int push_material(rt_builder* b, int kind, int tex, const double albedo[3], double param) {
    rt_material_desc m{};
    m.kind = kind;
    m.tex = tex;
    if (albedo) { m.albedo[0] = albedo[0]; m.albedo[1] = albedo[1]; m.albedo[2] = albedo[2]; }
    m.param = param;
    b->materials.push_back(m);
    return (int)b->materials.size() - 1;
}

int rt_mat_lambertian(rt_builder* b, int tex) {
    if (!b) return fail(RT_ERR_INVALID_ARGUMENT, "rt_mat_lambertian: builder is null");
    if (!valid_id(tex, b->textures.size())) return fail(RT_ERR_OUT_OF_RANGE, "rt_mat_lambertian: unknown texture id");
    return push_material(b, RT_MAT_LAMBERTIAN, tex, nullptr, 0.0);
}
int rt_mat_metal(rt_builder* b, const double albedo[3], double fuzz) {
    if (!b || !albedo) return fail(RT_ERR_INVALID_ARGUMENT, "rt_mat_metal: null argument");
    return push_material(b, RT_MAT_METAL, -1, albedo, fuzz);  // fuzz is not clamped (material.rs:49-51)
}
int rt_mat_dielectric(rt_builder* b, double ir) {
    if (!b) return fail(RT_ERR_INVALID_ARGUMENT, "rt_mat_dielectric: builder is null");
    return push_material(b, RT_MAT_DIELECTRIC, -1, nullptr, ir);
}
int rt_mat_diffuse_light(rt_builder* b, int tex) {
    if (!b) return fail(RT_ERR_INVALID_ARGUMENT, "rt_mat_diffuse_light: builder is null");
    if (!valid_id(tex, b->textures.size())) return fail(RT_ERR_OUT_OF_RANGE, "rt_mat_diffuse_light: unknown texture id");
    return push_material(b, RT_MAT_DIFFUSE_LIGHT, tex, nullptr, 0.0);
}
int rt_mat_isotropic(rt_builder* b, int tex) {
    if (!b) return fail(RT_ERR_INVALID_ARGUMENT, "rt_mat_isotropic: builder is null");
    if (!valid_id(tex, b->textures.size())) return fail(RT_ERR_OUT_OF_RANGE, "rt_mat_isotropic: unknown texture id");
    return push_material(b, RT_MAT_ISOTROPIC, tex, nullptr, 0.0);
}

// ----------------------------------------------------------------- hittables
int rt_hit_sphere(rt_builder* b, const double c[3], double radius, int mat) {
    if (!b || !c) return fail(RT_ERR_INVALID_ARGUMENT, "rt_hit_sphere: null argument");
    if (!valid_id(mat, b->materials.size())) return fail(RT_ERR_OUT_OF_RANGE, "rt_hit_sphere: unknown material id");
    rt_hittable_desc h = blank_hittable(RT_HIT_SPHERE);
    h.mat = mat;
    h.s0 = radius;
    double lo[3], hi[3];
    for (int k = 0; k < 3; ++k) { h.v0[k] = c[k]; lo[k] = c[k] - radius; hi[k] = c[k] + radius; }
    bbox_from_points(lo, hi, h.bbox);  // sphere.rs:24-32
    b->hittables.push_back(h);
    return (int)b->hittables.size() - 1;
}

int rt_hit_moving_sphere(rt_builder* b, const double c[3], const double target[3], double radius, int mat) {
    if (!b || !c || !target) return fail(RT_ERR_INVALID_ARGUMENT, "rt_hit_moving_sphere: null argument");
    if (!valid_id(mat, b->materials.size())) return fail(RT_ERR_OUT_OF_RANGE, "rt_hit_moving_sphere: unknown material id");
    rt_hittable_desc h = blank_hittable(RT_HIT_SPHERE);
    h.mat = mat;
    h.s0 = radius;
    h.flags = RT_FLAG_MOVING;
    double lo1[3], hi1[3], lo2[3], hi2[3], b1[6], b2[6];
    for (int k = 0; k < 3; ++k) {
        h.v0[k] = c[k];
        h.v1[k] = target[k] - c[k];  // center_vec (sphere.rs:42)
        lo1[k] = c[k] - radius; hi1[k] = c[k] + radius;
        lo2[k] = target[k] - radius; hi2[k] = target[k] + radius;
    }
    bbox_from_points(lo1, hi1, b1);
    bbox_from_points(lo2, hi2, b2);
    bbox_union(b1, b2, h.bbox);  // sphere.rs:36-44
    b->hittables.push_back(h);
    return (int)b->hittables.size() - 1;
}

int rt_hit_quad(rt_builder* b, const double q[3], const double u[3], const double v[3], int mat) {
    if (!b || !q || !u || !v) return fail(RT_ERR_INVALID_ARGUMENT, "rt_hit_quad: null argument");
    if (!valid_id(mat, b->materials.size())) return fail(RT_ERR_OUT_OF_RANGE, "rt_hit_quad: unknown material id");
    rt_hittable_desc h = blank_hittable(RT_HIT_QUAD);
    h.mat = mat;
    // quad.rs:23-39
    const double n[3] = {u[1] * v[2] - u[2] * v[1], u[2] * v[0] - u[0] * v[2], u[0] * v[1] - u[1] * v[0]};
    const double nn = n[0] * n[0] + n[1] * n[1] + n[2] * n[2];
    const double len_recip = 1.0 / std::sqrt(nn);  // vec3.rs:119-131: normalize = v * length().recip()
    const double inv_nn = 1.0 / nn;                // vec3.rs:244-249: Vec3 / FP = v * (1.0 / s)
    double corner[3];
    for (int k = 0; k < 3; ++k) {
        h.v0[k] = q[k]; h.v1[k] = u[k]; h.v2[k] = v[k];
        h.n[k] = n[k] * len_recip;
        h.v3[k] = n[k] * inv_nn;
        corner[k] = q[k] + u[k] + v[k];
    }
    h.s0 = h.n[0] * q[0] + h.n[1] * q[1] + h.n[2] * q[2];
    bbox_from_points(q, corner, h.bbox);  // quad.rs:41-43
    bbox_pad(h.bbox);
    b->hittables.push_back(h);
    return (int)b->hittables.size() - 1;
}

int rt_hit_list(rt_builder* b, const int* ids, int n) {
    if (!b || (n > 0 && !ids)) return fail(RT_ERR_INVALID_ARGUMENT, "rt_hit_list: null argument");
    if (n < 0) return fail(RT_ERR_INVALID_ARGUMENT, "rt_hit_list: negative count");
    rt_hittable_desc h = blank_hittable(RT_HIT_LIST);
    h.child = (int32_t)b->list_items.size();
    h.count = n;
    // HittableList derives Default, so its bbox starts as three [0,0] intervals and every add()
    // unions into that (hittable.rs:50-59, interval.rs:5, aabb.rs:9): the origin is always inside.
    for (int i = 0; i < n; ++i) {
        if (!valid_id(ids[i], b->hittables.size())) return fail(RT_ERR_OUT_OF_RANGE, "rt_hit_list: unknown hittable id");
        b->list_items.push_back(ids[i]);
        double u[6];
        bbox_union(h.bbox, b->hittables[ids[i]].bbox, u);
        std::memcpy(h.bbox, u, sizeof(u));
    }
    // A host that flattens its own objects cannot say "this list came from Quad::cube": recognise it. Six
    // consecutive quads of one material that are, bit for bit, what quad.rs:45-93 makes of some (min, max) pair.
    if (n == 6) {
        bool cube = true;
        for (int k = 0; k < 6 && cube; ++k) {
            const rt_hittable_desc& q = b->hittables[ids[k]];
            cube = ids[k] == ids[0] + k && q.kind == RT_HIT_QUAD && q.mat == b->hittables[ids[0]].mat;
        }
        if (cube) {
            const rt_hittable_desc* q = &b->hittables[ids[0]];
            const double mn[3] = {q[3].v0[0], q[3].v0[1], q[3].v0[2]}, mx[3] = {q[1].v0[0], q[4].v0[1], q[0].v0[2]};
            const double dx[3] = {mx[0] - mn[0], 0.0, 0.0}, dy[3] = {0.0, mx[1] - mn[1], 0.0}, dz[3] = {0.0, 0.0, mx[2] - mn[2]};
            const double ndx[3] = {-dx[0], -0.0, -0.0}, ndz[3] = {-0.0, -0.0, -dz[2]};
            const double q0[3] = {mn[0], mn[1], mx[2]}, q1[3] = {mx[0], mn[1], mx[2]}, q2[3] = {mx[0], mn[1], mn[2]},
                         q3[3] = {mn[0], mn[1], mn[2]}, q4[3] = {mn[0], mx[1], mx[2]}, q5[3] = {mn[0], mn[1], mn[2]};
            const double* want[6][3] = {{q0, dx, dy}, {q1, ndz, dy}, {q2, ndx, dy}, {q3, dz, dy}, {q4, dx, ndz}, {q5, dx, dz}};
            for (int k = 0; k < 6 && cube; ++k)
                for (int c = 0; c < 3 && cube; ++c)   // == treats -0.0 and 0.0 alike, as the quads' arithmetic does
                    cube = q[k].v0[c] == want[k][0][c] && q[k].v1[c] == want[k][1][c] && q[k].v2[c] == want[k][2][c];
            if (cube) {
                h.flags |= RT_FLAG_CUBE_LIST;
                for (int k = 0; k < 3; ++k) { h.v0[k] = mn[k]; h.v1[k] = mx[k]; }
            }
        }
    }
    b->hittables.push_back(h);
    return (int)b->hittables.size() - 1;
}

int rt_hit_cube(rt_builder* b, const double a[3], const double bb[3], int mat) {
    if (!b || !a || !bb) return fail(RT_ERR_INVALID_ARGUMENT, "rt_hit_cube: null argument");
    if (!valid_id(mat, b->materials.size())) return fail(RT_ERR_OUT_OF_RANGE, "rt_hit_cube: unknown material id");
    // quad.rs:45-93
    double mn[3], mx[3];
    for (int k = 0; k < 3; ++k) { mn[k] = std::fmin(a[k], bb[k]); mx[k] = std::fmax(a[k], bb[k]); }
    const double dx[3] = {mx[0] - mn[0], 0.0, 0.0}, dy[3] = {0.0, mx[1] - mn[1], 0.0}, dz[3] = {0.0, 0.0, mx[2] - mn[2]};
    const double ndx[3] = {-dx[0], -0.0, -0.0}, ndz[3] = {-0.0, -0.0, -dz[2]};
    const double q0[3] = {mn[0], mn[1], mx[2]}, q1[3] = {mx[0], mn[1], mx[2]}, q2[3] = {mx[0], mn[1], mn[2]},
                 q3[3] = {mn[0], mn[1], mn[2]}, q4[3] = {mn[0], mx[1], mx[2]}, q5[3] = {mn[0], mn[1], mn[2]};
    int ids[6];
    ids[0] = rt_hit_quad(b, q0, dx, dy, mat);
    ids[1] = rt_hit_quad(b, q1, ndz, dy, mat);
    ids[2] = rt_hit_quad(b, q2, ndx, dy, mat);
    ids[3] = rt_hit_quad(b, q3, dz, dy, mat);
    ids[4] = rt_hit_quad(b, q4, dx, ndz, mat);
    ids[5] = rt_hit_quad(b, q5, dx, dz, mat);
    for (int i = 0; i < 6; ++i) if (ids[i] < 0) return ids[i];
    const int list = rt_hit_list(b, ids, 6);
    if (list < 0) return list;
    rt_hittable_desc& h = b->hittables[list];
    h.flags |= RT_FLAG_CUBE_LIST;
    for (int k = 0; k < 3; ++k) { h.v0[k] = mn[k]; h.v1[k] = mx[k]; }
    return list;
}

int rt_hit_translate(rt_builder* b, int object, const double offset[3]) {
    if (!b || !offset) return fail(RT_ERR_INVALID_ARGUMENT, "rt_hit_translate: null argument");
    if (!valid_id(object, b->hittables.size())) return fail(RT_ERR_OUT_OF_RANGE, "rt_hit_translate: unknown hittable id");
    rt_hittable_desc h = blank_hittable(RT_HIT_TRANSLATE);
    h.child = object;
    const double* cb = b->hittables[object].bbox;
    for (int k = 0; k < 3; ++k) {
        h.v0[k] = offset[k];
        h.bbox[2 * k] = cb[2 * k] + offset[k];  // aabb.rs:87-97, interval.rs:49-58
        h.bbox[2 * k + 1] = cb[2 * k + 1] + offset[k];
    }
    b->hittables.push_back(h);
    return (int)b->hittables.size() - 1;
}

int rt_hit_rotate_y(rt_builder* b, int object, double angle) {
    const double theta = angle * kPi / 180.0;  // common.rs:6-8
    return rt_hit_rotate_y_sincos(b, object, std::sin(theta), std::cos(theta));
}

int rt_hit_rotate_y_sincos(rt_builder* b, int object, double sin_theta, double cos_theta) {
    if (!b) return fail(RT_ERR_INVALID_ARGUMENT, "rt_hit_rotate_y: builder is null");
    if (!valid_id(object, b->hittables.size())) return fail(RT_ERR_OUT_OF_RANGE, "rt_hit_rotate_y: unknown hittable id");
    rt_hittable_desc h = blank_hittable(RT_HIT_ROTATE_Y);
    h.child = object;
    // hittable.rs:120-157
    h.s0 = sin_theta;
    h.s1 = cos_theta;
    const double* cb = b->hittables[object].bbox;
    const double inf = std::numeric_limits<double>::infinity();
    double mn[3] = {inf, inf, inf}, mx[3] = {-inf, -inf, -inf};
    for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 2; ++j)
            for (int k = 0; k < 2; ++k) {
                const double x = (double)i * cb[1] + (1.0 - (double)i) * cb[0];
                const double y = (double)j * cb[3] + (1.0 - (double)j) * cb[2];
                const double z = (double)k * cb[5] + (1.0 - (double)k) * cb[4];
                const double tester[3] = {cos_theta * x + sin_theta * z, y, -sin_theta * x + cos_theta * z};
                for (int c = 0; c < 3; ++c) {
                    mn[c] = std::fmin(mn[c], tester[c]);
                    mx[c] = std::fmax(mx[c], tester[c]);
                }
            }
    bbox_from_points(mn, mx, h.bbox);
    b->hittables.push_back(h);
    return (int)b->hittables.size() - 1;
}

int rt_hit_constant_medium(rt_builder* b, int boundary, double density, int tex) {
    return rt_hit_constant_medium_nid(b, boundary, -1.0 / density, tex);   // constant_medium.rs:24
}

int rt_hit_constant_medium_nid(rt_builder* b, int boundary, double neg_inv_density, int tex) {
    if (!b) return fail(RT_ERR_INVALID_ARGUMENT, "rt_hit_constant_medium: builder is null");
    if (!valid_id(boundary, b->hittables.size())) return fail(RT_ERR_OUT_OF_RANGE, "rt_hit_constant_medium: unknown hittable id");
    if (!valid_id(tex, b->textures.size())) return fail(RT_ERR_OUT_OF_RANGE, "rt_hit_constant_medium: unknown texture id");
    rt_hittable_desc h = blank_hittable(RT_HIT_CONSTANT_MEDIUM);
    h.child = boundary;
    h.s0 = neg_inv_density;
    h.mat = rt_mat_isotropic(b, tex);      // constant_medium.rs:25
    std::memcpy(h.bbox, b->hittables[boundary].bbox, sizeof(h.bbox));  // constant_medium.rs:73-75
    b->hittables.push_back(h);
    return (int)b->hittables.size() - 1;
}

int rt_hit_bvh(rt_builder* b, const int* ids, int n) {
    if (!b || !ids) return fail(RT_ERR_INVALID_ARGUMENT, "rt_hit_bvh: null argument");
    if (n <= 0) return fail(RT_ERR_INVALID_ARGUMENT, "rt_hit_bvh: a BVH needs at least one object");  // bvh.rs would index objects[0] of an empty slice
    std::vector<int32_t> objs(ids, ids + n);
    for (int32_t id : objs)
        if (!valid_id(id, b->hittables.size())) return fail(RT_ERR_OUT_OF_RANGE, "rt_hit_bvh: unknown hittable id");
    const int32_t first = (int32_t)b->bvh_nodes.size();
    const int32_t root = bvh_node_from_list(b, objs.data(), n);
    rt_hittable_desc h = blank_hittable(RT_HIT_BVH);
    h.child = root;
    h.count = (int32_t)b->bvh_nodes.size() - first;
    std::memcpy(h.bbox, b->bvh_nodes[root].bbox, sizeof(h.bbox));  // bvh.rs:120-122
    b->hittables.push_back(h);
    return (int)b->hittables.size() - 1;
}

// A BVH the HOST built (the crate's own BVHNode::node_from_list run, bvh.rs:31-66: its thread_rng axis draws, its sort):
// `nodes` in pre-order, nodes[0] the root, left / right as indices into `nodes` (children after their parent),
// object = the leaf's hittable id or -1, bbox = the node's own (bvh.rs:64-66). The device then walks exactly the tree
// the Rust side holds; rt_hit_bvh above builds one with the same rule from a seeded stream instead.
int rt_hit_bvh_nodes(rt_builder* b, const rt_bvh_node_desc* nodes, int n) {
    if (!b || !nodes) return fail(RT_ERR_INVALID_ARGUMENT, "rt_hit_bvh_nodes: null argument");
    if (n <= 0) return fail(RT_ERR_INVALID_ARGUMENT, "rt_hit_bvh_nodes: a BVH needs at least one node");
    std::vector<char> seen((size_t)n, 0);
    int leaves = 0;
    for (int i = 0; i < n; ++i) {
        const rt_bvh_node_desc& nd = nodes[i];
        if (nd.object >= 0) {
            if (!valid_id(nd.object, b->hittables.size())) return fail(RT_ERR_OUT_OF_RANGE, "rt_hit_bvh_nodes: leaf names an unknown hittable");
            ++leaves;
        } else {
            if (nd.left <= i || nd.right <= i || nd.left >= n || nd.right >= n || nd.left == nd.right)
                return fail(RT_ERR_OUT_OF_RANGE, "rt_hit_bvh_nodes: children must follow their parent (pre-order)");
            if (seen[nd.left]++ || seen[nd.right]++) return fail(RT_ERR_INVALID_ARGUMENT, "rt_hit_bvh_nodes: a node has two parents");
        }
    }
    for (int i = 1; i < n; ++i) if (!seen[i]) return fail(RT_ERR_INVALID_ARGUMENT, "rt_hit_bvh_nodes: unreachable node");
    if (n != 2 * leaves - 1) return fail(RT_ERR_INVALID_ARGUMENT, "rt_hit_bvh_nodes: not a full binary tree");
    const int32_t first = (int32_t)b->bvh_nodes.size();
    for (int i = 0; i < n; ++i) {
        rt_bvh_node_desc nd = nodes[i];
        if (nd.object >= 0) { nd.left = -1; nd.right = -1; }
        else { nd.left += first; nd.right += first; }
        b->bvh_nodes.push_back(nd);
    }
    rt_hittable_desc h = blank_hittable(RT_HIT_BVH);
    h.child = first;
    h.count = n;
    std::memcpy(h.bbox, nodes[0].bbox, sizeof(h.bbox));  // bvh.rs:120-122
    b->hittables.push_back(h);
    return (int)b->hittables.size() - 1;
}

// BVHNode::new_from_objects with the sorting done on the GPU (bvh_build.cuh): same seeded axis stream, same topology
// rule, same node array as rt_hit_bvh - bit for bit (tests/test_gpu_bvh_build.py).
extern "C" int rt_bvh_axis_draws(int n);
extern "C" int rt_bvh_build_device(rt_context* c, const double* bboxes, int n, const int32_t* axes, rt_bvh_node_desc* nodes_out,
                                   int32_t* order_out);
int rt_hit_bvh_device(rt_builder* b, rt_context* ctx, const int* ids, int n) {
    if (!b || !ids || !ctx) return fail(RT_ERR_INVALID_ARGUMENT, "rt_hit_bvh_device: null argument");
    if (n <= 0) return fail(RT_ERR_INVALID_ARGUMENT, "rt_hit_bvh_device: a BVH needs at least one object");
    std::vector<double> boxes((size_t)n * 6);
    for (int i = 0; i < n; ++i) {
        if (!valid_id(ids[i], b->hittables.size())) return fail(RT_ERR_OUT_OF_RANGE, "rt_hit_bvh_device: unknown hittable id");
        std::memcpy(&boxes[(size_t)i * 6], b->hittables[ids[i]].bbox, 48);
    }
    std::vector<int32_t> axes((size_t)rt_bvh_axis_draws(n));
    for (int32_t& a : axes) a = b->axis_rng.range_inclusive(0, 2);   // the draws rt_hit_bvh makes, in its order (pre-order)
    std::vector<rt_bvh_node_desc> nodes((size_t)(2 * n - 1));
    const int made = rt_bvh_build_device(ctx, boxes.data(), n, axes.data(), nodes.data(), nullptr);
    if (made < 0) return made;
    const int32_t first = (int32_t)b->bvh_nodes.size();
    for (rt_bvh_node_desc nd : nodes) {
        if (nd.object >= 0) nd.object = ids[nd.object];   // position in `ids` -> hittable id
        else { nd.left += first; nd.right += first; }
        b->bvh_nodes.push_back(nd);
    }
    rt_hittable_desc h = blank_hittable(RT_HIT_BVH);
    h.child = first;
    h.count = 2 * n - 1;
    std::memcpy(h.bbox, nodes[0].bbox, sizeof(h.bbox));
    b->hittables.push_back(h);
    return (int)b->hittables.size() - 1;
}

int rt_builder_finish(rt_builder* b, int world, rt_scene_desc* out) {
    if (!b || !out) return fail(RT_ERR_INVALID_ARGUMENT, "rt_builder_finish: null argument");
    if (!valid_id(world, b->hittables.size())) return fail(RT_ERR_OUT_OF_RANGE, "rt_builder_finish: unknown world id");
    std::memset(out, 0, sizeof(*out));
    out->abi_version = RT_B200_ABI_VERSION;
    out->world = world;
    out->n_textures = (int32_t)b->textures.size();
    out->n_materials = (int32_t)b->materials.size();
    out->n_hittables = (int32_t)b->hittables.size();
    out->n_list_items = (int32_t)b->list_items.size();
    out->n_bvh_nodes = (int32_t)b->bvh_nodes.size();
    out->n_perlins = (int32_t)b->perlins.size();
    out->n_images = (int32_t)b->images.size();
    out->textures = b->textures.data();
    out->materials = b->materials.data();
    out->hittables = b->hittables.data();
    out->list_items = b->list_items.data();
    out->bvh_nodes = b->bvh_nodes.data();
    out->perlins = b->perlins.data();
    out->images = b->images.data();
    return RT_OK;
}

// -------------------------------------------------------------------- camera
void rt_camera_settings_default(rt_camera_settings* s) {
    if (!s) return;
    std::memset(s, 0, sizeof(*s));
    s->aspect_ratio = 16.0 / 9.0;  // camera.rs:21-37
    s->image_width = 400;
    s->samples_per_pixel = 100;
    s->max_depth = 50;
    s->vfov = 90.0;
    s->look_at[2] = -1.0;
    s->vup[1] = 1.0;
    s->defocus_angle = 0.0;
    s->focus_dist = 10.0;
}

int rt_camera_new(const rt_camera_settings* s, rt_camera_desc* c) {
    if (!s || !c) return fail(RT_ERR_INVALID_ARGUMENT, "rt_camera_new: null argument");
    if (s->image_width <= 0) return fail(RT_ERR_INVALID_ARGUMENT, "rt_camera_new: image_width must be positive");
    std::memset(c, 0, sizeof(*c));
    using rt_host::V3;
    // camera.rs:54-110, operation for operation.
    const double wf = (double)s->image_width;
    const int64_t image_height = (int64_t)(wf / s->aspect_ratio);  // `as usize` truncates; no >=1 clamp (camera.rs:69)
    if (image_height <= 0) return fail(RT_ERR_INVALID_ARGUMENT, "rt_camera_new: image_height truncates to 0");
    const double theta = s->vfov * kPi / 180.0;
    const double h = std::tan(theta / 2.0);
    const double viewport_height = 2.0 * h * s->focus_dist;
    const double viewport_width = viewport_height * (wf / (double)image_height);
    const V3 look_from(s->look_from), look_at(s->look_at), vup(s->vup);
    const V3 w = (look_from - look_at).normalize();
    const V3 u = vup.cross(w).normalize();
    const V3 v = w.cross(u);
    const V3 viewport_u = u * viewport_width;
    const V3 viewport_v = v * (-viewport_height);
    const V3 center = look_from;
    const V3 pixel_delta_u = viewport_u / wf;
    const V3 pixel_delta_v = viewport_v / (double)image_height;
    const V3 viewport_upper_left = center - w * s->focus_dist - viewport_u * 0.5 - viewport_v * 0.5;
    const V3 pixel00_loc = viewport_upper_left + (pixel_delta_u + pixel_delta_v) * 0.5;
    const double defocus_radius = s->focus_dist * std::tan((s->defocus_angle / 2.0) * kPi / 180.0);
    const V3 defocus_disk_u = u * defocus_radius;
    const V3 defocus_disk_v = v * defocus_radius;
    c->image_width = s->image_width;
    c->image_height = image_height;
    c->samples_per_pixel = s->samples_per_pixel;
    c->max_depth = s->max_depth;
    for (int k = 0; k < 3; ++k) c->background[k] = s->background[k];
    center.store(c->center);
    pixel00_loc.store(c->pixel00_loc);
    pixel_delta_u.store(c->pixel_delta_u);
    pixel_delta_v.store(c->pixel_delta_v);
    c->defocus_angle = s->defocus_angle;
    defocus_disk_u.store(c->defocus_disk_u);
    defocus_disk_v.store(c->defocus_disk_v);
    return RT_OK;
}

}  // extern "C"
