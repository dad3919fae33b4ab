// rt_scene_desc (the reference's object graph, f64) -> threaded f32 op stream (dev_scene.h).
// Host code, no CUDA. This is the "flatten" step the north star asks for: the BVH is built on
// the host with the reference's split rule (scene_builder.cpp) and linearised here.
#include "dev_scene.h"

#include "../../../include/rt_b200.h"

#include <cmath>
#include <cstring>
#include <limits>
#include <string>
#include <unordered_map>
#include <utility>

namespace rtdev {
namespace {

struct Box {
    double lo[3], hi[3];
    bool valid = false;
};

Box box_union(const Box& a, const Box& b) {
    if (!a.valid) return b;
    if (!b.valid) return a;
    Box r;
    r.valid = true;
    for (int c = 0; c < 3; ++c) {
        r.lo[c] = std::fmin(a.lo[c], b.lo[c]);
        r.hi[c] = std::fmax(a.hi[c], b.hi[c]);
    }
    return r;
}

float bits_to_float(uint32_t u) {
    float f;
    std::memcpy(&f, &u, 4);
    return f;
}
float int_to_float_bits(int32_t i) { return bits_to_float((uint32_t)i); }

// f64 -> f32 rounded outward with ~2 ulp of slack, so the f32 slab test stays conservative.
float round_down(double v) {
    float f = (float)v;
    const float ninf = -std::numeric_limits<float>::infinity();
    f = std::nextafterf(f, ninf);
    return std::nextafterf(f, ninf);
}
float round_up(double v) {
    float f = (float)v;
    const float pinf = std::numeric_limits<float>::infinity();
    f = std::nextafterf(f, pinf);
    return std::nextafterf(f, pinf);
}

struct Compiler {
    const rt_scene_desc* d;
    CompiledScene* out;
    CompileOptions opt;
    std::string err;
    int status = 0;
    // Enclosing instances, innermost last: word index of the OP_XFORM_ENTER and its COMPOSED world -> local transform
    // (local = R(x - a) + b). An instance nested in another instance's subtree is emitted with the composition, so the
    // device always transforms the world ray and needs no stack of saved rays; the matching OP_XFORM_EXIT names its parent.
    struct XfParams { double a[3], b[3], s, c; };
    std::vector<std::pair<int, XfParams>> xf_stack;
    bool in_xform() const { return !xf_stack.empty(); }
    static XfParams compose(const XfParams& P, const XfParams& C) {
        // lp = Rp(x - ap) + bp ; lc = Rc(lp - ac) + bc  =>  lc = (Rc Rp)(x - ap) + Rc(bp - ac) + bc
        XfParams X;
        for (int k = 0; k < 3; ++k) X.a[k] = P.a[k];
        X.s = C.s * P.c + C.c * P.s;
        X.c = C.c * P.c - C.s * P.s;
        const double q[3] = {P.b[0] - C.a[0], P.b[1] - C.a[1], P.b[2] - C.a[2]};
        X.b[0] = C.c * q[0] - C.s * q[2] + C.b[0];
        X.b[1] = q[1] + C.b[1];
        X.b[2] = C.s * q[0] + C.c * q[2] + C.b[2];
        return X;
    }
    std::vector<Box> tight_hittable;  // memo, by hittable id
    std::vector<Box> tight_node;      // memo, by bvh node index
    std::vector<char> have_h, have_n;
    std::vector<signed char> loose_h, loose_n;   // -1 unknown, 0 / 1: subtree holds a quad that sticks out of its own box
    double scale = 1.0;
    bool has_reference_boxes = false;    // an OP_INNER_REF was emitted: hits then depend on the order of the tests (compile_scene)
    std::vector<F4> hoisted;             // bodies of world-space media (appended after the world program)
    std::vector<int32_t> hoisted_at;     // offsets into `hoisted`

    bool fail(int code, const std::string& m) {
        if (status == 0) { status = code; err = m; }
        return false;
    }

    // Quad::new boxes the diagonal q .. q+u+v only (quad.rs:41-43): a parallelogram that is not axis aligned sticks
    // out of its own bounding box.
    static void quad_corners(const rt_hittable_desc& h, double c[4][3]) {
        for (int k = 0; k < 3; ++k) {
            c[0][k] = h.v0[k]; c[1][k] = h.v0[k] + h.v1[k]; c[2][k] = h.v0[k] + h.v2[k]; c[3][k] = h.v0[k] + h.v1[k] + h.v2[k];
        }
    }
    bool loose_quad(const rt_hittable_desc& h) const {
        double c[4][3];
        quad_corners(h, c);
        for (int i = 0; i < 4; ++i)
            for (int k = 0; k < 3; ++k) {
                const double eps = 1e-9 * std::fmax(1.0, std::fabs(c[i][k]));
                if (c[i][k] < h.bbox[2 * k] - eps || c[i][k] > h.bbox[2 * k + 1] + eps) return true;
            }
        return false;
    }
    bool loose(int id) {
        if (loose_h[id] >= 0) return loose_h[id] != 0;
        const rt_hittable_desc& h = d->hittables[id];
        bool l = false;
        switch (h.kind) {
            case RT_HIT_QUAD: l = loose_quad(h); break;
            case RT_HIT_LIST: for (int i = 0; i < h.count && !l; ++i) l = loose(d->list_items[h.child + i]); break;
            case RT_HIT_TRANSLATE: case RT_HIT_ROTATE_Y: case RT_HIT_CONSTANT_MEDIUM: l = loose(h.child); break;
            case RT_HIT_BVH: l = loose_node(h.child); break;
            default: break;
        }
        loose_h[id] = l ? 1 : 0;
        return l;
    }
    bool loose_node(int n) {
        if (loose_n[n] >= 0) return loose_n[n] != 0;
        const rt_bvh_node_desc& node = d->bvh_nodes[n];
        const bool l = node.object >= 0 ? loose(node.object) : (loose_node(node.left) || loose_node(node.right));
        loose_n[n] = l ? 1 : 0;
        return l;
    }

    // Tight bounds of the geometry in the hittable's own (outer) space. The reference's list boxes
    // always contain the origin (HittableList derives Default; hittable.rs:50-59) and its quad
    // boxes are padded (aabb.rs:35-53); neither changes which hits exist, so the device uses the
    // tight union (rounded outward) and culls more.
    Box tight(int id) {
        if (have_h[id]) return tight_hittable[id];
        const rt_hittable_desc& h = d->hittables[id];
        Box b;
        switch (h.kind) {
            case RT_HIT_SPHERE:
            case RT_HIT_QUAD:
                b.valid = true;
                for (int c = 0; c < 3; ++c) { b.lo[c] = h.bbox[2 * c]; b.hi[c] = h.bbox[2 * c + 1]; }
                if (h.kind == RT_HIT_QUAD && loose_quad(h)) {   // the boxes the device adds (lists, instances, media) must hold all of it
                    double q[4][3];
                    quad_corners(h, q);
                    for (int i = 0; i < 4; ++i)
                        for (int c = 0; c < 3; ++c) { b.lo[c] = std::fmin(b.lo[c], q[i][c]); b.hi[c] = std::fmax(b.hi[c], q[i][c]); }
                }
                break;
            case RT_HIT_LIST:
                for (int i = 0; i < h.count; ++i) b = box_union(b, tight(d->list_items[h.child + i]));
                break;
            case RT_HIT_TRANSLATE: {
                const Box cb = tight(h.child);
                b = cb;
                if (cb.valid) for (int c = 0; c < 3; ++c) { b.lo[c] = cb.lo[c] + h.v0[c]; b.hi[c] = cb.hi[c] + h.v0[c]; }
                break;
            }
            case RT_HIT_ROTATE_Y: {
                const Box cb = tight(h.child);
                if (cb.valid) {
                    const double inf = std::numeric_limits<double>::infinity();
                    b.valid = true;
                    for (int c = 0; c < 3; ++c) { b.lo[c] = inf; b.hi[c] = -inf; }
                    for (int i = 0; i < 2; ++i) for (int j = 0; j < 2; ++j) for (int k = 0; k < 2; ++k) {
                        const double x = i ? cb.hi[0] : cb.lo[0], y = j ? cb.hi[1] : cb.lo[1], z = k ? cb.hi[2] : cb.lo[2];
                        const double p[3] = {h.s1 * x + h.s0 * z, y, -h.s0 * x + h.s1 * z};  // hittable.rs:137-138
                        for (int c = 0; c < 3; ++c) { b.lo[c] = std::fmin(b.lo[c], p[c]); b.hi[c] = std::fmax(b.hi[c], p[c]); }
                    }
                }
                break;
            }
            case RT_HIT_CONSTANT_MEDIUM:
                b = tight(h.child);
                break;
            case RT_HIT_BVH:
                b = tight_of_node(h.child);
                break;
        }
        tight_hittable[id] = b;
        have_h[id] = 1;
        return b;
    }
    Box tight_of_node(int n) {
        if (have_n[n]) return tight_node[n];
        const rt_bvh_node_desc& node = d->bvh_nodes[n];
        Box b = node.object >= 0 ? tight(node.object) : box_union(tight_of_node(node.left), tight_of_node(node.right));
        tight_node[n] = b;
        have_n[n] = 1;
        return b;
    }

    int here() const { return (int)out->ops.size(); }
    void push(float x, float y, float z, float w) { out->ops.push_back(F4{x, y, z, w}); }

    // Cull box of `kind` (OP_INNER / OP_XFORM_ENTER: centre / half extent, padded; OP_INNER_REF: the reference's corners);
    // w1.w = skip (a word index until the final pass turns it into a link). Returns the index of word 1.
    int push_box_header(const Box& b, uint32_t kind, int size_words) {
        const uint32_t hdr = make_hdr(kind, kind == OP_INNER_REF ? FLAG_ALWAYS : 0u, (uint32_t)size_words);
        if (kind == OP_INNER_REF) has_reference_boxes = true;
        if (!b.valid) {  // empty subtree: a box nothing can hit (far < near on every axis, whatever the ray)
            const float inf = std::numeric_limits<float>::infinity();
            if (kind == OP_INNER_REF) { push(inf, inf, inf, bits_to_float(hdr)); push(-inf, -inf, -inf, 0.0f); }
            else { push(0.0f, 0.0f, 0.0f, bits_to_float(hdr)); push(-inf, -inf, -inf, 0.0f); }
            return here() - 1;
        }
        for (int c = 0; c < 3; ++c)
            if (std::isfinite(b.lo[c]) && std::isfinite(b.hi[c])) scale = std::fmax(scale, std::fmax(std::fabs(b.lo[c]), std::fabs(b.hi[c])));
        if (kind == OP_INNER_REF) {
            push(round_down(b.lo[0]), round_down(b.lo[1]), round_down(b.lo[2]), bits_to_float(hdr));
            push(round_up(b.hi[0]), round_up(b.hi[1]), round_up(b.hi[2]), 0.0f);
            return here() - 1;
        }
        float c[3], h[3];
        for (int k = 0; k < 3; ++k) {
            if (!std::isfinite(b.lo[k]) || !std::isfinite(b.hi[k])) { c[k] = 0.0f; h[k] = std::numeric_limits<float>::infinity(); continue; }
            const double cd = 0.5 * (b.lo[k] + b.hi[k]), hd = 0.5 * (b.hi[k] - b.lo[k]);
            c[k] = (float)cd;
            // covers the rounding of c, of the per-ray products and of the prim tests (dev_scene.h, CULL BOXES)
            const double pad = std::ldexp(std::fabs(cd) + hd, -21) + std::fabs((double)c[k] - cd) + 1e-30;
            h[k] = std::nextafterf((float)(hd + pad), std::numeric_limits<float>::infinity());
        }
        push(c[0], c[1], c[2], bits_to_float(hdr));
        push(h[0], h[1], h[2], 0.0f);
        return here() - 1;
    }
    void patch_skip(int word1) { out->ops[word1].w = int_to_float_bits(here()); }

    int add_precise(const rt_hittable_desc& h) {
        const int idx = (int)out->precise.size() / 2;
        out->precise.push_back(D4{h.v0[0], h.v0[1], h.v0[2], h.s0});
        out->precise.push_back(D4{h.v1[0], h.v1[1], h.v1[2], 0.0});
        return idx;
    }
    static bool wants_precise(const rt_hittable_desc& h) { return std::fabs(h.s0) > 200.0; }

    void emit_sphere(int id) {
        const rt_hittable_desc& h = d->hittables[id];
        uint32_t flags = 0;
        if (h.flags & RT_FLAG_MOVING) flags |= FLAG_MOVING;
        int pidx = 0;
        if (wants_precise(h)) { flags |= FLAG_PRECISE; pidx = add_precise(h); }
        push((float)h.v0[0], (float)h.v0[1], (float)h.v0[2], bits_to_float(make_hdr(OP_SPHERE, flags, (flags & FLAG_MOVING) ? 3 : 2)));
        push((float)h.s0, int_to_float_bits(h.mat), int_to_float_bits(id), int_to_float_bits(pidx));
        if (flags & FLAG_MOVING) push((float)h.v1[0], (float)h.v1[1], (float)h.v1[2], 0.0f);
    }

    void emit_quad(int id) {
        const rt_hittable_desc& h = d->hittables[id];
        const double* u = h.v1; const double* v = h.v2; const double* w = h.v3; const double* q = h.v0;
        // alpha = w.(p x v) = p.(v x w);  beta = w.(u x p) = p.(w x u)   with p = hit - q
        const double A[3] = {v[1] * w[2] - v[2] * w[1], v[2] * w[0] - v[0] * w[2], v[0] * w[1] - v[1] * w[0]};
        const double B[3] = {w[1] * u[2] - w[2] * u[1], w[2] * u[0] - w[0] * u[2], w[0] * u[1] - w[1] * u[0]};
        const double a0 = -(A[0] * q[0] + A[1] * q[1] + A[2] * q[2]);
        const double b0 = -(B[0] * q[0] + B[1] * q[1] + B[2] * q[2]);
        push((float)h.n[0], (float)h.n[1], (float)h.n[2], bits_to_float(make_hdr(OP_QUAD, 0, 4)));
        push((float)A[0], (float)A[1], (float)A[2], (float)a0);
        push((float)B[0], (float)B[1], (float)B[2], (float)b0);
        push((float)h.s0, int_to_float_bits(h.mat), int_to_float_bits(id), 0.0f);
    }

    // A list made by Quad::cube (quad.rs:45-93) whose six quads still are what cube() produced: consecutive ids,
    // one material, axis-aligned faces of the recorded min/max corners.
    bool is_cube(int id) const {
        if (!opt.box_primitives) return false;
        const rt_hittable_desc& h = d->hittables[id];
        if (h.kind != RT_HIT_LIST || !(h.flags & RT_FLAG_CUBE_LIST) || h.count != 6) return false;
        const int first = d->list_items[h.child];
        for (int k = 0; k < 6; ++k) {
            const int q = d->list_items[h.child + k];
            if (q != first + k || q < 0 || q >= d->n_hittables) return false;
            if (d->hittables[q].kind != RT_HIT_QUAD || d->hittables[q].mat != d->hittables[first].mat) return false;
        }
        for (int c = 0; c < 3; ++c) if (!(h.v0[c] < h.v1[c])) return false;   // a degenerate (flat) cube keeps its quads
        return true;
    }
    void emit_box(int id) {
        const rt_hittable_desc& h = d->hittables[id];
        const int first = d->list_items[h.child];
        // centre / half extent of the f32 corners (what the hit record is computed from), rounded to nearest, NOT padded:
        // this box is the primitive, not a cull box
        float cc[3], hh[3];
        for (int k = 0; k < 3; ++k) {
            const double lo = (double)(float)h.v0[k], hi = (double)(float)h.v1[k];
            cc[k] = (float)(0.5 * (lo + hi));
            hh[k] = (float)(0.5 * (hi - lo));
        }
        push(cc[0], cc[1], cc[2], bits_to_float(make_hdr(OP_BOX, 0, 4)));
        push(hh[0], hh[1], hh[2], 0.0f);   // .w: the fall-through link, written by the final pass
        push((float)h.v0[0], (float)h.v0[1], (float)h.v0[2], int_to_float_bits(first));
        push((float)h.v1[0], (float)h.v1[1], (float)h.v1[2], int_to_float_bits(d->hittables[first].mat));
        for (int c = 0; c < 3; ++c) scale = std::fmax(scale, std::fmax(std::fabs(h.v0[c]), std::fabs(h.v1[c])));
    }

    // Fold a chain of directly nested Translate / RotateY wrappers (outermost first) into local = R(x - a) + b.
    int fold_xform(int id, double a[3], double b[3], double* s, double* c) const {
        a[0] = a[1] = a[2] = b[0] = b[1] = b[2] = 0.0;
        *s = 0.0; *c = 1.0;
        bool rotated = false;
        int cur = id;
        while (d->hittables[cur].kind == RT_HIT_TRANSLATE || d->hittables[cur].kind == RT_HIT_ROTATE_Y) {
            const rt_hittable_desc& x = d->hittables[cur];
            if (x.kind == RT_HIT_TRANSLATE) {
                if (!rotated && b[0] == 0 && b[1] == 0 && b[2] == 0) for (int k = 0; k < 3; ++k) a[k] += x.v0[k];
                else for (int k = 0; k < 3; ++k) b[k] -= x.v0[k];
            } else {
                const double s2 = x.s0, c2 = x.s1;
                const double bx = c2 * b[0] - s2 * b[2], bz = s2 * b[0] + c2 * b[2];
                b[0] = bx; b[2] = bz;
                if (!rotated) { *s = s2; *c = c2; }
                else { const double ns = *s * c2 + *c * s2, nc = *c * c2 - *s * s2; *s = ns; *c = nc; }
                rotated = true;
            }
            cur = x.child;
        }
        return cur;
    }

    void emit_node(int n, bool in_boundary) {
        const rt_bvh_node_desc& node = d->bvh_nodes[n];
        if (loose_node(n)) {   // the reference's own box and test, at every node of the path down to the loose quad (dev_scene.h)
            Box rb;
            rb.valid = true;
            for (int c = 0; c < 3; ++c) { rb.lo[c] = node.bbox[2 * c]; rb.hi[c] = node.bbox[2 * c + 1]; }
            const int w1 = push_box_header(rb, OP_INNER_REF, 2);
            if (node.object >= 0) emit(node.object, in_boundary);
            else { emit_node(node.left, in_boundary); emit_node(node.right, in_boundary); }
            patch_skip(w1);
            return;
        }
        if (node.object >= 0) { emit(node.object, in_boundary); return; }
        const int w1 = push_box_header(tight_of_node(n), OP_INNER, 2);
        emit_node(node.left, in_boundary);
        emit_node(node.right, in_boundary);
        patch_skip(w1);
    }

    void emit(int id, bool in_boundary) {
        if (status) return;
        if (id < 0 || id >= d->n_hittables) { fail(RT_ERR_OUT_OF_RANGE, "scene references an unknown hittable id"); return; }
        const rt_hittable_desc& h = d->hittables[id];
        switch (h.kind) {
            case RT_HIT_SPHERE: emit_sphere(id); break;
            case RT_HIT_QUAD: emit_quad(id); break;
            case RT_HIT_LIST: {
                if (h.count == 0) break;
                if (is_cube(id)) { emit_box(id); break; }
                const int w1 = push_box_header(tight(id), OP_INNER, 2);
                for (int i = 0; i < h.count; ++i) emit(d->list_items[h.child + i], in_boundary);
                patch_skip(w1);
                break;
            }
            case RT_HIT_TRANSLATE:
            case RT_HIT_ROTATE_Y: {
                XfParams X;
                const int cur = fold_xform(id, X.a, X.b, &X.s, &X.c);       // relative to the enclosing space
                if (in_xform()) X = compose(xf_stack.back().second, X);     // world -> local
                const int w1 = push_box_header(tight(id), OP_XFORM_ENTER, 4);   // box in the enclosing space
                push((float)X.a[0], (float)X.a[1], (float)X.a[2], (float)X.s);
                push((float)X.b[0], (float)X.b[1], (float)X.b[2], (float)X.c);
                xf_stack.push_back({w1 - 1, X});
                emit(cur, in_boundary);
                xf_stack.pop_back();
                const int parent = in_xform() ? xf_stack.back().first : -1;
                push(int_to_float_bits(parent), 0.0f, 0.0f, bits_to_float(make_hdr(OP_XFORM_EXIT, FLAG_ALWAYS, 2)));
                push(0.0f, 0.0f, 0.0f, 0.0f);
                patch_skip(w1);
                break;
            }
            case RT_HIT_CONSTANT_MEDIUM: {
                if (in_boundary) { fail(RT_ERR_UNSUPPORTED, "a ConstantMedium used as the boundary of another medium is not supported"); return; }
                const rt_hittable_desc& bd = d->hittables[h.child];
                double xa[3], xb[3], xs, xc;
                const int inner = fold_xform(h.child, xa, xb, &xs, &xc);
                const bool analytic = bd.kind == RT_HIT_SPHERE || (!in_xform() && is_cube(inner));
                // A world-space medium with an analytic boundary is evaluated once at the start of every segment
                // instead of at its BVH position: its free-flight draw is keyed by (segment, medium), not by visit
                // order, and closest-hit is order independent, so the result is the same and every lane of a warp
                // runs it at the same time.
                const bool hoist = opt.hoist_media && analytic && !in_xform() && (int)hoisted_at.size() < kMaxHoistedMedia;
                std::vector<F4> saved;
                int w1 = -1;
                if (hoist) { saved.swap(out->ops); hoisted_at.push_back((int32_t)hoisted.size()); }
                else w1 = push_box_header(tight(id), OP_INNER, 2);   // the medium's box, then its body
                if (bd.kind == RT_HIT_SPHERE) {
                    push((float)h.s0, int_to_float_bits(h.mat), int_to_float_bits(id), bits_to_float(make_hdr(OP_MEDIUM, MEDIUM_BOUNDARY_SPHERE, 3)));
                    // A medium's entry/exit distances feed a random free-flight comparison, never a surface position, so
                    // f32 roots (relative error 1e-7) are enough even for the r = 5000 fog of final_scene.
                    uint32_t pidx = 0;
                    const bool precise = false;
                    push((float)bd.v0[0], (float)bd.v0[1], (float)bd.v0[2], (float)bd.s0);
                    const uint32_t aux = ((bd.flags & RT_FLAG_MOVING) ? FLAG_MOVING : 0u) | (precise ? FLAG_PRECISE : 0u);
                    push((float)bd.v1[0], (float)bd.v1[1], (float)bd.v1[2], int_to_float_bits((int32_t)(pidx | (aux << 24))));
                } else if (!in_xform() && is_cube(inner)) {
                    const rt_hittable_desc& cube = d->hittables[inner];
                    push((float)h.s0, int_to_float_bits(h.mat), int_to_float_bits(id), bits_to_float(make_hdr(OP_MEDIUM, MEDIUM_BOUNDARY_XBOX, 5)));
                    push((float)xa[0], (float)xa[1], (float)xa[2], (float)xs);
                    push((float)xb[0], (float)xb[1], (float)xb[2], (float)xc);
                    push((float)cube.v0[0], (float)cube.v0[1], (float)cube.v0[2], 0.0f);
                    push((float)cube.v1[0], (float)cube.v1[1], (float)cube.v1[2], 0.0f);
                } else {
                    push((float)h.s0, int_to_float_bits(h.mat), int_to_float_bits(id), bits_to_float(make_hdr(OP_MEDIUM, MEDIUM_BOUNDARY_PROGRAM, 3)));
                    const int wb = here();
                    push(0.0f, 0.0f, 0.0f, 0.0f);
                    push(0.0f, 0.0f, 0.0f, 0.0f);
                    const int bbegin = here();
                    // The program is run on the CURRENT ray - the local ray of the enclosing instance, if there is one - so
                    // instances inside it compose their transforms from the program's own frame, not from the world
                    // (traverse<false>() treats the ray it is handed as "world": an exit with parent -1 returns to it).
                    std::vector<std::pair<int, XfParams>> enclosing;
                    enclosing.swap(xf_stack);
                    emit(h.child, true);
                    xf_stack.swap(enclosing);
                    out->ops[wb].x = int_to_float_bits(bbegin);
                    out->ops[wb].y = int_to_float_bits(here());
                }
                if (hoist) {
                    hoisted.insert(hoisted.end(), out->ops.begin(), out->ops.end());
                    out->ops.swap(saved);
                } else {
                    patch_skip(w1);
                }
                break;
            }
            case RT_HIT_BVH: {
                out->bvh_hittable_ids.push_back(id);
                std::vector<int32_t> pre;
                collect_preorder(h.child, &pre);
                out->bvh_preorder_objects.push_back(pre);
                emit_node(h.child, in_boundary);
                break;
            }
            default:
                fail(RT_ERR_INVALID_ARGUMENT, "unknown hittable kind");
        }
    }

    void collect_preorder(int n, std::vector<int32_t>* o) {
        const rt_bvh_node_desc& node = d->bvh_nodes[n];
        o->push_back(node.object);
        if (node.object < 0) { collect_preorder(node.left, o); collect_preorder(node.right, o); }
    }
};


// ---------------------------------------------------------------------------------------------------------------
// Which cull boxes pay for themselves. The stream visits the primitives in the reference's order whatever boxes
// stand between them, so an OP_INNER can be dropped (its children are then tested whenever ITS parent passed)
// without changing any result: boxes only cull. The reference's tree splits on a random axis by box-min median
// (bvh.rs:31-66), so many of its nodes are nearly as large as their parent: a ray that reached the parent passes
// them almost surely and the test is wasted. With P(ray passes box n | it is inside kept ancestor a) taken as the
// surface-area ratio SA(n)/SA(a), the expected number of tests below a is
//     cost(n | a) = min( keep:     1 + SA(n)/SA(a) * sum_c cost(c | n),
//                        dissolve: sum_c cost(c | a) )
// which a memoised recursion minimises exactly (the ancestors of a node are few). Measured on the segments of real
// paths with the host-side stream walk (tools/opstream_cost.py): final_scene 21.3 -> 15.9 box tests per segment,
// random_balls 41.8 -> 31.4, cornell_box 6.8 -> 2.0, cornell_smoke 6.8 -> 0, closest hits bit-identical.
// OP_INNER_REF nodes are semantics, not culling, and always stay; medium boundary programs are left as they are.
struct IrNode {
    uint32_t kind = 0;
    int at = 0;            // word index in the unpruned stream
    int n_words = 0;       // words of the op itself (without children / exit)
    bool frozen = false;   // inside a medium boundary program: copied verbatim
    std::vector<IrNode> ch;
    double area = 0.0;     // box-headed ops
};

struct Pruner {
    const std::vector<F4>& in;
    std::vector<F4> out;
    std::unordered_map<uint64_t, std::vector<std::pair<double, double>>> memo;   // at -> [(ancestor area, cost)]
    std::vector<int> xf_pos;     // emitted positions of the enclosing OP_XFORM_ENTERs (OP_XFORM_EXIT names its parent)
    CompileOptions opt;
    Pruner(const std::vector<F4>& ops, const CompileOptions& o) : in(ops), opt(o) {}

    static uint32_t hdr_of(const F4& w) { uint32_t u; std::memcpy(&u, &w.w, 4); return u; }
    static int32_t int_of(float f) { int32_t v; std::memcpy(&v, &f, 4); return v; }
    static void set_int(float* f, int32_t v) { std::memcpy(f, &v, 4); }

    double box_area(int i) const {
        double ex, ey, ez;
        if (hdr_kind(hdr_of(in[i])) == OP_INNER_REF) {
            ex = (double)in[i + 1].x - in[i].x; ey = (double)in[i + 1].y - in[i].y; ez = (double)in[i + 1].z - in[i].z;
        } else {
            ex = 2.0 * in[i + 1].x; ey = 2.0 * in[i + 1].y; ez = 2.0 * in[i + 1].z;   // centre / half extent
        }
        if (!(ex >= 0.0) || !(ey >= 0.0) || !(ez >= 0.0)) return 0.0;          // the empty box
        const double a = 2.0 * (ex * ey + ey * ez + ex * ez);
        return a;                                                               // may be +inf
    }

    std::vector<IrNode> parse(int begin, int end, bool frozen) const {
        std::vector<IrNode> nodes;
        int i = begin;
        while (i < end) {
            const uint32_t hdr = hdr_of(in[i]), kind = hdr_kind(hdr), flags = hdr_flags(hdr);
            IrNode n;
            n.kind = kind; n.at = i; n.frozen = frozen;
            int next = i;
            switch (kind) {
                case OP_INNER: case OP_INNER_REF:
                    n.n_words = 2; n.area = box_area(i); next = int_of(in[i + 1].w);
                    n.ch = parse(i + 2, next, frozen);
                    break;
                case OP_XFORM_ENTER:
                    n.n_words = 4; n.area = box_area(i); next = int_of(in[i + 1].w);
                    n.ch = parse(i + 4, next - 2, frozen);                      // the matching OP_XFORM_EXIT is re-emitted by emit()
                    break;
                case OP_MEDIUM:
                    if ((int)flags == MEDIUM_BOUNDARY_PROGRAM) {
                        n.n_words = 3; next = int_of(in[i + 1].y);
                        n.ch = parse(int_of(in[i + 1].x), next, true);
                    } else {
                        n.n_words = (int)flags == MEDIUM_BOUNDARY_XBOX ? 5 : 3; next = i + n.n_words;
                    }
                    break;
                default: n.n_words = op_words(kind, flags); next = i + n.n_words; break;   // leaves (OP_XFORM_EXIT never appears here)
            }
            nodes.push_back(std::move(n));
            i = next;
        }
        return nodes;
    }

    static double pass_probability(double area, double ancestor) {
        if (!(ancestor > 0.0) || std::isinf(ancestor) || std::isinf(area)) return 1.0;
        return std::fmin(1.0, area / ancestor);
    }
    double cost_list(const std::vector<IrNode>& nodes, double ancestor) {
        double c = 0.0;
        for (const IrNode& n : nodes) c += cost(n, ancestor);
        return c;
    }
    bool keep(const IrNode& n, double ancestor) {
        const double k = 1.0 + pass_probability(n.area, ancestor) * cost_list(n.ch, n.area);
        return k <= cost_list(n.ch, ancestor);
    }
    // relative costs of one op for a lane (instructions of the op's body in the render kernel, OP_INNER = 1)
    double cost(const IrNode& n, double ancestor) {
        switch (n.kind) {
            case OP_SPHERE: return opt.cost_sphere;
            case OP_QUAD: return opt.cost_quad;
            case OP_BOX: return opt.cost_box;
            case OP_MEDIUM: return opt.cost_medium;
            default: break;
        }
        auto& slot = memo[(uint64_t)n.at];
        for (const auto& e : slot) if (e.first == ancestor) return e.second;
        double c;
        const double p = pass_probability(n.area, ancestor);
        if (n.kind == OP_INNER && !n.frozen) {
            const double k = 1.0 + p * cost_list(n.ch, n.area), dsv = cost_list(n.ch, ancestor);
            c = k <= dsv ? k : dsv;
        } else if (n.kind == OP_XFORM_ENTER) {
            c = opt.cost_xform + p * (cost_list(n.ch, n.area) + 0.5);
        } else {   // OP_INNER_REF, frozen OP_INNER
            c = 1.0 + p * cost_list(n.ch, n.area);
        }
        memo[(uint64_t)n.at].push_back({ancestor, c});
        return c;
    }

    void copy_words(const IrNode& n) { for (int k = 0; k < n.n_words; ++k) out.push_back(in[n.at + k]); }
    void emit_list(const std::vector<IrNode>& nodes, double ancestor) {
        for (const IrNode& n : nodes) emit(n, ancestor);
    }
    void emit(const IrNode& n, double ancestor) {
        switch (n.kind) {
            case OP_INNER: case OP_INNER_REF: {
                if (n.kind == OP_INNER && !n.frozen && !keep(n, ancestor)) { emit_list(n.ch, ancestor); return; }
                const int pos = (int)out.size();
                copy_words(n);
                emit_list(n.ch, n.area);
                set_int(&out[pos + 1].w, (int32_t)out.size());
                return;
            }
            case OP_XFORM_ENTER: {
                const int pos = (int)out.size();
                copy_words(n);
                xf_pos.push_back(pos);
                emit_list(n.ch, n.area);
                xf_pos.pop_back();
                const int exit_at = int_of(in[n.at + 1].w) - 2;
                out.push_back(in[exit_at]);
                out.push_back(in[exit_at + 1]);
                set_int(&out[out.size() - 2].x, xf_pos.empty() ? -1 : (int32_t)xf_pos.back());
                set_int(&out[pos + 1].w, (int32_t)out.size());
                return;
            }
            case OP_MEDIUM: {
                const int pos = (int)out.size();
                copy_words(n);
                if (!n.ch.empty() || hdr_flags(hdr_of(in[n.at])) == (uint32_t)MEDIUM_BOUNDARY_PROGRAM) {
                    set_int(&out[pos + 1].x, (int32_t)out.size());
                    std::vector<int> enclosing;          // the program runs on the current (local) ray: its instances' exits
                    enclosing.swap(xf_pos);              // name parents inside the program only, -1 = the ray it was handed
                    emit_list(n.ch, std::numeric_limits<double>::infinity());
                    xf_pos.swap(enclosing);
                    set_int(&out[pos + 1].y, (int32_t)out.size());
                }
                return;
            }
            default:
                copy_words(n);
        }
    }
};

void prune_stream(std::vector<F4>* ops, const CompileOptions& opt) {
    if (ops->empty()) return;
    Pruner p(*ops, opt);
    const std::vector<IrNode> tree = p.parse(0, (int)ops->size(), false);
    p.emit_list(tree, std::numeric_limits<double>::infinity());
    ops->swap(p.out);
}

}  // namespace

int compile_scene(const rt_scene_desc* desc, const CompileOptions& opt, CompiledScene* out, const char** err) {
    static thread_local std::string msg;
    if (!desc || !out) { msg = "compile_scene: null argument"; if (err) *err = msg.c_str(); return RT_ERR_INVALID_ARGUMENT; }
    if (desc->abi_version != RT_B200_ABI_VERSION) { msg = "rt_scene_desc.abi_version mismatch"; if (err) *err = msg.c_str(); return RT_ERR_INVALID_ARGUMENT; }
    Compiler c;
    c.d = desc;
    c.out = out;
    c.opt = opt;
    c.tight_hittable.resize(desc->n_hittables);
    c.have_h.assign(desc->n_hittables, 0);
    c.tight_node.resize(desc->n_bvh_nodes);
    c.have_n.assign(desc->n_bvh_nodes, 0);
    c.loose_h.assign(desc->n_hittables, -1);
    c.loose_n.assign(desc->n_bvh_nodes, -1);

    // validate references once so the emitters can index freely
    for (int i = 0; i < desc->n_hittables && !c.status; ++i) {
        const rt_hittable_desc& h = desc->hittables[i];
        auto bad = [&](const char* m) { c.fail(RT_ERR_OUT_OF_RANGE, m); };
        switch (h.kind) {
            case RT_HIT_SPHERE: case RT_HIT_QUAD:
                if (h.mat < 0 || h.mat >= desc->n_materials) bad("primitive references an unknown material");
                break;
            case RT_HIT_LIST:
                if (h.count < 0 || h.child < 0 || h.child + h.count > desc->n_list_items) bad("list range out of bounds");
                else for (int k = 0; k < h.count; ++k) { const int it = desc->list_items[h.child + k]; if (it < 0 || it >= i) { bad("list item must be an earlier hittable"); break; } }
                break;
            case RT_HIT_TRANSLATE: case RT_HIT_ROTATE_Y:
                if (h.child < 0 || h.child >= i) bad("instance child must be an earlier hittable");
                break;
            case RT_HIT_CONSTANT_MEDIUM:
                if (h.child < 0 || h.child >= i) bad("medium boundary must be an earlier hittable");
                else if (h.mat < 0 || h.mat >= desc->n_materials) bad("medium references an unknown material");
                break;
            case RT_HIT_BVH:
                if (h.child < 0 || h.child >= desc->n_bvh_nodes) bad("bvh root out of bounds");
                break;
            default: c.fail(RT_ERR_INVALID_ARGUMENT, "unknown hittable kind");
        }
    }
    for (int i = 0; i < desc->n_bvh_nodes && !c.status; ++i) {
        const rt_bvh_node_desc& n = desc->bvh_nodes[i];
        if (n.object >= 0) { if (n.object >= desc->n_hittables) c.fail(RT_ERR_OUT_OF_RANGE, "bvh leaf references an unknown hittable"); }
        else if (n.left <= i || n.right <= i || n.left >= desc->n_bvh_nodes || n.right >= desc->n_bvh_nodes)
            c.fail(RT_ERR_OUT_OF_RANGE, "bvh children must follow their parent (pre-order)");
    }
    for (int i = 0; i < desc->n_materials && !c.status; ++i) {
        const rt_material_desc& m = desc->materials[i];
        const bool needs_tex = m.kind == RT_MAT_LAMBERTIAN || m.kind == RT_MAT_DIFFUSE_LIGHT || m.kind == RT_MAT_ISOTROPIC;
        if (needs_tex && (m.tex < 0 || m.tex >= desc->n_textures)) c.fail(RT_ERR_OUT_OF_RANGE, "material references an unknown texture");
    }
    for (int i = 0; i < desc->n_textures && !c.status; ++i) {
        const rt_texture_desc& t = desc->textures[i];
        if (t.kind == RT_TEX_CHECKER && (t.a < 0 || t.a >= i || t.b < 0 || t.b >= i)) c.fail(RT_ERR_OUT_OF_RANGE, "checker children must be earlier textures");
        if (t.kind == RT_TEX_IMAGE && (t.a < 0 || t.a >= desc->n_images)) c.fail(RT_ERR_OUT_OF_RANGE, "texture references an unknown image");
        if (t.kind == RT_TEX_NOISE && (t.a < 0 || t.a >= desc->n_perlins)) c.fail(RT_ERR_OUT_OF_RANGE, "texture references an unknown perlin table");
    }
    if (!c.status && (desc->world < 0 || desc->world >= desc->n_hittables)) c.fail(RT_ERR_OUT_OF_RANGE, "world id out of range");

    if (!c.status) c.emit(desc->world, false);
    if (!c.status && out->ops.empty()) {  // empty world: one unhittable node keeps the kernels branch-free
        const float inf = std::numeric_limits<float>::infinity();
        out->ops.push_back(F4{0.0f, 0.0f, 0.0f, bits_to_float(make_hdr(OP_INNER, 0, 2))});
        out->ops.push_back(F4{-inf, -inf, -inf, int_to_float_bits(2)});
    }
    if (c.status) { msg = c.err; if (err) *err = msg.c_str(); return c.status; }
    // The reference's per-axis box test (OP_INNER_REF, kept for boxes a skewed quad sticks out of) compares every axis with
    // the CURRENT interval on its own, so what it lets through depends on how far the interval has been narrowed when the
    // node is reached: a medium evaluated before the traversal instead of at its place in the tree can hide such a quad
    // (found by the generated scenes of tools/fuzz_scenes.py). Scenes with such boxes keep their media in the stream.
    if (c.has_reference_boxes && !c.hoisted_at.empty()) {
        CompileOptions in_place = opt;
        in_place.hoist_media = false;
        *out = CompiledScene();
        return compile_scene(desc, in_place, out, err);
    }
    if (opt.prune_boxes) prune_stream(&out->ops, opt);

    // links (dev_scene.h): until here every skip is a word index and no header carries a class. This pass checks every
    // successor, writes the fall-through class into the headers and turns the skips into complete link words.
    {
        std::vector<F4>& ops = out->ops;
        if (ops.empty()) {   // nothing but hoisted media (or nothing at all): keep one unhittable node
            const float inf = std::numeric_limits<float>::infinity();
            ops.push_back(F4{0.0f, 0.0f, 0.0f, bits_to_float(make_hdr(OP_INNER, 0, 2))});
            ops.push_back(F4{-inf, -inf, -inf, int_to_float_bits(2)});
        }
        const int n = (int)ops.size();
        if ((uint64_t)(n + c.hoisted.size() + 2) * 16u >= (uint64_t)kSlabLimit) {
            msg = "scene too large: the op stream must stay below 256 MiB"; if (err) *err = msg.c_str(); return RT_ERR_UNSUPPORTED;
        }
        auto hdr_of = [&](int i) { uint32_t u; std::memcpy(&u, &ops[i].w, 4); return u; };
        auto int_of = [&](float f) { int32_t v; std::memcpy(&v, &f, 4); return v; };
        auto cls_at = [&](int i) -> uint32_t { return i >= n ? (uint32_t)CLS_SHADE : class_of_kind(hdr_kind(hdr_of(i))); };
        int i = 0;
        while (i < n) {
            const uint32_t hdr = hdr_of(i);
            const uint32_t kind = hdr_kind(hdr), flags = hdr_flags(hdr);
            if (kind > OP_INNER_REF) { msg = "internal: bad op kind in stream"; if (err) *err = msg.c_str(); return RT_ERR_INTERNAL; }
            const int size = op_words(kind, flags);
            const bool has_skip = kind == OP_INNER || kind == OP_INNER_REF || kind == OP_XFORM_ENTER;
            const int skip = has_skip ? int_of(ops[i + 1].w) : -1;
            int ft = i + size;
            if (kind == OP_MEDIUM && (int)flags == MEDIUM_BOUNDARY_PROGRAM) ft = int_of(ops[i + 1].y);   // run past the inline program
            // the kernels follow these links without bounds checks: every successor must lie in (i, n]
            if ((int)hdr_words(hdr) != size || i + size > n || ft <= i || ft > n || (has_skip && (skip <= i || skip > n))) {
                msg = "internal: op stream link out of range"; if (err) *err = msg.c_str(); return RT_ERR_INTERNAL;
            }
            if (kind == OP_MEDIUM && (int)flags == MEDIUM_BOUNDARY_PROGRAM && (int_of(ops[i + 1].x) != i + 3 || ft < i + 3)) {
                msg = "internal: medium boundary program out of range"; if (err) *err = msg.c_str(); return RT_ERR_INTERNAL;
            }
            const uint32_t nh = (hdr & 0x0fffffffu) | (cls_at(ft) << 28);
            std::memcpy(&ops[i].w, &nh, 4);
            if (has_skip) ops[i + 1].w = bits_to_float(make_link(skip, cls_at(skip)));
            if (kind == OP_BOX) ops[i + 1].w = bits_to_float(make_link(ft, cls_at(ft)));   // a missed cube goes on like a rejected cull box
            i += size;
        }
        if (i != n) { msg = "internal: op stream walk ended off the end"; if (err) *err = msg.c_str(); return RT_ERR_INTERNAL; }
        out->first_link = make_link(0, cls_at(0));
        out->n_world_words = n;
        for (int32_t off : c.hoisted_at) out->hoisted_media.push_back(n + off);
        ops.insert(ops.end(), c.hoisted.begin(), c.hoisted.end());
        ops.push_back(F4{0.0f, 0.0f, 0.0f, 0.0f});   // padding: the render kernel fetches two words at the cursor
        ops.push_back(F4{0.0f, 0.0f, 0.0f, 0.0f});   // unconditionally, including at the end of the world program
    }

    for (int i = 0; i < desc->n_materials; ++i) {
        const rt_material_desc& m = desc->materials[i];
        out->materials.push_back(F4{int_to_float_bits(m.kind), int_to_float_bits(m.tex), (float)m.param, 0.0f});
        out->materials.push_back(F4{(float)m.albedo[0], (float)m.albedo[1], (float)m.albedo[2], 0.0f});
    }
    for (int i = 0; i < desc->n_textures; ++i) {
        const rt_texture_desc& t = desc->textures[i];
        out->textures.push_back(F4{int_to_float_bits(t.kind), int_to_float_bits(t.a), int_to_float_bits(t.b), (float)t.scale});
        out->textures.push_back(F4{(float)t.color[0], (float)t.color[1], (float)t.color[2], 0.0f});
    }
    out->n_perlin = desc->n_perlins;
    for (int i = 0; i < desc->n_perlins; ++i) {
        const rt_perlin_desc& p = desc->perlins[i];
        for (int k = 0; k < 256; ++k) out->perlin_vec.push_back(F4{(float)p.ranvec[k][0], (float)p.ranvec[k][1], (float)p.ranvec[k][2], 0.0f});
        for (int k = 0; k < 256; ++k) out->perlin_perm.push_back((uint8_t)p.perm_x[k]);
        for (int k = 0; k < 256; ++k) out->perlin_perm.push_back((uint8_t)p.perm_y[k]);
        for (int k = 0; k < 256; ++k) out->perlin_perm.push_back((uint8_t)p.perm_z[k]);
    }
    out->scene_scale = (float)c.scale;
    return 0;
}

}  // namespace rtdev
