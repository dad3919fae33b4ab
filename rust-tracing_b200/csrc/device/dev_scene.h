// Device scene layout: one "threaded" op stream of float4 words.
//
// The reference walks Box<(Node,AABB)> trees recursively, always left child first
// (bvh.rs:90-113), with trait-object leaves (hittable.rs). Here the whole hittable graph —
// nested BVHs, lists, Translate/RotateY instances, media — is flattened at upload into a single
// array of variable-length ops laid out in that same depth-first, left-first order. Every op that
// can be culled carries a skip link (the word index just past its subtree), so traversal is a
// stackless loop `i = hit_box ? i + size : skip` that visits nodes in exactly the reference's
// order, which also preserves its tie rules on equal t (sphere: open interval, earlier wins;
// quad / medium: closed interval, later wins — sphere.rs:78, quad.rs:115).
//
// Word 0 of every op: xyz = payload, w = header. All ops are >= 2 words so the first two words can be fetched together.
//
// LINKS. The traversal cursor is a "link": byte offset of the op in the stream (word index * 16) | class of that op
// << 28. The slab class is 0, so for a lane that is at a box-headed op the link IS the address offset and the test
// "am I in the slab class" is one unsigned compare. Skip links are stored as complete link words; the fall-through link
// is `link + (hdr & kHdrFallThrough)`: the header's low byte is the op's size in bytes and its top nibble the class
// of the op that follows, and for OP_INNER (kind 0, flags 0) the whole header is that increment.
//
// CULL BOXES are stored as centre / half extent {c.xyz} {h.xyz}: with the per-ray constants inv = 1/d and
// oi = -o * inv the slab test is  tc = fma(c, inv, oi); th = h * inv; near = tc - |th|; far = tc + |th|  - twelve FMA-pipe
// instructions and no per-node sign selects (the ALU pipe, which runs at half rate, is what the traversal loop is
// bound by; profiles/r2_k1_region_breakdown.md). tc suffers cancellation that the classic (lo - o) * inv form does not,
// |err| <= ~2^-23 (|o| + |c| + h) |inv| per axis, so (i) h is padded by 2^-21 (|c| + h) at compile time and (ii) the
// kernel adds the per-ray term eps = 2^-21 max_k |o_k inv_k| (finite axes only) to the exit side of the comparison.
// Boxes only cull, so the test must never reject a box the exact test would pass; it may pass a few more.
#pragma once
#include <cstdint>
#include <vector>

#include "../../../include/rt_b200.h"

namespace rtdev {

enum OpKind : uint32_t {
    OP_INNER = 0,        // w0 = {c.xyz, hdr}   w1 = {h.xyz, skip link}                  size 2   BVH node / list box, padded
    OP_SPHERE = 1,       // w0 = {c.xyz, hdr}   w1 = {r, mat, prim_id, precise_idx}      size 2 (+1 if moving: w2 = {center_vec.xyz, 0})
    OP_QUAD = 2,         // w0 = {n.xyz, hdr}   w1 = {A.xyz, a0} w2 = {B.xyz, b0} w3 = {d, mat, prim_id, 0}   size 4
                         //   alpha = A.p + a0, beta = B.p + b0 with A = v x w, B = w x u (scalar triple product form of quad.rs:121-122)
    OP_XFORM_ENTER = 3,  // w0 = {c.xyz, hdr}   w1 = {h.xyz, skip link} w2 = {a.xyz, sin} w3 = {b.xyz, cos}       size 4
                         //   local = R(x - a) + b, R = rotate-y (hittable.rs:164-168), x = the WORLD ray: the transform is the
                         //   composition with every enclosing instance; the box (padded, centre / half extent) is in the
                         //   enclosing space; skip = the op after the matching exit
    OP_XFORM_EXIT = 4,   // w0 = {parent OP_XFORM_ENTER (word index, -1 = world space), 0, 0, hdr}   w1 = {0,0,0,0}   size 2, FLAG_ALWAYS
    OP_MEDIUM = 5,       // body of a ConstantMedium; always preceded by an OP_INNER holding its box (skip = past the medium)
                         // w0 = {neg_inv_density, mat, prim_id, hdr(flags = boundary kind)}
                         //   boundary sphere : w1 = {c.xyz, r}  w2 = {center_vec.xyz, precise_idx | aux<<24}             size 3
                         //   boundary program: w1 = {bbegin, bend (word indices), 0, 0} w2 = {0}; the program follows inline,
                         //                     the fall-through successor is the op at bend                              size 3
                         //   boundary xbox   : w1 = {a.xyz, sin} w2 = {b.xyz, cos} w3 = {min.xyz, 0} w4 = {max.xyz, 0}   size 5
    OP_BOX = 6,          // a Quad::cube list as ONE slab primitive:                                                      size 4
                         //   w0 = {c.xyz, hdr} w1 = {h.xyz, fall-through link} (centre / half extent of the exact corners, NOT padded;
                         //   the link makes "the ray misses this box" the same arithmetic as a rejected OP_INNER)
                         //   w2 = {min.xyz, first_quad_prim_id} w3 = {max.xyz, mat} (the exact corners: hit record, self-origin rule)
                         //   faces in quad.rs:45-93 order: 0 +z, 1 +x, 2 -z, 3 -x, 4 +y, 5 -y ; prim_id = first + face
    OP_INNER_REF = 7,    // w0 = {lo.xyz, hdr} w1 = {hi.xyz, skip link}: the reference's own node box and aabb.rs:64-84 verbatim
                         //   (per axis, never narrowed). Emitted for every BVH node - leaves included, bvh.rs:92 tests them too -
                         //   whose subtree holds a quad that sticks out of its own bounding box: Quad::new boxes only the
                         //   diagonal q .. q+u+v (quad.rs:41-43), so for a parallelogram that is not axis aligned the reference
                         //   culls part of the quad, and which part depends on exactly this test. Tight culling would differ.
};

constexpr uint32_t FLAG_MOVING = 1u;   // sphere has center_vec
constexpr uint32_t FLAG_PRECISE = 2u;  // sphere test runs in f64 (huge radius; SURVEY.md §7 "hard parts")
constexpr uint32_t FLAG_ALWAYS = 8u;   // box-headed op whose own code runs whatever the box arithmetic says (OP_XFORM_EXIT, OP_INNER_REF)

constexpr int MEDIUM_BOUNDARY_SPHERE = 0;
constexpr int MEDIUM_BOUNDARY_PROGRAM = 1;
constexpr int MEDIUM_BOUNDARY_XBOX = 2;
constexpr int kMaxHoistedMedia = 8;   // [Translate/RotateY chain of] a cube: entry/exit from one slab test in the cube's frame

// header bits: [0,8) size of the op in bytes, [8,12) kind, [12,16) flags, [28,32) class of the fall-through successor
// (for OP_MEDIUM with a boundary program: the class of the op at bend).
enum OpClass : uint32_t { CLS_SLAB = 0, CLS_SPHERE = 1, CLS_QUAD = 2, CLS_MEDIUM = 3, CLS_SHADE = 4, CLS_IDLE = 5 };
constexpr uint32_t kLinkMask = 0x0fffffffu;        // byte offset part of a link
constexpr uint32_t kHdrFallThrough = 0xf00000ffu;  // size in bytes | successor class << 28
constexpr uint32_t kHdrNotInner = 0x0fffff00u;     // zero for OP_INNER: then link + hdr is the fall-through link
constexpr uint32_t kHdrAlwaysRare = FLAG_ALWAYS << 12;
constexpr uint32_t kSlabLimit = 1u << 28;          // link < kSlabLimit  <=>  the lane is in the slab class
inline uint32_t make_hdr(uint32_t kind, uint32_t flags, uint32_t size_words) { return (size_words * 16u) | (kind << 8) | (flags << 12); }
inline uint32_t hdr_kind(uint32_t hdr) { return (hdr >> 8) & 15u; }
inline uint32_t hdr_flags(uint32_t hdr) { return (hdr >> 12) & 15u; }
inline uint32_t hdr_words(uint32_t hdr) { return (hdr & 0xffu) >> 4; }
inline uint32_t make_link(int word, uint32_t cls) { return ((uint32_t)word << 4) | (cls << 28); }
inline int link_word(uint32_t link) { return (int)((link & kLinkMask) >> 4); }
inline uint32_t class_of_kind(uint32_t kind) {
    return kind == OP_SPHERE ? CLS_SPHERE : kind == OP_QUAD ? CLS_QUAD : kind == OP_MEDIUM ? CLS_MEDIUM : CLS_SLAB;
}
inline int op_words(uint32_t kind, uint32_t flags) {
    switch (kind) {
        case OP_SPHERE: return (flags & FLAG_MOVING) ? 3 : 2;
        case OP_QUAD: case OP_XFORM_ENTER: case OP_BOX: return 4;
        case OP_MEDIUM: return (int)flags == MEDIUM_BOUNDARY_XBOX ? 5 : 3;
        default: return 2;   // OP_INNER, OP_INNER_REF, OP_XFORM_EXIT
    }
}

struct F4 { float x, y, z, w; };
struct D4 { double x, y, z, w; };

// Output of compile_scene (host memory), uploaded verbatim.
struct CompiledScene {
    std::vector<F4> ops;         // the op stream: world program = [0, n_world_words), then the hoisted media bodies
    int n_world_words = 0;
    std::vector<int32_t> hoisted_media;   // word indices (>= n_world_words) of OP_MEDIUM bodies evaluated at segment start
    std::vector<F4> materials;   // 2 words each: {kind, tex, param, 0} {albedo.xyz, 0}
    std::vector<F4> textures;    // 2 words each: {kind, a, b, scale} {color.xyz, 0}
    std::vector<F4> perlin_vec;  // 256 per table: {ranvec.xyz, 0}
    std::vector<uint8_t> perlin_perm;  // 768 per table: perm_x | perm_y | perm_z
    std::vector<D4> precise;     // per precise sphere: {c.xyz, r} {center_vec.xyz, 0}
    int n_perlin = 0;
    // word index of the op that starts each BVH hittable (for rt_bvh_export)
    std::vector<int32_t> bvh_hittable_ids;
    std::vector<std::vector<int32_t>> bvh_preorder_objects;
    float scene_scale = 1.0f;    // largest |coordinate| of finite geometry (parity tolerances)
    uint32_t first_link = 0;     // link of op 0 (offset 0 | its class << 28)
};

struct CompileOptions {
    bool box_primitives = true;   // false: emit cube lists as 6 quads (the reference's own structure), for A/B parity runs
    bool hoist_media = true;      // false: media stay in the op stream at their BVH position
    bool prune_boxes = true;      // drop cull boxes (OP_INNER) that cost more tests than they save (scene_compile.cpp, prune_stream)
    // prune_stream's cost of one leaf op relative to one OP_INNER test (what a lane-op of that kind costs the warp)
    double cost_sphere = 2.5, cost_quad = 1.5, cost_box = 6.0, cost_medium = 6.0, cost_xform = 3.0;
};

// Returns 0 or a negative rt_status; message in *err.
int compile_scene(const rt_scene_desc* desc, const CompileOptions& opt, CompiledScene* out, const char** err);

}  // namespace rtdev
