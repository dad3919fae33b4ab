"""The flattening, checked without a GPU: the op stream scene_compile.cpp produces (rt_scene_ops_export) is walked on
the host by tests/opstream.py - op for op what traverse<>() in rt_kernels.cuh does - and its closest hits must be the
f64 oracle's: same hit / miss, same primitive (up to exact ties between coplanar faces, where the oracle itself decides
by rounding noise), same t up to the f32 rounding of the stream's geometry. Covers skip links, visiting order, the
interval rules, instance folding, hoisted and in-stream media, Quad::cube slab primitives and OP_INNER_REF nodes.
"""
import os
import sys

import numpy as np
import pytest

from conftest import small_scene
import opstream

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
from gpu_probe import make_rays  # noqa: E402
from fuzz_scenes import random_scene  # noqa: E402

T_REL = 2e-5       # stream geometry is f32 (centres, radii, plane offsets, instance sin / cos), the walk itself f64


def compare(emu, ref, max_flips=0):
    flips = int((emu["hit"] != ref["hit"]).sum())
    assert flips <= max_flips, f"{flips} hit/miss flips"
    both = (emu["hit"] == 1) & (ref["hit"] == 1)
    if not both.any():
        return
    t_err = np.abs(emu["t"] - ref["t"])[both] / np.maximum(1.0, np.abs(ref["t"][both]))
    assert t_err.max() <= T_REL, t_err.max()
    other = emu["prim_id"][both] != ref["prim_id"][both]
    assert int((other & (t_err > 1e-9)).sum()) == 0            # a different primitive only as an exact tie in t
    assert int(other.sum()) <= max(2, int(both.sum()) // 200)


def random_rays(rt, rng, n, extent=14.0):
    rays = np.zeros(n, dtype=rt._abi.ray_dtype())
    rays["origin"] = rng.uniform(-extent, extent, (n, 3))
    d = rng.normal(size=(n, 3))
    rays["direction"] = d / np.linalg.norm(d, axis=1, keepdims=True) * rng.uniform(0.5, 2, (n, 1))
    rays["time"] = rng.random(n)
    return rays


@pytest.mark.parametrize("idx", range(9))
def test_stream_walk_matches_oracle_cli_scenes(rt, ob, earth, idx):
    s, cam = small_scene(rt, idx, earth)
    S = opstream.Stream(rt.scene_ops(s))
    rays = make_rays(cam, s.desc, 1 << 13, seed=7)
    counts = {}
    emu = opstream.hit_batch(S, rays, seed=7, counts=counts)
    compare(emu, ob.hit_batch(s.desc, rays, seed=7), max_flips=1)
    L = rt.scene_layout(s)
    assert (counts.get("box", 0) > 0) == (L["n_box"] > 0) and (counts.get("medium", 0) > 0) == (L["n_medium_hoisted"] + L["n_medium_in_stream"] > 0)


@pytest.mark.parametrize("seed", range(10))
def test_stream_walk_matches_oracle_random_scenes(rt, ob, seed):
    """Generated scenes (tools/fuzz_scenes.py). Most of their quads stick out of the diagonal box Quad::new gives them
    (quad.rs:41-43), so which part of them a BVH can see is decided by the reference's own per-axis box test."""
    s = random_scene(1000 + seed)
    S = opstream.Stream(rt.scene_ops(s))
    rays = random_rays(rt, np.random.default_rng(seed), 1 << 14)
    compare(opstream.hit_batch(S, rays, seed=seed), ob.hit_batch(s.desc, rays, seed=seed), max_flips=1)


RICH_SEEDS = [0, 1, 2, 5, 6, 11, 12, 14, 16, 18, 20, 21, 23, 24, 26, 36]


@pytest.mark.parametrize("seed", RICH_SEEDS)
def test_stream_walk_matches_oracle_rich_scenes(rt, ob, seed):
    """tools/fuzz_scenes.py::rich_scene: textured media, media inside instances and inside instanced groups (sphere, moving
    sphere, rotated cube and - seeds 2, 3, 11, 14, 16, 20, 23 - a rotated cube or an instanced group as the boundary of a
    medium that itself sits inside an instance: the boundary program then composes its instances from its own frame), a
    1000-unit ground sphere, a fog sphere around everything. Seed 24: a skewed quad under reference boxes (OP_INNER_REF) next
    to world-space media - what the reference's per-axis box test lets through depends on how far the interval has been
    narrowed when the node is reached, so such a scene keeps its media at their place in the stream (no hoisting)."""
    from fuzz_scenes import rich_scene
    s = rich_scene(5000 + seed)
    S = opstream.Stream(rt.scene_ops(s))
    rays = random_rays(rt, np.random.default_rng(seed), 1 << 13)
    compare(opstream.hit_batch(S, rays, seed=seed), ob.hit_batch(s.desc, rays, seed=seed), max_flips=1)


def test_reference_nodes_decide_what_a_skewed_quad_shows(rt, ob):
    """One skewed quad in a BVH: the reference culls the part outside its diagonal box (mostly - the per-axis test lets
    some of it through). The stream must reproduce the oracle exactly; a tight, geometrically complete box would not."""
    s = rt.Scene(bvh_seed=3)
    m = s.Lambertian(s.SolidColor(0.5, 0.5, 0.5))
    l = rt.HittableList()
    l.add(s.Quad((-2, -2, 0), (4, 0, 3), (0, 4, -3), m))       # corners q+u and q+v stick out in z
    l.add(s.Sphere((9, 0, 0), 1.0, m))
    s.finish(s.BVHNode(l))
    S = opstream.Stream(rt.scene_ops(s))
    kinds = np.array([k for _, k, _ in S.walk()])
    assert (kinds == opstream.OP_INNER_REF).sum() >= 2          # root and the quad's own leaf box
    rng = np.random.default_rng(1)
    n = 1 << 14
    rays = np.zeros(n, dtype=rt._abi.ray_dtype())
    rays["origin"] = np.column_stack([np.full(n, -9.0), rng.uniform(-3, 3, n), rng.uniform(-3, 3, n)])      # from the side:
    rays["direction"] = np.column_stack([np.full(n, 1.0), rng.uniform(-0.1, 0.1, n), rng.uniform(-0.1, 0.1, n)])   # the z slab decides
    ref = ob.hit_batch(s.desc, rays)
    compare(opstream.hit_batch(S, rays), ref)
    # the plain quad (a list world has no boxes, hittable.rs:61-79) is hit by more of these rays than the BVH one
    s2 = rt.Scene(bvh_seed=3)
    m2 = s2.Lambertian(s2.SolidColor(0.5, 0.5, 0.5))
    l2 = rt.HittableList()
    l2.add(s2.Quad((-2, -2, 0), (4, 0, 3), (0, 4, -3), m2))
    s2.finish(s2.List(l2))
    full = ob.hit_batch(s2.desc, rays)
    assert int(full["hit"].sum()) > int(ref["hit"].sum()) > 0
    compare(opstream.hit_batch(opstream.Stream(rt.scene_ops(s2)), rays), full)


def test_stream_walk_instances_media_and_nesting(rt, ob):
    """Wrappers in orders the CLI scenes do not use, media with a moving-sphere / rotated-cube / generic boundary,
    a medium inside an instance (stays in the stream), BVH in BVH, a list world."""
    rng = np.random.default_rng(8)
    s = rt.Scene(bvh_seed=5)
    white = s.Lambertian(s.SolidColor(0.7, 0.7, 0.7))
    glass = s.Dielectric(1.5)
    world = rt.HittableList()
    world.add(s.Translate(s.Sphere((0, 0, 0), 1.0, white), (4, 0, 0)))
    world.add(s.RotateY(s.cube((-1, -1, -1), (1, 2, 1), white), 30.0))
    world.add(s.RotateY(s.Translate(s.Quad((0, 0, 0), (2, 0, 0), (0, 2, 0), glass), (0, 3, 1)), -40.0))
    inner = rt.HittableList()
    for _ in range(40):
        c = rng.uniform(-3, 3, 3)
        inner.add(s.Sphere(c, 0.4, white, target=c + rng.uniform(-0.5, 0.5, 3)))
    world.add(s.Translate(s.RotateY(s.BVHNode(inner), 75.0), (-8, 0, -2)))
    nested = rt.HittableList()
    nested.add(s.BVHNode(inner))
    nested.add(s.Sphere((0, 8, 0), 2.0, glass))
    world.add(s.BVHNode(nested))
    c0 = np.array([8.0, 4.0, 0.0])
    world.add(s.ConstantMedium(s.Sphere(c0, 2.0, glass, target=c0 + (0, 1, 0)), 0.7, (1, 1, 1)))
    world.add(s.ConstantMedium(s.Translate(s.RotateY(s.cube((0, 0, 0), (3, 3, 3), white), 20.0), (-4, -6, 0)), 0.9, (0.2, 0.2, 0.2)))
    two = rt.HittableList()
    two.add(s.Sphere((0, -9, 6), 1.5, white))
    two.add(s.cube((-1, -11, 3), (1, -8, 5), white))
    world.add(s.ConstantMedium(s.BVHNode(two), 1.1, (0.5, 0.5, 0.5)))                                   # generic boundary
    world.add(s.Translate(s.ConstantMedium(s.Sphere((0, 0, 0), 1.5, glass), 0.8, (1, 1, 1)), (9, -7, -5)))   # medium in an instance
    s.finish(s.List(world))
    L = rt.scene_layout(s)
    assert L["n_medium_in_stream"] == 2 and L["n_medium_hoisted"] == 2
    S = opstream.Stream(rt.scene_ops(s))
    rays = random_rays(rt, rng, 1 << 15)
    ref = ob.hit_batch(s.desc, rays, seed=3)
    assert len(set(ref["prim_id"][ref["hit"] == 1])) > 30
    compare(opstream.hit_batch(S, rays, seed=3), ref, max_flips=1)


def test_stream_walk_nested_instances(rt, ob):
    """Instances inside other instances' subtrees, three deep (composed transforms, exits that return to the parent)."""
    from test_gpu_hits import nested_instances_scene
    rng = np.random.default_rng(21)
    s = nested_instances_scene(rt, rng)
    S = opstream.Stream(rt.scene_ops(s))
    rays = random_rays(rt, rng, 1 << 15)
    ref = ob.hit_batch(s.desc, rays, seed=4)
    assert len(set(ref["prim_id"][ref["hit"] == 1])) > 12
    compare(opstream.hit_batch(S, rays, seed=4), ref, max_flips=1)


@pytest.mark.parametrize("idx", [0, 6, 7, 8])
def test_pruning_changes_no_hit(rt, earth, idx):
    """prune_stream only removes cull boxes: walked on the same rays, the pruned and the unpruned stream must give
    bit-identical closest hits (same t, same primitive), and the pruned one must test fewer boxes."""
    s, cam = small_scene(rt, idx, earth)
    rays = make_rays(cam, s.desc, 1 << 13, seed=11)
    pruned = opstream.Stream(rt.scene_ops(s))
    full = opstream.Stream(rt.scene_ops(s, rt.layout_flags(prune=False)))
    ca, cb = {}, {}
    a = opstream.hit_batch(pruned, rays, seed=11, counts=ca)
    b = opstream.hit_batch(full, rays, seed=11, counts=cb)
    assert np.array_equal(a["hit"], b["hit"]) and np.array_equal(a["prim_id"], b["prim_id"]) and np.array_equal(a["t"], b["t"])
    assert ca.get("inner", 0) < cb.get("inner", 0)
    for k in ("sphere", "quad", "box"):          # leaves are tested at least as often: that is the trade the cost model makes
        assert ca.get(k, 0) >= cb.get(k, 0)


def test_ops_export_argument_checks(rt):
    import ctypes as C
    s, _ = small_scene(rt, 1)
    lib = rt._abi.lib()
    n = C.c_int64()
    assert lib.rt_scene_ops_export(None, 0, None, 0, C.byref(n), None, None, None, None) == rt._abi.RT_ERR_INVALID_ARGUMENT
    assert lib.rt_scene_ops_export(C.byref(s.desc), 0, None, 0, None, None, None, None, None) == rt._abi.RT_ERR_INVALID_ARGUMENT
    assert lib.rt_scene_ops_export(C.byref(s.desc), 0, None, 0, C.byref(n), None, None, None, None) == 0 and n.value > 2
    buf = (C.c_float * 8)()
    assert lib.rt_scene_ops_export(C.byref(s.desc), 0, buf, 2, C.byref(n), None, None, None, None) == rt._abi.RT_ERR_OUT_OF_RANGE


@pytest.mark.parametrize("seed", range(100, 124))
def test_pruning_changes_no_hit_random_scenes(rt, seed):
    """The same invariant over generated scenes (nested instances, skewed quads whose OP_INNER_REF nodes must survive,
    media with generic boundaries): pruned and unpruned streams give bit-identical closest hits."""
    s = random_scene(seed)
    rays = random_rays(rt, np.random.default_rng(seed), 1 << 12)
    pruned = opstream.Stream(rt.scene_ops(s))
    full = opstream.Stream(rt.scene_ops(s, rt.layout_flags(prune=False)))
    a = opstream.hit_batch(pruned, rays, seed=seed)
    b = opstream.hit_batch(full, rays, seed=seed)
    assert np.array_equal(a["hit"], b["hit"]) and np.array_equal(a["prim_id"], b["prim_id"]) and np.array_equal(a["t"], b["t"])
    assert pruned.n_world <= full.n_world


@pytest.mark.parametrize("seed", [0, 2, 6, 11, 14, 16, 24, 36])
def test_layout_switches_change_no_hit_rich_scenes(rt, seed):
    """rich_scene (media inside instances and instanced boundaries, reference boxes next to media, f64 spheres): the pruned
    and the unpruned stream, and the stream with every medium left at its place in the tree, give the same closest hits
    (bit-identical for pruning: boxes only cull; hoisting moves a medium's test, not its draw - the draw is keyed)."""
    from fuzz_scenes import rich_scene
    s = rich_scene(5000 + seed)
    rays = random_rays(rt, np.random.default_rng(seed), 1 << 12)
    a = opstream.hit_batch(opstream.Stream(rt.scene_ops(s)), rays, seed=seed)
    b = opstream.hit_batch(opstream.Stream(rt.scene_ops(s, rt.layout_flags(prune=False))), rays, seed=seed)
    c = opstream.hit_batch(opstream.Stream(rt.scene_ops(s, rt.layout_flags(hoist_media=False))), rays, seed=seed)
    assert np.array_equal(a["hit"], b["hit"]) and np.array_equal(a["prim_id"], b["prim_id"]) and np.array_equal(a["t"], b["t"])
    assert np.array_equal(a["hit"], c["hit"]) and np.array_equal(a["prim_id"], c["prim_id"]) and np.array_equal(a["t"], c["t"])


@pytest.mark.parametrize("seed", range(6))
def test_stream_walk_deeply_nested_instances(rt, ob, seed):
    """Five levels of Translate / RotateY around lists and BVHs: the composed world -> local transforms and the parent
    links of the exits must reproduce the oracle's recursive Translate::hit / RotateY::hit (hittable.rs:96-193)."""
    from test_gpu_hits import deeply_nested_scene
    s, rng = deeply_nested_scene(rt, seed)
    assert rt.scene_layout(s)["n_xform"] == 5
    S = opstream.Stream(rt.scene_ops(s))
    rays = random_rays(rt, rng, 1 << 14, extent=10.0)
    ref = ob.hit_batch(s.desc, rays, seed=1)
    assert int(ref["hit"].sum()) > 1000
    compare(opstream.hit_batch(S, rays, seed=1), ref, max_flips=1)
