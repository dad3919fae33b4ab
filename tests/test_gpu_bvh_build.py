"""BVHNode::node_from_list (bvh.rs:31-66) with the sorting on the device. The node array must be
the host build's, bit for bit: same leaf order, same topology, same axis per node, same boxes."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def nodes_of(rt, scene):
    d = scene.desc
    h = d.hittables[d.world]
    raw = C.string_at(C.addressof(d.bvh_nodes[h.child]), h.count * C.sizeof(rt._abi.BvhNodeDesc))
    return np.frombuffer(raw, dtype=np.dtype([("bbox", "f8", 6), ("left", "i4"), ("right", "i4"), ("object", "i4"), ("axis", "i4")]))


def build_both(rt, ctx, make_objects, seed):
    a = rt.Scene(bvh_seed=seed)
    a.finish(a.BVHNode(make_objects(a)))
    b = rt.Scene(bvh_seed=seed)
    b.finish(b.BVHNodeOnDevice(ctx, make_objects(b)))
    return a, b, nodes_of(rt, a), nodes_of(rt, b)


def same(na, nb):
    assert len(na) == len(nb)
    for f in ("left", "right", "object", "axis"):
        assert np.array_equal(na[f], nb[f]), f
    assert np.array_equal(na["bbox"], nb["bbox"])         # numerically equal (min / max of the same doubles)


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 7, 8, 31, 257, 1000, 5000])
def test_device_build_equals_host_build(rt, ctx, n):
    def make(s):
        rng = np.random.default_rng(n)
        m = s.Lambertian(s.SolidColor(0.5, 0.5, 0.5))
        l = rt.HittableList()
        for _ in range(n):
            l.add(s.Sphere(tuple(rng.uniform(-100, 100, 3)), float(rng.uniform(0.1, 5.0)), m))
        return l
    _, _, na, nb = build_both(rt, ctx, make, seed=2 + n)
    assert len(na) == 2 * n - 1
    same(na, nb)


def test_ties_keep_insertion_order_and_mixed_objects(rt, ctx):
    """Equal keys (a grid of cubes: many objects share a box minimum on every axis, main.rs:515-529) must come out in the
    order the host's stable sort leaves them; quads, cubes and instances carry their own boxes."""
    def make(s):
        m = s.Lambertian(s.SolidColor(0.5, 0.5, 0.5))
        l = rt.HittableList()
        for i in range(12):
            for j in range(12):
                l.add(s.cube((i * 10.0, 0.0, j * 10.0), (i * 10.0 + 10.0, 1.0 + (i * j) % 7, j * 10.0 + 10.0), m))
        l.add(s.Quad((5, 20, 5), (30, 0, 0), (0, 0, 30), m))
        l.add(s.Translate(s.RotateY(s.Sphere((0, 0, 0), 3.0, m), 30.0), (50, 10, 50)))
        return l
    _, _, na, nb = build_both(rt, ctx, make, seed=11)
    same(na, nb)


def test_cli_scenes_rebuilt_on_the_device(rt, ob, ctx, earth):
    """final_scene's three BVHs (400 cubes, 1000 spheres, the top level) from the device build: same nodes, hence the same
    flattened stream and the same hits."""
    s, cs = rt.builtin_scene(8, image_width=64, earth=earth)
    d = s.desc
    boxes_of = lambda h: np.array([list(d.hittables[d.bvh_nodes[h.child + k].object].bbox) for k in range(h.count)
                                   if d.bvh_nodes[h.child + k].object >= 0])
    lib = rt._abi.lib()
    checked = 0
    for i in range(d.n_hittables):
        h = d.hittables[i]
        if h.kind != rt._abi.RT_HIT_BVH:
            continue
        want = np.frombuffer(C.string_at(C.addressof(d.bvh_nodes[h.child]), h.count * 64),
                             dtype=np.dtype([("bbox", "f8", 6), ("left", "i4"), ("right", "i4"), ("object", "i4"), ("axis", "i4")]))
        # the leaves in the order the reference's in-place sort left them are NOT the insertion order; rebuild from the
        # insertion order instead: object ids ascending is how scenes.cpp adds them
        leaves = np.sort(want["object"][want["object"] >= 0])
        n = len(leaves)
        bb = np.ascontiguousarray([list(d.hittables[int(o)].bbox) for o in leaves], dtype=np.float64)
        # axis draws in pre-order of the calls = the axis fields of the nodes that drew one
        axes = np.ascontiguousarray(want["axis"][want["axis"] >= 0], dtype=np.int32)
        assert len(axes) == lib.rt_bvh_axis_draws(n)
        out = np.zeros(2 * n - 1, dtype=want.dtype)
        made = lib.rt_bvh_build_device(ctx._h, bb.ctypes.data, n, axes.ctypes.data, out.ctypes.data, None)
        assert made == 2 * n - 1
        got_obj = np.where(out["object"] >= 0, leaves[np.maximum(out["object"], 0)], -1)
        assert np.array_equal(got_obj, want["object"])
        assert np.array_equal(out["bbox"], want["bbox"])
        lr = want["object"] < 0
        assert np.array_equal(out["left"][lr] + h.child, want["left"][lr]) and np.array_equal(out["right"][lr] + h.child, want["right"][lr])
        checked += 1
    assert checked == 3
