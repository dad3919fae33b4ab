"""Host scene code (the product's flatten + BVH build) against the oracle's independent restatements:
BVH topology bit-exact on indices, every bounding box bit-exact, Camera::new bit-exact."""
import ctypes as C

import numpy as np
import pytest

from conftest import small_scene


def bvh_hittables(rt, desc):
    return [i for i in range(desc.n_hittables) if desc.hittables[i].kind == rt._abi.RT_HIT_BVH]


def preorder(desc, root):
    """(left, right, object, axis, bbox) of the subtree in pre-order with indices relative to root."""
    out = []

    def rec(n):
        node = desc.bvh_nodes[n]
        out.append((node.left, node.right, node.object, node.axis, tuple(node.bbox)))
        if node.object < 0:
            rec(node.left)
            rec(node.right)
    rec(root)
    return out


def leaf_inputs(rt, desc, bvh_id):
    """The object list a BVH was built from, in insertion order, with boxes; and its axis draws."""
    h = desc.hittables[bvh_id]
    nodes = preorder(desc, h.child)
    objs = sorted(n[2] for n in nodes if n[2] >= 0)   # ids were created in insertion order
    axes = [n[3] for n in nodes if n[3] >= 0]
    boxes = np.array([desc.hittables[o].bbox[:] for o in objs])
    return objs, boxes, axes, nodes


@pytest.mark.parametrize("idx", range(9))
def test_bvh_topology_matches_oracle_build(rt, ob, earth, idx):   # bvh.rs:31-66
    s, _ = small_scene(rt, idx, earth)
    d = s.desc
    for b in bvh_hittables(rt, d):
        objs, boxes, axes, nodes = leaf_inputs(rt, d, b)
        # ids are handed out in creation order, which is the insertion order of every list in main.rs
        ref = ob.bvh_build(boxes, axes)
        assert ref["axes_used"] == len(axes)
        base = d.hittables[b].child
        got_obj = np.array([n[2] for n in nodes])
        want_obj = np.array([objs[o] if o >= 0 else -1 for o in ref["object"]])
        if not np.array_equal(got_obj, want_obj):
            pytest.skip("insertion order != id order for this BVH; covered by test_bvh_topology_explicit")
        got_left = np.array([n[0] - base if n[0] >= 0 else -1 for n in nodes])
        got_right = np.array([n[1] - base if n[1] >= 0 else -1 for n in nodes])
        assert np.array_equal(got_left, ref["left"]) and np.array_equal(got_right, ref["right"])
        assert np.array_equal(np.array([n[4] for n in nodes]), ref["bbox"])


def test_bvh_topology_explicit(rt, ob):
    """A BVH over objects whose insertion order is known, with heavy ties (the 20x20 grid of final_scene
    shares minima on every axis, main.rs:515-529): ties keep insertion order; pairs swap unless strictly less."""
    rng = np.random.default_rng(5)
    s = rt.Scene(bvh_seed=99)
    m = s.Lambertian(s.SolidColor(0.5, 0.5, 0.5))
    lst = rt.HittableList()
    ids = []
    for i in range(7):
        for j in range(9):
            ids.append(s.Sphere((float(i), float(rng.integers(0, 3)), float(j % 3)), 0.25, m))
            lst.add(ids[-1])
    b = s.BVHNode(lst)
    s.finish(b)
    d = s.desc
    nodes = preorder(d, d.hittables[b].child)
    axes = [n[3] for n in nodes if n[3] >= 0]
    boxes = np.array([d.hittables[o].bbox[:] for o in ids])
    ref = ob.bvh_build(boxes, axes)
    assert len(nodes) == 2 * len(ids) - 1 == d.hittables[b].count
    assert [n[2] for n in nodes] == [ids[o] if o >= 0 else -1 for o in ref["object"]]
    assert [n[0] for n in nodes] == list(ref["left"]) and [n[1] for n in nodes] == list(ref["right"])
    assert np.array_equal(np.array([n[4] for n in nodes]), ref["bbox"])
    # balanced median split: depth = ceil(log2 n)
    depth = {0: 0}
    for k, n in enumerate(nodes):
        if n[2] < 0:
            depth[n[0]] = depth[n[1]] = depth[k] + 1
    assert max(depth.values()) == int(np.ceil(np.log2(len(ids))))


def test_span2_swaps_on_ties(rt):   # bvh.rs:45-49: comparator never returns Equal, so equal minima swap
    s = rt.Scene()
    m = s.Lambertian(s.SolidColor(0.5, 0.5, 0.5))
    a = s.Sphere((0, 0, 0), 1.0, m)
    b = s.Sphere((0, 0, 0), 1.0, m)
    l = rt.HittableList(); l.add(a); l.add(b)
    s.finish(s.BVHNode(l))
    n = s.desc.bvh_nodes
    assert (n[1].object, n[2].object) == (b, a)
    assert n[1].axis == -1 and n[2].axis == -1 and n[0].axis in (0, 1, 2)   # pair leaves draw no axis (bvh.rs:51-56)


@pytest.mark.parametrize("idx", range(9))
def test_bounding_boxes_bit_exact(rt, ob, earth, idx):
    s, _ = small_scene(rt, idx, earth)
    bbox_diff, derived_diff = ob.validate_scene(s.desc)
    assert bbox_diff == 0.0
    assert derived_diff < 1e-15


def test_list_box_contains_origin(rt):   # HittableList derives Default -> bbox starts at [0,0]^3 (hittable.rs:50-59)
    s = rt.Scene()
    m = s.Lambertian(s.SolidColor(0.5, 0.5, 0.5))
    c = s.cube((10, 10, 10), (20, 20, 20), m)
    s.finish(c)
    bb = s.desc.hittables[c].bbox
    assert bb[0] == 0.0 and bb[2] == 0.0 and bb[4] == 0.0 and bb[1] == 20.00005
    assert s.desc.hittables[c].flags & rt._abi.RT_FLAG_CUBE_LIST and s.desc.hittables[c].count == 6


def test_scene_sizes(rt, earth):   # BASELINE.md §2 / SURVEY §8(a) a7
    s, _ = small_scene(rt, 8, earth)
    d = s.desc
    counts = sorted(d.hittables[b].count for b in bvh_hittables(rt, d))
    assert counts == [21, 799, 1999]          # top 11 leaves, 400 boxes, 1000 spheres
    kinds = [d.hittables[i].kind for i in range(d.n_hittables)]
    A = rt._abi
    assert kinds.count(A.RT_HIT_QUAD) == 2401 and kinds.count(A.RT_HIT_SPHERE) == 1007
    assert kinds.count(A.RT_HIT_CONSTANT_MEDIUM) == 2
    s0, _ = small_scene(rt, 0)
    n_sph = [s0.desc.hittables[i].kind for i in range(s0.desc.n_hittables)].count(A.RT_HIT_SPHERE)
    assert 470 <= n_sph <= 489 and s0.desc.hittables[s0.desc.world].count == 2 * n_sph - 1
    s6, _ = small_scene(rt, 6)
    assert s6.desc.hittables[s6.desc.world].count == 15   # 8 leaves


def test_scene_seeds_are_deterministic(rt):
    a, _ = rt.builtin_scene(0, scene_seed=1)
    b, _ = rt.builtin_scene(0, scene_seed=1)
    c, _ = rt.builtin_scene(0, scene_seed=2)
    ha = bytes(C.string_at(a.desc.hittables, a.desc.n_hittables * C.sizeof(rt._abi.HittableDesc)))
    hb = bytes(C.string_at(b.desc.hittables, b.desc.n_hittables * C.sizeof(rt._abi.HittableDesc)))
    hc = bytes(C.string_at(c.desc.hittables, min(a.desc.n_hittables, c.desc.n_hittables) * C.sizeof(rt._abi.HittableDesc)))
    assert ha == hb and ha[:len(hc)] != hc


@pytest.mark.parametrize("idx,w,h", [(0, 600, 337), (1, 1200, 675), (4, 1200, 1200), (5, 600, 337), (6, 600, 600), (8, 800, 800)])
def test_reference_default_sizes(rt, earth, idx, w, h):   # the hard-coded CameraSettings literals of main.rs
    _, cs = rt.builtin_scene(idx, earth=earth if idx == 8 else None)
    cam = rt.Camera(cs)
    assert (cam.image_width, cam.image_height) == (w, h)
    assert cs.samples_per_pixel == {0: 128, 1: 128, 4: 128, 5: 1024, 6: 4096, 8: 8192}[idx]
    assert cs.max_depth == (40 if idx == 8 else 8)


def test_unknown_scene_index_is_random_balls(rt):   # main.rs:655
    a, _ = rt.builtin_scene(0)
    b, _ = rt.builtin_scene(42)
    assert a.desc.n_hittables == b.desc.n_hittables


@pytest.mark.parametrize("idx", range(9))
def test_camera_new_matches_oracle(rt, ob, earth, idx):
    _, cs = rt.builtin_scene(idx, earth=earth if idx in (2, 8) else None)
    a, b = rt.Camera(cs), ob.camera_new(cs)
    assert bytes(a) == bytes(b)


def test_perlin_tables(rt):   # perlin.rs:16-25,66-79
    s = rt.Scene()
    t = s.NoiseTexture(4.0, perlin_seed=3)
    s.finish(s.Sphere((0, 0, 0), 1.0, s.Lambertian(t)))
    p = s.desc.perlins[0]
    for perm in (p.perm_x, p.perm_y, p.perm_z):
        assert sorted(perm) == list(range(256))
    v = np.array([list(r) for r in p.ranvec])
    assert v.min() >= -1.0 and v.max() < 1.0 and abs(v.mean()) < 0.08
    assert not np.allclose(np.linalg.norm(v, axis=1), 1.0)   # NOT normalised (deviates from the book)


def test_earth_required(rt):
    with pytest.raises(rt._abi.RtError):
        s = rt._abi.SceneRequest(); s.scene = 2
        raw, d, cs = C.c_void_p(), rt._abi.SceneDesc(), rt.CameraSettings()
        rt._abi.check(rt._abi.lib().rt_scene_builtin(C.byref(s), C.byref(raw), C.byref(d), C.byref(cs)))


def test_bvh_imported_from_host_nodes_equals_the_built_one(rt, ob):
    """rt_hit_bvh_nodes: the drop-in's Rust host hands over the BVH it already built (bvh.rs:12-19). Re-importing the
    nodes of a tree built here must give the same description, the same device stream and the same hits."""
    import ctypes as C
    rng = np.random.default_rng(3)

    def objects(s):
        m = s.Lambertian(s.SolidColor(0.5, 0.5, 0.5))
        l = rt.HittableList()
        for _ in range(37):
            c = rng.uniform(-10, 10, 3)
            l.add(s.Sphere(tuple(c), float(rng.uniform(0.3, 1.5)), m))
        l.add(s.Quad((-3, -3, 4), (6, 0, 0), (0, 6, 0), m))
        return l
    state = rng.bit_generator.state
    a = rt.Scene(bvh_seed=9)
    a.finish(a.BVHNode(objects(a)))
    d = a.desc
    root = d.hittables[d.world].child
    nodes = []
    for k in range(d.hittables[d.world].count):
        nd = d.bvh_nodes[root + k]
        nodes.append((list(nd.bbox), nd.left - root if nd.left >= 0 else -1, nd.right - root if nd.right >= 0 else -1, nd.object))
    rng.bit_generator.state = state
    b = rt.Scene(bvh_seed=12345)           # the seed plays no role: no tree is built
    objects(b)
    b.finish(b.BVHFromNodes(nodes))
    wa, wb = rt.scene_ops(a)["words"], rt.scene_ops(b)["words"]
    assert wa.shape == wb.shape and np.array_equal(wa.view(np.uint32), wb.view(np.uint32))
    rays = np.zeros(4096, dtype=rt._abi.ray_dtype())
    rays["origin"] = rng.uniform(-12, 12, (4096, 3))
    rays["direction"] = rng.normal(size=(4096, 3))
    ha, hb = ob.hit_batch(a.desc, rays), ob.hit_batch(b.desc, rays)
    assert np.array_equal(ha["hit"], hb["hit"]) and np.array_equal(ha["prim_id"], hb["prim_id"]) and np.array_equal(ha["t"], hb["t"])
    # malformed trees are refused
    lib = rt._abi.lib()
    bad = (rt._abi.BvhNodeDesc * 3)()
    bad[0].left, bad[0].right, bad[0].object = 1, 1, -1
    bad[1].object = bad[2].object = 0
    assert lib.rt_hit_bvh_nodes(b._b, bad, 3) < 0
    bad[0].left, bad[0].right = 1, 2
    bad[2].object = 10 ** 6
    assert lib.rt_hit_bvh_nodes(b._b, bad, 3) == rt._abi.RT_ERR_OUT_OF_RANGE
    assert lib.rt_hit_bvh_nodes(b._b, bad, 2) < 0


def test_flattened_cube_lists_and_host_perlin_tables(rt):
    """What a host that flattens its own objects relies on: a HittableList of the six quads Quad::cube makes
    is recognised as a cube (one slab primitive on the device) without being told, a list that merely looks similar is not,
    and a NoiseTexture can carry the Perlin tables the host drew."""
    a = rt.Scene()
    m = a.Lambertian(a.SolidColor(0.5, 0.5, 0.5))
    a.finish(a.cube((1, 2, 3), (4, 6, 5), m))
    cube = a.desc.hittables[a.desc.world]
    quads = [a.desc.hittables[a.desc.list_items[cube.child + k]] for k in range(6)]
    b = rt.Scene()
    mb = b.Lambertian(b.SolidColor(0.5, 0.5, 0.5))
    l = rt.HittableList()
    for q in quads:
        l.add(b.Quad(tuple(q.v0), tuple(q.v1), tuple(q.v2), mb))
    b.finish(b.List(l))
    got = b.desc.hittables[b.desc.world]
    assert got.flags & rt._abi.RT_FLAG_CUBE_LIST and tuple(got.v0) == (1, 2, 3) and tuple(got.v1) == (4, 6, 5)
    assert rt.scene_layout(b)["n_box"] == 1 and rt.scene_layout(b)["n_quad"] == 0
    c = rt.Scene()
    mc = c.Lambertian(c.SolidColor(0.5, 0.5, 0.5))
    l = rt.HittableList()
    for k, q in enumerate(quads):
        u = tuple(q.v1) if k != 2 else (q.v1[0] * 0.5, q.v1[1], q.v1[2])        # one face shrunk: not a cube
        l.add(c.Quad(tuple(q.v0), u, tuple(q.v2), mc))
    c.finish(c.List(l))
    assert not (c.desc.hittables[c.desc.world].flags & rt._abi.RT_FLAG_CUBE_LIST)
    assert rt.scene_layout(c)["n_box"] == 0 and rt.scene_layout(c)["n_quad"] == 6
    # host-drawn Perlin tables end up in the description untouched
    s = rt.Scene()
    ref = s.NoiseTexture(4.0, perlin_seed=3)
    p = s.desc if s.desc else None
    s.finish(s.Sphere((0, 0, 0), 1.0, s.Lambertian(ref)))
    t0 = s.desc.perlins[0]
    rv = np.array([[t0.ranvec[k][c] for c in range(3)] for k in range(256)])
    px, py, pz = (np.array(t0.perm_x[:]), np.array(t0.perm_y[:]), np.array(t0.perm_z[:]))
    s2 = rt.Scene()
    s2.finish(s2.Sphere((0, 0, 0), 1.0, s2.Lambertian(s2.NoiseTextureFromTables(4.0, rv, px, py, pz))))
    t1 = s2.desc.perlins[0]
    assert bytes(t0) == bytes(t1)
    with pytest.raises(rt._abi.RtError):
        s2.NoiseTextureFromTables(4.0, rv, px + 300, py, pz)
