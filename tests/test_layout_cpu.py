"""The flattening that rt_scene_upload performs (scene description -> threaded op stream), dry-run on the host:
op counts per scene, the cube -> slab-primitive rewrite, media hoisting, f64 flags, and rejected nestings."""
import numpy as np
import pytest

from conftest import small_scene


@pytest.fixture
def unpruned(rt):
    """Layout flags for the stream as flattened, before the cull boxes that do not pay for themselves are dropped
    (prune_stream)."""
    return rt.layout_flags(prune=False)


def test_final_scene_layout(rt, unpruned):
    s, _ = small_scene(rt, 8, rt.synthetic_earth(64, 32))
    L = rt.scene_layout(s, unpruned)
    assert L["n_box"] == 400                      # 400 Quad::cube lists -> 400 slab primitives (2400 quads in the description)
    assert L["n_quad"] == 1                       # the light
    assert L["n_sphere"] == 1000 + 6              # box of spheres + moving, glass, metal, r=70 boundary, earth, noise
    assert L["n_inner"] == 399 + 999 + 10         # branch nodes of the three BVHs (leaves carry no box of their own)
    assert L["n_xform"] == 1 and L["n_bvh"] == 3
    assert L["n_medium_hoisted"] == 2 and L["n_medium_in_stream"] == 0
    assert L["n_precise_spheres"] == 0            # the r=5000 fog boundary only feeds a free-flight comparison: f32
    assert L["device_bytes"] < 400_000            # everything but full-size texels fits L1/L2 (earth here is a 64x32 stand-in)


def test_other_scenes_layout(rt, unpruned):
    L0 = rt.scene_layout(small_scene(rt, 0)[0], unpruned)
    assert L0["n_precise_spheres"] == 1 and L0["n_box"] == 0 and L0["n_sphere"] == L0["n_inner"] + 1
    L6 = rt.scene_layout(small_scene(rt, 6)[0], unpruned)
    assert (L6["n_quad"], L6["n_box"], L6["n_xform"], L6["n_inner"]) == (6, 2, 2, 7)
    L7 = rt.scene_layout(small_scene(rt, 7)[0], unpruned)
    assert (L7["n_quad"], L7["n_box"], L7["n_xform"], L7["n_medium_hoisted"]) == (6, 0, 0, 2)   # boxes only bound media


@pytest.mark.parametrize("idx", range(9))
def test_box_pruning_drops_only_cull_boxes(rt, idx):
    """prune_stream removes OP_INNER nodes whose expected saving is below their cost (surface-area model); every
    primitive, instance and medium stays, in the same order; OP_INNER_REF nodes (semantics, not culling) stay."""
    s, _ = small_scene(rt, idx, rt.synthetic_earth(64, 32) if idx in (2, 8) else None)
    pruned = rt.scene_layout(s)
    full = rt.scene_layout(s, rt.layout_flags(prune=False))
    for k in ("n_sphere", "n_quad", "n_box", "n_xform", "n_medium_in_stream", "n_medium_hoisted", "n_bvh", "n_precise_spheres"):
        assert pruned[k] == full[k], k
    assert pruned["n_inner"] <= full["n_inner"]
    assert pruned["n_words"] == full["n_words"] - 2 * (full["n_inner"] - pruned["n_inner"])
    if idx in (0, 6, 7, 8):
        assert pruned["n_inner"] < full["n_inner"]


def test_degenerate_cube_keeps_its_quads(rt):
    s = rt.Scene()
    m = s.Lambertian(s.SolidColor(0.5, 0.5, 0.5))
    s.finish(s.cube((0, 0, 0), (1, 0, 1), m))     # flat in y: not a slab primitive
    L = rt.scene_layout(s)
    assert L["n_box"] == 0 and L["n_quad"] == 6


def test_quads_that_stick_out_of_their_box_get_reference_nodes(rt, unpruned):
    """Quad::new boxes the diagonal q .. q+u+v only (quad.rs:41-43). For such a quad every BVH node above it, its own
    leaf included, is emitted with the reference's box and per-axis test (OP_INNER_REF) instead of a tight box."""
    def layout(skew):
        s = rt.Scene(bvh_seed=5)
        m = s.Lambertian(s.SolidColor(0.5, 0.5, 0.5))
        l = rt.HittableList()
        l.add(s.Quad((0, 0, 0), (1, 0, 0.5 * skew), (0, 1, -0.5 * skew), m))
        l.add(s.Quad((3, 0, 0), (1, 0, 0), (0, 1, 0), m))
        l.add(s.Sphere((6, 0, 0), 0.5, m))
        s.finish(s.BVHNode(l))
        return rt.scene_layout(s, unpruned)
    assert layout(0)["n_inner"] == 2          # root + one branch; leaves need no box of their own
    assert layout(1)["n_inner"] == 3          # + the leaf box of the skewed quad


def test_media_inside_instances_stay_in_the_stream(rt):
    s = rt.Scene()
    m = s.Lambertian(s.SolidColor(0.5, 0.5, 0.5))
    med = s.ConstantMedium(s.Sphere((0, 0, 0), 1.0, m), 0.5, (1, 1, 1))
    s.finish(s.Translate(med, (3, 0, 0)))
    L = rt.scene_layout(s)
    assert L["n_medium_in_stream"] == 1 and L["n_medium_hoisted"] == 0 and L["n_xform"] == 1


def test_nested_instances_flatten_to_composed_transforms(rt):
    """An instance inside another instance's subtree: both become OP_XFORM_ENTER ops, the inner one holding the
    composition (world -> its local space), its OP_XFORM_EXIT naming the outer one as parent."""
    s = rt.Scene()
    m = s.Lambertian(s.SolidColor(0.5, 0.5, 0.5))
    inner = rt.HittableList()
    inner.add(s.Translate(s.Sphere((0, 0, 0), 1.0, m), (1, 0, 0)))
    inner.add(s.Sphere((3, 0, 0), 1.0, m))
    s.finish(s.RotateY(s.BVHNode(inner), 90.0))
    assert rt.scene_layout(s)["n_xform"] == 2
    ops = rt.scene_ops(s)
    W, I = ops["words"], ops["words"].view(np.int32).reshape(-1, 4)
    kinds = (I[:ops["n_world_words"], 3] >> 8) & 15
    # walk op by op to find the real op starts (the header's low byte is the op's size in bytes)
    starts, i = [], 0
    while i < ops["n_world_words"]:
        starts.append(i)
        i += (int(I[i, 3]) & 0xFF) >> 4
    enter = [i for i in starts if kinds[i] == 3]
    exits = [i for i in starts if kinds[i] == 4]
    assert len(enter) == 2 and len(exits) == 2
    outer, nested = enter
    assert I[exits[0], 0] == outer and I[exits[1], 0] == -1          # inner exit -> parent, outer exit -> world
    # composed transform of the nested instance maps the world position of its sphere's centre to the local origin:
    # RotateY(90 deg) takes local (1, 0, 0) [the translated centre] to world (0, 0, -1) (hittable.rs:173-179)
    a, sn, b, cs = W[nested + 2, :3], W[nested + 2, 3], W[nested + 3, :3], W[nested + 3, 3]
    x = np.array([0.0, 0.0, -1.0]) - a
    local = np.array([cs * x[0] - sn * x[2], x[1], sn * x[0] + cs * x[2]]) + b
    assert np.allclose(local, 0.0, atol=1e-6)


def test_unsupported_nesting_rejected_on_the_host(rt):
    s2 = rt.Scene()
    m2 = s2.Lambertian(s2.SolidColor(0.5, 0.5, 0.5))
    inner_med = s2.ConstantMedium(s2.Sphere((0, 0, 0), 1.0, m2), 0.5, (1, 1, 1))
    s2.finish(s2.ConstantMedium(inner_med, 0.5, (1, 1, 1)))       # a medium bounding a medium
    with pytest.raises(rt._abi.RtError) as e2:
        rt.scene_layout(s2)
    assert e2.value.status == rt._abi.RT_ERR_UNSUPPORTED


def test_instances_inside_the_boundary_of_an_instanced_medium(rt, ob):
    """A generic boundary (list / BVH) that holds instances, around a medium that itself sits inside an instance: the
    boundary program runs on the enclosing instance's local ray, so its own instances compose their transforms from the
    program's frame and their exits name parents inside the program only (-1 = the ray the program was handed)."""
    import opstream
    s3 = rt.Scene()
    m3 = s3.Lambertian(s3.SolidColor(0.5, 0.5, 0.5))
    two = rt.HittableList()
    two.add(s3.Translate(s3.RotateY(s3.Sphere((0, 0, 0), 1.0, m3), 30.0), (1, 0, 0)))
    two.add(s3.Sphere((0, 2, 0), 1.0, m3))
    nested = rt.HittableList()
    nested.add(s3.Translate(s3.List(two), (0, 0, 1)))              # an instance holding an instance, all inside the boundary
    med = s3.ConstantMedium(s3.List(nested), 0.5, (1, 1, 1))
    s3.finish(s3.Translate(s3.RotateY(med, 20.0), (0.5, 0, 0)))
    L = rt.scene_layout(s3)
    assert L["n_medium_in_stream"] == 1 and L["n_xform"] == 3
    S = opstream.Stream(rt.scene_ops(s3))
    I = S.i
    prog = [i for i, k, fl in S.walk() if k == opstream.OP_MEDIUM][0]
    b0, b1 = int(I[prog + 1, 0]), int(I[prog + 1, 1])
    exits = [(i, int(I[i, 0])) for i, k, _ in S.walk() if k == opstream.OP_XFORM_EXIT]    # walk() steps through the program too
    inside = [p for i, p in exits if b0 <= i < b1]
    outside = [p for i, p in exits if not b0 <= i < b1]
    assert outside == [-1]                                          # the world program sees only the outer instance
    # inside the program: the nested instance names its parent IN the program, that parent names the program's own frame
    assert len(inside) == 2 and b0 <= inside[0] < b1 and inside[1] == -1
    rng = np.random.default_rng(4)
    n = 1 << 13
    rays = np.zeros(n, dtype=rt._abi.ray_dtype())
    rays["origin"] = rng.uniform(-4, 4, (n, 3))
    d = rng.normal(size=(n, 3))
    rays["direction"] = d / np.linalg.norm(d, axis=1, keepdims=True) * rng.uniform(0.5, 2, (n, 1))
    rays["time"] = rng.random(n)
    ref = ob.hit_batch(s3.desc, rays, seed=2)
    emu = opstream.hit_batch(S, rays, seed=2)
    assert int(ref["hit"].sum()) > 80
    assert int((emu["hit"] != ref["hit"]).sum()) <= 1
    both = (emu["hit"] == 1) & (ref["hit"] == 1)
    assert np.abs(emu["t"] - ref["t"])[both].max() <= 2e-5 * max(1.0, np.abs(ref["t"][both]).max())


@pytest.mark.parametrize("idx", range(9))
def test_every_cli_scene_flattens(rt, idx):
    s, _ = small_scene(rt, idx, rt.synthetic_earth(64, 32))
    L = rt.scene_layout(s)
    assert L["n_words"] >= 2 and L["n_bvh"] >= 1
