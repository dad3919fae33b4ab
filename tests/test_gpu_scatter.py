"""GPU parity, part 3: Material::emitted / Material::scatter (material.rs:26-138) record by record.

rt_scatter_batch runs the very shade() the render kernel calls on (incoming ray, hit record) pairs; the f64 oracle's
scatter() driven by the same keyed RNG must make the same decisions (scattered / absorbed, reflect / refract) and produce
the same ray, attenuation and emission up to f32 rounding. Covers what whole images cannot resolve: Metal with fuzz 1.0
(main.rs:561) absorbed below the surface, Dielectric total internal reflection and both outcomes of the Schlick draw
(material.rs:93-94: the draw happens only when there is no TIR), Isotropic, two-sided DiffuseLight, textured albedos.

Stated tolerance: decisions identical except where the deciding quantity lies within 1e-5 of its threshold (counted, at
most 0.1 % of a batch); directions 2e-5 of |direction| + 2e-6; attenuation / emission 1e-5 relative.
Not covered here: Lambertian's near_zero branch (vec3.rs:113-116, EPS 1e-8) - it needs normal + unit_vector to cancel
to 1e-8, which f32 cannot represent at |normal| = 1; the branch exists on the device and is unreachable in both.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def unit(v):
    return v / np.linalg.norm(v, axis=-1, keepdims=True)


def make_scene(rt, earth_small):
    s = rt.Scene()
    mats = {
        "lambert_solid": s.Lambertian(s.SolidColor(0.8, 0.3, 0.1)),
        "lambert_checker": s.Lambertian(s.CheckerTexture(0.32, (0.2, 0.3, 0.1), (0.9, 0.9, 0.9))),
        "lambert_noise": s.Lambertian(s.NoiseTexture(4.0)),
        "lambert_image": s.Lambertian(s.ImageTexture(earth_small)),
        "metal_mirror": s.Metal((0.8, 0.8, 0.9), 0.0),
        "metal_fuzzy": s.Metal((0.7, 0.6, 0.5), 0.3),
        "metal_fuzz_1": s.Metal((0.8, 0.8, 0.9), 1.0),       # main.rs:561
        "glass": s.Dielectric(1.5),
        "bubble": s.Dielectric(1.0 / 1.5),
        "light": s.DiffuseLight(s.SolidColor(7.0, 7.0, 7.0)),
        "light_checker": s.DiffuseLight(s.CheckerTexture(0.5, (4, 0, 0), (0, 4, 0))),
        "smoke": s.Isotropic(s.SolidColor(0.2, 0.4, 0.9)),
        "smoke_noise": s.Isotropic(s.NoiseTexture(0.1)),
    }
    l = rt.HittableList()
    for m in mats.values():                       # every material must be referenced by a primitive to be uploaded
        l.add(s.Sphere((0, 0, 0), 1.0, m))
    s.finish(s.List(l))
    return s, mats


def make_records(rt, rng, n, mat_ids, grazing=False, inside=None):
    A = rt._abi
    rays = np.zeros(n, dtype=A.ray_dtype())
    hits = np.zeros(n, dtype=A.hit_dtype())
    normal = unit(rng.normal(size=(n, 3)))
    # incoming direction in the hemisphere against the normal (a record's normal always faces the ray, hittable.rs:23-27)
    d = unit(rng.normal(size=(n, 3)))
    cos = (d * normal).sum(1)
    d = d - 2.0 * np.maximum(cos, 0.0)[:, None] * normal
    if grazing:                                   # cos(incidence) uniform in [0, 1): TIR (sin_theta large), Schlick from 0.04 to 1
        tang = unit(d - (d * normal).sum(1)[:, None] * normal)
        c = rng.uniform(0.0, 1.0, n)[:, None]
        d = unit(tang * np.sqrt(1.0 - c * c) - normal * c)
    length = rng.uniform(0.5, 2.0, n)[:, None]    # the reference never normalises ray directions
    p = rng.uniform(-50.0, 50.0, size=(n, 3))
    t = rng.uniform(0.1, 20.0, n)
    rays["origin"] = p - d * length * t[:, None]
    rays["direction"] = d * length
    rays["time"] = rng.uniform(0.0, 1.0, n)
    hits["p"], hits["normal"], hits["t"] = p, normal, t
    hits["u"], hits["v"] = rng.uniform(0, 1, n), rng.uniform(0, 1, n)
    hits["hit"] = 1
    hits["front_face"] = rng.integers(0, 2, n) if inside is None else (0 if inside else 1)
    hits["mat_id"] = rng.choice(mat_ids, n)
    hits["prim_id"] = 0
    return rays, hits


def compare(dev, ref, what):
    n = len(ref)
    flips = dev["scattered"] != ref["scattered"]
    assert flips.sum() <= max(1, n // 1000), f"{what}: {int(flips.sum())} scatter / absorb flips of {n}"
    both = (dev["scattered"] == 1) & (ref["scattered"] == 1)
    dd, dr = dev["ray_out"]["direction"][both], ref["ray_out"]["direction"][both]
    err = np.linalg.norm(dd - dr, axis=1) / (np.linalg.norm(dr, axis=1) * 2e-5 + 2e-6)
    # a reflect / refract decision that sits on its threshold (Schlick draw == reflectance, sin_theta * ratio == 1) may flip
    branch_flips = err > 1.0
    assert branch_flips.sum() <= max(1, n // 1000), f"{what}: {int(branch_flips.sum())} direction mismatches of {int(both.sum())}"
    ok = both.copy()
    ok[np.flatnonzero(both)[branch_flips]] = False
    assert np.allclose(dev["ray_out"]["origin"][ok], ref["ray_out"]["origin"][ok], rtol=1e-6, atol=1e-5)
    assert np.allclose(dev["ray_out"]["time"][ok], ref["ray_out"]["time"][ok], atol=1e-7)
    assert np.allclose(dev["attenuation"][ok], ref["attenuation"][ok], rtol=1e-5, atol=2e-3 if "noise" in what else 1e-6)
    assert np.allclose(dev["emitted"], ref["emitted"], rtol=1e-5, atol=1e-6)
    return int(both.sum()), int((ref["scattered"] == 0).sum())


@pytest.fixture(scope="module")
def scatter_scene(rt, ctx):
    s, mats = make_scene(rt, rt.synthetic_earth(256, 128, seed=11))
    ds = ctx.upload(s)
    yield s, mats, ds
    ds.close()


@pytest.mark.parametrize("group", ["lambert_solid,lambert_checker,lambert_image", "lambert_noise,smoke_noise", "metal_mirror,metal_fuzzy",
                                   "light,light_checker", "smoke"])
def test_scatter_parity_by_material(rt, ob, ctx, scatter_scene, group):
    s, mats, ds = scatter_scene
    rng = np.random.default_rng(sum(map(ord, group)))
    n = 20000 if "noise" not in group else 4000
    rays, hits = make_records(rt, rng, n, [int(mats[k]) for k in group.split(",")])
    pix, smp = rng.integers(0, 1 << 20, n).astype(np.uint32), rng.integers(0, 1 << 16, n).astype(np.uint32)
    for seg in (0, 3):
        dev = ctx.scatter_batch(ds, rays, hits, pix, smp, segment=seg, seed=5)
        ref = ob.scatter_batch(s.desc, rays, hits, pix, smp, segment=seg, seed=5)
        scattered, absorbed = compare(dev, ref, group)
        if group.startswith("light"):
            assert scattered == 0 and (dev["emitted"].max(axis=1) > 0).all()      # two-sided lights (material.rs:114-122)
        elif not group.startswith("metal"):
            assert absorbed == 0


def test_metal_fuzz_one_is_absorbed_below_the_surface(rt, ob, ctx, scatter_scene):
    """fuzz = 1.0 (main.rs:561): reflect + 1.0 * random_in_unit_sphere points below the surface for a good share of the
    draws, and those rays are absorbed (material.rs:59-62). Both outcomes must agree record by record."""
    s, mats, ds = scatter_scene
    rng = np.random.default_rng(77)
    n = 30000
    rays, hits = make_records(rt, rng, n, [int(mats["metal_fuzz_1"])], grazing=True, inside=False)
    pix, smp = rng.integers(0, 1 << 20, n).astype(np.uint32), rng.integers(0, 1 << 16, n).astype(np.uint32)
    dev = ctx.scatter_batch(ds, rays, hits, pix, smp, segment=1, seed=9)
    ref = ob.scatter_batch(s.desc, rays, hits, pix, smp, segment=1, seed=9)
    scattered, absorbed = compare(dev, ref, "metal_fuzz_1")
    assert absorbed > n // 20 and scattered > n // 2
    assert np.allclose(dev["attenuation"][dev["scattered"] == 1], (0.8, 0.8, 0.9), atol=1e-6)


@pytest.mark.parametrize("mat,inside", [("glass", False), ("glass", True), ("bubble", False), ("bubble", True)])
def test_dielectric_tir_and_schlick(rt, ob, ctx, scatter_scene, mat, inside):
    """ratio = front_face ? 1/ir : ir; ratio * sin_theta > 1 reflects without drawing, otherwise reflectance > U decides
    (material.rs:81-103). Leaving glass (ratio 1.5) at grazing angles gives TIR; the same batch holds refractions and both
    outcomes of the draw."""
    s, mats, ds = scatter_scene
    rng = np.random.default_rng(31 + int(inside))
    n = 30000
    rays, hits = make_records(rt, rng, n, [int(mats[mat])], grazing=True, inside=inside)
    pix, smp = rng.integers(0, 1 << 20, n).astype(np.uint32), rng.integers(0, 1 << 16, n).astype(np.uint32)
    dev = ctx.scatter_batch(ds, rays, hits, pix, smp, segment=2, seed=1)
    ref = ob.scatter_batch(s.desc, rays, hits, pix, smp, segment=2, seed=1)
    scattered, absorbed = compare(dev, ref, mat)
    assert absorbed == 0 and scattered >= n - n // 1000
    assert np.allclose(dev["attenuation"], 1.0)
    # classify the oracle's outcomes: reflected rays stay on the incoming side of the surface
    nrm = hits["normal"]
    out_side = (ref["ray_out"]["direction"] * nrm).sum(1) > 0
    ior = 1.5 if mat == "glass" else 1.0 / 1.5
    ratio = ior if inside else 1.0 / ior
    d_hat = unit(rays["direction"])
    cos_t = np.minimum(-(d_hat * nrm).sum(1), 1.0)
    tir = ratio * np.sqrt(1.0 - cos_t ** 2) > 1.0
    assert out_side[tir].all()                                        # TIR always reflects
    if ratio > 1.0:
        assert tir.sum() > n // 20
    no_tir = ~tir
    assert out_side[no_tir].sum() > 50 and (~out_side[no_tir]).sum() > n // 8     # both outcomes of the Schlick draw occur


def test_scatter_batch_argument_checks(rt, ctx, scatter_scene):
    s, mats, ds = scatter_scene
    A = rt._abi
    rays = np.zeros(1, dtype=A.ray_dtype()); hits = np.zeros(1, dtype=A.hit_dtype())
    rays["direction"] = (0, 0, -1); hits["normal"] = (0, 0, 1)
    hits["mat_id"] = 999
    with pytest.raises(A.RtError) as e:
        ctx.scatter_batch(ds, rays, hits, [0], [0])
    assert e.value.status == A.RT_ERR_OUT_OF_RANGE
    assert len(ctx.scatter_batch(ds, rays[:0], hits[:0], [], [])) == 0
