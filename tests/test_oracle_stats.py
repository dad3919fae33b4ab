"""Statistical pins of the oracle (SURVEY.md §8(c) "statistical pins"): they replace the seeds the reference
does not have. Also checks that the loop-free keyed samplers the device shares with the oracle (mode 0) have
the distributions of the reference's rejection loops (mode 1, vec3.rs:54-88)."""
import numpy as np
import pytest

from conftest import small_scene

N = 400_000


@pytest.mark.parametrize("mode", [0, 1])
def test_unit_vector_distribution(ob, mode):   # vec3.rs:63-65
    v = ob.sample(0, mode, 11, N)
    assert np.allclose(np.linalg.norm(v, axis=1), 1.0, atol=1e-12)
    assert np.abs(v.mean(axis=0)).max() < 4 / np.sqrt(N)
    assert np.allclose((v ** 2).mean(axis=0), 1 / 3, atol=4 * 0.3 / np.sqrt(N))
    # Lambertian: direction = n + unit vector; mean cosine of the normalised direction is 2/3
    d = v + np.array([0.0, 0.0, 1.0])
    cos = d[:, 2] / np.linalg.norm(d, axis=1)
    assert cos.mean() == pytest.approx(2 / 3, abs=4 * 0.24 / np.sqrt(N))


@pytest.mark.parametrize("mode", [0, 1])
def test_in_unit_sphere_distribution(ob, mode):   # vec3.rs:54-61
    v = ob.sample(1, mode, 12, N)
    r = np.linalg.norm(v, axis=1)
    assert r.max() < 1.0
    assert r.mean() == pytest.approx(0.75, abs=4 * 0.2 / np.sqrt(N))          # E|r| = 3/4
    assert (r < 0.5).mean() == pytest.approx(0.125, abs=4 * 0.33 / np.sqrt(N))   # volume ~ r^3
    assert np.abs(v.mean(axis=0)).max() < 4 * 0.45 / np.sqrt(N)


@pytest.mark.parametrize("mode", [0, 1])
def test_in_unit_disk_distribution(ob, mode):   # vec3.rs:77-88
    v = ob.sample(2, mode, 13, N)
    r = np.linalg.norm(v[:, :2], axis=1)
    assert r.max() < 1.0 and np.all(v[:, 2] == 0.0)
    assert r.mean() == pytest.approx(2 / 3, abs=4 * 0.24 / np.sqrt(N))
    assert (r < 0.5).mean() == pytest.approx(0.25, abs=4 * 0.44 / np.sqrt(N))


def test_medium_transmittance(rt, ob):   # constant_medium.rs:40-50: P(scatter inside a slab of length d) = 1 - exp(-rho d)
    s = rt.Scene()
    rho, thick = 0.05, 20.0
    box = s.cube((-100, -100, -thick), (100, 100, 0.0), s.Lambertian(s.SolidColor(1, 1, 1)))
    med = s.ConstantMedium(box, rho, (1.0, 1.0, 1.0))
    s.finish(med)
    n = 200_000
    rays = np.zeros(n, dtype=rt._abi.ray_dtype())
    rng = np.random.default_rng(3)
    rays["origin"] = np.c_[rng.uniform(-50, 50, n), rng.uniform(-50, 50, n), np.full(n, 10.0)]
    rays["direction"] = (0.0, 0.0, -2.5)          # un-normalised on purpose: distances scale by |d| (constant_medium.rs:46-51)
    h = ob.hit_batch(s.desc, rays, seed=5)
    p = (h["hit"] == 1).mean()
    want = 1 - np.exp(-rho * thick)
    assert p == pytest.approx(want, abs=4 * np.sqrt(want * (1 - want) / n))
    inside = h[h["hit"] == 1]
    assert np.all(inside["front_face"] == 0) and np.all(inside["prim_id"] == med)     # constant_medium.rs:52-58
    assert np.allclose(inside["normal"], (0.0, 0.0, 2.5))                             # normal = -direction, not normalised
    depth = -inside["p"][:, 2]
    assert depth.min() >= 0 and depth.max() <= thick
    assert depth.mean() == pytest.approx(1 / rho - thick * np.exp(-rho * thick) / (1 - np.exp(-rho * thick)), rel=0.02)


def test_furnace(rt, ob):
    """Camera inside a closed sphere that emits E and ... the crate's materials either emit or scatter, so
    use the series instead: a Lambertian sphere (albedo a) enclosing the camera, seen against nothing, gives 0;
    an emissive enclosure gives exactly E. A Lambertian enclosure around an emissive small sphere gives a value
    between E*a^k terms; here we pin the two closed forms."""
    s = rt.Scene()
    s.finish(s.Sphere((0, 0, 0), 5.0, s.DiffuseLight(s.SolidColor(2.0, 3.0, 4.0))))
    cam = rt.Camera(rt.CameraSettings(image_width=6, aspect_ratio=1.0, samples_per_pixel=4, max_depth=7))
    img, _ = ob.render(s.desc, cam, 0, 4)
    assert np.array_equal(img, np.broadcast_to(np.array([8.0, 12.0, 16.0]), img.shape))


def test_metal_and_dielectric_energy(rt, ob):
    # a glass sphere (attenuation 1, material.rs:82) on a constant background returns exactly the background
    s = rt.Scene()
    s.finish(s.Sphere((0, 0, -3), 1.0, s.Dielectric(1.5)))
    cam = rt.Camera(rt.CameraSettings(image_width=8, aspect_ratio=1.0, samples_per_pixel=8, max_depth=50, vfov=30.0,
                                      background=(0.25, 0.5, 1.0)))
    img, cnt = ob.render(s.desc, cam, 0, 8)
    assert cnt["depth_exhausted"] == 0
    assert np.allclose(img, np.array([2.0, 4.0, 8.0]), rtol=1e-15)
    assert cnt["dielectric"] > 0


@pytest.mark.parametrize("idx", [0, 6, 7])
def test_keyed_and_faithful_modes_agree(rt, ob, idx):
    """mode 0 (keyed RNG + loop-free samplers, what the device runs) and mode 1 (sequential RNG + the
    reference's rejection loops in the reference's call order) estimate the same image."""
    s, cam = small_scene(rt, idx, width=64)
    spp = 256
    a, _, sq_a = ob.render(s.desc, cam, 0, spp, seed=1, mode=0, want_sumsq=True)
    b, _, sq_b = ob.render(s.desc, cam, 0, spp, seed=2, mode=1, want_sumsq=True)
    lum = lambda im: (im * np.array([0.2126, 0.7152, 0.0722])).sum(axis=2) / spp
    la, lb = lum(a), lum(b)
    var = (sq_a / spp - la ** 2).clip(min=0) + (sq_b / spp - lb ** 2).clip(min=0)
    assert la.mean() == pytest.approx(lb.mean(), rel=0.02)
    rmse = np.sqrt(((la - lb) ** 2).mean())
    bound = 1.5 * np.sqrt(var.mean() / spp)
    assert rmse <= bound, (rmse, bound)


def test_screenshot_means(rt, ob):
    """Gross sanity against the reference's own screenshots (mean sRGB of screenshots/*.png,
    measured there). Unseeded renders at unknown spp, so this is a loose bound on deterministic scenes only."""
    want = {1: (0.556, 0.628, 0.571), 6: (0.294, 0.261, 0.230), 7: (0.522, 0.468, 0.421)}
    for idx, mean in want.items():
        s, cam = small_scene(rt, idx, width=96)
        spp = 128
        img, _ = ob.render(s.desc, cam, 0, spp, seed=0)
        got = ob.finalize_rgb8(img, spp).reshape(-1, 3).mean(axis=0) / 255.0
        assert np.abs(got - np.array(mean)).max() < 0.05, (idx, got, mean)
