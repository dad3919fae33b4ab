"""Pins the oracle (the CPU restatement) against the known-answer values SURVEY.md §8(c) derives analytically
from the reference formulas. The reference ships no tests or golden vectors of its own."""
import ctypes as C
import math

import numpy as np
import pytest


def one_ray(rt, o, d, time=0.0):
    r = np.zeros(1, dtype=rt._abi.ray_dtype())
    r["origin"], r["direction"], r["time"] = o, d, time
    return r


def sphere_scene(rt, center=(0, 0, -1), radius=0.5):
    s = rt.Scene()
    m = s.Lambertian(s.SolidColor(0.5, 0.5, 0.5))
    return s.finish(s.Sphere(center, radius, m))


def test_sphere_front_hit(rt, ob):   # sphere.rs:59-89
    s = sphere_scene(rt)
    h = ob.hit_batch(s.desc, one_ray(rt, (0, 0, 0), (0, 0, -1)))[0]
    assert h["hit"] == 1 and h["t"] == 0.5 and h["front_face"] == 1
    assert np.allclose(h["p"], (0, 0, -0.5), atol=0) and np.allclose(h["normal"], (0, 0, 1), atol=0)
    assert h["u"] == pytest.approx(0.25, abs=1e-15) and h["v"] == pytest.approx(0.5, abs=1e-15)


def test_sphere_inside_hit(rt, ob):
    s = sphere_scene(rt)
    h = ob.hit_batch(s.desc, one_ray(rt, (0, 0, -1), (0, 0, -1)))[0]
    assert h["hit"] == 1 and h["t"] == 0.5 and h["front_face"] == 0
    assert np.allclose(h["normal"], (0, 0, 1), atol=0)      # outward (0,0,-1) flipped against the ray
    assert h["u"] == pytest.approx(0.75, abs=1e-15) and h["v"] == pytest.approx(0.5, abs=1e-15)


# n = (-1, 0, z): the u seam. atan2(-n.z, n.x) is +pi for n.z = -0.0 (u = 1.0, SURVEY's value) and -pi for
# n.z = +0.0 (u = 0.0) because Rust's unary minus, like C's, turns +0.0 into -0.0. Both are pinned.
@pytest.mark.parametrize("n,uv", [((1, 0, 0), (0.5, 0.5)), ((0, 1, 0), (0.5, 1.0)), ((-1, 0, -0.0), (1.0, 0.5)),
                                  ((-1, 0, 0.0), (0.0, 0.5)), ((0, 0, 1), (0.25, 0.5)), ((0, 0, -1), (0.75, 0.5))])
def test_sphere_uv_table(ob, n, uv):   # sphere.rs:48-52
    u, v = C.c_double(), C.c_double()
    ob.lib().oracle_sphere_uv((C.c_double * 3)(*n), C.byref(u), C.byref(v))
    assert (u.value, v.value) == pytest.approx(uv, abs=1e-15)


def test_sphere_interval_is_open(rt, ob):   # ray_t.surrounds (sphere.rs:78, interval.rs:43-45)
    s = sphere_scene(rt)
    r = one_ray(rt, (0, 0, 0), (0, 0, -1))
    assert ob.hit_batch(s.desc, r, t_min=0.001, t_max=0.5)[0]["hit"] == 0     # near root == max: rejected, far root 1.5 > max
    assert ob.hit_batch(s.desc, r, t_min=0.5, t_max=10.0)[0]["t"] == 1.5      # near root == min: rejected, far root taken


def test_quad_hit(rt, ob):   # quad.rs:97-133, main.rs:254-259
    s = rt.Scene()
    m = s.Lambertian(s.SolidColor(0.2, 1.0, 0.2))
    s.finish(s.Quad((-2, -2, 0), (4, 0, 0), (0, 4, 0), m))
    h = ob.hit_batch(s.desc, one_ray(rt, (0, 0, 9), (0, 0, -1)))[0]
    assert h["hit"] == 1 and h["t"] == 9.0 and h["front_face"] == 1
    assert h["u"] == 0.5 and h["v"] == 0.5 and tuple(h["normal"]) == (0, 0, 1)
    # closed interval (quad.rs:115): t == max is accepted
    assert ob.hit_batch(s.desc, one_ray(rt, (0, 0, 9), (0, 0, -1)), t_max=9.0)[0]["hit"] == 1
    # parallel ray (quad.rs:110-112) and outside alpha/beta
    assert ob.hit_batch(s.desc, one_ray(rt, (0, 0, 9), (1, 0, 0)))[0]["hit"] == 0
    assert ob.hit_batch(s.desc, one_ray(rt, (2.5, 0, 9), (0, 0, -1)))[0]["hit"] == 0
    # no back-face culling (quad.rs:104-109): hit from behind, normal flipped against the ray
    hb = ob.hit_batch(s.desc, one_ray(rt, (0, 0, -9), (0, 0, 1)))[0]
    assert hb["hit"] == 1 and hb["front_face"] == 0 and tuple(hb["normal"]) == (0, 0, -1)


def test_aabb_per_axis_quirk(rt, ob):   # aabb.rs:64-84: ray_t is never narrowed between axes
    r = one_ray(rt, (-1, -3, 0.5), (1, 1, 0))
    box = (C.c_double * 6)(0, 1, 0, 1, 0, 1)
    assert ob.lib().oracle_aabb_hit(box, r.ctypes.data, 0.001, float("inf")) == 1   # a book slab test says miss
    r2 = one_ray(rt, (-1, 5, 0.5), (1, 1, 0))
    assert ob.lib().oracle_aabb_hit(box, r2.ctypes.data, 0.001, float("inf")) == 0


def test_checker(rt, ob):   # texture.rs:59-70
    s = rt.Scene()
    t = s.CheckerTexture(0.32, (1, 0, 0), (0, 0, 1))
    s.finish(s.Sphere((0, 0, 0), 1.0, s.Lambertian(t)))
    out = ob.texture_batch(s.desc, t, [[0, 0, 0.1, 0.1, 0.1], [0, 0, -0.1, 0.1, 0.1]])
    assert tuple(out[0]) == (1, 0, 0) and tuple(out[1]) == (0, 0, 1)


def test_gamma(ob):   # color.rs:12-27
    out = (C.c_double * 3)()
    ob.lib().oracle_rgb_to_color(255, 128, 0, out)
    assert out[0] == 1.0 and out[1] == pytest.approx(0.2195197180748679, abs=1e-15) and out[2] == 0.0
    px = (C.c_uint8 * 3)()
    for val, want in [(1.0, 255), (0.5, 186), (0.0, 0), (0.2, 123), (4.0, 255), (float("nan"), 0), (-1.0, 0)]:
        ob.lib().oracle_color_to_rgb((C.c_double * 3)(val, val, val), px)
        assert px[0] == want, (val, px[0])


def test_schlick(ob):   # material.rs:74-78
    f = ob.lib().oracle_reflectance
    assert f(1.0, 1.5) == pytest.approx(0.04, abs=1e-15)
    assert f(0.0, 1.5) == pytest.approx(1.0, abs=1e-15)
    assert f(0.5, 1 / 1.5) == pytest.approx(0.07, abs=1e-15)


def test_refract_reflect(ob):   # vec3.rs:91-101
    out = (C.c_double * 3)()
    k = math.sqrt(0.5)
    ob.lib().oracle_refract((C.c_double * 3)(k, -k, 0), (C.c_double * 3)(0, 1, 0), 1 / 1.5, out)
    assert tuple(out) == pytest.approx((0.4714045207910317, -0.8819171036881969, 0.0), abs=1e-15)
    ob.lib().oracle_reflect((C.c_double * 3)(1, -1, 0), (C.c_double * 3)(0, 1, 0), out)
    assert tuple(out) == (1, 1, 0)


def test_camera_final_scene(rt, ob):   # camera.rs:54-110 with main.rs:624-636
    _, cs = rt.builtin_scene("final_scene", image_width=800, earth=np.zeros((2, 2, 3), np.uint8))
    cam = ob.camera_new(cs)
    assert (cam.image_width, cam.image_height) == (800, 800)
    assert tuple(cam.pixel00_loc) == pytest.approx((478.28633100616247, 281.6351527147337, -589.3636307973846), abs=1e-10)
    assert tuple(cam.pixel_delta_u) == pytest.approx((-0.00863231205589697, 0, -0.00287743735196566), abs=1e-15)
    assert tuple(cam.pixel_delta_v) == pytest.approx((0, -0.00909925585665506, 0), abs=1e-15)


def test_camera_random_balls(rt, ob):   # main.rs:121-135 at W=400
    _, cs = rt.builtin_scene("random_balls", image_width=400)
    cam = ob.camera_new(cs)
    assert cam.image_height == 225
    assert tuple(cam.pixel00_loc) == pytest.approx((2.4070753285126214, 2.253536773358229, 3.7645238500996845), abs=1e-12)
    assert tuple(cam.defocus_disk_u) == pytest.approx((0.01177372383354866, 0, -0.05101946994537751), abs=1e-15)


@pytest.mark.parametrize("w,h", [(400, 225), (600, 337), (800, 450), (1200, 675)])
def test_image_height_truncation(rt, ob, w, h):   # camera.rs:69
    cam = ob.camera_new(rt.CameraSettings(image_width=w))
    assert cam.image_height == h


def test_rotate_y_bbox(rt, ob):   # hittable.rs:120-157 with main.rs:392-397
    s = rt.Scene()
    m = s.Lambertian(s.SolidColor(0.73, 0.73, 0.73))
    # a slanted quad whose (unpadded) box is exactly (0,0,0)-(165,330,165)
    r = s.RotateY(s.Quad((0, 0, 0), (165, 0, 0), (0, 330, 165), m), 15.0)
    # the real Cornell box: the cube's list box is the union of its quads' PADDED boxes (quad.rs:41-43,
    # aabb.rs:35-53), so it is 5e-5 larger on every side than the cube itself
    c = s.RotateY(s.cube((0, 0, 0), (165, 330, 165), m), 15.0)
    l = rt.HittableList(); l.add(r); l.add(c)
    s.finish(s.List(l))
    bb = s.desc.hittables[r].bbox
    assert tuple(bb) == pytest.approx((0.0, 202.0829037796122, 0.0, 330.0, -42.705142441915925, 159.37776133769628), abs=1e-12)
    cb = s.desc.hittables[c].bbox
    assert cb[2] == -5e-5 and cb[3] == 330.00005 and cb[1] > bb[1]
    assert ob.validate_scene(s.desc)[0] == 0.0


def test_depth_exhaustion_returns_black(rt, ob):   # renderer.rs:140-142
    # camera inside a closed Lambertian sphere with a non-black background: no path ever escapes, so after
    # max_depth bounces the tail contributes 0 (not the background).
    s = rt.Scene()
    s.finish(s.Sphere((0, 0, 0), 5.0, s.Lambertian(s.SolidColor(1, 1, 1))))
    cam = rt.Camera(rt.CameraSettings(image_width=8, aspect_ratio=1.0, samples_per_pixel=4, max_depth=5,
                                      background=(0.7, 0.8, 1.0)))
    img, cnt = ob.render(s.desc, cam, 0, 4)
    assert img.max() == 0.0 and cnt["depth_exhausted"] == cnt["paths"] and cnt["segments"] == 5 * cnt["paths"]


def test_light_seen_directly_and_two_sided(rt, ob):   # material.rs:114-122: emits from both faces, no scatter
    s = rt.Scene()
    light = s.DiffuseLight(s.SolidColor(4, 3, 2))
    s.finish(s.Quad((-50, -50, -5), (100, 0, 0), (0, 100, 0), light))
    for z_from in (0.0, -10.0):   # front and back of the quad
        cam = rt.Camera(rt.CameraSettings(image_width=4, aspect_ratio=1.0, samples_per_pixel=3, max_depth=5,
                                          look_from=(0, 0, z_from), look_at=(0, 0, -5)))
        img, _ = ob.render(s.desc, cam, 0, 3)
        assert np.array_equal(img, np.broadcast_to(np.array([12.0, 9.0, 6.0]), img.shape))


def test_convex_lambertian_on_constant_background(rt, ob):
    # a convex Lambertian sphere of albedo a under constant background B returns exactly a*B where it is seen
    s = rt.Scene()
    s.finish(s.Sphere((0, 0, -3), 1.0, s.Lambertian(s.SolidColor(0.5, 0.25, 1.0))))
    cam = rt.Camera(rt.CameraSettings(image_width=9, aspect_ratio=1.0, samples_per_pixel=2, max_depth=10, vfov=10.0,
                                      background=(0.8, 0.4, 0.2)))
    img, _ = ob.render(s.desc, cam, 0, 2)
    assert img[4, 4] == pytest.approx((2 * 0.4, 2 * 0.1, 2 * 0.2), rel=1e-15)
