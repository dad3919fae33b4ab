"""The oracle reproduces the committed fixtures (tests/golden/, made by tools/make_golden.py) exactly:
any drift of the CPU restatement shows up here, on the CPU, before it can move the GPU parity target."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def small_earth(rt):
    return rt.synthetic_earth(256, 128, seed=11)


def load_scene(rt, name):
    g = np.load(os.path.join(GOLD, f"{name}.npz"))
    s, cs = rt.builtin_scene(name, image_width=int(g["width"]), earth=small_earth(rt))
    return g, s, rt.Camera(cs)


@pytest.mark.parametrize("idx", range(9))
def test_oracle_matches_golden(rt, ob, idx):
    name = rt.SCENE_NAMES[idx]
    g, s, cam = load_scene(rt, name)
    hits = ob.hit_batch(s.desc, g["rays"], seed=7)
    for f in ("hit", "front_face", "prim_id", "mat_id"):
        assert np.array_equal(hits[f], g["hits"][f]), f
    for f in ("t", "p", "normal", "u", "v"):
        assert np.allclose(hits[f], g["hits"][f], rtol=1e-13, atol=1e-13), f
    img, cnt = ob.render(s.desc, cam, 0, 4, seed=0, mode=0)
    assert np.allclose(img, g["image_sum_4spp"], rtol=1e-12, atol=1e-12)
    assert np.array_equal(np.array([cnt[k] for k in ob.COUNTER_NAMES], dtype=np.uint64), g["counters"])


def test_oracle_textures_match_golden(rt, ob):
    g = np.load(os.path.join(GOLD, "textures.npz"))
    s = rt.Scene()
    t_chk = s.CheckerTexture(0.32, (0.2, 0.3, 0.1), (0.9, 0.9, 0.9))
    t_img = s.ImageTexture(small_earth(rt))
    t_noise = s.NoiseTexture(4.0, perlin_seed=3)
    s.finish(s.Sphere((0, 0, 0), 1.0, s.Lambertian(t_noise)))
    assert np.array_equal(ob.texture_batch(s.desc, t_chk, g["uvp"]), g["checker"])
    assert np.allclose(ob.texture_batch(s.desc, t_img, g["uvp"]), g["image"], rtol=1e-14, atol=0)
    assert np.allclose(ob.texture_batch(s.desc, t_noise, g["uvp"]), g["noise"], rtol=1e-12, atol=1e-13)
    # clamp + (W-1) scaling + truncation of texture.rs:84-89: corners land on corner texels
    e = small_earth(rt).astype(np.float64) / 255.0
    assert np.allclose(g["image"][0], e[-1, 0] ** 2.2) and np.allclose(g["image"][1], e[0, -1] ** 2.2)
    assert np.allclose(g["image"][4], e[0, 0] ** 2.2)      # u=-0.5 -> 0, v=2.0 -> 1 -> row 0


@pytest.mark.parametrize("name", ["random_balls", "final_scene"])
def test_oracle_camera_matches_golden(rt, ob, name):
    g = np.load(os.path.join(GOLD, f"camera_{name}.npz"))
    _, cs = rt.builtin_scene(name, image_width=int(g["width"]), earth=small_earth(rt))
    cam = rt.Camera(cs)
    rays = ob.get_ray_batch(cam, g["pixel"], g["sample"], seed=0)
    for f in ("origin", "direction", "time"):
        assert np.allclose(rays[f], g["rays"][f], rtol=1e-14, atol=1e-15), f
    if name == "final_scene":
        assert np.all(rays["origin"] == np.array(cs.look_from[:]))     # no defocus: every ray leaves the centre
    else:
        assert np.linalg.norm(rays["origin"] - np.array(cs.look_from[:]), axis=1).max() > 0   # defocus disk
    assert rays["time"].min() >= 0 and rays["time"].max() < 1
