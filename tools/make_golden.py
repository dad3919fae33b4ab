"""Generates tests/golden/*.npz from the CPU oracle (the reference itself cannot run here: no Rust toolchain,
and it has no golden vectors of its own). The fixtures freeze the oracle's answers on fixed seeds so that
(a) any later change to the oracle is caught on the CPU, and (b) the GPU tests can compare the device path
against committed numbers. Scenes that need the earth image use the synthetic stand-in so the fixtures do not
depend on the reference's assets.

    python tools/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

import rust_tracing_b200 as rt  # noqa: E402
from oracle import binding as ob  # noqa: E402
from gpu_probe import make_rays  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
SMALL_W = {0: 48, 1: 48, 2: 48, 3: 48, 4: 40, 5: 48, 6: 40, 7: 40, 8: 40}


def main():
    os.makedirs(OUT, exist_ok=True)
    earth = rt.synthetic_earth(256, 128, seed=11)   # small stand-in: fixtures stay tiny
    for idx, name in enumerate(rt.SCENE_NAMES):
        s, cs = rt.builtin_scene(idx, image_width=SMALL_W[idx], earth=earth)
        cam = rt.Camera(cs)
        rays = make_rays(cam, s.desc, 2048, seed=7)
        hits = ob.hit_batch(s.desc, rays, seed=7)
        img, cnt = ob.render(s.desc, cam, 0, 4, seed=0, mode=0)
        np.savez_compressed(os.path.join(OUT, f"{name}.npz"), rays=rays, hits=hits, image_sum_4spp=img,
                            counters=np.array([cnt[k] for k in ob.COUNTER_NAMES], dtype=np.uint64),
                            width=SMALL_W[idx])
        print(name, "hits", int(hits["hit"].sum()), "mean", img.mean())
    # textures: checker / image / noise on fixed points
    rng = np.random.default_rng(21)
    s = rt.Scene()
    t_chk = s.CheckerTexture(0.32, (0.2, 0.3, 0.1), (0.9, 0.9, 0.9))
    t_img = s.ImageTexture(earth)
    t_noise = s.NoiseTexture(4.0, perlin_seed=3)
    s.finish(s.Sphere((0, 0, 0), 1.0, s.Lambertian(t_noise)))
    uvp = np.c_[rng.random((4096, 2)), rng.uniform(-30, 30, (4096, 3))]
    uvp[:8, :2] = [[0, 0], [1, 1], [0, 1], [1, 0], [-0.5, 2.0], [0.5, 0.5], [0.999999, 0.000001], [0.25, 0.75]]
    np.savez_compressed(os.path.join(OUT, "textures.npz"), uvp=uvp,
                        checker=ob.texture_batch(s.desc, t_chk, uvp), image=ob.texture_batch(s.desc, t_img, uvp),
                        noise=ob.texture_batch(s.desc, t_noise, uvp))
    # camera rays
    for idx in (0, 8):
        _, cs = rt.builtin_scene(idx, image_width=SMALL_W[idx], earth=earth)
        cam = rt.Camera(cs)
        pix = rng.integers(0, cam.shape[0] * cam.shape[1], 1024)
        smp = rng.integers(0, 10000, 1024)
        np.savez_compressed(os.path.join(OUT, f"camera_{rt.SCENE_NAMES[idx]}.npz"), pixel=pix, sample=smp,
                            rays=ob.get_ray_batch(cam, pix, smp, seed=0), width=SMALL_W[idx])


if __name__ == "__main__":
    main()
