"""Short single-GPU program for ncu: uploads a scene and renders it a few times.

    python tools/ncu_target.py [--scene 8] [--width 800] [--spp 64] [--reps 3] [--depth 0]
"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rust_tracing_b200 as rt  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scene", type=int, default=8)
    ap.add_argument("--width", type=int, default=800)
    ap.add_argument("--spp", type=int, default=64)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--depth", type=int, default=0)
    a = ap.parse_args()
    earth, _ = rt.load_earth()
    s, cs = rt.builtin_scene(a.scene, image_width=a.width, max_depth=a.depth, earth=earth)
    cam = rt.Camera(cs)
    ctx = rt.Context(0)
    t0 = time.time()
    ds = ctx.upload(s)
    print(f"upload {1e3 * (time.time() - t0):.1f} ms", flush=True)
    t0 = time.time()
    ds.close()
    ds = ctx.upload(s)
    print(f"close+upload again {1e3 * (time.time() - t0):.1f} ms", flush=True)
    for r in range(a.reps):
        t0 = time.time()
        img = ctx.render(ds, cam, 0, a.spp, seed=r)
        dt = time.time() - t0
        st = ctx.stats()
        print(f"rep {r}: {cam.shape[0] * cam.shape[1] * a.spp / dt / 1e6:.1f} Mpaths/s  segs/path {st['segments'] / st['paths']:.2f}  mean {img[..., :3].mean() / a.spp:.4f}", flush=True)


if __name__ == "__main__":
    main()
