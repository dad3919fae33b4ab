"""Wavefront renderer (render_v4.cuh) against the megakernel (render_v3.cuh): same paths, same image.

Both trace the same keyed paths with the same device functions, so their SUM framebuffers may differ only by the
order of the f32 additions; the sample-count channel must be identical. Also prints Mpaths/s of each.

    gpurun -- python tools/wf_check.py [--quick]
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rust_tracing_b200 as rt  # noqa: E402

# (scene, width, spp, depth)
CASES = [(8, 30, 1, 40), (6, 64, 2, 50), (8, 800, 16, 40), (8, 200, 8, 40), (6, 600, 16, 50), (7, 300, 8, 50), (0, 400, 32, 50), (3, 403, 8, 0), (4, 200, 8, 0),
         (1, 100, 4, 0), (2, 100, 4, 0), (5, 201, 8, 0)]


def context(variant, **env):
    os.environ["RT_B200_KERNEL"] = str(variant)
    for k, v in env.items():
        os.environ[k] = str(v)
    c = rt.Context(0)
    for k in env:
        os.environ.pop(k, None)
    os.environ.pop("RT_B200_KERNEL", None)
    return c


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    a = ap.parse_args()
    c3, c4 = context(3), context(4)
    c4s = context(4, RT_B200_POOL=4096)      # a tiny pool: many iterations, reserve / window edge cases
    earth, _ = rt.load_earth()
    bad = 0
    for scene, width, spp, depth in (CASES[:4] if a.quick else CASES):
        s, cs = rt.builtin_scene(scene, image_width=width, max_depth=depth, earth=earth)
        cam = rt.Camera(cs)
        n = cam.shape[0] * cam.shape[1] * spp
        out = {}
        for name, c in (("v3", c3), ("v4", c4), ("v4 pool 4096", c4s)):
            if name == "v4 pool 4096" and n > 3e6:
                continue
            ds = c.upload(s)
            c.render(ds, cam, 0, min(spp, 2), seed=1)          # warm-up
            t0 = time.time()
            img = c.render(ds, cam, 3, spp, seed=5)
            dt = time.time() - t0
            st = c.stats()
            out[name] = (img, n / dt / 1e6, st)
            ds.close()
        ref = out["v3"][0]
        line = f"scene {scene} {cam.shape[1]}x{cam.shape[0]} spp {spp}:"
        for name, (img, mps, st) in out.items():
            line += f"  {name} {mps:.0f} Mpaths/s"
            if name == "v3":
                continue
            same_count = bool(np.array_equal(img[..., 3], ref[..., 3]))
            scale = np.maximum(np.abs(ref[..., :3]), 1e-3 * max(1.0, float(np.abs(ref[..., :3]).max())))
            rel = float((np.abs(img[..., :3] - ref[..., :3]) / scale).max())
            segs_same = st["segments"] == out["v3"][2]["segments"] and st["paths"] == out["v3"][2]["paths"]
            ok = same_count and rel < 1e-4 and segs_same
            bad += 0 if ok else 1
            line += f" [counts {'ok' if same_count else 'DIFFER'}, max rel diff {rel:.2e}, segments {'ok' if segs_same else str(st) + ' vs ' + str(out['v3'][2])}]"
        print(line, flush=True)
    print("failures:", bad)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
