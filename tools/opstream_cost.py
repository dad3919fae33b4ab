"""Ops a realistic ray population visits in the flattened traversal stream, counted on the host (tests/opstream.py).

Rays: the segments of real paths - camera rays, then whatever the f64 oracle scatters at each hit - of a scene, so the
mix of primary / secondary / in-medium rays is the render's own. Used to evaluate stream-layout changes (which cull
boxes to keep) before spending GPU time:     python tools/opstream_cost.py [--scene 8] [--paths 20000]
"""
import argparse
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import rust_tracing_b200 as rt  # noqa: E402
from oracle import binding as ob  # noqa: E402
import opstream  # noqa: E402


def path_segments(s, cam, n_paths, max_depth, seed=0):
    """All segments (rays) of n_paths keyed paths, traced by the oracle."""
    A = rt._abi
    rng = np.random.default_rng(seed)
    h, w = cam.shape
    pix = rng.integers(0, h * w, n_paths)
    smp = rng.integers(0, 1 << 20, n_paths)
    rays = ob.get_ray_batch(cam, pix, smp, seed=seed)
    lib = ob.lib()
    lib.oracle_scatter.restype = C.c_int
    out = []
    for seg in range(max_depth):
        if len(rays) == 0:
            break
        out.append(rays.copy())
        # media draws of the oracle's hit_batch are keyed by ray index, not by (pixel, sample, seg): fine for a ray population
        hits = ob.hit_batch(s.desc, rays, seed=seed + 17 * seg)
        nxt = np.zeros(len(rays), dtype=A.ray_dtype())
        keep = np.zeros(len(rays), dtype=bool)
        sc = A.RayDesc()
        for k in np.flatnonzero(hits["hit"] == 1):
            r = A.RayDesc.from_buffer_copy(rays[k].tobytes())
            hd = A.HitDesc.from_buffer_copy(hits[k].tobytes())
            ok = lib.oracle_scatter(C.byref(s.desc), C.byref(r), C.byref(hd), C.c_uint64(seed), C.c_uint32(int(pix[k])),
                                    C.c_uint32(int(smp[k])), C.c_uint32(seg), 0, C.byref(sc), None, None)
            if ok == 1:
                nxt[k] = np.frombuffer(bytes(sc), dtype=A.ray_dtype())[0]
                keep[k] = True
        rays, pix, smp = nxt[keep], pix[keep], smp[keep]
    return np.concatenate(out), len(out[0])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scene", type=int, default=8)
    ap.add_argument("--width", type=int, default=800)
    ap.add_argument("--paths", type=int, default=20000)
    ap.add_argument("--depth", type=int, default=40)
    a = ap.parse_args()
    earth, _ = rt.load_earth()
    s, cs = rt.builtin_scene(a.scene, image_width=a.width, max_depth=a.depth, earth=earth)
    cam = rt.Camera(cs)
    rays, n_paths = path_segments(s, cam, a.paths, a.depth)
    print(f"scene {a.scene}: {n_paths} paths, {len(rays)} segments ({len(rays) / n_paths:.2f} per path)")
    S = opstream.Stream(rt.scene_ops(s))
    counts = {}
    opstream.hit_batch(S, rays, counts=counts)
    print("ops per segment:", {k: round(v / len(rays), 2) for k, v in sorted(counts.items())})
    print("layout:", rt.scene_layout(s))


if __name__ == "__main__":
    main()
