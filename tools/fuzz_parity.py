"""Randomised device-vs-oracle hit parity over generated scenes (development sweep, not part of the test suite):
random mixes of spheres (static / moving), quads, cubes, Translate / RotateY instances, BVHs nested in BVHs and media,
checked with the same acceptance rules as tests/test_gpu_hits.py.   gpurun -- python tools/fuzz_parity.py [--seeds 20]"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import rust_tracing_b200 as rt  # noqa: E402
from oracle import binding as ob  # noqa: E402
from test_gpu_hits import check_hits  # noqa: E402
from fuzz_scenes import random_scene, rich_scene  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seeds", type=int, default=20)
    ap.add_argument("--rays", type=int, default=1 << 16)
    ap.add_argument("--first", type=int, default=0)
    ap.add_argument("--rich", action="store_true", help="fuzz_scenes.rich_scene: textures, f64 spheres, media in instances, big streams")
    a = ap.parse_args()
    ctx = rt.Context(0)
    bad = 0
    for seed in range(a.first, a.first + a.seeds):
        try:
            s = rich_scene(5000 + seed) if a.rich else random_scene(1000 + seed)
            ds = ctx.upload(s)
        except rt._abi.RtError as e:
            if e.status != rt._abi.RT_ERR_UNSUPPORTED:
                raise
            print(f"seed {seed}: not expressible in the device layout ({str(e)[:90]})", flush=True)
            continue
        rng = np.random.default_rng(seed)
        rays = np.zeros(a.rays, dtype=rt._abi.ray_dtype())
        rays["origin"] = rng.uniform(-14, 14, (a.rays, 3))
        d = rng.normal(size=(a.rays, 3))
        rays["direction"] = d / np.linalg.norm(d, axis=1, keepdims=True) * rng.uniform(0.5, 2, (a.rays, 1))
        rays["time"] = rng.random(a.rays)
        ref = ob.hit_batch(s.desc, rays, seed=seed)
        dev = ctx.hit_batch(ds, rays, seed=seed)
        try:
            st = check_hits(dev, ref, rays, max_inequivalent=6)
            print(f"seed {seed}: ok  hits {int((ref['hit'] == 1).sum())}  layout {rt.scene_layout(s)}  {st}", flush=True)
        except AssertionError as e:
            bad += 1
            print(f"seed {seed}: FAIL {e}  layout {rt.scene_layout(s)}", flush=True)
            kinds = ["sphere", "quad", "list", "translate", "rotate_y", "medium", "bvh"]
            flips = np.nonzero(dev["hit"] != ref["hit"])[0][:6]
            for k in flips:
                who = ref[k] if ref[k]["hit"] else dev[k]
                h = s.desc.hittables[int(who["prim_id"])]
                print(f"   ray {k}: ref.hit={ref[k]['hit']} dev.hit={dev[k]['hit']} prim={int(who['prim_id'])} kind={kinds[h.kind]} flags={h.flags} "
                      f"t={who['t']:.6f} p={np.round(who['p'], 4)} uv=({who['u']:.5f},{who['v']:.5f}) o={np.round(rays[k]['origin'], 3)} d={np.round(rays[k]['direction'], 3)}", flush=True)
        ds.close()
    print("failures:", bad)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
