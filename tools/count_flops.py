"""Algorithmic flops per path for every BASELINE config.

flops_per_path = sum_ops count_op x cost_op, with the op COUNTS produced by the CPU oracle (the
reference's own visit pattern: per-axis AABB quirk, left-first traversal, origin-inclusive list
boxes) and the fixed per-op COSTS of SURVEY.md §8(d). Writes profiles/flops_per_path.json, which
bench.py reads for roofline.achieved. Runs on the CPU only (oracle); ~10 minutes on 8 cores.

    python tools/count_flops.py [--scale 0.5] [--spp 64]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import rust_tracing_b200 as rt  # noqa: E402
from oracle import binding as ob  # noqa: E402

# SURVEY.md §8(d) cost table (add/sub/mul/cmp/min/max/abs/floor/cvt = 1, FMA = 2, div/rcp/sqrt = 1,
# transcendental = 1; RNG integer work, addressing and loads = 0).
COST = {
    "ray_setup": 3, "node_test": 24, "sphere_reject": 24, "sphere_accept": 58, "moving_extra": 6,
    "quad_parallel": 7, "quad_t_reject": 16, "quad_ab_reject": 57, "quad_accept": 66,
    "translate_miss": 3, "translate_hit": 6, "rotate_in": 12, "rotate_out": 12, "medium": 32,
    "get_ray": 31, "get_ray_defocus": 18, "lambertian": 42, "metal": 57, "dielectric": 60, "isotropic": 33,
    "tex_solid": 0, "tex_checker": 9, "tex_image": 12, "tex_noise": 850, "bounce": 9,
}

# (scene index, width, depth override or 0 = reference value)
CONFIGS = {
    "cfg1_random_balls": (0, 400, 50), "cfg2a_checker": (1, 800, 0), "cfg2b_earth": (2, 800, 0),
    "cfg2c_perlin": (3, 800, 0), "cfg3_cornell_box": (6, 600, 50), "cfg4_cornell_smoke": (7, 600, 0),
    "cfg5_final_scene": (8, 800, 0),
}


def flops_from_counters(c):
    p = c["paths"]
    f = 0.0
    f += COST["ray_setup"] * (c["segments"] + c["rotate_in"])
    f += COST["node_test"] * c["node_tests"]
    f += COST["sphere_reject"] * (c["sphere_tests"] - c["sphere_accepts"]) + COST["sphere_accept"] * c["sphere_accepts"]
    f += COST["moving_extra"] * c["moving_sphere_tests"]
    f += (COST["quad_parallel"] * c["quad_parallel"] + COST["quad_t_reject"] * c["quad_t_reject"]
          + COST["quad_ab_reject"] * c["quad_ab_reject"] + COST["quad_accept"] * c["quad_accepts"])
    f += COST["translate_miss"] * (c["translate_in"] - c["translate_hit"]) + COST["translate_hit"] * c["translate_hit"]
    f += COST["rotate_in"] * c["rotate_in"] + COST["rotate_out"] * c["rotate_hit"]
    f += COST["medium"] * c["medium_tests"]
    f += COST["get_ray"] * c["get_ray"] + COST["get_ray_defocus"] * c["get_ray_defocus"]
    for k in ("lambertian", "metal", "dielectric", "isotropic", "tex_solid", "tex_checker", "tex_image", "tex_noise"):
        f += COST[k] * c[k]
    f += COST["bounce"] * c["segments"]
    return f / p


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=0.5, help="resolution scale vs the BASELINE size")
    ap.add_argument("--spp", type=int, default=64)
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    earth, src = rt.load_earth()
    out_path = os.path.join(ROOT, "profiles", "flops_per_path.json")
    result = {"cost_table": COST, "earth": src, "scale": args.scale, "spp": args.spp, "configs": {}}
    if os.path.exists(out_path):
        with open(out_path) as f:
            result["configs"] = json.load(f).get("configs", {})
    for name, (idx, width, depth) in CONFIGS.items():
        if args.only and args.only not in name:
            continue
        s, cs = rt.builtin_scene(idx, image_width=int(width * args.scale), max_depth=depth, earth=earth)
        cam = rt.Camera(cs)
        t0 = time.time()
        _, cnt = ob.render(s.desc, cam, 0, args.spp, seed=0, mode=0)
        dt = time.time() - t0
        result["configs"][name] = {
            "scene": rt.SCENE_NAMES[idx], "width": int(cam.image_width), "height": int(cam.image_height),
            "max_depth": int(cam.max_depth), "spp": args.spp, "paths": cnt["paths"],
            "flops_per_path": flops_from_counters(cnt),
            "counters_per_path": {k: v / cnt["paths"] for k, v in cnt.items()},
            "oracle_seconds": dt, "oracle_threads": os.cpu_count(), "oracle_mpaths_per_s": cnt["paths"] / dt / 1e6,
        }
        print(name, result["configs"][name]["flops_per_path"], f"{dt:.1f}s", flush=True)
        with open(out_path, "w") as f:
            json.dump(result, f, indent=1)
        s.close()


if __name__ == "__main__":
    main()
