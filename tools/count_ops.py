"""Device op counts per path (instrumented kernel) for the BASELINE configs -> gpurun_out/device_ops.json"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rust_tracing_b200 as rt  # noqa: E402

CONFIGS = {"cfg1_random_balls": (0, 400, 50), "cfg2a_checker": (1, 800, 0), "cfg2b_earth": (2, 800, 0),
           "cfg2c_perlin": (3, 800, 0), "cfg3_cornell_box": (6, 600, 50), "cfg4_cornell_smoke": (7, 600, 0),
           "cfg5_final_scene": (8, 800, 0)}


def main():
    earth, _ = rt.load_earth()
    ctx = rt.Context(0)
    out = {}
    for name, (idx, width, depth) in CONFIGS.items():
        s, cs = rt.builtin_scene(idx, image_width=width, max_depth=depth, earth=earth)
        cam = rt.Camera(cs)
        ds = ctx.upload(s)
        c = ctx.count_ops(ds, cam, 0, 16, seed=0)
        per = {k: v / c["paths"] for k, v in c.items()}
        per["lane_utilisation"] = c["lane_ops"] / (32.0 * c["votes"])
        out[name] = per
        print(name, json.dumps({k: round(v, 3) for k, v in per.items()}), flush=True)
        ds.close()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "device_ops.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
