"""Development probe (not a test): device vs oracle on every CLI scene — hit parity on ray batches,
low-spp image agreement with the shared keyed RNG, and a first throughput number. Writes
gpurun_out/probe.json and PNGs.

    gpurun -- python tools/gpu_probe.py [--scenes 0,6,8] [--spp 64]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import rust_tracing_b200 as rt  # noqa: E402
from oracle import binding as ob  # noqa: E402


def make_rays(cam, desc, n, seed=7):
    """Half camera rays (oracle get_ray), half secondary-like rays inside the scene bounds."""
    rng = np.random.default_rng(seed)
    A = rt._abi
    h, w = cam.shape
    pix = rng.integers(0, h * w, n // 2)
    smp = rng.integers(0, 64, n // 2)
    cam_rays = ob.get_ray_batch(cam, pix, smp, seed=seed)
    world = desc.hittables[desc.world]
    box = np.array(world.bbox[:]).reshape(3, 2)
    lo, hi = box[:, 0], box[:, 1]
    span = hi - lo
    finite = np.isfinite(span) & (span < 1e4)
    lo = np.where(finite, lo, -600.0)
    hi = np.where(finite, hi, 600.0)
    m = n - n // 2
    sec = np.zeros(m, dtype=A.ray_dtype())
    sec["origin"] = lo + (hi - lo) * rng.random((m, 3))
    d = rng.normal(size=(m, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    sec["direction"] = d * rng.uniform(0.5, 2.0, (m, 1))
    sec["time"] = rng.random(m)
    return np.concatenate([cam_rays, sec])


def compare_hits(dev, ref, rays):
    both = (dev["hit"] == 1) & (ref["hit"] == 1)
    res = {"n": int(len(ref)), "ref_hits": int((ref["hit"] == 1).sum()), "hit_mismatch": int((dev["hit"] != ref["hit"]).sum())}
    if both.any():
        same = both & (dev["prim_id"] == ref["prim_id"])
        res["prim_mismatch"] = int((both & (dev["prim_id"] != ref["prim_id"])).sum())
        dlen = np.linalg.norm(rays["direction"], axis=1)
        scale = np.maximum(1.0, np.abs(ref["p"]).max(axis=1) + np.abs(rays["origin"]).max(axis=1))
        dt = np.abs(dev["t"] - ref["t"]) * dlen / scale / 2.0 ** -23
        res["t_ulp_scene_max"] = float(dt[same].max()) if same.any() else 0.0
        res["t_ulp_scene_p99"] = float(np.percentile(dt[same], 99)) if same.any() else 0.0
        res["p_abs_max"] = float(np.abs(dev["p"] - ref["p"])[same].max()) if same.any() else 0.0
        res["n_abs_max"] = float(np.abs(dev["normal"] - ref["normal"])[same].max()) if same.any() else 0.0
        res["uv_abs_max"] = float(max(np.abs(dev["u"] - ref["u"])[same].max(), np.abs(dev["v"] - ref["v"])[same].max())) if same.any() else 0.0
        res["front_mismatch"] = int((same & (dev["front_face"] != ref["front_face"])).sum())
        res["mat_mismatch"] = int((same & (dev["mat_id"] != ref["mat_id"])).sum())
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scenes", default="0,1,2,3,4,5,6,7,8")
    ap.add_argument("--spp", type=int, default=64)
    ap.add_argument("--rays", type=int, default=1 << 16)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out"))
    args = ap.parse_args()
    os.makedirs(args.out, exist_ok=True)
    from PIL import Image
    ctx = rt.Context(0)
    info = ctx.device_info()
    report = {"device": info, "fp32_peak_tflops": ctx.measure_fp32_peak(), "scenes": {}}
    print(report, flush=True)
    earth, earth_src = rt.load_earth()
    for idx in [int(x) for x in args.scenes.split(",")]:
        name = rt.SCENE_NAMES[idx]
        small_w = {0: 200, 1: 200, 2: 200, 3: 200, 4: 160, 5: 200, 6: 150, 7: 150, 8: 160}[idx]
        s, cs = rt.builtin_scene(idx, image_width=small_w, earth=earth)
        cam = rt.Camera(cs)
        ds = ctx.upload(s)
        rays = make_rays(cam, s.desc, args.rays)
        ref = ob.hit_batch(s.desc, rays)
        dev = ctx.hit_batch(ds, rays)
        entry = {"hits": compare_hits(dev, ref, rays)}
        # low-spp agreement with the shared keyed RNG
        spp = 4
        o_img, cnt = ob.render(s.desc, cam, 0, spp, seed=0, mode=0)
        d_img = ctx.render(ds, cam, 0, spp, seed=0)
        assert np.allclose(d_img[..., 3], spp), "sample count channel"
        rel = np.abs(d_img[..., :3] - o_img) / (np.abs(o_img) + 1e-3 * spp)
        entry["agree_1e-3"] = float((rel.max(axis=2) < 1e-3).mean())
        entry["mean_ratio"] = float(d_img[..., :3].mean() / max(o_img.mean(), 1e-12))
        Image.fromarray(rt.color_to_rgb8(d_img, spp)).save(os.path.join(args.out, f"dev_{name}.png"))
        Image.fromarray(ob.finalize_rgb8(o_img, spp).reshape(cam.shape + (3,))).save(os.path.join(args.out, f"ora_{name}.png"))
        # statistical agreement at moderate spp (different RNG seeds on purpose)
        o_img2, _, sq = ob.render(s.desc, cam, 0, args.spp, seed=1, mode=0, want_sumsq=True)
        d_img2 = ctx.render(ds, cam, 0, args.spp * 4, seed=2)
        entry["mean_lum_ratio_stat"] = float((d_img2[..., :3].mean() / (args.spp * 4)) / max(o_img2.mean() / args.spp, 1e-12))
        # throughput at the BASELINE size
        s2, cs2 = rt.builtin_scene(idx, image_width={0: 400, 1: 800, 2: 800, 3: 800, 4: 800, 5: 800, 6: 600, 7: 600, 8: 800}[idx],
                                   max_depth={0: 50, 6: 50}.get(idx, 0), earth=earth)
        cam2 = rt.Camera(cs2)
        ds2 = ctx.upload(s2)
        ctx.render(ds2, cam2, 0, 8, seed=0)
        t0 = time.time()
        spp_t = 64
        ctx.render(ds2, cam2, 0, spp_t, seed=0)
        dt = time.time() - t0
        st = ctx.stats()
        entry["mpaths_per_s"] = cam2.shape[0] * cam2.shape[1] * spp_t / dt / 1e6
        entry["segments_per_path"] = st["segments"] / max(1, st["paths"])
        entry["oracle_counters_per_path"] = {k: v / cnt["paths"] for k, v in cnt.items()}
        report["scenes"][name] = entry
        print(name, json.dumps(entry), flush=True)
        ds.close(); ds2.close(); s.close(); s2.close()
    report["earth"] = earth_src
    with open(os.path.join(args.out, "probe.json"), "w") as f:
        json.dump(report, f, indent=1)


if __name__ == "__main__":
    main()
