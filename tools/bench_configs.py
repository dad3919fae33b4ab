"""Every BASELINE.json config on one GPU, with the CPU oracle beside it and a converged-image
parity report. Writes gpurun_out/configs.json.

    gpurun -- python tools/bench_configs.py [--parity-spp 256]

Per config: GPU Mpaths/s at the config's full size and spp (CUDA events around rt_render_accumulate, scene resident,
3 warm-up launches), device-counted flops per path and roofline fraction, CPU oracle Mpaths/s on a bounded sample,
and image parity at 1/4 resolution: oracle at N_ref spp vs device at 16 x N_ref spp (independent seeds), per-channel
RMSE against the Monte-Carlo bound from the oracle's per-pixel variance, and the mean-luminance ratio.
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
from bench import DEVICE_COST, device_flops_per_path  # noqa: E402

CONFIGS = {  # name: (scene, width, spp, depth override)
    "cfg1_random_balls": (0, 400, 100, 50), "cfg2a_checker": (1, 800, 500, 0), "cfg2b_earth": (2, 800, 500, 0),
    "cfg2c_perlin": (3, 800, 500, 0), "cfg3_cornell_box": (6, 600, 1000, 50), "cfg4_cornell_smoke": (7, 600, 2000, 0),
    "cfg5_final_scene": (8, 800, 10000, 0),
}
LUM = np.array([0.2126, 0.7152, 0.0722])


def main():
    import torch
    ap = argparse.ArgumentParser()
    ap.add_argument("--parity-spp", type=int, default=256)
    ap.add_argument("--only", default="")
    ap.add_argument("--cfg5-spp", type=int, default=2000, help="spp used to TIME cfg5 here (bench.py times the full 10000)")
    args = ap.parse_args()
    earth, earth_src = rt.load_earth()
    ctx = rt.Context(0)
    peak = ctx.measure_fp32_peak()
    out = {"fp32_peak_tflops_measured": peak, "earth": earth_src, "cpu_threads": os.cpu_count(), "configs": {}}
    dev = torch.device("cuda:0")
    stream = torch.cuda.current_stream(dev)
    for name, (idx, width, spp, depth) in CONFIGS.items():
        if args.only and args.only not in name:
            continue
        s, cs = rt.builtin_scene(idx, image_width=width, samples_per_pixel=spp, max_depth=depth, earth=earth)
        cam = rt.Camera(cs)
        h, w = cam.shape
        ds = ctx.upload(s)
        fb = torch.zeros((h, w, 4), dtype=torch.float32, device=dev)
        t_spp = args.cfg5_spp if name.startswith("cfg5") else spp
        for _ in range(3):
            ctx.render_accumulate(ds, cam, 0, max(1, t_spp // 8), 0, fb.data_ptr(), stream.cuda_stream)
        torch.cuda.synchronize()
        times = []
        for rep in range(3):
            fb.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            ctx.render_accumulate(ds, cam, 0, t_spp, rep, fb.data_ptr(), stream.cuda_stream)
            e1.record(stream)
            e1.synchronize()
            times.append(e0.elapsed_time(e1) * 1e-3)
        t = min(times)
        paths = h * w * t_spp
        ops = ctx.count_ops(ds, cam, 0, 4, seed=0)
        fpp = device_flops_per_path(ops)
        entry = {"scene": rt.SCENE_NAMES[idx], "width": w, "height": h, "spp": spp, "timed_spp": t_spp, "max_depth": int(cam.max_depth),
                 "gpu_mpaths_per_s": paths / t / 1e6, "gpu_seconds": t, "flops_per_path_device": fpp,
                 "achieved_tflops": fpp * paths / t / 1e12, "roofline_frac": fpp * paths / t / 1e12 / peak,
                 "segments_per_path": ops["segments"] / ops["paths"]}
        # CPU oracle on a bounded sample of the same config
        t0 = time.perf_counter()
        _, cnt = ob.render(s.desc, cam, 0, 1, seed=0, mode=0)
        dt1 = time.perf_counter() - t0
        c_spp = max(1, min(32, int(6.0 / max(dt1, 1e-3))))
        t0 = time.perf_counter()
        _, cnt = ob.render(s.desc, cam, 0, c_spp, seed=0, mode=0)
        dt = time.perf_counter() - t0
        entry["cpu_mpaths_per_s"] = cnt["paths"] / dt / 1e6
        entry["cpu_sample"] = f"{c_spp} spp, {os.cpu_count()} threads, {dt:.1f} s"
        entry["speedup_vs_cpu"] = entry["gpu_mpaths_per_s"] / entry["cpu_mpaths_per_s"]
        ds.close()
        s.close()
        # converged-image parity at quarter resolution
        s2, cs2 = rt.builtin_scene(idx, image_width=max(32, width // 4), max_depth=depth, earth=earth)
        cam2 = rt.Camera(cs2)
        ds2 = ctx.upload(s2)
        n_ref = args.parity_spp
        n_gpu = 16 * n_ref
        ref, _, sq = ob.render(s2.desc, cam2, 0, n_ref, seed=101, mode=0, want_sumsq=True)
        devimg = ctx.render(ds2, cam2, 0, n_gpu, seed=202)
        ref_m, dev_m = ref / n_ref, devimg[..., :3] / n_gpu
        l_ref = (ref_m * LUM).sum(axis=2)
        var = (sq / n_ref - l_ref ** 2).clip(min=0) * n_ref / (n_ref - 1)
        bound = 1.5 * np.sqrt(var.mean() * (1.0 / n_ref + 1.0 / n_gpu))
        rmse_c = [float(np.sqrt(((ref_m[..., c] - dev_m[..., c]) ** 2).mean())) for c in range(3)]
        rmse_l = float(np.sqrt(((l_ref - (dev_m * LUM).sum(axis=2)) ** 2).mean()))
        entry["image_parity"] = {"size": list(cam2.shape), "n_ref": n_ref, "n_gpu": n_gpu, "rmse_rgb": rmse_c, "rmse_luminance": rmse_l,
                                 "bound_1p5_sigma_luminance": float(bound), "within_bound": bool(rmse_l <= bound),
                                 "mean_luminance_ratio": float((dev_m * LUM).sum(axis=2).mean() / l_ref.mean())}
        ds2.close()
        s2.close()
        out["configs"][name] = entry
        print(name, json.dumps(entry), flush=True)
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "configs.json"), "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
