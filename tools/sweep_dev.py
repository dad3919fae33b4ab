"""Same-GPU sweep of the A/B build's switches (csrc/librt_b200_dev.so, compiled with -DRT_B200_DEV: it reads RT_B200_*
at rt_context_create). One context per variant, renders alternate, CUDA-event timed.

    python tools/sweep_dev.py --scene 8 --spp 500 "RT_B200_SHADE_MIN=20" "RT_B200_SHADE_MIN=28,RT_B200_SLAB_FAST=12"
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("RT_B200_LIB", os.path.join(ROOT, "rust-tracing_b200", "csrc", "librt_b200_dev.so"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("variants", nargs="*")
    ap.add_argument("--scene", type=int, default=8)
    ap.add_argument("--spp", type=int, default=500)
    ap.add_argument("--rounds", type=int, default=3)
    a = ap.parse_args()
    import torch
    import rust_tracing_b200 as rt
    earth, _ = rt.load_earth()
    s, cs = rt.builtin_scene(a.scene, earth=earth)
    cam = rt.Camera(cs)
    runs = []
    for v in [""] + a.variants:
        env = dict(kv.split("=", 1) for kv in v.split(",") if kv)
        for k in [k for k in os.environ if k.startswith("RT_B200_") and k != "RT_B200_LIB"]:
            del os.environ[k]
        os.environ.update(env)
        ctx = rt.Context(0)
        runs.append((v or "(defaults)", ctx, ctx.upload(s)))
    h, w = cam.shape
    fb = torch.zeros((h, w, 4), dtype=torch.float32, device="cuda:0")
    stream = torch.cuda.current_stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    res = {n: [] for n, *_ in runs}
    for rnd in range(-1, a.rounds):
        for name, ctx, ds in (runs if rnd % 2 == 0 else runs[::-1]):
            fb.zero_()
            spp = a.spp if rnd >= 0 else max(1, a.spp // 8)
            e0.record(stream)
            ctx.render_accumulate(ds, cam, 0, spp, rnd + 7, fb.data_ptr(), stream.cuda_stream)
            e1.record(stream)
            e1.synchronize()
            if rnd >= 0:
                res[name].append(h * w * spp / e0.elapsed_time(e1) / 1e3)
    base = np.median(res[runs[0][0]])
    for name, *_ in runs:
        print(f"scene {a.scene}  {name:55s} {np.median(res[name]):8.1f} Mpaths/s  x{np.median(res[name]) / base:.3f}", flush=True)


if __name__ == "__main__":
    main()
