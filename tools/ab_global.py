"""Same-GPU comparison of the two homes of the op stream: shared memory (the product's choice when the stream fits, with the
instantiation specialised on the scene's features) against RT_LAYOUT_OPS_IN_GLOBAL (what a scene too large for shared
memory gets: the generic instantiation reading the stream through L1).   gpurun -- python tools/ab_global.py [--spp 200]"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rust_tracing_b200 as rt  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--spp", type=int, default=200)
    ap.add_argument("--rounds", type=int, default=3)
    ap.add_argument("--scenes", type=int, nargs="*", default=[0, 6, 7, 8])
    a = ap.parse_args()
    import torch
    ctx = rt.Context(0)
    earth, _ = rt.load_earth()
    for idx in a.scenes:
        s, cs = rt.builtin_scene(idx, earth=earth if idx in (2, 8) else None)
        cam = rt.Camera(cs)
        h, w = cam.shape
        fb = torch.zeros((h, w, 4), dtype=torch.float32, device="cuda")
        res = {}
        scenes = {"shared": ctx.upload(s), "global": ctx.upload(s, rt.layout_flags(ops_in_smem=False))}
        for name, ds in scenes.items():
            ctx.render_accumulate(ds, cam, 0, 8, 0, fb.data_ptr())
        torch.cuda.synchronize()
        for r in range(a.rounds):
            for name, ds in scenes.items():
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                ctx.render_accumulate(ds, cam, 0, a.spp, 0, fb.data_ptr())
                e1.record()
                torch.cuda.synchronize()
                res.setdefault(name, []).append(h * w * a.spp / e0.elapsed_time(e1) / 1e3)
        m = {k: float(np.median(v)) for k, v in res.items()}
        print(f"scene {idx} {w}x{h} spp {a.spp}: shared {m['shared']:8.1f} Mpaths/s   global {m['global']:8.1f} Mpaths/s   x{m['global'] / m['shared']:.3f}", flush=True)
        for ds in scenes.values():
            ds.close()


if __name__ == "__main__":
    main()
