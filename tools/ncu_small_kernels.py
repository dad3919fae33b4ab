"""Short program that launches the two HBM-bound helper kernels (expand_image_kernel at upload, finalize_kernel) a few
times, for an ncu capture: python tools/ncu_small_kernels.py"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import rust_tracing_b200 as rt  # noqa: E402

earth = rt.synthetic_earth()
ctx = rt.Context(0)
s = rt.Scene()
t = s.ImageTexture(earth)
s.finish(s.Sphere((0, 0, 0), 1.0, s.Lambertian(t)))
for _ in range(4):
    t0 = time.time()
    ds = ctx.upload(s)          # H2D of 61.4 MB RGB8 + expand_image_kernel (reads 61.4 MB, writes 327.7 MB)
    torch.cuda.synchronize()
    print(f"upload {1e3 * (time.time() - t0):.1f} ms")
    ds.close()
fb = torch.rand((800 * 800, 4), dtype=torch.float32, device="cuda") * 10000.0
for _ in range(4):
    out = ctx.finalize_rgb8(fb.data_ptr(), 800 * 800, 10000.0)   # finalize_kernel: reads 10.24 MB, writes 1.92 MB
print(out[:2])
