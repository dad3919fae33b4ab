"""Short program that launches the helper kernels a few times each, for an ncu capture (python tools/ncu_small_kernels.py):
  expand_image_kernel  RGB8 texels -> linear float4, once per scene upload          (HBM bound)
  finalize_kernel      color_to_rgb(sum / spp) -> RGB8                               (HBM bound, latency-sized)
  jpeg_idct_kernel, jpeg_colour_kernel   rt_jpeg_decode on the earth image   (HBM bound)
  bvh_rank_sort_kernel, bvh_pair_kernel, bvh_box_kernel   rt_hit_bvh_device on 1000 spheres (latency-sized)
tools/make_helper_summary.py turns the capture into profiles/r2_helper_kernels_ncu.md."""
import io
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import rust_tracing_b200 as rt  # noqa: E402

earth, src = rt.load_earth()
ctx = rt.Context(0)
s = rt.Scene()
t = s.ImageTexture(earth)
s.finish(s.Sphere((0, 0, 0), 1.0, s.Lambertian(t)))
for _ in range(3):
    t0 = time.time()
    ds = ctx.upload(s)          # H2D of 61.4 MB RGB8 + expand_image_kernel (reads 61.4 MB, writes 327.7 MB)
    torch.cuda.synchronize()
    print(f"upload {1e3 * (time.time() - t0):.1f} ms")
    ds.close()
fb = torch.rand((800 * 800, 4), dtype=torch.float32, device="cuda") * 10000.0
for _ in range(3):
    out = ctx.finalize_rgb8(fb.data_ptr(), 800 * 800, 10000.0)   # finalize_kernel: reads 10.24 MB, writes 1.92 MB
print(out[:2])

p = os.path.join(ROOT, "assets", "earth-large.jpg")
if os.path.exists(p):
    data = open(p, "rb").read()
else:
    from PIL import Image
    buf = io.BytesIO()
    Image.fromarray(earth).save(buf, format="JPEG", quality=90, subsampling=2)
    data = buf.getvalue()
for _ in range(3):
    t0 = time.time()
    img = ctx.jpeg_decode(data)
    print(f"rt_jpeg_decode {img.shape} {1e3 * (time.time() - t0):.1f} ms (host Huffman + H2D + kernels + D2H)")
t0 = time.time()
info, coef = rt.jpeg_entropy_decode(data)
print(f"rt_jpeg_entropy_decode alone {1e3 * (time.time() - t0):.1f} ms")

rng = np.random.default_rng(1)
for _ in range(3):
    b = rt.Scene()
    m = b.Lambertian(b.SolidColor((0.5, 0.5, 0.5)))
    l = rt.HittableList()
    for c in rng.uniform(0, 165, (1000, 3)):
        l.add(b.Sphere(tuple(c), 10.0, m))
    t0 = time.time()
    b.BVHNodeOnDevice(ctx, l)
    print(f"rt_hit_bvh_device 1000 spheres {1e3 * (time.time() - t0):.2f} ms")
