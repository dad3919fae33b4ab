"""Fixtures taken from the reference's own outputs (run in the build container, where /root/reference exists).

* tests/golden/reference/cornell_box.png, cornell_smoke.png: byte copies of /root/reference/screenshots/ - the only
  deterministic, current-code outputs the reference ships (main.rs:344-506: 600x600, 4096 spp, depth 8). They pin the
  oracle (tests/test_reference_screenshots.py). MIT-licensed data, see tests/golden/reference/NOTICE.
* tests/golden/reference/earth_blocks.npz: 15x15 block means (linear space) of screenshots/earth.png. That screenshot
  predates the committed scene (its background is a sky gradient, main.rs:163 has a constant), so it cannot pin
  radiance; its albedo PATTERN still pins the sphere u,v orientation and the image row flip (sphere.rs:48-52,
  texture.rs:83-92).
* assets/earth-large.jpg: a byte copy of the reference's texture (main.rs:179,591) so the GPU box, which has no
  /root/reference, renders the reference's own texels. Data, git-ignored (assets/), shipped by gpurun.
"""
import os
import shutil
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("RT_REFERENCE", "/root/reference")


def linear(rgb8):
    """Inverse of color_to_rgb (color.rs:12-19) at the centre of the byte's bin."""
    return ((np.asarray(rgb8, dtype=np.float64) + 0.5) / 256.0) ** 2.2


def block_means(x, b):
    h, w = x.shape[:2]
    return x[: h // b * b, : w // b * b].reshape(h // b, b, w // b, b, -1).mean(axis=(1, 3))


def copy_earth_asset():
    src = os.path.join(REF, "assets", "earth-large.jpg")
    dst = os.path.join(ROOT, "assets", "earth-large.jpg")
    if os.path.exists(src) and not os.path.exists(dst):
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
    return os.path.exists(dst)


def main():
    from PIL import Image
    out = os.path.join(ROOT, "tests", "golden", "reference")
    os.makedirs(out, exist_ok=True)
    for name in ("cornell_box", "cornell_smoke"):
        shutil.copyfile(os.path.join(REF, "screenshots", f"{name}.png"), os.path.join(out, f"{name}.png"))
    raw = np.asarray(Image.open(os.path.join(REF, "screenshots", "earth.png")).convert("RGB"))
    np.savez_compressed(os.path.join(out, "earth_blocks.npz"), block=15, shape=np.array(raw.shape[:2]),
                        linear_block_means=block_means(linear(raw), 15).astype(np.float32))
    print("earth asset present:", copy_earth_asset())


if __name__ == "__main__":
    sys.exit(main())
