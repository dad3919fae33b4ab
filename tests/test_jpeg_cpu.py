"""SURVEY.md 8(f) rank 4, host half and the oracle of the device half: ImageTexture::new's decode (texture.rs:76-80).

rt_jpeg_entropy_decode (markers + Huffman, host C++) feeds oracle/jpeg_oracle.py (numpy restatement of libjpeg's integer
IDCT, fancy upsampling and colour conversion); the result must equal PIL's (libjpeg-turbo) decode BYTE FOR BYTE - on the
reference's own asset and on generated files of every supported layout, odd sizes, optimised tables, restart markers."""
import io
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import jpeg_oracle  # noqa: E402


def pil_decode(data):
    from PIL import Image
    Image.MAX_IMAGE_PIXELS = None
    return np.asarray(Image.open(io.BytesIO(data)).convert("RGB"), dtype=np.uint8)


def encode(img, **kw):
    from PIL import Image
    buf = io.BytesIO()
    Image.fromarray(img).save(buf, format="JPEG", **kw)
    return buf.getvalue()


def jpeg_cases(rt):
    rng = np.random.default_rng(1)
    smooth = rt.synthetic_earth(640, 320, seed=5)
    noise = rng.integers(0, 256, (200, 300, 3), dtype=np.uint8)
    cases = []
    for (w, h) in [(640, 320), (333, 211), (17, 9), (5, 3), (2, 2), (1, 1)]:
        img = np.ascontiguousarray(smooth[:h, :w])
        for sub in (0, 1, 2):
            cases.append((f"{w}x{h} subsampling {sub}", encode(img, quality=92, subsampling=sub)))
        cases.append((f"{w}x{h} grey", encode(np.ascontiguousarray(img[..., 0]), quality=80)))
    for q in (5, 30, 100):
        for sub in (0, 1, 2):
            cases.append((f"noise q{q} subsampling {sub}", encode(noise, quality=q, subsampling=sub)))
    cases.append(("optimised Huffman tables", encode(noise, quality=90, subsampling=2, optimize=True)))
    cases.append(("restart markers", encode(noise, quality=90, subsampling=2, restart_marker_blocks=3)))
    cases.append(("restart markers 4:4:4", encode(noise, quality=90, subsampling=0, restart_marker_rows=1)))
    return cases


def test_entropy_decode_and_oracle_equal_pil(rt):
    for name, data in jpeg_cases(rt):
        info, coef = rt.jpeg_entropy_decode(data)
        out, ref = jpeg_oracle.decode(info, coef), pil_decode(data)
        assert out.shape == ref.shape, name
        assert np.array_equal(out, ref), f"{name}: {(out != ref).mean():.4f} of the bytes differ"


def test_reference_earth_image_equals_pil(rt):
    p = os.path.join(ROOT, "assets", "earth-large.jpg")
    if not os.path.exists(p):
        pytest.skip("assets/earth-large.jpg not shipped")
    data = open(p, "rb").read()
    info, coef = rt.jpeg_entropy_decode(data)
    assert (info.width, info.height, info.components) == (6400, 3200, 3)
    assert list(info.h_samp) == [2, 1, 1] and list(info.v_samp) == [2, 1, 1]          # 4:2:0
    assert np.array_equal(jpeg_oracle.decode(info, coef), pil_decode(data))


def test_unsupported_and_malformed_streams(rt):
    A = rt._abi
    img = rt.synthetic_earth(64, 32, seed=1)
    with pytest.raises(A.RtError) as e:
        rt.jpeg_entropy_decode(encode(img, progressive=True))
    assert e.value.status == A.RT_ERR_UNSUPPORTED
    with pytest.raises(A.RtError) as e:
        rt.jpeg_entropy_decode(b"not a jpeg at all, just bytes")
    assert e.value.status == A.RT_ERR_INVALID_ARGUMENT
    good = encode(img, quality=90)
    with pytest.raises(A.RtError):
        rt.jpeg_entropy_decode(good[: len(good) // 3])               # truncated inside the scan... or before it
    from PIL import Image
    buf = io.BytesIO()
    Image.fromarray(img).convert("CMYK").save(buf, format="JPEG")
    with pytest.raises(A.RtError) as e:
        rt.jpeg_entropy_decode(buf.getvalue())
    assert e.value.status == A.RT_ERR_UNSUPPORTED
