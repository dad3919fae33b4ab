"""CPU restatement of the back end of libjpeg's baseline decoder - TEST INFRASTRUCTURE ONLY (see oracle.cpp's header).

The reference decodes its image texture with the `image` crate (texture.rs:76-80; image = "0.24.7", Cargo.toml:14, no
lockfile, so the exact jpeg-decoder version is unpinned and absent from /root/reference). Every input of this repo's
oracle and device comes from PIL = libjpeg-turbo instead, so libjpeg-turbo's default decode is the
anchor of rt_jpeg_decode, and this file restates its published algorithms in numpy integer arithmetic:
  idct_islow        jidctint.c jpeg_idct_islow (with DEQUANTIZE): 13-bit constants, 2 extra bits between the passes
  upsample_*_fancy  jdsample.c h2v1_fancy_upsample / h2v2_fancy_upsample (context rows at the image edge: jdmainct.c)
  ycc_to_rgb        jdcolor.c build_ycc_rgb_table / ycc_rgb_convert
Pinned: tests/test_jpeg_cpu.py checks decode() byte for byte against PIL on the reference's own asset
(assets/earth-large.jpg) and on generated files of every supported chroma layout, odd sizes included.
Input: the quantised coefficients of rt_jpeg_entropy_decode.
"""
import numpy as np

C = dict(f0_298=2446, f0_390=3196, f0_541=4433, f0_765=6270, f0_899=7373, f1_175=9633, f1_501=12299, f1_847=15137, f1_961=16069,
         f2_053=16819, f2_562=20995, f3_072=25172)


def _descale(x, n):
    return (x + (1 << (n - 1))) >> n


def _pass(v, shift):
    """one 8-point pass over axis -1 of int64 array v"""
    z2, z3 = v[..., 2], v[..., 6]
    z1 = (z2 + z3) * C["f0_541"]
    tmp2 = z1 + z3 * (-C["f1_847"])
    tmp3 = z1 + z2 * C["f0_765"]
    z2, z3 = v[..., 0], v[..., 4]
    tmp0, tmp1 = (z2 + z3) << 13, (z2 - z3) << 13
    tmp10, tmp13, tmp11, tmp12 = tmp0 + tmp3, tmp0 - tmp3, tmp1 + tmp2, tmp1 - tmp2
    t0, t1, t2, t3 = v[..., 7], v[..., 5], v[..., 3], v[..., 1]
    z1, z2, z3, z4 = t0 + t3, t1 + t2, t0 + t2, t1 + t3
    z5 = (z3 + z4) * C["f1_175"]
    t0, t1, t2, t3 = t0 * C["f0_298"], t1 * C["f2_053"], t2 * C["f3_072"], t3 * C["f1_501"]
    z1, z2, z3, z4 = z1 * -C["f0_899"], z2 * -C["f2_562"], z3 * -C["f1_961"] + z5, z4 * -C["f0_390"] + z5
    t0, t1, t2, t3 = t0 + z1 + z3, t1 + z2 + z4, t2 + z2 + z3, t3 + z1 + z4
    out = np.stack([tmp10 + t3, tmp11 + t2, tmp12 + t1, tmp13 + t0, tmp13 - t0, tmp12 - t1, tmp11 - t2, tmp10 - t3], axis=-1)
    return _descale(out, shift)


def idct_islow(coef, quant):
    """coef: (..., 8, 8) int16 quantised, quant: (8, 8) -> (..., 8, 8) uint8 samples."""
    v = coef.astype(np.int64) * quant.astype(np.int64)
    ws = np.swapaxes(_pass(np.swapaxes(v, -1, -2), 13 - 2), -1, -2)      # pass 1: columns
    out = _pass(ws, 13 + 2 + 3)                                          # pass 2: rows
    return np.clip(out + 128, 0, 255).astype(np.uint8)


def plane_from_blocks(samples, blocks_h, blocks_w):
    return samples.reshape(blocks_h, blocks_w, 8, 8).transpose(0, 2, 1, 3).reshape(blocks_h * 8, blocks_w * 8)


def upsample_h2v1_fancy(p):
    p = p.astype(np.int32)
    h, w = p.shape
    if w <= 2:
        return np.repeat(p, 2, axis=1)
    out = np.empty((h, 2 * w), dtype=np.int32)
    prev = np.concatenate([p[:, :1], p[:, :-1]], axis=1)
    nxt = np.concatenate([p[:, 1:], p[:, -1:]], axis=1)
    out[:, 0::2] = (3 * p + prev + 1) >> 2
    out[:, 1::2] = (3 * p + nxt + 2) >> 2
    out[:, 0] = p[:, 0]
    out[:, -1] = p[:, -1]
    return out


def upsample_h2v2_fancy(p):
    p = p.astype(np.int32)
    h, w = p.shape
    if w <= 2:
        return np.repeat(np.repeat(p, 2, axis=0), 2, axis=1)
    above = np.concatenate([p[:1], p[:-1]], axis=0)
    below = np.concatenate([p[1:], p[-1:]], axis=0)
    out = np.empty((2 * h, 2 * w), dtype=np.int32)
    for v, far in ((0, above), (1, below)):
        s = 3 * p + far                                               # column sums
        last = np.concatenate([s[:, :1], s[:, :-1]], axis=1)
        nxt = np.concatenate([s[:, 1:], s[:, -1:]], axis=1)
        even = (3 * s + last + 8) >> 4
        odd = (3 * s + nxt + 7) >> 4
        even[:, 0] = (4 * s[:, 0] + 8) >> 4
        odd[:, -1] = (4 * s[:, -1] + 7) >> 4
        out[v::2, 0::2] = even
        out[v::2, 1::2] = odd
    return out


def ycc_to_rgb(y, cb, cr):
    y, u, w = y.astype(np.int32), cb.astype(np.int32) - 128, cr.astype(np.int32) - 128
    r = y + ((91881 * w + 32768) >> 16)
    g = y + ((-22554 * u + 32768 - 46802 * w) >> 16)
    b = y + ((116130 * u + 32768) >> 16)
    return np.clip(np.stack([r, g, b], axis=-1), 0, 255).astype(np.uint8)


def decode(info, coef):
    """info: rt_jpeg_info (ctypes), coef: int16 array -> uint8 (H, W, 3)."""
    W, H, n = info.width, info.height, info.components
    hmax, vmax = max(info.h_samp[:n]), max(info.v_samp[:n])
    planes = []
    for c in range(n):
        bw, bh = info.blocks_w[c], info.blocks_h[c]
        blocks = coef[info.coef_offset[c]: info.coef_offset[c] + bw * bh * 64].reshape(bh * bw, 8, 8)
        q = np.array(info.quant[c][:], dtype=np.int64).reshape(8, 8)
        p = plane_from_blocks(idct_islow(blocks, q), bh, bw)
        ds_w = (W * info.h_samp[c] + hmax - 1) // hmax
        ds_h = (H * info.v_samp[c] + vmax - 1) // vmax
        p = p[:ds_h, :ds_w]
        fx, fy = hmax // info.h_samp[c], vmax // info.v_samp[c]
        if (fx, fy) == (2, 1):
            p = upsample_h2v1_fancy(p)
        elif (fx, fy) == (2, 2):
            p = upsample_h2v2_fancy(p)
        elif (fx, fy) != (1, 1):
            raise ValueError("unsupported chroma layout")
        planes.append(p[:H, :W])
    if n == 1:
        return np.repeat(planes[0].astype(np.uint8)[..., None], 3, axis=-1)
    if info.adobe_rgb:
        return np.stack(planes, axis=-1).astype(np.uint8)
    return ycc_to_rgb(*planes)
