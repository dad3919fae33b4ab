"""The wavefront renderer (render_v4.cuh, RT_B200_KERNEL=4: shade / extend kernels over a pool of in-flight paths) traces
the same keyed paths with the same device functions as the megakernel (render_v3.cuh, the default), so the two SUM
framebuffers may differ only by the order of the f32 additions and their path / segment counts must be identical."""
import os

import numpy as np
import pytest

from conftest import small_scene

pytestmark = pytest.mark.gpu


def context(rt, **env):
    old = {k: os.environ.get(k) for k in env}
    os.environ.update({k: str(v) for k, v in env.items()})
    try:
        return rt.Context(0)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


@pytest.mark.parametrize("idx,spp", [(8, 6), (6, 8), (7, 8), (0, 16), (3, 8), (2, 4)])
def test_wavefront_matches_megakernel(rt, ctx, earth, idx, spp):
    s, cam = small_scene(rt, idx, earth)
    ref_ds = ctx.upload(s)
    ref = ctx.render(ref_ds, cam, 2, spp, seed=9)
    ref_stats = ctx.stats()
    ref_ds.close()
    for pool in (1 << 20, 2048):          # a tiny pool: many iterations, every reserve / window edge case
        wf = context(rt, RT_B200_KERNEL=4, RT_B200_POOL=pool)
        ds = wf.upload(s)
        img = wf.render(ds, cam, 2, spp, seed=9)
        st = wf.stats()
        ds.close()
        wf.close()
        assert np.array_equal(img[..., 3], ref[..., 3])                      # every path added exactly once
        assert (st["paths"], st["segments"]) == (ref_stats["paths"], ref_stats["segments"])
        scale = np.maximum(np.abs(ref[..., :3]), 1e-3 * max(1.0, float(np.abs(ref[..., :3]).max())))
        assert float((np.abs(img[..., :3] - ref[..., :3]) / scale).max()) < 1e-4
