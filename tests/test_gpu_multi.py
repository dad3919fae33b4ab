"""Multi-GPU through the C ABI (rt_render_multi: sample-range sharding inside one process, one ncclReduce to the first
device). Needs >= 2 GPUs; skipped on a single-GPU box (run with `gpurun --gpus 2`)."""
import numpy as np
import pytest

from conftest import small_scene

pytestmark = pytest.mark.gpu


def gpu_count():
    import torch
    return torch.cuda.device_count()


def test_render_multi_single_device_is_rt_render(rt, ctx):
    s, cam = small_scene(rt, 6, width=64)
    ds = ctx.upload(s)
    one = ctx.render(ds, cam, 3, 9, seed=2)
    multi, shares = rt.render_multi([ctx], [ds], cam, 3, 9, seed=2)
    assert shares == [9]
    assert np.allclose(multi, one, rtol=2e-6, atol=1e-6)
    with pytest.raises(rt._abi.RtError):
        rt.render_multi([ctx, ctx], [ds, ds], cam, 0, 4)          # one context per device
    with pytest.raises(rt._abi.RtError):
        rt.render_multi([ctx], [ds], cam, 0, 4, weights=[0.0])
    ds.close()


@pytest.mark.parametrize("idx", [6, 8])
def test_render_multi_two_gpus_equals_one(rt, ctx, earth, idx):
    if gpu_count() < 2:
        pytest.skip("needs 2 GPUs")
    s, cam = small_scene(rt, idx, earth, width=96)
    ctx1 = rt.Context(1)
    ds0, ds1 = ctx.upload(s), ctx1.upload(s)
    spp = 33
    one = ctx.render(ds0, cam, 0, spp, seed=4)
    for weights, want in ((None, [17, 16]), ([3.0, 1.0], [25, 8]), ([1.0, 1e-9], [33, 0])):
        multi, shares = rt.render_multi([ctx, ctx1], [ds0, ds1], cam, 0, spp, seed=4, weights=weights)
        assert shares == want
        assert np.all(multi[..., 3] == spp)
        # the same set of keyed paths; only the f32 summation order differs
        assert np.allclose(multi, one, rtol=1e-5, atol=1e-5 * spp)
    ds0.close(); ds1.close(); ctx1.close()
