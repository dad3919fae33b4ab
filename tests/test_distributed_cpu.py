"""N>1 host logic on the CPU: world_size-2 gloo processes shard the sample range, render their shares
(with the oracle standing in for the device kernel) and reduce to rank 0 through the product's own
render_sharded(); the result must equal the single-process render."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_samples(rt):
    from rust_tracing_b200.distributed import shard_samples
    for spp in (0, 1, 7, 8, 100, 10000):
        for world in (1, 2, 3, 4, 8):
            parts = [shard_samples(spp, r, world, sample_begin=5) for r in range(world)]
            assert parts[0][0] == 5
            assert sum(c for _, c in parts) == spp
            for (b0, c0), (b1, _) in zip(parts, parts[1:]):
                assert b0 + c0 == b1
            assert max(c for _, c in parts) - min(c for _, c in parts) <= 1
    assert shard_samples(10000, 3, 8) == (3750, 1250)
    with pytest.raises(ValueError):
        shard_samples(10, 2, 2)


def _worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import rust_tracing_b200 as rt
    from rust_tracing_b200.distributed import render_sharded
    from oracle import binding as ob
    s, cs = rt.builtin_scene("cornell_smoke", image_width=24, samples_per_pixel=9)
    cam = rt.Camera(cs)
    h, w = cam.shape

    def fn(begin, count, fb):
        img, _ = ob.render(s.desc, cam, begin, count, seed=4, mode=0, threads=1)
        fb[..., :3] += torch.from_numpy(img)
        fb[..., 3] += count

    fb = torch.zeros((h, w, 4), dtype=torch.float64)
    render_sharded(fn, fb, 9, rank, world)
    if rank == 0:
        np.save(out_path, fb.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_render_equals_single(rt, ob, tmp_path):
    out = str(tmp_path / "fb.npy")
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = np.load(out)
    s, cs = rt.builtin_scene("cornell_smoke", image_width=24, samples_per_pixel=9)
    cam = rt.Camera(cs)
    want, _ = ob.render(s.desc, cam, 0, 9, seed=4, mode=0)
    assert np.all(got[..., 3] == 9)
    assert np.allclose(got[..., :3], want, rtol=1e-12, atol=1e-12)   # f64 partial sums in a different order


def test_weighted_shards_cover_the_range_exactly():
    """shard_samples_weighted: the proportional split bench.py uses at N > 1 (calibrated on a warm-up step)."""
    from rust_tracing_b200.distributed import shard_samples_weighted
    for spp, weights in ((10000, [1.0, 1.02, 0.97, 1.0, 0.95, 1.01, 1.0, 0.99]), (7, [3.0, 1.0]), (5, [1.0, 0.0, 3.0]), (0, [1.0, 2.0]),
                         (13, [0.0, 0.0, 0.0]), (1, [0.2, 0.5, 0.3])):
        parts = shard_samples_weighted(spp, weights, sample_begin=11)
        assert len(parts) == len(weights)
        assert parts[0][0] == 11 and all(parts[k + 1][0] == parts[k][0] + parts[k][1] for k in range(len(parts) - 1))
        assert sum(c for _, c in parts) == spp and all(c >= 0 for _, c in parts)
        tot = sum(weights)
        if tot > 0:
            assert all(abs(c - spp * w / tot) <= 1.0 for (_, c), w in zip(parts, weights))
    with pytest.raises(ValueError):
        shard_samples_weighted(-1, [1.0])
    with pytest.raises(ValueError):
        shard_samples_weighted(4, [1.0, float("nan")])
