"""Multi-GPU sharding of the render loop: one process per GPU, the sample range of every
pixel is split across ranks, each rank accumulates its share into its own float4 SUM framebuffer, and the
partial framebuffers are summed onto rank 0 with ONE reduce (NCCL over NVLink on GPUs; gloo on CPU in tests).
The keyed RNG (pixel, sample) makes the union of the shards the same set of paths a single GPU traces.
"""
import torch
import torch.distributed as dist


def shard_samples(spp, rank, world_size, sample_begin=0):
    """Contiguous, balanced split of [sample_begin, sample_begin + spp): returns (begin, count) for `rank`.
    The first (spp % world_size) ranks get one extra sample; empty shards are allowed (count 0)."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("bad rank / world_size")
    if spp < 0:
        raise ValueError("spp must be >= 0")
    base, extra = divmod(spp, world_size)
    begin = sample_begin + rank * base + min(rank, extra)
    return begin, base + (1 if rank < extra else 0)


def shard_samples_weighted(spp, weights, sample_begin=0):
    """Contiguous split of [sample_begin, sample_begin + spp) in proportion to `weights` (one per rank: e.g. the paths/s
    each GPU reached in a warm-up step). Returns [(begin, count)] for every rank; counts add up to spp exactly (largest
    remainders get the leftover samples), so the union is the same set of (pixel, sample) paths a single GPU traces."""
    if spp < 0 or not weights or any((not w == w) or w < 0 for w in weights):
        raise ValueError("bad spp / weights")
    total = float(sum(weights))
    if total <= 0.0:
        return [shard_samples(spp, r, len(weights), sample_begin) for r in range(len(weights))]
    exact = [spp * w / total for w in weights]
    counts = [int(x) for x in exact]
    left = spp - sum(counts)
    for r in sorted(range(len(weights)), key=lambda r: exact[r] - counts[r], reverse=True)[:left]:
        counts[r] += 1
    out, b = [], sample_begin
    for c in counts:
        out.append((b, c))
        b += c
    return out


def reduce_to_root(framebuffer, group=None):
    """Sum the per-rank partial SUM framebuffers onto rank 0 (the only collective of the path)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.reduce(framebuffer, dst=0, op=dist.ReduceOp.SUM, group=group)
    return framebuffer


def render_sharded(render_fn, framebuffer, spp, rank, world_size, sample_begin=0, group=None):
    """render_fn(begin, count, framebuffer) accumulates samples [begin, begin+count) into `framebuffer`
    (a torch tensor, device float4 sums on GPUs). Returns the framebuffer; complete on rank 0 after the reduce."""
    begin, count = shard_samples(spp, rank, world_size, sample_begin)
    if count > 0:
        render_fn(begin, count, framebuffer)
    return reduce_to_root(framebuffer, group)


def render_on_gpu(ctx, dscene, cam, spp, seed, rank, world_size, framebuffer=None, sample_begin=0, group=None):
    """The GPU instance of render_sharded: rt_render_accumulate on torch's current stream, then the reduce."""
    h, w = cam.shape
    if framebuffer is None:
        framebuffer = torch.zeros((h, w, 4), dtype=torch.float32, device=f"cuda:{ctx.device_id}")
    stream = torch.cuda.current_stream(framebuffer.device).cuda_stream

    def fn(begin, count, fb):
        ctx.render_accumulate(dscene, cam, begin, count, seed, fb.data_ptr(), stream)

    return render_sharded(fn, framebuffer, spp, rank, world_size, sample_begin, group)
