"""Multi-GPU sharding: one process per GPU, no collective inside the wavefront loop.

The path shards the way the reference already does across threads (renderer.rs:63-95: independent
tiles, one gather at the end).  Here each rank renders a contiguous range of absolute sample
indices over the full image (Philox is keyed by the absolute sample index, so the union equals the
single-GPU image up to f32 summation order) or, alternatively, the tiles with index % world == rank;
the only communication is one framebuffer sum to rank 0 (torch.distributed: NCCL over NVLink on
GPUs, gloo in the CPU tests).
"""
from __future__ import annotations


def shard_samples(total_spp: int, rank: int, world: int) -> tuple[int, int]:
    """(first absolute sample, number of samples) of ``rank``: contiguous, remainder to the low ranks."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad rank/world")
    base, rem = divmod(total_spp, world)
    count = base + (1 if rank < rem else 0)
    begin = rank * base + min(rank, rem)
    return begin, count


def shard_params(make_params, total_spp: int, rank: int, world: int, mode: str = "samples", **kw):
    """Render parameters of one rank.  ``make_params`` is ``api.make_params``; the result always asks for raw sums
    (``RENDER_RAW_SUM``) so that shards add up; divide by ``total_spp`` after the reduce."""
    from . import capi
    flags = kw.pop("flags", 0) | capi.RENDER_RAW_SUM
    if mode == "samples":
        begin, count = shard_samples(total_spp, rank, world)
        # world > total_spp leaves some ranks without samples: sample_count 0 would mean "all of them" to the C ABI
        # (shimmer_b200.h), -1 is its explicit "none" (the rank still takes part in the reduce with a zero framebuffer)
        return make_params(spp=total_spp, sample_begin=begin, sample_count=count if count > 0 else -1, flags=flags, **kw), count
    if mode == "tiles":
        return make_params(spp=total_spp, tile_rank=rank, tile_world=world, flags=flags, **kw), total_spp
    raise ValueError("mode must be 'samples' or 'tiles'")


def reduce_framebuffer(fb, total_spp: int, dst: int = 0):
    """Sum the per-rank raw framebuffers onto ``dst`` and turn the sums into the mean (renderer.rs:147).
    ``fb`` is a torch tensor (CUDA with NCCL, CPU with gloo).  Returns the image on ``dst``, None elsewhere."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(fb, dst=dst, op=dist.ReduceOp.SUM)
        if dist.get_rank() != dst:
            return None
    return fb / float(total_spp)
