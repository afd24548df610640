"""Multi-GPU decomposition of one render: samples-per-pixel are partitioned across ranks (one process per GPU), every
rank renders its sample range of every pixel into an f64 accumulation buffer, and one reduce (NCCL over NVLink when the
buffers are CUDA tensors, gloo in the CPU tests) sums the buffers on rank 0, which runs the tone-map epilogue.

The reference has no counterpart (single process). Philox is keyed by the GLOBAL sample index, so the
union of the ranks' samples is the same set of paths for every world size.
"""
from __future__ import annotations


def partition_samples(spp: int, world_size: int, rank: int):
    """Contiguous, balanced split of sample indices [0, spp): returns (sample_offset, count) of `rank`.
    The first spp % world_size ranks get one extra sample; empty shares are allowed (spp < world_size)."""
    if world_size <= 0 or not (0 <= rank < world_size) or spp < 0:
        raise ValueError("bad partition arguments")
    base, extra = divmod(int(spp), int(world_size))
    count = base + (1 if rank < extra else 0)
    offset = rank * base + min(rank, extra)
    return offset, count


def reduce_accumulators(acc, dst: int = 0, group=None):
    """Sum the per-rank f64 accumulation buffers onto `dst` (torch.distributed.reduce). A no-op without a process group."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.reduce(acc, dst=dst, op=dist.ReduceOp.SUM, group=group)
    return acc


def render_share(ctx, scene, cam, make_params_fn, spp_total: int, world_size: int, rank: int, acc, want_stats=False):
    """Render this rank's share of `spp_total` samples into the CUDA f64 tensor `acc` ([H*W*3], zeroed by the caller).
    `make_params_fn(spp, sample_offset, spp_total)` builds the trt_params."""
    offset, count = partition_samples(spp_total, world_size, rank)
    if count == 0:
        return None
    p = make_params_fn(count, offset, spp_total)
    st = scene.render_accumulate(ctx, cam, p, acc.data_ptr(), want_stats=want_stats)
    _order_after_render(ctx)
    return st


def _order_after_render(ctx):
    """The render kernel runs on the ctx's stream; whatever consumes `acc` next (the NCCL reduce, a torch op) runs on
    torch's current stream. Unless the two are the same stream (`ctx.set_stream(torch.cuda.current_stream().cuda_stream)`,
    what bench.py does), wait for the render here so that the consumer cannot race with it."""
    try:
        import torch
        same = torch.cuda.is_available() and ctx.stream_handle is not None and \
            int(ctx.stream_handle) == int(torch.cuda.current_stream(ctx.device).cuda_stream)
    except Exception:
        same = False
    if not same:
        ctx.synchronize()
