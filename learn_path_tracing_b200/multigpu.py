"""Multi-GPU rendering: sample-split data parallelism + one NCCL reduce.

The reference has no multi-GPU path (it pins one device with CUDA_VISIBLE_DEVICES, 15_module.py:10).
Here every rank (one process per GPU, torchrun) holds a replica of the scene and renders a disjoint
range of sample indices for every pixel; the counter-based RNG is keyed on (seed, pixel, sample, bounce)
so the union over ranks is exactly the set of paths a single GPU would trace.  The per-GPU float4
accumulators are then summed onto rank 0 with one torch.distributed.reduce (NCCL over NVLink; gloo in
the CPU tests) and rank 0 runs the fused divide-by-spp + tonemap kernel.
"""
from __future__ import annotations

from . import _lib


def split_samples(spp: int, world_size: int, rank: int) -> tuple[int, int]:
    """(first sample index, number of samples) of `rank`: contiguous ranges, remainder to the low ranks."""
    base, rem = divmod(int(spp), int(world_size))
    count = base + (1 if rank < rem else 0)
    offset = rank * base + min(rank, rem)
    return offset, count


def reduce_accumulators(accum, dst: int = 0, group=None):
    """Sum the per-rank accumulators onto `dst` (in place).  No-op without an initialised process group."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.reduce(accum, dst=dst, op=dist.ReduceOp.SUM, group=group)
    return accum


def render_distributed(world, camera, spp: int = 8192, propagate_limit: int = 32, seed: int = 1, bsdf=None,
                       ctx=None, group=None, render_accum=None, postprocess=True):
    """render(world, camera) across all ranks of the process group.

    Returns the image ([W,H,3] float32, tonemapped when `postprocess`) on rank 0 and None elsewhere.
    `render_accum(offset, count) -> tensor[H*W,4]` can replace the CUDA renderer (the gloo/CPU tests inject
    the oracle there); by default it is Renderer.render on this rank's GPU.
    """
    import torch.distributed as dist
    ws = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    offset, count = split_samples(spp, ws, rank)
    w, h = camera.resolution
    renderer = None
    if render_accum is None:
        from .render import Renderer, default_context
        ctx = ctx or default_context()
        renderer = Renderer(w, h, ctx)
        model = getattr(bsdf, "shading_model", _lib.PT_SHADE_V2)
        renderer.render(world.device_scene(ctx), camera.to_struct(), count, propagate_limit, model, seed,
                        spp_offset=offset)
        accum = renderer.accum
    else:
        accum = render_accum(offset, count)
    reduce_accumulators(accum, 0, group)
    if rank != 0:
        return None
    if renderer is not None:
        if postprocess:
            return renderer.image(aces=True, gamma=2.2, total_spp=spp)
        return renderer.ctx.download_accum(accum.data_ptr(), w, h) / float(spp)
    img = accum[:, :3].reshape(h, w, 3).permute(1, 0, 2).contiguous().cpu().numpy() / float(spp)
    if postprocess:
        from .postprocessing import ACES_tonemapping, gamma_correction
        img = gamma_correction(ACES_tonemapping(img), 2.2)
    return img
