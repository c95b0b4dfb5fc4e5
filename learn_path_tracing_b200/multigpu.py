"""Multi-GPU rendering: sample-split data parallelism + one NCCL reduce.

The reference has no multi-GPU path (it pins one device with CUDA_VISIBLE_DEVICES, 15_module.py:10).
Here every rank (one process per GPU, torchrun) holds a replica of the scene and renders a disjoint
range of sample indices for every pixel; the counter-based RNG is keyed on (seed, pixel, sample, bounce)
so the union over ranks is exactly the set of paths a single GPU would trace.  The per-GPU accumulators are
then summed onto rank 0 with torch.distributed.reduce (NCCL over NVLink; gloo in the CPU tests) and rank 0
runs the fused divide-by-spp + tonemap kernel.

Nothing on this path waits for the GPU on the host: pt_render is enqueued without statistics, the reduce follows on
the same stream, and the only synchronisation is rank 0's read of the finished image.  `bands` > 1 renders the frame in
horizontal bands and reduces band k on a side stream while band k + 1 renders; measured on 2 B200s this LOSES to one
band (8_refract 34.9 / 33.6 / 31.5 Gpaths/s at 1 / 4 / 8 bands, Yoimiya 9.6 / 8.2 / 7.3): every band pays the ramp-up and
tail of its own persistent launch while the reduce it hides is ~0.1 ms — so 1 is the default everywhere.
Works for the v2 sphere worlds and for the legacy mesh worlds (the shading model follows the world type;
legacy/PT_in_one_weekend/15_module.py:1022-1036 is the loop it replaces).
"""
from __future__ import annotations

from . import _lib


def split_samples(spp: int, world_size: int, rank: int) -> tuple[int, int]:
    """(first sample index, number of samples) of `rank`: contiguous ranges, remainder to the low ranks."""
    base, rem = divmod(int(spp), int(world_size))
    count = base + (1 if rank < rem else 0)
    offset = rank * base + min(rank, rem)
    return offset, count


def _dist_active(group=None) -> bool:
    import torch.distributed as dist
    return dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1


def _reduce(t, dst, group):
    """dist.reduce(SUM) of a contiguous tensor.  NCCL reduces device memory in place over NVLink; a process group that
    cannot (gloo with CUDA tensors: the 1-GPU form of the multi-rank test) is served through a host copy."""
    import torch.distributed as dist
    if t.is_cuda and dist.get_backend(group) != "nccl":
        h = t.cpu()
        dist.reduce(h, dst=dst, op=dist.ReduceOp.SUM, group=group)
        if dist.get_rank(group) == dst:
            t.copy_(h)
        return
    dist.reduce(t, dst=dst, op=dist.ReduceOp.SUM, group=group)


def reduce_accumulators(accum, dst: int = 0, group=None, channels: int = 4):
    """Sum the per-rank accumulators ([H*W,4] float: r, g, b, contributing paths) onto `dst`, in place.  channels=3
    ships only the radiance (the path count is a diagnostic): 25 % fewer bytes over NVLink at the price of one
    packing copy on every rank.  No-op without an initialised process group."""
    import torch.distributed as dist
    if not _dist_active(group):
        return accum
    if channels >= accum.shape[-1]:
        _reduce(accum, dst, group)
        return accum
    rgb = accum[:, :channels].contiguous()
    _reduce(rgb, dst, group)
    if dist.get_rank(group) == dst:
        accum[:, :channels].copy_(rgb)
    return accum


def world_shading_model(world, bsdf=None) -> int:
    """Legacy (mesh / textured-sphere) worlds shade with gen_secondary_rays (15_module.py:994-1013); v2 worlds with the
    BSDF class the script names (default Metal/Dielectric)."""
    if getattr(world, "shading_model", None) is not None and bsdf is None:
        return int(world.shading_model)
    return int(getattr(bsdf, "shading_model", _lib.PT_SHADE_V2))


def render_split_reduce(renderer, scene, cam_struct, spp: int, max_depth: int, model: int, seed: int, group=None,
                        absorptivity: float = 0.25, bands: int = 1, channels: int = 3, flags: int = 0,
                        first_sample: int = 0, **render_kw):
    """This rank's share of `spp` samples into renderer.accum, then the sum over ranks onto rank 0 — all enqueued, nothing
    awaited.  Returns (sample offset, sample count) of this rank."""
    import torch
    import torch.distributed as dist
    active = _dist_active(group)
    ws = dist.get_world_size(group) if active else 1
    rank = dist.get_rank(group) if active else 0
    offset, count = split_samples(spp, ws, rank)
    offset += int(first_sample)  # progressive passes: this pass covers sample indices [first_sample, first_sample + spp)
    H, W = renderer.height, renderer.width
    bands = max(1, min(int(bands), H)) if active else 1
    if bands == 1:
        renderer.render(scene, cam_struct, count, max_depth, model, seed, spp_offset=offset, absorptivity=absorptivity,
                        flags=flags, want_stats=False, **render_kw)
        reduce_accumulators(renderer.accum, 0, group, channels)
        return offset, count
    main = torch.cuda.current_stream()
    side = renderer.side_stream()
    edges = [H * b // bands for b in range(bands + 1)]
    rows = renderer.accum.view(H, W, 4)
    for b in range(bands):
        y0, y1 = edges[b], edges[b + 1]
        renderer.render(scene, cam_struct, count, max_depth, model, seed, spp_offset=offset, absorptivity=absorptivity,
                        flags=flags, want_stats=False, rows=(y0, y1), count_samples=(b == bands - 1), **render_kw)
        ev = torch.cuda.Event()
        ev.record(main)
        with torch.cuda.stream(side):  # band b travels while band b + 1 renders
            side.wait_event(ev)
            _reduce(rows[y0:y1], 0, group)  # whole rows are contiguous: no packing
    main.wait_stream(side)
    return offset, count


def render_distributed(world, camera, spp: int = 8192, propagate_limit: int = 32, seed: int = 1, bsdf=None,
                       ctx=None, group=None, render_accum=None, postprocess=True, absorptivity: float = 0.25,
                       bands: int = 1, channels: int = 3):
    """render(world, camera) across all ranks of the process group (v2 World or legacy World).

    Returns the image ([W,H,3] float32; v2: ACES + gamma, legacy: gamma only, as the two reference scripts do; the
    linear mean when not `postprocess`) on rank 0 and None elsewhere.
    `render_accum(offset, count) -> tensor[H*W,4]` can replace the CUDA renderer (the gloo/CPU tests inject
    the oracle there); by default it is Renderer.render on this rank's GPU.
    """
    import torch.distributed as dist
    active = _dist_active(group)
    ws = dist.get_world_size(group) if active else 1
    rank = dist.get_rank(group) if active else 0
    w, h = camera.resolution
    model = world_shading_model(world, bsdf)
    legacy = model == _lib.PT_SHADE_LEGACY
    if render_accum is None:
        from .render import Renderer, default_context
        ctx = ctx or default_context()
        renderer = Renderer(w, h, ctx)
        render_split_reduce(renderer, world.device_scene(ctx), camera.to_struct(), spp, propagate_limit, model, seed, group,
                            absorptivity=absorptivity, bands=bands, channels=channels)
        if rank != 0:
            return None
        if postprocess:
            return renderer.image(aces=not legacy, gamma=2.2, total_spp=spp)
        return renderer.ctx.download_accum(renderer.accum.data_ptr(), w, h) / float(spp)
    offset, count = split_samples(spp, ws, rank)
    accum = render_accum(offset, count)
    reduce_accumulators(accum, 0, group, channels)
    if rank != 0:
        return None
    img = accum[:, :3].reshape(h, w, 3).permute(1, 0, 2).contiguous().cpu().numpy() / float(spp)
    if postprocess:
        from .postprocessing import ACES_tonemapping, gamma_correction
        img = gamma_correction(img if legacy else ACES_tonemapping(img), 2.2)
    return img
