"""render(): the reference's render loop (10_final/__main__.py:99-103) as ONE call into the device.

The reference launches get_rays + shader once per sample from Python (2*spp launches); here the host
hands (scene, camera, params) to pt_render and the wavefront runs on the GPU until every path of
every sample has terminated.  torch is used for device buffers and streams only.
"""
from __future__ import annotations

import os

import numpy as np

from . import _lib

_default_ctx = None


def default_context() -> _lib.Context:
    """Process-wide context on cuda:LOCAL_RANK, bound to torch's current stream."""
    global _default_ctx
    if _default_ctx is None:
        import torch
        if not torch.cuda.is_available():
            raise _lib.PtError("no CUDA device visible: learn_path_tracing_b200 has no CPU fallback")
        dev = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(dev)
        _default_ctx = _lib.Context(dev, torch.cuda.current_stream().cuda_stream)
    return _default_ctx


class Renderer:
    """Owns the device accumulators for one image and drives pt_render / pt_postprocess."""

    def __init__(self, width: int, height: int, ctx: _lib.Context | None = None, want_sq: bool = False):
        import torch
        self.torch = torch
        self.ctx = ctx or default_context()
        self.width, self.height = int(width), int(height)
        dev = torch.device("cuda", self.ctx.device)
        self.accum = torch.zeros((self.height * self.width, 4), dtype=torch.float32, device=dev)
        self.accum_sq = torch.zeros_like(self.accum) if want_sq else None
        self.spp_done = 0
        self.last_stats = None

    def clear(self):
        self.accum.zero_()
        if self.accum_sq is not None:
            self.accum_sq.zero_()
        self.spp_done = 0

    def render(self, scene: _lib.Scene, cam: _lib.PtCamera, spp: int, max_depth: int,
               shading_model: int = _lib.PT_SHADE_V2, seed: int = 1, spp_offset: int | None = None,
               absorptivity: float = 0.25, flags: int = 0, pool_capacity: int = 0, mode: int = 0,
               segments_per_launch: int = 0, shade_min: int = 0, serve_min: int = 0,
               want_stats: bool = True, rows: tuple[int, int] | None = None,
               count_samples: bool = True) -> _lib.PtStats | None:
        """Adds `spp` more samples per pixel into the accumulators (progressive, legacy render(moved=False)).
        want_stats=False returns None without waiting for the GPU (Renderer.stats() fetches them later).
        rows=(y0, y1) renders only that band of image rows (persistent kernel); count_samples=False leaves spp_done alone
        (all bands but the last of one banded pass)."""
        p = _lib.PtRenderParams()
        p.width, p.height = self.width, self.height
        p.spp = int(spp)
        p.spp_offset = int(self.spp_done if spp_offset is None else spp_offset)
        p.max_depth = int(max_depth)
        p.shading_model = int(shading_model)
        p.seed = int(seed) & 0xFFFFFFFF
        p.absorptivity = float(absorptivity)
        p.pool_capacity = int(pool_capacity)
        p.flags = int(flags) | (_lib.PT_FLAG_ACCUM_SQ if self.accum_sq is not None else 0)
        p.reserved[0] = int(mode)                 # 0 auto, 1 split extend/shade kernels, 2 fused k_paths, 3 persistent
        p.reserved[2] = int(shade_min)            # persistent: finished lanes that trigger shading/refill (0 = default)
        p.reserved[3] = int(serve_min)            # persistent: waiting lanes that trigger a service (0 = default)
        p.reserved[1] = int(segments_per_launch)  # fused: ray segments per path slot per launch (0 = 32)
        if rows is not None:
            p.reserved[4], p.reserved[5] = int(rows[0]), int(rows[1])
        self.ctx.set_stream(self.torch.cuda.current_stream().cuda_stream)
        st = self.ctx.render(scene, cam, p, self.accum.data_ptr(),
                             self.accum_sq.data_ptr() if self.accum_sq is not None else None, want_stats=want_stats)
        if count_samples:
            self.spp_done += int(spp)
        self.last_stats = st
        return st

    def side_stream(self):
        """A second CUDA stream of this renderer (reduce of one row band beside the rendering of the next)."""
        if getattr(self, "_side", None) is None:
            self._side = self.torch.cuda.Stream(device=self.accum.device)
        return self._side

    def stats(self) -> _lib.PtStats:
        """Statistics of the last render() (fetched now if that call did not wait for them)."""
        if self.last_stats is None:
            self.last_stats = self.ctx.render_stats()
        return self.last_stats

    def mean(self) -> np.ndarray:
        """Linear radiance estimate, Taichi field layout [W,H,3]."""
        return self.ctx.download_accum(self.accum.data_ptr(), self.width, self.height) / max(self.spp_done, 1)

    def moments(self):
        """(sum, sum of squares) as [W,H,3] float32 arrays, for Monte Carlo standard errors."""
        s = self.ctx.download_accum(self.accum.data_ptr(), self.width, self.height)
        q = self.ctx.download_accum(self.accum_sq.data_ptr(), self.width, self.height)
        return s, q

    def image(self, aces: bool = True, gamma: float = 2.2, total_spp: int | None = None) -> np.ndarray:
        """post_processing (ACES + gamma, __main__.py:90-96) or legacy gamma_correction (aces=False)."""
        n = total_spp if total_spp is not None else max(self.spp_done, 1)
        return self.ctx.postprocess_host(self.accum.data_ptr(), self.width, self.height, 1.0 / n, aces, gamma)


def render(world, camera, spp: int = 8192, propagate_limit: int = 32, seed: int = 1, bsdf=None,
           ctx: _lib.Context | None = None, return_stats: bool = False, postprocess: bool = True,
           pixel_grid: bool = False, absorptivity: float = 0.25, aces: bool = True):
    """Drop-in for the v2 scripts' render(world, camera) + post_processing(): returns the tonemapped
    image as a float32 [W,H,3] array in Taichi field layout (pass it to imwrite)."""
    ctx = ctx or default_context()
    w, h = camera.resolution
    model = getattr(bsdf, "shading_model", _lib.PT_SHADE_V2)
    r = Renderer(w, h, ctx)
    st = r.render(world.device_scene(ctx), camera.to_struct(), spp, propagate_limit, model, seed, absorptivity=absorptivity,
                  flags=_lib.PT_FLAG_PIXEL_GRID if pixel_grid else 0)  # stages 2-4: lattice rays, no jitter
    img = r.image(aces=aces, gamma=2.2) if postprocess else r.mean()  # stages <= 5 write the linear image
    return (img, st) if return_stats else img
