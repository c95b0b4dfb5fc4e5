"""World / Sphere of taichi_pathtracer stages 6-10 (10_final/world.py:5-60).

World keeps the reference's list semantics (construction from a list, add(), size, insertion order
decides ties).  World.hit / Sphere.hit run on the device: csrc/extend.cuh.
"""
from __future__ import annotations

import numpy as np

from . import _lib
from .dtypes import Sphere  # noqa: F401  (re-exported: `from world import World, Sphere`)


class World:
    def __init__(self, spheres=()):
        self.spheres = list(spheres)
        self._scene = None
        self._scene_ctx = None

    @property
    def size(self):
        return len(self.spheres)

    def add(self, sphere):
        self.spheres.append(sphere)
        self._scene = None

    # ---- device interface ------------------------------------------------------------------
    def arrays(self):
        """(center_radius float32 [n,4], materials MATERIAL_DTYPE [n]) in insertion order."""
        n = len(self.spheres)
        cr = np.zeros((n, 4), np.float32)
        mats = np.zeros(n, _lib.MATERIAL_DTYPE)
        for i, s in enumerate(self.spheres):
            cr[i, :3] = s.center
            cr[i, 3] = s.radius
            m = s.material
            mats[i] = (tuple(float(x) for x in m.albedo), m.roughness, m.metallic, m.ior, m.transparency, 0)
        return cr, mats

    def device_scene(self, ctx):
        if self._scene is None or self._scene_ctx is not ctx:
            sc = _lib.Scene(ctx)
            cr, mats = self.arrays()
            sc.set_spheres(cr, mats)
            sc.build()
            self._scene, self._scene_ctx = sc, ctx
        return self._scene

    def hit(self, rays, ctx=None):
        """World.hit over a ray batch [n,8] -> (prim_id int32 [n], t float32 [n])."""
        from .render import default_context
        ctx = ctx or default_context()
        ids, t, _ = ctx.trace_batch(self.device_scene(ctx), rays)
        return ids, t
