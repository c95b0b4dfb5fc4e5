"""Camera of taichi_pathtracer stages 6-10 (10_final/camera.py:38-93).

Same constructor, setters and angle conventions (degrees, yaw*pitch*roll).  The per-pixel work of
Camera.get_rays runs on the device (csrc/shade.cuh:camera_ray, fused into the path kernel); the host only derives the basis.
"""
from __future__ import annotations

import math

import numpy as np

from . import _lib
from .dtypes import Vec3f


def rotate(yaw, pitch, roll=0.0):
    """camera.py:6-26 — yaw/pitch/roll in DEGREES -> 3x3 (yaw @ pitch @ roll)."""
    y, p, r = math.radians(yaw), math.radians(pitch), math.radians(roll)
    yaw_t = np.array([[math.cos(y), 0, math.sin(y)], [0, 1, 0], [-math.sin(y), 0, math.cos(y)]])
    pitch_t = np.array([[1, 0, 0], [0, math.cos(p), -math.sin(p)], [0, math.sin(p), math.cos(p)]])
    roll_t = np.array([[math.cos(r), -math.sin(r), 0], [math.sin(r), math.cos(r), 0], [0, 0, 1]])
    return yaw_t @ pitch_t @ roll_t


class Camera:
    def __init__(self, resolution, fov=60, focal_length=1, aperture=0):
        self.resolution = (int(resolution[0]), int(resolution[1]))
        self.fov = float(fov)
        self.focal_length = float(focal_length)
        self.aperture = float(aperture)
        self.position = Vec3f(0)
        self.yaw = 0.0
        self.pitch = 0.0
        self.roll = 0.0

    def set_position(self, position):
        self.position = Vec3f(position)

    def set_direction(self, yaw, pitch, roll=0):
        self.yaw, self.pitch, self.roll = float(yaw), float(pitch), float(roll)

    def set_fov(self, fov):
        self.fov = fov

    def set_len(self, focal_length=1, aperture=0):
        self.focal_length = float(focal_length)
        self.aperture = float(aperture)

    def look_at(self, target, roll=0):
        d = (Vec3f(target) - self.position).normalized()  # camera.py:65-69
        self.yaw = math.degrees(math.atan2(-d[0], -d[2]))
        self.pitch = math.degrees(math.asin(d[1]))
        self.roll = float(roll)

    # ---- device interface ------------------------------------------------------------------
    def to_struct(self) -> _lib.PtCamera:
        """Basis + view extents (camera.py:79-85): view_width = 2*tan(radians(fov)/2)."""
        w, h = self.resolution
        trans = rotate(self.yaw, self.pitch, self.roll)
        view_w = 2.0 * math.tan(math.radians(self.fov) / 2.0)
        view_h = view_w * (h / w)
        c = _lib.PtCamera()
        c.pos[:] = [float(x) for x in self.position]
        c.front[:] = [float(x) for x in trans @ np.array([0.0, 0.0, -1.0])]
        c.right[:] = [float(x) for x in trans @ np.array([1.0, 0.0, 0.0])]
        c.up[:] = [float(x) for x in trans @ np.array([0.0, 1.0, 0.0])]
        c.view_w, c.view_h = view_w, view_h
        c.focal_length, c.aperture = self.focal_length, self.aperture
        return c

    def get_rays(self, sample=0, seed=1, ctx=None):
        """One camera ray per pixel (Camera.get_rays): returns float32 [H*W, 8] = o, tmin, d, tmax."""
        from .render import default_context
        ctx = ctx or default_context()
        w, h = self.resolution
        return ctx.generate_rays(self.to_struct(), w, h, int(sample), int(seed))
