"""Host-side value types of the reference's scene-description surface.

Mirrors taichi_pathtracer/10_final/dtypes.py:4-9 (Vec2f, Vec3f, Mat3f, Ray, Material, HitRecord) and
world.py:37-41 (Sphere) as plain numpy-backed Python objects: in the reference these are Taichi
struct/vector types; here they only describe the scene, the device layout lives in csrc/.
"""
from __future__ import annotations

import math

import numpy as np


class _Vec(np.ndarray):
    _n = 3

    def __new__(cls, value=0.0, *rest):
        if rest:  # Vec3f(x, y, z) as legacy/15_module.py:879 uses
            value = (value,) + rest
        a = np.empty(cls._n, np.float32).view(cls)
        a[...] = np.asarray(value, np.float32)  # scalar broadcasts like ti Vector(0)
        return a

    def __array_finalize__(self, obj):
        pass

    def norm(self) -> float:
        return float(math.sqrt(float(np.dot(self, self))))

    def normalized(self):
        return (self / np.float32(self.norm())).view(type(self))

    def dot(self, other):  # ti Vector.dot
        return float(np.dot(np.asarray(self), np.asarray(other)))

    def cross(self, other):
        return np.cross(np.asarray(self), np.asarray(other)).astype(np.float32).view(type(self))


class Vec2f(_Vec):
    _n = 2


class Vec3f(_Vec):
    _n = 3


class Vec2i(np.ndarray):
    def __new__(cls, value=0):
        a = np.empty(2, np.int32).view(cls)
        a[...] = np.asarray(value, np.int32)
        return a


def Mat3f(rows):
    return np.asarray(rows, np.float32).reshape(3, 3)


class Material:
    """Material(albedo, roughness, metallic, ior, transparency=0) — dtypes.py:8."""

    __slots__ = ("albedo", "roughness", "metallic", "ior", "transparency")

    def __init__(self, albedo=(0.0, 0.0, 0.0), roughness=0.0, metallic=0, ior=0.0, transparency=0):
        self.albedo = Vec3f(albedo)
        self.roughness = float(roughness)
        self.metallic = int(metallic)
        self.ior = float(ior)
        self.transparency = int(transparency)

    def __repr__(self):
        return (f"Material(albedo={list(map(float, self.albedo))}, roughness={self.roughness}, "
                f"metallic={self.metallic}, ior={self.ior}, transparency={self.transparency})")


class Sphere:
    """Sphere(center, radius, material) — world.py:37-41.

    Stage 6 passes a bare albedo vector as the third argument (6_diffuse/world.py:33-36); that is
    accepted and stored as a Lambertian material.
    """

    __slots__ = ("center", "radius", "material")

    def __init__(self, center, radius, material=None, albedo=None):
        self.center = Vec3f(center)
        self.radius = float(radius)
        if albedo is not None:
            material = albedo
        if material is None:
            material = Material()
        if not isinstance(material, Material):  # stage-6 form: Sphere(center, radius, albedo)
            material = Material(albedo=material, roughness=1.0, metallic=0, ior=1.5, transparency=0)
        self.material = material


class Ray:
    """Ray(ro, rd, l, end) — dtypes.py:7.  Host-side record used by tests and debugging only."""

    __slots__ = ("ro", "rd", "l", "end")

    def __init__(self, ro=(0, 0, 0), rd=(0, 0, -1), l=(1, 1, 1), end=0):
        self.ro, self.rd, self.l, self.end = Vec3f(ro), Vec3f(rd), Vec3f(l), int(end)


class HitRecord:
    """HitRecord(point, normal, t, material) — dtypes.py:9."""

    __slots__ = ("point", "normal", "t", "material")

    def __init__(self, point=(0, 0, 0), normal=(0, 0, 0), t=-1.0, material=None):
        self.point, self.normal, self.t = Vec3f(point), Vec3f(normal), float(t)
        self.material = material or Material()
