"""`from dtypes import Vec3f, Ray, Material` (taichi_pathtracer/10_final/dtypes.py:4-9) over learn_path_tracing_b200."""
import learn_path_tracing_b200 as L
from learn_path_tracing_b200.dtypes import HitRecord, Mat3f, Material, Vec2f  # noqa: F401

import _runtime


class Vec3f(L.Vec3f):
    @classmethod
    def field(cls, shape):
        """Vec3f.field(shape=resolution): the accumulation image."""
        return _runtime.ImageField(shape)


class Ray(L.Ray):
    @classmethod
    def field(cls, shape):
        """Ray.field(shape=resolution): filled by Camera.get_rays."""
        return _runtime.RayField(shape)
