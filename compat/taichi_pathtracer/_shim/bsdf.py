"""`from bsdf import MetalBSDF, DielectricBSDF` / `DiffuseBSDF` (10_final/bsdf.py:62-110, 6_diffuse/bsdf.py:20-26)."""
import learn_path_tracing_b200 as L


def _device_only(*a, **k):
    raise RuntimeError("BSDF.sample is Taichi device code; under the B200 shim scattering runs in libb200pt.so (csrc/shade.cuh)")


class MetalBSDF(L.MetalBSDF):
    sample = staticmethod(_device_only)


class DielectricBSDF(L.DielectricBSDF):
    sample = staticmethod(_device_only)


class DiffuseBSDF(L.DiffuseBSDF):
    sample = staticmethod(_device_only)
