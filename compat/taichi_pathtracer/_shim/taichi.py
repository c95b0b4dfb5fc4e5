"""`import taichi as ti` for the UNMODIFIED reference drivers (taichi_pathtracer/{6..10}_*/__main__.py):

    PYTHONPATH=compat/taichi_pathtracer/_shim python /path/to/reference/taichi_pathtracer/10_final

Only what those scripts touch: ti.init, ti.gpu/cpu, ti.func, ti.kernel, ti.template, ti.int8/f32/i32 and
ti.tools.imwrite.  See _runtime.py for what happens to the kernels."""
import numpy as _np

import _runtime

gpu, cuda, cpu = "gpu", "cuda", "cpu"
f32, i32, int8 = _np.float32, _np.int32, _np.int8


def init(arch=None, **kwargs):
    """ti.init(arch=ti.gpu): the only backend here is libb200pt.so on a B200; asking for ti.cpu does not change that
    (there is no CPU fallback) — the device is claimed when the first render is flushed."""
    return None


def template():
    return None


def func(fn):
    return _runtime.DeviceFunc(fn)


def kernel(fn):
    return _runtime.Kernel(fn)


class tools:  # noqa: N801  (ti.tools.imwrite)
    imwrite = staticmethod(_runtime.imwrite)


imwrite = _runtime.imwrite  # the older spelling
