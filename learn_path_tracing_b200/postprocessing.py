"""ACES + gamma of taichi_pathtracer (10_final/postprocessing.py:5-29) as host numpy helpers.

The device version (fused with the divide-by-spp that follows the NCCL reduce) is pt_postprocess;
these are for small arrays and tests.
"""
import numpy as np

_ACES_IN = np.array([[0.59719, 0.35458, 0.04823], [0.07600, 0.90834, 0.01566], [0.02840, 0.13383, 0.83777]],
                    np.float32)
_ACES_OUT = np.array([[1.60475, -0.53108, -0.07367], [-0.10208, 1.10813, -0.00605],
                      [-0.00327, -0.07276, 1.07602]], np.float32)


def ACES_tonemapping(color):
    c = np.asarray(color, np.float32)
    v = c @ _ACES_IN.T
    a = v * (v + np.float32(0.0245786)) - np.float32(0.000090537)
    b = v * (np.float32(0.983729) * v + np.float32(0.4329510)) + np.float32(0.238081)
    return np.maximum((a / b) @ _ACES_OUT.T, 0.0).astype(np.float32)


def gamma_correction(color, gamma):
    return np.power(np.asarray(color, np.float32), np.float32(1.0 / gamma))
