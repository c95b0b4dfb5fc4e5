"""ti.tools.imwrite / ti.imwrite replacement with Taichi's orientation.

Taichi fields are indexed image[i, j] = (x, y) with y UP; the PNG is clip(img,0,1)*255 -> uint8,
swapaxes(0,1)[::-1] (verified against outputs/8_refract.png).
"""
import numpy as np
from PIL import Image


def to_uint8(image) -> np.ndarray:
    """float [W,H,3] field -> uint8 [H,W,3] top-down, truncating cast like Taichi."""
    a = np.asarray(image)
    if a.dtype != np.uint8:
        a = (np.clip(a, 0.0, 1.0) * 255.0).astype(np.uint8)
    return np.ascontiguousarray(np.swapaxes(a, 0, 1)[::-1])


def imwrite(image, path):
    Image.fromarray(to_uint8(image)).save(path)


def imread(path) -> np.ndarray:
    """PNG -> float32 [W,H,3] field in [0,1] (inverse of imwrite up to quantisation)."""
    a = np.asarray(Image.open(path).convert("RGB"), np.float32) / 255.0
    return np.ascontiguousarray(np.swapaxes(a[::-1], 0, 1))
