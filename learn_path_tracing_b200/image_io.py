"""ti.tools.imwrite / ti.imwrite replacement with Taichi's orientation.

Taichi fields are indexed image[i, j] = (x, y) with y UP; the PNG is clip(img,0,1)*255 -> uint8,
swapaxes(0,1)[::-1] (verified against outputs/8_refract.png).
"""
import numpy as np
from PIL import Image


def to_uint8(image, rounding: bool = False) -> np.ndarray:
    """float [W,H,3] field -> uint8 [H,W,3] top-down.  v2 `ti.tools.imwrite` truncates (pinned by outputs/1_save_img.png
    and stages 2-4, byte for byte); the older `ti.imwrite` of the legacy scripts rounds to nearest (pinned by
    legacy/PT_in_one_weekend/{3_adding_a_sphere,4_objects}.png, byte for byte): rounding=True."""
    a = np.asarray(image)
    if a.dtype != np.uint8:
        a = np.clip(a, 0.0, 1.0).astype(np.float32) * np.float32(255.0)
        a = (np.round(a) if rounding else a).astype(np.uint8)
    return np.ascontiguousarray(np.swapaxes(a, 0, 1)[::-1])


def imwrite(image, path, rounding: bool = False):
    Image.fromarray(to_uint8(image, rounding)).save(path)


def imwrite_legacy(image, path):
    """`ti.imwrite(frame, path)` of the legacy scripts (15_module.py:1076): same orientation, rounding cast."""
    imwrite(image, path, rounding=True)


def imread(path) -> np.ndarray:
    """PNG -> float32 [W,H,3] field in [0,1] (inverse of imwrite up to quantisation)."""
    a = np.asarray(Image.open(path).convert("RGB"), np.float32) / 255.0
    return np.ascontiguousarray(np.swapaxes(a[::-1], 0, 1))
