"""BSDF selectors of taichi_pathtracer (10_final/bsdf.py:62-110, 6_diffuse/bsdf.py:20-26).

In the reference these classes hold @ti.func device code and are chosen inside propagate_once
(__main__.py:65-75).  Here the scattering code is hand-written CUDA (csrc/shade.cuh:scatter_v2); the classes
remain as the names driver scripts import and carry the shading-model id handed to pt_render.
"""
from . import _lib


class DiffuseBSDF:
    """l *= albedo; rd = normalize(n + uniform_on_sphere) — 6_diffuse/bsdf.py:20-26."""
    shading_model = _lib.PT_SHADE_V2_DIFFUSE


class MetalBSDF:
    """Schlick F0=albedo, slerp(mirror, lambert, roughness^2) micro-normal — bsdf.py:71-86."""
    shading_model = _lib.PT_SHADE_V2


class DielectricBSDF:
    """Schlick coin; albedo-tinted refraction / Lambert, else mirror — bsdf.py:89-110."""
    shading_model = _lib.PT_SHADE_V2


class NormalColor:
    """Stages 4-5 `ray_color`: 0.5 (normal + 1) on a hit, sky otherwise, no bounce — 5_anti_aliasing/__main__.py:19-28."""
    shading_model = _lib.PT_SHADE_V2_NORMALS


class LegacyStage7BSDF:
    """legacy/PT_in_one_weekend/7_reflect.py:187-209: cal_reflectivity_metal / _dielectirc, sample_reflect (lobe scaled by
    k = -d.n), sample_diffuse with throughput albedo * absorptivity — the untextured form of 15_module.py:281-334,994-1013."""
    shading_model = _lib.PT_SHADE_LEGACY_STAGE7


class LegacyStage6BSDF:
    """legacy/PT_in_one_weekend/6_diffuse.py:160-170: every hit scatters with sample_diffuse, l *= 0.5 * albedo."""
    shading_model = _lib.PT_SHADE_LEGACY_STAGE6
