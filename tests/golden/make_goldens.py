"""Regenerates tests/golden/*_320x180.png from the reference's committed renders.

Run in the build container (where /root/reference is mounted):
    python tests/golden/make_goldens.py [/root/reference]
Each golden is the reference's outputs/<stage>.png (1280x720, 8192 spp, ACES + gamma, 8 bit; stage 5:
100 spp, normals as colours, linear) box-
filtered 4x4 down to 320x180 so the CPU oracle can be checked against it in seconds.  The PNGs are
the reference's own output data (not source code); they pin camera, Sphere.hit incl. the far-root
rule, the v2 BSDFs, the sky, post-processing and imwrite orientation.
"""
import os
import sys

import numpy as np
from PIL import Image

STAGES = ["5_anti_aliasing", "6_diffuse", "7_reflect", "8_refract", "9_dof"]
# stages 2-4 are deterministic (one lattice ray per pixel, no random numbers): kept at their native size as exact pins
EXACT = ["1_save_img", "2_camera_and_ray", "3_adding_a_sphere", "4_objects"]   # stage 1 pins imwrite alone


def main(ref="/root/reference"):
    here = os.path.dirname(os.path.abspath(__file__))
    for s in STAGES:
        a = np.asarray(Image.open(os.path.join(ref, "outputs", s + ".png")).convert("RGB"), np.float64)
        h, w, _ = a.shape
        if (w, h) == (320, 180):  # the early stages were committed at 320x180: kept as they are
            b = a
        else:
            assert (w, h) == (1280, 720), (s, w, h)
            b = a.reshape(h // 4, 4, w // 4, 4, 3).mean(axis=(1, 3))
        out = os.path.join(here, f"{s}_320x180.png")
        Image.fromarray(np.round(b).astype(np.uint8)).save(out, optimize=True)
        print(out, os.path.getsize(out), "bytes")
    # the legacy tutorial stages with the legacy camera (half-angle fov, 15_module.py:397-401 has the same formula):
    # 3 and 4 are deterministic lattice renders, 5 is 100 spp of normals as colours
    # 6_diffuse.png is the converged (8192 spp) render of legacy 6_diffuse.py as committed: pins the legacy scattering
    # helpers shared with 15_module.py (sample_at_sphere, sample_diffuse), the t > 1e-3 hit rule and the 0.5 * albedo
    # throughput.  7_reflect.png was rendered with OTHER scene/camera settings than the committed 7_reflect.py (the horizon
    # lies 24.5 rows below the centre; the script's radius-10000 ground under a level camera puts it at the centre): only
    # its sky rows (camera with fov 45 = half angle, gamma 2.2, rounding cast) are usable as a pin.
    for s in ["3_adding_a_sphere", "4_objects", "5_anti_aliasing", "6_diffuse", "7_reflect"]:
        im = Image.open(os.path.join(ref, "legacy", "PT_in_one_weekend", s + ".png")).convert("RGB")
        out = os.path.join(here, f"legacy_{s}_{im.size[0]}x{im.size[1]}.png")
        im.save(out, optimize=True)
        print(out, os.path.getsize(out), "bytes")
    for s in EXACT:
        im = Image.open(os.path.join(ref, "outputs", s + ".png")).convert("RGB")
        out = os.path.join(here, f"{s}_{im.size[0]}x{im.size[1]}.png")
        im.save(out, optimize=True)
        print(out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main(*sys.argv[1:])
