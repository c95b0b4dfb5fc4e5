"""Stages 6-10 in the reference drivers' own idiom, run through the drop-in shim (compat/taichi_pathtracer/_shim).

The reference's scripts (taichi_pathtracer/{6..10}_*/__main__.py) run UNMODIFIED under the shim:

    PYTHONPATH=compat/taichi_pathtracer/_shim python /root/reference/taichi_pathtracer/10_final

This file is the same kind of program written against the same modules — `import taichi as ti`, `Vec3f.field`,
`Ray.field`, @ti.kernel bodies, the per-sample host loop, `ti.tools.imwrite` — for boxes where the reference checkout is
not mounted (the GPU box).  The scene comes from learn_path_tracing_b200.scenes instead of being
retyped; LPT_SPP overrides the sample count.  `python compat/taichi_pathtracer/10_final` from the repository root."""
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "_shim"))

import taichi as ti  # noqa: E402  (the shim)
from bsdf import DielectricBSDF, DiffuseBSDF, MetalBSDF  # noqa: E402,F401
from camera import Camera  # noqa: E402
from dtypes import Ray, Vec3f  # noqa: E402
from postprocessing import ACES_tonemapping, gamma_correction  # noqa: E402,F401

from learn_path_tracing_b200 import scenes  # noqa: E402

ti.init(arch=ti.gpu)

resolution = (1280, 720)
spp = int(os.environ.get("LPT_SPP", 8192))
propagate_limit = 32
image = Vec3f.field(shape=resolution)
rays = Ray.field(shape=resolution)


# The two kernel bodies below are Taichi device code in the reference (10_final/__main__.py:78-96); under the shim they
# are never executed — the shim reads which functions they reach and books the work for libb200pt.so.
@ti.kernel
def shade_glossy(world: ti.template(), rays: ti.template()):
    for i, j in rays:
        hit = world.hit(rays[i, j])
        (MetalBSDF if hit.material.metallic == 1 else DielectricBSDF).sample(rays[i, j], hit)


@ti.kernel
def shade_lambert(world: ti.template(), rays: ti.template()):
    for i, j in rays:
        DiffuseBSDF.sample(rays[i, j], world.hit(rays[i, j]))


@ti.kernel
def post_processing():
    for i, j in image:
        image[i, j] = gamma_correction(ACES_tonemapping(image[i, j]), 2.2)


def main(stage):
    world, preset = scenes.SCENES[stage](resolution)
    camera = Camera(resolution)            # the shim's Camera: get_rays(rays) stamps it into the field
    camera.set_position(preset.position)
    camera.set_direction(preset.yaw, preset.pitch, preset.roll)
    camera.set_fov(preset.fov)
    camera.set_len(preset.focal_length, preset.aperture)
    shader = shade_lambert if stage == "6_diffuse" else shade_glossy
    start = time.time()
    for _ in range(spp):                   # the reference's per-sample host loop: 2 x spp launches there, one here
        camera.get_rays(rays)
        shader(world, rays)
    post_processing()
    out = f"outputs/{stage}.png"
    ti.tools.imwrite(image, out)           # the image is first needed here: ONE pt_render call renders all booked samples
    dt = time.time() - start
    print(f"Time elapsed: {dt:.2f}s  ({resolution[0] * resolution[1] * spp / dt / 1e6:.0f} Mpaths/s end to end) -> {out}")
