"""Shared body of the drop-in v2 drivers: scene from learn_path_tracing_b200.scenes, render, write the PNG where the
reference script writes it (./outputs/<stage>.png relative to the cwd)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

from learn_path_tracing_b200 import DielectricBSDF, DiffuseBSDF, NormalColor, imwrite, render, scenes  # noqa: E402


def main(stage, resolution=(1280, 720), spp=8192, propagate_limit=32):
    spp = int(os.environ.get("LPT_SPP", spp))
    world, camera = scenes.SCENES[stage](resolution)
    start_time = time.time()
    early = stage in ("2_camera_and_ray", "3_adding_a_sphere", "4_objects")  # one lattice ray per pixel, normals / sky
    bsdf = {"5_anti_aliasing": NormalColor, "6_diffuse": DiffuseBSDF}.get(stage, NormalColor if early else DielectricBSDF)
    # stages <= 5 write the linear image (no post_processing() in those scripts)
    image, stats = render(world, camera, spp=1 if early else spp, propagate_limit=propagate_limit, bsdf=bsdf, return_stats=True,
                          postprocess=not early and stage != "5_anti_aliasing", pixel_grid=early)
    print(f"Time elapsed: {time.time() - start_time:.2f}s  ({stats.paths / stats.ms_total / 1e3:.0f} Mpaths/s)")
    os.makedirs("outputs", exist_ok=True)
    imwrite(image, f"outputs/{stage}.png")
