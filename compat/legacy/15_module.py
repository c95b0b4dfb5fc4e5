"""Drop-in for the script body of legacy/PT_in_one_weekend/15_module.py:1048-1076: OBJ -> World -> build -> save the
.world.npy scene cache -> progressive render(moved=False) passes -> 15_module.png.

    python compat/legacy/15_module.py [path/to/model.obj] [out.world.npy]

The reference's host-Python SAH build (minutes, 15_module.py:716-754) is replaced by the GPU LBVH, exported in the same
MeshBVHTree schema, so the saved file loads in the reference as well.  The EXR environment of the script
(cayley_interior_2k.exr) is not in the checkout; sky.png stands in when it is missing.
Knobs: LPT_RES=WxH (default 3000x2000), LPT_SPP samples per pass (8), LPT_PASSES (256 = the script's 32*8).
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

from learn_path_tracing_b200 import imwrite_legacy as imwrite  # noqa: E402  (the legacy ti.imwrite rounds)
from learn_path_tracing_b200.legacy import (Camera, LegacyRenderer, Vec3f, World, load_obj, resolve_asset,  # noqa: E402
                                            rotate)

resolution = tuple(int(v) for v in os.environ.get("LPT_RES", "3000x2000").split("x"))
spp = int(os.environ.get("LPT_SPP", 8))          # samples per render() pass (the script's `batch`)
passes = int(os.environ.get("LPT_PASSES", 32 * 8))
propagate_limit = 32

obj = sys.argv[1] if len(sys.argv) > 1 else "./models/Yoimiya/Yoimiya_ShapeChange.obj"
out = sys.argv[2] if len(sys.argv) > 2 else "Yoimiya.world.npy"

world = World()
env = "./textures/cayley_interior_2k.exr"
world.environments.add(env if resolve_asset(env) else "./textures/sky.png", 0)
positions, normals, texture_coords, indices, materials = load_obj(obj, 1, flip_z=True, flip_textcoord=True,
                                                                  transform=rotate(np.pi, 0))
for material in materials:
    world.textures.add(material["file_path"], material["id"])
world.add_mesh(positions, normals, texture_coords, indices)
world.set_environment(0)
t0 = time.time()
world.build()
world.save(out)
print(f"build + save {out}: {time.time() - t0:.2f}s ({len(indices)} faces)")

camera = Camera(resolution)
camera.set_fov(30)
camera.set_position(Vec3f([0, 8, -30]))
camera.look_at(Vec3f([0, 8, 0]))

renderer = LegacyRenderer(world, camera, spp=spp, propagate_limit=propagate_limit)
t0 = time.time()
for i in range(passes):
    frame = renderer.render(moved=False)
print(f"{renderer.total_spp} spp in {time.time() - t0:.2f}s")
imwrite(frame, "15_module.png")
