"""Drop-in for legacy/PT_in_one_weekend/14_mesh.py:985-1023: load a .world.npy scene cache, render, write 14_mesh.png.

    python compat/legacy/14_mesh.py [path/to/Yoimiya_ShapeChange.world.npy]

Module constants follow the reference script (resolution, spp, batch-free, propagate_limit, absorptivity 0.5).
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

from learn_path_tracing_b200 import imwrite_legacy as imwrite  # noqa: E402  (the legacy ti.imwrite rounds)
from learn_path_tracing_b200.legacy import Camera, LegacyRenderer, TextureManager, Vec3f, World  # noqa: E402

resolution = (3000, 2000)
spp = int(os.environ.get("LPT_SPP", 8192))
propagate_limit = 4          # 14_mesh.py:43

world = World()
world.load(sys.argv[1] if len(sys.argv) > 1 else "/root/reference/legacy/Yoimiya_ShapeChange.world.npy")
if world._env is None:       # old-format caches record no environment: the script supplied one (14_mesh.py:987-989)
    world.environments = TextureManager((256, 256))
    world.environments.add("./textures/sky.png", 0)
    world.environments.build()
    world.set_environment(0)
    world.load_textures()

camera = Camera(resolution)
camera.set_fov(30)
camera.set_position(Vec3f([0, 8, -30]))
camera.look_at(Vec3f([0, 8, 0]))

start_time = time.time()
frame = LegacyRenderer(world, camera, spp=spp, propagate_limit=propagate_limit, absorptivity=0.5).render()
print(f"Time elapsed: {time.time() - start_time:.2f}s")
imwrite(frame, "14_mesh.png")
