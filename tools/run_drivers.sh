#!/bin/bash
# Runs every drop-in driver once at reduced settings (GPU box): compat/taichi_pathtracer/{2..10}_* and
# compat/legacy/15_module.py (OBJ -> GPU LBVH -> .world.npy -> progressive render).  The legacy driver needs the model
# and sky.png under $LPT_ASSETS (copy assets/models/Yoimiya and assets/textures/sky.png of the reference checkout into
# ./_tmp_assets before calling gpurun; the directory is git-ignored).  Output: gpurun_out/drivers/.
set -e
export LPT_ASSETS=${LPT_ASSETS:-$PWD/_tmp_assets}
mkdir -p gpurun_out/drivers && cd gpurun_out/drivers
for s in 2_camera_and_ray 3_adding_a_sphere 4_objects 5_anti_aliasing 6_diffuse 7_reflect 8_refract 9_dof 10_final; do
  LPT_SPP=64 python ../../compat/taichi_pathtracer/$s 2>&1 | tail -1
done
LPT_RES=600x400 LPT_SPP=8 LPT_PASSES=4 python ../../compat/legacy/15_module.py 2>&1 | tail -3
python - <<'PY'
import numpy as np, sys
sys.path.insert(0, "../..")
from PIL import Image
from learn_path_tracing_b200 import worldnpy
for f in ["outputs/2_camera_and_ray.png","outputs/8_refract.png","outputs/10_final.png","15_module.png"]:
    a=np.asarray(Image.open(f)); print(f, a.shape, "mean", a.mean().round(2))
w = worldnpy.load_world("Yoimiya.world.npy")
print("saved world keys:", sorted(w.keys()) if isinstance(w, dict) else type(w))
PY
