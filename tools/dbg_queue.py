"""Debug: persistent vs queue kernel on small renders."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import learn_path_tracing_b200 as L
from learn_path_tracing_b200 import scenes
ctx = L.default_context()
for name, (W, H), spp in (("8_refract", (80, 45), 24), ("10_final", (80, 45), 24), ("10_final", (320, 180), 64)):
    world, cam = scenes.SCENES[name]((W, H))
    sc = world.device_scene(ctx)
    out = {}
    for mode in (L.PT_MODE_PERSIST, L.PT_MODE_QUEUE):
        r = L.Renderer(W, H, ctx)
        st = r.render(sc, cam.to_struct(), spp, 32, seed=5, mode=mode)
        acc = r.accum.cpu().numpy()
        out[mode] = (r.mean(), st, acc)
        print(name, W, H, spp, "mode", mode, "paths", st.paths, "segments", st.segments, "mean", float(r.mean().mean()),
              "counted paths", float(acc[:, 3].sum()), "ms", st.ms_total, "dbg", st.reserved[3], flush=True)
    a, b = out[L.PT_MODE_PERSIST][0], out[L.PT_MODE_QUEUE][0]
    print("  max abs diff", float(np.abs(a - b).max()), "rel mean", float(b.mean() / a.mean() - 1))
