import sys, time, os
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import learn_path_tracing_b200 as L
import bench
from learn_path_tracing_b200 import scenes
W,H,SPP,D=1920,1080,256,50
world, cam = scenes.scene_8_refract((W,H))
ctx = L.default_context()
scene = world.device_scene(ctx); cs = cam.to_struct()
r = L.Renderer(W,H,ctx)
flush = torch.empty(int(2*126e6)//4, dtype=torch.float32, device="cuda")
use_sampler = len(sys.argv)>1 and sys.argv[1]=="sampler"
use_timing = len(sys.argv)>2 and sys.argv[2]=="timing"
for _ in range(3):
    r.clear(); r.render(scene, cs, SPP, D)
if use_sampler:
    s = bench.ClockSampler(0); s.start()
for _ in range(3):
    flush.fill_(1.0); torch.cuda.synchronize()
    r.clear(); r.render(scene, cs, SPP, D, flags=L.PT_FLAG_TIMING if use_timing else 0)
    torch.cuda.synchronize()
if use_sampler: print(s.result())
def t(fn):
    torch.cuda.synchronize(); t0=time.perf_counter(); x=fn(); torch.cuda.synchronize(); return x,(time.perf_counter()-t0)*1e3
for i in range(5):
    world._scene=None
    sc,a = t(lambda: world.device_scene(ctx))
    rr,b = t(lambda: L.Renderer(W,H,ctx))
    _,c = t(lambda: rr.render(sc, cs, SPP, D))
    img,d = t(lambda: rr.image())
    print(f"iter {i}: build {a:.2f} alloc {b:.2f} render {c:.2f} image {d:.2f}")
