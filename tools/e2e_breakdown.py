import time, torch, numpy as np
import learn_path_tracing_b200 as L
from learn_path_tracing_b200 import scenes
W,H,SPP,D=1920,1080,256,50
world, cam = scenes.scene_8_refract((W,H))
ctx = L.default_context()
def T(label, fn, n=3):
    ts=[]
    for _ in range(n):
        torch.cuda.synchronize(); t0=time.perf_counter(); r=fn(); torch.cuda.synchronize(); ts.append((time.perf_counter()-t0)*1e3)
    print(f"{label:40s} {min(ts):8.2f} ms (min of {n}) {np.mean(ts):8.2f} mean"); return r
def build():
    world._scene=None; return world.device_scene(ctx)
sc=T("scene upload+build", build)
r=T("Renderer() alloc", lambda: L.Renderer(W,H,ctx))
T("render", lambda: (r.clear(), r.render(sc, cam.to_struct(), SPP, D))[1])
T("image() postprocess+D2H", lambda: r.image())
T("full L.render", lambda: L.render(world, cam, spp=SPP, propagate_limit=D, ctx=ctx))
print("bench-style e2e loop:")
for i in range(5):
    world._scene = None
    torch.cuda.synchronize(); t0=time.perf_counter()
    img = L.render(world, cam, spp=SPP, propagate_limit=D, seed=1, ctx=ctx)
    print(f"  iter {i}: {(time.perf_counter()-t0)*1e3:.2f} ms")
import threading
print("with a busy python thread (like the NVML sampler):")
stop=False
def spin():
    while not stop: time.sleep(0.1)
th=threading.Thread(target=spin,daemon=True); th.start()
for i in range(3):
    torch.cuda.synchronize(); t0=time.perf_counter()
    img = L.render(world, cam, spp=SPP, propagate_limit=D, seed=1, ctx=ctx)
    print(f"  iter {i}: {(time.perf_counter()-t0)*1e3:.2f} ms")
stop=True
