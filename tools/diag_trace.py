"""Timing breakdown of pt_trace_batch (host buffers) vs pt_trace_batch_device on a 4 Mi ray slice."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import learn_path_tracing_b200 as L

ctx = L.default_context()
sc = L.Scene(ctx)
sc.set_random_triangles(10_000_000, 12345, 0.004)
sc.build()
n = 4 * 2**20
rays = torch.empty((2 * n, 4), dtype=torch.float32, device="cuda")
ctx.random_rays_device(rays.data_ptr(), n, 54321)
hits = torch.empty((n, 4), dtype=torch.float32, device="cuda")
for flags in (0, L.PT_FLAG_COUNTERS, L.PT_FLAG_NO_SORT, L.PT_FLAG_TRACE_SIMPLE):
    for i in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        st = ctx.trace_batch_device(sc, rays.data_ptr(), n, hits.data_ptr(), flags)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"device flags={flags}: {dt*1e3:.2f} ms wall, {st.ms_total:.2f} ms events")
rh = rays.cpu().numpy().reshape(n, 8)
for i in range(3):
    t0 = time.perf_counter(); ids, t, st = ctx.trace_batch(sc, rh, counters=(i == 0)); dt = time.perf_counter() - t0
    print(f"host trace_batch: {dt*1e3:.2f} ms wall, {st.ms_total:.2f} ms events")
