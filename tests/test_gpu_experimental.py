"""Kernel forms that were built, measured and lost to the default kernels (`make EXPERIMENTAL=1` -> libb200pt_exp.so):
the 4-wide tree walk (PT_FLAG_TRACE_WIDE / PT_FLAG_WIDE).  The queue / dual render modes are covered by the mode lists of
test_gpu_parity.py when this library is loaded.  Collected only when the loaded library has them (tests/conftest.py):

    PT_LIB_PATH=$PWD/learn_path_tracing_b200/libb200pt_exp.so python -m pytest tests -m gpu -q

Round 2 on the B200 (profiles/r02_ab_wide.txt): all three tests green, the wide walk 5-9 % slower than the binary one on
the render workloads and 40 % slower on the 10 M-triangle batch."""
import numpy as np
import pytest

import learn_path_tracing_b200 as L
from learn_path_tracing_b200 import scenes

pytestmark = pytest.mark.gpu


def test_wide_tree_traversal_returns_identical_hit_records(ctx, monkeypatch):
    """EXPERIMENTAL k_trace_persist<.., WIDE> over the 4-wide copy of the tree (csrc/bvh4.h, PT_WIDE=1 at build time):
    the closest hit does not depend on the tree, so the records (t, prim, u, v) must equal the binary walk's bit for bit;
    the step count must fall (CPU prototype: 0.5-0.65x; measured on B200: 40.5 instead of 71.4 steps per ray, the
    counting kernels 0.62 instead of 0.82 ms for this batch)."""
    import torch
    monkeypatch.setenv("PT_WIDE", "1")
    n_tri, n_rays = 200_000, 400_000
    sc = L.Scene(ctx)
    sc.set_random_triangles(n_tri, 777, 0.02)
    sc.build()
    rays = torch.empty((2 * n_rays, 4), dtype=torch.float32, device="cuda")
    ctx.random_rays_device(rays.data_ptr(), n_rays, 999)
    a = torch.empty((n_rays, 4), dtype=torch.float32, device="cuda")
    b = torch.empty((n_rays, 4), dtype=torch.float32, device="cuda")
    sa = ctx.trace_batch_device(sc, rays.data_ptr(), n_rays, a.data_ptr(), L.PT_FLAG_COUNTERS)
    sb = ctx.trace_batch_device(sc, rays.data_ptr(), n_rays, b.data_ptr(), L.PT_FLAG_COUNTERS | L.PT_FLAG_TRACE_WIDE)
    torch.cuda.synchronize()
    assert torch.equal(a.view(torch.int32), b.view(torch.int32))
    assert (a[:, 1].view(torch.int32) >= 0).float().mean() > 0.2
    assert sb.prims_tested <= 1.1 * sa.prims_tested and sb.nodes_visited < 0.8 * sa.nodes_visited, (sa.nodes_visited, sb.nodes_visited)
    c = torch.empty((n_rays, 4), dtype=torch.float32, device="cuda")
    ctx.trace_batch_device(sc, rays.data_ptr(), n_rays, c.data_ptr(), L.PT_FLAG_TRACE_WIDE | L.PT_FLAG_NO_SORT)
    torch.cuda.synchronize()
    assert torch.equal(a.view(torch.int32), c.view(torch.int32))
    print(f"wide: {sb.nodes_visited / n_rays:.1f} steps/ray vs {sa.nodes_visited / n_rays:.1f}; {sb.ms_extend:.2f} ms vs {sa.ms_extend:.2f} ms (counting kernels)")
    monkeypatch.delenv("PT_WIDE")
    sc2 = L.Scene(ctx)
    sc2.set_random_triangles(1000, 1, 0.05)
    sc2.build()
    with pytest.raises(L.PtError):   # no wide tree in a scene built without PT_WIDE=1
        ctx.trace_batch_device(sc2, rays.data_ptr(), 1000, c.data_ptr(), L.PT_FLAG_TRACE_WIDE)


@pytest.mark.parametrize("name", ["10_final", "legacy_synthetic"])
def test_wide_tree_render_traces_the_same_paths(ctx, monkeypatch, name):
    """k_paths_persist<.., WIDE>: same RNG keys, same closest hits -> the same paths as the binary-tree kernel."""
    monkeypatch.setenv("PT_WIDE", "1")
    if name == "10_final":
        W, H = 160, 90
        world, cam = scenes.scene_10_final((W, H))
        model, kw = L.PT_SHADE_V2, {}
    else:
        from helpers import synthetic_legacy_world
        world, cam = synthetic_legacy_world()
        W, H = cam.resolution
        model, kw = L.PT_SHADE_LEGACY, {"absorptivity": 0.25}
    sc = world.device_scene(ctx)
    a, b = L.Renderer(W, H, ctx), L.Renderer(W, H, ctx)
    sa = a.render(sc, cam.to_struct(), 32, 32, model, seed=5, flags=L.PT_FLAG_COUNTERS, **kw)
    sb = b.render(sc, cam.to_struct(), 32, 32, model, seed=5, flags=L.PT_FLAG_COUNTERS | L.PT_FLAG_WIDE, **kw)
    assert sa.paths == sb.paths and abs(int(sa.segments) - int(sb.segments)) <= 1e-3 * sa.segments
    assert np.array_equal(a.accum.cpu().numpy()[:, 3], b.accum.cpu().numpy()[:, 3])
    assert np.allclose(a.mean(), b.mean(), rtol=2e-3, atol=2e-4)
    assert sb.nodes_visited < 0.85 * sa.nodes_visited, (sa.nodes_visited, sb.nodes_visited)
    print(f"{name}: wide {sb.nodes_visited / sb.segments:.2f} steps/segment vs {sa.nodes_visited / sa.segments:.2f}")
