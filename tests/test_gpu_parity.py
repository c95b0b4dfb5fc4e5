"""GPU parity tests (run on a B200 with `pytest -m gpu`): the CUDA path, called through the C-ABI,
against the CPU oracle on identical seeded inputs."""
import os

import numpy as np
import pytest

import learn_path_tracing_b200 as L
from learn_path_tracing_b200 import scenes

pytestmark = pytest.mark.gpu

# kernel forms of `make EXPERIMENTAL=1` builds (PT_LIB_PATH=.../libb200pt_exp.so): in the lists only when present
EXP = L._lib.has_experimental()
EXP_MODES = [L.PT_MODE_QUEUE, L.PT_MODE_DUAL] if EXP else []
EXP_CASES = [(L.PT_MODE_QUEUE, 0, 0, 0), (L.PT_MODE_QUEUE, 0, 0, 1), (L.PT_MODE_QUEUE, 0, 0, 32), (L.PT_MODE_DUAL, 0, 0, 0),
             (L.PT_MODE_DUAL, 0, 3, 0), (L.PT_MODE_DUAL, 0, 0, 1), (L.PT_MODE_DUAL, 0, 3, 32)] if EXP else []


def _z_scores(s, q, osum, osq, n):
    mu_g, mu_o = s / n, osum / n
    var_g = np.maximum(q / n - mu_g**2, 0) / n
    var_o = np.maximum(osq / n - mu_o**2, 0) / n
    return np.abs(mu_g - mu_o) / np.sqrt(var_g + var_o + 1e-12), mu_g, mu_o


# ---- camera -------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["8_refract", "9_dof", "10_final"])
def test_camera_rays_match_oracle(ctx, oracle, name):
    """Camera.get_rays (camera.py:71-93): same counter-based uniforms -> same rays to float rounding."""
    W, H = 160, 90
    _, cam = scenes.SCENES[name]((W, H))
    for sample in (0, 5):
        g = ctx.generate_rays(cam.to_struct(), W, H, sample, 11)
        o = oracle.generate_rays(cam.to_struct(), W, H, sample, 11)
        assert np.allclose(g[:, :3], o[:, :3], atol=2e-6)
        assert np.allclose(g[:, 4:7], o[:, 4:7], atol=2e-6)
        assert np.allclose(np.linalg.norm(g[:, 4:7], axis=1), 1.0, atol=1e-6)


# ---- fixed ray batch: hit ids bit-exact, t bit-exact (tolerance 1e-5 rel in the contract) ---------
def _secondary_rays(oracle, world, cam, W, H, n_bounce_seeds=2):
    """camera rays + rays leaving the first hit points in random directions (origin ON a surface)."""
    cr, mats = world.arrays()
    rays = oracle.generate_rays(cam.to_struct(), W, H, 0, 5)
    ids, t = oracle.trace_spheres(cr, mats, rays)
    hit = ids >= 0
    rng = np.random.default_rng(123)
    out = [rays]
    for _ in range(n_bounce_seeds):
        o = rays[hit, :3] + t[hit, None] * rays[hit, 4:7]
        d = rng.normal(size=o.shape).astype(np.float32)
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        sec = np.zeros((o.shape[0], 8), np.float32)
        sec[:, :3], sec[:, 3], sec[:, 4:7], sec[:, 7] = o, 1e-4, d, np.inf
        out.append(sec)
    return np.concatenate(out).astype(np.float32)


@pytest.mark.parametrize("name", ["6_diffuse", "8_refract", "10_final"])
def test_sphere_hits_bit_exact(ctx, oracle, name):
    W, H = 256, 144
    world, cam = scenes.SCENES[name]((W, H))
    rays = _secondary_rays(oracle, world, cam, W, H)
    cr, mats = world.arrays()
    oid, ot, ot64 = oracle.trace_spheres(cr, mats, rays, want_t64=True)
    gid, gt = world.hit(rays, ctx)
    assert np.array_equal(gid, oid)
    assert np.array_equal(gt, ot)  # bit-exact, stronger than the 1e-5 relative contract
    assert (oid >= 0).mean() > 0.3 and (oid < 0).any()
    # conditioning report vs float64 (informational bound): the reference's own f32 formula loses digits on
    # grazing rays and on the radius-10000 ground sphere; away from those it is within 1e-5 relative
    small = (oid >= 0) & (cr[np.maximum(oid, 0), 3] < 100)
    rel = np.abs(gt[small] - ot64[small]) / np.abs(ot64[small])
    assert np.median(rel) < 1e-5  # typical ray; grazing rays lose up to ~1e-1 in the reference's f32 formula itself


def test_bvh_is_used_for_the_random_scene(ctx):
    world, _ = scenes.scene_10_final((64, 36))
    sc = world.device_scene(ctx)
    n_nodes, n_prims, n_global = sc.bvh_info()
    assert n_prims == world.size and n_global == 1 and n_nodes == n_prims - n_global - 1
    nodes, glob = sc.bvh_download()
    assert list(glob) == [0]  # the radius-10000 ground sphere stays out of the Morton grid
    kids = nodes[:, 12:14].view(np.int32)
    leaves = np.sort(~kids[kids < 0])
    assert np.array_equal(leaves, np.arange(1, world.size))  # every other sphere is exactly one leaf
    inner = np.sort(kids[kids >= 0])
    assert np.array_equal(inner, np.arange(1, n_nodes))  # every inner node but the root has one parent


def test_trace_batch_edge_cases(ctx, oracle):
    world, cam = scenes.scene_8_refract((16, 9))
    ids, t = world.hit(np.zeros((0, 8), np.float32), ctx)  # empty batch
    assert ids.shape == (0,) and t.shape == (0,)
    # a ray that misses everything, a ray starting inside a glass sphere (far root), tmax clipping
    rays = np.array([[0, 50, 0, 1e-4, 0, 1, 0, np.inf],
                     [-0.5, 0.866, 0, 1e-4, 0, 0, 1, np.inf],
                     [0, 0.4, 4, 1e-4, 0, 0, -1, 1.0]], np.float32)
    ids, t = world.hit(rays, ctx)
    cr, mats = world.arrays()
    oid, ot = oracle.trace_spheres(cr, mats, rays[:2])
    assert ids[0] == -1 and t[0] == -1
    assert ids[1] == oid[1] == 3 and t[1] == ot[1] and abs(t[1] - 0.5) < 1e-6
    assert ids[2] == -1  # nearest sphere is 3.5 away, beyond tmax = 1


# ---- images: 3 sigma of the Monte Carlo standard error ------------------------------------------
@pytest.mark.parametrize("mode", [L.PT_MODE_FUSED, L.PT_MODE_SPLIT, L.PT_MODE_PERSIST] + EXP_MODES[1:])
@pytest.mark.parametrize("name,model,depth", [("6_diffuse", L.PT_SHADE_V2_DIFFUSE, 32), ("7_reflect", L.PT_SHADE_V2, 32),
                                              ("8_refract", L.PT_SHADE_V2, 50), ("9_dof", L.PT_SHADE_V2, 32),
                                              ("10_final", L.PT_SHADE_V2, 32)])
def test_image_within_3_sigma_of_oracle(ctx, oracle, name, model, depth, mode):
    W, H, SPP = 160, 90, 256
    world, cam = scenes.SCENES[name]((W, H))
    r = L.Renderer(W, H, ctx, want_sq=True)
    st = r.render(world.device_scene(ctx), cam.to_struct(), SPP, depth, model, seed=2, mode=mode)
    s, q = r.moments()
    osum, osq, ost = oracle.render(oracle.scene_from_world(world), cam.to_struct(), W, H, SPP, depth, model, seed=2,
                                   want_sq=True)
    z, mu_g, mu_o = _z_scores(s, q, osum, osq, SPP)
    assert st.paths == ost.paths == W * H * SPP
    assert abs(st.segments / ost.segments - 1.0) < 0.01
    assert (z > 3).mean() < 0.01, (z > 3).mean()  # 0.27 % expected by chance
    assert z.max() < 8.0, z.max()
    # aggregate RMSE of the tonemapped 8-bit images (reported in DESIGN.md)
    a = L.to_uint8(r.image()).astype(np.float64)
    b = L.to_uint8(oracle.postprocess(osum, 1.0 / SPP)).astype(np.float64)
    assert np.sqrt(((a - b) ** 2).mean()) < 4.0
    assert abs(mu_g.mean() / mu_o.mean() - 1.0) < 2e-3


def test_stage5_normals_match_oracle_and_reference_png(ctx, oracle):
    """PT_SHADE_V2_NORMALS (stages 4-5): same paths as the oracle (sums equal to fp32 order) and the reference's own
    outputs/5_anti_aliasing.png within silhouette noise."""
    import os
    from PIL import Image
    from conftest import GOLDEN
    W, H, SPP = 320, 180, 128
    world, cam = scenes.scene_5_anti_aliasing((W, H))
    r = L.Renderer(W, H, ctx)
    st = r.render(world.device_scene(ctx), cam.to_struct(), SPP, 32, L.PT_SHADE_V2_NORMALS, seed=3)
    osum, _, ost = oracle.render(oracle.scene_from_world(world), cam.to_struct(), W, H, SPP, 32, L.PT_SHADE_V2_NORMALS, seed=3)
    assert st.paths == ost.paths and st.segments == st.paths == ost.segments
    diff = np.abs(r.mean() - osum / SPP)   # same rays; a silhouette sample may flip hit/miss within float rounding
    assert np.quantile(diff, 0.99) < 1e-4 and diff.max() < 3.0 / SPP, (np.quantile(diff, 0.99), diff.max())
    gold = np.asarray(Image.open(os.path.join(GOLDEN, "5_anti_aliasing_320x180.png")).convert("RGB"), np.float64)
    d = L.to_uint8(r.mean()).astype(np.float64) - gold
    assert np.sqrt((d**2).mean()) < 1.5 and abs(d.mean()) < 0.3
    with pytest.raises(L.PtError):  # only the persistent kernel implements this model
        L.Renderer(W, H, ctx).render(world.device_scene(ctx), cam.to_struct(), 1, 32, L.PT_SHADE_V2_NORMALS, mode=L.PT_MODE_SPLIT)


@pytest.mark.parametrize("name,W,H", [("2_camera_and_ray", 1280, 720), ("3_adding_a_sphere", 1280, 720), ("4_objects", 320, 180)])
def test_deterministic_stages_match_reference_png_and_oracle(ctx, oracle, name, W, H):
    """Stages 2-4 (PT_FLAG_PIXEL_GRID, one lattice ray per pixel, no random numbers): the GPU image equals the
    reference's own PNG to the byte (last-ulp flips at a quantisation threshold aside) and the oracle to float rounding."""
    import os
    from PIL import Image
    from conftest import GOLDEN
    world, cam = scenes.SCENES[name]((W, H))
    img, st = L.render(world, cam, spp=1, bsdf=L.NormalColor, ctx=ctx, return_stats=True, postprocess=False, pixel_grid=True)
    assert st.paths == st.segments == W * H
    osum, _, _ = oracle.render(oracle.scene_from_world(world), cam.to_struct(), W, H, 1, 32, L.PT_SHADE_V2_NORMALS, seed=1,
                               flags=L.PT_FLAG_PIXEL_GRID)
    # same rays up to the last ulp (the kernel multiplies by 1/(W-1), normalises with rsqrt): colours agree to float
    # rounding, except next to the silhouette where t is ill-conditioned (and a lattice ray may flip hit/miss)
    diff = np.abs(img - osum)
    assert np.quantile(diff, 0.99) < 1e-5 and np.median(diff) < 1e-6, (np.quantile(diff, 0.99), np.median(diff), diff.max())
    gold = np.asarray(Image.open(os.path.join(GOLDEN, f"{name}_{W}x{H}.png")).convert("RGB"), np.int32)
    d = np.abs(L.to_uint8(img).astype(np.int32) - gold)
    assert (d == 0).mean() > 0.999, (d == 0).mean()
    assert (d.max(axis=2) > 1).mean() < 1e-4, (d.max(axis=2) > 1).mean()


@pytest.mark.parametrize("name", ["3_adding_a_sphere", "4_objects"])
def test_legacy_deterministic_stages_match_reference_png(ctx, name):
    """legacy/PT_in_one_weekend/{3_adding_a_sphere,4_objects}.png through the GPU path: legacy camera (half-angle fov),
    lattice rays, normals as colours, rounding cast of the old ti.imwrite."""
    import os
    from PIL import Image
    from conftest import GOLDEN
    from test_oracle_goldens import _legacy_stage
    W, H = 400, 225
    world, cam = _legacy_stage(name, W, H)
    r = L.Renderer(W, H, ctx)
    r.render(world.device_scene(ctx), cam.to_struct(), 1, 32, L.PT_SHADE_V2_NORMALS, flags=L.PT_FLAG_PIXEL_GRID)
    gold = np.asarray(Image.open(os.path.join(GOLDEN, f"legacy_{name}_{W}x{H}.png")).convert("RGB"), np.int32)
    d = np.abs(L.to_uint8(r.mean(), rounding=True).astype(np.int32) - gold)
    assert (d == 0).mean() > 0.999 and (d.max(axis=2) > 1).mean() < 1e-4, ((d == 0).mean(), d.max())


def test_fused_and_split_wavefronts_trace_the_same_paths(ctx):
    """Both wavefront forms key the RNG on (pixel, sample, bounce): same paths, images equal to summation order;
    small pools and short launches exercise regeneration, compaction and the tail."""
    W, H = 80, 45
    world, cam = scenes.scene_10_final((W, H))
    sc = world.device_scene(ctx)
    ref = None
    for mode, cap, k, tm in [(L.PT_MODE_SPLIT, 0, 0, 0), (L.PT_MODE_FUSED, 0, 0, 0), (L.PT_MODE_FUSED, 1024, 3, 0),
                             (L.PT_MODE_FUSED, 7000, 1, 0), (L.PT_MODE_SPLIT, 2048, 0, 0), (L.PT_MODE_PERSIST, 0, 0, 0),
                             (L.PT_MODE_PERSIST, 0, 0, 32), (L.PT_MODE_PERSIST, 0, 0, 1), (L.PT_MODE_AUTO, 0, 0, 0)] + EXP_CASES:
        r = L.Renderer(W, H, ctx)
        st = r.render(sc, cam.to_struct(), 24, 32, seed=5, mode=mode, pool_capacity=cap, segments_per_launch=k,
                      serve_min=tm)
        m = r.mean()
        if ref is None:
            ref, seg = m, st.segments
        assert st.paths == W * H * 24
        assert abs(int(st.segments) - int(seg)) <= 1e-3 * seg
        assert np.allclose(m, ref, rtol=2e-3, atol=2e-4)
        assert abs(m.mean() / ref.mean() - 1) < 1e-4


@pytest.mark.parametrize("size,spp", [((101, 53), 19), ((7, 3), 40), ((1, 1), 64), ((33, 130), 1)])
def test_odd_image_sizes_and_sample_counts_agree_across_kernels(ctx, size, spp):
    """Image sizes that are no multiple of the 8x4 tile, sample counts that are no multiple of the 16-sample work
    unit, images smaller than a warp: every kernel form renders the same paths (and exactly W*H*spp of them)."""
    W, H = size
    world, cam = scenes.scene_10_final((W, H))
    sc = world.device_scene(ctx)
    ref = None
    for mode in [L.PT_MODE_SPLIT, L.PT_MODE_FUSED, L.PT_MODE_PERSIST] + EXP_MODES:
        r = L.Renderer(W, H, ctx)
        st = r.render(sc, cam.to_struct(), spp, 32, seed=11, mode=mode)
        acc = r.accum.cpu().numpy()
        assert st.paths == W * H * spp
        if ref is None:
            ref, seg = acc, int(st.segments)
        assert int(st.segments) == seg or abs(int(st.segments) - seg) <= 1e-3 * seg
        assert np.array_equal(acc[:, 3], ref[:, 3])                    # contributing paths per pixel: exact
        assert np.allclose(acc[:, :3], ref[:, :3], rtol=2e-3, atol=2e-4 * spp)
    empty = L.Renderer(W, H, ctx)
    st0 = empty.render(sc, cam.to_struct(), 0, 32, seed=11)            # zero samples: nothing happens
    assert st0.paths == 0 and st0.segments == 0 and float(empty.accum.abs().sum()) == 0.0


def test_progressive_and_sample_split_equal_single_render(ctx):
    """spp_offset makes renders splittable (multi-GPU) and continuable (legacy render(moved=False)):
    the union of sample ranges is the same set of paths as one render."""
    W, H = 96, 54
    world, cam = scenes.scene_9_dof((W, H))
    sc = world.device_scene(ctx)
    a = L.Renderer(W, H, ctx)
    a.render(sc, cam.to_struct(), 32, 32, seed=4)
    b = L.Renderer(W, H, ctx)
    b.render(sc, cam.to_struct(), 8, 32, seed=4)              # samples 0..7
    b.render(sc, cam.to_struct(), 24, 32, seed=4)             # continues at 8
    assert b.spp_done == 32
    assert np.allclose(a.mean(), b.mean(), rtol=1e-4, atol=1e-5)  # equal up to fp32 atomic summation order


def test_depth_limit_drops_paths(ctx, oracle):
    W, H = 64, 36
    world, cam = scenes.scene_8_refract((W, H))
    r = L.Renderer(W, H, ctx)
    st = r.render(world.device_scene(ctx), cam.to_struct(), 16, 1, seed=1)  # propagate_limit = 1: only direct sky
    assert st.segments == st.paths
    osum, _, _ = oracle.render(oracle.scene_from_world(world), cam.to_struct(), W, H, 16, 1, seed=1)
    assert np.allclose(r.mean() * 16, osum, rtol=1e-4, atol=1e-4)


# ---- raw triangles: LBVH + Moller-Trumbore vs brute-force reference test ---------------------------
def test_random_triangles_match_bruteforce_oracle(ctx, oracle):
    n_tri, n_rays = 20000, 20000
    tris = oracle.random_triangles(n_tri, 12345, 0.03)
    rays = oracle.random_rays(n_rays, 54321)
    sc = L.Scene(ctx)
    sc.set_triangles(tris)
    sc.build()
    gid, gt, st = ctx.trace_batch(sc, rays, counters=True)
    oid, ot, ot2 = oracle.trace_triangles(tris, rays, want_second=True)
    agree = gid == oid
    # disagreements must be explainable as ties or edge grazes
    bad = np.flatnonzero(~agree)
    assert len(bad) <= max(2, int(2e-4 * n_rays)), len(bad)
    if len(bad):
        tg, wg = oracle.triangle_eval(tris, gid[bad], rays[bad])
        to, wo = oracle.triangle_eval(tris, oid[bad], rays[bad])
        tie = np.abs(gt[bad] - ot[bad]) <= 1e-5 * np.abs(ot[bad])
        edge = (np.abs(wg) < 1e-4) | (np.abs(wo) < 1e-4)
        assert np.all(tie | edge)
    hit = agree & (oid >= 0)
    assert hit.sum() > 0.2 * n_rays
    assert np.all(np.abs(gt[hit] - ot[hit]) <= 1e-5 * ot[hit])
    assert np.all(gt[agree & (oid < 0)] == -1)
    # the oracle walking the GPU-built tree with the reference triangle test finds the same hits
    nodes, glob = sc.bvh_download()
    assert len(glob) == 0 and nodes.shape[0] == n_tri - 1
    bid, bt, counts = oracle.trace_bvh2(nodes, tris, rays)
    assert np.array_equal(bid, oid) and np.array_equal(bt, ot)
    assert st.nodes_visited > 0 and st.prims_tested > 0


def _tree_stats(nodes, n_prims):
    """(max depth, leaves seen once each?) of a downloaded BVH2 (nodes [N,16], refs at floats 12, 13)."""
    kids = nodes[:, 12:14].copy().view(np.int32)
    depth = np.zeros(nodes.shape[0], np.int64)
    seen = np.zeros(n_prims, np.int64)
    order = [0]
    for i in order:          # breadth-first: parents before children
        for c in kids[i]:
            if c >= 0:
                depth[c] = depth[i] + 1
                order.append(int(c))
            else:
                seen[~c] += 1
    return int(depth.max()) + 1, len(order) == nodes.shape[0] and bool((seen == 1).all())


@pytest.mark.parametrize("builder", ["ploc", "lbvh", "sah", ""])   # "": all candidates are built, the lowest SAH cost is kept
def test_tree_builders_are_valid_and_equivalent(ctx, oracle, builder, monkeypatch):
    """The hierarchy builders (Karras LBVH, PLOC over the same Morton order, the host full-sweep SAH of tiny trees) emit a proper binary tree over every
    primitive, shallow enough for the 64-entry traversal stack — also with thousands of exactly duplicated triangles
    (the Genshin models double their two-sided faces) — and the closest hits do not depend on which one built it."""
    monkeypatch.setenv("PT_BUILDER", builder)
    tris = oracle.random_triangles(30000, 99, 0.03)
    tris[20000:] = tris[123]                       # 10000 exact copies of one triangle
    tris[10000:20000] = tris[:10000]               # and 10000 faces doubled
    rays = oracle.random_rays(20000, 7)
    sc = L.Scene(ctx)
    sc.set_triangles(tris)
    sc.build()
    nodes, glob = sc.bvh_download()
    assert len(glob) == 0 and nodes.shape[0] == len(tris) - 1
    depth, proper = _tree_stats(nodes, len(tris))
    assert proper and depth <= 56, (proper, depth)
    gid, gt, _ = ctx.trace_batch(sc, rays)
    oid, ot = oracle.trace_triangles(tris, rays)[:2]
    same = gid == oid                              # lowest id wins exact ties on both sides; edge grazes aside
    assert same.mean() > 0.9995, same.mean()
    hit = same & (oid >= 0)
    assert hit.mean() > 0.2 and np.all(np.abs(gt[hit] - ot[hit]) <= 1e-5 * ot[hit])
    print(f"{builder}: depth {depth}")


def test_host_trace_pipeline_equals_device_path(ctx):
    """pt_trace_batch (host rays in, host ids/t out): the chunked three-stream pipeline (staging threads, H2D | sort +
    trace + unpack | D2H) returns exactly what one pt_trace_batch_device call over the whole batch returns — several
    chunks, a ragged last one, result arrays reused across calls, a batch smaller than one chunk."""
    import torch
    n_tri = 200_000
    sc = L.Scene(ctx)
    sc.set_random_triangles(n_tri, 4242, 0.02)
    sc.build()
    for n_rays in (1_234_567, 70_001):
        rays = torch.empty((2 * n_rays, 4), dtype=torch.float32, device="cuda")
        ctx.random_rays_device(rays.data_ptr(), n_rays, 31337)
        hits = torch.empty((n_rays, 4), dtype=torch.float32, device="cuda")
        ctx.trace_batch_device(sc, rays.data_ptr(), n_rays, hits.data_ptr())
        torch.cuda.synchronize()
        h = hits.cpu().numpy()
        ref_id = h[:, 1].copy().view(np.int32)
        ref_t = np.where(ref_id >= 0, h[:, 0], np.float32(-1.0)).astype(np.float32)
        rays_h = rays.cpu().numpy().reshape(n_rays, 8)
        ids, t = np.full(n_rays, -7, np.int32), np.full(n_rays, -7.0, np.float32)
        for _ in range(2):   # the second call reuses staging and result arrays
            ctx.trace_batch(sc, rays_h, out=(ids, t))
            assert np.array_equal(ids, ref_id) and np.array_equal(t, ref_t)
            ids[:] = -7; t[:] = -7.0
        ids2, t2, _ = ctx.trace_batch(sc, rays_h)            # fresh result arrays
        assert np.array_equal(ids2, ref_id) and np.array_equal(t2, ref_t)
        assert (ref_id >= 0).mean() > 0.2
    with pytest.raises(L.PtError):
        ctx.trace_batch(sc, rays_h, out=(np.zeros(3, np.int32), np.zeros(3, np.float32)))


def test_trace_kernels_agree_bit_for_bit(ctx, oracle):
    """pt_trace_batch_device: the persistent while-while warps (ray-sorted or in batch order, any refill
    threshold) and the one-ray-per-thread kernel return identical (t, prim, u, v) records in batch order."""
    import torch
    n_tri, n_rays = 200_000, 300_000
    sc = L.Scene(ctx)
    sc.set_random_triangles(n_tri, 777, 0.02)
    sc.build()
    rays = torch.empty((2 * n_rays, 4), dtype=torch.float32, device="cuda")
    ctx.random_rays_device(rays.data_ptr(), n_rays, 999)
    # ray intervals: a third of the rays get a finite tmax (some shorter than their first hit), a third a large tmin
    g = torch.Generator(device="cuda").manual_seed(5)
    u = torch.rand(n_rays, generator=g, device="cuda")
    rays[1::2, 3] = torch.where(u < 0.33, 0.3 + 2.0 * torch.rand(n_rays, generator=g, device="cuda"), rays[1::2, 3])
    rays[0::2, 3] = torch.where(u > 0.66, 0.5 + torch.rand(n_rays, generator=g, device="cuda"), rays[0::2, 3])
    ref = None
    for flags in (L.PT_FLAG_TRACE_SIMPLE, 0, L.PT_FLAG_NO_SORT, L.PT_FLAG_NO_QNODES, 1 << 8 | 1 << 14, 32 << 8 | 32 << 14,
                  L.PT_FLAG_COUNTERS):
        hits = torch.full((n_rays, 4), 7.0, dtype=torch.float32, device="cuda")
        st = ctx.trace_batch_device(sc, rays.data_ptr(), n_rays, hits.data_ptr(), flags)
        torch.cuda.synchronize()
        h = hits.cpu().numpy().view(np.int32)
        if ref is None:
            ref = h
            assert (h[:, 1] >= 0).mean() > 0.2
        assert np.array_equal(h, ref), flags
        if flags & L.PT_FLAG_COUNTERS:
            assert st.nodes_visited > n_rays and st.prims_tested > 0
    # against the CPU oracle walking the same tree (bounded sample)
    nodes, _ = sc.bvh_download()
    tris = oracle.random_triangles(n_tri, 777, 0.02)
    r_h = rays[:2 * 20000].cpu().numpy().reshape(-1, 8)
    oid, ot, _ = oracle.trace_bvh2(nodes, tris, r_h)
    gid = ref[:20000, 1]
    plain = (r_h[:, 3] == np.float32(1e-4)) & np.isinf(r_h[:, 7])   # the oracle's tree walk knows only the default interval
    assert plain.sum() > 5000 and (gid[plain] == oid[plain]).mean() > 0.999
    t_g = ref[:20000, 0].view(np.float32)
    hit = gid >= 0
    assert np.all(t_g[hit] >= r_h[hit, 3]) and np.all(t_g[hit] <= r_h[hit, 7])   # every hit lies inside its ray's interval
    assert (~hit).mean() > 0.1 and hit.mean() > 0.1


def test_device_triangle_generator_matches_oracle(ctx, oracle):
    n = 5000
    sc = L.Scene(ctx)
    sc.set_random_triangles(n, 777, 0.01)
    sc.build()
    g = sc.triangles_download(n)
    t = oracle.random_triangles(n, 777, 0.01)
    assert np.array_equal(g[:, 0:3], t[:, 0:3])
    assert np.array_equal(g[:, 4:7], t[:, 3:6] - t[:, 0:3])
    assert np.array_equal(g[:, 8:11], t[:, 6:9] - t[:, 0:3])
