"""Shared builders for the legacy-path tests."""
import os

import numpy as np

from learn_path_tracing_b200 import legacy

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CACHE = os.path.join(ROOT, "scenes_cache")


def synthetic_legacy_world(seed=0, grid=14):
    """A small self-contained legacy scene: bumpy textured grid mesh (uv tiled beyond [0,1]), a +-50 ground plane
    (becomes 'global' primitives), two textured spheres (one transparent, with a normal map), a 3-area atlas with a
    non-square area (exercises bilinear's wrap-with-width quirk) and a non-square environment area."""
    rng = np.random.default_rng(seed)
    n = grid + 1
    xs, zs = np.meshgrid(np.linspace(-2, 2, n), np.linspace(-2, 2, n), indexing="ij")
    ys = 0.6 + 0.25 * np.sin(2.1 * xs) * np.cos(1.7 * zs) + 0.03 * rng.standard_normal(xs.shape)
    pos = np.stack([xs, ys, zs], -1).reshape(-1, 3).astype(np.float32)
    nrm = np.zeros_like(pos)
    nrm[:, 1] = 1.0
    nrm[:, 0] = -0.25 * 2.1 * np.cos(2.1 * xs).ravel() * np.cos(1.7 * zs).ravel()
    nrm[:, 2] = 0.25 * 1.7 * np.sin(2.1 * xs).ravel() * np.sin(1.7 * zs).ravel()
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    uv = (np.stack([xs, zs], -1).reshape(-1, 2) * 0.8 + 0.3).astype(np.float32)  # spans about [-1.3, 1.9]: wraps
    faces = []
    for i in range(grid):
        for j in range(grid):
            a, b, c, d = i * n + j, (i + 1) * n + j, (i + 1) * n + j + 1, i * n + j + 1
            tid = (i + j) % 2
            faces.append([a, a, a, b, b, b, c, c, c, tid])
            faces.append([a, a, a, c, c, c, d, d, d, tid])
    w = legacy.World()
    w.add_mesh(pos, nrm, uv, np.array(faces, np.int32))
    gp = np.array([[50, 0, -50], [-50, 0, -50], [-50, 0, 50], [50, 0, 50]], np.float32)
    gn = np.array([[0, 1, 0]], np.float32)
    gt = np.array([[0, 0], [10, 0], [10, 10], [0, 10]], np.float32)
    w.add_mesh(gp, gn, gt, np.array([[0, 0, 0, 1, 0, 1, 2, 0, 2, 1], [0, 0, 0, 2, 0, 2, 3, 0, 3, 1]], np.int32))
    w.add_sphere(legacy.Sphere([-1.2, 1.6, 0.3], 0.6, 0, 2))
    w.add_sphere(legacy.Sphere([1.3, 1.5, -0.4], 0.5, 1, 2))
    tex = rng.integers(0, 256, size=(160, 64, 8), dtype=np.uint8)
    tex[:, :, 3] = rng.integers(0, 256, size=(160, 64)) // 2 + 64   # roughness
    tex[:, :, 7] = np.where(rng.random((160, 64)) < 0.3, rng.integers(0, 256, size=(160, 64)), 0)  # metallic
    tex[:, :, 4:6] = 128 + rng.integers(-40, 40, size=(160, 64, 2))
    tex[:, :, 6] = 230
    areas = [[0, 0, 64, 64], [64, 0, 96, 64], [96, 0, 160, 32]]
    w.set_atlas(tex, areas, [1, 1, 0])
    ex, ey = np.meshgrid(np.linspace(0, 1, 64), np.linspace(0, 1, 32), indexing="ij")
    env = np.zeros((64, 64, 3), np.float32)
    env[:, :32, 0] = 0.4 + 0.6 * ey
    env[:, :32, 1] = 0.5 + 0.4 * np.sin(6.28 * ex) ** 2
    env[:, :32, 2] = 0.9
    w.set_environment(0)
    w.set_environment_image(env, [0, 0, 64, 32])
    cam = legacy.Camera((96, 64))
    cam.set_fov(22)
    cam.set_position(legacy.Vec3f([0.5, 2.6, -7.0]))
    cam.look_at(legacy.Vec3f([0, 1.0, 0]))
    cam.set_len(7.0, 0.05)
    return w, cam


def all_triangles(world):
    """tris9 [F,9] of every mesh face in primitive order, and the primitive-id offset of triangles."""
    tris = []
    for m in world.meshes:
        f = m["indices"]
        p = m["positions"]
        tris.append(np.concatenate([p[f[:, 0]], p[f[:, 3]], p[f[:, 6]]], axis=1))
    return np.concatenate(tris).astype(np.float32), len(world.spheres)


def cached_world(name):
    from learn_path_tracing_b200 import scene_cache
    path = os.path.join(CACHE, name + ".npz")
    if not os.path.exists(path):
        return None
    return scene_cache.load_cache(path)


def mesh_camera(resolution):
    """15_module.py:1068-1072."""
    cam = legacy.Camera(resolution)
    cam.set_fov(30)
    cam.set_position(legacy.Vec3f([0, 8, -30]))
    cam.look_at(legacy.Vec3f([0, 8, 0]))
    return cam


def hit_tolerance(tris, tri_ids, rays, t):
    """|dt| bound for one (ray, triangle) pair: the contract's 1e-5 relative, plus the unavoidable rounding of f32
    coordinates (a few ulps of the scene scale) amplified by 1/|d.N| when the ray grazes the triangle's plane."""
    T = tris[np.maximum(tri_ids, 0)]
    n = np.cross(T[:, 3:6] - T[:, 0:3], T[:, 6:9] - T[:, 0:3])
    n /= np.linalg.norm(n, axis=1, keepdims=True)
    dn = np.abs((rays[:, 4:7] * n).sum(1))
    scale = np.maximum(np.abs(rays[:, :3]).max(1), np.abs(T).max(1))
    return 1e-5 * np.abs(t) + 16 * 1.2e-7 * scale / np.maximum(dn, 1e-12)
