"""Compact scene caches (.npz) for legacy worlds: geometry (+ the reference's stored SAH tree), the 8-bit albedo
atlas, areas/flags and the environment map in one file, so a mesh scene can be rendered where neither the
reference checkout nor its image assets exist (the GPU box).  Written by tools/prepare_assets.py."""
from __future__ import annotations

import numpy as np

from . import legacy


def save_cache(path, world: "legacy.World", meta: dict | None = None):
    texels, areas, flags = world._atlas
    out = {"n_mesh": np.int32(len(world.meshes)), "areas": areas, "flags": flags,
           "albedo": np.ascontiguousarray(texels[:, :, 0:3]),
           "environment": np.int32(-1 if world.environment is None else world.environment)}
    # non-albedo channels are constant for plain diffuse maps and the documented fallback: store them only if not
    rest = texels[:, :, 3:8]
    if not np.all(rest == legacy.FALLBACK_TEXEL[3:8]):
        out["rest"] = np.ascontiguousarray(rest)
    if world._env is not None:
        out["env"], out["env_area"] = world._env[0], np.asarray(world._env[1], np.int32)
    if world.spheres:
        cr, tr, tx = world.sphere_arrays()
        out["sph_cr"], out["sph_tr"], out["sph_tx"] = cr, tr, tx
    for i, m in enumerate(world.meshes):
        out[f"m{i}_pos"], out[f"m{i}_nrm"], out[f"m{i}_uv"], out[f"m{i}_faces"] = (m["positions"], m["normals"],
                                                                                 m["texture_coords"], m["indices"])
        if m.get("tree") is not None:
            t = m["tree"]
            for k in ("left", "right", "low", "high", "data", "leaf_cut"):
                out[f"m{i}_tree_{k}"] = t[k]
            out[f"m{i}_tree_depth"] = np.int32(t["max_depth"])
    for k, v in (meta or {}).items():
        out[f"meta_{k}"] = np.asarray(v)
    np.savez_compressed(path, **out)


def load_cache(path) -> "legacy.World":
    z = np.load(path)
    w = legacy.World()
    alb = z["albedo"]
    texels = np.empty(alb.shape[:2] + (8,), np.uint8)
    texels[:, :, 0:3] = alb
    texels[:, :, 3:8] = z["rest"] if "rest" in z.files else legacy.FALLBACK_TEXEL[3:8]
    w.set_atlas(texels, z["areas"], z["flags"])
    env_id = int(z["environment"])
    w.environment = None if env_id < 0 else env_id
    if "env" in z.files:
        w.set_environment_image(z["env"], z["env_area"])
    if "sph_cr" in z.files:
        for cr, tr, tx in zip(z["sph_cr"], z["sph_tr"], z["sph_tx"]):
            w.spheres.append(legacy.Sphere(cr[:3], cr[3], tr, tx))
    for i in range(int(z["n_mesh"])):
        tree = None
        if f"m{i}_tree_left" in z.files:
            tree = {k: z[f"m{i}_tree_{k}"] for k in ("left", "right", "low", "high", "data", "leaf_cut")}
            tree["max_depth"] = int(z[f"m{i}_tree_depth"])
        w.add_mesh(z[f"m{i}_pos"], z[f"m{i}_nrm"], z[f"m{i}_uv"], z[f"m{i}_faces"], tree=tree)
    return w
