#!/usr/bin/env python
"""Builds scenes_cache/*.npz from the reference's data files (run by __graft_entry__.build() in the container where
/root/reference is mounted; the caches are git-ignored but travel to the GPU box with the gpurun snapshot).

    yoimiya_ground_{small,full}.npz   Yoimiya_ShapeChange.world.npy: ground plane + mesh, sky.png environment, the four
                                      diffuse maps (full: 2048^2 as recorded; small: 512^2 for tests), ground = documented
                                      constant fallback because granite-gray-white_albedo/_roughness/_normal are absent
    zhongli_{small,full}.npz          Zhongli.world.npy (old format): textures in MTL first-seen order (14_mesh.py:991-999)
    ganyu_{small,full}.npz            Ganyu.world.npy
    demo.npz                          demo.world.npy: one quad + one textured sphere (unit-test fixture)
No reference SOURCE is copied: only geometry arrays and images (data) are converted.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from learn_path_tracing_b200 import legacy, scene_cache  # noqa: E402

REF = "/root/reference"
OUT = os.path.join(ROOT, "scenes_cache")


def _with_sky(w):
    w.environments = legacy.TextureManager((256, 256))
    w.environments.add("./textures/sky.png", 0)
    w.environments.build()
    w.set_environment(0)


def new_format(name, tex_size):
    w = legacy.World()
    w.load(os.path.join(REF, "legacy", name + ".world.npy"), load_images=False)
    if tex_size != 2048:  # shrink every recorded area by the same factor
        f = 2048 // tex_size
        for c in w.textures.configs:
            a = c["area"].as_list()
            c["area"] = legacy.TextureArea([a[0] // f, a[1] // f], [a[2] // f, a[3] // f])
            c["size"] = (c["size"][0] // f, c["size"][1] // f)
    w.load_textures()
    return w


def old_format(name, obj, tex_size):
    w = legacy.World()
    w.load(os.path.join(REF, "legacy", name + ".world.npy"), load_images=False)
    _, _, _, _, textures = legacy.load_obj(os.path.join(REF, "assets", "models", obj, obj + ".obj"), 0, flip_z=True,
                                           flip_textcoord=True, transform=legacy.rotate(np.pi, 0))
    w.textures = legacy.TextureManager((tex_size * len(textures), tex_size))
    for i, t in enumerate(textures):  # 14_mesh.py:991-995: area i = [i*S, 0, (i+1)*S, S], image resized into it
        w.textures.configs.append({"file_path": t["file_path"], "size": (tex_size, tex_size), "id": t["id"],
                                   "area": legacy.TextureArea([i * tex_size, 0], [(i + 1) * tex_size, tex_size])})
    _with_sky(w)
    w.load_textures()
    return w


def main():
    if not os.path.isdir(REF):
        print("prepare_assets: /root/reference not mounted, nothing to do")
        return
    os.makedirs(OUT, exist_ok=True)
    jobs = [
        ("yoimiya_ground_small", lambda: new_format("Yoimiya_ShapeChange", 512)),
        ("yoimiya_ground_full", lambda: new_format("Yoimiya_ShapeChange", 2048)),
        ("zhongli_small", lambda: old_format("Zhongli", "Zhongli", 512)),
        ("zhongli_full", lambda: old_format("Zhongli", "Zhongli", 2048)),
        ("ganyu_small", lambda: old_format("Ganyu", "Ganyu", 512)),
        ("ganyu_full", lambda: old_format("Ganyu", "Ganyu", 2048)),
    ]
    for name, make in jobs:
        path = os.path.join(OUT, name + ".npz")
        if os.path.exists(path) and "--force" not in sys.argv:
            continue
        w = make()
        scene_cache.save_cache(path, w, {"source": name})
        print(f"prepare_assets: {path} {os.path.getsize(path) / 1e6:.1f} MB, {sum(len(m['indices']) for m in w.meshes)} triangles")
    # demo: quad + sphere, old format, no textures recorded: a procedural 64x64 checker stands in for texture 0
    path = os.path.join(OUT, "demo.npz")
    if not os.path.exists(path) or "--force" in sys.argv:
        w = legacy.World()
        w.load(os.path.join(REF, "legacy", "demo.world.npy"), load_images=False)
        xs, ys = np.meshgrid(np.arange(64), np.arange(64), indexing="ij")
        tex = np.empty((64, 64, 8), np.uint8)
        tex[...] = legacy.FALLBACK_TEXEL
        tex[:, :, 0] = np.where((xs // 8 + ys // 8) % 2, 230, 60)
        tex[:, :, 1] = 120
        tex[:, :, 3] = np.where(xs < 32, 255, 64)   # roughness
        tex[:, :, 7] = np.where(ys < 32, 0, 200)    # metallic
        w.set_atlas(tex, [[0, 0, 64, 64]], [1])
        _with_sky(w)
        # atlas set by hand above; only the environment comes from a file
        w._env = (legacy.load_environment_image("./textures/sky.png"), [0, 0, 256, 256])
        scene_cache.save_cache(path, w, {"source": "demo"})
        print(f"prepare_assets: {path}")


if __name__ == "__main__":
    main()
