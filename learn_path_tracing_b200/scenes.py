"""The reference's built-in scenes, restated as data.

Each returns (World, Camera-configuring function).  Sources: taichi_pathtracer/{6_diffuse,7_reflect,
8_refract,9_dof,10_final}/__main__.py.
"""
from __future__ import annotations

import random

from .camera import Camera
from .dtypes import Material, Sphere, Vec3f
from .world import World


def _six_spheres(with_glass=True, ground_y=-10000.5):
    s1 = Sphere(Vec3f([0.0, 0.0, 0.0]), 0.5, material=Material(albedo=Vec3f([0.25, 0.25, 0.5]), roughness=0.5, metallic=0, ior=1.5))
    s2 = Sphere(Vec3f([-1.0, 0.0, 0.0]), 0.5, material=Material(albedo=Vec3f([0.25, 0.5, 0.25]), roughness=0, metallic=1, ior=1.5))
    s3 = Sphere(Vec3f([1.0, 0.0, 0.0]), 0.5, material=Material(albedo=Vec3f([0.5, 0.25, 0.25]), roughness=0.5, metallic=1, ior=1.5))
    ground = Sphere(Vec3f([0, ground_y, 0.0]), 10000, material=Material(albedo=Vec3f([0.25, 0.25, 0.25]), roughness=0.5, metallic=0, ior=1.5))
    if not with_glass:
        return [s1, s2, s3, ground]
    s4 = Sphere(Vec3f([-0.5, 0.866, 0]), 0.5, material=Material(albedo=Vec3f([1, 1, 1]), roughness=0, metallic=0, ior=1.5, transparency=1))
    s5 = Sphere(Vec3f([0.5, 0.866, 0]), 0.5, material=Material(albedo=Vec3f([0.5, 1, 0.5]), roughness=0.5, metallic=0, ior=1.5, transparency=1))
    return [s1, s2, s3, s4, s5, ground]


def scene_2_camera_and_ray(resolution=(1280, 720)):
    """2_camera_and_ray/__main__.py:26-28 — an empty World(): the camera pitched up by 30 degrees looks at the sky."""
    w = World()
    cam = Camera(resolution)
    cam.set_direction(0, 30, 0)
    return w, cam


def scene_3_adding_a_sphere(resolution=(1280, 720)):
    """3_adding_a_sphere/__main__.py:27-38,49-51 — one sphere two units in front of a camera at the origin."""
    w = World([Sphere(Vec3f([0.0, 0.0, -2.0]), 0.5)])
    cam = Camera(resolution)
    cam.set_direction(0, 0)
    return w, cam


def scene_5_anti_aliasing(resolution=(1280, 720)):
    """5_anti_aliasing/__main__.py:44-50 (also 4_objects) — two plain spheres, normals shown as colours."""
    w = World([Sphere(Vec3f([0.0, 0.0, 0.0]), 0.5), Sphere(Vec3f([0, -100.5, 0]), 100)])
    cam = Camera(resolution)
    cam.set_direction(0, 0)
    cam.set_position(Vec3f([0, 0, 3]))
    return w, cam


def scene_legacy_6_diffuse(resolution=(400, 225)):
    """legacy/PT_in_one_weekend/6_diffuse.py:185-189 — two albedo-only spheres, legacy (half-angle) camera at the origin,
    fov 60, looking down -z; rendered with LegacyStage6BSDF, absorptivity 0.5, depth 100, gamma only."""
    from . import legacy
    w = World([Sphere(Vec3f([0.0, 0.0, -1.0]), 0.5, Vec3f([0.5, 0.5, 1.0])), Sphere(Vec3f([0, -100.5, -1]), 100, Vec3f([1, 1, 1]))])
    cam = legacy.Camera(resolution, fov=60)
    cam.set_direction(0, 0)
    return w, cam


def scene_legacy_7_reflect(resolution=(400, 225)):
    """legacy/PT_in_one_weekend/7_reflect.py:228-236 — three spheres on a metallic ground, legacy camera at (0, -0.5, 4)...
    (the committed 7_reflect.png is a 400x225 render; fov 45 is the HALF angle: view width 2 tan 45 = 2);
    rendered with LegacyStage7BSDF, absorptivity 0.5, depth 100, gamma only."""
    from . import legacy

    def mat(albedo, roughness, metallic):
        return Material(albedo=Vec3f(albedo), roughness=roughness, metallic=metallic, ior=1.5)
    w = World([Sphere(Vec3f([0.0, 0.0, 0.0]), 0.5, material=mat([0.5, 0.5, 1], 1, 0)),
               Sphere(Vec3f([-1.0, 0.0, 0.0]), 0.5, material=mat([0.5, 1, 0.5], 0, 1)),
               Sphere(Vec3f([1.0, 0.0, 0.0]), 0.5, material=mat([1, 0.5, 0.5], 0.25, 1)),
               Sphere(Vec3f([0, -10000.5, 0.0]), 10000, material=mat([0.5, 0.5, 0.5], 0.2, 1))])
    cam = legacy.Camera(resolution, fov=45)
    cam.set_direction(0, 0)
    cam.set_position(Vec3f([0, -0.5, 4]))
    return w, cam


def scene_6_diffuse(resolution=(1280, 720)):
    """6_diffuse/__main__.py:62-71 — Lambert-only spheres (albedo only)."""
    w = World([
        Sphere(Vec3f([0.0, 0.0, 0.0]), 0.5, Vec3f([0.25, 0.25, 0.5])),
        Sphere(Vec3f([-1.0, 0.0, 0.0]), 0.5, Vec3f([0.25, 0.5, 0.25])),
        Sphere(Vec3f([1.0, 0.0, 0.0]), 0.5, Vec3f([0.5, 0.25, 0.25])),
        Sphere(Vec3f([0, -10000.5, 0.0]), 10000, Vec3f([0.25, 0.25, 0.25])),
    ])
    cam = Camera(resolution)
    cam.set_direction(0, 0)
    cam.set_position(Vec3f([0, 0, 4]))
    return w, cam


def scene_7_reflect(resolution=(1280, 720)):
    """7_reflect/__main__.py:65-74."""
    cam = Camera(resolution)
    cam.set_direction(0, 0)
    cam.set_position(Vec3f([0, 0, 4]))
    return World(_six_spheres(with_glass=False)), cam


def scene_8_refract(resolution=(1280, 720)):
    """8_refract/__main__.py:65-79."""
    cam = Camera(resolution)
    cam.set_direction(0, 0)
    cam.set_position(Vec3f([0, 0.4, 4]))
    return World(_six_spheres()), cam


def scene_9_dof(resolution=(1280, 720)):
    """9_dof/__main__.py:69-80."""
    cam = Camera(resolution)
    cam.set_position(Vec3f([3, 0.5, 2]))
    cam.look_at(Vec3f([0.0, 0.35, 0.0]))
    cam.set_len(focal_length=cam.position.norm(), aperture=0.2)
    return World(_six_spheres()), cam


def random_scene(size=11, seed=None):
    """10_final/__main__.py:12-45.  The reference draws from the unseeded global `random`; pass a seed
    for a reproducible sphere list (the benchmark uses seed 20261018)."""
    rnd = random.Random(seed) if seed is not None else random
    world = World()
    world.add(Sphere(Vec3f([0, -10000, 0]), 10000, material=Material(albedo=Vec3f([0.25, 0.25, 0.25]), roughness=0.5, metallic=0, ior=1.5, transparency=0)))
    for a in range(-size, size):
        for b in range(-size, size):
            choose_mat = rnd.random()
            center = Vec3f([a + 0.9 * rnd.random(), 0.2, b + 0.9 * rnd.random()])
            if (center - Vec3f([4, 0.2, 0])).norm() > 0.9:
                albedo = Vec3f([rnd.random(), rnd.random(), rnd.random()])
                if choose_mat < 0.8:
                    m = Material(albedo=albedo, roughness=rnd.random(), metallic=0, ior=1.5, transparency=0)
                elif choose_mat < 0.95:
                    m = Material(albedo=0.5 + 0.5 * albedo, roughness=0.5 * rnd.random(), metallic=1, ior=0, transparency=0)
                else:
                    m = Material(albedo=0.75 + 0.25 * albedo, roughness=0.2 * rnd.random(), metallic=0, ior=1.5, transparency=1)
                world.add(Sphere(center, 0.2, material=m))
    world.add(Sphere(Vec3f([0, 1, 0]), 1.0, material=Material(albedo=Vec3f([1, 1, 1]), roughness=0, metallic=0, ior=1.5, transparency=1)))
    world.add(Sphere(Vec3f([-4, 1, 0]), 1.0, material=Material(albedo=Vec3f([0.4, 0.2, 0.1]), roughness=0.5, metallic=0, ior=1.5, transparency=0)))
    world.add(Sphere(Vec3f([4, 1, 0]), 1.0, material=Material(albedo=Vec3f([0.7, 0.6, 0.5]), roughness=0, metallic=1, ior=0, transparency=0)))
    return world


def scene_10_final(resolution=(1280, 720), seed=20261018):
    """10_final/__main__.py:106-112 with a seeded sphere list."""
    cam = Camera(resolution)
    cam.set_position(Vec3f([13, 2, 3]))
    cam.look_at(Vec3f([0, 0, 0]))
    cam.set_fov(40)
    cam.set_len(10, 0.2)
    return random_scene(seed=seed), cam


# stages 2-4 shoot one ray per pixel through the lattice i/(W-1), j/(H-1) (render(..., pixel_grid=True), spp=1)
SCENES = {"2_camera_and_ray": scene_2_camera_and_ray, "3_adding_a_sphere": scene_3_adding_a_sphere,
          "4_objects": scene_5_anti_aliasing, "5_anti_aliasing": scene_5_anti_aliasing, "6_diffuse": scene_6_diffuse, "7_reflect": scene_7_reflect, "8_refract": scene_8_refract,
          "9_dof": scene_9_dof, "10_final": scene_10_final,
          "legacy_6_diffuse": scene_legacy_6_diffuse, "legacy_7_reflect": scene_legacy_7_reflect}
