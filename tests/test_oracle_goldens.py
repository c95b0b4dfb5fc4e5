"""Pins the CPU oracle against the reference's own committed renders:
outputs/{6_diffuse,7_reflect,8_refract,9_dof}.png, box-filtered to 320x180 (tests/golden/make_goldens.py).
These pin camera (pinhole + thin lens), Sphere.hit incl. the far-root rule, the v2 BSDFs, the sky,
ACES + gamma and the imwrite orientation."""
import os

import numpy as np
import pytest
from PIL import Image

import learn_path_tracing_b200 as L
from learn_path_tracing_b200 import scenes
from conftest import GOLDEN

CASES = [("6_diffuse", L.PT_SHADE_V2_DIFFUSE), ("7_reflect", L.PT_SHADE_V2), ("8_refract", L.PT_SHADE_V2),
         ("9_dof", L.PT_SHADE_V2)]


@pytest.mark.parametrize("name,model", CASES)
def test_oracle_matches_reference_render(oracle, name, model):
    W, H, SPP = 320, 180, 192
    world, cam = scenes.SCENES[name]((W, H))
    acc, _, st = oracle.render(oracle.scene_from_world(world), cam.to_struct(), W, H, SPP, 32, model, seed=3)
    img = L.to_uint8(oracle.postprocess(acc, 1.0 / SPP, aces=True, gamma=2.2)).astype(np.float64)
    gold = np.asarray(Image.open(os.path.join(GOLDEN, f"{name}_320x180.png")).convert("RGB"), np.float64)
    d = img - gold
    rmse, bias = float(np.sqrt((d**2).mean())), float(d.mean())
    # measured at 256 spp: rmse 1.6-2.5 / 255, |bias| < 0.1; a flipped image gives rmse ~90
    assert rmse < 3.5, (name, rmse)
    assert abs(bias) < 0.5, (name, bias)
    assert st.paths == W * H * SPP


def test_oracle_matches_reference_stage5_normals(oracle):
    """outputs/5_anti_aliasing.png (320x180, 100 spp, normals as colours, no tonemapping): an almost noise-free pin of
    Camera.get_rays (pinhole, pixel jitter), Sphere.hit / World.hit, the sky and the imwrite orientation + quantisation."""
    W, H, SPP = 320, 180, 128
    world, cam = scenes.scene_5_anti_aliasing((W, H))
    acc, _, st = oracle.render(oracle.scene_from_world(world), cam.to_struct(), W, H, SPP, 32, L.PT_SHADE_V2_NORMALS, seed=3)
    img = L.to_uint8(acc / SPP).astype(np.float64)
    gold = np.asarray(Image.open(os.path.join(GOLDEN, "5_anti_aliasing_320x180.png")).convert("RGB"), np.float64)
    d = img - gold
    rmse, bias = float(np.sqrt((d**2).mean())), float(d.mean())
    assert st.segments == st.paths          # no bounce
    assert rmse < 1.5, rmse                 # only silhouette pixels carry sampling noise (measured 0.6)
    assert abs(bias) < 0.3, bias
    assert float(np.sqrt(((img[::-1] - gold) ** 2).mean())) > 20  # a vertically flipped image is far off


EXACT = [("2_camera_and_ray", 1280, 720), ("3_adding_a_sphere", 1280, 720), ("4_objects", 320, 180)]


@pytest.mark.parametrize("name,W,H", EXACT)
def test_oracle_reproduces_the_deterministic_stages_exactly(oracle, name, W, H):
    """outputs/{2_camera_and_ray,3_adding_a_sphere,4_objects}.png use no random numbers (one ray per pixel through the
    lattice i/(W-1), j/(H-1), 2_camera_and_ray/camera.py:67): noise-free known answers for rotate() incl. a 30 degree
    pitch, the field-of-view geometry, Sphere.hit / World.hit, the normal, the sky gradient and imwrite's orientation
    and uint8 quantisation.  The oracle must reproduce them to the byte, up to a last-ulp flip at a quantisation
    threshold (measured: 99.997-99.9998 % of the bytes equal, never off by more than 1)."""
    world, cam = scenes.SCENES[name]((W, H))
    acc, _, st = oracle.render(oracle.scene_from_world(world), cam.to_struct(), W, H, 1, 32, L.PT_SHADE_V2_NORMALS, seed=1,
                               flags=L.PT_FLAG_PIXEL_GRID)
    img = L.to_uint8(acc).astype(np.int32)
    gold = np.asarray(Image.open(os.path.join(GOLDEN, f"{name}_{W}x{H}.png")).convert("RGB"), np.int32)
    d = np.abs(img - gold)
    assert st.paths == st.segments == W * H
    assert d.max() <= 1, d.max()
    assert (d == 0).mean() > 0.9999, (d == 0).mean()
    assert np.abs(img[::-1] - gold).mean() > 3      # orientation matters


def _legacy_stage(name, W, H):
    """legacy/PT_in_one_weekend/{3_adding_a_sphere,4_objects,5_anti_aliasing}.py: camera at the origin looking down -z with
    the LEGACY field of view (view_width = 2 tan(fov pi / 180), the formula 15_module.py:397-401 still uses)."""
    from learn_path_tracing_b200 import Sphere, Vec3f, World, legacy
    spheres = [Sphere(Vec3f([0.0, 0.0, -1.0]), 0.5)]
    if name != "3_adding_a_sphere":
        spheres.append(Sphere(Vec3f([0, -100.5, -1]), 100))
    cam = legacy.Camera((W, H), fov=60)
    cam.set_direction(0, 0)
    return World(spheres), cam


@pytest.mark.parametrize("name", ["3_adding_a_sphere", "4_objects"])
def test_oracle_reproduces_the_legacy_deterministic_stages_to_the_byte(oracle, name):
    """The legacy tutorial's lattice renders pin the legacy camera's half-angle field of view and the ROUNDING cast of
    the old ti.imwrite (measured: every byte of both 400x225 PNGs equal; the truncating cast of v2 matches only 51-64 %)."""
    W, H = 400, 225
    world, cam = _legacy_stage(name, W, H)
    acc, _, _ = oracle.render(oracle.scene_from_world(world), cam.to_struct(), W, H, 1, 32, L.PT_SHADE_V2_NORMALS, seed=1,
                              flags=L.PT_FLAG_PIXEL_GRID)
    gold = np.asarray(Image.open(os.path.join(GOLDEN, f"legacy_{name}_{W}x{H}.png")).convert("RGB"), np.int32)
    d = np.abs(L.to_uint8(acc, rounding=True).astype(np.int32) - gold)
    assert d.max() <= 1 and (d == 0).mean() > 0.9999, (d.max(), (d == 0).mean())
    assert (L.to_uint8(acc).astype(np.int32) == gold).mean() < 0.7   # the v2 cast is NOT what the legacy ti.imwrite did


def test_oracle_matches_legacy_stage5_with_the_rounding_cast(oracle):
    """legacy 5_anti_aliasing.png (400x225, 100 spp, jittered legacy camera): unbiased with the rounding cast (measured
    bias -0.0001, rmse 0.46), half a level dark with the truncating one."""
    W, H, SPP = 400, 225, 128
    world, cam = _legacy_stage("5_anti_aliasing", W, H)
    acc, _, _ = oracle.render(oracle.scene_from_world(world), cam.to_struct(), W, H, SPP, 32, L.PT_SHADE_V2_NORMALS, seed=1)
    gold = np.asarray(Image.open(os.path.join(GOLDEN, f"legacy_5_anti_aliasing_{W}x{H}.png")).convert("RGB"), np.float64)
    d = L.to_uint8(acc / SPP, rounding=True).astype(np.float64) - gold
    assert np.sqrt((d**2).mean()) < 1.0 and abs(d.mean()) < 0.1, (np.sqrt((d**2).mean()), d.mean())
    dt = L.to_uint8(acc / SPP).astype(np.float64) - gold
    assert dt.mean() < -0.3


def test_imwrite_reproduces_stage1_png_exactly():
    """outputs/1_save_img.png: image[i, j] = (i/256, j/256, 0) through ti.tools.imwrite (1_save_img/__main__.py:10-19) pins
    the field orientation (x right, y up) and the truncating uint8 cast of our imwrite/to_uint8 to the byte."""
    i, j = np.meshgrid(np.arange(256), np.arange(256), indexing="ij")
    field = np.stack([i / 256, j / 256, np.zeros_like(i, float)], -1).astype(np.float32)
    gold = np.asarray(Image.open(os.path.join(GOLDEN, "1_save_img_256x256.png")).convert("RGB"))
    assert np.array_equal(L.to_uint8(field), gold)
    assert np.array_equal(L.to_uint8(L.imread(os.path.join(GOLDEN, "1_save_img_256x256.png"))), gold)  # imread inverts it


def test_pixel_grid_is_independent_of_seed_and_sample(oracle):
    world, cam = scenes.scene_3_adding_a_sphere((64, 36))
    a, _, _ = oracle.render(oracle.scene_from_world(world), cam.to_struct(), 64, 36, 1, 32, L.PT_SHADE_V2_NORMALS, seed=1,
                            flags=L.PT_FLAG_PIXEL_GRID)
    b, _, _ = oracle.render(oracle.scene_from_world(world), cam.to_struct(), 64, 36, 1, 32, L.PT_SHADE_V2_NORMALS, seed=77,
                            spp_offset=5, flags=L.PT_FLAG_PIXEL_GRID)
    assert np.array_equal(a, b)


def test_segments_per_path_matches_survey(oracle):
    """8_refract averages 2.33 ray segments per path."""
    world, cam = scenes.scene_8_refract((160, 90))
    _, _, st = oracle.render(oracle.scene_from_world(world), cam.to_struct(), 160, 90, 32, 32, L.PT_SHADE_V2, seed=1)
    assert abs(st.segments / st.paths - 2.33) < 0.05


def test_postprocess_matches_python_surface(oracle):
    rng = np.random.default_rng(0)
    a = rng.random((8, 5, 3), dtype=np.float32) * 2.0
    ref = L.gamma_correction(L.ACES_tonemapping(a), 2.2)
    out = oracle.postprocess(a, 1.0, aces=True, gamma=2.2)
    assert np.allclose(out, ref, rtol=2e-5, atol=2e-6)


def test_rng_is_counter_based_and_uniform(oracle):
    a = oracle.rng4(5, 7, 1, 99)
    b = oracle.rng4(5, 7, 1, 99)
    assert np.array_equal(a, b) and np.all((a >= 0) & (a < 1))
    assert not np.array_equal(a, oracle.rng4(5, 7, 2, 99))
    u = np.stack([oracle.rng4(i, 0, 0, 1) for i in range(4096)])
    assert abs(u.mean() - 0.5) < 0.01 and abs(u.var() - 1 / 12) < 0.005


def test_legacy_6_diffuse_png_pins_the_legacy_scattering_helpers(oracle):
    """legacy/PT_in_one_weekend/6_diffuse.png is the reference's converged render (400x225, 8192 spp, depth 100) of the
    script as committed (6_diffuse.py:12-14,160-170,185-189).  It uses the functions 15_module.py still carries —
    sample_at_sphere / sample_diffuse (:295-325), the half-angle legacy camera — plus the tutorial's hit rule (t > 1e-3,
    near root) and throughput 0.5 * albedo.  The oracle's PT_SHADE_LEGACY_STAGE6 at 256 spp, gamma 2.2, ROUNDING cast (the
    legacy ti.imwrite) lands on it: measured rmse 1.20 / 255 (Monte Carlo noise of 256 spp), bias -0.01, 8x8-box rmse
    0.15; with gamma 2.0 the bias is -5, with ACES + gamma -31, flipped > 30: a real pin of the model, not of the noise."""
    W, H, SPP = 400, 225, 256
    world, cam = scenes.scene_legacy_6_diffuse((W, H))
    acc, _, st = oracle.render(oracle.scene_from_world(world), cam.to_struct(), W, H, SPP, 100, L.PT_SHADE_LEGACY_STAGE6, seed=1,
                               absorptivity=0.5)
    gold = np.asarray(Image.open(os.path.join(GOLDEN, f"legacy_6_diffuse_{W}x{H}.png")).convert("RGB"), np.float64)
    img = L.to_uint8(np.power(acc / SPP, np.float32(1 / 2.2)), rounding=True).astype(np.float64)
    d = img - gold
    box = d[:224, :400].reshape(28, 8, 50, 8, 3).mean(axis=(1, 3))
    assert np.sqrt((d**2).mean()) < 1.6 and abs(d.mean()) < 0.08 and np.sqrt((box**2).mean()) < 0.3, \
        (np.sqrt((d**2).mean()), d.mean(), np.sqrt((box**2).mean()))
    assert np.sqrt(((img[::-1] - gold) ** 2).mean()) > 30
    wrong = L.to_uint8(np.sqrt(acc / SPP), rounding=True).astype(np.float64) - gold      # gamma 2.0 instead of 2.2
    assert abs(wrong.mean()) > 3
    assert 1.7 < st.segments / st.paths < 1.9


def test_legacy_7_reflect_png_sky_rows_pin_the_camera_and_gamma(oracle):
    """legacy/PT_in_one_weekend/7_reflect.png was NOT rendered from 7_reflect.py as committed (its horizon lies 24.5 rows
    below the image centre; the script's level camera over a radius-10000 ground puts it at the centre; the ground is
    diffuse there, metallic in the script), so its spheres cannot pin sample_reflect / cal_reflectivity_*.  Its sky rows
    still pin the legacy camera at fov 45 (HALF angle: view width 2 tan 45 = 2), the sky gradient, gamma 2.2 and the
    rounding cast: the top 40 rows are reproduced to <= 1 level."""
    W, H, SPP = 400, 225, 16
    world, cam = scenes.scene_legacy_7_reflect((W, H))
    acc, _, _ = oracle.render(oracle.scene_from_world(world), cam.to_struct(), W, H, SPP, 100, L.PT_SHADE_LEGACY_STAGE7, seed=1,
                              absorptivity=0.5)
    gold = np.asarray(Image.open(os.path.join(GOLDEN, f"legacy_7_reflect_{W}x{H}.png")).convert("RGB"), np.float64)
    img = L.to_uint8(np.power(acc / SPP, np.float32(1 / 2.2)), rounding=True).astype(np.float64)
    d = (img - gold)[:40]
    assert np.abs(d).max() <= 1 and (d == 0).mean() > 0.8, (np.abs(d).max(), (d == 0).mean())


def test_legacy_stage7_scattering_statistics(oracle):
    """PT_SHADE_LEGACY_STAGE7 (7_reflect.py:187-209) on the script's own scene: energy bookkeeping that follows from the
    source alone — a metallic hit multiplies by F >= albedo, a diffuse hit by albedo * absorptivity, nothing is added, so the
    image is bounded by the sky and darker with more absorption; the perfect mirror (roughness 0) sphere shows no noise at
    its centre beyond the sky's own variation."""
    W, H = 200, 112
    world, cam = scenes.scene_legacy_7_reflect((W, H))
    sc = oracle.scene_from_world(world)
    a, _, sa = oracle.render(sc, cam.to_struct(), W, H, 32, 100, L.PT_SHADE_LEGACY_STAGE7, seed=1, absorptivity=0.5)
    b, _, sb = oracle.render(sc, cam.to_struct(), W, H, 32, 100, L.PT_SHADE_LEGACY_STAGE7, seed=1, absorptivity=0.25)
    assert a.max() / 32 <= 1.0 + 1e-5 and a.min() >= 0
    assert b.mean() < a.mean() and sa.segments == sb.segments      # same paths, less throughput
