"""ctypes wrapper of oracle/libptoracle.so — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) may import this.
It reuses the struct layouts of include/pt_api.h through learn_path_tracing_b200._lib.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from learn_path_tracing_b200 import _lib as L

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libptoracle.so")


class OrcMesh(C.Structure):
    _fields_ = [("pos", C.c_void_p), ("nrm", C.c_void_p), ("uv", C.c_void_p), ("faces", C.c_void_p),
                ("nf", C.c_int32), ("n_nodes", C.c_int32), ("node_left", C.c_void_p), ("node_right", C.c_void_p),
                ("node_low", C.c_void_p), ("node_high", C.c_void_p), ("node_data", C.c_void_p),
                ("leaf_cut", C.c_void_p), ("max_depth", C.c_int32), ("_pad", C.c_int32)]


class OrcScene(C.Structure):
    _fields_ = [("sph_cr", C.c_void_p), ("sph_mat", C.c_void_p), ("n_sph", C.c_int32), ("n_tsph", C.c_int32),
                ("tsph_cr", C.c_void_p), ("tsph_transparency", C.c_void_p), ("tsph_tex", C.c_void_p),
                ("meshes", C.c_void_p), ("n_mesh", C.c_int32), ("tex_W", C.c_int32), ("texels", C.c_void_p),
                ("tex_areas", C.c_void_p), ("tex_flags", C.c_void_p), ("tex_H", C.c_int32), ("ntex", C.c_int32), ("env", C.c_void_p),
                ("env_W", C.c_int32), ("env_H", C.c_int32), ("env_area", C.c_int32 * 4)]


_lib = None


def build(force=False):
    if force or not os.path.exists(LIB_PATH):
        subprocess.check_call(["make", "-C", _HERE, "libptoracle.so"], stdout=subprocess.DEVNULL)
    return LIB_PATH


def load():
    global _lib
    if _lib is None:
        build()
        lib = C.CDLL(LIB_PATH)
        lib.orc_render.restype = C.c_int
        lib.orc_num_threads.restype = C.c_int
        _lib = lib
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f32(a, shape=None):
    a = np.ascontiguousarray(a, np.float32)
    return a.reshape(shape) if shape else a


def num_threads() -> int:
    return int(load().orc_num_threads())


def rng4(pixel, sample, stream, seed):
    out = (C.c_float * 4)()
    load().orc_rng4(C.c_uint32(pixel), C.c_uint32(sample), C.c_uint32(stream), C.c_uint32(seed), out)
    return np.array(out[:], np.float32)


def generate_rays(cam: L.PtCamera, width, height, sample, seed):
    rays = np.empty((height * width, 8), np.float32)
    load().orc_generate_rays(C.byref(cam), C.c_int(width), C.c_int(height), C.c_int(sample), C.c_uint32(seed), _p(rays))
    return rays


def trace_spheres(center_radius, materials, rays, want_t64=False):
    cr = _f32(center_radius, (-1, 4))
    mats = np.ascontiguousarray(materials, L.MATERIAL_DTYPE)
    rays = _f32(rays, (-1, 8))
    n = rays.shape[0]
    ids = np.empty(n, np.int32)
    t = np.empty(n, np.float32)
    t64 = np.empty(n, np.float64) if want_t64 else None
    load().orc_trace_spheres(_p(cr), _p(mats), C.c_int(cr.shape[0]), _p(rays), C.c_int64(n), _p(ids), _p(t), _p(t64))
    return (ids, t, t64) if want_t64 else (ids, t)


def trace_triangles(tris9, rays, want_second=False):
    tris = _f32(tris9, (-1, 9))
    rays = _f32(rays, (-1, 8))
    n = rays.shape[0]
    ids = np.empty(n, np.int32)
    t = np.empty(n, np.float32)
    t2 = np.empty(n, np.float32) if want_second else None
    load().orc_trace_triangles(_p(tris), C.c_int64(tris.shape[0]), _p(rays), C.c_int64(n), _p(ids), _p(t), _p(t2))
    return (ids, t, t2) if want_second else (ids, t)


def trace_bvh2(nodes16, tris9, rays):
    nodes = _f32(nodes16, (-1, 16))
    tris = _f32(tris9, (-1, 9))
    rays = _f32(rays, (-1, 8))
    n = rays.shape[0]
    ids = np.empty(n, np.int32)
    t = np.empty(n, np.float32)
    counts = (C.c_uint64 * 2)()
    load().orc_trace_bvh2(_p(nodes), C.c_int64(nodes.shape[0]), _p(tris), C.c_int64(tris.shape[0]), _p(rays),
                          C.c_int64(n), _p(ids), _p(t), counts)
    return ids, t, (int(counts[0]), int(counts[1]))


def trace_bvh4(wnodes32, tris9, rays):
    """Walk of a 4-wide tree (csrc/bvh4.h layout) with the reference triangle test -> (ids, t, (nodes, tests))."""
    nodes = _f32(wnodes32, (-1, 32))
    tris = _f32(tris9, (-1, 9))
    rays = _f32(rays, (-1, 8))
    n = rays.shape[0]
    ids = np.empty(n, np.int32)
    t = np.empty(n, np.float32)
    counts = (C.c_uint64 * 2)()
    load().orc_trace_bvh4(_p(nodes), C.c_int64(nodes.shape[0]), _p(tris), C.c_int64(tris.shape[0]), _p(rays),
                          C.c_int64(n), _p(ids), _p(t), counts)
    return ids, t, (int(counts[0]), int(counts[1]))


_bvh4 = None


def bvh4_collapse(nodes16):
    """The PRODUCT's host collapse (csrc/bvh4.h) compiled for the CPU (oracle/bvh4_host.cpp): BVH2 nodes -> wide nodes."""
    global _bvh4
    if _bvh4 is None:
        _bvh4 = C.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "libbvh4host.so"))
        _bvh4.bvh4_collapse_host.restype = C.c_int64
    nodes = _f32(nodes16, (-1, 16))
    out = np.zeros((nodes.shape[0] + 1, 32), np.float32)
    m = _bvh4.bvh4_collapse_host(_p(nodes), C.c_int64(nodes.shape[0]), _p(out), C.c_int64(out.shape[0]))
    if m < 0:
        raise RuntimeError("bvh4_collapse_host: output buffer too small")
    return out[:m].copy()


def triangle_eval(tris9, ids, rays):
    tris = _f32(tris9, (-1, 9))
    rays = _f32(rays, (-1, 8))
    ids = np.ascontiguousarray(ids, np.int32)
    n = rays.shape[0]
    t = np.empty(n, np.float32)
    wmin = np.empty(n, np.float32)
    load().orc_triangle_eval(_p(tris), _p(ids), _p(rays), C.c_int64(n), _p(t), _p(wmin))
    return t, wmin


def random_triangles(n, seed, edge_scale):
    out = np.empty((n, 9), np.float32)
    load().orc_random_triangles(C.c_int64(n), C.c_uint32(seed), C.c_float(edge_scale), _p(out))
    return out


def random_rays(n, seed):
    out = np.empty((n, 8), np.float32)
    load().orc_random_rays(C.c_int64(n), C.c_uint32(seed), _p(out))
    return out


class Scene:
    """Host scene description for the oracle; keeps every numpy array alive."""

    def __init__(self):
        self.c = OrcScene()
        self._keep = []
        self._meshes = []

    def _hold(self, a):
        self._keep.append(a)
        return _p(a)

    def set_spheres(self, center_radius, materials):
        cr = _f32(center_radius, (-1, 4))
        mats = np.ascontiguousarray(materials, L.MATERIAL_DTYPE)
        self.c.sph_cr, self.c.sph_mat, self.c.n_sph = self._hold(cr), self._hold(mats), cr.shape[0]
        return self

    def set_textured_spheres(self, center_radius, transparency, texture_id):
        cr = _f32(center_radius, (-1, 4))
        self.c.tsph_cr = self._hold(cr)
        self.c.tsph_transparency = self._hold(np.ascontiguousarray(transparency, np.int32))
        self.c.tsph_tex = self._hold(np.ascontiguousarray(texture_id, np.int32))
        self.c.n_tsph = cr.shape[0]
        return self

    def add_mesh(self, positions, normals, texcoords, faces, tree=None):
        """tree (optional) = dict(left, right, low, high, data, leaf_cut, max_depth) — the stored SAH tree."""
        m = OrcMesh()
        m.pos = self._hold(_f32(positions, (-1, 3)))
        m.nrm = self._hold(_f32(normals, (-1, 3)))
        m.uv = self._hold(_f32(texcoords, (-1, 2)))
        f = np.ascontiguousarray(faces, np.int32).reshape(-1, 10)
        m.faces, m.nf = self._hold(f), f.shape[0]
        if tree is not None:
            m.n_nodes = len(tree["left"])
            m.node_left = self._hold(np.ascontiguousarray(tree["left"], np.int32))
            m.node_right = self._hold(np.ascontiguousarray(tree["right"], np.int32))
            m.node_low = self._hold(_f32(tree["low"], (-1, 3)))
            m.node_high = self._hold(_f32(tree["high"], (-1, 3)))
            m.node_data = self._hold(np.ascontiguousarray(tree["data"], np.int32))
            m.leaf_cut = self._hold(np.ascontiguousarray(tree["leaf_cut"], np.int32))
            m.max_depth = int(tree.get("max_depth", 16))
        self._meshes.append(m)
        arr = (OrcMesh * len(self._meshes))(*self._meshes)
        self._mesh_arr = arr
        self.c.meshes = C.cast(arr, C.c_void_p)
        self.c.n_mesh = len(self._meshes)
        return self

    def set_texture_atlas(self, texels, areas, flags=None):
        tx = np.ascontiguousarray(texels, np.uint8)
        ar = np.ascontiguousarray(areas, np.int32).reshape(-1, 4)
        fl = np.zeros(ar.shape[0], np.int32) if flags is None else np.ascontiguousarray(flags, np.int32)
        self.c.texels, self.c.tex_W, self.c.tex_H = self._hold(tx), tx.shape[0], tx.shape[1]
        self.c.tex_areas, self.c.ntex = self._hold(ar), ar.shape[0]
        self.c.tex_flags = self._hold(fl)
        return self

    def set_environment(self, rgb, area=None):
        if rgb is None:
            self.c.env = None
            return self
        e = _f32(rgb)
        self.c.env, self.c.env_W, self.c.env_H = self._hold(e), e.shape[0], e.shape[1]
        ar = area if area is not None else [0, 0, e.shape[0], e.shape[1]]
        for i in range(4):
            self.c.env_area[i] = int(ar[i])
        return self

    def trace(self, rays):
        rays = _f32(rays, (-1, 8))
        n = rays.shape[0]
        ids = np.empty(n, np.int32)
        t = np.empty(n, np.float32)
        load().orc_trace_legacy(C.byref(self.c), _p(rays), C.c_int64(n), _p(ids), _p(t))
        return ids, t


def render(scene: Scene, cam: L.PtCamera, width, height, spp, max_depth, shading_model=L.PT_SHADE_V2, seed=1,
           spp_offset=0, absorptivity=0.25, want_sq=False, threads=0, flags=0, accum=None, accum_sq=None):
    """Returns (sum [W,H,3], sum of squares or None, PtStats).  Pass accum/accum_sq to keep adding."""
    p = L.PtRenderParams()
    p.width, p.height, p.spp, p.spp_offset = width, height, spp, spp_offset
    p.max_depth, p.shading_model, p.seed, p.absorptivity, p.flags = max_depth, shading_model, seed, absorptivity, flags
    if accum is None:
        accum = np.zeros((width, height, 3), np.float32)
    if want_sq and accum_sq is None:
        accum_sq = np.zeros((width, height, 3), np.float32)
    st = L.PtStats()
    rc = load().orc_render(C.byref(scene.c), C.byref(cam), C.byref(p), _p(accum), _p(accum_sq), C.byref(st),
                           C.c_int(threads))
    if rc != 0:
        raise RuntimeError(f"orc_render failed: {rc}")
    return accum, accum_sq, st


def postprocess(accum, scale, aces=True, gamma=2.2):
    a = _f32(accum)
    w, h = a.shape[0], a.shape[1]
    out = np.empty_like(a)
    load().orc_postprocess(_p(a), C.c_int(w), C.c_int(h), C.c_float(scale), C.c_int(int(aces)), C.c_float(gamma), _p(out))
    return out


def scene_from_world(world) -> Scene:
    cr, mats = world.arrays()
    return Scene().set_spheres(cr, mats)


def scene_from_legacy_world(world, use_stored_tree=True) -> Scene:
    """legacy.World -> oracle scene.  With use_stored_tree the oracle walks the reference's own SAH tree from the
    .world.npy exactly as MeshBVHTree.hit does (15_module.py:756-779); otherwise it loops over every face."""
    sc = Scene()
    if world.spheres:
        sc.set_textured_spheres(*world.sphere_arrays())
    for m in world.meshes:
        sc.add_mesh(m["positions"], m["normals"], m["texture_coords"], m["indices"],
                    tree=m["tree"] if use_stored_tree else None)
    if world._atlas is not None:
        sc.set_texture_atlas(*world._atlas)
    if world._env is not None:
        sc.set_environment(world._env[0], world._env[1])
    return sc
