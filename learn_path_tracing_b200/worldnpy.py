""".world.npy scene caches of the legacy tracer (World.save / World.load, 15_module.py:815-836).

np.save(filename, dict) -> NPY v1.0, dtype object, body = pickle.  Three of the reference's six files
embed taichi.lang.struct.Struct / taichi.lang.matrix.Matrix instances (the texture `area`s), so plain
np.load needs Taichi; this loader maps every taichi.* class to a stub and numpy.core.* to numpy._core.*
(the files were written by numpy 1.x).  Schema: SURVEY section 2.3.
"""
from __future__ import annotations

import pickle

import numpy as np
import numpy.lib.format as npfmt


class _Stub:
    """Stands in for taichi.lang.struct.Struct / taichi.lang.matrix.Matrix while unpickling."""

    def __init__(self, *a, **k):
        pass

    def __setstate__(self, state):
        self.__dict__.update(state if isinstance(state, dict) else {"_state": state})


# what a .world.npy may name besides taichi.* (stubbed): numpy's array/dtype/scalar reconstructors and plain containers.
# Anything else is refused — np.load(allow_pickle=True), which the reference uses, would execute it.
_ALLOWED = {
    ("numpy._core.multiarray", "_reconstruct"), ("numpy._core.multiarray", "scalar"), ("numpy", "ndarray"), ("numpy", "dtype"),
    ("numpy._core.numeric", "_frombuffer"), ("builtins", "dict"), ("builtins", "list"), ("builtins", "tuple"),
    ("builtins", "set"), ("builtins", "frozenset"), ("builtins", "int"), ("builtins", "float"), ("builtins", "str"),
    ("builtins", "bytes"), ("builtins", "bool"), ("builtins", "complex"), ("builtins", "slice"), ("collections", "OrderedDict"),
}


class _Unpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module.startswith("taichi"):
            return type(name, (_Stub,), {"__module__": module})
        if module.startswith("numpy.core"):
            module = module.replace("numpy.core", "numpy._core")
        if (module, name) in _ALLOWED or (module == "numpy" and name in np.sctypeDict):
            return super().find_class(module, name)
        raise pickle.UnpicklingError(f".world.npy names {module}.{name}: only numpy arrays, plain containers and taichi "
                                     "structs are loaded")


def _plain(x):
    """taichi Struct/Matrix stubs -> dicts/lists."""
    if isinstance(x, _Stub):
        d = x.__dict__
        if "entries" in d:
            e = d["entries"]
            return {k: _plain(v) for k, v in e.items()} if isinstance(e, dict) else [_plain(v) for v in e]
        return {k: _plain(v) for k, v in d.items()}
    if isinstance(x, dict):
        return {k: _plain(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return type(x)(_plain(v) for v in x)
    return x


def load_world(filename) -> dict:
    """The dict World.save wrote, with Taichi objects replaced by plain dicts/lists."""
    with open(filename, "rb") as f:
        npfmt.read_magic(f)
        npfmt.read_array_header_1_0(f)
        x = _Unpickler(f).load()
    if isinstance(x, np.ndarray):
        x = x.item()
    return _plain(x)


def _area(cfg):
    a = cfg["area"]
    low, high = a["low"], a["high"]
    return [int(low[0]), int(low[1]), int(high[0]), int(high[1])]


def texture_configs(manager_dump) -> list[dict]:
    """[{file_path, id, size, area=[lx,ly,hx,hy]}] from a TextureManager.dump() (15_module.py:494-495)."""
    out = []
    for c in manager_dump.get("configs", []):
        out.append({"file_path": c["file_path"], "id": int(c["id"]), "size": tuple(int(v) for v in c["size"]),
                    "area": _area(c)})
    return out


def mesh_arrays(mesh_dict) -> dict:
    """One meshes_bvhs entry -> numpy arrays: positions/normals/texcoords, faces [F,10] in leaf order and the
    stored SAH tree (left, right, low, high, data, leaf_cut, max_depth)."""
    nodes = mesh_dict["tree_nodes_field"]["data"]
    leaves = mesh_dict["tree_leaves_field"]["data"]
    faces = np.stack([leaves["a"]["p"], leaves["a"]["n"], leaves["a"]["t"], leaves["b"]["p"], leaves["b"]["n"],
                      leaves["b"]["t"], leaves["c"]["p"], leaves["c"]["n"], leaves["c"]["t"], leaves["texture_id"]],
                     axis=1).astype(np.int32)
    return {
        "positions": np.asarray(mesh_dict["positions_field"]["data"], np.float32).reshape(-1, 3),
        "normals": np.asarray(mesh_dict["normals_field"]["data"], np.float32).reshape(-1, 3),
        "texcoords": np.asarray(mesh_dict["texture_coords_field"]["data"], np.float32).reshape(-1, 2),
        "faces": faces,
        "tree": {
            "left": np.asarray(nodes["left"], np.int32), "right": np.asarray(nodes["right"], np.int32),
            "low": np.asarray(nodes["aabb"]["low"], np.float32).reshape(-1, 3),
            "high": np.asarray(nodes["aabb"]["high"], np.float32).reshape(-1, 3),
            "data": np.asarray(nodes["data"], np.int32),
            "leaf_cut": np.asarray(mesh_dict["tree_leaves_field_cut"]["data"], np.int32),
            "max_depth": int(mesh_dict["max_depth"]),
        },
    }


def sphere_arrays(bvh_dict) -> dict:
    """spheres_bvh entry -> center_radius [S,4], transparency [S], texture_id [S] (leaf order)."""
    leaves = bvh_dict["tree_leaves_field"]["data"]
    c = np.asarray(leaves["center"], np.float32).reshape(-1, 3)
    r = np.asarray(leaves["radius"], np.float32).reshape(-1, 1)
    return {"center_radius": np.concatenate([c, r], axis=1), "transparency": np.asarray(leaves["transparency"], np.int32),
            "texture_id": np.asarray(leaves["texture_id"], np.int32)}


def save_world(filename, meshes: list[dict], environment: int, textures: dict | None = None,
               environments: dict | None = None, spheres: dict | None = None):
    """Writes the reference's schema back (World.save, 15_module.py:815-821) with plain dict areas, so files
    interoperate both ways.  `meshes` entries are mesh_arrays()-shaped dicts (tree optional)."""
    def field(a):
        a = np.asarray(a)
        return {"data": a, "shape": list(a.shape[:1])}

    out = {"meshes_bvhs": [], "environment": int(environment)}
    for m in meshes:
        f = np.asarray(m["faces"], np.int32)
        tree = m.get("tree")
        if tree is None:  # a single leaf holding every face: valid input for the reference's traversal
            lo, hi = m["positions"].min(0), m["positions"].max(0)
            tree = {"left": np.array([-1], np.int32), "right": np.array([-1], np.int32), "low": lo[None], "high": hi[None],
                    "data": np.array([0], np.int32), "leaf_cut": np.array([0, len(f)], np.int32), "max_depth": 16}
        leaves = {k: {"p": f[:, 3 * i], "n": f[:, 3 * i + 1], "t": f[:, 3 * i + 2]} for i, k in enumerate("abc")}
        leaves["texture_id"] = f[:, 9]
        out["meshes_bvhs"].append({
            "max_depth": int(tree["max_depth"]),
            "tree_nodes_field": {"data": {"left": tree["left"], "right": tree["right"],
                                          "aabb": {"low": tree["low"], "high": tree["high"]}, "data": tree["data"]},
                                 "shape": [len(tree["left"])]},
            "tree_leaves_field": {"data": leaves, "shape": [len(f)]},
            "tree_leaves_field_cut": field(tree["leaf_cut"]),
            "positions_field": field(m["positions"]), "normals_field": field(m["normals"]),
            "texture_coords_field": field(m["texcoords"]),
        })
    # World.load (15_module.py:823-836) indexes data['textures'] / data['environments'] unconditionally: always present,
    # as empty TextureManager.dump()-shaped dicts when the world has none
    out["textures"] = textures if textures is not None else {"size": (0, 0), "configs": []}
    out["environments"] = environments if environments is not None else {"size": (0, 0), "configs": []}
    if spheres is not None:
        out["spheres_bvh"] = spheres
    np.save(filename, out, allow_pickle=True)
