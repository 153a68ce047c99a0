"""ctypes binding of include/dodrt_host.h (libdodrt_host.so): the host side above the ray-query path.

``HostScene`` mirrors the reference's registration order -- ``add_mesh* -> add_*shapes -> build_tree()``
(main.cpp:364-368) -- and ``upload()`` hands the finished arrays to the CUDA library through the C ABI
(``dodrt_scene_set_*``), exactly what the adapter in INTEGRATION.md does on the C++ side.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

from . import capi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DODRT_HOST_LIB") or os.path.join(_HERE, "lib", "libdodrt_host.so")
BUILD_KEEP_CREATION_ORDER = 1  # DODRT_HOST_BUILD_KEEP_CREATION_ORDER (include/dodrt_host.h)


class Config(C.Structure):
    """``dodrt_host_config`` = the reference's Config (config.h:4-14)."""

    _fields_ = [("height", C.c_uint32), ("width", C.c_uint32), ("epsilon", C.c_float), ("frustrum_max", C.c_float),
                ("intersect_cost", C.c_uint32), ("traversal_cost", C.c_uint32), ("empty_bonus", C.c_float),
                ("max_prims", C.c_uint32)]


class Sizes(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in ("num_triangles", "num_orig_lanes", "num_nodes", "num_lanes", "max_depth",
                                          "num_spheres", "num_planes", "num_cylinders", "num_boxes")]


_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: run __graft_entry__.build()")
        lib = C.CDLL(LIB_PATH)
        lib.dodrt_host_last_error.restype = C.c_char_p
        lib.dodrt_host_epsilon.restype = C.c_float
        lib.dodrt_host_num_meshes.restype = C.c_uint32
        for name in ("nodes", "tri_lanes", "prim_nums", "bounds", "tri_normals", "sphere_lanes", "sphere_colors",
                     "plane_lanes", "plane_colors", "cylinders", "box_lanes", "tri_attributes", "mesh_colors"):
            getattr(lib, f"dodrt_host_{name}").restype = C.c_void_p
        lib.dodrt_host_add_cylinder.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_void_p, C.c_void_p]
        lib.dodrt_host_add_sphere.argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.c_void_p]
        _lib = lib
    return _lib


def _check(rc: int):
    if rc != 0:
        raise RuntimeError(f"dodrt_host error {rc}: {load().dodrt_host_last_error().decode(errors='replace')}")


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def default_config() -> Config:
    cfg = Config()
    load().dodrt_host_config_defaults(C.byref(cfg))
    return cfg


def load_config(path: str) -> Config:
    cfg = Config()
    _check(load().dodrt_host_config_load(path.encode(), C.byref(cfg)))
    return cfg


def ray_tables(width: int, height: int):
    xs, ys = np.empty(width, np.float32), np.empty(height, np.float32)
    _check(load().dodrt_host_ray_tables(C.c_uint32(width), C.c_uint32(height), _ptr(xs), _ptr(ys)))
    return xs, ys


def standin_dragon(n: int = 660):
    """Deterministic stand-in for the missing assets/dragon.obj: (positions [(n+1)^2,3], indices [2n^2,3])."""
    pos = np.empty(((n + 1) * (n + 1), 3), np.float32)
    idx = np.empty((2 * n * n, 3), np.uint32)
    _check(load().dodrt_host_standin_dragon(C.c_uint32(n), _ptr(pos), _ptr(idx)))
    return pos, idx


def write_dodm(path: str, positions: np.ndarray, indices: np.ndarray):
    positions = np.ascontiguousarray(positions, np.float32).reshape(-1, 3)
    indices = np.ascontiguousarray(indices, np.uint32).reshape(-1, 3)
    _check(load().dodrt_host_write_dodm(path.encode(), _ptr(positions), C.c_uint32(len(positions)), _ptr(indices),
                                        C.c_uint32(len(indices))))


def _view(ptr, shape, dtype):
    n = int(np.prod(shape))
    if n == 0 or not ptr:
        return np.zeros(shape, dtype)
    buf = (C.c_char * (n * np.dtype(dtype).itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype).reshape(shape).copy()


class HostScene:
    def __init__(self, config: Optional[Config] = None):
        self._lib = load()
        self._h = C.c_void_p()
        _check(self._lib.dodrt_host_scene_create(C.byref(config) if config is not None else None, C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None):
            self._lib.dodrt_host_scene_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- registration ------------------------------------------------------------------------------
    @staticmethod
    def _xf(scale, translate):
        if scale is None and translate is None:
            return None
        t = translate if translate is not None else (0.0, 0.0, 0.0)
        return np.array([1.0 if scale is None else scale, t[0], t[1], t[2]], np.float32)

    def add_mesh(self, positions, indices, scale=None, translate=None):
        positions = np.ascontiguousarray(positions, np.float32).reshape(-1, 3)
        indices = np.ascontiguousarray(indices, np.uint32).reshape(-1, 3)
        xf = self._xf(scale, translate)
        _check(self._lib.dodrt_host_add_mesh(self._h, _ptr(positions), C.c_uint32(len(positions)), _ptr(indices),
                                             C.c_uint32(len(indices)), _ptr(xf)))

    def add_mesh_file(self, path: str, scale=None, translate=None):
        xf = self._xf(scale, translate)
        _check(self._lib.dodrt_host_add_mesh_file(self._h, path.encode(), _ptr(xf)))

    def add_sphere(self, pos, radius, color=(0, 0, 0)):
        pos, color = np.asarray(pos, np.float32), np.asarray(color, np.float32)
        _check(self._lib.dodrt_host_add_sphere(self._h, _ptr(pos), C.c_float(radius), _ptr(color)))

    def add_box(self, lo, hi):
        lo, hi = np.asarray(lo, np.float32), np.asarray(hi, np.float32)
        _check(self._lib.dodrt_host_add_box(self._h, _ptr(lo), _ptr(hi)))

    def add_reference_scene(self, seed: int = 1, num_spheres: int = 16):
        """srand(seed); generateSpheres(16); generatePlanes(); generateCylinders()  (main.cpp:364-366)"""
        _check(self._lib.dodrt_host_add_reference_scene(self._h, C.c_uint32(seed), C.c_uint32(num_spheres)))

    def add_analytic_scene(self, seed: int = 4, count: int = 10000):
        _check(self._lib.dodrt_host_add_analytic_scene(self._h, C.c_uint32(seed), C.c_uint32(count)))

    def build_tree(self, keep_creation_order: bool = False):
        """KDTree::buildTree() (kdtree.cpp:252-260).  By default incl. the lane re-order (triangle.cpp:349-367);
        with `keep_creation_order` the lanes stay as created and `upload` lets the GPU do the re-order
        (dodrt_scene_set_kdtree_indexed)."""
        _check(self._lib.dodrt_host_build_tree_ex(self._h, C.c_uint32(BUILD_KEEP_CREATION_ORDER if keep_creation_order else 0)))
        self._creation_order = bool(keep_creation_order)

    # ---- export ---------------------------------------------------------------------------------------
    def sizes(self) -> Sizes:
        z = Sizes()
        _check(self._lib.dodrt_host_sizes_get(self._h, C.byref(z)))
        return z

    def arrays(self, normals: bool = False, raw: bool = False) -> dict:
        """Copies of the exported arrays.  After build_tree(keep_creation_order=True) the per-lane arrays are gathered
        through prim_nums here (numpy) so that callers always see the re-ordered scene; `raw` returns them as stored."""
        z = self.sizes()
        L = self._lib
        out = dict(
            nodes=_view(L.dodrt_host_nodes(self._h), (z.num_nodes,), np.uint64),
            # re-ordered lanes, or (build_tree(keep_creation_order=True)) the lanes as created
            tri_lanes=_view(L.dodrt_host_tri_lanes(self._h),
                            (z.num_orig_lanes if getattr(self, "_creation_order", False) else z.num_lanes, 72), np.float32),
            prim_nums=_view(L.dodrt_host_prim_nums(self._h), (z.num_lanes if z.num_nodes else 0,), np.uint32),
            bounds=_view(L.dodrt_host_bounds(self._h), (6,), np.float32),
            sphere_lanes=_view(L.dodrt_host_sphere_lanes(self._h), ((z.num_spheres + 7) // 8, 4, 8), np.float32),
            sphere_colors=_view(L.dodrt_host_sphere_colors(self._h), (z.num_spheres, 3), np.float32),
            plane_lanes=_view(L.dodrt_host_plane_lanes(self._h), ((z.num_planes + 7) // 8, 6, 8), np.float32),
            plane_colors=_view(L.dodrt_host_plane_colors(self._h), (z.num_planes, 3), np.float32),
            cylinders=_view(L.dodrt_host_cylinders(self._h), (z.num_cylinders,), capi.CYL_DT),
            box_lanes=_view(L.dodrt_host_box_lanes(self._h), ((z.num_boxes + 7) // 8, 6, 8), np.float32),
            epsilon=float(L.dodrt_host_epsilon(self._h)),
            num_spheres=z.num_spheres, num_planes=z.num_planes, num_boxes=z.num_boxes, max_depth=z.max_depth,
            num_triangles=z.num_triangles,
        )
        if normals:
            nl = z.num_orig_lanes if getattr(self, "_creation_order", False) else z.num_lanes
            out["tri_normals"] = _view(L.dodrt_host_tri_normals(self._h), (nl * 8, 9), np.float32)
            out["tri_attributes"] = _view(L.dodrt_host_tri_attributes(self._h), (nl, 80), np.uint32)
            out["mesh_colors"] = _view(L.dodrt_host_mesh_colors(self._h), (L.dodrt_host_num_meshes(self._h), 3), np.float32)
        if getattr(self, "_creation_order", False) and not raw and z.num_nodes:
            pn = out["prim_nums"]
            out["tri_lanes"] = out["tri_lanes"][pn]
            if normals:
                out["tri_normals"] = out["tri_normals"].reshape(-1, 8, 9)[pn].reshape(-1, 9)
                out["tri_attributes"] = out["tri_attributes"][pn]
        return out

    def upload(self, device: int = 0, shading: bool = False) -> capi.Scene:
        """Copy the finished scene into one GPU through the C ABI (zero-copy from the host library's arrays);
        `shading` also uploads normals / colours for dodrt_render."""
        z = self.sizes()
        L = self._lib
        g = capi.Scene(device)
        cl = capi.load()
        indexed = getattr(self, "_creation_order", False)
        if z.num_nodes and indexed:
            capi._check(cl.dodrt_scene_set_kdtree_indexed(
                g._h, C.c_void_p(L.dodrt_host_nodes(self._h)), C.c_uint32(z.num_nodes),
                C.c_void_p(L.dodrt_host_tri_lanes(self._h)), C.c_uint32(z.num_orig_lanes),
                C.c_void_p(L.dodrt_host_prim_nums(self._h)), C.c_uint32(z.num_lanes), C.c_void_p(L.dodrt_host_bounds(self._h))))
        elif z.num_nodes:
            capi._check(cl.dodrt_scene_set_kdtree(g._h, C.c_void_p(L.dodrt_host_nodes(self._h)), C.c_uint32(z.num_nodes),
                                                  C.c_void_p(L.dodrt_host_tri_lanes(self._h)), C.c_uint32(z.num_lanes),
                                                  C.c_void_p(L.dodrt_host_bounds(self._h))))
        if z.num_spheres:
            capi._check(cl.dodrt_scene_set_spheres(g._h, C.c_void_p(L.dodrt_host_sphere_lanes(self._h)),
                                                   C.c_uint32(z.num_spheres)))
        if z.num_planes:
            capi._check(cl.dodrt_scene_set_planes(g._h, C.c_void_p(L.dodrt_host_plane_lanes(self._h)),
                                                  C.c_uint32(z.num_planes)))
        if z.num_cylinders:
            capi._check(cl.dodrt_scene_set_cylinders(g._h, C.c_void_p(L.dodrt_host_cylinders(self._h)),
                                                     C.c_uint32(z.num_cylinders)))
        if z.num_boxes:
            capi._check(cl.dodrt_scene_set_boxes(g._h, C.c_void_p(L.dodrt_host_box_lanes(self._h)), C.c_uint32(z.num_boxes)))
        g.set_epsilon(float(L.dodrt_host_epsilon(self._h)))
        if shading and indexed:
            capi._check(cl.dodrt_scene_set_shading_indexed(
                g._h, C.c_void_p(L.dodrt_host_tri_attributes(self._h)), C.c_uint32(z.num_orig_lanes),
                C.c_void_p(L.dodrt_host_prim_nums(self._h)), C.c_uint32(z.num_lanes),
                C.c_void_p(L.dodrt_host_mesh_colors(self._h)), C.c_uint32(L.dodrt_host_num_meshes(self._h)),
                C.c_void_p(L.dodrt_host_sphere_colors(self._h)), C.c_void_p(L.dodrt_host_plane_colors(self._h))))
        elif shading:
            capi._check(cl.dodrt_scene_set_shading(
                g._h, C.c_void_p(L.dodrt_host_tri_attributes(self._h)), C.c_uint32(z.num_lanes),
                C.c_void_p(L.dodrt_host_mesh_colors(self._h)), C.c_uint32(L.dodrt_host_num_meshes(self._h)),
                C.c_void_p(L.dodrt_host_sphere_colors(self._h)), C.c_void_p(L.dodrt_host_plane_colors(self._h))))
        return g
