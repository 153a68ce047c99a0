"""ctypes access to the two CPU checkers (TEST INFRASTRUCTURE, never imported by the product):

* ``Oracle``  -- oracle/liboracle.so, the plain-C restatement (oracle/dodrt_oracle.c).
* ``RefLib``  -- oracle/_ref/libdodrt_ref.so, the reference's own translation units behind
  oracle/ref_harness.cpp.  The reference keeps every shape in process-global, append-only vectors
  (triangle.h:59-61, sphere.cpp:21-23), so each ``RefLib()`` loads a PRIVATE copy of the library to
  get a fresh set of globals.
"""
from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "liboracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libdodrt_ref.so")

RAY_DT = np.dtype([("o", "<f4", 3), ("d", "<f4", 3), ("clip", "<f4"), ("flags", "<u4")])
HIT_DT = np.dtype([("t", "<f4"), ("prim", "<u4"), ("u", "<f4"), ("v", "<f4")])
CTR_DT = np.dtype([("nodes", "<u4"), ("leaves", "<u4"), ("lanes", "<u4"), ("max_stack", "<u4")])
REC_DT = np.dtype([("t", "<f4"), ("color", "<f4", 3), ("normal", "<f4", 3), ("point", "<f4", 3), ("hit", "<u4")])
CYL_DT = np.dtype([("base", "<f4", 3), ("axis", "<f4", 3), ("radius_sq", "<f4"), ("height", "<f4")])

MISS = 0xFFFFFFFF
KIND_SHIFT = 29
KIND_TRIANGLE, KIND_SPHERE, KIND_PLANE, KIND_CYLINDER, KIND_BOX = range(5)
CLS_SPHERE, CLS_PLANE, CLS_CYLINDER, CLS_TREE, CLS_BOX = 1, 2, 4, 8, 16
RAY_ANY = 1


def ensure_oracle_built() -> None:
    """(Re)build liboracle.so (and _ref when /root/reference exists).  Building the checker is not using it."""
    subprocess.run(["make", "-s", "-C", ORACLE_DIR, "all"], check=True, stdout=subprocess.DEVNULL)


def have_ref() -> bool:
    return os.path.exists(REF_SO)


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class _OrcScene(C.Structure):
    _fields_ = [
        ("nodes", C.c_void_p), ("num_nodes", C.c_uint32),
        ("tri_lanes", C.c_void_p), ("num_tri_lanes", C.c_uint32),
        ("bounds", C.c_float * 6),
        ("sphere_lanes", C.c_void_p), ("num_spheres", C.c_uint32),
        ("plane_lanes", C.c_void_p), ("num_planes", C.c_uint32),
        ("cylinders", C.c_void_p), ("num_cylinders", C.c_uint32),
        ("box_lanes", C.c_void_p), ("num_boxes", C.c_uint32),
        ("epsilon", C.c_float),
    ]


def pack_lanes(cols: np.ndarray) -> np.ndarray:
    """[N, K] float32 per-primitive columns -> reference lane layout [ceil(N/8), K, 8], zero padded
    (sphere.cpp:226-242, plane.cpp:204-222, triangle.cpp:262-292)."""
    cols = np.ascontiguousarray(cols, np.float32)
    n, k = cols.shape
    lanes = np.zeros(((n + 7) // 8, k, 8), np.float32)
    idx = np.arange(n)
    lanes[idx // 8, :, idx % 8] = cols
    return lanes


def sphere_lanes(spheres: np.ndarray) -> np.ndarray:
    """[N,4] x,y,z,radius -> lanes x[8] y[8] z[8] radiusSq[8]; radiusSq = r*r in fp32 (sphere.cpp:238)."""
    s = np.ascontiguousarray(spheres, np.float32).reshape(-1, 4).copy()
    s[:, 3] = s[:, 3] * s[:, 3]
    return pack_lanes(s)


class Scene:
    """Host-side arrays of one scene in the reference's native layouts."""

    def __init__(self, nodes=None, tri_lanes=None, bounds=None, spheres=None, planes=None, cylinders=None,
                 boxes=None, epsilon=1e-4):
        self.nodes = np.ascontiguousarray(nodes if nodes is not None else np.zeros(0, np.uint64), np.uint64)
        self.tri_lanes = np.ascontiguousarray(
            tri_lanes if tri_lanes is not None else np.zeros((0, 72), np.float32), np.float32).reshape(-1, 72)
        self.bounds = np.ascontiguousarray(bounds if bounds is not None else np.zeros(6, np.float32), np.float32)
        self.spheres = np.ascontiguousarray(
            spheres if spheres is not None else np.zeros((0, 4), np.float32), np.float32).reshape(-1, 4)
        self.planes = np.ascontiguousarray(
            planes if planes is not None else np.zeros((0, 6), np.float32), np.float32).reshape(-1, 6)
        self.cylinders = np.ascontiguousarray(cylinders if cylinders is not None else np.zeros(0, CYL_DT), CYL_DT)
        self.boxes = np.ascontiguousarray(
            boxes if boxes is not None else np.zeros((0, 6), np.float32), np.float32).reshape(-1, 6)
        self.epsilon = float(epsilon)
        self.sphere_lanes = sphere_lanes(self.spheres)
        self.plane_lanes = pack_lanes(self.planes)
        self.box_lanes = pack_lanes(self.boxes)

    @classmethod
    def from_host_arrays(cls, a: dict) -> "Scene":
        """Scene from the arrays dod_raytracer_b200.host.HostScene.arrays() exports (already in lane layout)."""
        s = cls.__new__(cls)
        s.nodes = np.ascontiguousarray(a["nodes"], np.uint64)
        s.tri_lanes = np.ascontiguousarray(a["tri_lanes"], np.float32).reshape(-1, 72)
        s.bounds = np.ascontiguousarray(a["bounds"], np.float32)
        s.sphere_lanes, s.plane_lanes, s.box_lanes = a["sphere_lanes"], a["plane_lanes"], a["box_lanes"]
        s.spheres = np.zeros((a["num_spheres"], 4), np.float32)  # only the counts are read below
        s.planes = np.zeros((a["num_planes"], 6), np.float32)
        s.boxes = np.zeros((a["num_boxes"], 6), np.float32)
        s.cylinders = np.ascontiguousarray(a["cylinders"]).view(CYL_DT)
        s.epsilon = float(a["epsilon"])
        return s

    def orc(self) -> _OrcScene:
        s = _OrcScene()
        s.nodes, s.num_nodes = _ptr(self.nodes), len(self.nodes)
        s.tri_lanes, s.num_tri_lanes = _ptr(self.tri_lanes), len(self.tri_lanes)
        s.bounds = (C.c_float * 6)(*[float(x) for x in self.bounds])
        s.sphere_lanes, s.num_spheres = _ptr(self.sphere_lanes), len(self.spheres)
        s.plane_lanes, s.num_planes = _ptr(self.plane_lanes), len(self.planes)
        s.cylinders, s.num_cylinders = _ptr(self.cylinders), len(self.cylinders)
        s.box_lanes, s.num_boxes = _ptr(self.box_lanes), len(self.boxes)
        s.epsilon = self.epsilon
        return s


class Oracle:
    def __init__(self):
        if not os.path.exists(ORACLE_SO):
            ensure_oracle_built()
        self.lib = C.CDLL(ORACLE_SO)

    def intersect(self, scene: Scene, rays: np.ndarray, classes: int, counters=False, nthreads=1):
        rays = np.ascontiguousarray(rays, RAY_DT)
        hits = np.zeros(len(rays), HIT_DT)
        ctrs = np.zeros(len(rays), CTR_DT) if counters else None
        s = scene.orc()
        self.lib.orc_intersect(C.byref(s), _ptr(rays), C.c_uint64(len(rays)), C.c_uint32(classes), _ptr(hits),
                               _ptr(ctrs), C.c_int(nthreads))
        return (hits, ctrs) if counters else hits

    def ray_tables(self, width: int, height: int):
        xs, ys = np.zeros(width, np.float32), np.zeros(height, np.float32)
        self.lib.orc_ray_tables(C.c_uint32(width), C.c_uint32(height), _ptr(xs), _ptr(ys))
        return xs, ys

    def primary_rays(self, width: int, height: int) -> np.ndarray:
        rays = np.zeros(width * height, RAY_DT)
        self.lib.orc_primary_rays(C.c_uint32(width), C.c_uint32(height), _ptr(rays))
        return rays

    def shadow_rays(self, points: np.ndarray, light) -> np.ndarray:
        points = np.ascontiguousarray(points, np.float32).reshape(-1, 3)
        light = np.ascontiguousarray(light, np.float32)
        rays = np.zeros(len(points), RAY_DT)
        for i in range(len(points)):  # small batches only
            self.lib.orc_shadow_ray(_ptr(points[i]), _ptr(light), C.c_void_p(rays.ctypes.data + i * RAY_DT.itemsize))
        return rays

    def trace_primary(self, scene: Scene, width: int, height: int, classes: int, counters=False, nthreads=1):
        n = width * height
        hits = np.zeros(n, HIT_DT)
        ctrs = np.zeros(n, CTR_DT) if counters else None
        s = scene.orc()
        self.lib.orc_trace_primary(C.byref(s), C.c_uint32(width), C.c_uint32(height), C.c_uint32(classes),
                                   _ptr(hits), _ptr(ctrs), C.c_int(nthreads))
        return (hits, ctrs) if counters else hits

    def trace_shadow(self, scene: Scene, width: int, height: int, classes: int, hits: np.ndarray, light,
                     counters=False, nthreads=1):
        n = width * height
        hits = np.ascontiguousarray(hits, HIT_DT)
        light = np.ascontiguousarray(light, np.float32)
        vis = np.zeros(n, np.uint8)
        ctrs = np.zeros(n, CTR_DT) if counters else None
        s = scene.orc()
        self.lib.orc_trace_shadow(C.byref(s), C.c_uint32(width), C.c_uint32(height), C.c_uint32(classes),
                                  _ptr(hits), _ptr(light), _ptr(vis), _ptr(ctrs), C.c_int(nthreads))
        return (vis, ctrs) if counters else vis


class RefLib:
    """A fresh instance of the reference (private copy of the .so => private globals)."""

    _in_place_taken = False

    def __init__(self):
        if not have_ref():
            raise FileNotFoundError(REF_SO)
        if not RefLib._in_place_taken:
            # the first instance of a process maps oracle/_ref/libdodrt_ref.so where it lies (so that whoever audits
            # the process's mappings sees the reference library, not an anonymous copy) ...
            RefLib._in_place_taken = True
            self._tmp = ""
            path = REF_SO
        else:
            # ... every further instance needs its own copy = its own set of the reference's process-global scene
            self._tmp = tempfile.mkdtemp(prefix="dodrt_ref_")
            path = os.path.join(self._tmp, f"libdodrt_ref_{id(self):x}.so")
            shutil.copy(REF_SO, path)
        self.lib = C.CDLL(path)
        self.lib.ref_num_spheres.restype = C.c_uint32
        self.lib.ref_add_sphere.argtypes = [C.c_void_p, C.c_float]
        self.lib.ref_add_mesh.argtypes = [C.c_char_p]

    def __del__(self):
        if getattr(self, "_tmp", ""):
            shutil.rmtree(self._tmp, ignore_errors=True)

    def set_config(self, width, height):
        self.lib.ref_set_config(C.c_uint(width), C.c_uint(height))

    def add_reference_spheres(self, seed=1, count=16):
        self.lib.ref_add_reference_spheres(C.c_uint(seed), C.c_uint(count))

    def add_reference_planes(self):
        self.lib.ref_add_reference_planes()

    def add_reference_cylinder(self):
        self.lib.ref_add_reference_cylinder()

    def add_spheres(self, spheres: np.ndarray):
        spheres = np.ascontiguousarray(spheres, np.float32).reshape(-1, 4)
        for s in spheres:
            pos = np.ascontiguousarray(s[:3])
            self.lib.ref_add_sphere(_ptr(pos), C.c_float(float(s[3])))

    def add_mesh(self, path: str) -> int:
        return self.lib.ref_add_mesh(path.encode())

    def build_tree(self):
        self.lib.ref_build_tree()

    def tree_sizes(self):
        v = [C.c_uint32() for _ in range(5)]
        self.lib.ref_tree_sizes(*[C.byref(x) for x in v])
        return dict(zip(["nodes", "lanes", "prim_nums", "max_depth", "triangles"], [x.value for x in v]))

    def export_tree(self):
        sz = self.tree_sizes()
        nodes = np.zeros(sz["nodes"], np.uint64)
        lanes = np.zeros((sz["lanes"], 72), np.float32)
        prim = np.zeros(sz["prim_nums"], np.uint32)
        bounds = np.zeros(6, np.float32)
        self.lib.ref_tree_export(_ptr(nodes), _ptr(lanes), _ptr(prim), _ptr(bounds))
        return nodes, lanes, prim, bounds

    def export_normals(self):
        sz = self.tree_sizes()
        out = np.zeros((sz["lanes"] * 8, 9), np.float32)
        self.lib.ref_normals_export(_ptr(out))
        return out

    def export_attributes(self):
        """(Triangle::Attributes lanes as uint32 [lanes, 80], mesh colours [meshes, 3])"""
        sz = self.tree_sizes()
        self.lib.ref_num_meshes.restype = C.c_uint32
        attrs = np.zeros((sz["lanes"], 80), np.uint32)
        colors = np.zeros((self.lib.ref_num_meshes(), 3), np.float32)
        self.lib.ref_attrs_export(_ptr(attrs), _ptr(colors))
        return attrs, colors

    def export_spheres(self):
        n = self.lib.ref_num_spheres()
        out = np.zeros((n, 7), np.float32)
        self.lib.ref_spheres_export(_ptr(out))
        return out

    def primary_rays(self, width, height):
        self.set_config(width, height)
        rays = np.zeros(width * height, RAY_DT)
        self.lib.ref_primary_rays(_ptr(rays))
        return rays

    def shadow_rays(self, points, light):
        points = np.ascontiguousarray(points, np.float32).reshape(-1, 3)
        light = np.ascontiguousarray(light, np.float32)
        rays = np.zeros(len(points), RAY_DT)
        self.lib.ref_shadow_rays(_ptr(points), C.c_uint64(len(points)), _ptr(light), _ptr(rays))
        return rays

    def intersect(self, rays, classes, nthreads=1):
        rays = np.ascontiguousarray(rays, RAY_DT)
        hits = np.zeros(len(rays), HIT_DT)
        self.lib.ref_intersect(_ptr(rays), C.c_uint64(len(rays)), C.c_uint32(classes), _ptr(hits), C.c_int(nthreads))
        return hits

    def intersect_records(self, rays, classes, nthreads=1):
        rays = np.ascontiguousarray(rays, RAY_DT)
        recs = np.zeros(len(rays), REC_DT)
        self.lib.ref_intersect_records(_ptr(rays), C.c_uint64(len(rays)), C.c_uint32(classes), _ptr(recs),
                                       C.c_int(nthreads))
        return recs

    def trace_frame(self, width, height, classes, light, nthreads=1):
        """reference per-pixel loop at bounce 0 with one light (ref_trace_frame): (t [n], visible [n])"""
        self.set_config(width, height)
        light = np.ascontiguousarray(light, np.float32)
        t = np.zeros(width * height, np.float32)
        vis = np.zeros(width * height, np.uint8)
        self.lib.ref_trace_frame(C.c_uint32(classes), _ptr(light), _ptr(t), _ptr(vis), C.c_int(nthreads))
        return t, vis

    def render(self, width, height, nthreads=0):
        self.set_config(width, height)
        img = np.zeros((height, width, 3), np.uint8)
        if nthreads <= 1:
            self.lib.ref_render_rows(_ptr(img), C.c_uint(0), C.c_uint(height))
        else:
            self.lib.ref_render_bands(_ptr(img), C.c_int(nthreads))
        return img


# ---- the reference's fixed scene parts, restated as data (main.cpp:52-129) -------------------------
def reference_planes() -> np.ndarray:
    """[6,6] position xyz, normal xyz (main.cpp:54-103)."""
    return np.array([
        [0, 0, 5, 0, 0, -1],
        [0, 0, -5, 0, 0, 1],
        [0, 5, 0, 0, -1, 0],
        [0, -5, 0, 0, 1, 0],
        [-5, 0, 0, 1, 0, 0],
        [5, 0, 0, -1, 0, 0],
    ], np.float32)


def reference_cylinder() -> np.ndarray:
    """main.cpp:113-117 + Cylinder::Cylinder (cylinder.cpp:223-229): axis = glm::normalize((2.2,5,2))."""
    a = np.array([2.2, 5, 2], np.float32)
    d = np.float32(np.float32(a[0] * a[0]) + np.float32(a[1] * a[1])) + np.float32(a[2] * a[2])
    inv = np.float32(1.0) / np.sqrt(np.float32(d))
    c = np.zeros(1, CYL_DT)
    c["base"] = [-2, 0, 2]
    c["axis"] = a * np.float32(inv)
    c["radius_sq"] = np.float32(1.5) * np.float32(1.5)
    c["height"] = 4.0
    return c


def same_bits(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    return np.ascontiguousarray(a, np.float32).view(np.uint32) == np.ascontiguousarray(b, np.float32).view(np.uint32)
