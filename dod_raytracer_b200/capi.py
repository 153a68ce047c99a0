"""ctypes binding of the C ABI in include/dodrt.h (libdodrt_cuda.so).

This is the thinnest possible layer: numpy arrays in the reference's own layouts go in, numpy
arrays come out, every non-zero status raises ``DodrtError`` with ``dodrt_last_error()``.  There is
no Python or CPU implementation of any query behind it -- if the library is missing or no CUDA
device is usable, calls fail loudly.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DODRT_LIB") or os.path.join(_HERE, "lib", "libdodrt_cuda.so")

# include/dodrt.h structs
RAY_DT = np.dtype([("o", "<f4", 3), ("d", "<f4", 3), ("clip", "<f4"), ("flags", "<u4")])
HIT_DT = np.dtype([("t", "<f4"), ("prim", "<u4"), ("u", "<f4"), ("v", "<f4")])
CYL_DT = np.dtype([("base", "<f4", 3), ("axis", "<f4", 3), ("radius_sq", "<f4"), ("height", "<f4")])

MISS = 0xFFFFFFFF
KIND_SHIFT = 29
KIND_TRIANGLE, KIND_SPHERE, KIND_PLANE, KIND_CYLINDER, KIND_BOX = range(5)
CLS_SPHERE, CLS_PLANE, CLS_CYLINDER, CLS_TREE, CLS_BOX = 1, 2, 4, 8, 16
CLS_ALL = 31
RAY_ANY = 1

EXPORTED_SYMBOLS = [
    "dodrt_abi_version", "dodrt_last_error", "dodrt_device_count", "dodrt_kernel_variant_available", "dodrt_scene_debug_stats",
    "dodrt_scene_create", "dodrt_scene_destroy", "dodrt_scene_set_kdtree", "dodrt_scene_set_kdtree_indexed",
    "dodrt_scene_set_shading_indexed", "dodrt_scene_set_spheres",
    "dodrt_scene_set_planes", "dodrt_scene_set_cylinders", "dodrt_scene_set_boxes", "dodrt_scene_set_epsilon",
    "dodrt_scene_set_kernel_variant", "dodrt_scene_set_shading", "dodrt_render",
    "dodrt_intersect", "dodrt_trace_primary", "dodrt_trace_shadow", "dodrt_trace_frame",
    "dodrt_intersect_device", "dodrt_trace_primary_device", "dodrt_trace_shadow_device",
    "dodrt_frame_assemble_device", "dodrt_frame_local_pixels", "dodrt_frame_pixel_map", "dodrt_scene_launch_count",
    "dodrt_trace_frame_device", "dodrt_frame_buffer_create", "dodrt_frame_buffer_export", "dodrt_frame_buffer_open",
    "dodrt_frame_buffer_attach", "dodrt_frame_buffer_pointers", "dodrt_frame_buffer_destroy",
    "dodrt_multi_create", "dodrt_multi_trace_frame", "dodrt_multi_render", "dodrt_multi_destroy",
]
ABI_VERSION = 2


class DodrtError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"dodrt error {code}: {message}")
        self.code = code


class Frame(C.Structure):
    """``dodrt_frame`` (include/dodrt.h)."""

    _fields_ = [
        ("width", C.c_uint32), ("height", C.c_uint32),
        ("tile_w", C.c_uint32), ("tile_h", C.c_uint32),
        ("first_tile", C.c_uint32), ("tile_stride", C.c_uint32),
        ("classes", C.c_uint32), ("compact", C.c_uint32),
        ("origin", C.c_float * 3),
    ]

    @classmethod
    def make(cls, width, height, classes=CLS_ALL, tile=(32, 32), first_tile=0, tile_stride=1, compact=0,
             origin=(0.0, 0.0, -4.9)):
        f = cls()
        f.width, f.height = int(width), int(height)
        f.tile_w, f.tile_h = int(tile[0]), int(tile[1])
        f.first_tile, f.tile_stride = int(first_tile), int(tile_stride)
        f.classes, f.compact = int(classes), int(compact)
        # (0,0,-4.9) as the reference writes it: a double literal narrowed to float (main.cpp:275,308)
        f.origin = (C.c_float * 3)(*[float(np.float32(x)) for x in origin])
        return f


_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    """Load libdodrt_cuda.so (built by ``__graft_entry__.build()`` / ``make -C dod_raytracer_b200/csrc``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(the CUDA path has no fallback)")
    lib = C.CDLL(LIB_PATH)
    lib.dodrt_last_error.restype = C.c_char_p
    for name in EXPORTED_SYMBOLS:
        getattr(lib, name)  # AttributeError if the ABI is incomplete
    if lib.dodrt_abi_version() != ABI_VERSION:
        raise ImportError(f"unexpected dodrt ABI version {lib.dodrt_abi_version()}")
    _lib = lib
    return lib


def _check(rc: int) -> None:
    if rc != 0:
        raise DodrtError(rc, load().dodrt_last_error().decode(errors="replace"))


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    return a.ctypes.data_as(C.c_void_p)


def variant_available(variant: int) -> bool:
    """Does the loaded build hold this kernel variant?  (product: 0, 3, 7; libdodrt_cuda_exp.so: 0-8)"""
    return bool(load().dodrt_kernel_variant_available(C.c_int(variant)))


def experiments_build() -> bool:
    """True for libdodrt_cuda_exp.so (-DDODRT_EXPERIMENTS: A/B variants, one-launch frame kernels, work splitting)."""
    return variant_available(1)


EXP_LIB_PATH = os.path.join(_HERE, "lib", "libdodrt_cuda_exp.so")


def device_count() -> int:
    n = C.c_int(0)
    _check(load().dodrt_device_count(C.byref(n)))
    return n.value


def frame_local_pixels(frame: Frame) -> int:
    n = C.c_uint64(0)
    _check(load().dodrt_frame_local_pixels(C.byref(frame), C.byref(n)))
    return n.value


def frame_pixel_map(frame: Frame) -> np.ndarray:
    n = frame_local_pixels(frame)
    out = np.zeros(n, np.uint32)
    _check(load().dodrt_frame_pixel_map(C.byref(frame), _ptr(out), C.c_uint64(n)))
    return out


class Scene:
    """One scene replica resident on one GPU (``dodrt_scene``)."""

    def __init__(self, device: int = 0):
        self._lib = load()
        self._h = C.c_void_p()
        self.device = device
        _check(self._lib.dodrt_scene_create(C.c_int(device), C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.dodrt_scene_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # ---- registration -------------------------------------------------------------------------
    def set_kdtree(self, nodes: np.ndarray, tri_lanes: np.ndarray, bounds: np.ndarray):
        nodes = np.ascontiguousarray(nodes, np.uint64)
        tri_lanes = np.ascontiguousarray(tri_lanes, np.float32).reshape(-1, 72)
        bounds = np.ascontiguousarray(bounds, np.float32)
        assert bounds.size == 6
        _check(self._lib.dodrt_scene_set_kdtree(self._h, _ptr(nodes), C.c_uint32(len(nodes)), _ptr(tri_lanes),
                                                C.c_uint32(len(tri_lanes)), _ptr(bounds)))

    def set_spheres(self, sphere_lanes: np.ndarray, count: int):
        lanes = np.ascontiguousarray(sphere_lanes, np.float32)
        assert lanes.size == ((count + 7) // 8) * 32
        _check(self._lib.dodrt_scene_set_spheres(self._h, _ptr(lanes), C.c_uint32(count)))

    def set_planes(self, plane_lanes: np.ndarray, count: int):
        lanes = np.ascontiguousarray(plane_lanes, np.float32)
        assert lanes.size == ((count + 7) // 8) * 48
        _check(self._lib.dodrt_scene_set_planes(self._h, _ptr(lanes), C.c_uint32(count)))

    def set_boxes(self, box_lanes: np.ndarray, count: int):
        lanes = np.ascontiguousarray(box_lanes, np.float32)
        assert lanes.size == ((count + 7) // 8) * 48
        _check(self._lib.dodrt_scene_set_boxes(self._h, _ptr(lanes), C.c_uint32(count)))

    def set_cylinders(self, cylinders: np.ndarray):
        cyl = np.ascontiguousarray(cylinders, CYL_DT)
        _check(self._lib.dodrt_scene_set_cylinders(self._h, _ptr(cyl), C.c_uint32(len(cyl))))

    def set_shading(self, tri_attributes: np.ndarray, mesh_colors: np.ndarray, sphere_colors: np.ndarray,
                    plane_colors: np.ndarray):
        """tri_attributes: the reference's Triangle::Attributes lanes (320 B each) as uint32 [lanes, 80]."""
        attrs = np.ascontiguousarray(tri_attributes, np.uint32).reshape(-1, 80)
        mc = np.ascontiguousarray(mesh_colors, np.float32).reshape(-1, 3)
        sc = np.ascontiguousarray(sphere_colors, np.float32).reshape(-1, 3)
        pc = np.ascontiguousarray(plane_colors, np.float32).reshape(-1, 3)
        _check(self._lib.dodrt_scene_set_shading(self._h, _ptr(attrs), C.c_uint32(len(attrs)), _ptr(mc),
                                                 C.c_uint32(len(mc)), _ptr(sc), _ptr(pc)))

    def render(self, frame: Frame, xs, ys, lights, depth: int = 10, out: Optional[np.ndarray] = None) -> np.ndarray:
        """dodrt_render: the reference's rayTrace for the whole frame; lights = [[x, y, z, intensity], ...]."""
        xs, ys = np.ascontiguousarray(xs, np.float32), np.ascontiguousarray(ys, np.float32)
        lights = np.ascontiguousarray(lights, np.float32).reshape(-1, 4)
        if out is not None:
            rgb = out
        elif frame.compact:
            rgb = np.empty((frame_local_pixels(frame), 3), np.uint8)
        else:
            rgb = np.empty((frame.height, frame.width, 3), np.uint8)
        _check(self._lib.dodrt_render(self._h, C.byref(frame), _ptr(xs), _ptr(ys), _ptr(lights), C.c_uint32(len(lights)),
                                      C.c_uint32(depth), _ptr(rgb)))
        return rgb

    def set_kernel_variant(self, variant: int):
        _check(self._lib.dodrt_scene_set_kernel_variant(self._h, C.c_int(variant)))

    def set_epsilon(self, eps: float):
        _check(self._lib.dodrt_scene_set_epsilon(self._h, C.c_float(eps)))

    # ---- host-buffer queries --------------------------------------------------------------------
    def intersect(self, rays: np.ndarray, classes: int = CLS_ALL, out: Optional[np.ndarray] = None) -> np.ndarray:
        rays = np.ascontiguousarray(rays, RAY_DT)
        hits = out if out is not None else np.empty(len(rays), HIT_DT)
        _check(self._lib.dodrt_intersect(self._h, _ptr(rays), C.c_uint64(len(rays)), C.c_uint32(classes), _ptr(hits)))
        return hits

    def _slots(self, frame: Frame) -> int:
        return frame_local_pixels(frame) if frame.compact else frame.width * frame.height

    def trace_primary(self, frame: Frame, xs: np.ndarray, ys: np.ndarray, out: Optional[np.ndarray] = None):
        xs, ys = np.ascontiguousarray(xs, np.float32), np.ascontiguousarray(ys, np.float32)
        assert len(xs) == frame.width and len(ys) == frame.height
        hits = out if out is not None else np.empty(self._slots(frame), HIT_DT)
        _check(self._lib.dodrt_trace_primary(self._h, C.byref(frame), _ptr(xs), _ptr(ys), _ptr(hits)))
        return hits

    def trace_shadow(self, frame: Frame, xs, ys, hits: np.ndarray, light, out: Optional[np.ndarray] = None):
        xs, ys = np.ascontiguousarray(xs, np.float32), np.ascontiguousarray(ys, np.float32)
        hits = np.ascontiguousarray(hits, HIT_DT)
        light = np.ascontiguousarray(light, np.float32)
        assert len(hits) == self._slots(frame) and light.size == 3
        vis = out if out is not None else np.empty(len(hits), np.uint8)
        _check(self._lib.dodrt_trace_shadow(self._h, C.byref(frame), _ptr(xs), _ptr(ys), _ptr(hits), _ptr(light),
                                            _ptr(vis)))
        return vis

    def trace_frame(self, frame: Frame, xs, ys, lights, hits_out: Optional[np.ndarray] = None,
                    vis_out: Optional[np.ndarray] = None):
        xs, ys = np.ascontiguousarray(xs, np.float32), np.ascontiguousarray(ys, np.float32)
        lights = np.ascontiguousarray(lights, np.float32).reshape(-1, 3)
        n = self._slots(frame)
        hits = hits_out if hits_out is not None else np.empty(n, HIT_DT)
        vis = vis_out if vis_out is not None else np.empty((len(lights), n), np.uint8)
        _check(self._lib.dodrt_trace_frame(self._h, C.byref(frame), _ptr(xs), _ptr(ys), _ptr(lights),
                                           C.c_uint32(len(lights)), _ptr(hits), _ptr(vis)))
        return hits, vis

    # ---- device-resident queries (raw device pointers, e.g. torch.Tensor.data_ptr()) -----------------
    def intersect_device(self, d_rays: int, num_rays: int, classes: int, d_hits: int, stream: int = 0):
        _check(self._lib.dodrt_intersect_device(self._h, C.c_void_p(d_rays), C.c_uint64(num_rays), C.c_uint32(classes),
                                                C.c_void_p(d_hits), C.c_void_p(stream)))

    def trace_primary_device(self, frame: Frame, d_xs: int, d_ys: int, d_hits: int, stream: int = 0):
        _check(self._lib.dodrt_trace_primary_device(self._h, C.byref(frame), C.c_void_p(d_xs), C.c_void_p(d_ys),
                                                    C.c_void_p(d_hits), C.c_void_p(stream)))

    def trace_shadow_device(self, frame: Frame, d_xs: int, d_ys: int, d_hits: int, light, d_visible: int,
                            stream: int = 0):
        light = np.ascontiguousarray(light, np.float32)
        _check(self._lib.dodrt_trace_shadow_device(self._h, C.byref(frame), C.c_void_p(d_xs), C.c_void_p(d_ys),
                                                   C.c_void_p(d_hits), _ptr(light), C.c_void_p(d_visible),
                                                   C.c_void_p(stream)))

    def frame_assemble_device(self, frame: Frame, d_compact_hits: int, d_compact_vis: int, slots_per_rank: int,
                              d_hits_out: int, d_vis_out: int, stream: int = 0):
        _check(self._lib.dodrt_frame_assemble_device(self._h, C.byref(frame), C.c_void_p(d_compact_hits),
                                                     C.c_void_p(d_compact_vis) if d_compact_vis else None,
                                                     C.c_uint64(slots_per_rank), C.c_void_p(d_hits_out),
                                                     C.c_void_p(d_vis_out) if d_vis_out else None, C.c_void_p(stream)))

    def trace_frame_device(self, frame: Frame, d_xs: int, d_ys: int, lights, d_hits: int, d_visible: int,
                           mirror: "Optional[FrameBuffer]" = None, stream: int = 0):
        """dodrt_trace_frame_device: primary + shadow queues of this call's tiles in one launch; `mirror` = a frame
        buffer view for this scene's GPU that also receives every result (row-major)."""
        lights = np.ascontiguousarray(lights, np.float32).reshape(-1, 3)
        _check(self._lib.dodrt_trace_frame_device(self._h, C.byref(frame), C.c_void_p(d_xs), C.c_void_p(d_ys), _ptr(lights),
                                                  C.c_uint32(len(lights)), C.c_void_p(d_hits),
                                                  C.c_void_p(d_visible) if d_visible else None,
                                                  mirror._h if mirror is not None else None, C.c_void_p(stream)))

    def debug_stats(self, enable: bool) -> np.ndarray:
        """dodrt_scene_debug_stats: counters since the last call (see include/dodrt.h), then on / off."""
        out = np.zeros(8, np.uint64)
        _check(self._lib.dodrt_scene_debug_stats(self._h, C.c_int(1 if enable else 0), _ptr(out)))
        return out

    def launch_count(self) -> int:
        n = C.c_uint64(0)
        _check(self._lib.dodrt_scene_launch_count(self._h, C.byref(n)))
        return n.value


class FrameBufferDesc(C.Structure):
    """``dodrt_frame_buffer_desc``: what another process needs to open a frame buffer (96 bytes, plain data)."""

    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("num_lights", C.c_uint32), ("device", C.c_uint32),
                ("bytes", C.c_uint64), ("ipc_handle", C.c_uint8 * 64), ("reserved", C.c_uint8 * 8)]

    def to_bytes(self) -> bytes:
        return bytes(self)

    @classmethod
    def from_bytes(cls, raw: bytes) -> "FrameBufferDesc":
        return cls.from_buffer_copy(raw)


class FrameBuffer:
    """``dodrt_frame_buffer``: a row-major frame in one GPU's HBM that kernels on other GPUs write into over NVLink.
    ``FrameBuffer.create(scene, ...)`` on the owner; ``attach`` (same process) / ``open`` (other process) elsewhere."""

    def __init__(self, handle, keepalive=None):
        self._lib = load()
        self._h = handle
        self._keep = keepalive

    @classmethod
    def create(cls, owner: Scene, width: int, height: int, num_lights: int) -> "FrameBuffer":
        h = C.c_void_p()
        _check(load().dodrt_frame_buffer_create(owner._h, C.c_uint32(width), C.c_uint32(height), C.c_uint32(num_lights),
                                                C.byref(h)))
        return cls(h, owner)

    @classmethod
    def attach(cls, user: Scene, owner_fb: "FrameBuffer") -> "FrameBuffer":
        h = C.c_void_p()
        _check(load().dodrt_frame_buffer_attach(user._h, owner_fb._h, C.byref(h)))
        return cls(h, (user, owner_fb))

    @classmethod
    def open(cls, user: Scene, desc: FrameBufferDesc) -> "FrameBuffer":
        h = C.c_void_p()
        _check(load().dodrt_frame_buffer_open(user._h, C.byref(desc), C.byref(h)))
        return cls(h, user)

    def export(self) -> FrameBufferDesc:
        d = FrameBufferDesc()
        _check(self._lib.dodrt_frame_buffer_export(self._h, C.byref(d)))
        return d

    def pointers(self):
        """(device address of the hit records, device address of the visibility bytes or 0)"""
        hits, vis = C.c_void_p(), C.c_void_p()
        _check(self._lib.dodrt_frame_buffer_pointers(self._h, C.byref(hits), C.byref(vis)))
        return hits.value or 0, vis.value or 0

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.dodrt_frame_buffer_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


class Multi:
    """``dodrt_multi``: one frame traced by several GPUs of this process, results in ONE host frame."""

    def __init__(self, scenes):
        self._lib = load()
        self._scenes = list(scenes)
        arr = (C.c_void_p * len(self._scenes))(*[s._h for s in self._scenes])
        self._h = C.c_void_p()
        _check(self._lib.dodrt_multi_create(arr, C.c_uint32(len(self._scenes)), C.byref(self._h)))

    def trace_frame(self, frame: Frame, xs, ys, lights, hits_out: Optional[np.ndarray] = None,
                    vis_out: Optional[np.ndarray] = None):
        xs, ys = np.ascontiguousarray(xs, np.float32), np.ascontiguousarray(ys, np.float32)
        lights = np.ascontiguousarray(lights, np.float32).reshape(-1, 3)
        n = frame.width * frame.height
        hits = hits_out if hits_out is not None else np.empty(n, HIT_DT)
        vis = vis_out if vis_out is not None else np.empty((len(lights), n), np.uint8)
        _check(self._lib.dodrt_multi_trace_frame(self._h, C.byref(frame), _ptr(xs), _ptr(ys), _ptr(lights),
                                                 C.c_uint32(len(lights)), _ptr(hits), _ptr(vis)))
        return hits, vis

    def render(self, frame: Frame, xs, ys, lights, depth: int = 10, out: Optional[np.ndarray] = None) -> np.ndarray:
        """dodrt_multi_render: the reference's rayTrace for the whole frame on all GPUs, one row-major rgb image."""
        xs, ys = np.ascontiguousarray(xs, np.float32), np.ascontiguousarray(ys, np.float32)
        lights = np.ascontiguousarray(lights, np.float32).reshape(-1, 4)
        rgb = out if out is not None else np.empty((frame.height, frame.width, 3), np.uint8)
        _check(self._lib.dodrt_multi_render(self._h, C.byref(frame), _ptr(xs), _ptr(ys), _ptr(lights), C.c_uint32(len(lights)),
                                            C.c_uint32(depth), _ptr(rgb)))
        return rgb

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.dodrt_multi_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
