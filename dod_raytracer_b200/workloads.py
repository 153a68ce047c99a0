"""Benchmark / parity workloads = BASELINE.json `configs`, built with the product's own host side.

Every workload is (scene recipe, frame size, shape classes, lights).  `dragon.obj` is absent from the
reference tree (.MISSING_LARGE_BLOBS), so the "dragon" workloads use the deterministic stand-in of
``host.standin_dragon`` unless ``DODRT_DRAGON_OBJ`` points at a real mesh; results always say which.

ALGORITHMIC_BYTES: per-ray traffic of the REFERENCE traversal (8 B per kd node fetched + 288 B per
triangle lane tested), counted by the oracle's instrumented restatement with
tests/tools/algorithmic_bytes.py and copied here (DESIGN.md "Algorithmic bytes").  They are scene
constants, not measurements of the GPU path.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field
from typing import Optional, Tuple

import numpy as np

from . import capi, host

LIGHT0 = (0.0, 0.0, -2.0)  # lights[0], main.cpp:284
# the reference's nine lights: position xyz, intensity (main.cpp:283-292)
REFERENCE_LIGHTS = ((0.0, 0.0, -2.0, 3.0), (4.0, 4.3, 3.3, 1.0), (-4.0, -2.95, 3.95, 1.0), (3.95, -4.2, 3.3, 1.0),
                    (-2.9, 4.2, 3.8, 1.0), (3.95, 2.8, -4.3, 1.0), (-3.0, -3.8, -3.3, 1.0), (4.2, -4.2, -3.4, 1.0),
                    (-2.9, 4.4, -3.5, 1.0))
REFERENCE_DEPTH = 10  # recursionDepth, main.cpp:301
CLS_REFERENCE = capi.CLS_SPHERE | capi.CLS_PLANE | capi.CLS_CYLINDER | capi.CLS_TREE


@dataclass
class Workload:
    name: str
    width: int
    height: int
    classes: int
    shadow: bool
    mesh: str  # "teapot" | "dragon" | "dragon16" | "none"
    reference_scene: bool = True  # 16 srand(1) spheres + 6 planes + cylinder (main.cpp:364-366)
    analytic: int = 0  # config 4: this many spheres and boxes
    dragon_n: int = 660
    description: str = ""
    lights: Tuple[Tuple[float, float, float], ...] = (LIGHT0,)
    # (nodes, lanes) per ray of the reference traversal: primary over all pixels, shadow over shadow rays
    primary_nodes_lanes: Optional[Tuple[float, float]] = None
    shadow_nodes_lanes: Optional[Tuple[float, float]] = None
    # config 4 (no kd-tree): bytes per ray of the accelerated sphere / box path incl. io (tests/tools/analytic_bytes.py)
    explicit_bytes: Optional[Tuple[float, float]] = None
    extra: dict = field(default_factory=dict)

    @property
    def pixels(self) -> int:
        return self.width * self.height

    def algorithmic_bytes(self, shadow_rays: int) -> Tuple[float, float]:
        """(primary kernel bytes, shadow kernel bytes) per launch over the whole frame."""
        if self.explicit_bytes:
            return self.pixels * self.explicit_bytes[0], shadow_rays * self.explicit_bytes[1]
        pn, pl = self.primary_nodes_lanes or (0.0, 0.0)
        sn, sl = self.shadow_nodes_lanes or (0.0, 0.0)
        primary = self.pixels * (8.0 * pn + 288.0 * pl + 16.0)  # + 16 B hit record out
        shadow = shadow_rays * (8.0 * sn + 288.0 * sl + 16.0 + 1.0)  # + hit record in, visibility out
        return primary, shadow


WORKLOADS = {
    # BASELINE.json configs[0]: the reference's own CPU-runnable case
    "teapot1080": Workload("teapot1080", 1920, 1080, CLS_REFERENCE, True, "teapot",
                           description="teapot.obj 1920x1080, reference scene, light0, primary+shadow"),
    # configs[1]
    "dragon1080_primary": Workload("dragon1080_primary", 1920, 1080, capi.CLS_TREE, False, "dragon",
                                   reference_scene=False,
                                   description="dragon stand-in 1920x1080 primary rays only, kd-tree"),
    # configs[2] -- the configuration BASELINE.json's metric is quoted on
    "dragon4k": Workload("dragon4k", 3840, 2160, CLS_REFERENCE, True, "dragon",
                         description="dragon stand-in (871,200 tris) 3840x2160, reference scene, light0, "
                                     "primary+shadow"),
    # configs[3]
    "analytic1080": Workload("analytic1080", 1920, 1080, capi.CLS_SPHERE | capi.CLS_BOX, True, "none",
                             reference_scene=False, analytic=10000,
                             description="10k spheres + 10k boxes, 1920x1080 primary+shadow, brute force"),
    # configs[4]
    "dragon16_8k": Workload("dragon16_8k", 7680, 4320, CLS_REFERENCE, True, "dragon16",
                            description="16 stand-in dragons on a 4x4 grid (13.9 M tris, one kd-tree) 7680x4320, "
                                        "reference scene, light0, primary+shadow"),
}


def _load_algorithmic_bytes():
    import json
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "algorithmic_bytes.json")
    if not os.path.exists(path):
        return
    for name, e in json.load(open(path)).items():
        w = WORKLOADS.get(name)
        if w is None:
            continue
        w.primary_nodes_lanes = (e["primary_nodes_per_ray"], e["primary_lanes_per_ray"])
        if "accelerated_primary" in e:
            w.explicit_bytes = (e["accelerated_primary"]["bytes"], e["accelerated_shadow"]["bytes"])
        if "shadow_nodes_per_ray" in e:
            w.shadow_nodes_lanes = (e["shadow_nodes_per_ray"], e["shadow_lanes_per_ray"])
        w.extra["counted"] = {k: e[k] for k in ("primary_hits", "shadow_rays", "shadow_visible", "triangles", "kd_nodes",
                                                 "tri_lanes", "max_depth") if k in e}


_load_algorithmic_bytes()

TEAPOT_FIXTURE = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden",
                              "teapot.dodm")
DRAGON_SCALE = 0.68  # keeps the camera (z=-4.9) and light0 (0,0,-2) outside the mesh


def mesh_label(w: Workload) -> str:
    if w.mesh in ("dragon", "dragon16") and os.environ.get("DODRT_DRAGON_OBJ"):
        return os.environ["DODRT_DRAGON_OBJ"]
    return {"dragon": f"stand-in displaced sphere n={w.dragon_n} ({2 * w.dragon_n ** 2} tris; dragon.obj is a "
                      "missing blob in the reference)",
            "dragon16": f"16 x stand-in n={w.dragon_n}", "teapot": "teapot (tests/golden/teapot.dodm)",
            "none": "none"}[w.mesh]


def write_mesh_files(w: Workload, directory: str):
    """The workload's meshes as files both the product loader and the reference's loader stand-in read
    (used by the CPU baseline, which has to build the same scene through the reference's own code).
    Returns [(path, scale, translate)] -- transforms are baked into the written positions."""
    out = []
    if w.mesh == "teapot":
        out.append(TEAPOT_FIXTURE)
    elif w.mesh in ("dragon", "dragon16"):
        real = os.environ.get("DODRT_DRAGON_OBJ")
        if real:
            out.append(real)
        else:
            pos, idx = host.standin_dragon(w.dragon_n)
            for k, (scale, tr) in enumerate(_dragon_instances(w)):
                p = (pos * np.float32(scale) + np.asarray(tr, np.float32)).astype(np.float32)
                path = os.path.join(directory, f"{w.name}_{k}.dodm")
                host.write_dodm(path, p, idx)
                out.append(path)
    return out


def _dragon_instances(w: Workload):
    if w.mesh == "dragon":
        return [(DRAGON_SCALE, (0.0, 0.0, 0.0))]
    # config 5: 4x4 grid, scale 0.4, pitch 2.2, in the z = 0 plane
    inst = []
    for gy in range(4):
        for gx in range(4):
            inst.append((0.4, ((gx - 1.5) * 2.2, (gy - 1.5) * 2.2, 0.0)))
    return inst


def build_host_scene(w: Workload, mesh_files=None, keep_creation_order: bool = False) -> host.HostScene:
    """Scene registration in the reference's order: spheres, planes, cylinder, meshes, buildTree (main.cpp:364-368).
    Meshes are added from `mesh_files` when given (so that the CPU baseline and the GPU path read the very same
    bytes), otherwise generated in memory with the same arithmetic."""
    hs = host.HostScene()
    if w.reference_scene:
        hs.add_reference_scene(1, 16)
    if w.analytic:
        hs.add_analytic_scene(4, w.analytic)
    if mesh_files is not None:
        for path in mesh_files:
            hs.add_mesh_file(path)
    elif w.mesh == "teapot":
        hs.add_mesh_file(TEAPOT_FIXTURE)
    elif w.mesh in ("dragon", "dragon16"):
        real = os.environ.get("DODRT_DRAGON_OBJ")
        if real:
            hs.add_mesh_file(real)
        else:
            pos, idx = host.standin_dragon(w.dragon_n)
            for scale, tr in _dragon_instances(w):
                p = (pos * np.float32(scale) + np.asarray(tr, np.float32)).astype(np.float32)
                hs.add_mesh(p, idx)
    if w.mesh != "none":
        # keep_creation_order: skip the host-side lane re-order, upload() lets the GPU do it (f-4)
        hs.build_tree(keep_creation_order=keep_creation_order)
    return hs
