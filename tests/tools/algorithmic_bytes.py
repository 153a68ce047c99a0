#!/usr/bin/env python
"""Counts the work the REFERENCE traversal does per ray for every benchmark workload (kd nodes fetched,
triangle lanes tested) with the oracle's instrumented restatement, and writes
dod_raytracer_b200/algorithmic_bytes.json.  These are the per-unit figures behind `roofline.achieved`
(B_ray = 8*nodes + 288*lanes + io).  Run:  python tests/tools/algorithmic_bytes.py [workload ...]
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
from dod_raytracer_b200 import workloads  # noqa: E402
from oracle_api import MISS, Oracle, Scene, ensure_oracle_built  # noqa: E402

OUT = os.path.join(ROOT, "dod_raytracer_b200", "algorithmic_bytes.json")


def oracle_scene(arrays) -> Scene:
    s = Scene.__new__(Scene)
    s.nodes, s.tri_lanes, s.bounds = arrays["nodes"], arrays["tri_lanes"], arrays["bounds"]
    s.sphere_lanes, s.plane_lanes, s.box_lanes = arrays["sphere_lanes"], arrays["plane_lanes"], arrays["box_lanes"]
    s.spheres = np.zeros((arrays["num_spheres"], 4), np.float32)
    s.planes = np.zeros((arrays["num_planes"], 6), np.float32)
    s.boxes = np.zeros((arrays["num_boxes"], 6), np.float32)
    s.cylinders = arrays["cylinders"]
    s.epsilon = arrays["epsilon"]
    return s


def main():
    ensure_oracle_built()
    orc = Oracle()
    names = sys.argv[1:] or ["teapot1080", "dragon1080_primary", "dragon4k", "analytic1080"]
    result = json.load(open(OUT)) if os.path.exists(OUT) else {}
    threads = os.cpu_count() or 1
    for name in names:
        w = workloads.WORKLOADS[name]
        t0 = time.time()
        hs = workloads.build_host_scene(w)
        arrays = hs.arrays()
        scene = oracle_scene(arrays)
        z = hs.sizes()
        hits, c = orc.trace_primary(scene, w.width, w.height, w.classes, counters=True, nthreads=threads)
        entry = dict(description=w.description, mesh=workloads.mesh_label(w), width=w.width, height=w.height,
                     triangles=z.num_triangles, kd_nodes=z.num_nodes, tri_lanes=z.num_lanes, max_depth=z.max_depth,
                     primary_rays=w.pixels, primary_hits=int((hits["prim"] != MISS).sum()),
                     primary_nodes_per_ray=float(c["nodes"].mean()), primary_lanes_per_ray=float(c["lanes"].mean()),
                     primary_max_stack=int(c["max_stack"].max()))
        entry["primary_bytes_per_ray"] = 8 * entry["primary_nodes_per_ray"] + 288 * entry["primary_lanes_per_ray"] + 16
        if w.shadow:
            vis, c2 = orc.trace_shadow(scene, w.width, w.height, w.classes, hits, np.array(w.lights[0], np.float32),
                                       counters=True, nthreads=threads)
            n = entry["primary_hits"]
            entry.update(shadow_rays=n, shadow_visible=int(vis.sum()),
                         shadow_nodes_per_ray=float(c2["nodes"].sum() / max(n, 1)),
                         shadow_lanes_per_ray=float(c2["lanes"].sum() / max(n, 1)),
                         shadow_max_stack=int(c2["max_stack"].max()))
            entry["shadow_bytes_per_ray"] = 8 * entry["shadow_nodes_per_ray"] + 288 * entry["shadow_lanes_per_ray"] + 17
        entry["counted_in_s"] = round(time.time() - t0, 1)
        result[name] = entry
        print(name, json.dumps(entry, indent=1))
        json.dump(result, open(OUT, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
