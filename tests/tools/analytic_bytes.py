#!/usr/bin/env python
"""Algorithmic bytes per ray of the ACCELERATED sphere / box path (the reference brute-forces all lanes,
so its own traversal has no per-ray byte figure worth a roofline; the accelerated variant reports what its culling
structure fetches): counts culling-BVH nodes and primitives per ray on the GPU (dodrt_scene_debug_stats) for the
analytic workload and writes them into dod_raytracer_b200/algorithmic_bytes.json.  Needs a GPU.
    bytes per ray = 32 B per BVH node fetched + 16 B per sphere tested + 24 B per box tested + io"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from dod_raytracer_b200 import capi, host, workloads  # noqa: E402

OUT = os.path.join(ROOT, "dod_raytracer_b200", "algorithmic_bytes.json")
name = sys.argv[1] if len(sys.argv) > 1 else "analytic1080"
w = workloads.WORKLOADS[name]
g = workloads.build_host_scene(w).upload(0)
xs, ys = host.ray_tables(w.width, w.height)
frame = capi.Frame.make(w.width, w.height, classes=w.classes)
light = np.array(w.lights[0], np.float32)
g.debug_stats(True)
hits = g.trace_primary(frame, xs, ys)
p = g.debug_stats(True).astype(np.float64)
vis = g.trace_shadow(frame, xs, ys, hits, light)
s = g.debug_stats(False).astype(np.float64)
n_hit = int((hits["prim"] != capi.MISS).sum())
result = json.load(open(OUT))
e = result.setdefault(name, {})


def per_ray(c, rays, io):
    return dict(bvh_nodes=(c[0] + c[2]) / rays, spheres_tested=c[1] / rays, boxes_tested=c[3] / rays,
                bytes=(32.0 * (c[0] + c[2]) + 16.0 * c[1] + 24.0 * c[3]) / rays + io)


e["accelerated_primary"] = per_ray(p, w.pixels, 16.0)
e["accelerated_shadow"] = per_ray(s, max(n_hit, 1), 17.0)
e["primary_bytes_per_ray"], e["shadow_bytes_per_ray"] = e["accelerated_primary"]["bytes"], e["accelerated_shadow"]["bytes"]
e["primary_hits"], e["shadow_rays"], e["shadow_visible"] = n_hit, n_hit, int(vis.sum())
e["note"] = ("accelerated variant (culling BVH, dodrt_prim_bvh.cuh): 32 B per node fetched + 16 B per sphere + 24 B per box "
             "tested + io, counted on the GPU by tests/tools/analytic_bytes.py; the reference's brute force tests all "
             f"{(w.analytic + 7) // 8} sphere lanes per ray")
json.dump(result, open(OUT, "w"), indent=1, sort_keys=True)
print(json.dumps(e, indent=1))
