#!/usr/bin/env python
"""Where the as-is frame of dodrt_render goes: time by depth and number of lights (dragon stand-in scene, 1080p).
    python tests/tools/render_breakdown.py [teapot|dragon]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT]
from dod_raytracer_b200 import capi, host, workloads  # noqa: E402

mesh = sys.argv[1] if len(sys.argv) > 1 else "dragon"
wl = workloads.WORKLOADS["teapot1080" if mesh == "teapot" else "dragon4k"]
g = workloads.build_host_scene(wl).upload(0, shading=True)
w, h = 1920, 1080
xs, ys = host.ray_tables(w, h)
frame = capi.Frame.make(w, h, classes=workloads.CLS_REFERENCE)
out = torch.empty((h, w, 3), dtype=torch.uint8, pin_memory=True).numpy()
L = np.array(workloads.REFERENCE_LIGHTS, np.float32)


def ms(depth, nl):
    best = 1e9
    for _ in range(4):
        t0 = time.perf_counter()
        g.render(frame, xs, ys, L[:nl], depth, out)
        best = min(best, time.perf_counter() - t0)
    return best * 1e3


for depth, nl in ((1, 0), (1, 1), (1, 9), (2, 0), (2, 9), (3, 9), (5, 9), (10, 0), (10, 1), (10, 9)):
    print(f"{mesh} depth {depth:2d} lights {nl}: {ms(depth, nl):8.2f} ms", flush=True)
