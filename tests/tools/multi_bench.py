#!/usr/bin/env python
"""The one-process multi-GPU entry points of the C ABI, timed on 1, 2, 4, ... GPUs of this box:
  * dodrt_multi_trace_frame: the dragon4k frame (primary + light0 shadow) into ONE pinned host frame, wall clock of the
    synchronous call (host buffers in, host buffers out = the e2e form of the metric), checked against the 1-GPU result;
  * dodrt_multi_render: the reference's as-is frame (1920x1080, 9 lights, 10 bounces) into one rgb image.
    python tests/tools/multi_bench.py [workload] [max_gpus]"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from dod_raytracer_b200 import capi, host, workloads  # noqa: E402


def pinned(shape, dtype):
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    return torch.empty(n, dtype=torch.uint8, pin_memory=True).numpy().view(dtype).reshape(shape)


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "dragon4k"
    have = torch.cuda.device_count()
    most = min(int(sys.argv[2]), have) if len(sys.argv) > 2 else have
    w = workloads.WORKLOADS[name]
    hs = workloads.build_host_scene(w, keep_creation_order=True)
    scenes = [hs.upload(d, shading=True) for d in range(most)]
    xs, ys = host.ray_tables(w.width, w.height)
    lights = np.array(w.lights, np.float32)[:1]
    frame = capi.Frame.make(w.width, w.height, classes=w.classes)
    hits, vis = pinned((w.pixels,), capi.HIT_DT), pinned((1, w.pixels), np.uint8)
    rxs, rys = host.ray_tables(1920, 1080)
    rframe = capi.Frame.make(1920, 1080, classes=workloads.CLS_REFERENCE)
    rgb = pinned((1080, 1920, 3), np.uint8)
    ref = ref_rgb = None
    out = {}
    n = 1
    while n <= most:
        with capi.Multi(scenes[:n]) as m:
            ts = []
            for k in range(8):
                t0 = time.perf_counter()
                m.trace_frame(frame, xs, ys, lights, hits, vis)
                ts.append(time.perf_counter() - t0)
            shadow = int((hits["prim"] != capi.MISS).sum())
            t = float(np.median(ts[3:]))
            same = True
            if ref is None:
                ref = (hits.tobytes(), vis.tobytes())
            else:
                same = (hits.tobytes(), vis.tobytes()) == ref
            rs = []
            for k in range(5):
                t0 = time.perf_counter()
                m.render(rframe, rxs, rys, workloads.REFERENCE_LIGHTS, workloads.REFERENCE_DEPTH, rgb)
                rs.append(time.perf_counter() - t0)
            rsame = True
            if ref_rgb is None:
                ref_rgb = rgb.tobytes()
            else:
                rsame = rgb.tobytes() == ref_rgb
            out[n] = {"trace_frame_ms": round(t * 1e3, 3), "Mrays/s": round((w.pixels + shadow) / t / 1e6, 1),
                      "identical_to_1gpu": same, "render_ms": round(float(np.median(rs[2:])) * 1e3, 2),
                      "render_identical_to_1gpu": rsame}
            print(f"{name} {n} GPU(s): dodrt_multi_trace_frame {t * 1e3:.3f} ms = {(w.pixels + shadow) / t / 1e6:.0f} Mrays/s "
                  f"(identical {same}); dodrt_multi_render 1080p 9 lights 10 bounces {out[n]['render_ms']} ms (identical {rsame})",
                  flush=True)
        n *= 2
    print(json.dumps({name: out}))


main()
