#!/usr/bin/env python
"""The frame of one benchmark workload, separate primary / shadow passes, a few times -- the short command ncu wraps:
    python tests/tools/profile_workload.py <workload> [frames] [variant|-1] [N]      (N: rank 0's share of an N-way tile split)
ncu:  ... && ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s <2*(frames-1)> -c 2 -o gpurun_out/prof_<workload> \
          python tests/tools/profile_workload.py <workload> <frames>
(skips the warm-up frames' launches and captures the last frame's primary and shadow kernels)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from dod_raytracer_b200 import capi, host, workloads  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "dragon4k"
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 3
w = workloads.WORKLOADS[name]
g = workloads.build_host_scene(w, keep_creation_order=True).upload(0)
if len(sys.argv) > 3 and int(sys.argv[3]) >= 0:
    g.set_kernel_variant(int(sys.argv[3]))
split = int(sys.argv[4]) if len(sys.argv) > 4 else 1
dev = torch.device("cuda:0")
st = torch.cuda.current_stream()
xs, ys = host.ray_tables(w.width, w.height)
d_xs, d_ys = torch.from_numpy(xs).to(dev), torch.from_numpy(ys).to(dev)
d_hits = torch.empty((w.pixels, 16), dtype=torch.uint8, device=dev)
d_vis = torch.zeros(w.pixels, dtype=torch.uint8, device=dev)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
frame = capi.Frame.make(w.width, w.height, classes=w.classes, first_tile=0, tile_stride=split, compact=1 if split > 1 else 0)
light = np.array(w.lights[0], np.float32)
for k in range(frames):
    flush.zero_()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record(st)
    g.trace_primary_device(frame, d_xs.data_ptr(), d_ys.data_ptr(), d_hits.data_ptr(), st.cuda_stream)
    e[1].record(st)
    if w.shadow:
        g.trace_shadow_device(frame, d_xs.data_ptr(), d_ys.data_ptr(), d_hits.data_ptr(), light, d_vis.data_ptr(), st.cuda_stream)
    e[2].record(st)
    torch.cuda.synchronize()
    print(f"{name} frame {k}: primary {e[0].elapsed_time(e[1]):.3f} ms, shadow {e[1].elapsed_time(e[2]):.3f} ms", flush=True)
