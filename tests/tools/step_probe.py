#!/usr/bin/env python
"""Per-step device times of back-to-back (or synchronised) primary passes of one workload: shows one-off allocation
spikes and launch gaps that an average hides.   python tests/tools/step_probe.py [workload] [sync]"""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from dod_raytracer_b200 import capi, host, workloads
w = workloads.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "dragon1080_primary"]
sync_each = len(sys.argv) > 2 and sys.argv[2] == "sync"
g = workloads.build_host_scene(w, keep_creation_order=True).upload(0)
dev = torch.device("cuda:0"); st = torch.cuda.current_stream()
xs, ys = host.ray_tables(w.width, w.height)
d_xs, d_ys = torch.from_numpy(xs).to(dev), torch.from_numpy(ys).to(dev)
d_hits = torch.empty((w.pixels, 16), dtype=torch.uint8, device=dev)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
frame = capi.Frame.make(w.width, w.height, classes=w.classes)
for _ in range(3):
    flush.zero_(); g.trace_primary_device(frame, d_xs.data_ptr(), d_ys.data_ptr(), d_hits.data_ptr(), st.cuda_stream)
torch.cuda.synchronize()
ev = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(12)]
for k in range(12):
    flush.zero_()
    ev[k][0].record(st); g.trace_primary_device(frame, d_xs.data_ptr(), d_ys.data_ptr(), d_hits.data_ptr(), st.cuda_stream); ev[k][1].record(st)
    if sync_each: torch.cuda.synchronize()
torch.cuda.synchronize()
print(("sync " if sync_each else "async") , " ".join(f"{e[0].elapsed_time(e[1]):.3f}" for e in ev))
