#!/usr/bin/env python
"""Times every 32x32 tile of the dragon4k frame as its own launch (32 batches on 32 warps, so the time is
about the longest batch of the tile): where is the critical path of a frame?"""
import sys, os
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from dod_raytracer_b200 import capi, host, workloads
w = workloads.WORKLOADS["dragon4k"]
g = workloads.build_host_scene(w).upload(0)
dev = torch.device("cuda:0"); st = torch.cuda.current_stream()
xs, ys = host.ray_tables(w.width, w.height)
d_xs, d_ys = torch.from_numpy(xs).to(dev), torch.from_numpy(ys).to(dev)
d_hits = torch.empty((w.pixels, 16), dtype=torch.uint8, device=dev); d_vis = torch.zeros(w.pixels, dtype=torch.uint8, device=dev)
light = np.array(w.lights[0], np.float32)
full = capi.Frame.make(w.width, w.height, classes=w.classes)
g.trace_primary_device(full, d_xs.data_ptr(), d_ys.data_ptr(), d_hits.data_ptr(), st.cuda_stream)  # hits for the shadow pass
torch.cuda.synchronize()
tx, ty = 120, 68
ntiles = tx * ty
res = np.zeros((ntiles, 2), np.float32)
ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(ntiles)]
for k in range(ntiles):
    f = capi.Frame.make(w.width, w.height, classes=w.classes, first_tile=k, tile_stride=ntiles)
    ev[k][0].record(st)
    g.trace_primary_device(f, d_xs.data_ptr(), d_ys.data_ptr(), d_hits.data_ptr(), st.cuda_stream)
    ev[k][1].record(st)
    g.trace_shadow_device(f, d_xs.data_ptr(), d_ys.data_ptr(), d_hits.data_ptr(), light, d_vis.data_ptr(), st.cuda_stream)
    ev[k][2].record(st)
torch.cuda.synchronize()
for k in range(ntiles):
    res[k] = (ev[k][0].elapsed_time(ev[k][1]), ev[k][1].elapsed_time(ev[k][2]))
for name, col in (("primary", 0), ("shadow", 1)):
    t = res[:, col] * 1e3
    order = np.argsort(-t)
    print(f"{name}: per-tile us  median {np.median(t):.1f}  p90 {np.percentile(t, 90):.1f}  p99 {np.percentile(t, 99):.1f}  max {t.max():.1f}")
    print("   heaviest tiles (tx,ty,us):", [(int(k % tx), int(k // tx), round(float(t[k]), 1)) for k in order[:8]])
np.save(os.path.join(ROOT, "gpurun_out", "tile_cost_map.npy"), res)
