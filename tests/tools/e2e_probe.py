#!/usr/bin/env python
"""Host-buffer frame call (dodrt_trace_frame, pinned buffers) on rank 0's share of an N-way split, staged copies vs direct
stores by the kernel (DODRT_ZEROCOPY=0 / 1), plus the raw D2H rate.   python tests/tools/e2e_probe.py [workload] [N,N,..]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from dod_raytracer_b200 import capi, distributed, host, workloads  # noqa: E402

w = workloads.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "dragon4k"]
splits = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "1,2,4,8").split(",")]
g = workloads.build_host_scene(w, keep_creation_order=True).upload(0)
xs, ys = host.ray_tables(w.width, w.height)
lights = np.array(w.lights, np.float32)[: (1 if w.shadow else 0)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for n in splits:
    frame = distributed.rank_frame(w.width, w.height, w.classes, 0, n)
    slots = capi.frame_local_pixels(frame) if n > 1 else w.pixels
    h_hits = torch.empty((slots, 16), dtype=torch.uint8, pin_memory=True).numpy().reshape(-1).view(capi.HIT_DT)
    h_vis = torch.empty((max(len(lights), 1), slots), dtype=torch.uint8, pin_memory=True).numpy()
    out = {}
    for mode in ("0", "1"):
        os.environ["DODRT_ZEROCOPY"] = mode
        ts = []
        for k in range(9):
            flush.zero_()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            g.trace_frame(frame, xs, ys, lights, h_hits, h_vis)
            ts.append(time.perf_counter() - t0)
        out[mode] = float(np.median(ts[3:])) * 1e3
    print(f"{w.name} N={n} rank 0 ({slots * 17 / 1e6:.1f} MB): staged {out['0']:.3f} ms   zero-copy {out['1']:.3f} ms", flush=True)
d = torch.empty(w.pixels * 17, dtype=torch.uint8, device="cuda")
hbuf = torch.empty(w.pixels * 17, dtype=torch.uint8, pin_memory=True)
torch.cuda.synchronize()
t0 = time.perf_counter()
hbuf.copy_(d, non_blocking=True)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print("raw D2H %.2f ms = %.1f GB/s" % (dt * 1e3, d.numel() / dt / 1e9))
