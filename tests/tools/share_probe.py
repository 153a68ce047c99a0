#!/usr/bin/env python
"""Frame-share kernel time, one-launch frame kernel vs separate primary / shadow passes, for rank 0..k of an N-way
tile split emulated on ONE GPU (no peer writes).  usage: share_probe.py [workload] [N,N,...] [ranks-per-N]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from dod_raytracer_b200 import capi, distributed, host, workloads  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "dragon4k"
    splits = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "1,2,4,8").split(",")]
    nranks = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    w = workloads.WORKLOADS[name]
    g = workloads.build_host_scene(w, keep_creation_order=True).upload(0)
    dev = torch.device("cuda:0")
    xs, ys = host.ray_tables(w.width, w.height)
    d_xs, d_ys = torch.from_numpy(xs).to(dev), torch.from_numpy(ys).to(dev)
    lights = np.array(w.lights, np.float32)[: (1 if w.shadow else 0)]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream()
    tile = tuple(int(x) for x in os.environ.get("SHARE_TILE", "32x32").split("x"))  # tile shape of the split (A/B)
    for n in splits:
        spr = distributed.slots_per_rank(w.width, w.height, n, tile=tile) if n > 1 else w.pixels
        d_hits = torch.empty((spr, 16), dtype=torch.uint8, device=dev)
        d_vis = torch.empty((1, spr), dtype=torch.uint8, device=dev)
        for r in range(min(nranks, n)):
            f = distributed.rank_frame(w.width, w.height, w.classes, r, n, tile=tile)
            modes = os.environ.get("SHARE_MODES", "fused,queues,separate").split(",")
            out = {m: (float("nan"), float("nan")) for m in ("fused", "queues", "separate")}
            for mode in modes:
                os.environ["DODRT_FUSED"] = "0" if mode == "separate" else "1"
                os.environ["DODRT_FRAME_QUEUES"] = "1" if mode == "queues" else "0"
                ts = []
                for rep in range(8):
                    flush.zero_()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(st)
                    g.trace_frame_device(f, d_xs.data_ptr(), d_ys.data_ptr(), lights, d_hits.data_ptr(), d_vis.data_ptr(), None,
                                         st.cuda_stream)
                    e1.record(st)
                    torch.cuda.synchronize()
                    ts.append(e0.elapsed_time(e1))
                out[mode] = (float(np.median(ts[2:])), float(np.min(ts[2:])))
            print(f"{name} N={n} rank {r}: block-fused {out['fused'][0]:.3f} (min {out['fused'][1]:.3f}) ms   tile queues "
                  f"{out['queues'][0]:.3f} (min {out['queues'][1]:.3f}) ms   separate {out['separate'][0]:.3f} "
                  f"(min {out['separate'][1]:.3f}) ms", flush=True)


main()
