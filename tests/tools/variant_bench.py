#!/usr/bin/env python
"""A/B timing of the traversal kernel variants on one workload (GPU box).  Device-resident buffers, CUDA
events, L2 flushed between launches; checks that every variant returns the same bytes as variant 0.
    python tests/tools/variant_bench.py [workload] [reps]"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from dod_raytracer_b200 import capi, host, workloads  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "dragon4k"
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    variants = [int(v) for v in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0, 1, 2]
    w = workloads.WORKLOADS[name]
    hs = workloads.build_host_scene(w)
    g = hs.upload(0)
    dev = torch.device("cuda:0")
    xs, ys = host.ray_tables(w.width, w.height)
    d_xs, d_ys = torch.from_numpy(xs).to(dev), torch.from_numpy(ys).to(dev)
    d_hits = torch.empty((w.pixels, 16), dtype=torch.uint8, device=dev)
    d_vis = torch.empty(w.pixels, dtype=torch.uint8, device=dev)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    frame = capi.Frame.make(w.width, w.height, classes=w.classes)
    light = np.array(w.lights[0], np.float32)
    st = torch.cuda.current_stream()
    ref = None
    out = {}
    for v in variants:
        g.set_kernel_variant(v)
        tp, ts = [], []
        for r in range(reps + 2):
            flush.zero_()
            e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            e[0].record(st)
            g.trace_primary_device(frame, d_xs.data_ptr(), d_ys.data_ptr(), d_hits.data_ptr(), st.cuda_stream)
            e[1].record(st)
            if w.shadow:
                g.trace_shadow_device(frame, d_xs.data_ptr(), d_ys.data_ptr(), d_hits.data_ptr(), light, d_vis.data_ptr(),
                                      st.cuda_stream)
            e[2].record(st)
            torch.cuda.synchronize()
            if r >= 2:
                tp.append(e[0].elapsed_time(e[1]))
                ts.append(e[1].elapsed_time(e[2]))
        res = (d_hits.cpu().numpy().tobytes(), d_vis.cpu().numpy().tobytes() if w.shadow else b"")
        if ref is None:
            ref = res
        same = res == ref
        nshadow = int((np.frombuffer(res[0], capi.HIT_DT)["prim"] != capi.MISS).sum()) if w.shadow else 0
        tot = min(tp) + (min(ts) if w.shadow else 0.0)
        out[v] = dict(primary_ms=min(tp), shadow_ms=min(ts) if w.shadow else None, frame_ms=tot,
                      mrays=(w.pixels + nshadow) / tot / 1e3, identical_to_first=same)
        print(f"variant {v}: primary {min(tp):.3f} ms  shadow {min(ts) if w.shadow else 0:.3f} ms  "
              f"-> {out[v]['mrays']:.0f} Mrays/s  identical={same}", flush=True)
    print(json.dumps({name: out}))


if __name__ == "__main__":
    main()
