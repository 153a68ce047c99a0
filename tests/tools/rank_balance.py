#!/usr/bin/env python
"""Per-rank kernel times of an N-way tile split, emulated rank by rank on ONE GPU (no collectives)."""
import sys, os
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from dod_raytracer_b200 import capi, distributed, host, workloads

def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    tile = tuple(int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "32x32").split("x"))
    w = workloads.WORKLOADS["dragon4k"]
    g = workloads.build_host_scene(w).upload(0)
    dev = torch.device("cuda:0")
    xs, ys = host.ray_tables(w.width, w.height)
    d_xs, d_ys = torch.from_numpy(xs).to(dev), torch.from_numpy(ys).to(dev)
    spr = distributed.slots_per_rank(w.width, w.height, n, tile)
    d_hits = torch.empty((spr, 16), dtype=torch.uint8, device=dev)
    d_vis = torch.empty(spr, dtype=torch.uint8, device=dev)
    light = np.array(w.lights[0], np.float32)
    st = torch.cuda.current_stream()
    rows = []
    for r in range(n):
        f = distributed.rank_frame(w.width, w.height, w.classes, r, n, tile)
        best = None
        for rep in range(4):
            e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            e[0].record(st)
            g.trace_primary_device(f, d_xs.data_ptr(), d_ys.data_ptr(), d_hits.data_ptr(), st.cuda_stream)
            e[1].record(st)
            g.trace_shadow_device(f, d_xs.data_ptr(), d_ys.data_ptr(), d_hits.data_ptr(), light, d_vis.data_ptr(), st.cuda_stream)
            e[2].record(st)
            torch.cuda.synchronize()
            t = (e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2]))
            best = t if best is None or sum(t) < sum(best) else best
        rows.append(best)
        print(f"rank {r}: primary {best[0]:.3f} ms shadow {best[1]:.3f} ms")
    p, s = np.array(rows).T
    print(f"N={n} tile={tile}: sum {p.sum() + s.sum():.3f} ms, max rank {(p + s).max():.3f} ms, ideal {(p.sum() + s.sum()) / n:.3f} ms")

main()
