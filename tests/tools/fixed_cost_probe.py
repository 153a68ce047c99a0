import sys, os
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from dod_raytracer_b200 import capi, host, workloads
w = workloads.WORKLOADS["dragon4k"]
g = workloads.build_host_scene(w).upload(0)
dev = torch.device("cuda:0"); st = torch.cuda.current_stream()
light = np.array(w.lights[0], np.float32)
for (W, H) in [(32, 32), (256, 256), (1024, 1024), (3840, 2160)]:
    xs, ys = host.ray_tables(W, H)
    d_xs, d_ys = torch.from_numpy(xs).to(dev), torch.from_numpy(ys).to(dev)
    d_hits = torch.empty((W * H, 16), dtype=torch.uint8, device=dev); d_vis = torch.empty(W * H, dtype=torch.uint8, device=dev)
    for cls, name in ((w.classes, "all"), (8, "tree"), (7, "analytic")):
        f = capi.Frame.make(W, H, classes=cls)
        best = None
        for rep in range(5):
            e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            e[0].record(st); g.trace_primary_device(f, d_xs.data_ptr(), d_ys.data_ptr(), d_hits.data_ptr(), st.cuda_stream)
            e[1].record(st); g.trace_shadow_device(f, d_xs.data_ptr(), d_ys.data_ptr(), d_hits.data_ptr(), light, d_vis.data_ptr(), st.cuda_stream)
            e[2].record(st); torch.cuda.synchronize()
            t = (e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2]))
            best = t if best is None or sum(t) < sum(best) else best
        print(f"{W}x{H} {name:9s} primary {best[0]*1e3:8.1f} us  shadow {best[1]*1e3:8.1f} us")
