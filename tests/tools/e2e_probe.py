import sys, time, os
sys.path.insert(0, '.')
import numpy as np, torch
from dod_raytracer_b200 import capi, host, workloads
w = workloads.WORKLOADS['dragon4k']
hs = workloads.build_host_scene(w); g = hs.upload(0)
xs, ys = host.ray_tables(w.width, w.height)
frame = capi.Frame.make(w.width, w.height, classes=w.classes)
h_hits = torch.empty((w.pixels, 16), dtype=torch.uint8, pin_memory=True).numpy().reshape(-1).view(capi.HIT_DT)
h_vis = torch.empty((1, w.pixels), dtype=torch.uint8, pin_memory=True).numpy()
lights = np.array(w.lights, np.float32)
for _ in range(3): g.trace_frame(frame, xs, ys, lights, h_hits, h_vis)
ts=[]
for _ in range(8):
    t0=time.perf_counter(); g.trace_frame(frame, xs, ys, lights, h_hits, h_vis); ts.append(time.perf_counter()-t0)
print('bands', os.environ.get('DODRT_BANDS'), 'e2e ms min %.3f med %.3f'%(min(ts)*1e3, sorted(ts)[len(ts)//2]*1e3))
# raw D2H speed
d = torch.empty(w.pixels*17, dtype=torch.uint8, device='cuda'); hbuf = torch.empty(w.pixels*17, dtype=torch.uint8, pin_memory=True)
torch.cuda.synchronize(); t0=time.perf_counter(); hbuf.copy_(d, non_blocking=True); torch.cuda.synchronize(); dt=time.perf_counter()-t0
print('raw D2H %.2f ms = %.1f GB/s'%(dt*1e3, d.numel()/dt/1e9))
