"""Helper process of tests/test_gpu_frame.py::test_frame_buffer_shared_with_another_process: opens the frame buffer
another process exported and traces rank `rank` of a `world`-way tile split of the teapot frame into it."""
import sys

import numpy as np
import torch

from dod_raytracer_b200 import capi, host
from gpu_util import upload
from scenes import LIGHT0, teapot_scene

desc = capi.FrameBufferDesc.from_bytes(open(sys.argv[1], "rb").read())
rank, world = int(sys.argv[2]), int(sys.argv[3])
device = int(sys.argv[4]) if len(sys.argv) > 4 else 0
w, h = desc.width, desc.height
xs, ys = host.ray_tables(w, h)
dev = torch.device("cuda", device)
torch.cuda.set_device(dev)
with upload(teapot_scene(full=True), device) as g:
    view = capi.FrameBuffer.open(g, desc)
    f = capi.Frame.make(w, h, classes=15, first_tile=rank, tile_stride=world, compact=1)
    slots = capi.frame_local_pixels(f)
    d_xs, d_ys = torch.from_numpy(xs).to(dev), torch.from_numpy(ys).to(dev)
    d_hits = torch.empty((slots, 16), dtype=torch.uint8, device=dev)
    d_vis = torch.empty((1, slots), dtype=torch.uint8, device=dev)
    g.trace_frame_device(f, d_xs.data_ptr(), d_ys.data_ptr(), LIGHT0[None, :], d_hits.data_ptr(), d_vis.data_ptr(), view,
                         torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    view.close()
print("peer ok")
