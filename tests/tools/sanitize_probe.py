#!/usr/bin/env python
"""Small end-to-end run meant for compute-sanitizer (memcheck / initcheck): teapot scene, a 320x180 frame, a compact
1-of-3 tile split and a 3-bounce render through the plain kernel, the donating kernel, the donating kernel with
every ray forced through the queue, the single-copy exact stage and the ray pool; all five must return the same bytes.
    compute-sanitizer --tool memcheck python tests/tools/sanitize_probe.py
(compute-sanitizer is closed on this round's GPU pool, so only the plain run -- identical bytes -- has been done.)"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
from dod_raytracer_b200 import capi, host, workloads  # noqa: E402


def main():
    w, h = 320, 180
    xs, ys = host.ray_tables(w, h)
    light = np.array([workloads.LIGHT0], np.float32)
    ref = None
    for variant, always in ((3, "0"), (7, "0"), (7, "1"), (8, "0"), (4, "0")):
        os.environ["DODRT_DONATE_ALWAYS"] = always
        hs = host.HostScene()
        hs.add_reference_scene(1, 16)
        hs.add_mesh_file(workloads.TEAPOT_FIXTURE)
        hs.build_tree(keep_creation_order=True)
        with hs.upload(0, shading=True) as g:
            g.set_kernel_variant(variant)
            frame = capi.Frame.make(w, h, classes=workloads.CLS_REFERENCE)
            hits, vis = g.trace_frame(frame, xs, ys, light)
            split = capi.Frame.make(w, h, classes=workloads.CLS_REFERENCE, first_tile=1, tile_stride=3, compact=1)
            g.trace_frame(split, xs, ys, light)
            rgb = g.render(frame, xs, ys, workloads.REFERENCE_LIGHTS, 3)
            out = (hits.tobytes(), vis.tobytes(), rgb.tobytes())
            ref = ref or out
            assert out == ref, (variant, always)
        print("variant", variant, "always", always, "ok", flush=True)


main()
